for c in 0 3 7; do LTR_TRACE_CHAIN=$c timeout 300 python tools/chain_trace.py 2>&1 | grep -v "^sig_attention" | tail -14; done
timeout 300 python tools/token_trace.py 2>&1 | tail -26
