# Run on the GPU box (gpurun): full GPU test suite, the chain-forced subset, one bench line; TRACE_CHAINS="0 3 7" / TOKEN_TRACE=1 add clock64 timelines
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
LTR_CHAIN_MIN_TILES=1 timeout 600 python -m pytest tests -m gpu -q -x -k "forward or varlen or full_size or cfg3 or pair or plumbing or shipped or ffn_width" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu > gpurun_out/check_bench_default.json 2> gpurun_out/check_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/check_bench_default.json").read().strip().splitlines()[-1]); print("default", round(d["value"]), d["ms_per_step"], d["gpu_launches"], {k:(round(v["avg_launch_ms"]*1e3,1), v["launches_per_step"]) for k,v in d["roofline_by_class"].items()}, d.get("output_check"))
except Exception as e:
    print("default failed", e); print(open("gpurun_out/check_bench.err").read()[-600:])
PY
for c in ${TRACE_CHAINS:-}; do LTR_TRACE_CHAIN=$c timeout 300 python tools/chain_trace.py 2>&1 | grep -v "^sig_attention" | tail -14; done
if [ -n "${TOKEN_TRACE:-}" ]; then timeout 300 python tools/token_trace.py 2>&1 | tail -26; fi
