# Run on the GPU box (gpurun): everything the round-end record needs - tests, bench of all workloads, ncu summaries, launch list, latency, traces
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
LTR_CHAIN_MIN_TILES=1 timeout 600 python -m pytest tests -m gpu -q -x -k "forward or varlen or full_size or cfg3 or pair or plumbing or shipped or ffn_width" 2>&1 | tail -2
bash tools/gpu_profile.sh r2g 2>&1 | grep -E "rc=" 
for w in cfg2 cfg3 cfg4; do timeout 300 python bench.py --no-cpu --workload $w > gpurun_out/r2g_bench_$w.json 2> gpurun_out/r2g_bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/r2g_bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['value']), d['ms_per_step'], d.get('output_check'))" || tail -5 gpurun_out/r2g_bench_$w.err; done
timeout 300 python tools/latency_b1.py > gpurun_out/r2g_latency_b1.json 2> gpurun_out/r2g_latency.err; tail -c 700 gpurun_out/r2g_latency_b1.json
for c in 0 3 7; do LTR_TRACE_CHAIN=$c timeout 300 python tools/chain_trace.py 2>&1 | grep -v "^sig_attention" | tail -16; done > gpurun_out/r2g_chain_trace.txt 2>&1
timeout 300 python tools/token_trace.py > gpurun_out/r2g_token_trace.txt 2>&1
rm -f gpurun_out/r2g_src4.ncu-rep
du -sm gpurun_out
