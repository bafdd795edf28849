LTR_CHAIN_MIN_TILES=1 timeout 600 python -m pytest tests -m gpu -q -x -k "forward or varlen or full_size or cfg3 or pair or plumbing or shipped" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/chain_trace.py 2>&1 | tail -9 | head -7
for v in "default" "LTR_GEMM_PAIR=0"; do
  if [ "$v" = default ]; then envs=""; else envs="$v"; fi
  env $envs timeout 300 python bench.py --no-cpu > gpurun_out/r2i_bench_$v.json 2> gpurun_out/r2i_bench.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2i_bench_$v.json").read().strip().splitlines()[-1]); print("$v", round(d["value"]), d["ms_per_step"], d["gpu_launches"], {k:(round(v["avg_launch_ms"]*1e3,1), v["launches_per_step"]) for k,v in d["roofline_by_class"].items()})
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/r2i_bench.err").read()[-600:])
PY
done
for g in p2p nccl; do LTR_BENCH_GATHER=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu > gpurun_out/r2i_bench_n2_$g.json 2> gpurun_out/r2i_bench_n2_$g.err; python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench_n2_$g.json').read().strip().splitlines()[-1]); print('$g', round(d['value']), d['ms_per_step'], d['config'].get('count_gather'))" || grep -E "Error|error" gpurun_out/r2i_bench_n2_$g.err | tail -8; done
