for v in "LTR_BENCH_GATHER=off" "LTR_BENCH_GATHER=p2p" "LTR_BENCH_GATHER=p2p LTR_BENCH_KEEP=1"; do env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu --steps 40 > gpurun_out/r2h_n2.json 2> gpurun_out/r2h_n2.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h_n2.json').read().strip().splitlines()[-1]); print('$v', round(d['value']), d['ms_per_step'], d['config'].get('count_gather'))" || grep -E "Error|error" gpurun_out/r2h_n2.err | tail -8; done
timeout 300 python bench.py --no-cpu --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('single', round(d['value']), d['ms_per_step'])"
