# Run under `gpurun --gpus 2`: the bench with both count-exchange paths + the reference arm under torchrun
for g in p2p nccl; do LTR_BENCH_GATHER=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu > gpurun_out/r2g_bench_n2_$g.json 2> gpurun_out/r2g_bench_n2_$g.err; python -c "
import json
d=json.loads(open('gpurun_out/r2g_bench_n2_$g.json').read().strip().splitlines()[-1]); print('$g', round(d['value']), d['ms_per_step'], d['config'].get('count_gather'), d['e2e']['value'])" || grep -E "Error|error" gpurun_out/r2g_bench_n2_$g.err | tail -8; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
