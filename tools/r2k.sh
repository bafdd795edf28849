timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
LTR_CHAIN_MIN_TILES=1 timeout 600 python -m pytest tests -m gpu -q -x -k "forward or varlen or full_size or cfg3 or pair or plumbing or shipped" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu > gpurun_out/r2r_bench_default.json 2> gpurun_out/r2r_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2r_bench_default.json").read().strip().splitlines()[-1]); print("default", round(d["value"]), d["ms_per_step"], d["gpu_launches"], {k:(round(v["avg_launch_ms"]*1e3,1), v["launches_per_step"]) for k,v in d["roofline_by_class"].items()})
except Exception as e:
    print("default failed", e); print(open("gpurun_out/r2r_bench.err").read()[-600:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm_|token_fused|sig_attention|match_|desc_tiles|small_mlp|final_norm|argmin|mutual|segmean|dist_kernel|gather_wait' -s 22 -c 44 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --profile-only --steps 1 --warmup 1 > /dev/null 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r2r_launches.csv
