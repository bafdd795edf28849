#!/usr/bin/env bash
# One GPU-box session: matcher tests first (new kernels), full GPU suite, bench lines of the workloads.
#   tools/gpu_round.sh <tag> [workloads...]
set -uo pipefail
TAG=${1:-run}; shift || true
WLS=${@:-cfg1 cfg4}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
timeout 600 python -m pytest tests -m gpu -q -x -k "tc_matcher or nn_matcher or cfg4 or distmat or encoder_tiles" > $OUT/${TAG}_pytest_match.log 2>&1; echo "pytest matcher rc=$?"; tail -5 $OUT/${TAG}_pytest_match.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest all rc=$?"; tail -8 $OUT/${TAG}_pytest.log
LTR_CHAIN_MIN_TILES=1 timeout 900 python -m pytest tests -m gpu -q -k "forward or varlen or full_size or cfg3 or pair or plumbing or shipped" > $OUT/${TAG}_pytest_chain1.log 2>&1; echo "pytest (chain forced on small batches) rc=$?"; tail -3 $OUT/${TAG}_pytest_chain1.log
for wl in $WLS; do
  timeout 600 python bench.py --workload $wl > $OUT/${TAG}_bench_$wl.json 2> $OUT/${TAG}_bench_$wl.err; echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", round(d["value"]), "pairs/s", round(d["ms_per_step"],4), "ms/step e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "shares", d["kernel_time_shares"], "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "check", d["output_check"])
except Exception as e:
    print("$wl: no bench line", e); print(open("$OUT/${TAG}_bench_$wl.err").read()[-1500:])
PY
done
