#!/usr/bin/env bash
# Run on the GPU box (under gpurun): bench line + ncu launch list + two `ncu --set full` captures
# (cache flushed = ncu default, and --cache-control none = live L2 behaviour) of one step's kernels.
#   tools/gpu_profile.sh <tag> [extra bench args]
set -uo pipefail
TAG=${1:-run}; shift || true
OUT=gpurun_out
mkdir -p $OUT
KREG='regex:gemm_img|gemm_chain|token_fused|sig_attention|match_tc|match_tail|desc_tiles'
python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 $OUT/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:ltr' -s 22 -c 44 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --profile-only --steps 1 --warmup 1 "$@" > /dev/null 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "$KREG" -s 21 -c 21 -f -o $OUT/${TAG}_full_flush \
    python bench.py --profile-only --steps 1 --warmup 1 "$@" > $OUT/${TAG}_ncu1.log 2>&1; echo "ncu flush rc=$?"
ncu --set full --clock-control none --cache-control none -k "$KREG" -s 21 -c 21 -f -o $OUT/${TAG}_full_live \
    python bench.py --profile-only --steps 1 --warmup 1 "$@" > $OUT/${TAG}_ncu2.log 2>&1; echo "ncu live rc=$?"
ls -la $OUT | tail -8
