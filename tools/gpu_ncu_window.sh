#!/bin/bash
# ncu --set full capture of selected kernels inside the profiler window of the bench (one replayed posterior solve).
# Usage: bash tools/gpu_ncu_window.sh <tag> <kernel regex> <skip> <count>
TAG=${1:-w}; KRE=${2:-k_gemm2}; SKIP=${3:-0}; COUNT=${4:-30}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --inflight 1"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$KRE -s $SKIP -c $COUNT \
    -f -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu_full.log | cut -c1-200
