#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of one timed bench step.
TAG=${1:-ll}; SKIP=${2:-10170}; COUNT=${3:-3400}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -2 $OUT/${TAG}_plain.log | cut -c1-300; wc -l $OUT/${TAG}_launches.csv
