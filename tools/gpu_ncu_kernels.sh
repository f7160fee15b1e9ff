#!/bin/bash
# ncu --set full captures of individual latency-critical kernels inside the bench command (tuning evidence).
TAG=${1:-k}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/${TAG}_$1 $CMD > $OUT/${TAG}_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap potrf k_potrf64 150 2
cap apply k_apply_inv 150 2
cap fwd k_fwd_step 150 2
cap bwd k_bwd_step 10 2
cap fsmall k_front_factor_small 16 2
cap ssmall k_front_selinv_small 0 2
ls -la $OUT | grep ${TAG}_
