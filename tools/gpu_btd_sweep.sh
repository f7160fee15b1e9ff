#!/bin/bash
# Dense block-tridiagonal sweep of SURVEY.md section 8(d) config 5 (b x N blocks) on N_GPU ranks of one box.
# Usage (under gpurun --gpus G): bash tools/gpu_btd_sweep.sh G "1024x1024 4096x1024 8192x256" [--seq]
G=${1:-2}; CFGS=${2:-"1024x1024"}; SEQ=${3:-}
OUT=gpurun_out; mkdir -p $OUT
export GMRFB_POOL_MAX_GB=150
for cfg in $CFGS; do
  set -- ${cfg/x/ }
  if [ "$G" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517"; fi
  timeout 600 $TR tools/bench_btd_dist.py --b $1 --N $2 $SEQ > $OUT/r02_btd_sweep_$1x$2_n$G.json 2> $OUT/r02_btd_sweep_$1x$2_n$G.err
  echo "b=$1 N=$2 ranks=$G rc=$?"; tail -1 $OUT/r02_btd_sweep_$1x$2_n$G.json | cut -c1-700; tail -2 $OUT/r02_btd_sweep_$1x$2_n$G.err | cut -c1-300
done
