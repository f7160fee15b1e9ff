#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus G): bench.py and the time-sharded block-tridiagonal factor/solve over NCCL
# at every N in the list (N <= G).  Usage: bash tools/gpu_multi.sh "1 2 4" tag ["2048x128 4096x64"]
NS=${1:-"1 2"}; TAG=${2:-mg}; CFGS=${3:-"2048x128 4096x64"}
OUT=gpurun_out; mkdir -p $OUT
for N in $NS; do
  if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
  [ -n "${SKIP_BENCH:-}" ] || timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
  [ -n "${SKIP_BENCH:-}" ] || python - <<PY
import json
for l in open("$OUT/${TAG}_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N value",round(d["value"],2),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"],2),"ok",d["parity_check"]["ok"])
PY
  for cfg in $CFGS; do
    set -- ${cfg/x/ }
    SEQ=""; [ "$N" = "1" ] && SEQ="--seq"
    timeout 600 $TR tools/bench_btd_dist.py --b $1 --N $2 $SEQ > $OUT/${TAG}_btd_dist_$1x$2_n$N.json 2> $OUT/${TAG}_btd_dist_$1x$2_n$N.err; echo "btd b=$1 N=$2 ranks=$N rc=$?"; tail -1 $OUT/${TAG}_btd_dist_$1x$2_n$N.json | cut -c1-400
  done
done
