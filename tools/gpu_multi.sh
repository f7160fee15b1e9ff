#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus N): bench.py at N ranks, time-sharded block-tridiagonal factor/solve over NCCL.
N=${1:-2}; TAG=${2:-mg}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time $TR bench.py --gpus $N --steps 3 --warmup 3 ) > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
for l in open("$OUT/${TAG}_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"ok",d["parity_check"]["ok"])
PY
$TR tools/bench_btd_dist.py --b 2048 --N 64 --seq > $OUT/${TAG}_btd_dist_2048_n$N.json 2> $OUT/${TAG}_btd_dist_2048_n$N.err; echo "btd 2048 rc=$?"; tail -1 $OUT/${TAG}_btd_dist_2048_n$N.json
$TR tools/bench_btd_dist.py --b 4096 --N 32 --seq > $OUT/${TAG}_btd_dist_4096_n$N.json 2> $OUT/${TAG}_btd_dist_4096_n$N.err; echo "btd 4096 rc=$?"; tail -1 $OUT/${TAG}_btd_dist_4096_n$N.json
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench N=1 rc=$?"
python tools/bench_btd_dist.py --b 2048 --N 64 > $OUT/${TAG}_btd_dist_2048_n1.json 2>&1; tail -1 $OUT/${TAG}_btd_dist_2048_n1.json
python tools/bench_btd_dist.py --b 4096 --N 32 > $OUT/${TAG}_btd_dist_4096_n1.json 2>&1; tail -1 $OUT/${TAG}_btd_dist_4096_n1.json
tail -3 $OUT/${TAG}_bench_n$N.err
