#!/bin/bash
# Throughput against the number of independent posterior problems in flight per GPU.
OUT=gpurun_out; mkdir -p $OUT
for B in ${@:-1 2 3 4 5 6}; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --inflight $B > $OUT/inflight_$B.json 2> $OUT/inflight_$B.err; echo "B=$B rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/inflight_$B.json"))
print("B=$B value",round(d["value"],2),"ms/step",round(d["ms_per_step"],2),"single",round(d["config"]["single_solve_latency_ms"],2),"e2e",round(d["e2e"]["value"],2),"ok",d["parity_check"]["ok"])
PY
done
