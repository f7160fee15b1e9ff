#!/bin/bash
# A/B of environment-selected tuning switches on the bench step (one problem in flight). Usage: gpu_ab.sh "VAR=val" ...
OUT=gpurun_out; mkdir -p $OUT
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg GMRFB_PROFILE_DUMP=$OUT/ab_${i}_dump.csv timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --inflight 1 > $OUT/ab_${i}.json 2> $OUT/ab_${i}.err
  python - <<PY
import json
d=json.load(open("$OUT/ab_${i}.json"))
kp={r["name"]:r for r in d["kernel_profile"]}
print("[$cfg] ms_per_step %.2f"%d["ms_per_step"], " ".join("%s %.2fms %.1fTF"%(k.split()[0],kp[k]["ms"],kp[k].get("tflops",0)) for k in kp if "gemm" in k), "ok", d["parity_check"]["ok"])
PY
done
