#!/bin/bash
# Sweep of the GEMM tile-configuration threshold (GMRFB_GEMM_SMALL_MAX) on the bench step, one problem in flight.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/sw_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $OUT/sw_pytest.txt
for th in 0 296 592 1184 2368 1000000; do
  GMRFB_GEMM_SMALL_MAX=$th GMRFB_PROFILE_DUMP=$OUT/sw_${th}_dump.csv timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --inflight 1 > $OUT/sw_${th}.json 2> $OUT/sw_${th}.err
  python - <<PY
import json
d=json.load(open("$OUT/sw_${th}.json"))
kp={r["name"]:r for r in d["kernel_profile"]}
print("th=$th ms_per_step %.2f"%d["ms_per_step"], " ".join("%s %.2fms %.1fTF"%(k.split()[0],kp[k]["ms"],kp[k].get("tflops",0)) for k in kp if "gemm" in k))
PY
done
