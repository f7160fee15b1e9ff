#!/bin/bash
# A/B of the end-to-end leg's worker stagger (bench.py --e2e-stagger-ms) + re-run of the Lagrange-triangle tests.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_fem2d.py -q -m gpu > $OUT/fem2d_pytest2.txt 2>&1; echo "pytest rc=$?"; tail -4 $OUT/fem2d_pytest2.txt
for S in 0 6 12 25; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-extras --e2e-stagger-ms $S > $OUT/stagger_$S.json 2> $OUT/stagger_$S.err
  echo "stagger $S rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/stagger_$S.json"))
print("stagger", $S, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms_e2e", round(d["e2e"]["ms_per_step"],1))
PY
done
