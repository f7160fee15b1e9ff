#!/bin/bash
# Sanity record after the last edits: the Lagrange-triangle GPU tests (incl. the P2 dataset loop and the error paths)
# and the default bench line at the driver's step counts.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_fem2d.py -q -m gpu > $OUT/r02i_pytest_fem2d.txt 2>&1; echo "pytest rc=$?"; tail -6 $OUT/r02i_pytest_fem2d.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r02i_bench.json 2> $OUT/r02i_bench.err; echo "bench rc=$?"; tail -c 600 $OUT/r02i_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02i_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "detail", {k: d["detail"][k] for k in ("analyze_s", "analyze_with_given_perm_s", "single_solve_latency_ms")})
PY
