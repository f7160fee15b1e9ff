#!/bin/bash
# First GPU job of the next round: time the orderings of profiles/r01_ordering_study.md on the bench step
# (one problem in flight for the per-kernel profile, four for the throughput), then refresh the evidence.
# Usage (repo root on the GPU box): bash tools/gpu_orderings.sh <tag>
TAG=${1:-ord}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.txt
run() {  # name, env, extra bench args
  for B in 1 4; do
    env $2 GMRFB_PROFILE_DUMP=$OUT/${TAG}_$1_b${B}_dump.csv timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --inflight $B $3 \
      > $OUT/${TAG}_$1_b$B.json 2> $OUT/${TAG}_$1_b$B.err
    python - <<PY
import json
d=json.load(open("$OUT/${TAG}_$1_b$B.json"))
print("[$1] inflight $B: value %.2f solves/s, %.2f ms/step, single %.2f ms, nnz_L %.1fM, flops %.3g, ok %s" % (d["value"], d["ms_per_step"], d["config"]["single_solve_latency_ms"], d["config"]["nnz_L"]/1e6, d["config"]["factor_flops"], d["parity_check"]["ok"]))
PY
  done
}
run cover_default "GMRFB_ND_COVER=1" ""
run boundary_separators "GMRFB_ND_COVER=0" ""
run graph_nd "GMRFB_ND_COVER=1" "--ordering ndgraph"
run graph_nd_balance "GMRFB_ND_COVER=1 GMRFB_ND_BALANCE=0.4" "--ordering ndgraph"
