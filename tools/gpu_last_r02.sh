#!/bin/bash
# Last record of round 2 on the shipped defaults (after the parallel nested dissection): GPU suite, the bench line at the
# driver's step counts, the P2 Darcy configuration with device-resident normals.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r02h_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02h_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r02h_bench.json 2> $OUT/r02h_bench.err; echo "bench rc=$?"
timeout 300 python tools/bench_fem2d.py --nel 300 --out $OUT/r02h_fem2d_darcy_p2.json > $OUT/r02h_fem2d.log 2>&1; echo "fem2d rc=$?"; tail -1 $OUT/r02h_fem2d.log
GMRFB_SYM_TIMING=1 timeout 200 python tools/probe/symtime.py > $OUT/r02h_symtime.txt 2>&1; echo "symtime rc=$?"; tail -12 $OUT/r02h_symtime.txt
