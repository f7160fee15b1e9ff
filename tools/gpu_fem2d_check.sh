#!/bin/bash
# First GPU run of the Lagrange-triangle assembler and the sparse-product plan: their GPU tests (no -x: every failure
# is wanted), the P2 Darcy bench, and config 5 at 8 steps on the shipped wide-inverse bounds.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 420 python -m pytest tests/test_gpu_fem2d.py -q -m gpu > $OUT/fem2d_pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 $OUT/fem2d_pytest.txt
timeout 300 python tools/bench_fem2d.py --nel 300 --out $OUT/fem2d_darcy_p2.json > $OUT/fem2d_bench.log 2>&1; echo "bench rc=$?"; tail -3 $OUT/fem2d_bench.log
timeout 240 python tools/bench_config5.py --nx 256 --steps 8 --out $OUT/config5_n8.json > $OUT/config5_n8.log 2>&1; echo "config5 rc=$?"; tail -2 $OUT/config5_n8.log
