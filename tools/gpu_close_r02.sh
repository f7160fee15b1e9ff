#!/bin/bash
# Closing record of round 2 on the shipped code (threaded plan builders included): the full GPU suite, the P2 Darcy
# configuration (plan-construction times), and an ncu capture of the new assembly / sparse-product kernels.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/r02j_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02j_pytest.txt
timeout 200 python tools/bench_fem2d.py --nel 300 --out $OUT/r02j_fem2d_darcy_p2.json > $OUT/r02j_fem2d.log 2>&1; echo "fem2d rc=$?"; tail -1 $OUT/r02j_fem2d.log
timeout 240 ncu --set full --clock-control none -k regex:"k_fem2d_stiffness|k_fem2d_cubic_J|k_spgemm|k_fem2d_geom" -c 6 -f -o $OUT/r02j_fem2d_prof \
    python tools/bench_fem2d.py --nel 300 --no-cpu --problems 2 > $OUT/r02j_ncu_fem2d.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py full $OUT/r02j_fem2d_prof.ncu-rep $OUT/r02j_ncu_fem2d > /dev/null 2>&1; echo "summ rc=$?"
rm -f $OUT/r02j_fem2d_prof.ncu-rep
cat $OUT/r02j_ncu_fem2d.md 2>/dev/null | head -12
