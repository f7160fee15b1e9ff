#!/bin/bash
# 1-minute smoke of the rebuilt library (entry points wrapped in function-try-blocks): smoke() and the core sparse tests.
set -u
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02k_smoke.txt
timeout 50 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_fem2d.py -x -q -m gpu > gpurun_out/r02k_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02k_pytest.txt
