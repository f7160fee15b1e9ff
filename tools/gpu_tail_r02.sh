#!/bin/bash
# Last seconds of the round-2 GPU budget: the block-tridiagonal and P1 / 1-D assembly tests on the final library (the
# function-try-block change touched btd.cu / fem.cu / fem1d.cu mechanically; tools/gpu_smoke_r02.sh covered the rest).
set -u
mkdir -p gpurun_out
timeout 24 python -m pytest tests/test_gpu_btd.py tests/test_gpu_fem.py -x -q -m gpu > gpurun_out/r02m_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02m_pytest.txt
