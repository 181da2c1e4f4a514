#!/bin/bash
# round 2, GPU call 31: back-transformation with four rows' loads in flight per warp: parity + timing
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_twostage.py -m gpu -q -x -k "band_reduce or regress" > gpurun_out/r02_pytest31.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest31.log
CHD_PROF_CATS=1 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress31.log 2>&1; tail -3 gpurun_out/r02_regress31.log
