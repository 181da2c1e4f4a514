#!/bin/bash
# round 2, GPU call 26: bulge chasing with a guard-free copy of the interior hops and constant-offset addressing
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest26.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest26.log
CHD_PROF_CATS=1 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress26.log 2>&1; tail -3 gpurun_out/r02_regress26.log
CHD_PROF_CATS=1 python tools/prof_regress.py 600 8 quadratic > gpurun_out/r02_regress26_n600.log 2>&1; tail -3 gpurun_out/r02_regress26_n600.log
