#!/bin/bash
# round 2, GPU call 27: bulge chasing, three-hop lag with register prefetch of the next hop's blocks
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest27.log
CHD_PROF_CATS=1 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress27.log 2>&1; tail -3 gpurun_out/r02_regress27.log
CHD_PROF_CATS=1 CHD_CHASE_LAG=2 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress27_lag2.log 2>&1; tail -3 gpurun_out/r02_regress27_lag2.log
CHD_PROF_CATS=1 CHD_CHASE_LAG=3 python tools/prof_regress.py 600 8 quadratic > gpurun_out/r02_regress27_n600_lag3.log 2>&1; tail -3 gpurun_out/r02_regress27_n600_lag3.log
