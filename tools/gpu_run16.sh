#!/bin/bash
# round 2, GPU call 16: DMMA small-product kernels + two-row QR wave policy: parity, timing A/B
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest16.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest16.log
export CHD_PROF_CATS=1
python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress16.log 2>&1; tail -3 gpurun_out/r02_regress16.log
CHD_SMALL_DMMA=0 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress16_fma.log 2>&1; tail -3 gpurun_out/r02_regress16_fma.log
CHD_QR_TWO=2 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress16_two2.log 2>&1; tail -3 gpurun_out/r02_regress16_two2.log
CHD_QR_TWO=0 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress16_two0.log 2>&1; tail -3 gpurun_out/r02_regress16_two0.log
python tools/prof_regress.py 8192 32 gaussian > gpurun_out/r02_regress16_T32.log 2>&1; tail -3 gpurun_out/r02_regress16_T32.log
