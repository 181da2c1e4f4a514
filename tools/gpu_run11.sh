#!/bin/bash
# round 2, GPU call 11: trailing update with two 64-row CTAs per SM vs one 128-row CTA
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest11.log
tail -5 gpurun_out/r02_pytest11.log
export CHD_PROF_CATS=1
timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress11.log 2>&1; cat gpurun_out/r02_regress11.log
CHD_SYR2K_TMA_CTAS=1 timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress11_c1.log 2>&1; cat gpurun_out/r02_regress11_c1.log
