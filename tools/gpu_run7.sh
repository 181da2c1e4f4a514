#!/bin/bash
# round 2, GPU call 7: symm TMA with the partials kernel split off: 1 vs 2 CTAs per SM, vs cp.async symm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest7.log
tail -5 gpurun_out/r02_pytest7.log
export CHD_PROF_CATS=1
CHD_SYMM_TMA=0 timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress7_symm0.log 2>&1; cat gpurun_out/r02_regress7_symm0.log
CHD_SYMM_TMA_CTAS=1 timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress7_c1.log 2>&1; cat gpurun_out/r02_regress7_c1.log
CHD_SYMM_TMA_CTAS=2 timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress7_c2.log 2>&1; cat gpurun_out/r02_regress7_c2.log
CHD_SYMM_TMA_CTAS=2 timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress7_c2_T64.log 2>&1; cat gpurun_out/r02_regress7_c2_T64.log
