#!/bin/bash
# round 2, GPU call 6: TMA symm kernel: parity + timing against the cp.async kernel + ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest6.log
tail -12 gpurun_out/r02_pytest6.log
export CHD_PROF_CATS=1
CHD_SYMM_TMA=0 timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress_symm0.log 2>&1; cat gpurun_out/r02_regress_symm0.log
CHD_SYMM_TMA=1 timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress_symm1.log 2>&1; cat gpurun_out/r02_regress_symm1.log
unset CHD_PROF_CATS
PR="python tools/prof_regress.py 8192 32 quadratic"
$PR > gpurun_out/r02_prof_plain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:symm_tma_kernel -s 280 -c 2 -o gpurun_out/r02_symm_tma $PR > gpurun_out/r02_ncu_symm_tma.log 2>&1
tail -3 gpurun_out/r02_ncu_symm_tma.log
