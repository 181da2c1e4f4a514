#!/bin/bash
# round 2, GPU call 3: TMA trailing-update kernel: parity (two-stage tests) + timing against the cp.async kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest3.log
tail -15 gpurun_out/r02_pytest3.log
export CHD_PROF_CATS=1
CHD_SYR2K_TMA=0 timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress_tma0.log 2>&1; cat gpurun_out/r02_regress_tma0.log
CHD_SYR2K_TMA=1 timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress_tma1.log 2>&1; cat gpurun_out/r02_regress_tma1.log
CHD_SYR2K_TMA=1 CHD_KD=1 timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress_tma1_kd1.log 2>&1; cat gpurun_out/r02_regress_tma1_kd1.log
