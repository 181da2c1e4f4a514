#!/bin/bash
# round 2, GPU call 10: 5-barrier chase kernel: parity + timing vs the first formulation
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest10.log
tail -5 gpurun_out/r02_pytest10.log
export CHD_PROF_CATS=1
timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress10.log 2>&1; cat gpurun_out/r02_regress10.log
CHD_CHASE=1 timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress10_c1.log 2>&1; cat gpurun_out/r02_regress10_c1.log
timeout 300 python tools/prof_regress.py 8192 8 quadratic > gpurun_out/r02_regress10_T8.log 2>&1; cat gpurun_out/r02_regress10_T8.log
