#!/bin/bash
# round 2, GPU call 19: DMMA thin update, W and Z halves as separate CTAs: parity + timing
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest19.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest19.log
export CHD_PROF_CATS=1
python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress19.log 2>&1; tail -3 gpurun_out/r02_regress19.log
unset CHD_PROF_CATS
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --kill 1 --csv --log-file gpurun_out/r02_launches_regress64c.csv python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_ncu_regress64c.log 2>&1; echo "ncu rc=$?"
python tools/ncu_launch_summary.py gpurun_out/r02_launches_regress64c.csv | cut -c1-200
