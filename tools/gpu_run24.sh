#!/bin/bash
# round 2, GPU call 24: W / Z-update as two kernels sized for occupancy (80 / 64 registers): parity + timing + ncu
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_twostage.py -m gpu -q -x > gpurun_out/r02_pytest24.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest24.log
PR="python tools/prof_regress.py 8192 64 quadratic"
CHD_PROF_CATS=1 $PR > gpurun_out/r02_regress24.log 2>&1; tail -3 gpurun_out/r02_regress24.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --kill 1 --csv --log-file gpurun_out/r02_launches_regress64d.csv $PR > /dev/null 2>&1; echo "ncu rc=$?"
python tools/ncu_launch_summary.py gpurun_out/r02_launches_regress64d.csv | cut -c1-200
