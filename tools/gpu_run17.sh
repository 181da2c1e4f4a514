#!/bin/bash
# round 2, GPU call 17: FP64 tensor + vector pipe probe, per-kernel launch list of the regression at 64 in flight
set -x
mkdir -p gpurun_out
python tools/fp64_mix_probe.py > gpurun_out/r02_fp64_mix_probe.json 2>&1; cat gpurun_out/r02_fp64_mix_probe.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 1300 --kill 1 --csv --log-file gpurun_out/r02_launches_regress64.csv python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_ncu_regress64.log 2>&1; echo "ncu rc=$?"
python tools/ncu_launch_summary.py gpurun_out/r02_launches_regress64.csv 150 | cut -c1-400
