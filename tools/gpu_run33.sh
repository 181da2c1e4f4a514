#!/bin/bash
# round 2, GPU call 33: ncu of the remaining per-regression kernels (band solve, back-transformation, gamma search)
set -x
mkdir -p gpurun_out
PR="python tools/prof_regress.py 8192 32 quadratic"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"band_solve|backtransform2|gamma_grid|gamma_finish" -c 4 -o gpurun_out/r02_tail_kernels $PR > gpurun_out/r02_ncu_tail.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_tail_kernels.ncu-rep
