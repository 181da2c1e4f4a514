#!/bin/bash
# round 2, GPU call 13: ncu of the final chase kernel and of the small per-panel kernels
set -x
mkdir -p gpurun_out
PR="python tools/prof_regress.py 8192 32 quadratic"
$PR > gpurun_out/r02_prof_plain13.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chase_cta2_kernel -s 1 -c 1 -o gpurun_out/r02_chase2 $PR > gpurun_out/r02_ncu_chase2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"symm_partials_kernel|zupdate_kernel|wform_kernel|panel_qr_reg_kernel|syr2k_tma_kernel" -s 1200 -c 6 -o gpurun_out/r02_small $PR > gpurun_out/r02_ncu_small.log 2>&1
tail -2 gpurun_out/r02_ncu_small.log
