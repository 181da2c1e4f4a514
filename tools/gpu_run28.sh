#!/bin/bash
# round 2, GPU call 28: ncu of the W-formation / Z-update DMMA kernels (renamed after the final script was written)
set -x
mkdir -p gpurun_out
PR="python tools/prof_regress.py 8192 32 quadratic"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wform_dmma|zupd_dmma" -s 40 -c 4 -o gpurun_out/r02_wform_zupd $PR > gpurun_out/r02_ncu_wform_zupd.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_wform_zupd.ncu-rep
