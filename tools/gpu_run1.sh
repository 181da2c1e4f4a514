#!/bin/bash
# round 2, GPU call 1: parity tests, stream-overlap probe, ncu of the latency-bound kernels (final versions)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
timeout 600 python tools/overlap_probe.py --steps 3 --modes seq,k3,h6 > gpurun_out/r02_overlap1.log 2>&1; echo "overlap rc=$?"
tail -5 gpurun_out/r02_overlap1.log
PR="python tools/prof_regress.py 8192 32 quadratic"
$PR > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chase_cta_kernel -s 1 -c 1 -o gpurun_out/r02_chase $PR > gpurun_out/r02_ncu_chase.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"panel_qr_reg_kernel|wform_kernel" -s 540 -c 4 -o gpurun_out/r02_qr_wform $PR > gpurun_out/r02_ncu_qr.log 2>&1
cat gpurun_out/r02_prof_plain.log
ls -la gpurun_out | tail -8
