#!/bin/bash
# round 2, GPU call 22: what a Gaussian prune step costs outside the regression (launch list at 64 in flight) + ncu of
# gauss_shared_kernel
set -x
mkdir -p gpurun_out
B="python bench.py --kernel gaussian --steps 1 --warmup 1 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel"
K='regex:gauss|gram|gmat|gdot|act_|quad_rows|members|transpose|cluster_union|removed|active_vars|compact|gather|choose|normal|split|gamma'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r02_launches_gauss_step.csv $B > gpurun_out/r02_ncu_gauss_step.log 2>&1; echo "ncu rc=$?"
python tools/ncu_launch_summary.py gpurun_out/r02_launches_gauss_step.csv 1 | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gauss_shared_kernel -s 2 -c 1 -o gpurun_out/r02_gauss_shared $B > gpurun_out/r02_ncu_gauss_shared.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r02_gauss_shared.ncu-rep
