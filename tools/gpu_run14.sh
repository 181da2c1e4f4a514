#!/bin/bash
# round 2, GPU call 14: parallel members_kernel (clusters), C3 fit time + steady-state launch list, C5-size probes
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fit.py -m gpu -q -x -k "activ or cluster or possible" > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest14.log
python tools/fit_configs.py --configs c1,c2,c3 > gpurun_out/r02_fit_configs14.log 2>&1; echo "fit rc=$?"; cat gpurun_out/r02_fit_configs14.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 1500 --kill 1 --csv --log-file gpurun_out/r02_launches_c3.csv python tools/fit_configs.py --configs c3 > gpurun_out/r02_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
python tools/ncu_launch_summary.py gpurun_out/r02_launches_c3.csv | head -40
timeout 600 python tools/twostage_perf.py 16384 8 > gpurun_out/r02_c5_twostage.log 2>&1; echo "c5 twostage rc=$?"; cat gpurun_out/r02_c5_twostage.log
timeout 900 python bench.py --n 16384 --N 2048 --kernel gaussian --targets 8 --steps 2 --warmup 1 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo "c5 bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_c5.err; cut -c1-1200 gpurun_out/r02_bench_c5.json
