#!/bin/bash
# ncu evidence for round 1 (run under gpurun): launch list of the bench command + one full capture of the
# dominant kernel.  Each ncu run follows a plain run of the same command that exited 0.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-per-kernel"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
PROBE="python tools/perf_probe.py --n 8192 --N 128 --kinds 1"
$PROBE > gpurun_out/plain_probe.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tridiag_kernel -c 1 -o gpurun_out/tridiag_n8192_v3 $PROBE > gpurun_out/ncu_probe.log 2>&1
PROBE2="python tools/perf_probe.py --n 2048 --N 32 --kinds 2"
$PROBE2 > gpurun_out/plain_probe2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gauss_pairs|gram_kernel" -c 2 -o gpurun_out/gauss_n2048 $PROBE2 > gpurun_out/ncu_probe2.log 2>&1
ls -la gpurun_out/
