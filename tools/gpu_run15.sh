#!/bin/bash
# round 2, GPU call 15: two-rows-per-thread panel QR (parity, C4 wave experiment, C5 size), C5 bench line
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_twostage.py -m gpu -q -x -k "band_reduce_backtransform or regress_two_stage" > gpurun_out/r02_pytest15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest15.log
export CHD_PROF_CATS=1
python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress15.log 2>&1; tail -3 gpurun_out/r02_regress15.log
CHD_QR_REG_MAXCS=8 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress15_cs8.log 2>&1; tail -3 gpurun_out/r02_regress15_cs8.log
CHD_QR_REG_MAXCS=4 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress15_cs4.log 2>&1; tail -3 gpurun_out/r02_regress15_cs4.log
python tools/prof_regress.py 16384 16 quadratic > gpurun_out/r02_regress15_c5.log 2>&1; tail -3 gpurun_out/r02_regress15_c5.log
CHD_QR_REG=0 python tools/prof_regress.py 16384 8 quadratic > gpurun_out/r02_regress15_c5_fallback.log 2>&1; tail -3 gpurun_out/r02_regress15_c5_fallback.log
unset CHD_PROF_CATS
timeout 900 python bench.py --n 16384 --N 2048 --kernel gaussian --targets 16 --steps 2 --warmup 1 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo "c5 bench rc=$?"; tail -c 800 gpurun_out/r02_bench_c5.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_c5.json") if l.startswith("{")][-1])
r = d["roofline"]
print(d["value"], d["e2e"]["value"], d["ms_per_step"], r["frac"], r["stage1"]["frac_of_dmma_peak"])
print({k: (round(v["ms"], 1), round(v["share_of_step"], 3)) for k, v in r["kernels"].items()})
PY
