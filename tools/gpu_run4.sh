#!/bin/bash
# round 2, GPU call 4: shared Gaussian activations (parity + timing), chase waves, ncu of the TMA trailing update
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fit.py -m gpu -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest4.log
tail -25 gpurun_out/r02_pytest4.log
export CHD_PROF_CATS=1
for w in 16 8; do
  CHD_CHASE_WAVE=$w timeout 300 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress_wave$w.log 2>&1; cat gpurun_out/r02_regress_wave$w.log
done
unset CHD_PROF_CATS
timeout 900 python bench.py --steps 3 --warmup 1 --no-fit-wall --no-strong --no-cpu-baseline > gpurun_out/r02_bench_quick4.json 2> gpurun_out/r02_bench_quick4.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_quick4.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_quick4.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_kernel", d["per_kernel"])
    print({k: (round(v["ms"], 1), round(v["share_of_step"], 3)) for k, v in d["roofline"]["kernels"].items()})
    print("roofline frac", d["roofline"]["frac"], "stage1", d["roofline"]["stage1"]["frac_of_dmma_peak"])
except Exception as e:
    print("bench parse failed", e)
PY
PR="python tools/prof_regress.py 8192 32 quadratic"
$PR > gpurun_out/r02_prof_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:syr2k_tma_kernel -s 140 -c 2 -o gpurun_out/r02_syr2k_tma $PR > gpurun_out/r02_ncu_syr2k_tma.log 2>&1
tail -3 gpurun_out/r02_ncu_syr2k_tma.log
