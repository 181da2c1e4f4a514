#!/bin/bash
# round 2, GPU call 9: Gram downdate + chase progress caching: parity (kernels, fit, twostage), timing, 2 chase variants
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_twostage.py tests/test_gpu_fit.py -m gpu -q > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest9.log
tail -6 gpurun_out/r02_pytest9.log
export CHD_PROF_CATS=1
timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress9.log 2>&1; cat gpurun_out/r02_regress9.log
CHD_CHASE=0 timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress9_warp.log 2>&1; cat gpurun_out/r02_regress9_warp.log
unset CHD_PROF_CATS
timeout 1800 python bench.py --steps 3 --warmup 3 --no-fit-wall --no-cpu-baseline --no-strong > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r02_bench9.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench9.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_kernel", d["per_kernel"], "s1", d["stage1_regressions_per_sec"])
    print({k: (round(v["ms"], 1), round(v["share_of_step"], 3)) for k, v in d["roofline"]["kernels"].items()})
    print("roofline", d["roofline"]["kernel"][:40], d["roofline"]["frac"], "stage1", d["roofline"]["stage1"]["frac_of_dmma_peak"])
except Exception as e:
    print("bench parse failed", e)
PY
