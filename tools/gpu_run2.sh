#!/bin/bash
# round 2, GPU call 2: full parity suite on the deferred-update path, KD=1 vs KD=2 timing, short bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest2.log
tail -15 gpurun_out/r02_pytest2.log
export CHD_PROF_CATS=1
CHD_KD=1 python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress_kd1.log 2>&1; cat gpurun_out/r02_regress_kd1.log
python tools/prof_regress.py 8192 32 quadratic > gpurun_out/r02_regress_kd2.log 2>&1; cat gpurun_out/r02_regress_kd2.log
python bench.py --steps 3 --warmup 1 --no-fit-wall --no-strong --no-cpu-baseline > gpurun_out/r02_bench_quick2.json 2> gpurun_out/r02_bench_quick2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_quick2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_quick2.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_kernel", d["per_kernel"])
    print({k: (round(v["ms"], 1), round(v["share_of_step"], 3)) for k, v in d["roofline"]["kernels"].items()})
    print("roofline frac", d["roofline"]["frac"], "stage1", d["roofline"]["stage1"]["frac_of_dmma_peak"])
except Exception as e:
    print("bench parse failed", e)
PY
