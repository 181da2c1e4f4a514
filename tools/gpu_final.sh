#!/bin/bash
# round 2: final GPU evidence -- full parity suite, smoke, bench line (all features), reference arm, launch list
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_final.log
tail -4 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 2400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_final.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_final.json") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_kernel", d["per_kernel"], "s1", d["stage1_regressions_per_sec"])
    print({k: (round(v["ms"], 1), round(v["share_of_step"], 3)) for k, v in d["roofline"]["kernels"].items()})
    print("roofline", d["roofline"]["kernel"][:30], d["roofline"]["frac"], d["roofline"]["gemm_kernels"], "stage1", d["roofline"]["stage1"]["frac_of_dmma_peak"])
    print("fit_wall", {k: v["fit_s"] for k, v in d["fit_wall_s"].items()}, "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 900 python bench.py --impl reference --steps 8 --warmup 1 --ref-budget-s 150 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
B="python bench.py --steps 1 --warmup 1 --targets 8 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel"
$B > gpurun_out/r02_plain_bench_launches2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 14000 -c 8000 --csv --log-file gpurun_out/r02_launches_bench_final.csv $B > gpurun_out/r02_ncu_launches2.log 2>&1
ls -la gpurun_out/r02_launches_bench_final.csv
