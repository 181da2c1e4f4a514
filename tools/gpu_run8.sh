#!/bin/bash
# round 2, GPU call 8: full GPU suite, bench at the new defaults, ncu of the final symm kernel, launch list of a bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest8.log
tail -4 gpurun_out/r02_pytest8.log
CHD_PROF_CATS=1 timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress8_T64.log 2>&1; cat gpurun_out/r02_regress8_T64.log
timeout 1800 python bench.py --steps 3 --warmup 3 --no-fit-wall --no-cpu-baseline > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r02_bench8.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench8.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_kernel", d["per_kernel"], "s1", d["stage1_regressions_per_sec"])
    print({k: (round(v["ms"], 1), round(v["share_of_step"], 3)) for k, v in d["roofline"]["kernels"].items()})
    print("roofline", d["roofline"]["kernel"][:40], d["roofline"]["frac"], "stage1", d["roofline"]["stage1"]["frac_of_dmma_peak"])
    print("strong", d.get("strong_scaling", {}).get("node_prune_steps_per_sec"))
except Exception as e:
    print("bench parse failed", e)
PY
PR="python tools/prof_regress.py 8192 32 quadratic"
$PR > gpurun_out/r02_prof_plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:symm_tma_kernel -s 280 -c 1 -o gpurun_out/r02_symm_tma2 $PR > gpurun_out/r02_ncu_symm_tma2.log 2>&1
tail -2 gpurun_out/r02_ncu_symm_tma2.log
B="python bench.py --steps 1 --warmup 1 --targets 16 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel"
$B > gpurun_out/r02_plain_bench_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log; ls -la gpurun_out/r02_launches_bench.csv
