#!/bin/bash
# round 2, final GPU evidence (second half of the round): full parity suite, smoke, bench line, reference arm, C5 line,
# ncu captures of the new kernels, launch list of one mixed step
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_final.log
tail -4 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
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
timeout 700 python bench.py --impl reference --steps 6 --warmup 1 --ref-budget-s 110 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --n 16384 --N 2048 --kernel gaussian --targets 16 --steps 2 --warmup 1 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo "c5 rc=$?"
PR="python tools/prof_regress.py 8192 32 quadratic"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wz_dmma|zpart_dmma|partials_dmma|pupd_dmma|panel_qr_reg" -s 300 -c 10 -o gpurun_out/r02_small_dmma $PR > gpurun_out/r02_ncu_small_dmma.log 2>&1; echo "ncu small rc=$?"
B="python bench.py --steps 1 --warmup 1 --targets 8 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 14000 -c 5500 --kill 1 --csv --log-file gpurun_out/r02_launches_bench_final.csv $B > gpurun_out/r02_ncu_launches_final.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/r02_launches_bench_final.csv
