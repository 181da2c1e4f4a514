#!/bin/bash
# round 2, GPU call 5: DMMA occupancy probe, 64 targets in flight, full bench line (all features), reference arm
set -x
mkdir -p gpurun_out
python tools/dmma_probe.py > gpurun_out/r02_dmma_probe.json 2>&1; cat gpurun_out/r02_dmma_probe.json
CHD_PROF_CATS=1 timeout 300 python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress_T64.log 2>&1; cat gpurun_out/r02_regress_T64.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_full5.json 2> gpurun_out/r02_bench_full5.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_full5.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_full5.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_kernel", d["per_kernel"])
    print("fit_wall", json.dumps(d.get("fit_wall_s"))[:900])
    print("strong", d.get("strong_scaling"))
    print("cpu", {k: v for k, v in d.get("cpu_baseline", {}).items() if k != "sample"})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 --ref-budget-s 60 > gpurun_out/r02_bench_ref5.json 2> gpurun_out/r02_bench_ref5.err; echo "ref rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_ref5.json')); print({k: d[k] for k in ('value','steps','ms_per_step','timed_region_s','setup_s')}, d['cpu_baseline']['cores'], d['cpu_baseline']['per_kernel'])"
