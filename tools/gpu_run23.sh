#!/bin/bash
# round 2, GPU call 23: 64-target shape of gauss_shared_kernel: parity + A/B on the Gaussian chain
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "shared_gaussian or activations" > gpurun_out/r02_pytest23.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest23.log
B="python bench.py --kernel gaussian --steps 2 --warmup 1 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel"
$B > gpurun_out/r02_bench23_g64.json 2> gpurun_out/r02_bench23_g64.err; echo "rc=$?"
CHD_GAUSS_SHARED=2 $B > gpurun_out/r02_bench23_g32.json 2> gpurun_out/r02_bench23_g32.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("g64", "g32"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r02_bench23_{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 1))
    except Exception as e:
        print(f, "failed", e)
PY
