#!/bin/bash
# round 2, GPU call 21: three chains on three streams (32 targets each), chase CTAs capped so that a GEMM CTA fits beside them
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --targets 32 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel"
$B > gpurun_out/r02_bench21_t32.json 2> gpurun_out/r02_bench21_t32.err; echo "rc=$?"
$B --streams 3 > gpurun_out/r02_bench21_t32s3.json 2> gpurun_out/r02_bench21_t32s3.err; echo "rc=$?"
CHD_CHASE_CTAS_PER_SM=5 $B --streams 3 > gpurun_out/r02_bench21_t32s3c5.json 2> gpurun_out/r02_bench21_t32s3c5.err; echo "rc=$?"
CHD_CHASE_CTAS_PER_SM=3 $B --streams 3 > gpurun_out/r02_bench21_t32s3c3.json 2> gpurun_out/r02_bench21_t32s3c3.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("t32", "t32s3", "t32s3c5", "t32s3c3"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r02_bench21_{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 1))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r02_bench21_t32s3c5.err
