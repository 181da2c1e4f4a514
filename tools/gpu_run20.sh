#!/bin/bash
# round 2, GPU call 20: bench line with the new kernels; chains on three streams with room left for the chase CTAs
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-fit-wall --no-cpu-baseline --no-strong --no-per-kernel"
$B > gpurun_out/r02_bench20.json 2> gpurun_out/r02_bench20.err; echo "rc=$?"
$B --streams 3 > gpurun_out/r02_bench20_s3.json 2> gpurun_out/r02_bench20_s3.err; echo "rc=$?"
CHD_CHASE_CTAS_PER_SM=5 $B --streams 3 > gpurun_out/r02_bench20_s3c5.json 2> gpurun_out/r02_bench20_s3c5.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("r02_bench20", "r02_bench20_s3", "r02_bench20_s3c5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 1),
              {k: round(v["ms"] / (3 * 64 * 2), 2) for k, v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r02_bench20_s3c5.err
