#!/bin/bash
# round 2, GPU call 25: how the bulge chasing scales with resident CTAs per SM
set -x
mkdir -p gpurun_out
for c in 8 6 4; do
CHD_PROF_CATS=1 CHD_CHASE_CTAS_PER_SM=$c python tools/prof_regress.py 8192 64 quadratic > gpurun_out/r02_regress25_c$c.log 2>&1; tail -2 gpurun_out/r02_regress25_c$c.log | head -1
done
