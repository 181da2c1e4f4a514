#!/bin/bash
# round 2, GPU call 30: compute-sanitizer memcheck over the kernels added in the second half of the round
set -x
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_kernels.py -m gpu -q -x -k "(band_reduce and (97- or 333- or 1000-two1 or 333-fma)) or (shared_gaussian and 40-37-)" > gpurun_out/r02_sanitizer_memcheck2.log 2>&1; echo "sanitizer rc=$?"
tail -15 gpurun_out/r02_sanitizer_memcheck2.log
