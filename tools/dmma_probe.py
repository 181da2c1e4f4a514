"""DMMA issue-rate ceiling as a function of resident warps per SM and independent accumulator chains per warp."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from computationalhypergraphdiscovery_b200 import _native as nat  # noqa: E402

nat.Plan(64, 4, 4)
out = {"peak_probe": nat.dmma_peak_tflops(20000)}
for w in (4, 8, 12, 16, 32):
    for ch in (2, 4, 8):
        out[f"warps{w}_chains{ch}"] = nat.dmma_occupancy_probe_tflops(w, ch, 20000)
print(json.dumps(out, indent=1))
