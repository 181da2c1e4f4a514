"""Do the FP64 tensor pipe (DMMA) and the FP64 vector pipe (DFMA) add up?  mode 0: DMMA only, 1: DFMA only,
2: both interleaved at equal flops in every warp (chd_fp64_probe).  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from computationalhypergraphdiscovery_b200 import _native as nat  # noqa: E402

nat.Plan(64, 4, 4)
out = {name: nat.fp64_probe_tflops(mode, 20000) for mode, name in ((0, "dmma_only"), (1, "dfma_only"), (2, "interleaved"))}
out["note"] = ("TFLOP/s of 148 x 4 CTAs x 8 warps; interleaved = 8 DMMA + 64 DFMA per iteration and lane group "
               "(equal flops): a sum close to dmma_only + dfma_only means separate execution resources")
print(json.dumps(out, indent=1))
