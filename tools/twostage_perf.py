"""Times the two-stage regression pieces at full size (CUDA events), for tuning.  Usage:
python tools/twostage_perf.py [n] [T]"""
import sys
import time

import numpy as np
import torch

from computationalhypergraphdiscovery_b200 import _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
what = sys.argv[3] if len(sys.argv) > 3 else "all"
N = 64
plan = nat.Plan(n, N, N)
dev = plan.device
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn((n, N), dtype=torch.float64, device=dev, generator=g)
desc = nat.make_desc(nat.KERNEL_GAUSSIAN, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)
w = torch.ones((T, N), dtype=torch.uint8, device=dev)
w[:, 0] = 0
ga = X[:, :1].T.repeat(T, 1).contiguous()


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        K = plan.gram(desc, X, w)
        torch.cuda.synchronize()
        e0.record()
        fn(K)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


S = torch.randn((T, n, 100), dtype=torch.float64, device=dev, generator=g)


def reduce_only(K=None):
    if K is None:
        K = plan.gram(desc, X, w)
    return plan.band_reduce(K, ga, S.clone())


ms = timed(reduce_only)
print(f"band_reduce n={n} T={T}: {ms:.1f} ms total, {ms / T:.1f} ms/matrix, "
      f"{4 / 3 * n ** 3 * T / ms / 1e9:.2f} TFLOP/s credited")
K = plan.gram(desc, X, w)
band, b0 = plan.band_reduce(K, ga, S.clone())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("band_chase", lambda: plan.band_chase(band)),
                 ("band_solve", lambda: plan.band_solve(band, torch.full((T,), 0.5, dtype=torch.float64, device=dev),
                                                        S.clone())),
                 ("backtransform", lambda: plan.band_backtransform(K, S[:, :, 0].contiguous(), b0))):
    fn()
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.2f} ms total, {e0.elapsed_time(e1) / T:.2f} ms/matrix")

for two in (0, 1):
    plan.set_two_stage(two)
    keys = torch.zeros((T, 2), dtype=torch.int32, device=dev)
    gmin = torch.full((T,), 1e-9, dtype=torch.float64, device=dev)
    best = 1e30
    for _ in range(2):
        K = plan.gram(desc, X, w)
        torch.cuda.synchronize()
        e0.record()
        r = plan.regress(K, ga, keys, gmin, mode=1)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"regress two_stage={two}: {best:.1f} ms total, {best / T:.1f} ms/matrix; noise[0]={r['noise'][0].item():.12g} "
          f"gamma[0]={r['gamma'][0].item():.8g}")
