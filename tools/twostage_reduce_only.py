"""One band_reduce + chase + solve + back-transform at full size (for ncu launch lists).  python tools/twostage_reduce_only.py [n] [T]"""
import sys

import torch

from computationalhypergraphdiscovery_b200 import _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
N = 64
plan = nat.Plan(n, N, N)
dev = plan.device
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn((n, N), dtype=torch.float64, device=dev, generator=g)
desc = nat.make_desc(nat.KERNEL_GAUSSIAN, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)
w = torch.ones((T, N), dtype=torch.uint8, device=dev)
w[:, 0] = 0
ga = X[:, :1].T.repeat(T, 1).contiguous()
S = torch.randn((T, n, 100), dtype=torch.float64, device=dev, generator=g)
K = plan.gram(desc, X, w)
band, b0 = plan.band_reduce(K, ga, S)
d, e = plan.band_chase(band)
r = plan.band_solve(band, torch.full((T,), 0.5, dtype=torch.float64, device=dev), S)
out = plan.band_backtransform(K, r["x"], b0)
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))
