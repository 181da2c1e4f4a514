"""Minimal harness for ncu captures of the regression kernels at the bench's size: T Gram matrices (quadratic mode,
n x n), one warm regression, one profiled regression.  python tools/prof_regress.py [n] [T] [kernel]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from computationalhypergraphdiscovery_b200 import _native as nat  # noqa: E402
from computationalhypergraphdiscovery_b200.synthetic import make_data  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
kind = sys.argv[3] if len(sys.argv) > 3 else "quadratic"
N = 64
X, _ = make_data(n, N, seed=0)
X = np.ascontiguousarray((X - X.mean(0)) / X.std(0))
plan = nat.Plan(n, N, N)
dev = plan.device
Xd = torch.from_numpy(X).to(dev)
desc = {"linear": nat.make_desc(0, 2.0), "quadratic": nat.make_desc(1, 1.0, alpha=1.0),
        "gaussian": nat.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)}[kind]
w = torch.ones((T, N), dtype=torch.uint8, device=dev)
ga = torch.empty((T, n), dtype=torch.float64, device=dev)
for t in range(T):
    w[t, t % N] = 0
    ga[t] = Xd[:, t % N]
keys = torch.zeros((T, 2), dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    K = plan.gram(desc, Xd, w)
    gmin = torch.full((T,), 1e-9, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    if rep == 1 and os.environ.get("CHD_PROF_CATS"):
        nat.profile_enable(True)
    e0.record()
    r = plan.regress(K, ga, keys, gmin, mode=1)
    e1.record()
    torch.cuda.synchronize()
    if rep == 1 and os.environ.get("CHD_PROF_CATS"):
        cats = nat.profile_read_all()
        nat.profile_enable(False)
        print("   per matrix (ms): " + ", ".join(f"{k} {v[0] / T:.2f}" for k, v in cats.items()), flush=True)
    print(f"rep {rep}: regress n={n} T={T} {kind}: {e0.elapsed_time(e1):.1f} ms, {e0.elapsed_time(e1) / T:.2f} ms/matrix, "
          f"noise[0]={r['noise'][0].item():.12g} gamma[0]={r['gamma'][0].item():.8g}", flush=True)
    del K
