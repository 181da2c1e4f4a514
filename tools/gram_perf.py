"""Times chd_gram at the bench shape (n = 8192, N = 512): python tools/gram_perf.py [T]"""
import sys
import torch
from computationalhypergraphdiscovery_b200 import _native as nat

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, N = 8192, 512
plan = nat.Plan(n, N, N)
dev = plan.device
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn((n, N), dtype=torch.float64, device=dev, generator=g)
w = torch.ones((T, N), dtype=torch.uint8, device=dev)
w[:, 0] = 0
out = torch.zeros((T, n, plan.ldk), dtype=torch.float64, device=dev)
for name, desc in (("linear", nat.make_desc(0, 2.0)), ("quadratic", nat.make_desc(1, 1.0, alpha=1.0)),
                   ("gaussian", nat.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0))):
    plan.gram(desc, X, w, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.gram(desc, X, w, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {ms:.1f} ms for {T} matrices, {ms / T:.2f} ms/matrix, {n * n * (N - 1) * T / ms / 1e9:.1f} TFLOP/s (lower triangle)")
