"""One Gaussian activation pass at the bench shape (for ncu): python tools/act_once.py [T]"""
import sys
import torch
from computationalhypergraphdiscovery_b200 import _native as nat

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, N = 8192, 512
plan = nat.Plan(n, N, N)
dev = plan.device
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn((n, N), dtype=torch.float64, device=dev, generator=g)
desc = nat.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)
cl = torch.arange(N, dtype=torch.int32, device=dev)
w0 = torch.ones((T, N), dtype=torch.uint8, device=dev)
w0[:, 0] = 0
alive = torch.ones((T, N), dtype=torch.uint8, device=dev)
yb = torch.randn((T, n), dtype=torch.float64, device=dev, generator=g)
ga = X[:, :1].T.repeat(T, 1).contiguous()
act, pruned = plan.activations(desc, X, cl, w0, alive, yb, ga, prune=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
act, pruned = plan.activations(desc, X, cl, w0, alive, yb, ga, prune=False)
e1.record()
torch.cuda.synchronize()
print(f"gaussian activations T={T}: {e0.elapsed_time(e1):.1f} ms, {e0.elapsed_time(e1) / T:.1f} ms/matrix, act[0,1]={act[0, 1].item():.6g}")
