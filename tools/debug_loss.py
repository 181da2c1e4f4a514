import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from computationalhypergraphdiscovery_b200 import _native as nat
from computationalhypergraphdiscovery_b200.synthetic import make_data
from oracle import kernels as okern, regression as oreg
import tridiag_model as tm
import twostage_model as ts
np.set_printoptions(precision=6, linewidth=200)
n = 260
X, names = make_data(n, 9, seed=3)
X = (X - X.mean(0)) / X.std(0)
spec = okern.make_specs([("linear", 0.5, None), ("quadratic", 0.5, None), ("gaussian", 0.7, 1.5)])[2]
for active in ([5], [6, 7, 8], [0, 1, 2, 5]):
    w = np.zeros(9); w[active] = 1
    ga = X[:, 4].copy()
    K = okern.gram(spec, X, X, w)
    lam, V = np.linalg.eigh(K)
    P = V.T @ ga
    grid = oreg.GRID[::50]
    ref = np.array([oreg.make_loss(lam, P)(g)[0] for g in grid])
    plan = nat.Plan(n, 9, 9)
    dev = plan.device
    def kd():
        Kd = torch.zeros((1, n, plan.ldk), dtype=torch.float64, device=dev)
        Kd[0, :, :n] = torch.from_numpy(K).to(dev)
        return Kd
    gad = torch.from_numpy(ga[None]).to(dev).contiguous()
    d1, e1, b1 = plan.tridiag(kd(), gad)
    band, b0 = plan.band_reduce(kd(), gad)
    d2, e2 = plan.band_chase(band)
    d1, e1, d2, e2 = (x.cpu().numpy()[0] for x in (d1, e1, d2, e2))
    l1 = np.array([tm.loss_and_grad(d1, e1, g)[0] for g in grid])
    l2 = np.array([tm.loss_and_grad(d2, e2, g)[0] for g in grid])
    # model: numpy two-stage
    m = ts.stage1(K, ga, 32)
    md, me = ts.stage2(m["band"])
    l3 = np.array([tm.loss_and_grad(md, np.r_[me, 0.0], g)[0] for g in grid])
    # band from CUDA -> numpy stage2
    md4, me4 = ts.stage2(band.cpu().numpy()[0][:, :33])
    l4 = np.array([tm.loss_and_grad(md4, np.r_[me4, 0.0], g)[0] for g in grid])
    print("active", active, "lam range", lam[0], lam[-1])
    print(" grid           ", grid[::20])
    print(" ref loss       ", ref[::20])
    print(" one-stage relerr", (np.abs(l1 - ref) / np.abs(ref))[::20], "max", (np.abs(l1 - ref) / np.abs(ref)).max())
    print(" two-stage relerr", (np.abs(l2 - ref) / np.abs(ref))[::20], "max", (np.abs(l2 - ref) / np.abs(ref)).max())
    print(" numpy 2st relerr", (np.abs(l3 - ref) / np.abs(ref))[::20], "max", (np.abs(l3 - ref) / np.abs(ref)).max())
    print(" cuda band+np chase", (np.abs(l4 - ref) / np.abs(ref)).max())
    lamT = np.linalg.eigvalsh(np.diag(d2) + np.diag(e2[:n-1], 1) + np.diag(e2[:n-1], -1))
    print(" eig err two-stage", np.abs(lamT - lam).max(), "one-stage", np.abs(np.linalg.eigvalsh(np.diag(d1) + np.diag(e1[:n-1], 1) + np.diag(e1[:n-1], -1)) - lam).max())
