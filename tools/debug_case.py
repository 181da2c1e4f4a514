import os, sys
import numpy as np
import networkx as nx
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from computationalhypergraphdiscovery_b200 import _native as nat
from computationalhypergraphdiscovery_b200.synthetic import make_data
from oracle import chd as ochd, kernels as okern, regression as oreg
import tridiag_model as tm
np.set_printoptions(precision=10, linewidth=200)
X, names = make_data(260, 9, seed=3)
n = 260
clusters = [names[0:3], names[3:5], names[5:6], names[6:9]]
rng = np.random.default_rng(4)
pe = nx.DiGraph()
pe.add_nodes_from(names)
for a in names:
    for b in names:
        if a != b and rng.random() > 0.2:
            pe.add_edge(a, b)
adj = np.array(nx.adjacency_matrix(pe, nodelist=names).todense())
res = ochd.fit(X, names, kernels=[("linear", 0.5, None), ("quadratic", 0.5, None), ("gaussian", 0.7, 1.5)],
               clusters=clusters, adjacency=adj, gamma_min=1e-8, kernel_chooser=("threshold", 0.9),
               mode_chooser=("threshold", 0.5))
ch = res["chains"][2]
print("targets", ch["targets"], "active", ch["active"][2])
Xn = (X - X.mean(0)) / X.std(0)
# last step of target v4: which variables are active?  anc = per-step variable masks (T, steps+1, N)
w = ch["active"][2][-1].astype(float)
print("w", w)
spec = okern.make_specs([("linear", 0.5, None), ("quadratic", 0.5, None), ("gaussian", 0.7, 1.5)])[2]
K = okern.gram(spec, Xn, Xn, w)
ga = Xn[:, 4].copy()
lam, V = np.linalg.eigh(K)
P = V.T @ ga
lg = oreg.make_loss(lam, P)
ref = oreg.grid_losses(lam, P)
print("oracle grid argmin", np.argmin(ref), oreg.GRID[np.argmin(ref)], ref.min())
plan = nat.Plan(n, 9, 9)
dev = plan.device
def kd():
    Kd = torch.zeros((1, n, plan.ldk), dtype=torch.float64, device=dev)
    Kd[0, :, :n] = torch.from_numpy(K).to(dev)
    return Kd
gad = torch.from_numpy(ga[None]).to(dev).contiguous()
band, b0 = plan.band_reduce(kd(), gad)
d2, e2 = plan.band_chase(band)
d2, e2 = d2.cpu().numpy()[0], e2.cpu().numpy()[0]
d1, e1, _ = plan.tridiag(kd(), gad)
d1, e1 = d1.cpu().numpy()[0], e1.cpu().numpy()[0]
for nm, (d, e) in (("one", (d1, e1)), ("two", (d2, e2))):
    l = np.array([tm.loss_and_grad(d, e, g)[0] for g in oreg.GRID[::10]])
    r = ref[::10]
    i = np.argmin(l)
    print(nm, "grid argmin", i * 10, oreg.GRID[i * 10], l.min(), "max relerr", (np.abs(l - r) / r).max())
idx = [np.searchsorted(oreg.GRID, g) for g in (1e-5, 5e-5, 1.4e-4, 3e-4, 5.5e-4, 1e-3, 1e-2)]
print("ref losses at", [oreg.GRID[i] for i in idx], [ref[i] for i in idx])
for two in (0, 1):
    plan.set_two_stage(two)
    keys = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    gmin = torch.full((1,), 1e-8, dtype=torch.float64, device=dev)
    r = plan.regress(kd(), gad, keys, gmin, mode=0)
    print("two_stage", two, "gamma", r["gamma"].item(), "noise", r["noise"].item(), "status", r["status"].item())
info = {}
print("oracle", oreg.find_gamma(lam, P, 1e-8, info), info)
print("---- GPU gram, several gamma_min")
from oracle import bfgs as obfgs
desc = nat.make_desc(2, spec.scale, alpha=spec.alpha, quad_scale=spec.quad_scale, l=spec.l)
Xd = torch.from_numpy(np.ascontiguousarray(Xn)).to(dev)
wd = torch.from_numpy(w.astype(np.uint8)[None]).to(dev)
for gm in (1e-9, 1e-8):
    for two in (0, 1):
        plan.set_two_stage(two)
        Kg = plan.gram(desc, Xd, wd)
        gmin = torch.full((1,), gm, dtype=torch.float64, device=dev)
        r = plan.regress(Kg, gad, keys, gmin, mode=0)
        print("gmin", gm, "two_stage", two, "gamma", r["gamma"].item(), "noise", r["noise"].item(), "status", r["status"].item())
    kept = lam[~(lam < gm)]
    print("  oracle median", np.median(kept), "kept", kept.size, "find_gamma", oreg.find_gamma(lam, P, gm))
# numpy BFGS on the two tridiagonals from the median
for nm, (d, e) in (("one", (d1, e1)), ("two", (d2, e2))):
    f = lambda g: tm.loss_and_grad(d, e, g)
    for start in (0.00025547157396289523, 0.000140806):
        print(nm, "bfgs from", start, obfgs.minimize_bfgs_1d(f, start))
print("ref bfgs", obfgs.minimize_bfgs_1d(lg, 0.00025547157396289523), obfgs.minimize_bfgs_1d(lg, 0.000140806))
srt = np.sort(lam)
print("eigs around median", srt[(srt > 5e-5) & (srt < 1e-3)])
