"""Parity in the regimes the bench runs in:

* chains longer than the 50-step refresh cadence -- the `step % 50 == 0` gamma_min refresh of
  _GraphDiscoveryMain.py:622-631 and, for the Gaussian mode, the incremental product map
  P <- P / (1 + E_removed) with its 50-step rebuild -- against `oracle.chd.fit`, for each of the three kernels
  (a user KernelChooser forces the kernel so that every mode gets a 62-step chain);
* n = 8192 (BASELINE.json configs[3] sample size): Gram + regression of one target per kernel against the oracle's
  LAPACK eigendecomposition (interpolatory.py:24-135), then one prune step's activations / argmin
  (helper_functions.py:166-176).

Both report the worst observed deviation (printed; run with -s to see it) next to the 1e-8 contract.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import chd as ochd  # noqa: E402
from oracle import kernels as okern, regression as oreg, jaxprng  # noqa: E402
from test_gpu_fit import _compare_with_oracle, _force_chooser, flip_report  # noqa: E402

KERNELS = [("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)]


@pytest.mark.parametrize("k,n,ntargets", [(0, 200, 6), (1, 200, 6), (2, 150, 3)])
def test_62_step_chain_matches_oracle(k, n, ntargets):
    """N = C = 64 -> L = 62 prune steps: crosses the step-50 refresh; the Gaussian P-cache is divided 49 times
    between its rebuilds."""
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(n, 64, seed=11)
    targets = names[:ntargets]
    gd = CHD.GraphDiscovery(X, names, verbose=False)
    gd.fit(targets=targets, kernel_chooser=_force_chooser(k))
    res = ochd.fit(X, names, kernels=KERNELS, targets=targets, kernel_chooser=("force", k))
    ch, mine = res["chains"][k], gd.last_fit["chains"][k]
    assert ch["noises"].shape == (ntargets, 63) and mine["noises"].shape == (ntargets, 63)
    _compare_with_oracle(gd, res, targets, {})
    worst = {q: float(np.nanmax(np.abs(mine[q] - ch[q]) / np.abs(ch[q]))) for q in ("noises", "Z_lows", "gammas")}
    m = ch["activations"] != 2.0
    worst["activations"] = float(np.max(np.abs(mine["activations"][m] - ch["activations"][m])
                                        / np.abs(ch["activations"][m])))
    print(f"\n62-step chain, kernel {KERNELS[k][0]}: worst relative deviation from the oracle {worst}; "
          f"{flip_report()}")


@pytest.fixture(scope="module")
def full():
    from computationalhypergraphdiscovery_b200 import _native
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    out = {}
    for N in (64, 12):
        X, _ = make_data(8192, N, seed=0)
        X = np.ascontiguousarray((X - X.mean(0)) / X.std(0))
        plan = _native.Plan(8192, N, N)
        out[N] = (plan, X, torch.from_numpy(X).to(plan.device))
    return _native, out


@pytest.mark.parametrize("k", [0, 1, 2])
def test_regression_and_prune_step_at_n8192_match_oracle(full, k):
    """One target at n = 8192: gamma, noise, Z bounds, yb, lambda_min of the CUDA path (16-CTA cluster panel QR,
    split-K symm, bulge chasing -- the code paths the bench runs) against one LAPACK eigh; then the activations and
    the argmin of the following prune step.  The Gaussian mode uses N = 12 variables (its oracle activations cost
    C full n x n products each); linear / quadratic use N = 64."""
    nat, plans = full
    n = 8192
    N = 12 if k == 2 else 64
    plan, X, Xd = plans[N]
    dev = plan.device
    spec = okern.make_specs(KERNELS)[k]
    desc = [nat.make_desc(0, 2.0), nat.make_desc(1, 1.0, alpha=1.0),
            nat.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)][k]
    tgt = 3
    w = np.ones((1, N), dtype=np.uint8)
    w[0, tgt] = 0
    key = jaxprng.split(jaxprng.PRNGKey(5), 2)[1:]
    ga = np.ascontiguousarray(X[:, [tgt]].T)
    w_d = torch.from_numpy(w).to(dev)
    ga_d = torch.from_numpy(ga).to(dev)
    keys_d = torch.from_numpy(key.astype(np.uint32).view(np.int32).copy()).to(dev)
    K = plan.gram(desc, Xd, w_d)
    Kref = okern.gram(spec, X, X, w[0].astype(float))
    Kh = K[0, :, :n].cpu().numpy()
    assert np.max(np.abs(Kh - Kref)) <= 1e-13 * np.max(np.abs(Kref))
    gmin = torch.full((1,), 1e-9, dtype=torch.float64, device=dev)
    r = plan.regress(K, ga_d, keys_d, gmin, mode=1, base=1e-9)
    torch.cuda.synchronize()
    info = {}
    yb, noise, zl, zh, gamma, nkey = oreg.perform_regression_and_find_gamma(Kref, ga[0], 1e-9, key[0], info)
    got = {q: r[q][0].item() for q in ("noise", "z_low", "z_high", "gamma", "lambda_min")}
    dev_rel = dict(gamma=abs(got["gamma"] - gamma) / abs(gamma), noise=abs(got["noise"] - noise) / abs(noise),
                   z_low=abs(got["z_low"] - zl) / abs(zl), z_high=abs(got["z_high"] - zh) / abs(zh),
                   yb=float(np.linalg.norm(r["yb"][0].cpu().numpy() - yb) / np.linalg.norm(yb)),
                   lambda_min=abs(got["lambda_min"] - info["eigenvalues"][0]) / abs(info["eigenvalues"][-1]))
    print(f"\nn=8192 {spec.name}: gamma={gamma:.6g} noise={noise:.6g} relative deviations {dev_rel}")
    assert (r["keys"][0].cpu().numpy().view(np.uint32) == nkey).all()
    assert dev_rel["gamma"] <= 1e-6, (got["gamma"], gamma, info["best"])
    assert dev_rel["noise"] <= 1e-8 and dev_rel["z_low"] <= 1e-8 and dev_rel["z_high"] <= 1e-8
    assert dev_rel["yb"] <= 1e-7
    assert dev_rel["lambda_min"] <= 1e-11
    # ---- one prune step: activations of the full set, argmin, removed row (HELP:166-176)
    im = np.eye(N)
    am = im.copy()
    am[:, tgt] = 0
    act_ref = ochd.activations(spec, X, yb, ga[0], am, False)
    cl = torch.arange(N, dtype=torch.int32, device=dev)
    alive = torch.ones((1, N), dtype=torch.uint8, device=dev)
    # same inputs on both sides: the oracle's yb (the CUDA yb was compared above)
    yb_d = torch.from_numpy(np.ascontiguousarray(yb[None, :])).to(dev)
    act, pruned = plan.activations(desc, Xd, cl, w_d, alive, yb_d, ga_d, True)
    torch.cuda.synchronize()
    act = act[0].cpu().numpy()
    m = act_ref != 2.0
    assert (act[~m] == 2.0).all()
    strict = float(np.max(np.abs(act[m] - act_ref[m]) / np.abs(act_ref[m])))
    # activations far below the largest one are differences of O(max) terms: floor the denominator at 1e-4 max
    floor = 1e-4 * float(np.max(np.abs(act_ref[m])))
    worst = float(np.max(np.abs(act[m] - act_ref[m]) / np.maximum(np.abs(act_ref[m]), floor)))
    print(f"n=8192 {spec.name}: worst relative activation deviation {worst:.3g} (strict, per element: {strict:.3g})")
    assert worst <= 1e-8
    assert int(pruned[0].item()) == int(np.argmin(act_ref))
    assert int(alive[0, int(np.argmin(act_ref))].item()) == 0
