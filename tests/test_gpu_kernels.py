"""Parity of every CUDA building block (through the C ABI) against the NumPy oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import jaxprng, kernels as okern, regression as oreg, chd as ochd  # noqa: E402
import tridiag_model as tm  # noqa: E402


def _native():
    from computationalhypergraphdiscovery_b200 import _native
    return _native


def _keys_t(keys, dev):
    return torch.from_numpy(np.asarray(keys, dtype=np.uint32).view(np.int32).copy()).to(dev)


def _data(n, N, seed=0):
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(n, N, seed)
    X = (X - X.mean(0)) / X.std(0)
    return np.ascontiguousarray(X), names


SPECS = okern.make_specs([("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)])


def _desc(nat, spec):
    kind = {"linear": 0, "quadratic": 1, "gaussian": 2}[spec.name]
    return nat.make_desc(kind, spec.scale, alpha=spec.alpha or 0.0, quad_scale=spec.quad_scale or 0.0,
                         l=spec.l or 1.0)


def test_split_and_normal_match_jax_prng():
    nat = _native()
    plan = nat.Plan(64, 4, 4)
    keys = np.array([[0, 0], [4146024105, 967050713], [123, 456789]], dtype=np.uint32)
    nk, sk = plan.split(_keys_t(keys, plan.device))
    for t in range(3):
        ref = jaxprng.split(keys[t])
        assert (nk[t].cpu().numpy().view(np.uint32) == ref[0]).all()
        assert (sk[t].cpu().numpy().view(np.uint32) == ref[1]).all()
    for rows, cols in [(1000, 4), (257, 100)]:
        out = plan.normal(_keys_t(keys, plan.device), rows, cols).cpu().numpy()
        for t in range(3):
            ref = jaxprng.normal(keys[t], (rows, cols))
            np.testing.assert_allclose(out[t], ref, rtol=0, atol=5e-15)


@pytest.mark.parametrize("n,N", [(100, 5), (333, 11), (64, 40)])
def test_gram_matches_oracle(n, N):
    nat = _native()
    X, _ = _data(n, N, 1)
    plan = nat.Plan(n, N, N)
    Xd = torch.from_numpy(X).to(plan.device)
    rng = np.random.default_rng(2)
    w = (rng.random((3, N)) < 0.7).astype(np.uint8)
    w[0] = 1
    w[2] = 0
    w[2, 0] = 1
    for spec in SPECS:
        K = plan.gram(_desc(nat, spec), Xd, torch.from_numpy(w).to(plan.device)).cpu().numpy()[:, :, :n]
        for t in range(3):
            ref = okern.gram(spec, X, X, w[t].astype(float))
            err = np.abs(K[t] - ref).max() / np.abs(ref).max()
            assert err < 1e-14, (spec, t, err)
            assert np.array_equal(K[t], K[t].T)


def test_cross_gram_and_predict():
    nat = _native()
    n, N, nq = 150, 6, 77
    X, _ = _data(n, N, 3)
    Xq = np.ascontiguousarray(_data(nq, N, 4)[0])
    plan = nat.Plan(n, N, N)
    dev = plan.device
    w = np.ones((2, N), dtype=np.uint8)
    w[1, 2] = 0
    yb = np.random.default_rng(5).standard_normal((2, n))
    for spec in SPECS:
        out = plan.predict(_desc(nat, spec), torch.from_numpy(X).to(dev), torch.from_numpy(Xq).to(dev),
                           torch.from_numpy(w).to(dev), torch.from_numpy(yb).to(dev)).cpu().numpy()
        for t in range(2):
            ref = okern.gram(spec, Xq, X, w[t].astype(float)) @ (-yb[t])
            np.testing.assert_allclose(out[t], ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())


@pytest.mark.parametrize("n,G,threads,bulk", [(97, 2, 0, 0), (200, 4, 0, 0), (310, 8, 0, 0), (256, 3, 0, 0),
                                               (333, 5, 256, 1), (290, 4, 512, 0), (301, 6, 512, 1), (1100, 0, 0, 0),
                                               (1700, 0, 0, 0), (2300, 0, 0, 1)])
def test_tridiag_matches_householder_model(n, G, threads, bulk, monkeypatch):
    """(threads, bulk) select the kernel variant: 256 threads x 2 CTAs/SM or 512 x 1, register or
    cp.async.bulk (UBLKCP) symv data path; 0 = the plan's default."""
    if threads:
        monkeypatch.setenv("CHD_TD_THREADS", str(threads))
    monkeypatch.setenv("CHD_TD_BULK", str(bulk))
    monkeypatch.setenv("CHD_TD_LEVELS", "1" if n >= 1536 else "0")   # also exercise the multi-level restart
    nat = _native()
    N = 6
    X, _ = _data(n, N, 6)
    plan = nat.Plan(n, N, N, ctas_per_matrix=G)
    dev = plan.device
    w = np.ones(N)
    w[0] = 0
    T = 3
    Ks, gas = [], []
    for t, spec in enumerate(SPECS):
        Ks.append(okern.gram(spec, X, X, w))
        gas.append(X[:, t])
    Kd = torch.zeros((T, n, plan.ldk), dtype=torch.float64, device=dev)
    Kd[:, :, :n] = torch.from_numpy(np.stack(Ks)).to(dev)
    S = np.random.default_rng(7).standard_normal((T, n, 100))
    Bd = torch.from_numpy(S).to(dev).contiguous()
    d, e, b0 = plan.tridiag(Kd, torch.from_numpy(np.stack(gas)).to(dev).contiguous(), Bd)
    d, e, b0, B = d.cpu().numpy(), e.cpu().numpy(), b0.cpu().numpy(), Bd.cpu().numpy()
    for t in range(T):
        md, me, mb0, QT, _Q = tm.householder_tridiag(Ks[t], gas[t])
        scale = np.abs(Ks[t]).max() * n
        assert abs(b0[t] - mb0) < 1e-12 * abs(mb0)
        lam = np.linalg.eigvalsh(Ks[t])
        lamT = np.linalg.eigvalsh(np.diag(d[t]) + np.diag(e[t, :n - 1], 1) + np.diag(e[t, :n - 1], -1))
        assert np.abs(lam - lamT).max() < 1e-13 * scale, (t, np.abs(lam - lamT).max())
        # spectral functionals that the regression consumes: e0^T (T+g)^-1 e0 and B^T (T+g)^-1 B
        for g in (1e-3, 0.5, 50.0):
            x_dev = tm.tridiag_solve(d[t], e[t, :n - 1], g, np.eye(n)[:, 0])
            ref = gas[t] @ np.linalg.solve(Ks[t] + g * np.eye(n), gas[t])
            assert abs(b0[t] ** 2 * x_dev[0] - ref) < 1e-9 * abs(ref), (t, g)
            U = tm.tridiag_solve(d[t], e[t, :n - 1], g, B[t])
            ref2 = np.einsum("ij,ij->j", S[t], np.linalg.solve(Ks[t] + g * np.eye(n), S[t]))
            np.testing.assert_allclose(np.einsum("ij,ij->j", B[t], U), ref2, rtol=1e-8)


@pytest.mark.parametrize("n,N,G", [(180, 6, 0), (420, 6, 4), (257, 9, 3)])
def test_regress_matches_oracle(n, N, G):
    nat = _native()
    X, _ = _data(n, N, 8)
    plan = nat.Plan(n, N, N, ctas_per_matrix=G)
    dev = plan.device
    Xd = torch.from_numpy(X).to(dev)
    targets = [0, 1, 2]
    w = np.ones((3, N), dtype=np.uint8)
    for t, idx in enumerate(targets):
        w[t, idx] = 0
    keys = jaxprng.split(jaxprng.PRNGKey(11), 4)[1:]
    ga = np.ascontiguousarray(X[:, targets].T)
    for spec in SPECS:
        K = plan.gram(_desc(nat, spec), Xd, torch.from_numpy(w).to(dev))
        gmin = torch.full((3,), 1e-9, dtype=torch.float64, device=dev)
        r = plan.regress(K, torch.from_numpy(ga).to(dev), _keys_t(keys, dev), gmin, mode=1, base=1e-9)
        torch.cuda.synchronize()
        for t in range(3):
            Kref = okern.gram(spec, X, X, w[t].astype(float))
            info = {}
            yb, noise, zl, zh, gamma, nkey = oreg.perform_regression_and_find_gamma(Kref, ga[t], 1e-9, keys[t], info)
            got = {k: r[k][t].item() for k in ("noise", "z_low", "z_high", "gamma", "lambda_min")}
            assert (r["keys"][t].cpu().numpy().view(np.uint32) == nkey).all()
            assert abs(got["gamma"] - gamma) <= 1e-6 * abs(gamma), (spec, t, got["gamma"], gamma, info["best"])
            assert abs(got["noise"] - noise) <= 1e-8 * abs(noise), (spec, t, got["noise"], noise)
            assert abs(got["z_low"] - zl) <= 1e-8 * abs(zl), (spec, t)
            assert abs(got["z_high"] - zh) <= 1e-8 * abs(zh), (spec, t)
            ybd = r["yb"][t].cpu().numpy()
            assert np.linalg.norm(ybd - yb) <= 1e-7 * np.linalg.norm(yb), (spec, t)
            assert abs(got["lambda_min"] - info["eigenvalues"][0]) < 1e-11 * abs(info["eigenvalues"][-1])


@pytest.mark.parametrize("clusters", [False, True])
def test_activations_match_oracle(clusters):
    nat = _native()
    n, N = 150, 8
    X, names = _data(n, N, 9)
    cl = [[names[0], names[3]], [names[1]], [names[2], names[4], names[5]], [names[6]], [names[7]]] if clusters else None
    im = ochd.make_index_matrix(names, cl)
    C = im.shape[0]
    cluster_of = np.argmax(im, axis=0).astype(np.int32)
    plan = nat.Plan(n, N, C)
    dev = plan.device
    rng = np.random.default_rng(10)
    T = 3
    w0 = np.ones((T, N), dtype=np.uint8)
    alive = np.ones((T, C), dtype=np.uint8)
    for t in range(T):
        w0[t, t] = 0
    alive[1, 1] = 0
    alive[2, C - 1] = 0
    yb = rng.standard_normal((T, n)) * 0.1
    ga = np.ascontiguousarray(X[:, :T].T)
    yb = yb - 0.5 * ga  # keeps energy = -ga.yb positive
    for spec in SPECS:
        alive_d = torch.from_numpy(alive.copy()).to(dev)
        act, pruned = plan.activations(_desc(nat, spec), torch.from_numpy(X).to(dev),
                                       torch.from_numpy(cluster_of).to(dev), torch.from_numpy(w0).to(dev), alive_d,
                                       torch.from_numpy(yb).to(dev), torch.from_numpy(ga).to(dev), prune=True)
        act = act.cpu().numpy()
        for t in range(T):
            am = im * w0[t][None, :] * alive[t][:, None]
            ref = ochd.activations(spec, X, yb[t], ga[t], am, clusters)
            np.testing.assert_allclose(act[t], ref, rtol=1e-9, atol=1e-12)
            j = int(np.argmin(ref))
            assert int(pruned[t]) == j
            assert alive_d[t, j].item() == 0




@pytest.mark.parametrize("clusters,N,T,shape", [(False, 40, 37, ""), (False, 40, 37, "2"), (False, 40, 70, ""),
                                                (False, 260, 33, ""), (True, 8, 8, ""), (False, 8, 3, "")])
def test_shared_gaussian_activations_match_oracle(clusters, N, T, shape, monkeypatch):
    """gauss_shared_kernel (one evaluation of the per-pair factors for the whole batch + DMMA contraction; needs the
    cached product maps, so it is reached through chd_prune_chain) against the oracle's activations
    (helper_functions.py:113-132, kernels.py:270-278): more than one target group (T = 37 > 32), clustered variables
    (targets inside a multi-variable cluster are flagged and served by gauss_pairs_kernel), few targets.  Batches of more
    than 32 targets take the 64-target shape of the kernel (T = 37: one partly filled group, T = 70: two groups,
    N = 260: two passes over the clusters); shape "2" forces the 32-target shape (two groups at T = 37)."""
    if shape:
        monkeypatch.setenv("CHD_GAUSS_SHARED", shape)
    nat = _native()
    n = 150
    X, names = _data(n, N, 13)
    if clusters:
        cl = [[names[0], names[3]], [names[1]], [names[2], names[4], names[5]], [names[6]], [names[7]]]
    else:
        cl = None
    im = ochd.make_index_matrix(names, cl)
    C = im.shape[0]
    cluster_of = np.argmax(im, axis=0).astype(np.int32)
    plan = nat.Plan(n, N, C)
    dev = plan.device
    spec = SPECS[2]
    desc = _desc(nat, spec)
    rng = np.random.default_rng(14)
    w0 = np.ones((T, N), dtype=np.uint8)
    for t in range(T):
        w0[t, t % N] = 0
    w0[0, (N // 2) % N] = 0                      # an impossible edge
    alive = np.ones((T, C), dtype=np.uint8)
    alive[1 % T, 1] = 0
    ga = np.ascontiguousarray(X[:, np.arange(T) % N].T)
    yb = rng.standard_normal((T, n)) * 0.1 - 0.5 * ga
    f64 = torch.float64
    alive_d = torch.from_numpy(alive.copy()).to(dev)
    pc = dict(P=torch.empty((T, n, plan.ldk), dtype=f64, device=dev), valid=0)
    out = plan.prune_chain(desc, torch.from_numpy(X).to(dev), torch.from_numpy(cluster_of).to(dev),
                           torch.from_numpy(w0).to(dev), alive_d, torch.from_numpy(ga).to(dev),
                           torch.from_numpy(yb.copy()).to(dev), torch.zeros((T, 2), dtype=torch.int32, device=dev),
                           torch.zeros(T, dtype=f64, device=dev), torch.full((T,), 1e-9, dtype=f64, device=dev), 1, 1,
                           keep_yb=False, p_cache=pc)
    torch.cuda.synchronize()
    act = out["act"][:, 0].cpu().numpy()
    pruned = out["pruned"][:, 0].cpu().numpy()
    for t in (range(T) if N <= 64 else (0, T // 2, T - 1)):     # the oracle costs C full n x n x N products per target
        am = im * w0[t][None, :] * alive[t][:, None]
        ref = ochd.activations(spec, X, yb[t], ga[t], am, clusters)
        np.testing.assert_allclose(act[t], ref, rtol=1e-9, atol=1e-12)
        assert int(pruned[t]) == int(np.argmin(ref))


@pytest.mark.parametrize("n,N", [(333, 11), (130, 70)])
def test_gram_downdate_from_shared_core_matches_oracle(n, N):
    """Batches of >= 4 targets take the downdate form  L = X X^T - X_r X_r^T  for the targets with fewer removed than
    active variables (gram.cu); the others keep the direct form.  Same parity bar as the direct Gram, plus bitwise
    symmetry."""
    nat = _native()
    X, _ = _data(n, N, 3)
    plan = nat.Plan(n, N, N)
    Xd = torch.from_numpy(X).to(plan.device)
    rng = np.random.default_rng(4)
    T = 7
    w = np.ones((T, N), dtype=np.uint8)
    w[0, 0] = 0                                      # stage-1 like: one removed
    w[1, rng.choice(N, size=N // 4, replace=False)] = 0
    w[2] = (rng.random(N) < 0.5)
    w[3] = (rng.random(N) < 0.2)                     # mostly removed: direct form
    w[4] = 0
    w[4, 1] = 1
    w[5] = 1                                         # nothing removed
    w[6, N - 1] = 0
    for spec in SPECS:
        K = plan.gram(_desc(nat, spec), Xd, torch.from_numpy(w).to(plan.device)).cpu().numpy()[:, :, :n]
        for t in range(T):
            ref = okern.gram(spec, X, X, w[t].astype(float))
            err = np.abs(K[t] - ref).max() / np.abs(ref).max()
            assert err < 5e-14, (spec, t, err)
            assert np.array_equal(K[t], K[t].T)
