"""Parity of the two-stage reduction kernels (band.cu, chase.cu, bandsolve.cu; through the C ABI) against the NumPy
model tests/twostage_model.py and the oracle (interpolatory.py:24-135 restated in oracle/regression.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import jaxprng, kernels as okern, regression as oreg  # noqa: E402
import tridiag_model as tm  # noqa: E402
import twostage_model as ts  # noqa: E402


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


def _problem(n, seed=6):
    N = 6
    X, _ = _data(n, N, seed)
    w = np.ones(N)
    w[0] = 0
    Ks = [okern.gram(spec, X, X, w) for spec in SPECS]
    gas = [X[:, t].copy() for t in range(3)]
    return Ks, gas


@pytest.mark.parametrize("n,qr", [(3, ""), (20, ""), (33, ""), (34, ""), (70, ""), (97, ""), (200, ""), (333, ""),
                                  (1100, ""), (2300, "fallback"), (1000, "two1"), (1100, "two1"), (2300, "two2"),
                                  (2300, "two_auto"), (333, "fma"), (1100, "fma")])
def test_band_reduce_backtransform_match_model(n, qr, monkeypatch):
    """Matrices 0-2 are the (numerically rank-deficient) mode Gram matrices: Q is not unique there, so only the
    quantities the regression consumes are compared; matrix 3 is well conditioned and is compared element-wise
    with the NumPy model (same algorithm, same reflector conventions).  `qr` selects the panel-QR kernel the large
    panels go through: the register kernel (default), its two-rows-per-thread variant (the n > 8192 path, reached
    here by capping the cluster size or by CHD_QR_TWO=2) or the shared-memory / in-place fallback; "fma" keeps the
    FMA forms of the per-panel small products (default: DMMA, wz_dmma_kernel / partials_dmma_kernel)."""
    if qr == "fallback":
        monkeypatch.setenv("CHD_QR_REG", "0")
        monkeypatch.setenv("CHD_QR_MAXCS", "2")     # forces panel rows out of shared memory (in-place rows in L2)
    elif qr == "two_auto":
        monkeypatch.setenv("CHD_QR_TWO", "2")            # two rows per thread whenever the cluster can be halved
    elif qr.startswith("two"):
        monkeypatch.setenv("CHD_QR_REG_MAXCS", qr[3:])   # 512 < rows per CTA <= 1024: second row in shared memory
    elif qr == "fma":
        monkeypatch.setenv("CHD_SMALL_DMMA", "0")        # FMA forms of the per-panel small products
    nat = _native()
    Ks, gas = _problem(n)
    rng = np.random.default_rng(5)
    F = rng.standard_normal((n, n))
    Ks.append(F @ F.T / n + np.eye(n))
    gas.append(rng.standard_normal(n))
    T = 4
    plan = nat.Plan(n, 6, 6)
    dev = plan.device
    Kd = torch.zeros((T, n, plan.ldk), dtype=torch.float64, device=dev)
    Kd[:, :, :n] = torch.from_numpy(np.stack(Ks)).to(dev)
    if plan.ldk > n:
        Kd[:, :, n:] = float("nan")                  # the padding column must not leak into the result
    S = rng.standard_normal((T, n, 100))
    Bd = torch.from_numpy(S).to(dev).contiguous()
    band, b0 = plan.band_reduce(Kd, torch.from_numpy(np.stack(gas)).to(dev).contiguous(), Bd)
    torch.cuda.synchronize()
    band, b0, Bq = band.cpu().numpy(), b0.cpu().numpy(), Bd.cpu().numpy()
    g = 0.5
    e0 = np.eye(n)[:, 0]
    xb = np.stack([np.linalg.solve(ts.band_to_dense(band[t, :, :nat.BAND + 1]) + g * np.eye(n), e0) for t in range(T)])
    out = plan.band_backtransform(Kd, torch.from_numpy(xb).to(dev), torch.from_numpy(b0).to(dev)).cpu().numpy()
    for t in range(T):
        scale = np.abs(Ks[t]).max() * n
        got = band[t, :, :nat.BAND + 1]
        Tb = ts.band_to_dense(got)
        lam = np.linalg.eigvalsh(Ks[t])
        assert np.abs(lam - np.linalg.eigvalsh(Tb)).max() < 1e-13 * scale
        assert abs(b0[t] ** 2 - gas[t] @ gas[t]) < 1e-12 * (gas[t] @ gas[t])
        Kg = Ks[t] + g * np.eye(n)
        ref = np.linalg.solve(Kg, gas[t])
        assert abs(b0[t] ** 2 * xb[t, 0] - gas[t] @ ref) < 1e-9 * abs(gas[t] @ ref), t
        assert np.linalg.norm(out[t] + ref) < 1e-9 * np.linalg.norm(ref), t      # yb = -(K+g)^-1 ga
        U = np.linalg.solve(Tb + g * np.eye(n), Bq[t])
        ref2 = np.einsum("ij,ij->j", S[t], np.linalg.solve(Kg, S[t]))
        np.testing.assert_allclose(np.einsum("ij,ij->j", Bq[t], U), ref2, rtol=1e-8)
    t = 3
    m = ts.stage1(Ks[t], gas[t], nat.BAND, S[t])
    scale = np.abs(Ks[t]).max() * n
    assert abs(b0[t] - m["beta0"]) < 1e-12 * abs(m["beta0"])
    assert np.abs(band[t, :, :nat.BAND + 1] - m["band"]).max() < 1e-10 * scale
    assert np.abs(Bq[t] - m["Bq"]).max() < 1e-9 * np.abs(S[t]).max() * n
    ref = -m["beta0"] * ts.apply_Q(m, xb[t])
    assert np.linalg.norm(out[t] - ref) < 1e-10 * np.linalg.norm(ref) * n


@pytest.mark.parametrize("n", [3, 33, 40, 97, 200, 700, 1500])
def test_band_chase_keeps_spectrum_and_first_row(n):
    nat = _native()
    rng = np.random.default_rng(n)
    b = nat.BAND
    T = 3
    bands = np.zeros((T, n, nat.BAND_LD))
    for t in range(T):
        for d in range(min(b, n - 1) + 1):
            bands[t, d:, d] = rng.standard_normal(n - d) * (3.0 if d == 0 else 1.0 / (1 + d))
    plan = nat.Plan(n, 4, 4)
    d, e = plan.band_chase(torch.from_numpy(bands).to(plan.device))
    torch.cuda.synchronize()
    d, e = d.cpu().numpy(), e.cpu().numpy()
    for t in range(T):
        A = ts.band_to_dense(bands[t, :, :b + 1])
        lam = np.linalg.eigvalsh(A)
        Tt = np.diag(d[t]) + np.diag(e[t, :n - 1], 1) + np.diag(e[t, :n - 1], -1)
        assert np.abs(np.linalg.eigvalsh(Tt) - lam).max() < 1e-13 * np.abs(lam).max() * n
        md, me = ts.stage2(bands[t, :, :b + 1])
        assert np.abs(d[t] - md).max() < 1e-10 * np.abs(lam).max() * n
        assert np.abs(np.abs(e[t, :n - 1]) - np.abs(me)).max() < 1e-10 * np.abs(lam).max() * n
        g = 2.0 + np.abs(lam).max()
        F = np.linalg.solve(A + g * np.eye(n), np.eye(n)[:, 0])[0]
        eta, _, _ = tm.eta_jets(d[t], e[t], g)
        assert abs(1.0 / eta - F) < 1e-12 * abs(F)


@pytest.mark.parametrize("n", [5, 33, 100, 333, 1000])
def test_band_solve_matches_dense_solve(n):
    nat = _native()
    rng = np.random.default_rng(n + 1)
    b = nat.BAND
    T = 2
    bands = np.zeros((T, n, nat.BAND_LD))
    for t in range(T):
        for d in range(min(b, n - 1) + 1):
            bands[t, d:, d] = rng.standard_normal(n - d) * (2.0 if d == 0 else 0.05)
    S = rng.standard_normal((T, n, 100))
    gu = np.array([40.0, 55.0])
    plan = nat.Plan(n, 4, 4)
    dev = plan.device
    Bq = torch.from_numpy(S).to(dev).contiguous()
    r = plan.band_solve(torch.from_numpy(bands).to(dev), torch.from_numpy(gu).to(dev), Bq)
    torch.cuda.synchronize()
    for t in range(T):
        A = ts.band_to_dense(bands[t, :, :b + 1]) + gu[t] * np.eye(n)
        x = np.linalg.solve(A, np.eye(n)[:, 0])
        np.testing.assert_allclose(r["x"][t].cpu().numpy(), x, rtol=1e-10, atol=1e-14 * np.abs(x).max())
        assert abs(r["noise"][t].item() - gu[t] * (x @ x) / x[0]) < 1e-11 * gu[t]
        Z = np.linalg.solve(A, S[t])
        nz = np.sort(gu[t] * np.sum(Z * Z, 0) / np.sum(S[t] * Z, 0))
        assert abs(r["z_low"][t].item() - nz[5]) < 1e-10 * nz[5]
        assert abs(r["z_high"][t].item() - nz[95]) < 1e-10 * nz[95]


@pytest.mark.parametrize("n,N", [(180, 6), (420, 6), (257, 9), (40, 5), (21, 4), (1031, 5)])
def test_regress_two_stage_matches_oracle(n, N):
    nat = _native()
    X, _ = _data(n, N, 8)
    plan = nat.Plan(n, N, N)
    plan.set_two_stage(True)
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


@pytest.mark.parametrize("tma", ["1", "0"])
@pytest.mark.parametrize("n", [70, 333, 1100])
def test_deferred_group_update_equals_per_panel_update(n, tma, monkeypatch):
    """The grouped trailing update (KD = 2, default) and the per-panel update (CHD_KD=1) are the same orthogonal
    reduction: band form, Q^T S, reflectors and beta0 agree to rounding on a well-conditioned matrix -- with the
    TMA + mbarrier kernels (band_tma.cu: trailing update and Y = A22 V with the paired schedule; default) and with the
    cp.async ones (CHD_SYR2K_TMA=0, CHD_SYMM_TMA=0)."""
    monkeypatch.setenv("CHD_SYR2K_TMA", tma)
    monkeypatch.setenv("CHD_SYMM_TMA", tma)
    nat = _native()
    rng = np.random.default_rng(9)
    F = rng.standard_normal((n, n))
    K = F @ F.T / n + np.eye(n)
    ga = rng.standard_normal(n)
    S = rng.standard_normal((1, n, 100))
    res = []
    for kd in ("1", "2"):
        monkeypatch.setenv("CHD_KD", kd)
        plan = nat.Plan(n, 6, 6)
        dev = plan.device
        Kd = torch.zeros((1, n, plan.ldk), dtype=torch.float64, device=dev)
        Kd[0, :, :n] = torch.from_numpy(K).to(dev)
        Bd = torch.from_numpy(S).to(dev).contiguous()
        band, b0 = plan.band_reduce(Kd, torch.from_numpy(ga[None, :]).to(dev).contiguous(), Bd)
        torch.cuda.synchronize()
        res.append((band.cpu().numpy()[0, :, :nat.BAND + 1], b0.cpu().numpy(), Bd.cpu().numpy(),
                    np.tril(Kd.cpu().numpy()[0, :, :n], -(nat.BAND + 1))))
    (b1, beta1, B1, V1), (b2, beta2, B2, V2) = res
    model = ts.stage1_deferred(K, ga, nat.BAND, S[0], 2)
    scale = np.abs(K).max()
    assert np.abs(b1 - b2).max() <= 1e-12 * scale * n
    assert np.abs(b2 - model["band"]).max() <= 1e-12 * scale * n
    assert abs(beta1[0] - beta2[0]) <= 1e-14 * abs(beta1[0])
    assert np.abs(B1 - B2).max() <= 1e-11 * np.abs(S).max() * np.sqrt(n)
    assert np.abs(B2[0] - model["Bq"]).max() <= 1e-11 * np.abs(S).max() * np.sqrt(n)
    assert np.abs(V1 - V2).max() <= 1e-10           # reflectors stored below the band
