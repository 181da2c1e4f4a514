"""CPU proof that the two-stage formulation (tests/twostage_model.py: dense -> band by block reflectors, band ->
tridiagonal by bulge chasing for the loss, banded Cholesky solves for everything that needs vectors) equals the
eigendecomposition formulation of the reference (interpolatory.py:24-135, restated in oracle/regression.py)."""
import numpy as np
import pytest

from oracle import kernels, regression, jaxprng
import tridiag_model as tm
import twostage_model as ts


@pytest.fixture(scope="module")
def problem():
    rng = np.random.default_rng(1)
    n, N = 150, 6
    X = rng.standard_normal((n, N))
    X[:, 0] = X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(n)
    X = (X - X.mean(0)) / X.std(0)
    specs = kernels.make_specs([("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)])
    w = np.ones(N)
    w[0] = 0
    return X, X[:, 0].copy(), specs, w


@pytest.mark.parametrize("b", [4, 8, 32])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_twostage_equals_eigen_formulation(problem, k, b):
    X, ga, specs, w = problem
    n = X.shape[0]
    K = kernels.gram(specs[k], X, X, w)
    lam, V = np.linalg.eigh(K)
    P = V.T @ ga
    lg = regression.make_loss(lam, P)
    key = jaxprng.PRNGKey(3)
    S = jaxprng.normal(key, (n, 100))
    s1 = ts.stage1(K, ga, b, S)
    scale = abs(lam).max()
    assert s1["resid"] < 1e-12 * scale
    assert abs(s1["beta0"] ** 2 - ga @ ga) < 1e-12 * (ga @ ga)
    Tb = ts.band_to_dense(s1["band"])
    assert np.abs(np.linalg.eigvalsh(Tb) - lam).max() < 1e-13 * scale * n
    d, e = ts.stage2(s1["band"])
    Tt = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.abs(np.linalg.eigvalsh(Tt) - lam).max() < 1e-13 * scale * n
    for g in (0.1, 5.0, 2e3):
        f, fp = lg(g)
        f2, fp2 = tm.loss_and_grad(d, e, g)
        assert abs(f - f2) < 1e-10 * abs(f)
        assert abs(fp - fp2) < 1e-7 * max(abs(fp), 1e-6)
    g = 0.05
    yb, noise = regression.solve_variationnal(ga, g, lam, V)
    L = ts.band_cholesky(s1["band"], g)
    e0 = np.zeros(n)
    e0[0] = 1.0
    _, x = ts.band_solve(L, e0)
    yb2 = -s1["beta0"] * ts.apply_Q(s1, x)
    assert np.linalg.norm(yb - yb2) < 1e-9 * np.linalg.norm(yb)
    assert abs(g * (x @ x) / x[0] - noise) < 1e-9 * noise
    zl, zh = regression.Z_test(g, lam, V, key)
    Y, Z = ts.band_solve(L, s1["Bq"])
    nz = np.sort(g * np.sum(Z * Z, 0) / np.sum(Y * Y, 0))
    assert abs(nz[5] - zl) < 1e-9 * zl and abs(nz[95] - zh) < 1e-9 * zh


@pytest.mark.parametrize("n,b", [(33, 32), (34, 32), (65, 32), (70, 8), (9, 4), (5, 4), (3, 4)])
def test_twostage_ragged_sizes(n, b):
    rng = np.random.default_rng(n)
    F = rng.standard_normal((n, n + 3))
    K = F @ F.T / n
    ga = rng.standard_normal(n)
    lam = np.linalg.eigvalsh(K)
    s1 = ts.stage1(K, ga, b)
    assert s1["resid"] < 1e-12 * lam.max()
    d, e = ts.stage2(s1["band"])
    Tt = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.abs(np.linalg.eigvalsh(Tt) - lam).max() < 1e-12 * lam.max()
    g = 0.3
    F1 = ga @ np.linalg.solve(K + g * np.eye(n), ga)
    eta, _, _ = tm.eta_jets(d, np.r_[e, 0.0], g)
    assert abs(s1["beta0"] ** 2 / eta - F1) < 1e-11 * abs(F1)


@pytest.mark.parametrize("m,b", [(200, 32), (40, 32), (33, 32), (20, 32), (2, 32), (64, 8)])
def test_one_exchange_panel_qr_equals_householder_qr(m, b):
    """The single-reduction-per-column formulation of the CUDA panel QR equals the textbook Householder QR."""
    rng = np.random.default_rng(m)
    P = rng.standard_normal((m, b))
    R1, V1, tau1, T1 = ts.panel_qr(P)
    R2, V2, tau2, T2 = ts.panel_qr_one_exchange(P)
    np.testing.assert_allclose(tau2, tau1, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(V2, V1, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(R2, R1, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(T2, T1, rtol=1e-10, atol=1e-12)
    Q = np.eye(m) - V2 @ T2 @ V2.T
    np.testing.assert_allclose(Q.T @ Q, np.eye(m), atol=1e-12)
    np.testing.assert_allclose(Q.T @ P, R2, atol=1e-11)


@pytest.mark.parametrize("n,b", [(60, 8), (40, 32), (33, 32), (5, 4)])
def test_band_ldl_equals_dense_solve(n, b):
    """The square-root-free banded factorisation of bandsolve.cu, incl. an indefinite shift (negative pivots)."""
    rng = np.random.default_rng(n)
    band = np.zeros((n, b + 1))
    for d in range(min(b, n - 1) + 1):
        band[d:, d] = rng.standard_normal(n - d) * (2.0 if d == 0 else 0.3)
    A = ts.band_to_dense(band)
    R = rng.standard_normal((n, 7))
    for g in (15.0, -0.37):
        Lg = ts.band_ldl(band, g)
        y, z = ts.band_ldl_solve(Lg, R)
        ref = np.linalg.solve(A + g * np.eye(n), R)
        np.testing.assert_allclose(z, ref, rtol=1e-8, atol=1e-10 * np.abs(ref).max())
        np.testing.assert_allclose(np.sum(y * y / Lg[:, :1], 0), np.sum(R * ref, 0), rtol=1e-8)
        if g > 0:
            assert (Lg[:, 0] > 0).all()


@pytest.mark.parametrize("n,b,seed", [(40, 4, 0), (40, 4, 1), (70, 8, 2), (37, 8, 3), (30, 32, 4), (90, 32, 5)])
def test_chase_pipeline_rule_is_sufficient(n, b, seed):
    """The hop tasks of chase.cu, on the compact band storage, executed in a random order that respects only the
    kernels' dependency rule (hop h of sweep j after hop h+1 of sweep j-1) give the sequential result."""
    rng = np.random.default_rng(seed)
    band = np.zeros((n, b + 1))
    for d in range(min(b, n - 1) + 1):
        band[d:, d] = rng.standard_normal(n - d) * (3.0 if d == 0 else 1.0)
    d_seq, e_seq = ts.stage2(band)
    d_par, e_par = ts.ChaseTasks(band).run(rng)
    np.testing.assert_allclose(d_par, d_seq, rtol=0, atol=1e-11 * np.abs(d_seq).max())
    np.testing.assert_allclose(np.abs(e_par), np.abs(e_seq), rtol=0, atol=1e-11 * np.abs(d_seq).max())


@pytest.mark.parametrize("n,b", [(70, 8), (64, 8), (37, 4), (9, 4), (131, 32), (200, 32)])
@pytest.mark.parametrize("kd", [1, 2, 3])
def test_deferred_trailing_update_equals_per_panel_update(n, b, kd):
    """The grouped (deferred) trailing update of csrc/band.cu -- stale matrix + thin block-column update + Y / M
    corrections + one rank-2*kd*b update per group -- gives the same band form, reflectors and Q^T S as the
    per-panel update."""
    rng = np.random.default_rng(n + b)
    Xr = rng.standard_normal((n, n))
    K = Xr @ Xr.T / n + np.eye(n)
    ga = rng.standard_normal(n)
    S = rng.standard_normal((n, 5))
    r1 = ts.stage1(K, ga, b, S)
    r2 = ts.stage1_deferred(K, ga, b, S, kd)
    scale = np.abs(r1["band"]).max()
    assert np.abs(r1["band"] - r2["band"]).max() <= 1e-12 * scale
    assert np.abs(r1["Bq"] - r2["Bq"]).max() <= 1e-11 * np.abs(S).max()
    assert len(r1["panels"]) == len(r2["panels"])
    for (ra, Va, Ta), (rb, Vb, Tb) in zip(r1["panels"], r2["panels"]):
        assert ra == rb and np.abs(Va - Vb).max() <= 1e-10 and np.abs(Ta - Tb).max() <= 1e-10
