"""CPU proof that the narrow-band formulation (tests/band_model.py) equals the eigendecomposition formulation of
the reference (interpolatory.py:24-135, restated in oracle/regression.py) for half-widths B = 1, 2, 4."""
import numpy as np
import pytest

from oracle import kernels, regression, jaxprng
import band_model as bm


@pytest.fixture(scope="module")
def problem():
    rng = np.random.default_rng(1)
    n, N = 90, 6
    X = rng.standard_normal((n, N))
    X[:, 0] = X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(n)
    X = (X - X.mean(0)) / X.std(0)
    specs = kernels.make_specs([("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)])
    w = np.ones(N)
    w[0] = 0
    return X, X[:, 0].copy(), specs, w


@pytest.mark.parametrize("B", [1, 2, 4])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_band_formulation_equals_eigen_formulation(problem, k, B):
    X, ga, specs, w = problem
    n = X.shape[0]
    K = kernels.gram(specs[k], X, X, w)
    lam, V = np.linalg.eigh(K)
    P = V.T @ ga
    lg = regression.make_loss(lam, P)
    band, gt, refl, resid = bm.band_reduce(K, ga, B)
    assert resid < 1e-12 * abs(lam).max()
    assert abs(gt @ gt - ga @ ga) < 1e-12 * (ga @ ga)
    Tb = sum(np.diag(band[t, t:], -t) + (np.diag(band[t, t:], t) if t else 0) for t in range(B + 1))
    assert np.abs(np.linalg.eigvalsh(Tb) - lam).max() < 1e-13 * abs(lam).max() * n
    for g in (0.1, 5.0, 2e3):
        f, fp = lg(g)
        f2, fp2 = bm.loss_and_grad(band, gt, g)
        assert abs(f - f2) < 1e-10 * abs(f)
        assert abs(fp - fp2) < 1e-7 * max(abs(fp), 1e-6)
    g = 0.05
    yb, noise = regression.solve_variationnal(ga, g, lam, V)
    rhs = np.zeros(n)
    rhs[:B] = gt
    x = bm.band_solve(band, g, rhs)
    yb2 = -bm.apply_Q(refl, x)
    assert np.linalg.norm(yb - yb2) < 1e-9 * np.linalg.norm(yb)
    assert abs(g * (x @ x) / (rhs @ x) - noise) < 1e-9 * noise
    key = jaxprng.PRNGKey(3)
    zl, zh = regression.Z_test(g, lam, V, key)
    Bq = bm.apply_QT(refl, jaxprng.normal(key, (n, 100)))
    U = bm.band_solve(band, g, Bq)
    nz = np.sort(g * np.sum(U * U, 0) / np.sum(Bq * U, 0))
    assert abs(nz[5] - zl) < 1e-9 * zl and abs(nz[95] - zh) < 1e-9 * zh
    kept = lam[~(lam < 1e-9)]
    c0 = bm.inertia_below(band, 1e-9)
    m = n - c0
    assert m == kept.size
    med = bm.kth_eigenvalue(band, c0 + m // 2) if m % 2 else 0.5 * (
        bm.kth_eigenvalue(band, c0 + m // 2 - 1) + bm.kth_eigenvalue(band, c0 + m // 2))
    assert abs(med - np.median(kept)) < 1e-10 * abs(lam).max()
