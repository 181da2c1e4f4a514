"""CPU proof that the eigenvector-free tridiagonal formulation the CUDA path implements (DESIGN.md
section 3; NumPy model in tests/tridiag_model.py) equals the reference's eigendecomposition
formulation (interpolatory.py:24-135, restated in oracle/regression.py)."""
import numpy as np
import pytest

from oracle import kernels, regression, jaxprng
import tridiag_model as tm


@pytest.fixture(scope="module")
def problem():
    rng = np.random.default_rng(1)
    n, N = 160, 6
    X = rng.standard_normal((n, N))
    X[:, 0] = X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(n)
    X = (X - X.mean(0)) / X.std(0)
    specs = kernels.make_specs([("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)])
    w = np.ones(N)
    w[0] = 0
    return X, X[:, 0].copy(), specs, w


@pytest.mark.parametrize("k", [0, 1, 2])
def test_tridiagonal_formulation_equals_eigen_formulation(problem, k):
    X, ga, specs, w = problem
    n = X.shape[0]
    K = kernels.gram(specs[k], X, X, w)
    lam, V = np.linalg.eigh(K)
    P = V.T @ ga
    lg = regression.make_loss(lam, P)
    d, e, b0, QT, Q = tm.householder_tridiag(K, ga)
    assert abs(b0 * b0 - ga @ ga) < 1e-12 * (ga @ ga)
    lamT = np.linalg.eigvalsh(np.diag(d) + np.diag(e, 1) + np.diag(e, -1))
    assert np.abs(lamT - lam).max() < 1e-13 * abs(lam).max() * n
    for g in (0.1, 5.0, 2e3):
        f, fp = lg(g)
        f2, fp2 = tm.loss_and_grad(d, e, g)
        assert abs(f - f2) < 1e-10 * abs(f)
        assert abs(fp - fp2) < 1e-7 * max(abs(fp), 1e-6)
    g = 0.05
    yb, noise = regression.solve_variationnal(ga, g, lam, V)
    x = tm.tridiag_solve(d, e, g, np.eye(n)[:, 0])
    yb2 = -b0 * Q(x)
    assert np.linalg.norm(yb - yb2) < 1e-9 * np.linalg.norm(yb)
    assert abs(g * (x @ x) / x[0] - noise) < 1e-9 * noise
    key = jaxprng.PRNGKey(3)
    zl, zh = regression.Z_test(g, lam, V, key)
    B = QT(jaxprng.normal(key, (n, 100)))
    U = tm.tridiag_solve(d, e, g, B)
    nz = np.sort(g * np.sum(U * U, 0) / np.sum(B * U, 0))
    assert abs(nz[5] - zl) < 1e-9 * zl and abs(nz[95] - zh) < 1e-9 * zh
    kept = lam[~(lam < 1e-9)]
    c0 = tm.sturm_count(d, e, 1e-9)
    m = n - c0
    assert m == kept.size
    med = tm.kth_eigenvalue(d, e, c0 + m // 2) if m % 2 else 0.5 * (
        tm.kth_eigenvalue(d, e, c0 + m // 2 - 1) + tm.kth_eigenvalue(d, e, c0 + m // 2))
    assert abs(med - np.median(kept)) < 1e-10 * abs(lam).max()
