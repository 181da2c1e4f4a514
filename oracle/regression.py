"""Oracle (test infrastructure): per-node kernel ridge regression, gamma search,
variational solve and Monte-Carlo Z-test, restated from REG
(src/ComputationalHypergraphDiscovery/interpolatory.py, byte-identical twin
non_interpolatory.py) on top of numpy.linalg.eigh (LAPACK syevd), as the
reference does with jnp.linalg.eigh (REG:48).
"""
import numpy as np

from . import jaxprng
from .bfgs import minimize_bfgs_1d

GRID = 10 ** np.linspace(-6.0, 6.0, 10000)   # jnp.logspace(-6, 6, 10000)  REG:129
Z_SAMPLES = 100                              # REG:98


def make_loss(eigenvalues, Pga):
    """loss(gamma) and d loss/d gamma  (REG:124-126; gradient = what jax autodiff
    returns for that expression)."""
    p2 = Pga ** 2

    def loss_and_grad(g):
        with np.errstate(all="ignore"):
            den = eigenvalues + g
            r = g / den
            rp = eigenvalues / den ** 2
            num = np.sum(p2 * r * r)
            d = np.sum(r * p2)
            f = num / d
            dnum = np.sum(2.0 * p2 * r * rp)
            dd = np.sum(rp * p2)
            grad = (dnum * d - num * dd) / d ** 2
        return float(f), float(grad)

    return loss_and_grad


def grid_losses(eigenvalues, Pga, chunk=500):
    """jax.vmap(loss)(test_vals)  REG:130."""
    p2 = Pga ** 2
    out = np.empty(GRID.size)
    with np.errstate(all="ignore"):
        for s in range(0, GRID.size, chunk):
            g = GRID[s:s + chunk, None]
            r = g / (eigenvalues[None, :] + g)
            out[s:s + chunk] = np.sum(p2[None, :] * r * r, axis=1) / np.sum(r * p2[None, :], axis=1)
    return out


def _argmin_nanfirst(a):
    """jnp.argmin: first occurrence; NaN is treated as the minimum."""
    nan = np.isnan(a)
    if nan.any():
        return int(np.argmax(nan))
    return int(np.argmin(a))


def find_gamma(eigenvalues, Pga, gamma_min, info=None):
    """REG:110-135."""
    kept = eigenvalues[~(eigenvalues < gamma_min)]
    median = float(np.median(kept)) if kept.size else float("nan")   # nanmedian REG:122
    test = grid_losses(eigenvalues, Pga)
    best = _argmin_nanfirst(test)
    guess = median if best == 0 else float(GRID[best])               # REG:132
    x, success, k = minimize_bfgs_1d(make_loss(eigenvalues, Pga), guess)  # REG:133
    if info is not None:
        info.update(best=best, guess=guess, median=median, success=success, k=k, x=float(x))
    return float(x) if success else median                            # REG:134


def solve_variationnal(ga, gamma, eigenvalues, eigenvectors):
    """REG:61-82."""
    Pga = eigenvectors.T @ ga
    coeffs = gamma / (eigenvalues + gamma)
    Pc = Pga * coeffs
    noise = np.dot(Pc, Pc) / np.dot(Pc, Pga)
    yb = -(eigenvectors @ (Pga / (eigenvalues + gamma)))
    return yb, float(noise)


def Z_test(gamma, eigenvalues, eigenvectors, key):
    """REG:85-107."""
    N = Z_SAMPLES
    samples = jaxprng.normal(key, (eigenvectors.shape[0], N))
    Pgas = eigenvectors.T @ samples
    coeffs = gamma / (eigenvalues + gamma)
    Pc = Pgas * coeffs[:, None]
    noises = np.sum(Pc * Pc, axis=0) / np.sum(Pc * Pgas, axis=0)
    B = np.sort(noises)                     # NaNs last, like jnp.sort
    return float(B[int(0.05 * N)]), float(B[int(0.95 * N)])


def perform_regression_and_find_gamma(K, ga, gamma_min, key, info=None):
    """REG:24-58.  Returns (yb, noise, Z_low, Z_high, gamma_raw, new_key)."""
    ks = jaxprng.split(key)
    key, subkey = ks[0], ks[1]
    eigenvalues, eigenvectors = np.linalg.eigh(K)
    Pga = eigenvectors.T @ ga
    gamma = find_gamma(eigenvalues, Pga, gamma_min, info)
    gamma_used = max(gamma, gamma_min) if not np.isnan(gamma) else float("nan")
    yb, noise = solve_variationnal(ga, gamma_used, eigenvalues, eigenvectors)
    Z_low, Z_high = Z_test(gamma_used, eigenvalues, eigenvectors, subkey)
    if info is not None:
        info.update(eigenvalues=eigenvalues, gamma_used=gamma_used)
    return yb, noise, Z_low, Z_high, gamma, key
