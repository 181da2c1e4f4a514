"""NumPy model of the narrow-band formulation: the Gram matrix is reduced only to a
symmetric band of half-width B (never to a tridiagonal), and every quantity the regression consumes is obtained
from banded UL factorisations of T_b + gamma I.  B = 1 is the tridiagonal formulation of tridiag_model.py.
Test helper: proves on the CPU that the band formulation equals the reference's eigendecomposition formulation
(interpolatory.py:24-135) and serves as the checker for the CUDA band kernels."""
import numpy as np


def band_reduce(K, ga, B):
    """Householder reduction of the bordered matrix [0 ga^T; ga K] to half-bandwidth B.
    Returns band (B+1, n) with band[t, k] = T_b[k, k-t], gt (B,) = first B entries of Q^T ga (the rest is 0),
    and the reflectors [(pivot_row, v, tau)] (K-index) of Q = H_0 H_1 ..."""
    n = K.shape[0]
    M = np.zeros((n + 1, n + 1))
    M[0, 1:] = ga
    M[1:, 0] = ga
    M[1:, 1:] = K
    refl = []
    for p in range(n + 1):
        q = p + B                      # pivot row (bordered index)
        if q >= n:                     # fewer than two rows left: nothing to annihilate
            break
        u = M[q:, p].copy()
        alpha, xnorm = u[0], np.linalg.norm(u[1:])
        if xnorm == 0.0:
            continue
        beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
        tau = (beta - alpha) / beta
        v = u / (alpha - beta)
        v[0] = 1.0
        S = M[q:, :]
        S -= tau * np.outer(v, v @ S)
        S = M[:, q:]
        S -= tau * np.outer(S @ v, v)
        refl.append((q - 1, v, tau))
    T = M[1:, 1:]
    band = np.zeros((B + 1, n))
    for t in range(B + 1):
        band[t, t:] = np.diagonal(T, -t)
    gt = M[1:B + 1, 0].copy()
    resid = np.abs(np.tril(T, -(B + 1))).max() if n > B + 1 else 0.0
    return band, gt, refl, resid


def apply_QT(refl, X):
    X = np.array(X, dtype=np.float64, copy=True)
    for (s, v, tau) in refl:
        X[s:] -= tau * (np.outer(v, v @ X[s:]) if X.ndim == 2 else v * (v @ X[s:]))
    return X


def apply_Q(refl, x):
    x = np.array(x, dtype=np.float64, copy=True)
    for (s, v, tau) in reversed(refl):
        x[s:] -= tau * v * (v @ x[s:])
    return x


class Jet:
    """value + first two derivatives w.r.t. gamma (forward mode)."""
    __slots__ = ("v", "d1", "d2")

    def __init__(self, v, d1=0.0, d2=0.0):
        self.v, self.d1, self.d2 = v, d1, d2

    @staticmethod
    def lift(x):
        return x if isinstance(x, Jet) else Jet(x)

    def __add__(self, o):
        o = Jet.lift(o)
        return Jet(self.v + o.v, self.d1 + o.d1, self.d2 + o.d2)
    __radd__ = __add__

    def __sub__(self, o):
        o = Jet.lift(o)
        return Jet(self.v - o.v, self.d1 - o.d1, self.d2 - o.d2)

    def __rsub__(self, o):
        return Jet.lift(o) - self

    def __mul__(self, o):
        o = Jet.lift(o)
        return Jet(self.v * o.v, self.d1 * o.v + self.v * o.d1, self.d2 * o.v + 2 * self.d1 * o.d1 + self.v * o.d2)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Jet.lift(o)
        q = self.v / o.v
        q1 = (self.d1 - q * o.d1) / o.v
        q2 = (self.d2 - 2 * q1 * o.d1 - q * o.d2) / o.v
        return Jet(q, q1, q2)

    def __rtruediv__(self, o):
        return Jet.lift(o) / self


def band_ul(band, g):
    """UL factorisation T_b + g I = U D U^T (U unit upper triangular, bandwidth B), bottom-up.
    g may be a float or a Jet.  Returns D (list, index k) and Uc (list of lists: Uc[k][t-1] = U[k-t, k])."""
    B, n = band.shape[0] - 1, band.shape[1]
    D = [None] * n
    Uc = [None] * n
    for k in range(n - 1, -1, -1):
        dk = float(band[0, k]) + g
        for m in range(1, B + 1):
            if k + m < n:
                u = Uc[k + m][m - 1]          # U[k, k+m]
                dk = dk - u * u * D[k + m]
        col = []
        for t in range(1, B + 1):             # U[k-t, k]
            if k - t < 0:
                col.append(0.0)
                continue
            s = float(band[t, k])            # T[k, k-t] = M[k-t, k]
            for m in range(1, B - t + 1):
                if k + m < n:
                    s = s - Uc[k + m][t + m - 1] * Uc[k + m][m - 1] * D[k + m]
            col.append(s / dk)
        D[k], Uc[k] = dk, col
    return D, Uc


def F_of(band, gt, g):
    """F(g) = gt^T (T_b + g)^-1 gt  (gt supported on the first B rows); g float or Jet."""
    B = band.shape[0] - 1
    D, Uc = band_ul(band, g)
    z = [None] * B
    F = 0.0
    for k in range(B - 1, -1, -1):
        zk = gt[k]
        for m in range(1, B - k):
            zk = zk - Uc[k + m][m - 1] * z[k + m]
        z[k] = zk
        F = F + zk * zk / D[k]
    return F


def loss_and_grad(band, gt, g):
    """loss = sum (P r)^2 / sum (r P^2) = -g F'/F and its derivative (interpolatory.py:124-126)."""
    F = F_of(band, gt, Jet(g, 1.0, 0.0))
    loss = -g * F.d1 / F.v
    grad = -F.d1 / F.v - g * F.d2 / F.v + g * F.d1 * F.d1 / (F.v * F.v)
    return loss, grad


def band_solve(band, g, Bm):
    """(T_b + g I)^-1 Bm for a matrix / vector right-hand side."""
    B, n = band.shape[0] - 1, band.shape[1]
    D, Uc = band_ul(band, g)
    Z = np.array(Bm, dtype=np.float64, copy=True)
    for k in range(n - 1, -1, -1):             # U z = b  (back substitution, bottom-up)
        for m in range(1, B + 1):
            if k + m < n:
                Z[k] = Z[k] - Uc[k + m][m - 1] * Z[k + m]
    for k in range(n):
        Z[k] = Z[k] / D[k]
    for k in range(n):                         # U^T x = y  (top-down)
        for t in range(1, B + 1):
            if k - t >= 0:
                Z[k] = Z[k] - Uc[k][t - 1] * Z[k - t]
    return Z


def inertia_below(band, x):
    """number of eigenvalues of T_b strictly below x (negative pivots of T_b - x I)."""
    D, _ = band_ul(band, -x)
    return int(sum(1 for d in D if d < 0))


def kth_eigenvalue(band, k, iters=200):
    B, n = band.shape[0] - 1, band.shape[1]
    r = np.abs(band[0]).copy()
    for t in range(1, B + 1):
        r[t:] += np.abs(band[t, t:])
        r[:n - t] += np.abs(band[t, t:])
    lo, hi = -r.max(), r.max()
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        if inertia_below(band, mid) > k:
            hi = mid
        else:
            lo = mid
    return 0.5 * (lo + hi)
