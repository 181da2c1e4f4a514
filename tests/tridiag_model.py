"""NumPy model of the eigenvector-free regression the CUDA path implements
(DESIGN.md section 3).  Test helper: proves on the CPU that the tridiagonal
formulation equals the reference's eigendecomposition formulation (REG:24-135).

  [0 ga^T; ga K]  --Householder-->  T (d, e), beta0, Q       (Q^T ga = beta0 e_0)
  F(g) = e_0^T (T+g)^-1 e_0 = 1/eta_0,   eta_k = d_k+g - e_k^2/eta_{k+1}
  sum_i P_i^2 r_i   = |ga|^2 g F          (r = g/(lambda+g))
  sum_i P_i^2 r_i^2 = |ga|^2 g^2 (-F')
  loss(g) = g * eta_0'/eta_0,  eta_k' = 1 + (e_k/eta_{k+1})^2 eta_{k+1}'
  yb = -beta0 Q (T+g)^-1 e_0;  noise = g |yb|^2 / (-ga.yb)
  Z_j = g |u_j|^2 / (b_j.u_j),  b_j = Q^T s_j, u_j = (T+g)^-1 b_j
"""
import numpy as np


def householder_tridiag(K, ga):
    """Returns d (n), e (n-1), beta0, and a function applying Q^T / Q."""
    n = K.shape[0]
    A = K.copy()
    refl = []  # (start, v, tau)

    def reflector(u):
        alpha = u[0]
        xnorm = np.linalg.norm(u[1:])
        if xnorm == 0.0:
            return np.r_[1.0, np.zeros(u.size - 1)], 0.0, alpha
        beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
        tau = (beta - alpha) / beta
        v = u / (alpha - beta)
        v[0] = 1.0
        return v, tau, beta

    v, tau, beta0 = reflector(ga.copy())
    refl.append((0, v, tau))
    w = tau * (A @ v)
    w += (-0.5 * tau * (w @ v)) * v
    A -= np.outer(v, w) + np.outer(w, v)
    d = np.zeros(n)
    e = np.zeros(n - 1)
    for j in range(n - 1):
        d[j] = A[j, j]
        u = A[j + 1:, j].copy()
        v, tau, beta = reflector(u)
        e[j] = beta
        refl.append((j + 1, v, tau))
        S = A[j + 1:, j + 1:]
        w = tau * (S @ v)
        w += (-0.5 * tau * (w @ v)) * v
        S -= np.outer(v, w) + np.outer(w, v)
    d[n - 1] = A[n - 1, n - 1]

    def apply_QT(B):
        B = B.copy()
        for (s, v, tau) in refl:
            B[s:] -= tau * np.outer(v, v @ B[s:]) if B.ndim == 2 else tau * v * (v @ B[s:])
        return B

    def apply_Q(x):
        x = x.copy()
        for (s, v, tau) in reversed(refl):
            x[s:] -= tau * v * (v @ x[s:])
        return x

    return d, e, beta0, apply_QT, apply_Q


def eta_jets(d, e, g):
    """Backward recurrence for eta_0 and its first two gamma-derivatives."""
    n = d.size
    eta, eta1, eta2 = d[n - 1] + g, 1.0, 0.0
    for k in range(n - 2, -1, -1):
        q = e[k] / eta
        q2 = q * q
        new2 = q2 * (eta2 - 2.0 * eta1 * eta1 / eta)
        new1 = 1.0 + q2 * eta1
        eta = d[k] + g - e[k] * q
        eta1, eta2 = new1, new2
    return eta, eta1, eta2


def loss_and_grad(d, e, g):
    eta, eta1, eta2 = eta_jets(d, e, g)
    f = g * eta1 / eta
    fp = eta1 / eta + g * eta2 / eta - g * eta1 * eta1 / (eta * eta)
    return f, fp


def sturm_count(d, e, x):
    """number of eigenvalues of T strictly below x."""
    cnt = 0
    q = d[0] - x
    if q < 0:
        cnt += 1
    for k in range(1, d.size):
        if q == 0.0:
            q = 1e-300
        q = d[k] - x - e[k - 1] * e[k - 1] / q
        if q < 0:
            cnt += 1
    return cnt


def kth_eigenvalue(d, e, k, iters=200):
    r = np.abs(d).copy()
    r[:-1] += np.abs(e)
    r[1:] += np.abs(e)
    lo, hi = -r.max(), r.max()
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        if sturm_count(d, e, mid) > k:
            hi = mid
        else:
            lo = mid
    return 0.5 * (lo + hi)


def tridiag_solve(d, e, g, B):
    """(T + g I)^-1 B via LDL^T (B: (n,) or (n, m))."""
    n = d.size
    piv = np.empty(n)
    l = np.empty(n - 1)
    piv[0] = d[0] + g
    for k in range(n - 1):
        l[k] = e[k] / piv[k]
        piv[k + 1] = d[k + 1] + g - l[k] * e[k]
    Z = np.array(B, dtype=np.float64, copy=True)
    for k in range(1, n):
        Z[k] -= l[k - 1] * Z[k - 1]
    Z[n - 1] /= piv[n - 1]
    for k in range(n - 2, -1, -1):
        Z[k] = Z[k] / piv[k] - l[k] * Z[k + 1]
    return Z
