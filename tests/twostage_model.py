"""NumPy model of the two-stage reduction the CUDA path implements.  Test helper: proves on
the CPU that the two-stage formulation equals the reference's eigendecomposition formulation
(interpolatory.py:24-135) and serves as the checker for the CUDA kernels of csrc/band.cu, chase.cu, bandsolve.cu.

  stage 0   H_{-1} ga = beta0 e_0,  K1 = H_{-1} K H_{-1}
  stage 1   K1 -> symmetric band T_b of half-width b: per panel k (columns c0 = k b .. c0+b-1) a Householder QR of
            the block below the band, Q_k = I - V T V^T (compact WY), and the two-sided rank-2b update
            A22 -= V W^T + W V^T,  W = (Y - 1/2 V T^T (V^T Y)) T,  Y = A22 V         -- all GEMM-shaped
  stage 2   T_b -> tridiagonal T (values only) by Householder bulge chasing; row/column 0 is never touched, so
            e_0 is preserved and  beta0^2 e_0^T (T+g)^-1 e_0 = ga^T (K+g)^-1 ga  (the loss only needs d, e)
  solve     everything that needs vectors (x, the Z-test) is solved in band coordinates with a banded Cholesky of
            T_b + g I: the stage-2 transformations are never applied to a vector.
"""
import numpy as np


def reflector(u):
    """LAPACK dlarfg: returns v (v[0] = 1), tau, beta with (I - tau v v^T) u = beta e_0."""
    alpha = u[0]
    xnorm = np.linalg.norm(u[1:]) if u.size > 1 else 0.0
    if xnorm == 0.0:
        v = np.zeros(u.size)
        v[0] = 1.0
        return v, 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
    tau = (beta - alpha) / beta
    v = u / (alpha - beta)
    v[0] = 1.0
    return v, tau, beta


def panel_qr(P):
    """Unblocked Householder QR of an m x b block (m >= 1).  Returns R-over-zeros (m x b), V (m x b unit lower
    trapezoidal, zero columns where no reflector exists), tau (b) and the compact-WY factor T (b x b upper)."""
    m, b = P.shape
    P = P.copy()
    V = np.zeros((m, b))
    tau = np.zeros(b)
    for t in range(min(b, m - 1)):
        v, tau[t], beta = reflector(P[t:, t].copy())
        V[t:, t] = v
        P[t, t] = beta
        P[t + 1:, t] = 0.0
        if t + 1 < b:
            w = tau[t] * (v @ P[t:, t + 1:])
            P[t:, t + 1:] -= np.outer(v, w)
    G = V.T @ V
    T = np.zeros((b, b))
    for t in range(b):
        T[t, t] = tau[t]
        if t:
            T[:t, t] = -tau[t] * (T[:t, :t] @ G[:t, t])
    return P, V, tau, T


def stage1(K, ga, b, S=None):
    """Returns dict(band (n x (b+1)): band[i, t] = T_b[i, i-t]; beta0; panels: list of (r0, V, T); v0, tau0;
    Bq = Q^T S)."""
    n = K.shape[0]
    A = K.copy()
    Bq = None if S is None else S.copy()
    v0, tau0, beta0 = reflector(ga.copy())
    w = tau0 * (A @ v0)
    w += (-0.5 * tau0 * (w @ v0)) * v0
    A -= np.outer(v0, w) + np.outer(w, v0)
    if Bq is not None:
        Bq -= tau0 * np.outer(v0, v0 @ Bq)
    panels = []
    k = 0
    while True:
        c0 = k * b
        r0 = c0 + b
        if n - r0 < 2:
            break
        R, V, tau, T = panel_qr(A[r0:, c0:r0])
        A[r0:, c0:r0] = R
        A[c0:r0, r0:] = R.T
        A22 = A[r0:, r0:]
        Y = A22 @ V
        M = V.T @ Y
        W = (Y - 0.5 * V @ (T.T @ M)) @ T
        A22 -= V @ W.T + W @ V.T
        if Bq is not None:
            Bq[r0:] -= V @ (T.T @ (V.T @ Bq[r0:]))
        panels.append((r0, V, T))
        k += 1
    band = np.zeros((n, b + 1))
    for t in range(min(b + 1, n)):
        band[t:, t] = np.diagonal(A, -t)
    resid = np.abs(np.tril(A, -(b + 1))).max() if n > b + 1 else 0.0
    return dict(band=band, beta0=beta0, panels=panels, v0=v0, tau0=tau0, Bq=Bq, resid=resid)


def stage1_deferred(K, ga, b, S=None, kd=2):
    """stage1 with the trailing update DEFERRED over groups of `kd` consecutive panels, the way csrc/band.cu runs it:
    slot 0 is H_{-1} (a block reflector with one non-zero column), slots 1.. are the QR panels; slots (0,1), (2,3), ...
    form groups.  Inside a group the trailing matrix stays STALE: the next panel's block column is brought up to date
    by a thin update, its Y = A22 V is computed from the stale matrix plus the corrections
        Y <- Y_stale - V_a (W_a^T V) - W_a (V_a^T V),    M = V^T Y = M_stale - G2^T G1 - G1^T G2,
    and ONE rank-2*kd*b update  A22 -= sum_q V_q W_q^T + W_q V_q^T  closes the group (twice the arithmetic intensity
    of the per-panel update for kd = 2).  Only the lower triangle is referenced, like the kernels do."""
    n = K.shape[0]
    A = np.tril(K.copy())
    Bq = None if S is None else S.copy()
    v0, tau0, beta0 = reflector(ga.copy())

    def symm_lower(r0, V):            # Y = A22 V from the lower triangle of A[r0:, r0:]
        L = A[r0:, r0:]
        return L @ V + np.tril(L, -1).T @ V

    panels = []
    slots = []                        # (r0, V (n - r0 x b), T (b x b))
    V0 = np.zeros((n, b))
    V0[:, 0] = v0
    T0 = np.zeros((b, b))
    T0[0, 0] = tau0
    slots.append((0, V0, T0))
    pend = []                         # pending (r0, V, W) of the current group, rows indexed from their own r0
    k = 0
    while True:
        if slots:
            r0, V, T = slots.pop()
        else:
            c0 = k * b
            r0 = c0 + b
            if n - r0 < 2:
                break
            # thin update of this panel's block column [c0:, c0:r0] with the pending reflectors of the group
            for (ra, Va, Wa) in pend:
                o = c0 - ra
                blk = Va[o:] @ Wa[o:o + b].T + Wa[o:] @ Va[o:o + b].T
                A[c0:, c0:r0] -= blk
            A[c0:r0, c0:r0] = np.tril(A[c0:r0, c0:r0])
            R, V, tau, T = panel_qr(A[r0:, c0:r0])
            A[r0:, c0:r0] = R
            panels.append((r0, V, T))
            k += 1
        Y = symm_lower(r0, V)
        M = V.T @ Y
        for (ra, Va, Wa) in pend:     # corrections for the stale trailing matrix
            o = r0 - ra
            G1 = Wa[o:].T @ V
            G2 = Va[o:].T @ V
            Y -= Va[o:] @ G1 + Wa[o:] @ G2
            M -= G2.T @ G1 + G1.T @ G2
        W = (Y - 0.5 * V @ (T.T @ M)) @ T
        if Bq is not None:
            Bq[r0:] -= V @ (T.T @ (V.T @ Bq[r0:]))
        pend.append((r0, V, W))
        next_exists = n - ((k + 1) * b) >= 2
        if len(pend) == kd or not next_exists:
            for (ra, Va, Wa) in pend:
                o = r0 - ra
                A[r0:, r0:] -= np.tril(Va[o:] @ Wa[o:].T + Wa[o:] @ Va[o:].T)
            pend = []
    band = np.zeros((n, b + 1))
    for t in range(min(b + 1, n)):
        band[t:, t] = np.diagonal(A, -t)
    return dict(band=band, beta0=beta0, panels=panels, v0=v0, tau0=tau0, Bq=Bq)


def apply_Q(s1, x):
    """Q x with Q = H_{-1} Q_0 Q_1 ... (the back-transformation yb = -beta0 Q x_b)."""
    x = np.array(x, dtype=np.float64, copy=True)
    for (r0, V, T) in reversed(s1["panels"]):
        x[r0:] -= V @ (T @ (V.T @ x[r0:]))
    x -= s1["tau0"] * s1["v0"] * (s1["v0"] @ x)
    return x


def band_to_dense(band):
    n, b1 = band.shape
    A = np.zeros((n, n))
    for t in range(min(b1, n)):
        A += np.diag(band[t:, t], -t)
        if t:
            A += np.diag(band[t:, t], t)
    return A


def stage2(band):
    """Band -> tridiagonal by Householder bulge chasing (Murata-Horikoshi / Lang), values only.
    Sweep j annihilates column j below the sub-diagonal with a reflector on rows j+1..j+b; the fill-in
    (one b x b bulge per hop) is chased down the band: each hop h applies the incoming reflector two-sided to
    the diagonal block I_h, from the right to the block below it, and generates the next reflector from that
    block's first column.  Returns d (n), e (n-1)."""
    n, b1 = band.shape
    b = b1 - 1
    A = band_to_dense(band)
    if b > 1:
        for j in range(n - 2):
            lo = j + 1
            hi = min(j + b, n - 1)            # inclusive
            if hi <= lo:
                break
            v, tau, beta = reflector(A[lo:hi + 1, j].copy())
            A[lo:hi + 1, j] = 0.0
            A[lo, j] = beta
            A[j, lo:hi + 1] = A[lo:hi + 1, j]
            while True:
                I = slice(lo, hi + 1)
                D = A[I, I]
                p = tau * (D @ v)
                p += (-0.5 * tau * (p @ v)) * v
                D -= np.outer(v, p) + np.outer(p, v)
                lo2 = hi + 1
                hi2 = min(hi + b, n - 1)
                if lo2 > n - 1:
                    break
                I2 = slice(lo2, hi2 + 1)
                Sb = A[I2, I]
                Sb -= tau * np.outer(Sb @ v, v)
                if hi2 > lo2:
                    v2, tau2, beta2 = reflector(Sb[:, 0].copy())
                    Sb[:, 0] = 0.0
                    Sb[0, 0] = beta2
                    Sb[:, 1:] -= tau2 * np.outer(v2, v2 @ Sb[:, 1:])
                else:
                    v2, tau2 = np.ones(1), 0.0
                A[I, I2] = Sb.T
                if hi2 <= lo2:
                    break
                v, tau, lo, hi = v2, tau2, lo2, hi2
    return np.diagonal(A).copy(), np.diagonal(A, -1).copy()


def band_cholesky(band, g):
    """Lower banded Cholesky T_b + g I = L L^T; L[i, t] = L(i, i-t)."""
    n, b1 = band.shape
    b = b1 - 1
    L = np.zeros((n, b1))
    for i in range(n):
        for t in range(min(b, i), -1, -1):          # column j = i - t, ascending j
            j = i - t
            s = band[i, t] + (g if t == 0 else 0.0)
            for p in range(1, b + 1):               # common columns j - p
                if j - p < 0 or t + p > b:
                    break
                s -= L[i, t + p] * L[j, p]
            L[i, t] = np.sqrt(s) if t == 0 else s / L[j, 0]
    return L


def band_solve(L, R):
    """Returns (y, z): y = L^-1 R, z = L^-T y  (R: (n,) or (n, m))."""
    n, b1 = L.shape
    b = b1 - 1
    y = np.array(R, dtype=np.float64, copy=True)
    for i in range(n):
        for p in range(1, min(b, i) + 1):
            y[i] = y[i] - L[i, p] * y[i - p]
        y[i] = y[i] / L[i, 0]
    z = y.copy()
    for i in range(n - 1, -1, -1):
        for p in range(1, b + 1):
            if i + p < n:
                z[i] = z[i] - L[i + p, p] * z[i + p]
        z[i] = z[i] / L[i, 0]
    return y, z


def panel_qr_one_exchange(P):
    """The formulation panel_qr_reg_kernel (csrc/band.cu) uses: per column ONE reduction over the rows -- the unscaled
    products u[c] = sum_{r > t} x_r P[r, c] (x = column t) for ALL columns c, next to row t itself.  u[t] is the
    squared norm the reflector needs, u[c > t] gives v^T P through  w_c = tau (P[t, c] + scale u[c]),  and u[c < t]
    gives the Gram entries  v_c . v_t = V[t, c] + scale u[c]  of the compact-WY factor (the finished columns hold
    their reflectors in place).  Returns what panel_qr returns."""
    m, b = P.shape
    P = P.copy()
    tau = np.zeros(b)
    G = np.zeros((b, b))
    nref = min(b, m - 1)
    for t in range(nref):
        x = P[:, t].copy()
        x[:t + 1] = 0.0
        u = x @ P                      # the one reduction (all b columns)
        row = P[t].copy()              # from the owner of row t
        ssq, alpha = u[t], row[t]
        if ssq == 0.0:
            beta, tau[t], scale = alpha, 0.0, 0.0
        else:
            beta = -np.copysign(np.sqrt(alpha * alpha + ssq), alpha)
            tau[t] = (beta - alpha) / beta
            scale = 1.0 / (alpha - beta)
        v = x * scale
        v[t] = 1.0
        w = tau[t] * (row + scale * u)
        P[t:, t + 1:] -= np.outer(v[t:], w[t + 1:])
        G[:t, t] = row[:t] + scale * u[:t]
        P[t, t] = beta
        P[t + 1:, t] = v[t + 1:]
    V = np.tril(P[:, :b], -1)
    V[:, nref:] = 0.0
    for t in range(nref):
        V[t, t] = 1.0
    T = np.zeros((b, b))
    for t in range(b):
        T[t, t] = tau[t]
        if t:
            T[:t, t] = -tau[t] * (T[:t, :t] @ G[:t, t])
    R = np.triu(P[:, :b]) if m >= b else np.triu(P)
    Rfull = np.zeros_like(P)
    Rfull[:R.shape[0]] = R[:m]
    return Rfull, V, tau, T


def band_ldl(band, g):
    """The factorisation band_solve_kernel (csrc/bandsolve.cu) runs: T_b + g I = L D L^T, left-looking by columns,
    no pivoting and no square roots (negative pivots pass through).  Returns Lg with Lg[i, 0] = D(i) and
    Lg[i, d] = L(i, i-d) for d >= 1."""
    n, b1 = band.shape
    b = b1 - 1
    Lg = np.zeros((n, b1))
    for k in range(n):
        # c_p = D(k-p) L(k, k-p)
        c = np.zeros(b + 1)
        for p in range(1, min(b, k) + 1):
            c[p] = Lg[k - p, 0] * Lg[k, p]
        s = np.zeros(b + 1)
        for d in range(0, b + 1):
            if k + d >= n:
                break
            s[d] = band[k + d, d] + (g if d == 0 else 0.0)
            for p in range(1, min(b - d, k) + 1):
                s[d] -= Lg[k + d, d + p] * c[p]
        Lg[k, 0] = s[0]
        for d in range(1, b + 1):
            if k + d < n:
                Lg[k + d, d] = s[d] / s[0]
    return Lg


def band_ldl_solve(Lg, R):
    """(y, z) with L y = R and z = L^-T D^-1 y; then  R . z = sum y^2 / D  (the Z-test denominator)."""
    n, b1 = Lg.shape
    b = b1 - 1
    y = np.array(R, dtype=np.float64, copy=True)
    for i in range(n):
        for p in range(1, min(b, i) + 1):
            y[i] = y[i] - Lg[i, p] * y[i - p]
    z = np.empty_like(y)
    for i in range(n - 1, -1, -1):
        z[i] = y[i] / Lg[i, 0]
        for p in range(1, b + 1):
            if i + p < n:
                z[i] = z[i] - Lg[i + p, p] * z[i + p]
    return y, z


class ChaseTasks:
    """Stage 2 cut into the hop tasks the CUDA kernels run (csrc/chase.cu), on the same compact storage
    (bw[i, d] = A(i, i-d), 2b sub-diagonals) and with the same dependency rule: hop h of sweep j may run once hop h-1
    of sweep j and hop h+1 of sweep j-1 are done (or sweep j-1 has finished).  `run(order_rng)` executes the hops in a
    RANDOM order that respects only that rule -- the result must not depend on the order."""

    def __init__(self, band):
        n, b1 = band.shape
        self.n, self.b = n, b1 - 1
        self.bw = np.zeros((n, 2 * self.b + 2))
        self.bw[:, :b1] = band
        self.state = {}      # sweep -> (h, lo, hi, v, tau) of its next hop, or None when finished
        self.done = {}       # sweep -> hops finished (large when the sweep is over)

    def A(self, r, c):
        return self.bw[r, r - c] if r >= c else self.bw[c, c - r]

    def setA(self, r, c, x):
        if r >= c:
            self.bw[r, r - c] = x
        else:
            self.bw[c, c - r] = x

    def ready(self, j):
        if j in self.done and self.done[j] >= 10 ** 9:
            return False
        h = self.state[j][0] if j in self.state else 0
        return j == 0 or self.done.get(j - 1, 0) >= h + 2

    def step(self, j):
        n, b = self.n, self.b
        if j not in self.state:                       # first reflector of the sweep
            lo, hi = j + 1, min(j + b, n - 1)
            x = np.array([self.A(r, j) for r in range(lo, hi + 1)])
            v, tau, beta = reflector(x)
            for i, r in enumerate(range(lo, hi + 1)):
                self.setA(r, j, beta if i == 0 else 0.0)
            self.state[j] = (0, lo, hi, v, tau)
        h, lo, hi, v, tau = self.state[j]
        I = list(range(lo, hi + 1))
        D = np.array([[self.A(r, c) for c in I] for r in I])
        p = tau * (D @ v)
        p += (-0.5 * tau * (p @ v)) * v
        D -= np.outer(v, p) + np.outer(p, v)
        for a, r in enumerate(I):
            for c_, c in enumerate(I):
                if c <= r:
                    self.setA(r, c, D[a, c_])
        lo2, hi2 = hi + 1, min(hi + b, n - 1)
        last = hi2 - lo2 + 1 <= 1
        v2, tau2 = None, 0.0
        if lo2 <= n - 1:
            I2 = list(range(lo2, hi2 + 1))
            S = np.array([[self.A(r, c) for c in I] for r in I2])
            S -= tau * np.outer(S @ v, v)
            if len(I2) >= 2:
                v2, tau2, beta2 = reflector(S[:, 0].copy())
                S[:, 0] = 0.0
                S[0, 0] = beta2
                S[:, 1:] -= tau2 * np.outer(v2, v2 @ S[:, 1:])
            for a, r in enumerate(I2):
                for c_, c in enumerate(I):
                    self.setA(r, c, S[a, c_])
        if last:
            self.done[j] = 10 ** 9
            self.state[j] = None
        else:
            self.done[j] = h + 1
            self.state[j] = (h + 1, lo2, hi2, v2, tau2)

    def run(self, rng):
        n = self.n
        sweeps = list(range(0, n - 2)) if self.b > 1 else []
        started = 0
        while True:
            cand = [j for j in sweeps[:started + 1] if self.ready(j)]
            if not cand:
                if started + 1 < len(sweeps):
                    started += 1
                    continue
                break
            j = cand[int(rng.integers(len(cand)))]
            self.step(j)
            if j == sweeps[min(started, len(sweeps) - 1)] and started + 1 < len(sweeps):
                started += 1
        return self.bw[:, 0].copy(), self.bw[1:, 1].copy()
