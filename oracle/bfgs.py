"""Oracle (test infrastructure): jax.scipy.optimize.minimize(method="BFGS") for d=1.

Third-party dependency restated: jax==0.4.28 `jax/_src/scipy/optimize/bfgs.py`
(minimize_bfgs) and `line_search.py` (line_search, _zoom, _cubicmin, _quadmin);
the reference's only call site is REG:133 with a 1-vector x0.  Scalars are used
throughout because d == 1 (H is a scalar, p = -H g).

`fun_and_grad(x) -> (f, g)` must return the loss and its derivative.
"""
import math


def _cubicmin(a, fa, fpa, b, fb, c, fc):
    with_nan = float("nan")
    try:
        C = fpa
        db = b - a
        dc = c - a
        denom = (db * dc) ** 2 * (db - dc)
        d2_0 = fb - fa - C * db
        d2_1 = fc - fa - C * dc
        A = (dc ** 2 * d2_0 - db ** 2 * d2_1) / denom
        B = (-(dc ** 3) * d2_0 + db ** 3 * d2_1) / denom
        radical = B * B - 3.0 * A * C
        if radical < 0:
            return with_nan
        return a + (-B + math.sqrt(radical)) / (3.0 * A)
    except (ZeroDivisionError, OverflowError, ValueError):
        return with_nan


def _quadmin(a, fa, fpa, b, fb):
    try:
        db = b - a
        B = (fb - fa - fpa * db) / (db ** 2)
        return a - fpa / (2.0 * B)
    except (ZeroDivisionError, OverflowError):
        return float("nan")


def _zoom(phi_fn, wolfe_one, wolfe_two, a_lo, phi_lo, dphi_lo, a_hi, phi_hi, dphi_hi, g_0):
    """line_search._zoom. Returns (failed, a_star, phi_star, dphi_star, g_star)."""
    done = False
    failed = False
    j = 0
    a_rec = (a_lo + a_hi) / 2.0
    phi_rec = (phi_lo + phi_hi) / 2.0
    a_star, phi_star, dphi_star, g_star = 1.0, phi_lo, dphi_lo, g_0
    while (not done) and (not failed):
        dalpha = a_hi - a_lo
        a = min(a_hi, a_lo)
        b = max(a_hi, a_lo)
        cchk = 0.2 * dalpha
        qchk = 0.1 * dalpha
        failed = failed or (dalpha <= 1e-10)  # signed; 1e-10 for 64-bit
        a_c = _cubicmin(a_lo, phi_lo, dphi_lo, a_hi, phi_hi, a_rec, phi_rec)
        use_cubic = (j > 0) and (a_c > a + cchk) and (a_c < b - cchk)
        a_q = _quadmin(a_lo, phi_lo, dphi_lo, a_hi, phi_hi)
        use_quad = (not use_cubic) and (a_q > a + qchk) and (a_q < b - qchk)
        if use_cubic:
            a_j = a_c
        elif use_quad:
            a_j = a_q
        else:
            a_j = (a_lo + a_hi) / 2.0
        phi_j, dphi_j, g_j = phi_fn(a_j)
        hi_to_j = wolfe_one(a_j, phi_j) or (phi_j >= phi_lo)
        star_to_j = wolfe_two(dphi_j) and (not hi_to_j)
        hi_to_lo = (dphi_j * (a_hi - a_lo) >= 0.0) and (not hi_to_j) and (not star_to_j)
        lo_to_j = (not hi_to_j) and (not star_to_j)
        if hi_to_j:
            a_rec, phi_rec = a_hi, phi_hi
            a_hi, phi_hi, dphi_hi = a_j, phi_j, dphi_j
        if star_to_j:
            done = True
            a_star, phi_star, dphi_star, g_star = a_j, phi_j, dphi_j, g_j
        if hi_to_lo:
            a_rec, phi_rec = a_hi, phi_hi
            a_hi, phi_hi, dphi_hi = a_lo, phi_lo, dphi_lo
        if lo_to_j:
            a_rec, phi_rec = a_lo, phi_lo
            a_lo, phi_lo, dphi_lo = a_j, phi_j, dphi_j
        j += 1
        failed = failed or (j >= 30)
    return failed, a_star, phi_star, dphi_star, g_star


def line_search(fun_and_grad, xk, pk, old_fval, old_old_fval, gfk, c1=1e-4, c2=0.9, maxiter=10):
    """Strong-Wolfe line search. Returns (failed, a_k, f_k, g_k)."""

    def phi_fn(t):
        f, g = fun_and_grad(xk + t * pk)
        return f, g * pk, g

    phi_0 = old_fval
    dphi_0 = gfk * pk
    try:
        cand = 1.01 * 2 * (phi_0 - old_old_fval) / dphi_0
    except ZeroDivisionError:
        cand = float("nan")
    start_value = 1.0 if cand > 1 else cand

    def wolfe_one(a_i, phi_i):
        return phi_i > phi_0 + c1 * a_i * dphi_0

    def wolfe_two(dphi_i):
        return abs(dphi_i) <= -c2 * dphi_0

    done = False
    failed = False
    i = 1
    a_i1, phi_i1, dphi_i1 = 0.0, phi_0, dphi_0
    a_star, phi_star, g_star = 0.0, phi_0, gfk
    while (not done) and (i <= maxiter) and (not failed):
        a_i = start_value if i == 1 else a_i1 * 2.0
        phi_i, dphi_i, g_i = phi_fn(a_i)
        to_zoom1 = wolfe_one(a_i, phi_i) or ((phi_i >= phi_i1) and (i > 1))
        to_i = wolfe_two(dphi_i) and (not to_zoom1)
        to_zoom2 = (dphi_i >= 0.0) and (not to_zoom1) and (not to_i)
        if to_zoom1:
            zf, a_star, phi_star, _, g_star = _zoom(
                phi_fn, wolfe_one, wolfe_two, a_i1, phi_i1, dphi_i1, a_i, phi_i, dphi_i, gfk)
            done = True
            failed = failed or zf
        if to_i:
            done = True
            a_star, phi_star, g_star = a_i, phi_i, g_i
        if to_zoom2:
            zf, a_star, phi_star, _, g_star = _zoom(
                phi_fn, wolfe_one, wolfe_two, a_i, phi_i, dphi_i, a_i1, phi_i1, dphi_i1, gfk)
            done = True
            failed = failed or zf
        i += 1
        a_i1, phi_i1, dphi_i1 = a_i, phi_i, dphi_i
    return (failed or (not done)), a_star, phi_star, g_star


def minimize_bfgs_1d(fun_and_grad, x0, gtol=1e-5, line_search_maxiter=10):
    """minimize_bfgs for a 1-vector. Returns (x, success, k)."""
    maxiter = 200
    f_k, g_k = fun_and_grad(x0)
    x_k = x0
    H_k = 1.0
    converged = abs(g_k) < gtol
    failed = False
    old_old_fval = f_k + abs(g_k) / 2.0
    k = 0
    while (not converged) and (not failed) and (k < maxiter):
        p_k = -H_k * g_k
        failed, a_k, f_kp1, g_kp1 = line_search(
            fun_and_grad, x_k, p_k, f_k, old_old_fval, g_k, maxiter=line_search_maxiter)
        s_k = a_k * p_k
        x_kp1 = x_k + s_k
        y_k = g_kp1 - g_k
        ys = y_k * s_k
        rho_k = (1.0 / ys) if ys != 0.0 else math.copysign(float("inf"), ys if ys == ys else 1.0)
        if math.isfinite(rho_k):
            w = 1.0 - rho_k * s_k * y_k
            H_k = w * H_k * w + rho_k * s_k * s_k
        converged = abs(g_kp1) < gtol
        old_old_fval = f_k
        x_k, f_k, g_k = x_kp1, f_kp1, g_kp1
        k += 1
    return x_k, (converged and not failed), k
