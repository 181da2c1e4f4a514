// Scalar restatement of jax 0.4.28 `minimize(method="BFGS")` for a 1-vector with its strong-Wolfe line search
// (jax/_src/scipy/optimize/{bfgs,line_search}.py; cf. oracle/bfgs.py), generic in the loss evaluator
// EV: void operator()(double gamma, double& loss, double& dloss) const.
#pragma once
#include "common.cuh"

namespace chd {

__device__ inline double cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc) {
  const double C = fpa;
  const double db = b - a, dc = c - a;
  const double denom = (db * dc) * (db * dc) * (db - dc);
  const double d20 = fb - fa - C * db, d21 = fc - fa - C * dc;
  const double A = (dc * dc * d20 - db * db * d21) / denom;
  const double B = (-(dc * dc * dc) * d20 + db * db * db * d21) / denom;
  const double radical = B * B - 3.0 * A * C;
  return a + (-B + sqrt(radical)) / (3.0 * A);
}

__device__ inline double quadmin(double a, double fa, double fpa, double b, double fb) {
  const double db = b - a;
  const double B = (fb - fa - fpa * db) / (db * db);
  return a - fpa / (2.0 * B);
}

template <class EV>
struct LineFn {
  const EV* ev;
  double xk, pk;
  __device__ void operator()(double t, double& phi, double& dphi, double& g) const {
    (*ev)(xk + t * pk, phi, g);
    dphi = g * pk;
  }
};

struct Wolfe {
  double phi_0, dphi_0;
  __device__ bool one(double a_i, double phi_i) const { return phi_i > phi_0 + 1e-4 * a_i * dphi_0; }
  __device__ bool two(double dphi_i) const { return fabs(dphi_i) <= -0.9 * dphi_0; }
};

template <class EV>
__device__ bool zoom(const LineFn<EV>& fn, const Wolfe& wf, double a_lo, double phi_lo, double dphi_lo, double a_hi,
                     double phi_hi, double dphi_hi, double g_0, double& a_star, double& phi_star,
                     double& g_star) {
  bool done = false, failed = false;
  int j = 0;
  double a_rec = (a_lo + a_hi) / 2.0, phi_rec = (phi_lo + phi_hi) / 2.0;
  a_star = 1.0; phi_star = phi_lo; g_star = g_0;
  while (!done && !failed) {
    const double dalpha = a_hi - a_lo;
    const double a = fmin(a_hi, a_lo), b = fmax(a_hi, a_lo);
    const double cchk = 0.2 * dalpha, qchk = 0.1 * dalpha;
    failed = failed || (dalpha <= 1e-10);
    const double a_c = cubicmin(a_lo, phi_lo, dphi_lo, a_hi, phi_hi, a_rec, phi_rec);
    const bool use_cubic = (j > 0) && (a_c > a + cchk) && (a_c < b - cchk);
    const double a_q = quadmin(a_lo, phi_lo, dphi_lo, a_hi, phi_hi);
    const bool use_quad = (!use_cubic) && (a_q > a + qchk) && (a_q < b - qchk);
    const double a_j = use_cubic ? a_c : (use_quad ? a_q : (a_lo + a_hi) / 2.0);
    double phi_j, dphi_j, g_j;
    fn(a_j, phi_j, dphi_j, g_j);
    const bool hi_to_j = wf.one(a_j, phi_j) || (phi_j >= phi_lo);
    const bool star_to_j = wf.two(dphi_j) && !hi_to_j;
    const bool hi_to_lo = (dphi_j * (a_hi - a_lo) >= 0.0) && !hi_to_j && !star_to_j;
    const bool lo_to_j = !hi_to_j && !star_to_j;
    if (hi_to_j) {
      a_rec = a_hi; phi_rec = phi_hi;
      a_hi = a_j; phi_hi = phi_j; dphi_hi = dphi_j;
    }
    if (star_to_j) {
      done = true;
      a_star = a_j; phi_star = phi_j; g_star = g_j;
    }
    if (hi_to_lo) {
      a_rec = a_hi; phi_rec = phi_hi;
      a_hi = a_lo; phi_hi = phi_lo; dphi_hi = dphi_lo;
    }
    if (lo_to_j) {
      a_rec = a_lo; phi_rec = phi_lo;
      a_lo = a_j; phi_lo = phi_j; dphi_lo = dphi_j;
    }
    ++j;
    failed = failed || (j >= 30);
  }
  return failed;
}

// returns failed; outputs a_k, f_k, g_k
template <class EV>
__device__ bool line_search(const EV& T, double xk, double pk, double old_fval, double old_old_fval, double gfk,
                            double& a_k, double& f_k, double& g_k) {
  LineFn<EV> fn{&T, xk, pk};
  const double phi_0 = old_fval, dphi_0 = gfk * pk;
  Wolfe wf{phi_0, dphi_0};
  const double cand = 1.01 * 2.0 * (phi_0 - old_old_fval) / dphi_0;
  const double start_value = (cand > 1.0) ? 1.0 : cand;
  bool done = false, failed = false;
  int i = 1;
  double a_i1 = 0.0, phi_i1 = phi_0, dphi_i1 = dphi_0;
  double a_star = 0.0, phi_star = phi_0, g_star = gfk;
  while (!done && i <= 10 && !failed) {
    const double a_i = (i == 1) ? start_value : a_i1 * 2.0;
    double phi_i, dphi_i, g_i;
    fn(a_i, phi_i, dphi_i, g_i);
    const bool z1 = wf.one(a_i, phi_i) || ((phi_i >= phi_i1) && (i > 1));
    const bool to_i = wf.two(dphi_i) && !z1;
    const bool z2 = (dphi_i >= 0.0) && !z1 && !to_i;
    if (z1) {
      const bool zf = zoom(fn, wf, a_i1, phi_i1, dphi_i1, a_i, phi_i, dphi_i, gfk, a_star, phi_star, g_star);
      done = true;
      failed = failed || zf;
    }
    if (to_i) {
      done = true;
      a_star = a_i; phi_star = phi_i; g_star = g_i;
    }
    if (z2) {
      const bool zf = zoom(fn, wf, a_i, phi_i, dphi_i, a_i1, phi_i1, dphi_i1, gfk, a_star, phi_star, g_star);
      done = true;
      failed = failed || zf;
    }
    ++i;
    a_i1 = a_i; phi_i1 = phi_i; dphi_i1 = dphi_i;
  }
  a_k = a_star; f_k = phi_star; g_k = g_star;
  return failed || !done;
}

// minimize_bfgs for a 1-vector; returns success
template <class EV>
__device__ bool bfgs_1d(const EV& T, double x0, double& x_out) {
  double f_k, g_k;
  T(x0, f_k, g_k);
  double x_k = x0, H_k = 1.0;
  bool converged = fabs(g_k) < 1e-5, failed = false;
  double old_old = f_k + fabs(g_k) / 2.0;
  int k = 0;
  while (!converged && !failed && k < 200) {
    const double p_k = -H_k * g_k;
    double a_k, f_n, g_n;
    failed = line_search(T, x_k, p_k, f_k, old_old, g_k, a_k, f_n, g_n);
    const double s_k = a_k * p_k;
    const double x_n = x_k + s_k;
    const double y_k = g_n - g_k;
    const double rho = 1.0 / (y_k * s_k);
    if (isfinite(rho)) {
      const double w = 1.0 - rho * s_k * y_k;
      H_k = w * H_k * w + rho * s_k * s_k;
    }
    converged = fabs(g_n) < 1e-5;
    old_old = f_k;
    x_k = x_n; f_k = f_n; g_k = g_n;
    ++k;
  }
  x_out = x_k;
  return converged && !failed;
}


}  // namespace chd
