// Gamma search, variational solve, noise ratio and Z-test on a symmetric BAND form T_b = Q^T K Q of half-width B
// (narrow-band formulation; NumPy model and CPU proof: tests/band_model.py).
// Generalises gamma.cu (B = 1).  Everything is a banded UL factorisation T_b + g I = U D U^T, U unit upper
// triangular with bandwidth B, computed bottom-up with a B x B sliding window:
//   D_k      = T_kk + g - sum_{m=1..B} U_{k,k+m}^2 D_{k+m}
//   U_{k-t,k} = (T_{k,k-t} - sum_{m=1..B-t} U_{k-t,k+m} U_{k,k+m} D_{k+m}) / D_k          t = 1..B
//   F(g)     = gt^T (T_b + g)^-1 gt = sum_{k<B} z_k^2 / D_k,   U z = gt   (gt = Q^T ga lives on the first B rows)
//   loss(g)  = -g F'/F,  loss' = -F'/F - g F''/F + g F'^2/F^2            (interpolatory.py:124-126)
// with F', F'' from forward-mode jets pushed through the same recurrence; inertia counts (negative D_k) give
// the median / smallest eigenvalue; the final factor is stored for x = (T_b+g)^-1 gt and the 100 Z-test solves.
#include "bfgs.cuh"
#include "internal.cuh"

namespace chd {

template <int O>
struct J {
  double v, d1, d2;
};
template <int O>
__device__ __forceinline__ J<O> jconst(double x) { return J<O>{x, 0.0, 0.0}; }
template <int O>
__device__ __forceinline__ J<O> jsub(const J<O>& a, const J<O>& b) {
  J<O> r;
  r.v = a.v - b.v;
  if (O >= 1) r.d1 = a.d1 - b.d1;
  if (O >= 2) r.d2 = a.d2 - b.d2;
  return r;
}
template <int O>
__device__ __forceinline__ J<O> jadd(const J<O>& a, const J<O>& b) {
  J<O> r;
  r.v = a.v + b.v;
  if (O >= 1) r.d1 = a.d1 + b.d1;
  if (O >= 2) r.d2 = a.d2 + b.d2;
  return r;
}
template <int O>
__device__ __forceinline__ J<O> jmul(const J<O>& a, const J<O>& b) {
  J<O> r;
  r.v = a.v * b.v;
  if (O >= 1) r.d1 = a.d1 * b.v + a.v * b.d1;
  if (O >= 2) r.d2 = a.d2 * b.v + 2.0 * a.d1 * b.d1 + a.v * b.d2;
  return r;
}
template <int O>
__device__ __forceinline__ J<O> jrecip(const J<O>& a) {
  J<O> r;
  const double q = 1.0 / a.v;
  r.v = q;
  if (O >= 1) r.d1 = -a.d1 * q * q;
  if (O >= 2) r.d2 = (2.0 * a.d1 * a.d1 * q - a.d2) * q * q;
  return r;
}

// sliding window of the last B columns of U and D (columns k+1 .. k+B while column k is being formed)
template <int O, int B>
struct BandWin {
  J<O> U[B][B];   // U[m][t] = U_{k+1+m-(t+1), k+1+m}
  J<O> D[B];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int m = 0; m < B; ++m) {
      D[m] = jconst<O>(1.0);
#pragma unroll
      for (int t = 0; t < B; ++t) U[m][t] = jconst<O>(0.0);
    }
  }
  // bcol[t] = T_b[k][k-t] (t = 0..B; zero when k < t).  Returns D_k and writes column k of U (ucol[t] = U_{k-1-t,k}).
  // pivmin > 0 activates the Sturm-count guard (tiny pivots are replaced by -pivmin).
  __device__ __forceinline__ J<O> step(const double* bcol, const J<O>& g, J<O>* ucol, double pivmin = 0.0) {
    J<O> dk = g;
    dk.v += bcol[0];
#pragma unroll
    for (int m = 0; m < B; ++m) dk = jsub(dk, jmul(jmul(U[m][m], U[m][m]), D[m]));
    if (pivmin > 0.0 && fabs(dk.v) <= pivmin) dk.v = -pivmin;
    const J<O> inv = jrecip(dk);
#pragma unroll
    for (int t = 0; t < B; ++t) {
      J<O> s = jconst<O>(bcol[t + 1]);
#pragma unroll
      for (int m = 0; m + t + 1 < B; ++m) s = jsub(s, jmul(jmul(U[m][t + 1 + m], U[m][m]), D[m]));
      ucol[t] = jmul(s, inv);
    }
#pragma unroll
    for (int m = B - 1; m > 0; --m) {
      D[m] = D[m - 1];
#pragma unroll
      for (int t = 0; t < B; ++t) U[m][t] = U[m - 1][t];
    }
    D[0] = dk;
#pragma unroll
    for (int t = 0; t < B; ++t) U[0][t] = ucol[t];
    return dk;
  }
  // after column 0 has been processed: F = gt^T (T_b + g)^-1 gt
  __device__ __forceinline__ J<O> quadform(const double* gt) const {
    J<O> z[B];
    J<O> F = jconst<O>(0.0);
#pragma unroll
    for (int k = B - 1; k >= 0; --k) {
      J<O> zk = jconst<O>(gt[k]);
#pragma unroll
      for (int j = k + 1; j < B; ++j) zk = jsub(zk, jmul(U[j][j - k - 1], z[j]));
      z[k] = zk;
      F = jadd(F, jmul(jmul(zk, zk), jrecip(D[k])));
    }
    return F;
  }
};

struct BandT {
  const double* band;   // (B+1) x n : band[t*n + k] = T_b[k][k-t]
  const double* gt;     // B
  int n;
};

template <int O, int B>
__device__ J<O> band_F(const BandT& T, double g) {
  BandWin<O, B> W;
  W.init();
  J<O> gj = jconst<O>(g);
  if (O >= 1) gj.d1 = 1.0;
  J<O> ucol[B];
  for (int k = T.n - 1; k >= 0; --k) {
    double bc[B + 1];
#pragma unroll
    for (int t = 0; t <= B; ++t) bc[t] = (k >= t) ? T.band[(size_t)t * T.n + k] : 0.0;
    W.step(bc, gj, ucol);
  }
  return W.quadform(T.gt);
}

template <int B>
struct BandEval {
  BandT T;
  __device__ void operator()(double g, double& f, double& fp) const {
    const J<2> F = band_F<2, B>(T, g);
    f = -g * F.d1 / F.v;
    fp = -F.d1 / F.v - g * F.d2 / F.v + g * F.d1 * F.d1 / (F.v * F.v);
  }
};

template <int B>
__device__ int band_inertia(const BandT& T, double x, double pivmin) {
  BandWin<0, B> W;
  W.init();
  const J<0> gj = jconst<0>(-x);
  J<0> ucol[B];
  int cnt = 0;
  for (int k = T.n - 1; k >= 0; --k) {
    double bc[B + 1];
#pragma unroll
    for (int t = 0; t <= B; ++t) bc[t] = (k >= t) ? T.band[(size_t)t * T.n + k] : 0.0;
    const J<0> dk = W.step(bc, gj, ucol, pivmin);
    cnt += (dk.v < 0.0);
  }
  return cnt;
}

constexpr int GB_THREADS = 128;
constexpr int GB_CHUNK = 256;

template <int B>
__global__ void __launch_bounds__(GB_THREADS) gamma_band_grid_kernel(const double* __restrict__ band,
                                                                     const double* __restrict__ gt, int n,
                                                                     const double* __restrict__ grid, int ngrid,
                                                                     double* __restrict__ loss) {
  __shared__ double sb[(B + 1) * GB_CHUNK];
  const int t = blockIdx.y;
  const double* bt = band + (size_t)t * (B + 1) * n;
  const int gi = blockIdx.x * GB_THREADS + threadIdx.x;
  const double g = (gi < ngrid) ? grid[gi] : 1.0;
  BandWin<1, B> W;
  W.init();
  J<1> gj = jconst<1>(g);
  gj.d1 = 1.0;
  J<1> ucol[B];
  for (int hi = n - 1; hi >= 0; hi -= GB_CHUNK) {
    const int lo = max(hi - GB_CHUNK + 1, 0);
    const int cnt = hi - lo + 1;
    __syncthreads();
    for (int q = threadIdx.x; q < cnt * (B + 1); q += GB_THREADS) {
      const int tt = q / cnt, kk = q - tt * cnt;
      sb[tt * GB_CHUNK + kk] = (lo + kk >= tt) ? bt[(size_t)tt * n + lo + kk] : 0.0;
    }
    __syncthreads();
    for (int q = cnt - 1; q >= 0; --q) {
      double bc[B + 1];
#pragma unroll
      for (int tt = 0; tt <= B; ++tt) bc[tt] = sb[tt * GB_CHUNK + q];
      W.step(bc, gj, ucol);
    }
  }
  const J<1> F = W.quadform(gt + (size_t)t * B);
  if (gi < ngrid) loss[(size_t)t * ngrid + gi] = -g * F.d1 / F.v;
}

// k-th smallest eigenvalue (0-based) of T_b by inertia multisection; all threads of the block participate.
template <int B>
__device__ double band_kth_eigenvalue(const BandT& T, int kth, double glo, double ghi, double pivmin, int* scnt,
                                      double* sx) {
  const int nt = blockDim.x, tid = threadIdx.x;
  double lo = glo, hi = ghi;
  for (int it = 0; it < 64; ++it) {
    const double w = hi - lo;
    if (!(w > 4.0 * 2.220446049250313e-16 * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin)) break;
    const double x = lo + w * ((double)(tid + 1) / (double)(nt + 1));
    scnt[tid] = band_inertia<B>(T, x, pivmin);
    sx[tid] = x;
    __syncthreads();
    double nlo = lo, nhi = hi;
    if (scnt[0] > kth) {
      nhi = sx[0];
    } else if (scnt[nt - 1] <= kth) {
      nlo = sx[nt - 1];
    } else {
      int a = 0, b = nt - 1;
      while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (scnt[mid] <= kth) a = mid; else b = mid;
      }
      nlo = sx[a];
      nhi = sx[b];
    }
    __syncthreads();
    lo = nlo; hi = nhi;
  }
  return 0.5 * (lo + hi);
}

struct BandFinishArgs {
  const double* band;   // T x (B+1) x n
  const double* gt;     // T x B
  const double* loss;   // T x ngrid
  const double* grid;
  int ngrid;
  double* Bq;           // T x n x ldb  (Q^T S), overwritten
  int ldb, nz;
  int gamma_min_mode;
  double gamma_min_base;
  double* gamma_min;
  int want_lambda_min;
  double* lambda_min;
  double* Dg;           // T x n       scratch: pivots of the final factor
  double* Ug;           // T x n x B   scratch: columns of U
  double* x;            // T x n out: (T_b + g)^-1 gt
  double* noise; double* z_low; double* z_high; double* gamma;
  size_t out_stride;
  int32_t* status;
  int n;
};

template <int B>
__global__ void __launch_bounds__(256) gamma_band_finish_kernel(BandFinishArgs a) {
  __shared__ double s_val[256];
  __shared__ int s_idx[256];
  __shared__ int s_cnt[256];
  __shared__ double s_x[256];
  __shared__ double s_sc[8];
  __shared__ double s_z[128];
  const int t = blockIdx.x, tid = threadIdx.x, n = a.n;
  BandT T{a.band + (size_t)t * (B + 1) * n, a.gt + (size_t)t * B, n};

  // Gershgorin interval and pivmin
  double gl = 1e300, gh = -1e300, emax = 0.0;
  for (int k = tid; k < n; k += blockDim.x) {
    double off = 0.0;
    for (int tt = 1; tt <= B; ++tt) {
      if (k >= tt) { const double e = fabs(T.band[(size_t)tt * n + k]); off += e; emax = fmax(emax, e * e); }
      if (k + tt < n) off += fabs(T.band[(size_t)tt * n + k + tt]);
    }
    gl = fmin(gl, T.band[k] - off);
    gh = fmax(gh, T.band[k] + off);
  }
  s_val[tid] = gl; s_x[tid] = gh; s_z[tid & 127] = 0.0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s_val[tid] = fmin(s_val[tid], s_val[tid + o]); s_x[tid] = fmax(s_x[tid], s_x[tid + o]); }
    __syncthreads();
  }
  gl = s_val[0]; gh = s_x[0];
  __syncthreads();
  s_val[tid] = emax;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_val[tid] = fmax(s_val[tid], s_val[tid + o]);
    __syncthreads();
  }
  emax = s_val[0];
  __syncthreads();
  const double pivmin = 2.2250738585072014e-308 * fmax(1.0, emax);
  const double span = fmax(gh - gl, 1e-300);
  const double glo = gl - 1e-3 * span - 2.0 * pivmin, ghi = gh + 1e-3 * span + 2.0 * pivmin;

  double gmin = a.gamma_min[t];
  if (a.want_lambda_min || a.gamma_min_mode == 1) {
    const double lmin = band_kth_eigenvalue<B>(T, 0, glo, ghi, pivmin, s_cnt, s_x);
    if (a.lambda_min && tid == 0) a.lambda_min[t] = lmin;
    if (a.gamma_min_mode == 1) {
      gmin = (lmin + a.gamma_min_base < 0.0) ? -2.0 * lmin : a.gamma_min_base;
      if (tid == 0) a.gamma_min[t] = gmin;
    }
  }

  // argmin over the grid: first occurrence, NaN counts as the minimum (jnp.argmin)
  {
    const double* L = a.loss + (size_t)t * a.ngrid;
    double bv = 0.0; int bi = -1; bool bnan = false;
    for (int i = tid; i < a.ngrid; i += blockDim.x) {
      const double v = L[i];
      const bool vn = isnan(v);
      if (bi < 0) { bv = v; bi = i; bnan = vn; }
      else if (!bnan && (vn || v < bv)) { bv = v; bi = i; bnan = vn; }
    }
    s_val[tid] = bv; s_idx[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) {
        const double v1 = s_val[tid], v2 = s_val[tid + o];
        const int i1 = s_idx[tid], i2 = s_idx[tid + o];
        bool take2;
        if (i2 < 0) take2 = false;
        else if (i1 < 0) take2 = true;
        else {
          const bool n1 = isnan(v1), n2 = isnan(v2);
          if (n1 || n2) take2 = n2 && (!n1 || i2 < i1);
          else take2 = (v2 < v1) || (v2 == v1 && i2 < i1);
        }
        if (take2) { s_val[tid] = v2; s_idx[tid] = i2; }
      }
      __syncthreads();
    }
  }
  const int best = s_idx[0];
  __syncthreads();

  int st = 0;
  double gamma = 0.0;
  bool need_median = (best == 0);
  bool success = false;
  if (!need_median) {
    if (tid == 0) {
      double xo;
      const bool ok = bfgs_1d(BandEval<B>{T}, a.grid[best], xo);
      s_sc[0] = xo; s_sc[1] = ok ? 1.0 : 0.0;
    }
    __syncthreads();
    gamma = s_sc[0];
    success = s_sc[1] != 0.0;
    need_median = !success;
    __syncthreads();
  }
  if (need_median) {
    if (tid == 0) s_cnt[0] = band_inertia<B>(T, gmin, pivmin);
    __syncthreads();
    const int c0 = s_cnt[0];
    __syncthreads();
    const int mk = n - c0;
    double med;
    if (mk <= 0) {
      med = nan("");
    } else if (mk & 1) {
      med = band_kth_eigenvalue<B>(T, c0 + mk / 2, glo, ghi, pivmin, s_cnt, s_x);
    } else {
      const double m1 = band_kth_eigenvalue<B>(T, c0 + mk / 2 - 1, glo, ghi, pivmin, s_cnt, s_x);
      const double m2 = band_kth_eigenvalue<B>(T, c0 + mk / 2, glo, ghi, pivmin, s_cnt, s_x);
      med = 0.5 * (m1 + m2);
    }
    if (best == 0) {
      if (tid == 0) {
        double xo;
        const bool ok = bfgs_1d(BandEval<B>{T}, med, xo);
        s_sc[0] = ok ? xo : med; s_sc[1] = ok ? 1.0 : 0.0;
      }
      __syncthreads();
      gamma = s_sc[0];
      success = s_sc[1] != 0.0;
      __syncthreads();
    } else {
      gamma = med;
    }
    if (!success) st |= CHD_ST_USED_MEDIAN;
  }
  if (success) st |= CHD_ST_BFGS_SUCCESS;
  if (best == 0 || best == a.ngrid - 1) st |= CHD_ST_GRID_EDGE;

  // final factor with gamma_used = max(gamma, gamma_min), x = (T_b + gu)^-1 gt
  const double gu = isnan(gamma) ? gamma : fmax(gamma, gmin);
  double* Dg = a.Dg + (size_t)t * n;
  double* Ug = a.Ug + (size_t)t * n * B;
  double* xs = a.x + (size_t)t * n;
  if (tid == 0) {
    BandWin<0, B> W;
    W.init();
    const J<0> gj = jconst<0>(gu);
    J<0> ucol[B];
    for (int k = n - 1; k >= 0; --k) {
      double bc[B + 1];
#pragma unroll
      for (int tt = 0; tt <= B; ++tt) bc[tt] = (k >= tt) ? T.band[(size_t)tt * n + k] : 0.0;
      const J<0> dk = W.step(bc, gj, ucol);
      Dg[k] = dk.v;
#pragma unroll
      for (int tt = 0; tt < B; ++tt) Ug[(size_t)k * B + tt] = ucol[tt].v;   // U_{k-1-tt, k}
    }
    // U z = gt (z vanishes below row B), y = z / D, U^T x = y
    double zt[B];
#pragma unroll
    for (int k = B - 1; k >= 0; --k) {
      double zk = T.gt[k];
#pragma unroll
      for (int j = k + 1; j < B; ++j) zk -= Ug[(size_t)j * B + (j - k - 1)] * zt[j];
      zt[k] = zk;
    }
    double xw[B];   // last B entries of x (xw[0] = x_{k-1})
#pragma unroll
    for (int q = 0; q < B; ++q) xw[q] = 0.0;
    double ssq = 0.0, gx = 0.0;
    for (int k = 0; k < n; ++k) {
      double xk = (k < B) ? zt[k < B ? k : 0] / Dg[k] : 0.0;
#pragma unroll
      for (int tt = 0; tt < B; ++tt) xk -= Ug[(size_t)k * B + tt] * xw[tt];
      xs[k] = xk;
      ssq += xk * xk;
      if (k < B) gx += T.gt[k < B ? k : 0] * xk;
#pragma unroll
      for (int q = B - 1; q > 0; --q) xw[q] = xw[q - 1];
      xw[0] = xk;
    }
    a.noise[(size_t)t * a.out_stride] = gu * ssq / gx;
    a.gamma[(size_t)t * a.out_stride] = gamma;
    a.status[t] = st;
  }
  __syncthreads();

  // Z-test: z_j = (T_b + gu)^-1 b_j for the nz columns of Bq = Q^T S
  if (a.Bq && a.nz > 0) {
    double* Bt = a.Bq + (size_t)t * n * a.ldb;
    if (tid < a.nz) {
      const int j = tid;
      double zw[B];   // zw[m] = z_{k+1+m}
#pragma unroll
      for (int q = 0; q < B; ++q) zw[q] = 0.0;
      double den = 0.0;
      for (int k = n - 1; k >= 0; --k) {
        double zk = Bt[(size_t)k * a.ldb + j];
#pragma unroll
        for (int m = 0; m < B; ++m)
          if (k + 1 + m < n) zk -= Ug[(size_t)(k + 1 + m) * B + m] * zw[m];   // U_{k, k+1+m}
        const double yk = zk / Dg[k];
        den += zk * yk;
        Bt[(size_t)k * a.ldb + j] = yk;
#pragma unroll
        for (int q = B - 1; q > 0; --q) zw[q] = zw[q - 1];
        zw[0] = zk;
      }
      double xw[B];
#pragma unroll
      for (int q = 0; q < B; ++q) xw[q] = 0.0;
      double num = 0.0;
      for (int k = 0; k < n; ++k) {
        double xk = Bt[(size_t)k * a.ldb + j];
#pragma unroll
        for (int tt = 0; tt < B; ++tt) xk -= Ug[(size_t)k * B + tt] * xw[tt];
        num += xk * xk;
#pragma unroll
        for (int q = B - 1; q > 0; --q) xw[q] = xw[q - 1];
        xw[0] = xk;
      }
      s_z[j] = gu * num / den;
    }
    __syncthreads();
    if (tid < a.nz) {
      const double v = s_z[tid];
      const bool vn = isnan(v);
      int rank = 0;
      for (int i = 0; i < a.nz; ++i) {
        const double o = s_z[i];
        const bool on = isnan(o);
        bool less;
        if (on || vn) less = (!on && vn) || (on && vn && i < tid);
        else less = (o < v) || (o == v && i < tid);
        rank += less;
      }
      if (rank == (int)(0.05 * a.nz)) a.z_low[(size_t)t * a.out_stride] = v;
      if (rank == (int)(0.95 * a.nz)) a.z_high[(size_t)t * a.out_stride] = v;
    }
  }
}

template <int B>
static int gamma_band_launch_t(const Plan* pl, const double* band, const double* gt, const double* grid, double* loss,
                               BandFinishArgs a, int T, cudaStream_t st) {
  const int n = pl->n;
  gamma_band_grid_kernel<B><<<dim3((CHD_GRID_POINTS + GB_THREADS - 1) / GB_THREADS, T), GB_THREADS, 0, st>>>(
      band, gt, n, grid, CHD_GRID_POINTS, loss);
  CHD_LAUNCHED();
  gamma_band_finish_kernel<B><<<T, 256, 0, st>>>(a);
  CHD_LAUNCHED();
  return CHD_OK;
}

int gamma_band_launch(const Plan* pl, int B, const double* band, const double* gt, const double* grid, double* loss,
                      double* Bq, int ldb, int nz, int gamma_min_mode, double gamma_min_base, double* gamma_min,
                      int want_lambda_min, double* lambda_min, double* Dg, double* Ug, double* x, double* noise,
                      double* z_low, double* z_high, double* gamma, size_t out_stride, int32_t* status, int T,
                      cudaStream_t st) {
  BandFinishArgs a;
  a.band = band; a.gt = gt; a.loss = loss; a.grid = grid; a.ngrid = CHD_GRID_POINTS; a.Bq = Bq; a.ldb = ldb; a.nz = nz;
  a.gamma_min_mode = gamma_min_mode; a.gamma_min_base = gamma_min_base; a.gamma_min = gamma_min;
  a.want_lambda_min = want_lambda_min; a.lambda_min = lambda_min; a.Dg = Dg; a.Ug = Ug; a.x = x; a.noise = noise;
  a.z_low = z_low; a.z_high = z_high; a.gamma = gamma; a.out_stride = out_stride; a.status = status; a.n = pl->n;
  switch (B) {
    case 1: return gamma_band_launch_t<1>(pl, band, gt, grid, loss, a, T, st);
    case 2: return gamma_band_launch_t<2>(pl, band, gt, grid, loss, a, T, st);
    case 4: return gamma_band_launch_t<4>(pl, band, gt, grid, loss, a, T, st);
    default: set_error("gamma_band: unsupported half-bandwidth %d", B); return CHD_EINVAL;
  }
}

}  // namespace chd
