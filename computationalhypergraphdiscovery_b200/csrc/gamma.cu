// Gamma search, variational solve, noise ratio and Monte-Carlo Z-test on the tridiagonal form.
// Replaces find_gamma / solve_variationnal / Z_test (interpolatory.py:61-135) of the reference.
//
// With T = Q^T K Q (d, e), Q^T ga = beta0 e_0 and B = Q^T S (tridiag.cu):
//   eta_k(g)  = d_k + g - e_k^2 / eta_{k+1}(g)      (pivots of the UL factorisation of T + g I)
//   F(g)      = e_0^T (T + g)^-1 e_0 = 1 / eta_0
//   sum_i P_i^2 r_i   = beta0^2 g F,   sum_i P_i^2 r_i^2 = -beta0^2 g^2 F'     (r_i = g/(lambda_i+g), P = V^T ga)
//   loss(g)   = g eta_0' / eta_0                                    == interpolatory.py:124-126
//   x = (T+g)^-1 e_0:  yb = -beta0 Q x,  noise = g |x|^2 / x_0      == interpolatory.py:76-80
//   Z_j = g |z_j|^2 / (b_j . z_j),  z_j = (T+g)^-1 b_j              == interpolatory.py:98-105
//   median / smallest eigenvalue: Sturm-count multisection on T     == nanmedian interpolatory.py:122, eigvalsh
// The BFGS + strong-Wolfe line search is a scalar restatement of jax 0.4.28
// (jax/_src/scipy/optimize/{bfgs,line_search}.py), cf. oracle/bfgs.py.
#include "internal.cuh"
#include "bfgs.cuh"

namespace chd {

constexpr int GR_THREADS = 128;
constexpr int GR_CHUNK = 512;

// loss on the 10^4-point grid: one gamma per thread, (d, e) streamed through shared memory.
__global__ void __launch_bounds__(GR_THREADS) gamma_grid_kernel(const double* __restrict__ d,
                                                                 const double* __restrict__ e, int n,
                                                                 const double* __restrict__ grid, int ngrid,
                                                                 double* __restrict__ loss) {
  __shared__ double sd[GR_CHUNK], se[GR_CHUNK];
  const int t = blockIdx.y;
  const double* dt = d + (size_t)t * n;
  const double* et = e + (size_t)t * n;
  const int gi = blockIdx.x * GR_THREADS + threadIdx.x;
  const double g = (gi < ngrid) ? grid[gi] : 1.0;
  double eta = dt[n - 1] + g, eta1 = 1.0;
  // k runs n-2 .. 0 in chunks
  for (int hi = n - 2; hi >= 0; hi -= GR_CHUNK) {
    const int lo = max(hi - GR_CHUNK + 1, 0);
    const int cnt = hi - lo + 1;
    __syncthreads();
    for (int q = threadIdx.x; q < cnt; q += GR_THREADS) {
      sd[q] = dt[lo + q];
      se[q] = et[lo + q];
    }
    __syncthreads();
    for (int q = cnt - 1; q >= 0; --q) {
      const double ek = se[q];
      const double qq = ek / eta;
      eta1 = fma(qq * qq, eta1, 1.0);
      eta = (sd[q] + g) - ek * qq;
    }
  }
  if (gi < ngrid) loss[(size_t)t * ngrid + gi] = g * eta1 / eta;
}

// ------------------------------------------------------------------ scalar pieces (thread 0)
struct Tri {
  const double* d;
  const double* e;
  int n;
};

__device__ void loss_grad(const Tri& T, double g, double& f, double& fp) {
  const int n = T.n;
  double eta = T.d[n - 1] + g, eta1 = 1.0, eta2 = 0.0;
  for (int k = n - 2; k >= 0; --k) {
    const double ek = T.e[k];
    const double r = 1.0 / eta;
    const double q = ek * r;
    const double q2 = q * q;
    const double n2 = q2 * (eta2 - 2.0 * eta1 * eta1 * r);
    const double n1 = 1.0 + q2 * eta1;
    eta = (T.d[k] + g) - ek * q;
    eta1 = n1;
    eta2 = n2;
  }
  f = g * eta1 / eta;
  fp = eta1 / eta + g * eta2 / eta - g * eta1 * eta1 / (eta * eta);
}

struct TriEval {
  Tri T;
  __device__ void operator()(double g, double& f, double& fp) const { loss_grad(T, g, f, fp); }
};

// ------------------------------------------------------------------ Sturm multisection (whole block)
__device__ __forceinline__ int sturm_count(const Tri& T, double x, double pivmin) {
  int cnt = 0;
  double q = T.d[0] - x;
  if (fabs(q) <= pivmin) q = -pivmin;
  cnt += (q < 0.0);
  for (int k = 1; k < T.n; ++k) {
    const double ek = T.e[k - 1];
    q = (T.d[k] - x) - ek * ek / q;
    if (fabs(q) <= pivmin) q = -pivmin;
    cnt += (q < 0.0);
  }
  return cnt;
}

// k-th smallest eigenvalue (0-based) of T; all threads of the block participate.
__device__ double kth_eigenvalue(const Tri& T, int kth, double glo, double ghi, double pivmin, int* scnt,
                                 double* sx) {
  const int nt = blockDim.x, tid = threadIdx.x;
  double lo = glo, hi = ghi;  // count(lo) <= kth < count(hi)
  for (int it = 0; it < 64; ++it) {
    const double w = hi - lo;
    if (!(w > 4.0 * 2.220446049250313e-16 * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin)) break;
    const double x = lo + w * ((double)(tid + 1) / (double)(nt + 1));
    scnt[tid] = sturm_count(T, x, pivmin);
    sx[tid] = x;
    __syncthreads();
    // largest x_j with count <= kth becomes lo; smallest with count > kth becomes hi
    double nlo = lo, nhi = hi;
    if (scnt[0] > kth) {
      nhi = sx[0];
    } else if (scnt[nt - 1] <= kth) {
      nlo = sx[nt - 1];
    } else {
      // binary search for the crossing (counts are monotone in x)
      int a = 0, b = nt - 1;  // scnt[a] <= kth < scnt[b]
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

struct FinishArgs {
  const double* d;
  const double* e;
  const double* beta0;
  const double* loss;   // T x ngrid
  const double* grid;
  int ngrid;
  double* B;            // T x n x ldb  (Q^T S), overwritten
  int ldb, nz;
  int gamma_min_mode;
  double gamma_min_base;
  double* gamma_min;    // T in/out
  int want_lambda_min;
  double* lambda_min;   // T out (may be null)
  double* eta;          // T x n scratch
  double* uu;           // T x n scratch
  double* x;            // T x n out: (T + g)^-1 e_0
  double* noise; double* z_low; double* z_high; double* gamma;  // each [t * out_stride]
  size_t out_stride;
  int32_t* status;
  int n;
  double* gu_out;       // T or null.  Not null: two-stage path -- gamma_used is handed to bandsolve.cu, which does the
                        // solve and the Z-test in band coordinates; this kernel stops after the gamma search
};

__global__ void __launch_bounds__(256) gamma_finish_kernel(FinishArgs a) {
  __shared__ double s_val[256];
  __shared__ int s_idx[256];
  __shared__ int s_cnt[256];
  __shared__ double s_x[256];
  __shared__ double s_sc[8];
  __shared__ double s_z[128];
  const int t = blockIdx.x, tid = threadIdx.x, n = a.n;
  Tri T{a.d + (size_t)t * n, a.e + (size_t)t * n, n};

  // Gershgorin interval and pivmin
  double gl = 1e300, gh = -1e300, emax = 0.0;
  for (int k = tid; k < n; k += blockDim.x) {
    const double el = (k > 0) ? fabs(T.e[k - 1]) : 0.0, er = (k < n - 1) ? fabs(T.e[k]) : 0.0;
    gl = fmin(gl, T.d[k] - el - er);
    gh = fmax(gh, T.d[k] + el + er);
    emax = fmax(emax, er * er);
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

  // smallest eigenvalue and the gamma_min rule of stage 1 (MAIN:476-479)
  double gmin = a.gamma_min[t];
  if (a.want_lambda_min || a.gamma_min_mode == 1) {
    const double lmin = kth_eigenvalue(T, 0, glo, ghi, pivmin, s_cnt, s_x);
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

  // median of the eigenvalues >= gamma_min (nanmedian, REG:122) -- only where the reference uses it
  int st = 0;
  double gamma = 0.0;
  bool need_median = (best == 0);
  bool success = false;
  if (!need_median) {
    if (tid == 0) {
      double xo;
      const bool ok = bfgs_1d(TriEval{T}, a.grid[best], xo);
      s_sc[0] = xo; s_sc[1] = ok ? 1.0 : 0.0;
    }
    __syncthreads();
    gamma = s_sc[0];
    success = s_sc[1] != 0.0;
    need_median = !success;
    __syncthreads();
  }
  if (need_median) {
    // number of eigenvalues strictly below gamma_min
    if (tid == 0) s_cnt[0] = sturm_count(T, gmin, pivmin);
    __syncthreads();
    const int c0 = s_cnt[0];
    __syncthreads();
    const int mk = n - c0;
    double med;
    if (mk <= 0) {
      med = nan("");
    } else if (mk & 1) {
      med = kth_eigenvalue(T, c0 + mk / 2, glo, ghi, pivmin, s_cnt, s_x);
    } else {
      const double m1 = kth_eigenvalue(T, c0 + mk / 2 - 1, glo, ghi, pivmin, s_cnt, s_x);
      const double m2 = kth_eigenvalue(T, c0 + mk / 2, glo, ghi, pivmin, s_cnt, s_x);
      med = 0.5 * (m1 + m2);
    }
    if (best == 0) {
      // BFGS from the median (REG:132-134); on failure the median itself
      if (tid == 0) {
        double xo;
        const bool ok = bfgs_1d(TriEval{T}, med, xo);
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

  // final solve with gamma_used = max(gamma, gamma_min)  (REG:52, 76-80)
  const double gu = isnan(gamma) ? gamma : fmax(gamma, gmin);
  if (a.gu_out) {
    if (tid == 0) {
      a.gu_out[t] = gu;
      a.gamma[(size_t)t * a.out_stride] = gamma;
      a.status[t] = st;
    }
    return;
  }
  double* eta = a.eta + (size_t)t * n;
  double* uu = a.uu + (size_t)t * n;
  double* xs = a.x + (size_t)t * n;
  if (tid == 0) {
    double et = T.d[n - 1] + gu;
    eta[n - 1] = et;
    for (int k = n - 2; k >= 0; --k) {
      const double u = T.e[k] / et;
      uu[k] = u;
      et = (T.d[k] + gu) - T.e[k] * u;
      eta[k] = et;
    }
    double xk = 1.0 / et, ssq = 0.0;
    const double x0 = xk;
    xs[0] = xk;
    ssq = xk * xk;
    for (int k = 1; k < n; ++k) {
      xk = -uu[k - 1] * xk;
      xs[k] = xk;
      ssq += xk * xk;
    }
    a.noise[(size_t)t * a.out_stride] = gu * ssq / x0;
    a.gamma[(size_t)t * a.out_stride] = gamma;
    a.status[t] = st;
  }
  __syncthreads();
  __threadfence_block();

  // Z-test: z_j = (T + gu)^-1 b_j for the nz columns of B = Q^T S
  if (a.B && a.nz > 0) {
    double* Bt = a.B + (size_t)t * n * a.ldb;
    if (tid < a.nz) {
      const int j = tid;
      double tk = Bt[(size_t)(n - 1) * a.ldb + j];
      double den = tk * tk / eta[n - 1];
      Bt[(size_t)(n - 1) * a.ldb + j] = tk / eta[n - 1];
      for (int k = n - 2; k >= 0; --k) {
        tk = Bt[(size_t)k * a.ldb + j] - uu[k] * tk;
        const double sk = tk / eta[k];
        den += tk * sk;
        Bt[(size_t)k * a.ldb + j] = sk;
      }
      double z = Bt[j];
      double num = z * z;
      for (int k = 1; k < n; ++k) {
        z = Bt[(size_t)k * a.ldb + j] - uu[k - 1] * z;
        num += z * z;
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

int gamma_launch(const Plan* pl, const double* d, const double* e, const double* beta0, const double* grid,
                 double* loss, double* B, int ldb, int nz, int gamma_min_mode, double gamma_min_base,
                 double* gamma_min, int want_lambda_min, double* lambda_min, double* eta, double* uu, double* x,
                 double* noise, double* z_low, double* z_high, double* gamma, size_t out_stride, int32_t* status,
                 int T, cudaStream_t st, double* gu_out) {
  const int n = pl->n;
  gamma_grid_kernel<<<dim3((CHD_GRID_POINTS + GR_THREADS - 1) / GR_THREADS, T), GR_THREADS, 0, st>>>(
      d, e, n, grid, CHD_GRID_POINTS, loss);
  CHD_LAUNCHED();
  FinishArgs a;
  a.d = d; a.e = e; a.beta0 = beta0; a.loss = loss; a.grid = grid; a.ngrid = CHD_GRID_POINTS;
  a.B = B; a.ldb = ldb; a.nz = nz; a.gamma_min_mode = gamma_min_mode; a.gamma_min_base = gamma_min_base;
  a.gamma_min = gamma_min; a.want_lambda_min = want_lambda_min; a.lambda_min = lambda_min;
  a.eta = eta; a.uu = uu; a.x = x; a.noise = noise; a.z_low = z_low; a.z_high = z_high; a.gamma = gamma;
  a.out_stride = out_stride; a.status = status; a.n = n; a.gu_out = gu_out;
  gamma_finish_kernel<<<T, 256, 0, st>>>(a);
  CHD_LAUNCHED();
  return CHD_OK;
}

}  // namespace chd
