// Variational solve, noise ratio and Monte-Carlo Z-test in BAND coordinates (two-stage path):
//   T_b + g I = L D L^T   (banded, half-width NB, no pivoting, no square roots: negative pivots pass through
//                          like in the reference's eigen-formulation)
//   x   = (T_b + g)^-1 e_0:   yb = -beta0 Q x (band.cu),  noise = g |x|^2 / x_0      == interpolatory.py:76-80
//   Z_j = g |z_j|^2 / (b_j . z_j),  z_j = (T_b + g)^-1 b_j,  b_j = Q^T s_j             == interpolatory.py:98-105
// One CTA per matrix.  Factorisation: left-looking by columns, lane = row offset inside the column, the last NB
// columns of L in a shared-memory ring.  Solves: lane = right-hand side (100 Z samples + e_0), the last NB
// solution values of every right-hand side in registers.  NumPy model: tests/twostage_model.py.
#include "internal.cuh"

namespace chd {

constexpr int SB = CHD_BAND;
constexpr int SLD = CHD_BAND_LD;
constexpr int BS_THREADS = 128;
constexpr int BS_RING = SB + 1;        // columns kept in the ring
constexpr int BS_RS = SB + 2;          // ring row stride

struct BandSolveArgs {
  const double* bandS;   // T x n x SLD   bandS[i][d] = T_b(i, i-d)
  double* Lg;            // T x n x SLD   L(i, i-d) (d >= 1), Lg[i][0] = D(i)
  const double* gu;      // T  gamma_used
  const double* gamma;   // T  raw gamma (only copied through)
  double* Bq;            // T x n x ldb   in: Q^T S, overwritten
  int ldb, nz;
  double* x;             // T x n  out: (T_b + g)^-1 e_0
  double* noise; double* z_low; double* z_high;
  size_t out_stride;
  int n;
};

__global__ void __launch_bounds__(BS_THREADS) band_solve_kernel(BandSolveArgs a) {
  __shared__ double ring[BS_RING * BS_RS];     // ring[(col % BS_RING)][d] = L(col + d, col), d = 0 -> D(col)
  __shared__ double cs[SB + 1];
  __shared__ double Ls[2 * SB * SLD];          // staged rows of Lg for the solves
  __shared__ double s_z[128];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = a.n;
  const double* bs = a.bandS + (size_t)t * n * SLD;
  double* Lg = a.Lg + (size_t)t * n * SLD;
  const double g = a.gu[t];

  // ---------------- factorisation (warp 0)
  if (wid == 0) {
    for (int e = lane; e < BS_RING * BS_RS; e += 32) ring[e] = 0.0;
    __syncwarp();
    for (int k = 0; k < n; ++k) {
      const int km = k % BS_RING;
      // c_p = D(k-p) L(k, k-p), p = 1..NB  (lane p-1)
      {
        const int p = lane + 1;
        double c = 0.0;
        if (k - p >= 0) {
          int sl = km - p;
          sl += (sl >> 31) & BS_RING;
          const double* col = ring + sl * BS_RS;
          c = col[0] * col[p];
        }
        cs[p] = c;
      }
      // lanes d = 0..31 own L(k+d, k); d = 32 is done by lane 0 afterwards (single term-free entry)
      double s = 0.0, s32 = 0.0;
      {
        const int d = lane;
        if (k + d < n) s = bs[(size_t)(k + d) * SLD + d] + ((d == 0) ? g : 0.0);
        if (lane == 0 && k + SB < n) s32 = bs[(size_t)(k + SB) * SLD + SB];
      }
      __syncwarp();
      {
        const int d = lane;
        // sum over previous columns k-p that reach row k+d: d + p <= NB
        double s2 = 0.0;
        for (int p = min(SB - d, k); p >= 1; --p) {
          int sl = km - p;
          sl += (sl >> 31) & BS_RING;
          const double term = ring[sl * BS_RS + d + p] * cs[p];
          if (p & 1) s -= term; else s2 -= term;
        }
        s += s2;
      }
      const double Dk = __shfl_sync(0xffffffffu, s, 0);
      const double inv = 1.0 / Dk;
      const double lv = (lane == 0) ? Dk : s * inv;
      double* colk = ring + km * BS_RS;
      colk[lane] = lv;
      if (lane == 0) colk[SB] = s32 * inv;
      // row-format copy for the solves
      if (k + lane < n) Lg[(size_t)(k + lane) * SLD + lane] = lv;
      if (lane == 0 && k + SB < n) Lg[(size_t)(k + SB) * SLD + SB] = s32 * inv;
      __syncwarp();
    }
  }
  __syncthreads();
  __threadfence_block();

  // ---------------- solves: thread z < nz -> Z sample z, thread nz -> e_0
  const int nz = a.nz;
  double* Bt = a.Bq ? a.Bq + (size_t)t * n * a.ldb : nullptr;
  double* xs = a.x + (size_t)t * n;
  const bool isz = Bt && tid < nz;
  const bool ise = tid == nz;
  const bool act = isz || ise;
  double hist[SB];
#pragma unroll
  for (int q = 0; q < SB; ++q) hist[q] = 0.0;
  double den = 0.0, num = 0.0;
  // forward: y[i] = b[i] - sum_p L(i, i-p) y[i-p];  den = sum y^2 / D
  for (int i0 = 0; i0 < n; i0 += SB) {
    __syncthreads();
    for (int e = tid; e < SB * SLD; e += BS_THREADS) {
      const int r = e / SLD, d = e % SLD;
      Ls[e] = (i0 + r < n && d <= SB && d <= i0 + r) ? Lg[(size_t)(i0 + r) * SLD + d] : 0.0;
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int r = 0; r < SB; ++r) {
        const int i = i0 + r;
        if (i < n) {
          double y = isz ? Bt[(size_t)i * a.ldb + tid] : ((i == 0) ? 1.0 : 0.0);
          const double* Lr = Ls + r * SLD;
          // oldest terms first, the value of the previous row last (shortest dependency chain)
#pragma unroll
          for (int p = SB; p >= 1; --p) y = fma(-Lr[p], hist[(r - p + 2 * SB) % SB], y);
          hist[r] = y;
          den = fma(y * y, 1.0 / Lr[0], den);
          if (isz) Bt[(size_t)i * a.ldb + tid] = y; else xs[i] = y;
        }
      }
    }
  }
  // backward: z[i] = y[i] / D(i) - sum_p L(i+p, i) z[i+p];  num = sum z^2
#pragma unroll
  for (int q = 0; q < SB; ++q) hist[q] = 0.0;
  const int nblk = (n + SB - 1) / SB;
  double x0 = 0.0;
  for (int b = nblk - 1; b >= 0; --b) {
    const int i0 = b * SB;
    __syncthreads();
    for (int e = tid; e < 2 * SB * SLD; e += BS_THREADS) {
      const int r = e / SLD, d = e % SLD;
      Ls[e] = (i0 + r < n && d <= SB && d <= i0 + r) ? Lg[(size_t)(i0 + r) * SLD + d] : 0.0;
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int r = SB - 1; r >= 0; --r) {
        const int i = i0 + r;
        if (i < n) {
          const double y = isz ? Bt[(size_t)i * a.ldb + tid] : xs[i];
          double z = y / Ls[r * SLD];
#pragma unroll
          for (int p = SB; p >= 1; --p) z = fma(-Ls[(r + p) * SLD + p], hist[(r + p) % SB], z);
          hist[r] = z;
          num = fma(z, z, num);
          if (ise) { xs[i] = z; if (i == 0) x0 = z; }
        }
      }
    }
  }
  if (ise) {
    a.noise[(size_t)t * a.out_stride] = g * num / x0;
  }
  if (Bt && nz > 0) {
    __syncthreads();
    if (tid < 128) s_z[tid] = isz ? g * num / den : 0.0;
    __syncthreads();
    if (isz) {
      const double v = s_z[tid];
      const bool vn = isnan(v);
      int rank = 0;
      for (int i = 0; i < nz; ++i) {
        const double o = s_z[i];
        const bool on = isnan(o);
        bool less;
        if (on || vn) less = (!on && vn) || (on && vn && i < tid);
        else less = (o < v) || (o == v && i < tid);
        rank += less;
      }
      if (rank == (int)(0.05 * nz)) a.z_low[(size_t)t * a.out_stride] = v;
      if (rank == (int)(0.95 * nz)) a.z_high[(size_t)t * a.out_stride] = v;
    }
  }
}

int band_solve_launch(const Plan* pl, const double* bandS, double* Lg, const double* gu, double* Bq, int ldb, int nz,
                      double* x, double* noise, double* z_low, double* z_high, size_t out_stride, int T,
                      cudaStream_t st) {
  if (nz >= BS_THREADS) return CHD_EINVAL;
  BandSolveArgs a;
  a.bandS = bandS; a.Lg = Lg; a.gu = gu; a.gamma = nullptr; a.Bq = Bq; a.ldb = ldb; a.nz = Bq ? nz : 0; a.x = x;
  a.noise = noise; a.z_low = z_low; a.z_high = z_high; a.out_stride = out_stride; a.n = pl->n;
  int rc;
  if ((rc = profile_mark_begin(PROF_SOLVE, st))) return rc;
  band_solve_kernel<<<T, BS_THREADS, 0, st>>>(a);
  CHD_LAUNCHED();
  if ((rc = profile_mark_end(PROF_SOLVE, st))) return rc;
  return CHD_OK;
}

}  // namespace chd
