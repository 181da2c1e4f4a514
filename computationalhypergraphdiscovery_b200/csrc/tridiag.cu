// Batched Householder tridiagonalisation of the bordered matrix [0 ga^T; ga K]  (persistent,
// cooperative kernel; G CTAs per matrix, several matrices per launch) and the matching
// back-transformation.  Together with gamma.cu this replaces the reference's
// jnp.linalg.eigh(K) + V^T ga + V^T S + V c  (interpolatory.py:48-51, 76-80, 100): the path only
// consumes quantities that are functions of the tridiagonal T = Q^T K Q, beta0 e_0 = Q^T ga and
// B = Q^T S, so no eigenvector is ever formed.
//
// Algorithm: blocked dsytrd-style reduction (panel width nb): inside a panel the trailing matrix
// is not touched -- every column needs one symmetric matrix-vector product with it (HBM/L2 bound),
// corrected by the panel's V/W; after the panel a rank-2nb update A -= V W^T + W V^T runs on the
// FP64 tensor cores (DMMA m8n8k4).  The matrix is stored full (both triangles) so that every row
// slice is owned by exactly one CTA and the mat-vec needs no cross-CTA reduction.  The Z-test
// block B (n x 100) is transformed on the fly (B <- H_c B) inside the same three phases.
#include <utility>
#include <vector>

#include "internal.cuh"

namespace chd {

constexpr int TD_THREADS = 512;
constexpr int TD_RB = 16;         // rows per ownership chunk (= warps per CTA)
constexpr int TD_MAXNB = 32;      // panel width
constexpr int TD_BC = 128;        // max columns of B
constexpr int TD_TILE = 64;       // syr2k tile
constexpr int TD_TS = TD_TILE + 8;  // staging stride (== 8 mod 16: conflict-free fragment loads)
constexpr int TD_PART = 2 + 2 * TD_MAXNB + TD_BC;

struct TridiagArgs {
  double* K;
  size_t strideK;
  int ldk;
  const double* ga;  // T x n
  double* B;         // T x n x ldb or null
  int ldb, nbc;
  double* d;      // T x n
  double* e;      // T x n
  double* beta0;  // T
  double* tau;    // T x n     tau[c+1] for column c = -1..n-2
  double* v0;     // T x n     first reflector (column -1)
  double* u;      // T x n     scratch: updated column
  double* Vp;     // T x nb x n
  double* Wp;     // T x nb x n   final W
  double* Wraw;   // T x nb x n   W before the alpha correction
  double* part;   // T x G x TD_PART
  double* pgram;  // T x n x nb   pgram[(c+1)*nb + t] = v_c . v_t for t < k (same panel)
  unsigned* bar;  // T
  int n, G, nb, LR, nv;
};

struct GroupBarrier {
  unsigned* ctr;
  unsigned G;
  unsigned epoch;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += 1;
      __threadfence();
      atomicAdd(ctr, 1u);
      const unsigned target = epoch * G;
      unsigned long long spins = 0;
      while (true) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr));
        if (v >= target) break;
        if (++spins > (1ull << 34)) { asm volatile("trap;"); }
      }
      __threadfence();
    }
    __syncthreads();
  }
};

__device__ __forceinline__ int owner_of(int i, int G) { return (i / TD_RB) % G; }
__device__ __forceinline__ int global_row(int lr, int g, int G) {
  return ((lr / TD_RB) * G + g) * TD_RB + (lr % TD_RB);
}

__global__ void __launch_bounds__(TD_THREADS, 1) tridiag_kernel(TridiagArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int n = a.n, G = a.G, nb = a.nb, LR = a.LR, ldk = a.ldk;
  const int g = blockIdx.x, m = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  double* A = a.K + (size_t)m * a.strideK;
  const double* ga = a.ga + (size_t)m * n;
  double* Bm = a.B ? a.B + (size_t)m * n * a.ldb : nullptr;
  const int ldb = a.ldb, nbc = Bm ? a.nbc : 0;
  double* dd = a.d + (size_t)m * n;
  double* ee = a.e + (size_t)m * n;
  double* tauv = a.tau + (size_t)m * n;
  double* v0 = a.v0 + (size_t)m * n;
  double* ug = a.u + (size_t)m * n;
  double* Vp = a.Vp + (size_t)m * nb * n;
  double* Wp = a.Wp + (size_t)m * nb * n;
  double* Wraw = a.Wraw + (size_t)m * nb * n;
  double* part = a.part + (size_t)m * G * TD_PART;
  double* mypart = part + (size_t)g * TD_PART;
  double* pgram = a.pgram + (size_t)m * n * nb;

  // ---- shared memory carve-up
  const int stage_doubles = 2 * (2 * TD_MAXNB) * TD_TS;
  const int r0_doubles = max(a.nv + 2 * LR, stage_doubles);
  double* v = smem;                 // nv (aliased by the syr2k staging)
  double* ys = v + a.nv;            // LR
  double* vloc = ys + LR;           // LR
  double* Pa = smem;                // staging alias
  double* Pb = smem + (2 * TD_MAXNB) * TD_TS;
  double* Vs = smem + r0_doubles;   // nb x LR
  double* Ws = Vs + (size_t)nb * LR;
  double* cV = Ws + (size_t)nb * LR;  // 32
  double* cW = cV + TD_MAXNB;
  double* Vc = cW + TD_MAXNB;
  double* Wc = Vc + TD_MAXNB;
  double* cB = Wc + TD_MAXNB;       // 128
  double* Bred = cB + TD_BC;        // 4 x 128
  double* red = Bred + 4 * TD_BC;   // 40
  double* sc = red + 40;            // scalars: [0]=alpha_prev [1]=tau [2]=beta [3]=scale

  GroupBarrier bar{a.bar + m, (unsigned)G, 0u};

  for (int i = tid; i < a.nv; i += TD_THREADS) v[i] = 0.0;
  for (int i = tid; i < 2 * nb * LR; i += TD_THREADS) Vs[i] = 0.0;
  __syncthreads();

  int jj0 = -1;          // first column of the current panel
  double tau_prev = 0.0;

  for (int c = -1; c <= n - 2; ++c) {
    const int k = c - jj0;   // panel vectors already built
    const int R0 = c + 1;

    // ================= phase A: finish W_{k-1}, update column c, partial norm
    if (k > 0) {
      double s = 0.0;
      if (wid == 0) {
        for (int q = lane; q < G; q += 32) s += ldcg(part + (size_t)q * TD_PART + 1);
        s = warp_sum(s);
        if (lane == 0) sc[0] = -0.5 * tau_prev * s;
      }
      __syncthreads();
      const double alpha = sc[0];
      for (int lr = tid; lr < LR; lr += TD_THREADS) {
        const int i = global_row(lr, g, G);
        const double wf = Ws[(k - 1) * LR + lr] + alpha * Vs[(k - 1) * LR + lr];
        Ws[(k - 1) * LR + lr] = wf;
        if (i < n) Wp[(size_t)(k - 1) * n + i] = wf;
      }
      if (tid < k) {
        Vc[tid] = ldcg(Vp + (size_t)tid * n + c);
        Wc[tid] = (tid < k - 1) ? ldcg(Wp + (size_t)tid * n + c)
                                : ldcg(Wraw + (size_t)tid * n + c) + alpha * ldcg(Vp + (size_t)tid * n + c);
      }
    }
    __syncthreads();
    {
      double ss = 0.0;
      for (int lr = tid; lr < LR; lr += TD_THREADS) {
        const int i = global_row(lr, g, G);
        if (i >= R0 && i < n) {
          double x;
          if (c < 0) {
            x = ga[i];
          } else {
            x = ldcg(A + (size_t)c * ldk + i);
            for (int t = 0; t < k; ++t) x -= Vs[t * LR + lr] * Wc[t] + Ws[t * LR + lr] * Vc[t];
          }
          ug[i] = x;
          if (i > R0) ss += x * x;
        }
      }
      ss = block_sum(ss, red);
      if (tid == 0) {
        mypart[0] = ss;
        if (g == 0 && c >= 0) {
          double x = ldcg(A + (size_t)c * ldk + c);
          for (int t = 0; t < k; ++t) x -= 2.0 * Vc[t] * Wc[t];
          dd[c] = x;
        }
      }
    }
    bar.sync();

    // ================= phase B: reflector, symv, panel/B partial products
    {
      if (wid == 0) {
        double s = 0.0;
        for (int q = lane; q < G; q += 32) s += ldcg(part + (size_t)q * TD_PART + 0);
        s = warp_sum(s);
        if (lane == 0) {
          const double alpha0 = ldcg(ug + R0);
          double beta, tau, scale;
          if (s == 0.0) {
            beta = alpha0; tau = 0.0; scale = 0.0;
          } else {
            beta = -copysign(sqrt(alpha0 * alpha0 + s), alpha0);
            tau = (beta - alpha0) / beta;
            scale = 1.0 / (alpha0 - beta);
          }
          sc[1] = tau; sc[2] = beta; sc[3] = scale;
        }
      }
      __syncthreads();
      const double tau = sc[1], scale = sc[3];
      if (g == 0 && tid == 0) {
        if (c < 0) a.beta0[m] = sc[2]; else ee[c] = sc[2];
        tauv[c + 1] = tau;
      }
      for (int i = R0 + tid; i < n; i += TD_THREADS) v[i] = (i == R0) ? 1.0 : ldcg(ug + i) * scale;
      if (tid == 0 && R0 > 0) v[R0 - 1] = 0.0;
      __syncthreads();
      for (int lr = tid; lr < LR; lr += TD_THREADS) {
        const int i = global_row(lr, g, G);
        const double x = (i >= R0 && i < n) ? v[i] : 0.0;
        vloc[lr] = x;
        Vs[k * LR + lr] = x;
        if (i < n) {
          Vp[(size_t)k * n + i] = x;
          if (i >= R0) {
            if (c < 0) v0[i] = x; else A[(size_t)c * ldk + i] = x;
          }
        }
      }
      __syncthreads();
      // symv over the (not yet updated) trailing block: one warp per owned row
      const int jb = R0 & ~1;
      const int je = n & ~1;   // an odd last column is added by lane 0 (never touch the padding)
      for (int lr = wid; lr < LR; lr += TD_THREADS / 32) {
        const int i = global_row(lr, g, G);
        if (i < R0 || i >= n) continue;
        const double2* row = reinterpret_cast<const double2*>(A + (size_t)i * ldk);
        const double2* vv = reinterpret_cast<const double2*>(v);
        double s0 = 0.0, s1 = 0.0;
        int j2 = (jb >> 1) + lane;
        const int e2 = je >> 1;
        for (; j2 + 96 < e2; j2 += 128) {
          double2 x0 = ldcg2(row + j2), x1 = ldcg2(row + j2 + 32), x2 = ldcg2(row + j2 + 64),
                  x3 = ldcg2(row + j2 + 96);
          double2 y0 = vv[j2], y1 = vv[j2 + 32], y2 = vv[j2 + 64], y3 = vv[j2 + 96];
          s0 += x0.x * y0.x; s1 += x0.y * y0.y;
          s0 += x1.x * y1.x; s1 += x1.y * y1.y;
          s0 += x2.x * y2.x; s1 += x2.y * y2.y;
          s0 += x3.x * y3.x; s1 += x3.y * y3.y;
        }
        for (; j2 < e2; j2 += 32) {
          double2 x0 = ldcg2(row + j2);
          double2 y0 = vv[j2];
          s0 += x0.x * y0.x; s1 += x0.y * y0.y;
        }
        if ((n & 1) && lane == 0) s0 += ldcg(A + (size_t)i * ldk + (n - 1)) * v[n - 1];
        const double s = warp_sum(s0 + s1);
        if (lane == 0) ys[lr] = s;
      }
      // panel partial products  pV_t = V_t . v,  pW_t = W_t . v   (t < k)
      for (int task = wid; task < 2 * k; task += TD_THREADS / 32) {
        const double* src = (task < k) ? (Vs + (size_t)task * LR) : (Ws + (size_t)(task - k) * LR);
        double s = 0.0;
        for (int lr = lane; lr < LR; lr += 32) s += src[lr] * vloc[lr];
        s = warp_sum(s);
        if (lane == 0) mypart[2 + task] = s;
      }
      // B^T v partial
      if (nbc > 0) {
        const int col = tid & (TD_BC - 1), grp = tid >> 7;
        double s = 0.0;
        if (col < nbc) {
          for (int lr = grp; lr < LR; lr += 4) {
            const double x = vloc[lr];
            if (x != 0.0) {
              const int i = global_row(lr, g, G);
              s += ldcg(Bm + (size_t)i * ldb + col) * x;
            }
          }
        }
        Bred[grp * TD_BC + col] = s;
        __syncthreads();
        if (tid < nbc)
          mypart[2 + 2 * TD_MAXNB + tid] = Bred[tid] + Bred[TD_BC + tid] + Bred[2 * TD_BC + tid] + Bred[3 * TD_BC + tid];
      }
    }
    bar.sync();

    // ================= phase C: w' = tau (y - V (W^T v) - W (V^T v)), B update
    {
      const double tau = sc[1];
      const int nval = 2 * k + nbc;
      for (int q = wid; q < nval; q += TD_THREADS / 32) {
        const int slot = (q < 2 * k) ? (2 + q) : (2 + 2 * TD_MAXNB + (q - 2 * k));
        double s = 0.0;
        for (int h = lane; h < G; h += 32) s += ldcg(part + (size_t)h * TD_PART + slot);
        s = warp_sum(s);
        if (lane == 0) {
          if (q < k) cV[q] = s;
          else if (q < 2 * k) cW[q - k] = s;
          else cB[q - 2 * k] = s;
        }
      }
      __syncthreads();
      if (g == 0 && tid < k) pgram[(size_t)(c + 1) * nb + tid] = cV[tid];
      double dot = 0.0;
      for (int lr = tid; lr < LR; lr += TD_THREADS) {
        const int i = global_row(lr, g, G);
        double wv = 0.0;
        if (i >= R0 && i < n) {
          double y = ys[lr];
          for (int t = 0; t < k; ++t) y -= Vs[t * LR + lr] * cW[t] + Ws[t * LR + lr] * cV[t];
          wv = tau * y;
          dot += wv * vloc[lr];
        }
        Ws[k * LR + lr] = wv;
        if (i < n) Wraw[(size_t)k * n + i] = wv;
      }
      dot = block_sum(dot, red);
      if (tid == 0) mypart[1] = dot;
      if (nbc > 0 && tau != 0.0) {
        const int col = tid & (TD_BC - 1), grp = tid >> 7;
        if (col < nbc) {
          const double cb = tau * cB[col];
          for (int lr = grp; lr < LR; lr += 4) {
            const double x = vloc[lr];
            if (x != 0.0) {
              const int i = global_row(lr, g, G);
              double* p = Bm + (size_t)i * ldb + col;
              *p = ldcg(p) - x * cb;
            }
          }
        }
      }
      tau_prev = tau;
    }
    bar.sync();

    // ================= panel end: trailing update on the FP64 tensor cores
    if (k + 1 == nb || c == n - 2) {
      const int kp = k + 1;
      {  // finish W_{kp-1}
        if (wid == 0) {
          double s = 0.0;
          for (int q = lane; q < G; q += 32) s += ldcg(part + (size_t)q * TD_PART + 1);
          s = warp_sum(s);
          if (lane == 0) sc[0] = -0.5 * tau_prev * s;
        }
        __syncthreads();
        const double alpha = sc[0];
        for (int lr = tid; lr < LR; lr += TD_THREADS) {
          const int i = global_row(lr, g, G);
          const double wf = Ws[(kp - 1) * LR + lr] + alpha * Vs[(kp - 1) * LR + lr];
          if (i < n) Wp[(size_t)(kp - 1) * n + i] = wf;
        }
      }
      bar.sync();
      const int Rn = c + 1;   // trailing block is rows/cols > c
      const int t0 = Rn / TD_TILE, t1 = (n + TD_TILE - 1) / TD_TILE;
      const int nt = t1 - t0;
      const int kk_end = (2 * kp + 3) & ~3;
      const int wm = wid >> 2, wn = wid & 3;
      const int fr = lane >> 2, fk = lane & 3;
      for (int tile = g; tile < nt * nt; tile += G) {
        const int ti = t0 + tile / nt, tj = t0 + tile % nt;
        const int i0 = ti * TD_TILE, j0 = tj * TD_TILE;
        __syncthreads();
        for (int q = tid; q < kk_end * TD_TILE; q += TD_THREADS) {
          const int kk = q / TD_TILE, r = q % TD_TILE;
          double pa = 0.0, pb = 0.0;
          if (kk < 2 * kp) {
            const double* srcA = (kk < kp) ? (Vp + (size_t)kk * n) : (Wp + (size_t)(kk - kp) * n);
            const double* srcB = (kk < kp) ? (Wp + (size_t)kk * n) : (Vp + (size_t)(kk - kp) * n);
            if (i0 + r < n) pa = ldcg(srcA + i0 + r);
            if (j0 + r < n) pb = ldcg(srcB + j0 + r);
          }
          Pa[kk * TD_TS + r] = pa;
          Pb[kk * TD_TS + r] = pb;
        }
        __syncthreads();
        double acc[2][2][2] = {};
        for (int kk = 0; kk < kk_end; kk += 4) {
          double af[2], bf[2];
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            af[x] = Pa[(kk + fk) * TD_TS + wm * 16 + x * 8 + fr];
            bf[x] = Pb[(kk + fk) * TD_TS + wn * 16 + x * 8 + fr];
          }
#pragma unroll
          for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 2; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int i = i0 + wm * 16 + x * 8 + fr;
#pragma unroll
          for (int y = 0; y < 2; ++y) {
            const int j = j0 + wn * 16 + y * 8 + 2 * fk;
            if (i >= Rn && i < n) {
              double* p = A + (size_t)i * ldk + j;
              if (j >= Rn && j + 1 < n) {
                double2 o = ldcg2(reinterpret_cast<const double2*>(p));
                o.x -= acc[x][y][0];
                o.y -= acc[x][y][1];
                *reinterpret_cast<double2*>(p) = o;
              } else {
                if (j >= Rn && j < n) p[0] = ldcg(p) - acc[x][y][0];
                if (j + 1 >= Rn && j + 1 < n) p[1] = ldcg(p + 1) - acc[x][y][1];
              }
            }
          }
        }
      }
      __syncthreads();
      // staging aliased v/ys/vloc: restore the invariant v[i < R0_next] == 0 (rest is rewritten)
      for (int i = tid; i < a.nv; i += TD_THREADS) v[i] = 0.0;
      for (int i = tid; i < 2 * nb * LR; i += TD_THREADS) Vs[i] = 0.0;
      bar.sync();
      jj0 = c + 1;
    }
  }
  if (g == 0 && tid == 0) {
    dd[n - 1] = ldcg(A + (size_t)(n - 1) * ldk + (n - 1));
    ee[n - 1] = 0.0;
  }
}

// ------------------------------------------------------------------ back-transformation
// out = scale * Q x with Q = H_{-1} H_0 ... H_{n-2}: reflectors are applied last-to-first, one panel
// (nb reflectors) at a time: c_t = v_t . x in one pass, the short triangular recurrence with the
// stored panel Gram rows, then x -= sum_t s_t v_t in a second pass.  One CTA per matrix.
struct BackArgs {
  const double* K;
  size_t strideK;
  int ldk;
  const double* v0;
  const double* tau;
  const double* pgram;
  const double* x;      // T x n
  const double* beta0;  // T: out = -beta0 * Q x
  double* out;          // T x n (stride out_stride)
  size_t out_stride;
  int n, nb;
};

__global__ void __launch_bounds__(512) backtransform_kernel(BackArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int n = a.n, nb = a.nb, ldk = a.ldk;
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const double* A = a.K + (size_t)m * a.strideK;
  const double* v0 = a.v0 + (size_t)m * n;
  const double* tau = a.tau + (size_t)m * n;
  const double* pg = a.pgram + (size_t)m * n * nb;
  double* x = smem;            // n
  double* cc = x + n;          // nb
  double* ss = cc + TD_MAXNB;  // nb
  for (int i = tid; i < n; i += blockDim.x) x[i] = a.x[(size_t)m * n + i];
  __syncthreads();
  // columns c = -1..n-2 live in panels [-1, nb-2], [nb-1, 2nb-2], ...
  const int ncol = n;  // number of reflectors
  const int npanel = (ncol + nb - 1) / nb;
  for (int pnl = npanel - 1; pnl >= 0; --pnl) {
    const int cfirst = -1 + pnl * nb;
    const int kp = min(nb, ncol - pnl * nb);
    // c_t = v_t . x  (rows > c_t)
    for (int t = wid; t < kp; t += nw) {
      const int c = cfirst + t;
      const double* vr = (c < 0) ? v0 : (A + (size_t)c * ldk);
      double s = 0.0;
      for (int i = c + 1 + lane; i < n; i += 32) s += vr[i] * x[i];
      s = warp_sum(s);
      if (lane == 0) cc[t] = s;
    }
    __syncthreads();
    if (tid == 0) {
      for (int t = kp - 1; t >= 0; --t) {
        double s = cc[t];
        for (int t2 = t + 1; t2 < kp; ++t2) s -= pg[(size_t)(cfirst + t2 + 1) * nb + t] * ss[t2];
        ss[t] = tau[cfirst + t + 1] * s;
      }
    }
    __syncthreads();
    for (int i = cfirst + 1 + tid; i < n; i += blockDim.x) {
      double s = x[i];
      for (int t = 0; t < kp; ++t) {
        const int c = cfirst + t;
        if (i > c) {
          const double* vr = (c < 0) ? v0 : (A + (size_t)c * ldk);
          s -= ss[t] * vr[i];
        }
      }
      x[i] = s;
    }
    __syncthreads();
  }
  const double sc = -a.beta0[m];
  for (int i = tid; i < n; i += blockDim.x) a.out[(size_t)m * a.out_stride + i] = sc * x[i];
}

// ------------------------------------------------------------------ host side
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static size_t g_prof_used = 0;

int profile_enable(int on) {
  g_prof_on = on != 0;
  g_prof_used = 0;
  return CHD_OK;
}

int profile_read(double* total_ms, int* launches) {
  double tot = 0.0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    CHD_CUDA(cudaEventSynchronize(g_prof_events[i].second));
    float ms = 0.f;
    CHD_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i].first, g_prof_events[i].second));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int)g_prof_used;
  return CHD_OK;
}

size_t tridiag_smem_bytes(const Plan* pl) {
  const int chunks = (pl->n + TD_RB - 1) / TD_RB;
  const int LR = (chunks + pl->G - 1) / pl->G * TD_RB;
  const int nv = (int)align_up(pl->n + 2, 2);
  const int stage = 2 * (2 * TD_MAXNB) * TD_TS;
  const int r0 = max(nv + 2 * LR, stage);
  size_t doubles = (size_t)r0 + 2 * (size_t)pl->nb * LR + 4 * TD_MAXNB + TD_BC + 4 * TD_BC + 40 + 8;
  return doubles * sizeof(double);
}

size_t tridiag_workspace(const Plan* pl, int T) {
  const size_t n = pl->n, nb = pl->nb;
  size_t b = 0;
  b += 3 * align_up(T * n * sizeof(double), 256);             // tau, v0, u
  b += 3 * align_up(T * nb * n * sizeof(double), 256);        // Vp, Wp, Wraw
  b += align_up((size_t)T * pl->G * TD_PART * sizeof(double), 256);
  b += align_up(T * n * nb * sizeof(double), 256);            // pgram
  b += align_up(T * sizeof(unsigned), 256);
  return b + 4096;
}

bool tridiag_carve(const Plan* pl, int T, Carver& cv, TridiagBuffers& tb) {
  const size_t n = pl->n, nb = pl->nb;
  tb.tau = cv.take<double>(T * n);
  tb.v0 = cv.take<double>(T * n);
  tb.u = cv.take<double>(T * n);
  tb.Vp = cv.take<double>(T * nb * n);
  tb.Wp = cv.take<double>(T * nb * n);
  tb.Wraw = cv.take<double>(T * nb * n);
  tb.part = cv.take<double>((size_t)T * pl->G * TD_PART);
  tb.pgram = cv.take<double>(T * n * nb);
  tb.bar = cv.take<unsigned>(T);
  return cv.ok;
}

int tridiag_launch(const Plan* pl, double* K, int ldk, const double* ga, int T, double* B, int ldb, int nbc,
                   double* d, double* e, double* beta0, const TridiagBuffers& tb, cudaStream_t st) {
  const int n = pl->n;
  if (nbc > TD_BC) return CHD_EINVAL;
  const int chunks = (n + TD_RB - 1) / TD_RB;
  TridiagArgs a;
  a.K = K; a.strideK = (size_t)n * ldk; a.ldk = ldk; a.ga = ga; a.B = B; a.ldb = ldb; a.nbc = nbc;
  a.d = d; a.e = e; a.beta0 = beta0; a.tau = tb.tau; a.v0 = tb.v0; a.u = tb.u; a.Vp = tb.Vp; a.Wp = tb.Wp;
  a.Wraw = tb.Wraw; a.part = tb.part; a.pgram = tb.pgram; a.bar = tb.bar;
  a.n = n; a.G = pl->G; a.nb = pl->nb;
  a.LR = (chunks + pl->G - 1) / pl->G * TD_RB;
  a.nv = (int)align_up(n + 2, 2);
  const size_t smem = tridiag_smem_bytes(pl);
  static bool attr_set = false;
  if (!attr_set) {
    CHD_CUDA(cudaFuncSetAttribute(tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_optin));
    attr_set = true;
  }
  if (smem > pl->smem_optin) {
    set_error("tridiag: %zu B shared memory needed, %zu available (raise ctas_per_matrix)", smem, pl->smem_optin);
    return CHD_EINVAL;
  }
  CHD_CUDA(cudaMemsetAsync(tb.bar, 0, sizeof(unsigned) * T, st));
  for (int t0 = 0; t0 < T; t0 += pl->wave) {
    const int cnt = (T - t0 < pl->wave) ? (T - t0) : pl->wave;
    TridiagArgs b = a;
    b.K = K + (size_t)t0 * a.strideK;
    b.ga = ga + (size_t)t0 * n;
    b.B = B ? B + (size_t)t0 * n * ldb : nullptr;
    b.d = d + (size_t)t0 * n; b.e = e + (size_t)t0 * n; b.beta0 = beta0 + t0;
    b.tau = tb.tau + (size_t)t0 * n; b.v0 = tb.v0 + (size_t)t0 * n; b.u = tb.u + (size_t)t0 * n;
    b.Vp = tb.Vp + (size_t)t0 * pl->nb * n; b.Wp = tb.Wp + (size_t)t0 * pl->nb * n;
    b.Wraw = tb.Wraw + (size_t)t0 * pl->nb * n;
    b.part = tb.part + (size_t)t0 * pl->G * TD_PART;
    b.pgram = tb.pgram + (size_t)t0 * n * pl->nb;
    b.bar = tb.bar + t0;
    void* params[] = {&b};
    if (g_prof_on) {
      if (g_prof_used == g_prof_events.size()) {
        cudaEvent_t e0, e1;
        CHD_CUDA(cudaEventCreate(&e0));
        CHD_CUDA(cudaEventCreate(&e1));
        g_prof_events.emplace_back(e0, e1);
      }
      CHD_CUDA(cudaEventRecord(g_prof_events[g_prof_used].first, st));
    }
    CHD_CUDA(cudaLaunchCooperativeKernel((void*)tridiag_kernel, dim3(pl->G, cnt), dim3(TD_THREADS), params, smem, st));
    g_launches.fetch_add(1);
    if (g_prof_on) {
      CHD_CUDA(cudaEventRecord(g_prof_events[g_prof_used].second, st));
      ++g_prof_used;
    }
  }
  return CHD_OK;
}

int backtransform_launch(const Plan* pl, const double* K, int ldk, const TridiagBuffers& tb, const double* x,
                         const double* beta0, double* out, size_t out_stride, int T, cudaStream_t st) {
  BackArgs a;
  a.K = K; a.strideK = (size_t)pl->n * ldk; a.ldk = ldk; a.v0 = tb.v0; a.tau = tb.tau; a.pgram = tb.pgram;
  a.x = x; a.beta0 = beta0; a.out = out; a.out_stride = out_stride; a.n = pl->n; a.nb = pl->nb;
  const size_t smem = ((size_t)pl->n + 2 * TD_MAXNB + 8) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    CHD_CUDA(cudaFuncSetAttribute(backtransform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pl->smem_optin));
    attr_set = true;
  }
  backtransform_kernel<<<T, 512, smem, st>>>(a);
  CHD_LAUNCHED();
  return CHD_OK;
}

}  // namespace chd
