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
#include <stdlib.h>
#include <utility>
#include <vector>

#include "internal.cuh"

namespace chd {

constexpr int TD_RB = 16;         // rows per ownership chunk (= warps per CTA)
constexpr int TD_MAXNB = 32;      // panel width
constexpr int TD_BC = 128;        // max columns of B
constexpr int TD_TN = 128;        // syr2k tile: (THREADS / 4) rows x TD_TN cols, warps of 32 x 32
constexpr int TD_KC = 16;         // K chunk staged per pass
constexpr int TD_NS = 2;          // stages of the symv tile ring
constexpr int TD_SB = TD_TN + 8;  // staging stride (== 8 mod 16: conflict-free fragment loads)
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
  double* rsum;   // T x n     scratch: row part of the symv (owner rows)
  double* Vp;     // T x nb x n
  double* Wp;     // T x nb x n
  double* part;   // T x G x TD_PART
  double* colpart;  // T x G x n  column part of the symv, per CTA
  double* pgram;  // T x n x nb   pgram[(c+1)*nb + t] = v_c . v_t for t < k (same panel)
  unsigned* bar;  // T
  unsigned long long* timing;  // optional (debug): 8 accumulated phase times in ns, CTA 0 of matrix 0
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
      // release-arrive / acquire-poll on a monotonically increasing counter (no reset needed)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      const unsigned target = epoch * G;
      unsigned long long spins = 0;
      while (true) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) break;
        if (++spins > (1ull << 34)) { asm volatile("trap;"); }
      }
    }
    __syncthreads();
  }
};

// ---- bulk asynchronous copy (TMA engine, non-tensor form) + mbarrier helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

// Row chunks of TD_RB rows are dealt to the G CTAs in zigzag order (0..G-1, G-1..0, ...) so that the
// triangular row lengths balance across CTAs.
__device__ __forceinline__ int chunk_row0(int ch, int g, int G) {
  const int gg = (ch & 1) ? (G - 1 - g) : g;
  return (ch * G + gg) * TD_RB;
}
__device__ __forceinline__ int global_row(int lr, int g, int G) {
  return chunk_row0(lr / TD_RB, g, G) + (lr % TD_RB);
}

// Lower-triangle storage: only A[i][j], j <= i, of the trailing block is kept up to date; the strictly upper
// part of row c receives the reflector of column c.
// THREADS = 256: two CTAs (of different matrices) per SM, one streams while the other sits in a barrier;
// THREADS = 512: one CTA per SM (large n, where the per-CTA shared memory does not allow two).
template <int TD_THREADS, bool TD_BULK>
__global__ void __launch_bounds__(TD_THREADS, 512 / TD_THREADS) tridiag_kernel(TridiagArgs a) {
  constexpr int TD_TM = TD_THREADS / 4;          // syr2k tile rows
  constexpr int TD_SA = TD_TM + 8;
  constexpr int TD_NGRP = TD_THREADS / 128;      // row groups of the B-block loops
  constexpr int TD_NH = TD_THREADS / 256;
  extern __shared__ __align__(16) double smem[];
  const int n = a.n, G = a.G, nb = a.nb, LR = a.LR, ldk = a.ldk, nv = a.nv;
  const int g = blockIdx.x, m = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = TD_THREADS / 32;

  double* A = a.K + (size_t)m * a.strideK;
  const double* ga = a.ga + (size_t)m * n;
  double* Bm = a.B ? a.B + (size_t)m * n * a.ldb : nullptr;
  const int ldb = a.ldb, nbc = Bm ? a.nbc : 0;
  double* dd = a.d + (size_t)m * n;
  double* ee = a.e + (size_t)m * n;
  double* tauv = a.tau + (size_t)m * n;
  double* v0 = a.v0 + (size_t)m * n;
  double* ug = a.u + (size_t)m * n;
  double* rsum = a.rsum + (size_t)m * n;
  double* Vp = a.Vp + (size_t)m * nb * n;
  double* Wp = a.Wp + (size_t)m * nb * n;
  double* part = a.part + (size_t)m * G * TD_PART;
  double* mypart = part + (size_t)g * TD_PART;
  double* colpart = a.colpart + (size_t)m * G * n;
  double* mycol = colpart + (size_t)g * n;
  double* pgram = a.pgram + (size_t)m * n * nb;

  // ---- shared memory carve-up.  The reflector v is never cached whole: v_j = ug[j] * scale is recomputed
  // from the updated column in L2 where needed, which keeps the footprint small enough for two CTAs per SM.
  double* ys = smem;                         // LR
  double* vloc = ys + LR;                    // LR   v at the owned rows
  double* rowpart = vloc + LR;               // NW x LR
  double* Vs = rowpart + (size_t)NW * LR;    // nb x LR
  double* Ws = Vs + (size_t)nb * LR;
  double* cV = Ws + (size_t)nb * LR;         // 32 each
  double* cW = cV + TD_MAXNB;
  double* Vc = cW + TD_MAXNB;
  double* Wc = Vc + TD_MAXNB;
  double* cB = Wc + TD_MAXNB;                // 128
  double* Bred = cB + TD_BC;                 // 512
  double* red = Bred + 512;                  // 40
  double* sc = red + 40;                     // [0]=alpha [1]=tau [2]=beta [3]=scale [4]=vAv
  uint64_t* fullbar = reinterpret_cast<uint64_t*>(sc + 8);   // TD_NS mbarriers
  double* ring = sc + 8 + 8;                 // TD_NS x 16 x (32 NW) symv tiles; aliased by the syr2k staging
  double* Pa = ring;
  double* Pb = Pa + TD_KC * TD_SA;
  double* ycol = mycol;                      // column part of the symv lives in this CTA's global (L2) slice

  GroupBarrier bar{a.bar + m, (unsigned)G, 0u};
  int ringq = 0;
  if (TD_BULK && tid == 0) {
    for (int u = 0; u < TD_NS; ++u) mbar_init(&fullbar[u], 1);
    fence_mbar_init();
  }
  (void)ringq;

  for (int i = tid; i < (2 + NW + 2 * nb) * LR; i += TD_THREADS) smem[i] = 0.0;
  __syncthreads();

  // ---- prologue: "column -1" is ga
  {
    double ss = 0.0;
    for (int lr = tid; lr < LR; lr += TD_THREADS) {
      const int i = global_row(lr, g, G);
      if (i < n) {
        const double x = ga[i];
        ug[i] = x;
        if (i > 0) ss += x * x;
      }
    }
    ss = block_sum(ss, red);
    if (tid == 0) mypart[0] = ss;
  }
  bar.sync();

  int jj0 = -1;  // first column of the current panel
  const bool timed = a.timing && g == 0 && m == 0 && tid == 0;
  unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
#define TD_TICK(slot)                                              \
  if (timed) {                                                     \
    unsigned long long now_;                                       \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));      \
    tacc[slot] += now_ - tprev;                                    \
    tprev = now_;                                                  \
  }
  if (timed) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tprev));

  for (int c = -1; c <= n - 2; ++c) {
    const int k = c - jj0;  // panel vectors already built
    const int R0 = c + 1;

    // ================= phase B: reflector, symv (lower triangle), partial products
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
    const double tau = sc[1];
    {
      const double scale = sc[3];
      if (g == 0 && tid == 0) {
        if (c < 0) a.beta0[m] = sc[2]; else ee[c] = sc[2];
        tauv[c + 1] = tau;
      }
      for (int i = R0 + tid; i < n; i += TD_THREADS) ycol[i] = 0.0;
    }
    for (int lr = tid; lr < LR; lr += TD_THREADS) {
      const int i = global_row(lr, g, G);
      const double x = (i < R0 || i >= n) ? 0.0 : ((i == R0) ? 1.0 : ldcg(ug + i) * sc[3]);
      vloc[lr] = x;
      Vs[k * LR + lr] = x;
      if (i < n) {
        Vp[(size_t)k * n + i] = x;
        if (i >= R0) {
          if (c < 0) v0[i] = x; else A[(size_t)c * ldk + i] = x;
        }
      }
    }
    TD_TICK(0)
    __syncthreads();
    double vav = 0.0;   // this lane's share of sum_j v_j * (column part of y)_j
    // ---- symv on the stored lower triangle, two interchangeable data paths:
    //  TD_BULK : tiles of 16 rows x (32 NW) columns are streamed into a shared-memory ring by bulk asynchronous
    //            copies (cp.async.bulk + mbarrier complete_tx; SASS: UBLKCP), one 16 x 32 sub-block per warp;
    //  default : every warp loads its 16 x 32 sub-blocks straight into registers (16 independent loads per
    //            lane), no block-level synchronisation inside the product.
    if constexpr (TD_BULK)
    {
      constexpr int TW = 32 * NW;
      const int jbase = R0 & ~31;
      const double scale_ = sc[3];
      const int nch = LR / TD_RB;
      auto tiles_in = [&](int ch) -> int {
        const int i0 = chunk_row0(ch, g, G);
        if (i0 + TD_RB - 1 < R0 || i0 >= n) return 0;
        return (min(i0 + TD_RB - 1, n - 1) - jbase) / TW + 1;
      };
      int ich = 0, itt = 0, iq = ringq;   // next tile to issue (thread 0); ringq = tiles consumed so far
      auto issue_next = [&]() {
        while (ich < nch && itt >= tiles_in(ich)) { ++ich; itt = 0; }
        if (ich >= nch) return;
        const int i0 = chunk_row0(ich, g, G);
        const int jmax = min(i0 + TD_RB - 1, n - 1);
        const int c_lo = jbase + TW * itt;
        const int ncols = (min(TW, jmax + 1 - c_lo) + 1) & ~1;
        const int r_lo = max(i0, R0) - i0, r_hi = min(i0 + TD_RB, n) - i0;
        const int st = iq % TD_NS;
        const uint32_t bytes_row = (uint32_t)ncols * 8u;
        fence_proxy_async();
        mbar_expect_tx(&fullbar[st], bytes_row * (uint32_t)(r_hi - r_lo));
        for (int r = r_lo; r < r_hi; ++r)
          bulk_g2s(ring + ((size_t)st * TD_RB + r) * TW, A + (size_t)(i0 + r) * ldk + c_lo, bytes_row, &fullbar[st]);
        ++itt;
        ++iq;
      };
      if (tid == 0)
        for (int u = 0; u < TD_NS; ++u) issue_next();
      int q = ringq;   // tiles consumed so far by this CTA (mbarrier phase bookkeeping)
      for (int ch = 0; ch < nch; ++ch) {
        const int nt = tiles_in(ch);
        if (nt == 0) continue;
        const int i0 = chunk_row0(ch, g, G);
        const int jmax = min(i0 + TD_RB - 1, n - 1);
        double racc[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) racc[r] = 0.0;
        for (int tt = 0; tt < nt; ++tt, ++q) {
          const int st = q % TD_NS;
          const uint32_t parity = (uint32_t)((q / TD_NS) & 1);
          const int j = jbase + TW * tt + 32 * wid + lane;
          const bool jok = (j >= R0) && (j <= jmax);
          double vj = 0.0, yold = 0.0;
          if (jok) {
            vj = (j == R0) ? 1.0 : ldcg(ug + j) * scale_;
            yold = ldcg(ycol + j);
          }
          mbar_wait(&fullbar[st], parity);
          if (jbase + TW * tt + 32 * wid <= jmax) {   // warp-uniform: this sub-block has columns to process
            const double* tp = ring + (size_t)st * TD_RB * TW + 32 * wid + lane;
            double colacc = 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const int i = i0 + r;
              const bool ok = jok && (i >= R0) && (i < n) && (j <= i);
              const double av = ok ? tp[r * TW] : 0.0;
              racc[r] = fma(av, vj, racc[r]);
              if (j < i) colacc = fma(av, vloc[ch * TD_RB + r], colacc);
            }
            if (jok) {
              ycol[j] = yold + colacc;
              vav = fma(vj, colacc, vav);
            }
          }
          __syncthreads();
          if (tid == 0) issue_next();
        }
        // butterfly reduction of the 16 row accumulators over the warp (16 shuffles)
        double t8[8], t4[4], t2[2], t1;
        {
          const bool hi = lane & 16;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const double send = hi ? racc[r] : racc[r + 8];
            const double keep = hi ? racc[r + 8] : racc[r];
            t8[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool hi = lane & 8;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const double send = hi ? t8[r] : t8[r + 4];
            const double keep = hi ? t8[r + 4] : t8[r];
            t4[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        {
          const bool hi = lane & 4;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const double send = hi ? t4[r] : t4[r + 2];
            const double keep = hi ? t4[r + 2] : t4[r];
            t2[r] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
        {
          const bool hi = lane & 2;
          const double send = hi ? t2[0] : t2[1];
          const double keep = hi ? t2[1] : t2[0];
          t1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
        if ((lane & 1) == 0) {
          const int r = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
          rowpart[wid * LR + ch * TD_RB + r] = t1;
        }
      }
      ringq = q;
    }
    else
    {
      const int jbase = R0 & ~31;
      const double scale_ = sc[3];
      for (int ch = 0; ch < LR / TD_RB; ++ch) {
        const int i0 = chunk_row0(ch, g, G);
        if (i0 + TD_RB - 1 < R0 || i0 >= n) continue;
        double racc[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) racc[r] = 0.0;
        const int jmax = min(i0 + TD_RB - 1, n - 1);
        const int nstrips = (jmax - jbase) / 32 + 1;
        for (int s = wid; s < nstrips; s += NW) {
          const int js = jbase + 32 * s;
          const int j = js + lane;
          double colacc = 0.0;
          if (js >= R0 && js + 31 < i0 && i0 >= R0 && i0 + TD_RB - 1 < n) {
            const double uj = ldcg(ug + j);
            const double* p = A + (size_t)i0 * ldk + j;
            double av[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) av[r] = ldcg(p + (size_t)r * ldk);
            const double yold = ldcg(ycol + j);
            const double vj = (j == R0) ? 1.0 : uj * scale_;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              racc[r] = fma(av[r], vj, racc[r]);
              colacc = fma(av[r], vloc[ch * TD_RB + r], colacc);
            }
            ycol[j] = yold + colacc;
            vav = fma(vj, colacc, vav);
          } else {
            const bool jok = (j >= R0) && (j <= jmax);
            const double vj = !jok ? 0.0 : ((j == R0) ? 1.0 : ldcg(ug + j) * scale_);
            double av[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const int i = i0 + r;
              const bool ok = jok && (i >= R0) && (i < n) && (j <= i);
              av[r] = ok ? ldcg(A + (size_t)i * ldk + j) : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const int i = i0 + r;
              racc[r] = fma(av[r], vj, racc[r]);
              if (j < i) colacc = fma(av[r], vloc[ch * TD_RB + r], colacc);
            }
            if (jok) {
              ycol[j] = ldcg(ycol + j) + colacc;
              vav = fma(vj, colacc, vav);
            }
          }
        }
        // butterfly reduction of the 16 row accumulators over the warp (16 shuffles)
        double t8[8], t4[4], t2[2], t1;
        {
          const bool hi = lane & 16;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const double send = hi ? racc[r] : racc[r + 8];
            const double keep = hi ? racc[r + 8] : racc[r];
            t8[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool hi = lane & 8;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const double send = hi ? t8[r] : t8[r + 4];
            const double keep = hi ? t8[r + 4] : t8[r];
            t4[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        {
          const bool hi = lane & 4;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const double send = hi ? t4[r] : t4[r + 2];
            const double keep = hi ? t4[r + 2] : t4[r];
            t2[r] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
        {
          const bool hi = lane & 2;
          const double send = hi ? t2[0] : t2[1];
          const double keep = hi ? t2[1] : t2[0];
          t1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
        if ((lane & 1) == 0) {
          const int r = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
          rowpart[wid * LR + ch * TD_RB + r] = t1;
        }
      }
    }
    __syncthreads();
    TD_TICK(1)
    for (int lr = tid; lr < LR; lr += TD_THREADS) {
      const int i = global_row(lr, g, G);
      double s = 0.0;
      if (i >= R0 && i < n) {
#pragma unroll
        for (int w = 0; w < NW; ++w) s += rowpart[w * LR + lr];
        rsum[i] = s;
      }
      ys[lr] = s;
    }
    const double colv = block_sum(vav, red);
    // panel partial products pV_t = V_t . v, pW_t = W_t . v (t < k) and the local share of v^T A v
    for (int task = wid; task < 2 * k + 1; task += NW) {
      double s = 0.0;
      if (task < 2 * k) {
        const double* src = (task < k) ? (Vs + (size_t)task * LR) : (Ws + (size_t)(task - k) * LR);
        for (int lr = lane; lr < LR; lr += 32) s += src[lr] * vloc[lr];
      } else {
        for (int lr = lane; lr < LR; lr += 32) s += ys[lr] * vloc[lr];
      }
      s = warp_sum(s);
      if (lane == 0) mypart[(task < 2 * k) ? (2 + task) : 1] = (task < 2 * k) ? s : s + colv;
    }
    if (nbc > 0) {
      const int col = tid & (TD_BC - 1), grp = tid >> 7;
      double s = 0.0;
      if (col < nbc) {
        for (int lr0 = grp; lr0 < LR; lr0 += 8 * TD_NGRP) {
          double bv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int lr = lr0 + TD_NGRP * u;
            const bool ok = lr < LR && vloc[lr < LR ? lr : 0] != 0.0;
            bv[u] = ok ? ldcg(Bm + (size_t)global_row(lr, g, G) * ldb + col) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int lr = lr0 + TD_NGRP * u;
            if (lr < LR) s = fma(bv[u], vloc[lr], s);
          }
        }
      }
      Bred[grp * TD_BC + col] = s;
      __syncthreads();
      if (tid < nbc)
      {
        double bs = 0.0;
#pragma unroll
        for (int gq = 0; gq < TD_NGRP; ++gq) bs += Bred[gq * TD_BC + tid];
        mypart[2 + 2 * TD_MAXNB + tid] = bs;
      }
    }
    TD_TICK(2)
    bar.sync();
    TD_TICK(3)

    // ================= phase C: w = tau (A v - V (W^T v) - W (V^T v)) + alpha v, B update
    {
      const int nval = 2 * k + nbc + 1;   // <= 193
      {
        // every value is the sum of G per-CTA partials: 2 threads per value, independent loads
        const int q = tid & 255, h0 = tid >> 8;
        double s = 0.0;
        if (q < nval) {
          int slot;
          if (q < 2 * k) slot = 2 + q;
          else if (q < 2 * k + nbc) slot = 2 + 2 * TD_MAXNB + (q - 2 * k);
          else slot = 1;
          for (int hb = h0; hb < G; hb += 8 * TD_NH) {
            double pv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int h = hb + u * TD_NH;
              pv[u] = (h < G) ? ldcg(part + (size_t)h * TD_PART + slot) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s += pv[u];
          }
        }
        Bred[h0 * 256 + q] = s;
      }
      if (tid < k) {   // panel vectors at the pivot row of the next column (final since >= 1 barrier)
        Vc[tid] = ldcg(Vp + (size_t)tid * n + R0);
        Wc[tid] = ldcg(Wp + (size_t)tid * n + R0);
      }
      // column part of y for the owned rows: 4 threads per row
      for (int idx = tid; idx < 4 * LR; idx += TD_THREADS) {
        const int lr = idx % LR, p4 = idx / LR;
        const int i = global_row(lr, g, G);
        double y = 0.0;
        if (i >= R0 && i < n)
          for (int qb = p4; qb < G; qb += 32) {
            double pv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int q = qb + 4 * u;
              pv[u] = (q < G) ? ldcg(colpart + (size_t)q * n + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) y += pv[u];
          }
        rowpart[p4 * LR + lr] = y;
      }
      __syncthreads();
      if (tid < nval) {
        double s = Bred[tid];
        if (TD_NH > 1) s += Bred[256 + tid];
        if (tid < k) cV[tid] = s;
        else if (tid < 2 * k) cW[tid - k] = s;
        else if (tid < 2 * k + nbc) cB[tid - 2 * k] = s;
        else sc[4] = s;
      }
      __syncthreads();
      if (tid == 0) {
        double cc = 0.0;
        for (int t = 0; t < k; ++t) cc += cV[t] * cW[t];
        sc[0] = -0.5 * tau * tau * (sc[4] - 2.0 * cc);
      }
      if (g == 0 && tid < k) pgram[(size_t)(c + 1) * nb + tid] = cV[tid];
      __syncthreads();
      const double alpha = sc[0];
      for (int lr = tid; lr < LR; lr += TD_THREADS) {
        const int i = global_row(lr, g, G);
        double wv = 0.0;
        if (i >= R0 && i < n) {
          double y = ys[lr] + ((rowpart[lr] + rowpart[LR + lr]) + (rowpart[2 * LR + lr] + rowpart[3 * LR + lr]));
          for (int t = 0; t < k; ++t) y -= Vs[t * LR + lr] * cW[t] + Ws[t * LR + lr] * cV[t];
          wv = tau * y + alpha * vloc[lr];
        }
        Ws[k * LR + lr] = wv;
        if (i < n) Wp[(size_t)k * n + i] = wv;
      }
      if (nbc > 0 && tau != 0.0) {
        const int col = tid & (TD_BC - 1), grp = tid >> 7;
        if (col < nbc) {
          const double cb = tau * cB[col];
          // batches of 8 rows: all loads are issued before the first store (the loop would otherwise
          // serialise on load -> store round trips)
          for (int lr0 = grp; lr0 < LR; lr0 += 8 * TD_NGRP) {
            double bv[8];
            double* pp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int lr = lr0 + TD_NGRP * u;
              const bool ok = lr < LR && vloc[lr < LR ? lr : 0] != 0.0;
              pp[u] = ok ? Bm + (size_t)global_row(lr, g, G) * ldb + col : nullptr;
              bv[u] = ok ? ldcg(pp[u]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (pp[u]) *pp[u] = bv[u] - vloc[lr0 + TD_NGRP * u] * cb;
          }
        }
      }
      const bool panel_end = (k + 1 == nb) || (c == n - 2);
      if (!panel_end) {
        // W_k at the pivot row of the next column (row R0), needed by every CTA
        if (wid == 1) {
          double s = 0.0;
          for (int q = lane; q < G; q += 32) s += ldcg(colpart + (size_t)q * n + R0);
          s = warp_sum(s);
          if (lane == 0) {
            double y = s + ldcg(rsum + R0);
            for (int t = 0; t < k; ++t) y -= Vc[t] * cW[t] + Wc[t] * cV[t];
            Wc[k] = tau * y + alpha;   // v[R0] == 1
            Vc[k] = 1.0;
          }
        }
      }
      __syncthreads();
      TD_TICK(4)

      // ================= panel end: rank-2kp update of the lower triangle on the FP64 tensor cores
      if (panel_end) {
        const int kp = k + 1;
        bar.sync();
        const int Rn = c + 1;  // trailing block is rows/cols > c
        const int rb0 = Rn / TD_TM, nrb = (n + TD_TM - 1) / TD_TM - rb0;
        const int cb0 = Rn / TD_TN;
        const int kk_end = (2 * kp + 3) & ~3;
        const int wm = wid >> 2, wn = wid & 3;   // (THREADS/128) x 4 warps, warp tile 32 x 32
        const int fr = lane >> 2, fk = lane & 3;
        int tile_id = 0;
        for (int rb = 0; rb < nrb; ++rb) {
          const int i0 = (rb0 + rb) * TD_TM;
          const int ncb = (i0 + TD_TM - 1) / TD_TN - cb0 + 1;   // column blocks touching the lower triangle
          for (int cb = 0; cb < ncb; ++cb, ++tile_id) {
            if (tile_id % G != g) continue;
            const int j0 = (cb0 + cb) * TD_TN;
            double acc[4][4][2];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
            for (int kc = 0; kc < kk_end; kc += TD_KC) {
              const int kn_ = min(TD_KC, kk_end - kc);
              __syncthreads();
              for (int q = tid; q < kn_ * TD_TN; q += TD_THREADS) {
                const int kl = q / TD_TN, r = q % TD_TN, kk = kc + kl;
                double pa = 0.0, pb = 0.0;
                if (kk < 2 * kp) {
                  const double* srcA = (kk < kp) ? (Vp + (size_t)kk * n) : (Wp + (size_t)(kk - kp) * n);
                  const double* srcB = (kk < kp) ? (Wp + (size_t)kk * n) : (Vp + (size_t)(kk - kp) * n);
                  if (r < TD_TM && i0 + r < n) pa = ldcg(srcA + i0 + r);
                  if (j0 + r < n) pb = ldcg(srcB + j0 + r);
                }
                if (r < TD_TM) Pa[kl * TD_SA + r] = pa;
                Pb[kl * TD_SB + r] = pb;
              }
              __syncthreads();
              for (int kl = 0; kl < kn_; kl += 4) {
                double af[4], bf[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  af[x] = Pa[(kl + fk) * TD_SA + wm * 32 + x * 8 + fr];
                  bf[x] = Pb[(kl + fk) * TD_SB + wn * 32 + x * 8 + fr];
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                  for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
              }
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const int i = i0 + wm * 32 + x * 8 + fr;
              if (i >= Rn && i < n) {
                // all loads of the row first, then the stores (otherwise every element is a serial round trip)
                double2 o[4];
                bool full[4];
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                  const int j = j0 + wn * 32 + y * 8 + 2 * fk;
                  full[y] = (j >= Rn) && (j + 1 <= i);
                  const double* p = A + (size_t)i * ldk + j;
                  if (full[y]) {
                    o[y] = ldcg2(reinterpret_cast<const double2*>(p));
                  } else {
                    o[y].x = (j >= Rn && j <= i) ? ldcg(p) : 0.0;
                    o[y].y = (j + 1 >= Rn && j + 1 <= i) ? ldcg(p + 1) : 0.0;
                  }
                }
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                  const int j = j0 + wn * 32 + y * 8 + 2 * fk;
                  double* p = A + (size_t)i * ldk + j;
                  if (full[y]) {
                    *reinterpret_cast<double2*>(p) = make_double2(o[y].x - acc[x][y][0], o[y].y - acc[x][y][1]);
                  } else {
                    if (j >= Rn && j <= i) p[0] = o[y].x - acc[x][y][0];
                    if (j + 1 >= Rn && j + 1 <= i) p[1] = o[y].y - acc[x][y][1];
                  }
                }
              }
            }
          }
        }
        __syncthreads();
        // drop the finished panel
        for (int i = tid; i < 2 * nb * LR; i += TD_THREADS) Vs[i] = 0.0;
        bar.sync();
        jj0 = c + 1;
      }

      TD_TICK(5)
      // ================= phase A for the next column cn = c + 1
      const int cn = c + 1;
      if (cn <= n - 2) {
        const int kn = cn - jj0;
        const int R1 = cn + 1;
        double ss = 0.0;
        for (int lr = tid; lr < LR; lr += TD_THREADS) {
          const int i = global_row(lr, g, G);
          if (i >= R1 && i < n) {
            double x = ldcg(A + (size_t)i * ldk + cn);
            for (int t = 0; t < kn; ++t) x -= Vs[t * LR + lr] * Wc[t] + Ws[t * LR + lr] * Vc[t];
            ug[i] = x;
            if (i > R1) ss += x * x;
          }
        }
        ss = block_sum(ss, red);
        if (tid == 0) {
          mypart[0] = ss;
          if (g == 0) {
            double x = ldcg(A + (size_t)cn * ldk + cn);
            for (int t = 0; t < kn; ++t) x -= 2.0 * Vc[t] * Wc[t];
            dd[cn] = x;
          }
        }
        TD_TICK(6)
        bar.sync();
        TD_TICK(7)
      }
    }
  }
  if (timed)
    for (int q = 0; q < 8; ++q) a.timing[q] = tacc[q];
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
  double* gg = ss + TD_MAXNB;  // nb x nb panel Gram block
  for (int i = tid; i < n; i += blockDim.x) x[i] = a.x[(size_t)m * n + i];
  __syncthreads();
  // columns c = -1..n-2 live in panels [-1, nb-2], [nb-1, 2nb-2], ...
  const int ncol = n;  // number of reflectors
  const int npanel = (ncol + nb - 1) / nb;
  for (int pnl = npanel - 1; pnl >= 0; --pnl) {
    const int cfirst = -1 + pnl * nb;
    const int kp = min(nb, ncol - pnl * nb);
    // panel Gram block (v_t2 . v_t, t < t2) -> shared memory, so that the short recurrence below is not a chain
    // of dependent global loads
    for (int q = tid; q < kp * kp; q += blockDim.x) {
      const int t2 = q / kp, t = q % kp;
      gg[t2 * TD_MAXNB + t] = (t < t2) ? pg[(size_t)(cfirst + t2 + 1) * nb + t] : 0.0;
    }
    // c_t = v_t . x  (rows > c_t): one warp per reflector, 8 independent loads per lane in flight
    for (int t = wid; t < kp; t += nw) {
      const int c = cfirst + t;
      const double* vr = (c < 0) ? v0 : (A + (size_t)c * ldk);
      double s = 0.0;
      int i = c + 1 + lane;
      for (; i + 7 * 32 < n; i += 8 * 32) {
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = vr[i + 32 * u];
#pragma unroll
        for (int u = 0; u < 8; ++u) s = fma(r[u], x[i + 32 * u], s);
      }
      for (; i < n; i += 32) s = fma(vr[i], x[i], s);
      s = warp_sum(s);
      if (lane == 0) cc[t] = s;
    }
    __syncthreads();
    if (tid == 0) {
      for (int t = kp - 1; t >= 0; --t) {
        double s = cc[t];
        for (int t2 = t + 1; t2 < kp; ++t2) s -= gg[t2 * TD_MAXNB + t] * ss[t2];
        ss[t] = tau[cfirst + t + 1] * s;
      }
    }
    __syncthreads();
    for (int i = cfirst + 1 + tid; i < n; i += blockDim.x) {
      double s = x[i];
      for (int t0 = 0; t0 < kp; t0 += 8) {
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + u, c = cfirst + t;
          r[u] = (t < kp && i > c) ? ((c < 0) ? v0[i] : A[(size_t)c * ldk + i]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (t0 + u < kp) s = fma(-ss[t0 + u], r[u], s);
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
struct ProfSpan { cudaEvent_t e0, e1; int cat; };
static std::vector<ProfSpan> g_prof_events;
static size_t g_prof_used = 0;

int profile_enable(int on) {
  g_prof_on = on != 0;
  g_prof_used = 0;
  return CHD_OK;
}

bool profile_on() { return g_prof_on; }

// Brackets one launch (category `cat`) by CUDA events on its own stream.
int profile_mark_begin(int cat, cudaStream_t st) {
  if (!g_prof_on) return CHD_OK;
  if (g_prof_used == g_prof_events.size()) {
    ProfSpan sp;
    CHD_CUDA(cudaEventCreate(&sp.e0));
    CHD_CUDA(cudaEventCreate(&sp.e1));
    sp.cat = cat;
    g_prof_events.push_back(sp);
  }
  g_prof_events[g_prof_used].cat = cat;
  CHD_CUDA(cudaEventRecord(g_prof_events[g_prof_used].e0, st));
  return CHD_OK;
}

int profile_mark_end(int cat, cudaStream_t st) {
  (void)cat;
  if (!g_prof_on) return CHD_OK;
  CHD_CUDA(cudaEventRecord(g_prof_events[g_prof_used].e1, st));
  ++g_prof_used;
  return CHD_OK;
}

int profile_read_cat(int cat, double* total_ms, int* launches) {
  double tot = 0.0;
  int cnt = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    if (cat >= 0 && g_prof_events[i].cat != cat) continue;
    CHD_CUDA(cudaEventSynchronize(g_prof_events[i].e1));
    float ms = 0.f;
    CHD_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i].e0, g_prof_events[i].e1));
    tot += ms;
    ++cnt;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = cnt;
  return CHD_OK;
}

int profile_read(double* total_ms, int* launches) { return profile_read_cat(PROF_TRIDIAG, total_ms, launches); }

size_t tridiag_smem_bytes(const Plan* pl) {
  const int chunks = (pl->n + TD_RB - 1) / TD_RB;
  const int LR = (chunks + pl->G - 1) / pl->G * TD_RB;
  const int stage = TD_KC * (pl->threads / 4 + 8 + TD_SB);
  const int ring = pl->bulk ? TD_NS * TD_RB * pl->threads : 0;   // TD_NS x 16 rows x (32 * warps) columns
  size_t doubles = (size_t)(2 + pl->threads / 32 + 2 * pl->nb) * LR + 4 * TD_MAXNB + TD_BC + 512 + 40 + 8 + 8 +
                   (size_t)max(stage, ring);
  return doubles * sizeof(double);
}

size_t tridiag_workspace(const Plan* pl, int T) {
  const size_t n = pl->n, nb = pl->nb;
  size_t b = 0;
  b += 4 * align_up(T * n * sizeof(double), 256);             // tau, v0, u, rsum
  b += 2 * align_up(T * nb * n * sizeof(double), 256);        // Vp, Wp
  b += align_up((size_t)T * pl->G * n * sizeof(double), 256); // colpart
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
  tb.rsum = cv.take<double>(T * n);
  tb.Vp = cv.take<double>(T * nb * n);
  tb.Wp = cv.take<double>(T * nb * n);
  tb.colpart = cv.take<double>((size_t)T * pl->G * n);
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
  a.rsum = tb.rsum; a.colpart = tb.colpart; a.part = tb.part; a.pgram = tb.pgram; a.bar = tb.bar;
  a.n = n; a.G = pl->G; a.nb = pl->nb;
  a.LR = (chunks + pl->G - 1) / pl->G * TD_RB;
  a.nv = (int)align_up(n + 34, 2);
  const size_t smem = tridiag_smem_bytes(pl);
  static unsigned long long* dbg = nullptr;
  static const bool dbg_on = getenv("CHD_TD_TIMING") != nullptr;
  if (dbg_on && !dbg) CHD_CUDA(cudaMalloc(&dbg, 8 * sizeof(unsigned long long)));
  a.timing = dbg_on ? dbg : nullptr;
  void* kfn = pl->bulk ? (pl->threads == 256 ? (void*)tridiag_kernel<256, true> : (void*)tridiag_kernel<512, true>)
                       : (pl->threads == 256 ? (void*)tridiag_kernel<256, false> : (void*)tridiag_kernel<512, false>);
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
    b.rsum = tb.rsum + (size_t)t0 * n;
    b.colpart = tb.colpart + (size_t)t0 * pl->G * n;
    b.part = tb.part + (size_t)t0 * pl->G * TD_PART;
    b.pgram = tb.pgram + (size_t)t0 * n * pl->nb;
    b.bar = tb.bar + t0;
    void* params[] = {&b};
    int prc;
    if ((prc = profile_mark_begin(PROF_TRIDIAG, st))) return prc;
    CHD_CUDA(cudaLaunchCooperativeKernel(kfn, dim3(pl->G, cnt), dim3(pl->threads), params, smem, st));
    g_launches.fetch_add(1);
    if ((prc = profile_mark_end(PROF_TRIDIAG, st))) return prc;
  }
  if (dbg_on) {
    unsigned long long h[8];
    CHD_CUDA(cudaStreamSynchronize(st));
    CHD_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
    const char* nm[8] = {"setup", "symv", "partials", "sync2", "phaseC", "syr2k", "phaseA", "sync1"};
    fprintf(stderr, "[tridiag timing, last wave, ms]");
    for (int q = 0; q < 8; ++q) fprintf(stderr, " %s=%.2f", nm[q], h[q] * 1e-6);
    fprintf(stderr, "\n");
  }
  return CHD_OK;
}

int backtransform_launch(const Plan* pl, const double* K, int ldk, const TridiagBuffers& tb, const double* x,
                         const double* beta0, double* out, size_t out_stride, int T, cudaStream_t st) {
  BackArgs a;
  a.K = K; a.strideK = (size_t)pl->n * ldk; a.ldk = ldk; a.v0 = tb.v0; a.tau = tb.tau; a.pgram = tb.pgram;
  a.x = x; a.beta0 = beta0; a.out = out; a.out_stride = out_stride; a.n = pl->n; a.nb = pl->nb;
  const size_t smem = ((size_t)pl->n + 2 * TD_MAXNB + TD_MAXNB * TD_MAXNB + 8) * sizeof(double);
  backtransform_kernel<<<T, 512, smem, st>>>(a);
  CHD_LAUNCHED();
  return CHD_OK;
}

int tridiag_device_init(Plan* pl) {
  void* fns[5] = {(void*)tridiag_kernel<256, false>, (void*)tridiag_kernel<512, false>,
                  (void*)tridiag_kernel<256, true>, (void*)tridiag_kernel<512, true>, (void*)backtransform_kernel};
  for (void* f : fns)
    CHD_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_optin));
  return CHD_OK;
}

}  // namespace chd
