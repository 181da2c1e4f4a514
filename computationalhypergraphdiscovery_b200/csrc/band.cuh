// Shared declarations of the dense -> band stage (band.cu: cp.async kernels + host orchestration; band_tma.cu: the
// TMA + mbarrier warp-specialised GEMM kernels).
#pragma once
#include "internal.cuh"

namespace chd {

constexpr int NB = CHD_BAND;
constexpr int KD = 2;               // panels per deferred-update group
constexpr int LDV = NB * KD;        // row stride of the V / W operand arrays: [panel q = 0 | panel q = 1]
constexpr int SY_BM = 128;          // rows of a symm row block
constexpr int SY_ZW = 128;          // padded width of the Z-block partials
constexpr int SY_MAXSPLIT = 4;      // split-K factor of the symm (partial Y buffers)

struct SymmArgs {
  const double* A; size_t strideA; int ldk;
  const double* V;   // T x n x LDV, offset to this panel's column block
  double* Y;         // T x n x NB
  const double* Bq;  // T x n x ldb or null
  int ldb, nz;
  double* Mpart;     // T x nblk x NB x NB    (nblk = row blocks x splits)
  double* Zpart;     // T x nzblk x NB x SY_ZW (nzblk = row blocks)
  // second panel of a deferred-update group: the first panel's V_a, W_a (T x n x LDV, column block 0) and the
  // partials of G1 = W_a^T V, G2 = V_a^T V (T x nzblk x 2 x NB x NB); null otherwise
  const double* Va; const double* Wa; double* Gpart;
  int n, r0, nblk, nzblk;
};

// Epilogue shared by both symm kernels, run by 256 threads (ctid = 0..255, 8 warps) that hold the 128 x 32 block of Y
// in DMMA accumulator layout (warp w: rows 16 w + 8 x + fr, columns 8 y + 2 fk + {0,1}): stores the Y rows (split-K
// partial `sp` of `S`), then the partial products of this row block: V_I^T Y_I (Mpart), V_I^T B_I (Zpart, sp == 0)
// and, for the second panel of a group, W_a^T V and V_a^T V (Gpart, sp == 0).  Ys, Vs: SY_BM x (NB + 1) doubles of
// shared memory each; `sync` synchronises the 256 threads.
template <class Sync>
__device__ __forceinline__ void symm_epilogue(const SymmArgs& a, const double (&acc)[2][4][2], double* Ys, double* Vs,
                                              int ctid, int t, int bx, int sp, int S, int T, Sync sync) {
  const int lane = ctid & 31, wid = ctid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = a.n, i0 = a.r0 + SY_BM * bx;
  const double* V = a.V + (size_t)t * n * LDV;
  double* Y = a.Y + ((size_t)sp * T + t) * n * NB;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int rl = 16 * wid + 8 * x + fr, i = i0 + rl;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int c = y * 8 + 2 * fk;
      Ys[rl * (NB + 1) + c] = acc[x][y][0];
      Ys[rl * (NB + 1) + c + 1] = acc[x][y][1];
      if (i < n) *reinterpret_cast<double2*>(Y + (size_t)i * NB + c) = make_double2(acc[x][y][0], acc[x][y][1]);
    }
  }
  for (int e = ctid; e < SY_BM * NB; e += 256) {
    const int rl = e / NB, c = e % NB;
    Vs[rl * (NB + 1) + c] = (i0 + rl < n) ? V[(size_t)(i0 + rl) * LDV + c] : 0.0;
  }
  sync();
  {
    double mp[4] = {0.0, 0.0, 0.0, 0.0};
    for (int rl = 0; rl < SY_BM; ++rl) {
      const double yv = Ys[rl * (NB + 1) + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) mp[u] = fma(Vs[rl * (NB + 1) + wid + 8 * u], yv, mp[u]);
    }
    double* Mp = a.Mpart + ((size_t)t * a.nblk + bx * S + sp) * NB * NB;
#pragma unroll
    for (int u = 0; u < 4; ++u) Mp[(wid + 8 * u) * NB + lane] = mp[u];
  }
  // deferred-update group, second panel: partials of G1 = W_a^T V and G2 = V_a^T V over this row block
  if (a.Gpart && sp == 0) {
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const double* src = (which == 0 ? a.Wa : a.Va) + (size_t)t * n * LDV;
      sync();                                // Ys (Y rows, then W_a rows) has been consumed
      for (int e = ctid; e < SY_BM * NB; e += 256) {
        const int rl = e / NB, c = e % NB;
        Ys[rl * (NB + 1) + c] = (i0 + rl < n) ? src[(size_t)(i0 + rl) * LDV + c] : 0.0;
      }
      sync();
      double gp[4] = {0.0, 0.0, 0.0, 0.0};
      for (int rl = 0; rl < SY_BM; ++rl) {
        const double vv = Vs[rl * (NB + 1) + lane];
#pragma unroll
        for (int u = 0; u < 4; ++u) gp[u] = fma(Ys[rl * (NB + 1) + wid + 8 * u], vv, gp[u]);
      }
      double* Gp = a.Gpart + (((size_t)t * a.nzblk + bx) * 2 + which) * NB * NB;
#pragma unroll
      for (int u = 0; u < 4; ++u) Gp[(wid + 8 * u) * NB + lane] = gp[u];
    }
  }
  if (a.Bq && a.nz > 0 && sp == 0) {
    const double* Bq = a.Bq + (size_t)t * n * a.ldb;
    const int z = ctid & 127, c1g = ctid >> 7;   // 2 groups of 16 reflector columns
    double zp[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) zp[u] = 0.0;
    if (z < a.nz) {
      for (int rl = 0; rl < SY_BM && i0 + rl < n; ++rl) {
        const double bv = Bq[(size_t)(i0 + rl) * a.ldb + z];
#pragma unroll
        for (int u = 0; u < 16; ++u) zp[u] = fma(Vs[rl * (NB + 1) + 16 * c1g + u], bv, zp[u]);
      }
    }
    double* Zp = a.Zpart + ((size_t)t * a.nzblk + bx) * NB * SY_ZW;
#pragma unroll
    for (int u = 0; u < 16; ++u) Zp[(16 * c1g + u) * SY_ZW + z] = zp[u];
  }
}

// TMA kernels (band_tma.cu)
int symm_tma_launch(const Plan* pl, const SymmArgs& a, int T, int S, cudaStream_t st);

}  // namespace chd
