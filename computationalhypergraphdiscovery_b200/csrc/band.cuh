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
  double* Mpart;     // T x nblk x NB x NB     (nblk = row blocks)
  double* Zpart;     // T x nblk x NB x SY_ZW
  // second panel of a deferred-update group: the first panel's V_a, W_a (T x n x LDV, column block 0) and the
  // partials of G1 = W_a^T V, G2 = V_a^T V (T x nblk x 2 x NB x NB); null otherwise
  const double* Va; const double* Wa; double* Gpart;
  int n, r0, nblk;
  int nsplit, T;     // split-K partial buffers of Y (summed into buffer 0 by symm_partials_kernel), targets
};

// Epilogue shared by both symm kernels, run by 256 threads (ctid = 0..255, 8 warps) that hold the 128 x 32 block of Y
// in DMMA accumulator layout (warp w: rows 16 w + 8 x + fr, columns 8 y + 2 fk + {0,1}): stores the Y rows of split-K
// partial `sp`.  The products of the row block with V, B and the group's first panel (V^T Y, V^T B, Gram blocks) are
// computed by symm_partials_kernel, at full occupancy, instead of at the tail of every GEMM CTA.
__device__ __forceinline__ void symm_store_y(const SymmArgs& a, const double (&acc)[2][4][2], int ctid, int t, int bx,
                                             int sp, int T) {
  const int lane = ctid & 31, wid = ctid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = a.n, i0 = a.r0 + SY_BM * bx;
  double* Y = a.Y + ((size_t)sp * T + t) * n * NB;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int i = i0 + 16 * wid + 8 * x + fr;
    if (i < n) {
#pragma unroll
      for (int y = 0; y < 4; ++y)
        *reinterpret_cast<double2*>(Y + (size_t)i * NB + y * 8 + 2 * fk) = make_double2(acc[x][y][0], acc[x][y][1]);
    }
  }
}

// TMA kernels (band_tma.cu)
int symm_tma_launch(const Plan* pl, const SymmArgs& a, int T, int S, cudaStream_t st);

}  // namespace chd
