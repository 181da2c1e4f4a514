// Trailing update of the dense -> band stage as a TMA-fed, warp-specialised DMMA kernel (sm_100a):
//     A22 -= Z Z'^T,   Z = [V_a V_b | W_a W_b],  Z' = [W_a W_b | V_a V_b]     (K = 128 for a group of two panels)
// replaces, together with the rest of band.cu, the O(n^3) part of jnp.linalg.eigh (interpolatory.py:48).
//
//   * one CTA = 1 producer warp + 8 DMMA consumer warps; it owns a 128-row block I and walks a chunk of 16-column
//     tiles J of that block row (lower triangle only);
//   * every tile of A, of Z_I (once per CTA) and of Z'_J (once per tile) travels global -> shared memory as
//     16-column boxes through tensor maps (cp.async.bulk.tensor, SASS UTMALDG) with the 128-byte swizzle, which makes
//     every DMMA fragment load bank-conflict free without padding;
//   * a 3-stage ring of {A tile, Z'_J tile} is handed from the producer to the consumers through mbarrier full / empty
//     pairs (SASS SYNCS): the consumers never meet at a CTA-wide barrier, the producer runs up to three tiles ahead,
//     so the HBM stream (A read + write) and the DMMA pipe overlap instead of alternating;
//   * accumulators start at -A (read from the staged tile), results go from registers to global memory.
// FP64 has no tcgen05 kind: DMMA m8n8k4 (mma.sync) is the FP64 tensor instruction of sm_100a.
#include "band.cuh"
#include "tma.cuh"

namespace chd {

namespace {

constexpr int TM_BN = 16;            // tile columns (one 16-column box)
constexpr int TM_CH = 32;            // tiles per CTA (512 columns)
constexpr int BOX_ROWS = 16;         // every box is 16 columns x 16 rows (2 KB)
constexpr int BOX_BYTES = 16 * BOX_ROWS * 8;

// BM = rows of the CTA's block (128: one CTA per SM with a 3-stage ring; 64: two CTAs per SM with 2-stage rings, so that
// the Z_I load and the pipeline fill of one CTA overlap with the other's main loop), G = panels of the group
template <int G, int BM> struct TmCfg {
  static constexpr int NST = (BM == 128) ? 3 : 2;       // ring depth
  static constexpr int CONS_WARPS = BM / 16;
  static constexpr int THREADS = 32 * (1 + CONS_WARPS);
  static constexpr int KT = 2 * G * NB;                 // K of the update (64 or 128)
  static constexpr int KB = KT / 16;                    // 16-column boxes along K
  static constexpr int ZI_BYTES = KB * BM * 128;        // Z_I: KB box-columns of BM rows
  static constexpr int ZJ_BYTES = KB * TM_BN * 128;     // Z'_J: KB box-columns of 16 rows
  static constexpr int A_BYTES = BM * 128;              // A tile
  static constexpr int STAGE_BYTES = ZJ_BYTES + A_BYTES;
  static constexpr int BAR_OFF = ZI_BYTES + NST * STAGE_BYTES;
  static constexpr int SMEM = BAR_OFF + 64;             // the dynamic shared array is declared 1024-byte aligned
};

struct TmArgs {
  double* A; size_t strideA; int ldk;
  int n, r0;
};

template <int G, int BM>
__global__ void __launch_bounds__(TmCfg<G, BM>::THREADS, BM == 128 ? 1 : 2)
syr2k_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapV,
                 const __grid_constant__ CUtensorMap mapW, TmArgs a) {
  using C = TmCfg<G, BM>;
  constexpr int TM_NST = C::NST;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // swizzle period: box stacks start on 1024-byte boundaries
  uint8_t* sm = smem_raw;
  uint8_t* zi = sm;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::BAR_OFF);
  uint64_t* full = bars;                 // TM_NST
  uint64_t* empty = bars + TM_NST;       // TM_NST
  uint64_t* zi_full = bars + 2 * TM_NST;

  const int t = blockIdx.z;
  const int n = a.n, r0 = a.r0;
  const int i0 = r0 + BM * blockIdx.y;
  // 16-column tiles of this block row that touch the lower triangle: j0 = r0 + 16 jt <= i0 + BM - 1, j0 < n
  const int njt = min((i0 + BM - r0) / TM_BN, (n - r0 + TM_BN - 1) / TM_BN);
  const int jt_lo = TM_CH * blockIdx.x, jt_hi = min(njt, jt_lo + TM_CH);
  if (jt_lo >= jt_hi) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < TM_NST; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, C::CONS_WARPS);
    }
    tma::mbar_init(zi_full, 1);
    tma::fence_mbar_init();
  }
  __syncthreads();

  if (wid == 0) {
    // ===================== producer warp: all global -> shared traffic =====================
    if (lane == 0) {
      tma::prefetch_map(&mapA);
      tma::prefetch_map(&mapV);
      tma::prefetch_map(&mapW);
      tma::mbar_expect_tx(zi_full, C::ZI_BYTES);
    }
    __syncwarp();
    // Z_I = [V_I | W_I]: box-column kb (16 K-columns) x BM / 16 row boxes; lanes share the copies
    for (int b = lane; b < C::KB * (BM / BOX_ROWS); b += 32) {
      const int kb = b / (BM / BOX_ROWS), rb = b % (BM / BOX_ROWS);
      const bool fromV = kb < C::KB / 2;
      const int col = 16 * (fromV ? kb : kb - C::KB / 2);
      tma::load_box(zi + kb * (BM * 128) + rb * BOX_BYTES, fromV ? &mapV : &mapW, col, i0 + BOX_ROWS * rb, t,
                    zi_full);
    }
    for (int jt = jt_lo; jt < jt_hi; ++jt) {
      const int it = jt - jt_lo, s = it % TM_NST, use = it / TM_NST;
      if (use > 0) tma::mbar_wait(empty + s, (use - 1) & 1);      // consumers have drained the previous use
      uint8_t* st = sm + C::ZI_BYTES + s * C::STAGE_BYTES;
      const int j0 = r0 + TM_BN * jt;
      if (lane == 0) tma::mbar_expect_tx(full + s, C::STAGE_BYTES);
      __syncwarp();
      // Z'_J = [W_J | V_J]: KB boxes of 16 rows; the A tile: BM / 16 boxes of 16 rows
      for (int b = lane; b < C::KB + BM / BOX_ROWS; b += 32) {
        if (b < C::KB) {
          const bool fromW = b < C::KB / 2;
          const int col = 16 * (fromW ? b : b - C::KB / 2);
          tma::load_box(st + b * BOX_BYTES, fromW ? &mapW : &mapV, col, j0, t, full + s);
        } else {
          const int rb = b - C::KB;
          tma::load_box(st + C::ZJ_BYTES + rb * BOX_BYTES, &mapA, j0, i0 + BOX_ROWS * rb, t, full + s);
        }
      }
    }
    return;
  }

  // ===================== consumer warps: DMMA =====================
  const int cw = wid - 1;                    // rows 16 cw .. 16 cw + 15 of the block, all 16 tile columns
  const int fr = lane >> 2, fk = lane & 3;
  // swizzled byte offsets of this lane's fragment element inside a 128-byte row: column kk + fk of a 16-column box,
  // kk = 0, 4, 8, 12  (chunk = ((kk + fk) / 2) ^ fr, see tma::swz128; rows used here are = fr mod 8)
  uint32_t foff[4];
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) foff[c4] = tma::swz128(fr, 4 * c4 + fk);
  const uint8_t* zi_row = zi + (16 * cw) * 128;               // foff carries the lane's row fr; + 8 x rows: + 1024
  double* Ag = a.A + (size_t)t * a.strideA;
  tma::mbar_wait(zi_full, 0);
  for (int jt = jt_lo; jt < jt_hi; ++jt) {
    const int it = jt - jt_lo, s = it % TM_NST, use = it / TM_NST;
    const uint8_t* st = sm + C::ZI_BYTES + s * C::STAGE_BYTES;
    const uint8_t* zj_row = st;                               // foff carries the lane's row fr; + 8 y rows: + 1024
    const uint8_t* at = st + C::ZJ_BYTES;
    tma::mbar_wait(full + s, use & 1);
    // accumulators start at -A: c0, c1 = columns 8 y + 2 fk, + 1 of row 16 cw + 8 x + fr
    double acc[2][2][2];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const double2 o = *reinterpret_cast<const double2*>(at + tma::swz128(16 * cw + 8 * x + fr, 8 * y + 2 * fk));
        acc[x][y][0] = -o.x;
        acc[x][y][1] = -o.y;
      }
#pragma unroll
    for (int kb = 0; kb < C::KB; ++kb) {
      const uint8_t* zik = zi_row + kb * (BM * 128);
      const uint8_t* zjk = zj_row + kb * BOX_BYTES;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        double af[2], bf[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) af[x] = *reinterpret_cast<const double*>(zik + x * 1024 + foff[c4]);
#pragma unroll
        for (int y = 0; y < 2; ++y) bf[y] = *reinterpret_cast<const double*>(zjk + y * 1024 + foff[c4]);
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
      }
    }
    // this warp is done with the stage: hand it back to the producer before the (slow) global stores
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(empty + s);
    const int j0 = r0 + TM_BN * jt;
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int i = i0 + 16 * cw + 8 * x + fr;
      if (i < n) {
#pragma unroll
        for (int y = 0; y < 2; ++y) {
          const int j = j0 + 8 * y + 2 * fk;
          double* p = Ag + (size_t)i * a.ldk + j;
          if (j + 1 <= i) *reinterpret_cast<double2*>(p) = make_double2(-acc[x][y][0], -acc[x][y][1]);
          else if (j <= i) p[0] = -acc[x][y][0];
        }
      }
    }
  }
}

// ------------------------------------------------------------------ Y = A22 V  (TMA + mbarrier)
// The symmetric product from the lower triangle only, as a warp-specialised kernel: a CTA owns the 128 rows of row
// block I and accumulates Y_I over the other row blocks P of the trailing matrix -- P < I contributes the direct block
// A[I, P] V_P, P > I the transposed block A[P, I]^T V_P, P = I both halves of the diagonal block (masked).
//   * PAIRED SCHEDULE: at step s the CTA of block I works on P = (s - I) mod nb; the CTA of block P works on I at the
//     same step, i.e. both read the SAME 128 x 128 block of A (one as it is, one transposed) at the same time, so the
//     second read hits L2: the lower triangle streams from HBM once per panel instead of twice;
//   * tiles: direct chunks (128 rows x 32 k) as 16-column boxes with the 128-byte swizzle (conflict-free row-major
//     fragment loads), transposed chunks (32 k x 128 rows) and the V chunks (32 k x 32) as 8-column unswizzled boxes
//     (a DMMA fragment then reads 4 k-rows x 64 B = 256 contiguous bytes) -- all through tensor maps (UTMALDG);
//   * 4-stage ring with mbarrier full / empty pairs, 1 producer warp + 8 DMMA consumer warps, no CTA-wide barrier in the
//     main loop; the epilogue only stores the Y rows (the V^T Y / V^T B / Gram partials: symm_partials_kernel).
constexpr int SM_A_BYTES = SY_BM * 32 * 8;        // 32 KB: one chunk of A (either orientation)
constexpr int SM_V_BYTES = 32 * NB * 8;           // 8 KB: V chunk
constexpr int SM_STAGE = SM_A_BYTES + SM_V_BYTES;
constexpr int SM_THREADS = 32 * (1 + 8);
// SM_NST stages per CTA: 4 with one CTA per SM, 2 with two CTAs per SM (16 consumer warps per SM, the epilogue and the
// pipeline fill of one CTA overlap with the other's main loop)
template <int SM_NST> constexpr int symm_tma_smem() { return SM_NST * SM_STAGE + 64; }

template <int SM_NST>
__global__ void __launch_bounds__(SM_THREADS, SM_NST == 2 ? 2 : 1)
symm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAt,
                const __grid_constant__ CUtensorMap mapV, SymmArgs a) {
  constexpr int SM_BAR_OFF = SM_NST * SM_STAGE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + SM_NST;
  const int t = blockIdx.y, I = blockIdx.x, sp = blockIdx.z, S = gridDim.z, nb = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = a.n, r0 = a.r0;
  const int i0 = r0 + SY_BM * I;
  // split-K over the steps: this CTA handles steps [s_lo, s_hi) of its row block
  const int s_lo = (int)((long long)nb * sp / S), s_hi = (int)((long long)nb * (sp + 1) / S);
  // number of chunks: 4 per step, 8 for the step that meets the diagonal block (s = 2 I mod nb)
  int nit = 4 * (s_hi - s_lo);
  {
    const int sd = (2 * I) % nb;
    if (sd >= s_lo && sd < s_hi) nit += 4;
  }
  if (tid == 0) {
    for (int s = 0; s < SM_NST; ++s) {
      tma::mbar_init(full + s, 1);
      tma::mbar_init(empty + s, 8);
    }
    tma::fence_mbar_init();
  }
  __syncthreads();

  if (wid == 0) {
    // ===================== producer warp =====================
    if (lane == 0) {
      tma::prefetch_map(&mapA);
      tma::prefetch_map(&mapAt);
      tma::prefetch_map(&mapV);
    }
    int s = s_lo, rem_in_step = 0;
    for (int it = 0; it < nit; ++it) {
      int P = (s - I) % nb;
      if (P < 0) P += nb;
      const int cnt = (P == I) ? 8 : 4;
      const bool tr = (P > I) || (P == I && rem_in_step >= 4);
      const int c = rem_in_step & 3;
      const int st = it % SM_NST, use = it / SM_NST;
      if (use > 0) tma::mbar_wait(empty + st, (use - 1) & 1);
      uint8_t* sb = sm + st * SM_STAGE;
      const int j0 = r0 + SY_BM * P + 32 * c;          // first k index (row of V, column / row of A) of the chunk
      if (lane == 0) tma::mbar_expect_tx(full + st, SM_STAGE);
      __syncwarp();
      if (lane < 16) {
        if (!tr) {        // direct: box (bc, rb): columns j0 + 16 bc, rows i0 + 16 rb
          const int bc = lane >> 3, rb = lane & 7;
          tma::load_box(sb + bc * (SY_BM * 128) + rb * 2048, &mapA, j0 + 16 * bc, i0 + 16 * rb, t, full + st);
        } else {          // transposed: box b: columns i0 + 8 b, rows j0 .. j0 + 31
          tma::load_box(sb + lane * 2048, &mapAt, i0 + 8 * lane, j0, t, full + st);
        }
      } else if (lane < 20) {
        const int y = lane - 16;                        // V chunk: columns 8 y, rows j0 .. j0 + 31
        tma::load_box(sb + SM_A_BYTES + y * 2048, &mapV, 8 * y, j0, t, full + st);
      }
      if (++rem_in_step == cnt) { rem_in_step = 0; ++s; }
    }
    return;
  }

  // ===================== consumer warps =====================
  const int cw = wid - 1, ctid = tid - 32;
  const int fr = lane >> 2, fk = lane & 3;
  uint32_t foff[4];                                     // direct fragment: column 4 c4 + fk of a swizzled 16-column box
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) foff[c4] = tma::swz128(fr, 4 * c4 + fk);
  const uint32_t toff = fk * 64 + fr * 8;               // transposed / V fragment inside an 8-column box: row fk, col fr
  double acc[2][4][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
  {
    int s = s_lo, rem_in_step = 0;
    for (int it = 0; it < nit; ++it) {
      int P = (s - I) % nb;
      if (P < 0) P += nb;
      const int cnt = (P == I) ? 8 : 4;
      const bool mk = (P == I);
      const bool tr = (P > I) || (mk && rem_in_step >= 4);
      const int c = rem_in_step & 3;
      if (++rem_in_step == cnt) { rem_in_step = 0; ++s; }
      const int st = it % SM_NST, use = it / SM_NST;
      const uint8_t* sb = sm + st * SM_STAGE;
      const uint8_t* vb = sb + SM_A_BYTES + toff;
      tma::mbar_wait(full + st, use & 1);
      const int jl0 = 32 * c;                           // k offset of the chunk inside the (diagonal) block
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {                  // 8 k-steps of 4
        double af[2], bf[4];
#pragma unroll
        for (int y = 0; y < 4; ++y) bf[y] = *reinterpret_cast<const double*>(vb + y * 2048 + ks * 256);
        if (!tr) {
          const uint8_t* ab = sb + (ks >> 2) * (SY_BM * 128) + (16 * cw) * 128 + foff[ks & 3];
#pragma unroll
          for (int x = 0; x < 2; ++x) af[x] = *reinterpret_cast<const double*>(ab + x * 1024);
        } else {
          const uint8_t* ab = sb + (2 * cw) * 2048 + ks * 256 + toff;
#pragma unroll
          for (int x = 0; x < 2; ++x) af[x] = *reinterpret_cast<const double*>(ab + x * 2048);
        }
        if (mk) {
          // diagonal block: the direct half keeps j <= i, the transposed half j > i (local indices inside the block)
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const int il = 16 * cw + 8 * x + fr, jl = jl0 + 4 * ks + fk;
            if (tr ? (jl <= il) : (jl > il)) af[x] = 0.0;
          }
        }
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
      }
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(empty + st);
    }
  }
  // ---- epilogue: the Y rows of this split-K partial, straight from the accumulators
  symm_store_y(a, acc, ctid, t, I, sp, (int)gridDim.y);
}

}  // namespace

int band_tma_device_init(Plan* pl) {
  pl->syr2k_tma = 1;
  if (const char* e = getenv("CHD_SYR2K_TMA")) pl->syr2k_tma = atoi(e) ? 1 : 0;
  if (!tma::encode_fn()) pl->syr2k_tma = 0;
  pl->syr2k_tma_ctas = 2;
  if (const char* e = getenv("CHD_SYR2K_TMA_CTAS")) pl->syr2k_tma_ctas = atoi(e) == 1 ? 1 : 2;
  CHD_CUDA(cudaFuncSetAttribute(syr2k_tma_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                TmCfg<1, 128>::SMEM));
  CHD_CUDA(cudaFuncSetAttribute(syr2k_tma_kernel<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                TmCfg<2, 128>::SMEM));
  CHD_CUDA(cudaFuncSetAttribute(syr2k_tma_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                TmCfg<1, 64>::SMEM));
  CHD_CUDA(cudaFuncSetAttribute(syr2k_tma_kernel<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                TmCfg<2, 64>::SMEM));
  pl->symm_tma = 1;
  if (const char* e = getenv("CHD_SYMM_TMA")) pl->symm_tma = atoi(e) ? 1 : 0;
  if (!tma::encode_fn()) pl->symm_tma = 0;
  pl->symm_tma_ctas = 2;
  if (const char* e = getenv("CHD_SYMM_TMA_CTAS")) pl->symm_tma_ctas = atoi(e) == 1 ? 1 : 2;
  CHD_CUDA(cudaFuncSetAttribute(symm_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, symm_tma_smem<4>()));
  CHD_CUDA(cudaFuncSetAttribute(symm_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, symm_tma_smem<2>()));
  return CHD_OK;
}

// Y = A[r0:, r0:] V with the partials of the epilogue (see SymmArgs); S = split-K factor (grid z)
int symm_tma_launch(const Plan* pl, const SymmArgs& a, int T, int S, cudaStream_t st) {
  const int n = pl->n, m = n - a.r0;
  CUtensorMap mapA, mapAt, mapV;
  int rc;
  if ((rc = tma::make_map_f64(&mapA, a.A, (uint64_t)n, (uint64_t)n, (uint64_t)T, (uint64_t)a.ldk, 16, 16))) return rc;
  if ((rc = tma::make_map_f64(&mapAt, a.A, (uint64_t)n, (uint64_t)n, (uint64_t)T, (uint64_t)a.ldk, 32, 8))) return rc;
  if ((rc = tma::make_map_f64(&mapV, a.V, (uint64_t)NB, (uint64_t)n, (uint64_t)T, (uint64_t)LDV, 32, 8))) return rc;
  const int nblk = (m + SY_BM - 1) / SY_BM;
  if (pl->symm_tma_ctas == 2)
    symm_tma_kernel<2><<<dim3(nblk, T, S), SM_THREADS, symm_tma_smem<2>(), st>>>(mapA, mapAt, mapV, a);
  else
    symm_tma_kernel<4><<<dim3(nblk, T, S), SM_THREADS, symm_tma_smem<4>(), st>>>(mapA, mapAt, mapV, a);
  CHD_LAUNCHED();
  return CHD_OK;
}

// A[r0:, r0:] -= sum over the G panels of the group (V, W: T x n x LDV, column blocks 0 .. G-1), lower tiles only
int syr2k_tma_launch(const Plan* pl, double* K, int ldk, int T, int r0, int G, const double* V, const double* W,
                     cudaStream_t st) {
  const int n = pl->n, m = n - r0;
  CUtensorMap mapA, mapV, mapW;
  int rc;
  if ((rc = tma::make_map_f64(&mapA, K, (uint64_t)n, (uint64_t)n, (uint64_t)T, (uint64_t)ldk, BOX_ROWS))) return rc;
  if ((rc = tma::make_map_f64(&mapV, V, (uint64_t)LDV, (uint64_t)n, (uint64_t)T, (uint64_t)LDV, BOX_ROWS))) return rc;
  if ((rc = tma::make_map_f64(&mapW, W, (uint64_t)LDV, (uint64_t)n, (uint64_t)T, (uint64_t)LDV, BOX_ROWS))) return rc;
  TmArgs a;
  a.A = K; a.strideA = (size_t)n * ldk; a.ldk = ldk; a.n = n; a.r0 = r0;
  const int ntile = (m + TM_BN - 1) / TM_BN;
  if (pl->syr2k_tma_ctas == 2) {
    const dim3 grid((ntile + TM_CH - 1) / TM_CH, (m + 63) / 64, T);
    if (G == 1)
      syr2k_tma_kernel<1, 64><<<grid, TmCfg<1, 64>::THREADS, TmCfg<1, 64>::SMEM, st>>>(mapA, mapV, mapW, a);
    else
      syr2k_tma_kernel<2, 64><<<grid, TmCfg<2, 64>::THREADS, TmCfg<2, 64>::SMEM, st>>>(mapA, mapV, mapW, a);
  } else {
    const dim3 grid((ntile + TM_CH - 1) / TM_CH, (m + 127) / 128, T);
    if (G == 1)
      syr2k_tma_kernel<1, 128><<<grid, TmCfg<1, 128>::THREADS, TmCfg<1, 128>::SMEM, st>>>(mapA, mapV, mapW, a);
    else
      syr2k_tma_kernel<2, 128><<<grid, TmCfg<2, 128>::THREADS, TmCfg<2, 128>::SMEM, st>>>(mapA, mapV, mapW, a);
  }
  CHD_LAUNCHED();
  return CHD_OK;
}

}  // namespace chd
