// Stage 1 of the two-stage reduction: batched dense -> band (half-width NB = 32) reduction of
//   K1 = H_{-1} K H_{-1}   (H_{-1} ga = beta0 e_0)
// by block Householder reflectors, with the Z-test block carried along (B <- Q^T S), and the
// matching back-transformation  yb = -beta0 Q x.  Together with chase.cu / bandsolve.cu / gamma.cu
// this replaces the reference's jnp.linalg.eigh(K) + V^T ga + V^T S + V c
// (interpolatory.py:48-51, 76-80, 100); DESIGN.md section 3b, NumPy model: tests/twostage_model.py.
//
// Per panel p (columns c0 .. c0+NB-1, rows r0 = c0+NB .. n-1 below the band), all T matrices of the
// batch in one launch each (blockIdx.y / .z = matrix):
//   panel_qr_kernel  Householder QR of the (n-r0) x NB block, one thread-block CLUSTER per matrix:
//                    the block lives in the distributed shared memory of the cluster, column norms
//                    and v^T P are exchanged through DSMEM + cluster barriers.  Emits R (in place),
//                    V (in place below R, and as a dense n x NB operand) and the compact-WY factor.
//   symm_kernel      Y = A22 V on the FP64 tensor cores (DMMA m8n8k4), lower triangle only: each
//                    CTA owns 128 rows and walks the row of tiles (direct left of the diagonal,
//                    transposed right of it), cp.async double-buffered; epilogue V^T Y and V^T B partials
//   reduce_kernel    fixed-order sum of the per-CTA partials (deterministic, no atomics)
//   wform_kernel     W = (Y - 1/2 V T^T (V^T Y)) T,  B -= V T^T (V^T B)
//   syr2k_kernel     A22 -= V W^T + W V^T on the lower tiles (DMMA, K = 2 NB)
// Everything is GEMM-shaped: 4/3 n^3 flop on the tensor cores.
//
// DEFERRED trailing update (default, CHD_KD=2): panels are processed in groups of KD = 2 (slot 0 = H_{-1}, slots
// 1.. = QR panels; groups (0,1), (2,3), ...).  Inside a group the trailing matrix stays stale: the second panel's
// block column is brought up to date by panel_update_kernel, its Y = A22 V comes from the stale matrix plus the
// corrections  Y -= V_a (W_a^T V) + W_a (V_a^T V)  (the two 32 x 32 Gram blocks are partials of the symm epilogue,
// M = V^T Y follows from them without a second reduction), and ONE update with K = 4 NB = 128
//   A22 -= [V_a V_b | W_a W_b] [W_a W_b | V_a V_b]^T
// closes the group: the trailing matrix is read 3x and written 1x per TWO panels instead of 4x / 2x, and the update
// runs at 16 flop per byte of trailing-matrix traffic instead of 8 (NumPy model + proof of equivalence:
// tests/twostage_model.py::stage1_deferred).
#include <cooperative_groups.h>

#include <stdlib.h>

#include "band.cuh"

namespace cg = cooperative_groups;

namespace chd {

constexpr int QR_THREADS = 256;
constexpr int QR_S = NB + 1;        // shared-memory row stride of the panel slice
constexpr int QR_MAXCS = 16;
constexpr int QR_XW = NB + 2;       // exchange slot: NB values + (ssq, alpha)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------ H_{-1}
struct Refl0Args {
  const double* ga;   // T x n
  double* V;          // T x n x LDV, columns 0..NB-1 (column 0 = v0, the rest 0)
  double* v0;         // T x n
  double* Tf;         // T x tf_stride (slot 0 = diag(tau0, 0, ...))
  size_t tf_stride;
  double* tau0;       // T
  double* beta0;      // T
  int n;
};

__global__ void __launch_bounds__(256) refl0_kernel(Refl0Args a) {
  __shared__ double red[40];
  const int t = blockIdx.x, tid = threadIdx.x, n = a.n;
  const double* ga = a.ga + (size_t)t * n;
  double ss = 0.0;
  for (int i = 1 + tid; i < n; i += blockDim.x) ss += ga[i] * ga[i];
  ss = block_sum(ss, red);
  const double alpha = ga[0];
  double beta, tau, scale;
  if (ss == 0.0) {
    beta = alpha; tau = 0.0; scale = 0.0;
  } else {
    beta = -copysign(sqrt(alpha * alpha + ss), alpha);
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  double* V = a.V + (size_t)t * n * LDV;
  double* v0 = a.v0 + (size_t)t * n;
  for (int i = tid; i < n; i += blockDim.x) {
    const double v = (i == 0) ? 1.0 : ga[i] * scale;
    v0[i] = v;
  }
  for (int e = tid; e < n * NB; e += blockDim.x) {
    const int i = e / NB, c = e % NB;
    V[(size_t)i * LDV + c] = (c == 0) ? ((i == 0) ? 1.0 : ga[i] * scale) : 0.0;
  }
  double* Tf = a.Tf + (size_t)t * a.tf_stride;
  for (int e = tid; e < NB * NB; e += blockDim.x) Tf[e] = (e == 0) ? tau : 0.0;
  if (tid == 0) { a.tau0[t] = tau; a.beta0[t] = beta; }
}

// ------------------------------------------------------------------ panel QR (cluster kernel)
struct QrArgs {
  double* A; size_t strideA; int ldk;
  double* V;        // T x n x LDV, already offset to this panel's column block (q NB)
  double* Tf;       // T x tf_stride, already offset to this panel's slot
  size_t tf_stride;
  int n, r0, c0;
  int RP;           // panel rows per CTA
  int RS;           // of which kept in shared memory (the rest is worked on in place, in L2)
};

__global__ void __launch_bounds__(QR_THREADS) panel_qr_kernel(QrArgs a) {
  extern __shared__ __align__(16) double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int g = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = QR_THREADS / 32;
  const int n = a.n, r0 = a.r0, c0 = a.c0, ldk = a.ldk, RP = a.RP, RS = a.RS;
  const int m = n - r0;
  const int nref = min(NB, m - 1);
  double* A = a.A + (size_t)t * a.strideA;

  double* xch = smem;                              // 2 x QR_MAXCS x QR_XW
  double* red = xch + 2 * QR_MAXCS * QR_XW;        // NW x QR_XW
  double* mine = red + NW * QR_XW;                 // QR_XW
  double* tot = mine + QR_XW;                      // QR_XW
  double* ws = tot + QR_XW;                        // NB
  double* taus = ws + NB;                          // NB
  double* Gs = taus + NB;                          // NB x QR_S   G[c][t] = v_c . v_t (c < t)
  double* Ts = Gs + NB * QR_S;                     // NB x QR_S
  double* Ps = Ts + NB * QR_S;                     // RS x QR_S

  const int pr0 = g * RP;
  const int nrows = max(0, min(m - pr0, RP));
  double* Ag = A + (size_t)(r0 + pr0) * ldk + c0;  // first owned row, first panel column
  auto rowp = [&](int lr) -> double* { return (lr < RS) ? (Ps + lr * QR_S) : (Ag + (size_t)lr * ldk); };

  for (int e = tid; e < min(nrows, RS) * NB; e += QR_THREADS) {
    const int lr = e / NB, c = e % NB;
    Ps[lr * QR_S + c] = Ag[(size_t)lr * ldk + c];
  }
  for (int e = tid; e < NB * QR_S; e += QR_THREADS) { Gs[e] = 0.0; Ts[e] = 0.0; }
  if (tid < NB) taus[tid] = 0.0;
  __syncthreads();

  int par = 0;
  // exchange `nv` doubles of `mine` with every CTA of the cluster; afterwards tot[] holds the fixed-order sums
  auto exchange = [&](int nv) {
    if (CS == 1) {
      __syncthreads();
      if (tid < nv) tot[tid] = mine[tid];
      __syncthreads();
      return;
    }
    __syncthreads();
    for (int e = tid; e < CS * nv; e += QR_THREADS) {
      const int dst = e / nv, q = e % nv;
      double* remote = cluster.map_shared_rank(xch + (par * QR_MAXCS + g) * QR_XW, dst);
      remote[q] = mine[q];
    }
    cluster.sync();
    if (tid < nv) {
      double s = 0.0;
      for (int q = 0; q < CS; ++q) s += xch[(par * QR_MAXCS + q) * QR_XW + tid];
      tot[tid] = s;
    }
    par ^= 1;
    __syncthreads();
  };

  // norm of column 0 below its diagonal and the diagonal element
  {
    double sq = 0.0, al = 0.0;
    for (int lr = tid; lr < nrows; lr += QR_THREADS) {
      const int pr = pr0 + lr;
      const double x = rowp(lr)[0];
      if (pr > 0) sq += x * x; else al = x;
    }
    sq = warp_sum(sq); al = warp_sum(al);
    if (lane == 0) { red[wid * QR_XW] = sq; red[wid * QR_XW + 1] = al; }
    __syncthreads();
    if (tid < 2) {
      double s = 0.0;
      for (int w = 0; w < NW; ++w) s += red[w * QR_XW + tid];
      mine[tid] = s;
    }
    exchange(2);
  }

  for (int tc = 0; tc < nref; ++tc) {
    const double ssq = tot[0], alpha = tot[1];
    double beta, tau, scale;
    if (ssq == 0.0) {
      beta = alpha; tau = 0.0; scale = 0.0;
    } else {
      beta = -copysign(sqrt(alpha * alpha + ssq), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    __syncthreads();   // everybody has read tot[]
    if (tid == 0) taus[tc] = tau;
    // ---- w[c] = sum_r v[r] P[r][c]   (lanes c > tc); lanes c < tc get G[c][tc] = v_c . v_tc for free
    {
      double acc = 0.0;
      for (int lr = wid; lr < nrows; lr += NW) {
        const int pr = pr0 + lr;
        if (pr < tc) continue;
        const double* row = rowp(lr);
        const double vr = (pr == tc) ? 1.0 : row[tc] * scale;
        // columns c < tc hold v_c in place (rows pr > c; here pr >= tc > c)
        acc = fma(vr, row[lane], acc);
      }
      red[wid * QR_XW + lane] = acc;
      __syncthreads();
      if (tid < NB) {
        double s = 0.0;
        for (int w = 0; w < NW; ++w) s += red[w * QR_XW + tid];
        mine[tid] = s;
      }
      exchange(NB);
      if (tid < NB) {
        ws[tid] = (tid > tc) ? tau * tot[tid] : 0.0;
        if (tid < tc) Gs[tid * QR_S + tc] = tot[tid];
      }
      __syncthreads();
    }
    // ---- P[:, c > tc] -= v w^T, v stored in place; norm / diagonal of the next column on the fly
    {
      const bool more = (tc + 1 < nref);
      double sq = 0.0, al = 0.0;
      const double wl = ws[lane];
      for (int lr = wid; lr < nrows; lr += NW) {
        const int pr = pr0 + lr;
        if (pr < tc) continue;
        double* row = rowp(lr);
        const double vr = (pr == tc) ? 1.0 : row[tc] * scale;
        __syncwarp();
        if (lane > tc) {
          const double x = row[lane] - vr * wl;
          row[lane] = x;
          if (lane == tc + 1) {
            if (pr > tc + 1) sq = fma(x, x, sq);
            else if (pr == tc + 1) al = x;
          }
        } else if (lane == tc) {
          row[tc] = (pr == tc) ? beta : vr;
        }
      }
      if (more) {
        sq = __shfl_sync(0xffffffffu, sq, tc + 1);
        al = __shfl_sync(0xffffffffu, al, tc + 1);
        if (lane == 0) { red[wid * QR_XW] = sq; red[wid * QR_XW + 1] = al; }
        __syncthreads();
        if (tid < 2) {
          double s = 0.0;
          for (int w = 0; w < NW; ++w) s += red[w * QR_XW + tid];
          mine[tid] = s;
        }
        exchange(2);
      } else {
        __syncthreads();
      }
    }
  }

  // ---- compact-WY factor: T[t][t] = tau_t, T[:t, t] = -tau_t T[:t,:t] G[:t, t]   (dlarft, forward columnwise)
  if (wid == 0) {
    for (int tc = 0; tc < nref; ++tc) {
      const double tau = taus[tc];
      double s = 0.0;
      if (lane < tc)
        for (int k = lane; k < tc; ++k) s = fma(Ts[lane * QR_S + k], Gs[k * QR_S + tc], s);
      __syncwarp();
      if (lane < tc) Ts[lane * QR_S + tc] = -tau * s;
      if (lane == tc) Ts[tc * QR_S + tc] = tau;
      __syncwarp();
    }
  }
  __syncthreads();
  if (g == 0) {
    double* Tf = a.Tf + (size_t)t * a.tf_stride;
    for (int e = tid; e < NB * NB; e += QR_THREADS) Tf[e] = Ts[(e / NB) * QR_S + (e % NB)];
  }
  // ---- outputs: the factored block in place (R over V), V as a dense operand
  double* V = a.V + (size_t)t * n * LDV + (size_t)(r0 + pr0) * LDV;
  for (int e = tid; e < nrows * NB; e += QR_THREADS) {
    const int lr = e / NB, c = e % NB, pr = pr0 + lr;
    const double x = rowp(lr)[c];
    if (lr < RS) Ag[(size_t)lr * ldk + c] = x;
    V[(size_t)lr * LDV + c] = (c >= nref || pr < c) ? 0.0 : ((pr == c) ? 1.0 : x);
  }
  if (CS > 1) cluster.sync();   // no CTA may exit while its shared memory can still be written remotely
}

// ------------------------------------------------------------------ panel QR, register-resident variant
// One thread per panel row (the row's NB values live in registers), RG_THREADS rows per CTA, clusters of up to 16
// CTAs.  Per column ONE cluster-wide exchange: the unscaled products  u[c] = sum_{r > tc} x_r P[r][c]  (x = column
// tc) for all NB columns at once -- u[tc] is the squared norm the reflector needs, u[c > tc] gives v^T P, u[c < tc]
// gives the Gram entries v_c . v_tc of the compact-WY factor -- together with row tc itself (from its owner).  The
// row sums over the 32 rows of a warp use a transposing butterfly (31 shuffles for 32 columns).
// TWO = true (panels with more than 16 x 512 rows, i.e. n > 8192): every thread owns a SECOND row (local rows
// RG_THREADS ..) that lives in shared memory in natural column order; its products join the first row's before the
// butterfly, so the exchanges per column stay the same and a 16-CTA cluster covers 16384 rows.  Second rows are never
// pivot rows (their panel row index is >= RG_THREADS > NB).
constexpr int RG_THREADS = 512;
constexpr int RG_NW = RG_THREADS / 32;
constexpr int RG_XW = 2 * NB;
constexpr int RG_S2 = NB + 1;                                    // second-row stride: conflict-free column walks
constexpr size_t RG_SMEM2 = (size_t)RG_THREADS * RG_S2 * sizeof(double);

template <bool TWO>
__global__ void __launch_bounds__(RG_THREADS, 1) panel_qr_reg_kernel(QrArgs a) {
  extern __shared__ __align__(16) double rg_dyn[];               // TWO: RG_THREADS x RG_S2 second rows
  __shared__ __align__(16) double xch[2 * QR_MAXCS * RG_XW];
  __shared__ double red[RG_NW * NB];
  __shared__ double rowbuf[NB];
  __shared__ double tots[2 * RG_XW];
  __shared__ double taus[NB];
  __shared__ double Gs[NB * QR_S];
  cg::cluster_group cluster = cg::this_cluster();
  const int g = (int)cluster.block_rank(), CS = (int)cluster.num_blocks();
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = a.n, r0 = a.r0, c0 = a.c0, ldk = a.ldk;
  const int m = n - r0;
  const int nref = min(NB, m - 1);
  double* A = a.A + (size_t)t * a.strideA;
  const int pr = g * a.RP + tid;
  const bool valid = (tid < a.RP) && (pr < m);
  double* rowg = A + (size_t)(r0 + (valid ? pr : 0)) * ldk + c0;
  const int pr2 = pr + RG_THREADS;
  const bool valid2 = TWO && (tid + RG_THREADS < a.RP) && (pr2 < m);
  double* rowg2 = A + (size_t)(r0 + (valid2 ? pr2 : 0)) * ldk + c0;
  double* PB = rg_dyn + tid * RG_S2;

  // the row lives in registers, ROTATED: after tc columns P[q] holds column (q + tc) mod NB, so that the column
  // being eliminated is always P[0] (static register indices inside a rolled loop: the fully unrolled variant
  // is 49k instructions and instruction-cache bound)
  double P[NB];
#pragma unroll
  for (int c = 0; c < NB; c += 2) {
    double2 v2 = make_double2(0.0, 0.0);
    if (valid) v2 = *reinterpret_cast<const double2*>(rowg + c);
    P[c] = v2.x; P[c + 1] = v2.y;
  }
  if (TWO) {
#pragma unroll
    for (int c = 0; c < NB; c += 2) {
      double2 v2 = make_double2(0.0, 0.0);
      if (valid2) v2 = *reinterpret_cast<const double2*>(rowg2 + c);
      PB[c] = v2.x; PB[c + 1] = v2.y;                            // an absent row is all zeros: it contributes nothing
    }
  }
  for (int e = tid; e < NB * QR_S; e += RG_THREADS) Gs[e] = 0.0;
  if (tid < NB) taus[tid] = 0.0;

  int par = 0;
  __syncthreads();

#pragma unroll 1
  for (int tc = 0; tc < NB; ++tc) {
    if (tc < nref) {
      const double x = (valid && pr > tc) ? P[0] : 0.0;
      const double xb = TWO ? PB[tc] : 0.0;
      // ---- transposing butterfly: lane q ends up with sum over the warp's rows of x * P[q]
      double t16[16], t8[8], t4[4], t2[2], t1;
      {
        const bool hi = lane & 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          double a0 = x * P[c], a1 = x * P[c + 16];
          if (TWO) {                                             // P is rotated by tc, the second row is not
            a0 = fma(xb, PB[(c + tc) & (NB - 1)], a0);
            a1 = fma(xb, PB[(c + 16 + tc) & (NB - 1)], a1);
          }
          const double keep = hi ? a1 : a0, send = hi ? a0 : a1;
          t16[c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool hi = lane & 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const double keep = hi ? t16[c + 8] : t16[c], send = hi ? t16[c] : t16[c + 8];
          t8[c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool hi = lane & 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double keep = hi ? t8[c + 4] : t8[c], send = hi ? t8[c] : t8[c + 4];
          t4[c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
      }
      {
        const bool hi = lane & 2;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const double keep = hi ? t4[c + 2] : t4[c], send = hi ? t4[c] : t4[c + 2];
          t2[c] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
      }
      {
        const bool hi = lane & 1;
        const double keep = hi ? t2[1] : t2[0], send = hi ? t2[0] : t2[1];
        t1 = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
      red[wid * NB + lane] = t1;
      if (valid && pr == tc) {
#pragma unroll
        for (int q = 0; q < NB; ++q) rowbuf[(q + tc) & (NB - 1)] = P[q];   // row tc, natural column order
      }
      __syncthreads();
      // 64 values per CTA go round the cluster: u[c] (threads 0..31) and row tc from its owner (threads 32..63);
      // written straight into every CTA's slot, summed in fixed order after ONE cluster barrier
      double* tot = tots + par * RG_XW;
      if (tid < RG_XW) {
        double val;
        int slot;
        if (tid < NB) {
          double sum = 0.0;
#pragma unroll
          for (int w = 0; w < RG_NW; ++w) sum += red[w * NB + tid];
          val = sum;
          slot = (tid + tc) & (NB - 1);                // back to natural column indices
        } else {
          val = (g == tc / a.RP) ? rowbuf[tid - NB] : 0.0;
          slot = tid;
        }
        if (CS == 1) {
          tot[slot] = val;
        } else {
          for (int dst = 0; dst < CS; ++dst)
            cluster.map_shared_rank(xch + (par * QR_MAXCS + g) * RG_XW, dst)[slot] = val;
        }
      }
      if (CS == 1) {
        __syncthreads();
      } else {
        cluster.sync();
        if (tid < RG_XW) {
          double sum = 0.0;
          for (int q = 0; q < CS; ++q) sum += xch[(par * QR_MAXCS + q) * RG_XW + tid];
          tot[tid] = sum;
        }
        __syncthreads();
      }
      par ^= 1;
      const double ssq = tot[tc], alpha = tot[NB + tc];
      double beta, tau, scale;
      if (ssq == 0.0) {
        beta = alpha; tau = 0.0; scale = 0.0;
      } else {
        beta = -copysign(sqrt(alpha * alpha + ssq), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      const double v = (!valid || pr < tc) ? 0.0 : ((pr == tc) ? 1.0 : x * scale);
      const double tv = tau * v;
      const double vb = xb * scale, tvb = tau * vb;
#pragma unroll
      for (int q = 1; q < NB; ++q) {
        if (q < NB - tc) {                        // columns c = tc + q > tc still to be eliminated
          const int c = tc + q;
          const double f = fma(scale, tot[c], tot[NB + c]);
          P[q] = fma(-tv, f, P[q]);
          if (TWO) PB[c] = fma(-tvb, f, PB[c]);
        }
      }
      if (valid && pr >= tc) P[0] = (pr == tc) ? beta : v;
      if (TWO) PB[tc] = vb;
      if (wid == 0) {
        if (lane < tc) Gs[lane * QR_S + tc] = fma(scale, tot[lane], tot[NB + lane]);
        if (lane == 0) taus[tc] = tau;
      }
      // no barrier here: tot is double buffered, red / rowbuf are rewritten only after this column's barriers
    }
    // rotate: the finished column moves to the end
    {
      const double p0 = P[0];
#pragma unroll
      for (int q = 0; q < NB - 1; ++q) P[q] = P[q + 1];
      P[NB - 1] = p0;
    }
  }

  __syncthreads();
  // ---- compact-WY factor (dlarft, forward columnwise): lane a owns row a of T
  if (wid == 0 && g == 0) {
    double Tr[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) Tr[k] = 0.0;
#pragma unroll
    for (int tc = 0; tc < NB; ++tc) {
      const double tau = taus[tc];
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < tc; ++k) s = fma(Tr[k], Gs[k * QR_S + tc], s);   // Tr[k] == 0 for k < lane
      Tr[tc] = (lane < tc) ? -tau * s : ((lane == tc) ? tau : 0.0);
    }
    double* Tf = a.Tf + (size_t)t * a.tf_stride;
#pragma unroll
    for (int k = 0; k < NB; ++k) Tf[lane * NB + k] = Tr[k];
  }
  // ---- outputs: the factored block in place (R over V), V as a dense operand
  if (valid) {
    double* V = a.V + (size_t)t * n * LDV + (size_t)(r0 + pr) * LDV;
#pragma unroll
    for (int c = 0; c < NB; c += 2) {
      *reinterpret_cast<double2*>(rowg + c) = make_double2(P[c], P[c + 1]);
      const double v0 = (c >= nref || pr < c) ? 0.0 : ((pr == c) ? 1.0 : P[c]);
      const double v1 = (c + 1 >= nref || pr < c + 1) ? 0.0 : ((pr == c + 1) ? 1.0 : P[c + 1]);
      *reinterpret_cast<double2*>(V + c) = make_double2(v0, v1);
    }
  }
  if (TWO && valid2) {
    double* V = a.V + (size_t)t * n * LDV + (size_t)(r0 + pr2) * LDV;
#pragma unroll
    for (int c = 0; c < NB; c += 2) {
      *reinterpret_cast<double2*>(rowg2 + c) = make_double2(PB[c], PB[c + 1]);
      *reinterpret_cast<double2*>(V + c) = make_double2(c < nref ? PB[c] : 0.0, c + 1 < nref ? PB[c + 1] : 0.0);
    }
  }
  if (CS > 1) cluster.sync();
}

// ------------------------------------------------------------------ Y = A22 V
constexpr int SY_ST = SY_BM + 8;    // transposed tile [k][i] stride (== 8 mod 16)
constexpr int SY_SV = NB + 8;       // V chunk [k][c] stride (== 8 mod 16)
// per k-chunk width KC (32: two CTAs per SM; 64: one CTA per SM, half the barriers per DMMA)
template <int KC> struct SymmCfg {
  static constexpr int SD = KC + 4;   // direct tile [i][k] stride (== 4 mod 8: conflict-free fragment loads)
  static constexpr int TILE = (SY_BM * SD > KC * SY_ST) ? SY_BM * SD : KC * SY_ST;
  static constexpr int STAGE = TILE + KC * SY_SV;
};
constexpr int SY_NST = 2;

// one staged k-chunk (SY_KC columns) of the row block: direct tile [i][k], transposed tile [k][i], optionally
// masked against the diagonal (only the chunks that cross it)
template <bool TR, bool MK, int KC>
__device__ __forceinline__ void symm_chunk(double (&acc)[2][4][2], const double* __restrict__ tile,
                                           const double* __restrict__ vs, int wid, int fr, int fk, int i0, int j0) {
  constexpr int SD = SymmCfg<KC>::SD;
  const double* ap = TR ? tile + fk * SY_ST + 16 * wid + fr : tile + (16 * wid + fr) * SD + fk;
  const double* vp = vs + fk * SY_SV + fr;
#pragma unroll
  for (int kk = 0; kk < KC; kk += 4) {
    double af[2], bf[4];
#pragma unroll
    for (int y = 0; y < 4; ++y) bf[y] = vp[kk * SY_SV + y * 8];
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      af[x] = TR ? ap[kk * SY_ST + 8 * x] : ap[8 * x * SD + kk];
      if (MK) {
        const int i = i0 + 16 * wid + 8 * x + fr, j = j0 + kk + fk;
        if (TR ? (j <= i) : (j > i)) af[x] = 0.0;
      }
    }
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
  }
}

template <int KC>
__global__ void __launch_bounds__(256, KC == 32 ? 2 : 1) symm_kernel(SymmArgs a) {
  constexpr int SY_KC = KC, SY_SD = SymmCfg<KC>::SD, SY_TILE = SymmCfg<KC>::TILE, SY_STAGE = SymmCfg<KC>::STAGE;
  extern __shared__ __align__(16) double smem[];
  const int t = blockIdx.y, bx = blockIdx.x, sp = blockIdx.z, S = gridDim.z;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = a.n, r0 = a.r0, ldk = a.ldk;
  const double* A = a.A + (size_t)t * a.strideA;
  const double* V = a.V + (size_t)t * n * LDV;
  const int i0 = r0 + SY_BM * bx;
  const int m = n - r0;
  const int nq = (m + SY_KC - 1) / SY_KC;
  const int qd0 = (i0 - r0) / SY_KC;
  const int nd = min(SY_BM / SY_KC, nq - qd0);
  const int niter = nq + nd;
  // split-K: this CTA handles iterations [it_lo, it_hi) of the row block; the S partial Y's are summed by wform
  const int it_lo = (int)((long long)niter * sp / S), it_hi = (int)((long long)niter * (sp + 1) / S);

  // iteration -> (chunk, transposed, masked)
  auto decode = [&](int it, int& q, bool& tr, bool& mk) {
    if (it < qd0) { q = it; tr = false; mk = false; }
    else if (it < qd0 + nd) { q = it; tr = false; mk = true; }
    else if (it < qd0 + 2 * nd) { q = it - nd; tr = true; mk = true; }
    else { q = it - nd; tr = true; mk = false; }
  };
  // per-thread copy geometry (fixed over the loop)
  constexpr int DPR = KC / 2, DRS = 256 / DPR;    // 16-byte chunks per direct-tile row, rows per pass
  const int drow = tid / DPR, dpart = tid % DPR;  // direct tile: rows drow + DRS s, 16-byte chunk dpart
  const int vrow = tid >> 4, vpart = tid & 15;    // V chunk: rows vrow + 16 s
  const int trow = tid >> 6, tpart = tid & 63;    // transposed tile: rows trow + 4 s, chunk tpart
  const double* dsrc = A + (size_t)(i0 + drow) * ldk + 2 * dpart;
  const double* tsrc = A + (size_t)trow * ldk + i0 + 2 * tpart;
  const bool tcol_ok = i0 + 2 * tpart < n;
  auto issue = [&](int it) {
    int q; bool tr, mk;
    decode(it, q, tr, mk);
    (void)mk;
    double* tile = smem + (size_t)((it - it_lo) % SY_NST) * SY_STAGE;
    double* vs = tile + SY_TILE;
    const int j0 = r0 + SY_KC * q;
    if (!tr) {
      const bool cok = j0 + 2 * dpart < n;
#pragma unroll
      for (int s = 0; s < SY_BM / DRS; ++s) {
        const bool ok = cok && (i0 + drow + DRS * s < n);
        cp_async16(tile + (drow + DRS * s) * SY_SD + 2 * dpart, ok ? dsrc + (size_t)DRS * s * ldk + j0 : A, ok);
      }
    } else {
#pragma unroll
      for (int s = 0; s < KC / 4; ++s) {
        const bool ok = tcol_ok && (j0 + trow + 4 * s < n);
        cp_async16(tile + (trow + 4 * s) * SY_ST + 2 * tpart, ok ? tsrc + (size_t)(j0 + 4 * s) * ldk : A, ok);
      }
    }
#pragma unroll
    for (int s = 0; s < KC / 16; ++s) {
      const int row = vrow + 16 * s;
      const bool ok = (j0 + row < n);
      cp_async16(vs + row * SY_SV + 2 * vpart, ok ? V + (size_t)(j0 + row) * LDV + 2 * vpart : V, ok);
    }
  };

  double acc[2][4][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;

  if (it_lo < it_hi) issue(it_lo);
  cp_async_commit();
  for (int it = it_lo; it < it_hi; ++it) {
    if (it + 1 < it_hi) issue(it + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    int q; bool tr, mk;
    decode(it, q, tr, mk);
    const double* tile = smem + (size_t)((it - it_lo) % SY_NST) * SY_STAGE;
    const double* vs = tile + SY_TILE;
    const int j0 = r0 + SY_KC * q;
    if (!mk) {
      if (!tr) symm_chunk<false, false, KC>(acc, tile, vs, wid, fr, fk, i0, j0);
      else symm_chunk<true, false, KC>(acc, tile, vs, wid, fr, fk, i0, j0);
    } else {
      if (!tr) symm_chunk<false, true, KC>(acc, tile, vs, wid, fr, fk, i0, j0);
      else symm_chunk<true, true, KC>(acc, tile, vs, wid, fr, fk, i0, j0);
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // ---- epilogue: the Y rows of this split-K partial (the products with V / B follow in symm_partials_kernel)
  symm_store_y(a, acc, tid, t, bx, sp, (int)gridDim.y);
}

// ------------------------------------------------------------------ partial products of a row block
// One CTA per 128-row block: sums the split-K partials of Y (result back into buffer 0), then
//   Mpart = V_I^T Y_I,  Zpart = V_I^T B_I,  and for the second panel of a group  Gpart = {W_a,I^T V_I, V_a,I^T V_I};
// reduce_kernel adds the blocks in fixed order (deterministic, no atomics).
__global__ void __launch_bounds__(256) symm_partials_kernel(SymmArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* Ys = smem;                       // SY_BM x (NB + 1)
  double* Vs = Ys + SY_BM * (NB + 1);      // SY_BM x (NB + 1)
  const int t = blockIdx.y, bx = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = a.n, i0 = a.r0 + SY_BM * bx;
  const double* V = a.V + (size_t)t * n * LDV;
  // blockIdx.z = 1: the Z-block product V_I^T B_I only (its own CTA, so that its long load chain overlaps with the
  // Y / Gram half of the same row block)
  if (blockIdx.z == 1) {
    if (!(a.Bq && a.nz > 0)) return;
    for (int e = tid; e < SY_BM * NB; e += 256) {
      const int rl = e / NB, c = e % NB;
      Vs[rl * (NB + 1) + c] = (i0 + rl < n) ? V[(size_t)(i0 + rl) * LDV + c] : 0.0;
    }
    const double* Bq = a.Bq + (size_t)t * n * a.ldb;
    const int z = tid & 127, c1g = tid >> 7;   // 2 groups of 16 reflector columns
    double zp[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) zp[u] = 0.0;
    // batches of 16 rows, the next batch's loads in flight while the current one is multiplied
    double bv[16], nx[16];
    auto fetch = [&](int rb, double (&dst)[16]) {
#pragma unroll
      for (int q = 0; q < 16; ++q)
        dst[q] = (z < a.nz && i0 + rb + q < n) ? Bq[(size_t)(i0 + rb + q) * a.ldb + z] : 0.0;
    };
    fetch(0, bv);
    __syncthreads();
#pragma unroll 1
    for (int rb = 0; rb < SY_BM; rb += 16) {
      if (rb + 16 < SY_BM) fetch(rb + 16, nx);
#pragma unroll
      for (int q = 0; q < 16; ++q)
#pragma unroll
        for (int u = 0; u < 16; ++u) zp[u] = fma(Vs[(rb + q) * (NB + 1) + 16 * c1g + u], bv[q], zp[u]);
#pragma unroll
      for (int q = 0; q < 16; ++q) bv[q] = nx[q];
    }
    if (z < a.nz) {
      double* Zp = a.Zpart + ((size_t)t * a.nblk + bx) * NB * SY_ZW;
#pragma unroll
      for (int u = 0; u < 16; ++u) Zp[(16 * c1g + u) * SY_ZW + z] = zp[u];
    }
    return;
  }
  double* Y0 = a.Y + (size_t)t * n * NB;
  for (int e = tid; e < SY_BM * NB; e += 256) {
    const int rl = e / NB, c = e % NB;
    const bool ok = i0 + rl < n;
    double yv = 0.0;
    if (ok) {
      for (int sp = 0; sp < a.nsplit; ++sp) yv += a.Y[((size_t)sp * a.T + t) * n * NB + (size_t)(i0 + rl) * NB + c];
      if (a.nsplit > 1) Y0[(size_t)(i0 + rl) * NB + c] = yv;
    }
    Ys[rl * (NB + 1) + c] = yv;
    Vs[rl * (NB + 1) + c] = ok ? V[(size_t)(i0 + rl) * LDV + c] : 0.0;
  }
  __syncthreads();
  {
    double mp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int rl = 0; rl < SY_BM; ++rl) {
      const double yv = Ys[rl * (NB + 1) + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) mp[u] = fma(Vs[rl * (NB + 1) + wid + 8 * u], yv, mp[u]);
    }
    double* Mp = a.Mpart + ((size_t)t * a.nblk + bx) * NB * NB;
#pragma unroll
    for (int u = 0; u < 4; ++u) Mp[(wid + 8 * u) * NB + lane] = mp[u];
  }
  if (a.Gpart) {
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const double* src = (which == 0 ? a.Wa : a.Va) + (size_t)t * n * LDV;
      __syncthreads();                       // Ys (Y rows, then W_a rows) has been consumed
      for (int e = tid; e < SY_BM * NB; e += 256) {
        const int rl = e / NB, c = e % NB;
        Ys[rl * (NB + 1) + c] = (i0 + rl < n) ? src[(size_t)(i0 + rl) * LDV + c] : 0.0;
      }
      __syncthreads();
      double gp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
      for (int rl = 0; rl < SY_BM; ++rl) {
        const double vv = Vs[rl * (NB + 1) + lane];
#pragma unroll
        for (int u = 0; u < 4; ++u) gp[u] = fma(Ys[rl * (NB + 1) + wid + 8 * u], vv, gp[u]);
      }
      double* Gp = a.Gpart + (((size_t)t * a.nblk + bx) * 2 + which) * NB * NB;
#pragma unroll
      for (int u = 0; u < 4; ++u) Gp[(wid + 8 * u) * NB + lane] = gp[u];
    }
  }
}
constexpr size_t SP_SMEM = (size_t)2 * SY_BM * (NB + 1) * sizeof(double);

// ------------------------------------------------------------------ fixed-order reduction of the partials
struct ReduceArgs {
  const double* Mpart; const double* Zpart; const double* Gpart;
  double* Mred;   // T x NB x NB
  double* Zred;   // T x NB x SY_ZW
  double* Gred;   // T x 2 x NB x NB
  int nblk, has_z, has_g;
};

__global__ void __launch_bounds__(256) reduce_kernel(ReduceArgs a) {
  const int t = blockIdx.y;
  const int e = blockIdx.x * 256 + threadIdx.x;
  constexpr int NM = NB * NB, NZ = NB * SY_ZW;
  if (e < NM) {
    const double* p = a.Mpart + (size_t)t * a.nblk * NM + e;
    double s = 0.0;
    for (int b = 0; b < a.nblk; ++b) s += p[(size_t)b * NM];
    a.Mred[(size_t)t * NM + e] = s;
  } else if (e < NM + NZ) {
    if (!a.has_z) return;
    const int ez = e - NM;
    const double* p = a.Zpart + (size_t)t * a.nblk * NZ + ez;
    double s = 0.0;
    for (int b = 0; b < a.nblk; ++b) s += p[(size_t)b * NZ];
    a.Zred[(size_t)t * NZ + ez] = s;
  } else if (e < NM + NZ + 2 * NM && a.has_g) {
    const int eg = e - NM - NZ;
    const double* p = a.Gpart + (size_t)t * a.nblk * 2 * NM + eg;
    double s = 0.0;
    for (int b = 0; b < a.nblk; ++b) s += p[(size_t)b * 2 * NM];
    a.Gred[(size_t)t * 2 * NM + eg] = s;
  }
}

// ------------------------------------------------------------------ the panel's small matrices (one CTA per matrix)
//   M  = V^T Y            (as reduced; second panel of a group: M -= G2^T G1 + G1^T G2, the stale-matrix correction)
//   TM = T^T M,  TZ = T^T (V^T B)
// computed ONCE per matrix and panel (the row-block kernels below only stream)
struct SmallArgs {
  const double* Tf; size_t tf_stride;      // this panel's slot
  const double* Mred; const double* Zred; const double* Gred;
  double* TM;     // T x NB x NB
  double* TZ;     // T x NB x SY_ZW
  int has_z, has_g;
};

__global__ void __launch_bounds__(256) smallmat_kernel(SmallArgs a) {
  __shared__ double Ts[NB * (NB + 1)], Ms[NB * (NB + 1)], G1s[NB * (NB + 1)], G2s[NB * (NB + 1)];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double* Tf = a.Tf + (size_t)t * a.tf_stride;
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    Ts[r * (NB + 1) + c] = Tf[e];
    Ms[r * (NB + 1) + c] = a.Mred[(size_t)t * NB * NB + e];
    if (a.has_g) {
      G1s[r * (NB + 1) + c] = a.Gred[(size_t)t * 2 * NB * NB + e];
      G2s[r * (NB + 1) + c] = a.Gred[(size_t)t * 2 * NB * NB + NB * NB + e];
    }
  }
  __syncthreads();
  if (a.has_g) {
    // M[r][c] -= sum_k G2[k][r] G1[k][c] + G1[k][r] G2[k][c]      (rows r = wid + 8 u, column c = lane)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int k = 0; k < NB; ++k) {
      const double g1c = G1s[k * (NB + 1) + lane], g2c = G2s[k * (NB + 1) + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        acc[u] = fma(G2s[k * (NB + 1) + wid + 8 * u], g1c, fma(G1s[k * (NB + 1) + wid + 8 * u], g2c, acc[u]));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) Ms[(wid + 8 * u) * (NB + 1) + lane] -= acc[u];
    __syncthreads();
  }
  {
    // TM[r][c] = sum_k T[k][r] M[k][c]   (T upper triangular with explicit zeros)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int k = 0; k < NB; ++k) {
      const double mv = Ms[k * (NB + 1) + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fma(Ts[k * (NB + 1) + wid + 8 * u], mv, acc[u]);
    }
    double* TM = a.TM + (size_t)t * NB * NB;
#pragma unroll
    for (int u = 0; u < 4; ++u) TM[(wid + 8 * u) * NB + lane] = acc[u];
  }
  if (a.has_z) {
    const double* Zr = a.Zred + (size_t)t * NB * SY_ZW;
    double* TZ = a.TZ + (size_t)t * NB * SY_ZW;
    const int z = tid & 127, hg = tid >> 7;
    double acc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) acc[u] = 0.0;
#pragma unroll 2
    for (int k = 0; k < NB; ++k) {
      const double zv = Zr[k * SY_ZW + z];
#pragma unroll
      for (int u = 0; u < 16; ++u) acc[u] = fma(Ts[k * (NB + 1) + hg + 2 * u], zv, acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) TZ[(hg + 2 * u) * SY_ZW + z] = acc[u];
  }
}

// ------------------------------------------------------------------ W = (Y - 1/2 V T^T M) T  per 64-row block
struct WformArgs {
  const double* V;      // T x n x LDV, this panel's column block
  const double* Y;      // nsplit partial buffers of T x n x NB
  double* W;            // T x n x LDV, this panel's column block
  int nsplit, T;
  const double* Tf; size_t tf_stride;            // this panel's slot
  const double* TM;     // T x NB x NB
  const double* Gred;   // T x 2 x NB x NB  (G1, G2) with Va / Wa: stale-matrix correction, or null
  const double* Va; const double* Wa;            // T x n x LDV, column block 0
  int n, r0;
};

constexpr int WF_BM = 64;
constexpr int WF_RS = NB + 1;
constexpr size_t WF_SMEM = (size_t)(2 * NB * WF_RS + 2 * WF_BM * WF_RS) * sizeof(double);

// out[r][c] (rows wid + 8 q, column lane) += sign * sum_k R[r][k] Mx[k][c]
__device__ __forceinline__ void wf_mac(double (&acc)[WF_BM / 8], const double* __restrict__ R,
                                       const double* __restrict__ Mx, int wid, int lane, double sign) {
  double part[WF_BM / 8];
#pragma unroll
  for (int q = 0; q < WF_BM / 8; ++q) part[q] = 0.0;
#pragma unroll 4
  for (int k = 0; k < NB; ++k) {
    const double mv = Mx[k * WF_RS + lane];
#pragma unroll
    for (int q = 0; q < WF_BM / 8; ++q) part[q] = fma(R[(wid + 8 * q) * WF_RS + k], mv, part[q]);
  }
#pragma unroll
  for (int q = 0; q < WF_BM / 8; ++q) acc[q] = fma(sign, part[q], acc[q]);
}

__global__ void __launch_bounds__(256) wform_kernel(WformArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* Mx0 = smem;                       // NB x WF_RS  (two matrix slots, two row-operand slots: every phase
  double* Mx1 = Mx0 + NB * WF_RS;           //  reads one pair while the next pair is being filled)
  double* R0 = Mx1 + NB * WF_RS;            // WF_BM x WF_RS
  double* R1 = R0 + WF_BM * WF_RS;
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = a.n, i0 = a.r0 + WF_BM * blockIdx.x;
  const bool corr = a.Gred != nullptr;
  const double* Tf = a.Tf + (size_t)t * a.tf_stride;
  const double* TM = a.TM + (size_t)t * NB * NB;
  const double* V = a.V + (size_t)t * n * LDV;
  double* W = a.W + (size_t)t * n * LDV;
  auto load_mat = [&](double* dst, const double* src) {
    for (int e = tid; e < NB * NB; e += 256) dst[(e / NB) * WF_RS + (e % NB)] = src[e];
  };
  auto load_rows = [&](double* dst, const double* src) {      // src: T x n x LDV operand, rows i0 ..
    for (int e = tid; e < WF_BM * NB; e += 256) {
      const int rl = e / NB, c = e % NB;
      dst[rl * WF_RS + c] = (i0 + rl < n) ? src[(size_t)(i0 + rl) * LDV + c] : 0.0;
    }
  };
  // accumulators start at the summed split-K partials of Y
  double acc[WF_BM / 8];
#pragma unroll
  for (int q = 0; q < WF_BM / 8; ++q) {
    const int i = i0 + wid + 8 * q;
    double yv = 0.0;
    if (i < n)
      for (int sp = 0; sp < a.nsplit; ++sp) yv += a.Y[((size_t)sp * a.T + t) * n * NB + (size_t)i * NB + lane];
    acc[q] = yv;
  }
  if (corr) {
    const double* G = a.Gred + (size_t)t * 2 * NB * NB;
    load_mat(Mx0, G);                                     // G1 = W_a^T V
    load_rows(R0, a.Va + (size_t)t * n * LDV);
    load_mat(Mx1, G + NB * NB);                           // G2 = V_a^T V
    load_rows(R1, a.Wa + (size_t)t * n * LDV);
    __syncthreads();
    wf_mac(acc, R0, Mx0, wid, lane, -1.0);                // Y -= V_a G1
    wf_mac(acc, R1, Mx1, wid, lane, -1.0);                // Y -= W_a G2
    __syncthreads();
  }
  load_mat(Mx0, TM);
  load_rows(R0, V);
  load_mat(Mx1, Tf);
  __syncthreads();
  wf_mac(acc, R0, Mx0, wid, lane, -0.5);                  // X1 = Y - 1/2 V (T^T M)
#pragma unroll
  for (int q = 0; q < WF_BM / 8; ++q) R1[(wid + 8 * q) * WF_RS + lane] = acc[q];
  __syncthreads();
  double wv[WF_BM / 8];
#pragma unroll
  for (int q = 0; q < WF_BM / 8; ++q) wv[q] = 0.0;
  wf_mac(wv, R1, Mx1, wid, lane, 1.0);                    // W = X1 T
#pragma unroll
  for (int q = 0; q < WF_BM / 8; ++q) {
    const int i = i0 + wid + 8 * q;
    if (i < n) W[(size_t)i * LDV + lane] = wv[q];
  }
}

// ------------------------------------------------------------------ B -= V (T^T V^T B)  per 64-row block
struct ZupdArgs {
  const double* V;      // T x n x LDV, this panel's column block
  const double* TZ;     // T x NB x SY_ZW
  double* Bq; int ldb, nz;
  int n, r0;
};
constexpr size_t ZU_SMEM = (size_t)(NB * SY_ZW + WF_BM * WF_RS) * sizeof(double);

__global__ void __launch_bounds__(256) zupdate_kernel(ZupdArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* TZs = smem;                       // NB x SY_ZW
  double* Vs = TZs + NB * SY_ZW;            // WF_BM x WF_RS
  const int t = blockIdx.y, tid = threadIdx.x;
  const int n = a.n, i0 = a.r0 + WF_BM * blockIdx.x;
  const double* V = a.V + (size_t)t * n * LDV;
  const double* TZ = a.TZ + (size_t)t * NB * SY_ZW;
  double* Bq = a.Bq + (size_t)t * n * a.ldb;
  const int z = tid & 127, hg = tid >> 7;
  // batches of 8 rows (8 independent chains); the next batch's loads are in flight while the current one is multiplied,
  // the first batch's while the operands are staged
  double bv[8], nx[8];
  auto fetch = [&](int ub, double (&dst)[8]) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + hg + 2 * (ub + u);
      dst[u] = (z < a.nz && i < n) ? Bq[(size_t)i * a.ldb + z] : 0.0;
    }
  };
  fetch(0, bv);
  for (int e = tid; e < NB * SY_ZW; e += 256) TZs[e] = TZ[e];
  for (int e = tid; e < WF_BM * NB; e += 256) {
    const int rl = e / NB, c = e % NB;
    Vs[rl * WF_RS + c] = (i0 + rl < n) ? V[(size_t)(i0 + rl) * LDV + c] : 0.0;
  }
  __syncthreads();
  if (z >= a.nz) return;
#pragma unroll 1
  for (int ub = 0; ub < WF_BM / 2; ub += 8) {
    if (ub + 8 < WF_BM / 2) fetch(ub + 8, nx);
    double sv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) sv[u] = 0.0;
#pragma unroll 4
    for (int k = 0; k < NB; ++k) {
      const double zv = TZs[k * SY_ZW + z];
#pragma unroll
      for (int u = 0; u < 8; ++u) sv[u] = fma(Vs[(hg + 2 * (ub + u)) * WF_RS + k], zv, sv[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + hg + 2 * (ub + u);
      if (i < n) Bq[(size_t)i * a.ldb + z] = bv[u] - sv[u];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) bv[u] = nx[u];
  }
}

// ------------------------------------------------------------------ the same small products on the FP64 tensor cores
// The per-panel products of a row block with 32 x 32 / 32 x 100 matrices are GEMM-shaped (K = 32 or K = the block's
// rows); as FMA loops they ran at 8 TFLOP/s with the FP64 pipe as the limiter (ncu: 27 % FP64 pipe, 15-25 % of the DRAM
// bandwidth).  Here they are DMMA m8n8k4 chains: a warp owns 8 rows (wform_dmma_kernel, zupd_dmma_kernel) or a pair / a strip of 8 x 8
// output tiles (partials_dmma_kernel); 32 x 32 operands sit in shared memory with stride NB + 4 and the Z operand with
// stride SY_ZW + 4 (== 4 mod 16: the four k-rows of a B fragment land in different bank groups), row operands are
// read from global memory straight into fragment position.  CHD_SMALL_DMMA=0 keeps the FMA kernels.
constexpr int DM_S = NB + 4;
constexpr int DM_SZ = SY_ZW + 4;
constexpr size_t WZ_SMEM = (size_t)(4 * NB * DM_S) * sizeof(double);
constexpr size_t PD_SMEM = (size_t)(2 * SY_BM * DM_S) * sizeof(double);

// B -= V (T^T V^T B) for one 64-row block: warp w owns rows 8 w .. 8 w + 7, blockIdx.z picks 7 of the 13 column tiles
// (28 accumulator registers instead of 52: four CTAs per SM -- the kernel is one round of global loads, one DMMA
// chain and one round of stores per warp, so it lives on occupancy).
constexpr int ZU_T = 7;                           // column tiles per CTA
constexpr int ZU_S = 8 * ZU_T + 4;                // stride of the staged 32 x 56 piece of T^T V^T B (== 12 mod 16)
__global__ void __launch_bounds__(256, 4) zupd_dmma_kernel(ZupdArgs zu) {
  __shared__ __align__(16) double TZs[NB * ZU_S];
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = zu.n, i = zu.r0 + WF_BM * blockIdx.x + 8 * wid + fr;
  const bool ok = i < n;
  const int cb = 8 * ZU_T * blockIdx.z;           // first column of this CTA
  const double* V = zu.V + (size_t)t * n * LDV;
  double* Brow = zu.Bq + (size_t)t * n * zu.ldb + (size_t)i * zu.ldb + cb;
  double zc[ZU_T][2];
#pragma unroll
  for (int y = 0; y < ZU_T; ++y) {
    const int c = cb + 8 * y + 2 * fk;
    double2 bv = make_double2(0.0, 0.0);
    if (ok && c < zu.nz) bv = *reinterpret_cast<const double2*>(Brow + 8 * y + 2 * fk);
    zc[y][0] = bv.x; zc[y][1] = bv.y;
  }
  double av[NB / 4];
#pragma unroll
  for (int s_ = 0; s_ < NB / 4; ++s_) av[s_] = ok ? -V[(size_t)i * LDV + 4 * s_ + fk] : 0.0;
  const double* TZ = zu.TZ + (size_t)t * NB * SY_ZW + cb;
  for (int e = tid; e < NB * (4 * ZU_T); e += 256) {     // 32 rows x 28 double2
    const int k = e / (4 * ZU_T), c = 2 * (e % (4 * ZU_T));
    *reinterpret_cast<double2*>(TZs + k * ZU_S + c) = *reinterpret_cast<const double2*>(TZ + k * SY_ZW + c);
  }
  __syncthreads();
#pragma unroll
  for (int s_ = 0; s_ < NB / 4; ++s_)
#pragma unroll
    for (int y = 0; y < ZU_T; ++y) dmma884(zc[y][0], zc[y][1], av[s_], TZs[(4 * s_ + fk) * ZU_S + 8 * y + fr]);
  if (ok) {
#pragma unroll
    for (int y = 0; y < ZU_T; ++y) {
      const int c = cb + 8 * y + 2 * fk;
      if (c < zu.nz) *reinterpret_cast<double2*>(Brow + 8 * y + 2 * fk) = make_double2(zc[y][0], zc[y][1]);
    }
  }
}

// W = (Y - V_a G1 - W_a G2 - 1/2 V T^T M) T for one 64-row block: warp w owns rows 8 w .. 8 w + 7
__global__ void __launch_bounds__(256, 3) wform_dmma_kernel(WformArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = a.n, i = a.r0 + WF_BM * blockIdx.x + 8 * wid + fr;
  const bool ok = i < n;
  const double* V = a.V + (size_t)t * n * LDV;
  // this lane's A fragments of V (row i, k = 4 s + fk)
  double av[NB / 4];
#pragma unroll
  for (int s_ = 0; s_ < NB / 4; ++s_) av[s_] = ok ? V[(size_t)i * LDV + 4 * s_ + fk] : 0.0;
  double* G1s = smem;                       // NB x DM_S each
  double* G2s = G1s + NB * DM_S;
  double* TMs = G2s + NB * DM_S;
  double* Ts = TMs + NB * DM_S;
  const bool corr = a.Gred != nullptr;
  // accumulators (row i, columns 8 y + 2 fk + {0, 1}) = Y
  double acc[NB / 8][2];
#pragma unroll
  for (int y = 0; y < NB / 8; ++y) {
    acc[y][0] = 0.0; acc[y][1] = 0.0;
    if (ok)
      for (int sp = 0; sp < a.nsplit; ++sp) {
        const double2 yv = *reinterpret_cast<const double2*>(a.Y + ((size_t)sp * a.T + t) * n * NB + (size_t)i * NB +
                                                             8 * y + 2 * fk);
        acc[y][0] += yv.x; acc[y][1] += yv.y;
      }
  }
  double a1[NB / 4], a2[NB / 4];
  if (corr) {
    const double* Va = a.Va + (size_t)t * n * LDV;
    const double* Wa = a.Wa + (size_t)t * n * LDV;
#pragma unroll
    for (int s_ = 0; s_ < NB / 4; ++s_) {
      a1[s_] = ok ? -Va[(size_t)i * LDV + 4 * s_ + fk] : 0.0;
      a2[s_] = ok ? -Wa[(size_t)i * LDV + 4 * s_ + fk] : 0.0;
    }
  }
  {
    const double* Tf = a.Tf + (size_t)t * a.tf_stride;
    const double* TM = a.TM + (size_t)t * NB * NB;
    const double* G = corr ? a.Gred + (size_t)t * 2 * NB * NB : nullptr;
    for (int e = tid; e < NB * NB; e += 256) {
      const int o = (e / NB) * DM_S + (e % NB);
      TMs[o] = TM[e];
      Ts[o] = Tf[e];
      if (corr) { G1s[o] = G[e]; G2s[o] = G[NB * NB + e]; }
    }
  }
  __syncthreads();
  if (corr) {
#pragma unroll
    for (int s_ = 0; s_ < NB / 4; ++s_)
#pragma unroll
      for (int y = 0; y < NB / 8; ++y) {
        dmma884(acc[y][0], acc[y][1], a1[s_], G1s[(4 * s_ + fk) * DM_S + 8 * y + fr]);   // Y -= V_a G1
        dmma884(acc[y][0], acc[y][1], a2[s_], G2s[(4 * s_ + fk) * DM_S + 8 * y + fr]);   // Y -= W_a G2
      }
  }
#pragma unroll
  for (int s_ = 0; s_ < NB / 4; ++s_)
#pragma unroll
    for (int y = 0; y < NB / 8; ++y)
      dmma884(acc[y][0], acc[y][1], -0.5 * av[s_], TMs[(4 * s_ + fk) * DM_S + 8 * y + fr]);   // X1 = Y - 1/2 V (T^T M)
  // W = X1 T: the accumulator tile (columns 2 fk + {0, 1}) turned into A fragments (column fk) by two shuffles per k-step
  double wv[NB / 8][2];
#pragma unroll
  for (int y = 0; y < NB / 8; ++y) { wv[y][0] = 0.0; wv[y][1] = 0.0; }
#pragma unroll
  for (int j = 0; j < NB / 4; ++j) {
    const int src = (lane & ~3) | (2 * (j & 1) + (fk >> 1));
    const double v0 = __shfl_sync(0xffffffffu, acc[j >> 1][0], src);
    const double v1 = __shfl_sync(0xffffffffu, acc[j >> 1][1], src);
    const double xa = (fk & 1) ? v1 : v0;
#pragma unroll
    for (int y = 0; y < NB / 8; ++y) dmma884(wv[y][0], wv[y][1], xa, Ts[(4 * j + fk) * DM_S + 8 * y + fr]);
  }
  if (ok) {
    double* W = a.W + (size_t)t * n * LDV + (size_t)i * LDV;
#pragma unroll
    for (int y = 0; y < NB / 8; ++y) *reinterpret_cast<double2*>(W + 8 * y + 2 * fk) = make_double2(wv[y][0], wv[y][1]);
  }
}

// Zpart = V_I^T B_I of a 128-row block (32 x nz, K = 128 rows) in two halves of 64 rows: the rows of B and V are
// staged in shared memory by 16-byte cp.async copies (whole rows, coalesced; B stride 108 == 12 mod 16 keeps the DMMA
// fragment loads conflict-free), three CTAs per SM so that one CTA's copies overlap the others' DMMA chains.  Warp
// (x, yh): rows 8 x .. 8 x + 7 of the product, column tiles 7 yh .. (13 tiles cover nz <= 104).
constexpr int ZP_HB = 64;
constexpr int ZP_SB = 108;
constexpr size_t ZP_SMEM = (size_t)ZP_HB * (DM_S + ZP_SB) * sizeof(double);

__global__ void __launch_bounds__(256, 3) zpart_dmma_kernel(SymmArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* Vs = smem;                       // ZP_HB x DM_S
  double* Bs = Vs + ZP_HB * DM_S;          // ZP_HB x ZP_SB
  const int t = blockIdx.y, bx = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = a.n, i0 = a.r0 + SY_BM * bx;
  const double* V = a.V + (size_t)t * n * LDV;
  const double* Bq = a.Bq + (size_t)t * n * a.ldb;
  const int x = wid & 3, yh = wid >> 2;
  constexpr int ZQ = 7;
  const int nq = yh ? 13 - ZQ : ZQ;        // column tiles of this warp
  const int chunks = a.nz >> 1;            // 16-byte pieces of a row of B
  double acc[ZQ][2];
#pragma unroll
  for (int q = 0; q < ZQ; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
#pragma unroll 1
  for (int h = 0; h < SY_BM / ZP_HB; ++h) {
    const int rh = i0 + ZP_HB * h;
    if (h) __syncthreads();                // the previous half has been consumed
    for (int e = tid; e < ZP_HB * (NB / 2); e += 256) {
      const int rl = e / (NB / 2), part = e % (NB / 2);
      const bool ok = rh + rl < n;
      cp_async16(Vs + rl * DM_S + 2 * part, ok ? V + (size_t)(rh + rl) * LDV + 2 * part : V, ok);
    }
    for (int e = tid; e < ZP_HB * chunks; e += 256) {
      const int rl = e / chunks, part = e - rl * chunks;
      const bool ok = rh + rl < n;
      cp_async16(Bs + rl * ZP_SB + 2 * part, ok ? Bq + (size_t)(rh + rl) * a.ldb + 2 * part : Bq, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll 4
    for (int s_ = 0; s_ < ZP_HB / 4; ++s_) {
      const double av = Vs[(4 * s_ + fk) * DM_S + 8 * x + fr];
#pragma unroll
      for (int q = 0; q < ZQ; ++q)
        if (q < nq) dmma884(acc[q][0], acc[q][1], av, Bs[(4 * s_ + fk) * ZP_SB + 8 * (ZQ * yh + q) + fr]);
    }
  }
  double* Zp = a.Zpart + ((size_t)t * a.nblk + bx) * NB * SY_ZW + (size_t)(8 * x + fr) * SY_ZW;
#pragma unroll
  for (int q = 0; q < ZQ; ++q) {
    const int c = 8 * (ZQ * yh + q) + 2 * fk;
    if (q < nq && c < a.nz) Zp[c] = acc[q][0];
    if (q < nq && c + 1 < a.nz) Zp[c + 1] = acc[q][1];
  }
}

// Partial products of a 128-row block (the DMMA form of symm_partials_kernel, same outputs): sum of the split-K
// partials of Y, Mpart = V_I^T Y_I, Gpart = {W_a,I^T V_I, V_a,I^T V_I}
__global__ void __launch_bounds__(256) partials_dmma_kernel(SymmArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* Ys = smem;                       // SY_BM x DM_S
  double* Vs = Ys + SY_BM * DM_S;          // SY_BM x DM_S
  const int t = blockIdx.y, bx = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = a.n, i0 = a.r0 + SY_BM * bx;
  const double* V = a.V + (size_t)t * n * LDV;
  const int x = wid & 3, yp = wid >> 2;    // output rows 8 x .. (a column block of V^T), half of the output columns
  double* Y0 = a.Y + (size_t)t * n * NB;
  for (int e = tid; e < SY_BM * NB; e += 256) {
    const int rl = e / NB, c = e % NB;
    const bool ok = i0 + rl < n;
    double yv = 0.0;
    if (ok) {
      for (int sp = 0; sp < a.nsplit; ++sp) yv += a.Y[((size_t)sp * a.T + t) * n * NB + (size_t)(i0 + rl) * NB + c];
      if (a.nsplit > 1) Y0[(size_t)(i0 + rl) * NB + c] = yv;
    }
    Ys[rl * DM_S + c] = yv;
    Vs[rl * DM_S + c] = ok ? V[(size_t)(i0 + rl) * LDV + c] : 0.0;
  }
  __syncthreads();
  // out[r][c] = sum_rows L[row][r] R[row][c] for r = 8 x + fr, c = 16 yp + 8 j + 2 fk + {0, 1}
  auto tn_product = [&](const double* L, const double* R, double* out) {
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 8
    for (int s_ = 0; s_ < SY_BM / 4; ++s_) {
      const double av = L[(4 * s_ + fk) * DM_S + 8 * x + fr];
      dmma884(acc[0][0], acc[0][1], av, R[(4 * s_ + fk) * DM_S + 16 * yp + fr]);
      dmma884(acc[1][0], acc[1][1], av, R[(4 * s_ + fk) * DM_S + 16 * yp + 8 + fr]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
      *reinterpret_cast<double2*>(out + (8 * x + fr) * NB + 16 * yp + 8 * j + 2 * fk) = make_double2(acc[j][0], acc[j][1]);
  };
  tn_product(Vs, Ys, a.Mpart + ((size_t)t * a.nblk + bx) * NB * NB);            // V^T Y
  if (a.Gpart) {
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const double* src = (which == 0 ? a.Wa : a.Va) + (size_t)t * n * LDV;
      __syncthreads();                       // Ys (Y rows, then W_a rows) has been consumed
      for (int e = tid; e < SY_BM * NB; e += 256) {
        const int rl = e / NB, c = e % NB;
        Ys[rl * DM_S + c] = (i0 + rl < n) ? src[(size_t)(i0 + rl) * LDV + c] : 0.0;
      }
      __syncthreads();
      tn_product(Ys, Vs, a.Gpart + (((size_t)t * a.nblk + bx) * 2 + which) * NB * NB);   // W_a^T V, V_a^T V
    }
  }
}

// ------------------------------------------------------------------ thin update of the next panel's block column
// Second panel of a deferred-update group (columns c0 .. c0+NB-1, rows >= c0): bring the block column up to date with
// the first panel's reflector before its QR,  A[i][j] -= V_a[i] . W_a[j] + W_a[i] . V_a[j]   (lower part, i >= j).
struct PupdArgs {
  double* A; size_t strideA; int ldk;
  const double* Va; const double* Wa;     // T x n x LDV, column block 0
  int n, c0;
};
constexpr int PU_BM = 64;
constexpr size_t PU_SMEM = (size_t)(2 * PU_BM * WF_RS + 2 * NB * WF_RS) * sizeof(double);

__global__ void __launch_bounds__(256) panel_update_kernel(PupdArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* VI = smem;                        // PU_BM x WF_RS
  double* WI = VI + PU_BM * WF_RS;
  double* VJ = WI + PU_BM * WF_RS;          // NB x WF_RS  (rows c0 .. c0+NB-1 of V_a / W_a)
  double* WJ = VJ + NB * WF_RS;
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = a.n, c0 = a.c0, i0 = c0 + PU_BM * blockIdx.x;
  const double* Va = a.Va + (size_t)t * n * LDV;
  const double* Wa = a.Wa + (size_t)t * n * LDV;
  double* A = a.A + (size_t)t * a.strideA;
  for (int e = tid; e < PU_BM * NB; e += 256) {
    const int rl = e / NB, c = e % NB;
    const bool ok = i0 + rl < n;
    VI[rl * WF_RS + c] = ok ? Va[(size_t)(i0 + rl) * LDV + c] : 0.0;
    WI[rl * WF_RS + c] = ok ? Wa[(size_t)(i0 + rl) * LDV + c] : 0.0;
  }
  for (int e = tid; e < NB * NB; e += 256) {
    const int rl = e / NB, c = e % NB;
    const bool ok = c0 + rl < n;
    VJ[rl * WF_RS + c] = ok ? Va[(size_t)(c0 + rl) * LDV + c] : 0.0;
    WJ[rl * WF_RS + c] = ok ? Wa[(size_t)(c0 + rl) * LDV + c] : 0.0;
  }
  __syncthreads();
  // thread: column j = c0 + lane, rows i0 + wid + 8 q
  double acc[PU_BM / 8];
#pragma unroll
  for (int q = 0; q < PU_BM / 8; ++q) acc[q] = 0.0;
#pragma unroll 4
  for (int k = 0; k < NB; ++k) {
    const double wj = WJ[lane * WF_RS + k], vj = VJ[lane * WF_RS + k];
#pragma unroll
    for (int q = 0; q < PU_BM / 8; ++q)
      acc[q] = fma(VI[(wid + 8 * q) * WF_RS + k], wj, fma(WI[(wid + 8 * q) * WF_RS + k], vj, acc[q]));
  }
  const int j = c0 + lane;
#pragma unroll
  for (int q = 0; q < PU_BM / 8; ++q) {
    const int i = i0 + wid + 8 * q;
    if (i < n && j < n && j <= i) A[(size_t)i * a.ldk + j] -= acc[q];
  }
}

// DMMA form of the thin update: warp w of a 64-row block owns rows 8 w .. 8 w + 7 (A fragments of V_a / W_a straight
// from global memory, accumulators = the 8 x 32 piece of the block column), the 32 rows c0 .. of W_a / V_a are the B
// operands in shared memory.
constexpr size_t PUD_SMEM = (size_t)(2 * NB * DM_S) * sizeof(double);

__global__ void __launch_bounds__(256) pupd_dmma_kernel(PupdArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* VJ = smem;                        // NB x DM_S  (rows c0 .. c0+NB-1 of V_a / W_a)
  double* WJ = VJ + NB * DM_S;
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int n = a.n, c0 = a.c0, i = c0 + PU_BM * blockIdx.x + 8 * wid + fr;
  const bool ok = i < n;
  const double* Va = a.Va + (size_t)t * n * LDV;
  const double* Wa = a.Wa + (size_t)t * n * LDV;
  double* Arow = a.A + (size_t)t * a.strideA + (size_t)i * a.ldk + c0;
  for (int e = tid; e < NB * NB; e += 256) {
    const int rl = e / NB, c = e % NB;
    const bool okj = c0 + rl < n;
    VJ[rl * DM_S + c] = okj ? Va[(size_t)(c0 + rl) * LDV + c] : 0.0;
    WJ[rl * DM_S + c] = okj ? Wa[(size_t)(c0 + rl) * LDV + c] : 0.0;
  }
  double va[NB / 4], wa[NB / 4], acc[NB / 8][2];
#pragma unroll
  for (int s_ = 0; s_ < NB / 4; ++s_) {
    va[s_] = ok ? -Va[(size_t)i * LDV + 4 * s_ + fk] : 0.0;
    wa[s_] = ok ? -Wa[(size_t)i * LDV + 4 * s_ + fk] : 0.0;
  }
#pragma unroll
  for (int y = 0; y < NB / 8; ++y) {
    double2 av = make_double2(0.0, 0.0);
    if (ok) av = *reinterpret_cast<const double2*>(Arow + 8 * y + 2 * fk);
    acc[y][0] = av.x; acc[y][1] = av.y;
  }
  __syncthreads();
#pragma unroll
  for (int s_ = 0; s_ < NB / 4; ++s_)
#pragma unroll
    for (int y = 0; y < NB / 8; ++y) {
      dmma884(acc[y][0], acc[y][1], va[s_], WJ[(8 * y + fr) * DM_S + 4 * s_ + fk]);   // -= V_a[i] . W_a[j]
      dmma884(acc[y][0], acc[y][1], wa[s_], VJ[(8 * y + fr) * DM_S + 4 * s_ + fk]);   // -= W_a[i] . V_a[j]
    }
  if (ok) {
#pragma unroll
    for (int y = 0; y < NB / 8; ++y) {
      const int j = c0 + 8 * y + 2 * fk;                     // lower part only: j <= i
      if (j + 1 <= i) *reinterpret_cast<double2*>(Arow + 8 * y + 2 * fk) = make_double2(acc[y][0], acc[y][1]);
      else if (j <= i) Arow[8 * y + 2 * fk] = acc[y][0];
    }
  }
}

// ------------------------------------------------------------------ A22 -= sum_q V_q W_q^T + W_q V_q^T  (lower tiles)
// Row-persistent DMMA kernel: one CTA keeps Z_I = [V_I | W_I] (all G panels of the group: K = 2 G NB columns) of its
// 128-row block in shared memory and walks a chunk of BN-column tiles of that block row; the next tile's
// Z'_J = [W_J | V_J] (cp.async, double buffered) and its A values (registers) are fetched while the current tile is on
// the tensor cores; accumulators start at -A, the result is stored as -acc.
constexpr int S3_BM = 128;   // tile rows

struct Syr2kArgs {
  double* A; size_t strideA; int ldk;
  const double* V; const double* W;   // T x n x LDV, column blocks 0 .. G-1
  int n, r0;
};

// BN = tile columns, CH = tiles per CTA, G = panels in the update (K = 2 G NB)
template <int BN, int CH, int G>
__global__ void __launch_bounds__(256, 1) syr2k_row_kernel(Syr2kArgs a) {
  constexpr int NY = BN / 16;          // 8-column fragments per warp
  constexpr int KT = 2 * G * NB;       // K of the update
  constexpr int SS = KT + 4;           // shared-memory row stride (== 4 mod 8: conflict-free fragment loads)
  constexpr int CPR = KT / 2;          // 16-byte chunks per operand row
  extern __shared__ __align__(16) double smem[];
  const int t = blockIdx.z;
  const int n = a.n, r0 = a.r0, ldk = a.ldk;
  const int i0 = r0 + S3_BM * blockIdx.y;
  // BN-column tiles of this block row that touch the lower triangle: j0 = r0 + BN jt <= i0 + 127, j0 < n
  const int njt = min((i0 + S3_BM - r0) / BN, (n - r0 + BN - 1) / BN);
  const int jt_lo = CH * blockIdx.x, jt_hi = min(njt, jt_lo + CH);
  if (jt_lo >= jt_hi) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  double* A = a.A + (size_t)t * a.strideA;
  const double* V = a.V + (size_t)t * n * LDV;
  const double* W = a.W + (size_t)t * n * LDV;
  double* As = smem;                      // S3_BM x SS : [V_I | W_I]
  double* Bs = As + S3_BM * SS;           // 2 x BN x SS : [W_J | V_J]
  const int wm = wid >> 1, wn = wid & 1;  // 4 x 2 warps, warp tile 32 x BN/2
  // operand row: chunks 0 .. CPR/2-1 from the first array, the rest from the second (G NB columns each)
  auto stage_rows = [&](double* dst, int row0, int rows, const double* first, const double* second) {
    for (int c = tid; c < rows * CPR; c += 256) {
      const int row = c / CPR, part = c % CPR;
      const bool ok = row0 + row < n;
      const double* base = (part < CPR / 2) ? first : second;
      const int col = 2 * (part % (CPR / 2));
      cp_async16(dst + row * SS + 2 * part, ok ? base + (size_t)(row0 + row) * LDV + col : base, ok);
    }
  };
  stage_rows(As, i0, S3_BM, V, W);
  auto issue_b = [&](int jt) {
    stage_rows(Bs + (size_t)((jt - jt_lo) & 1) * BN * SS, r0 + BN * jt, BN, W, V);
  };
  double nxt[4][NY][2];
  const size_t rowstep = (size_t)8 * ldk;
  auto load_a = [&](int jt) {
    const int j0 = r0 + BN * jt;
    const int ib = i0 + 32 * wm + fr, jb = j0 + (BN / 2) * wn + 2 * fk;
    const double* p0 = A + (size_t)ib * ldk + jb;
    if (j0 + BN <= i0 && i0 + S3_BM <= n) {   // tile strictly below the diagonal, all rows present: no masks
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < NY; ++y) {
          const double2 o = ldcg2(reinterpret_cast<const double2*>(p0 + x * rowstep + 8 * y));
          nxt[x][y][0] = o.x;
          nxt[x][y][1] = o.y;
        }
    } else {
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int i = ib + 8 * x;
#pragma unroll
        for (int y = 0; y < NY; ++y) {
          const int j = jb + 8 * y;
          const double* p = p0 + x * rowstep + 8 * y;
          double2 o = make_double2(0.0, 0.0);
          if (i < n) {
            if (j + 1 <= i) o = ldcg2(reinterpret_cast<const double2*>(p));
            else if (j <= i) o.x = ldcg(p);
          }
          nxt[x][y][0] = o.x;
          nxt[x][y][1] = o.y;
        }
      }
    }
  };
  issue_b(jt_lo);
  cp_async_commit();
  load_a(jt_lo);
  const double* ap = As + (32 * wm + fr) * SS + fk;
  for (int jt = jt_lo; jt < jt_hi; ++jt) {
    cp_async_wait<0>();
    __syncthreads();
    if (jt + 1 < jt_hi) issue_b(jt + 1);
    cp_async_commit();
    double acc[4][NY][2];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < NY; ++y) { acc[x][y][0] = -nxt[x][y][0]; acc[x][y][1] = -nxt[x][y][1]; }
    if (jt + 1 < jt_hi) load_a(jt + 1);
    const double* bp = Bs + (size_t)((jt - jt_lo) & 1) * BN * SS + ((BN / 2) * wn + fr) * SS + fk;
#pragma unroll 4
    for (int kk = 0; kk < KT; kk += 4) {
      double af[4], bf[NY];
#pragma unroll
      for (int x = 0; x < 4; ++x) af[x] = ap[8 * x * SS + kk];
#pragma unroll
      for (int y = 0; y < NY; ++y) bf[y] = bp[8 * y * SS + kk];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < NY; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
    }
    const int j0 = r0 + BN * jt;
    const int ib = i0 + 32 * wm + fr, jb = j0 + (BN / 2) * wn + 2 * fk;
    double* p0 = A + (size_t)ib * ldk + jb;
    if (j0 + BN <= i0 && i0 + S3_BM <= n) {
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < NY; ++y)
          *reinterpret_cast<double2*>(p0 + x * rowstep + 8 * y) = make_double2(-acc[x][y][0], -acc[x][y][1]);
    } else {
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int i = ib + 8 * x;
        if (i < n) {
#pragma unroll
          for (int y = 0; y < NY; ++y) {
            const int j = jb + 8 * y;
            double* p = p0 + x * rowstep + 8 * y;
            if (j + 1 <= i) *reinterpret_cast<double2*>(p) = make_double2(-acc[x][y][0], -acc[x][y][1]);
            else if (j <= i) p[0] = -acc[x][y][0];
          }
        }
      }
    }
  }
}
template <int BN, int G>
constexpr size_t syr2k_smem() { return (size_t)(S3_BM + 2 * BN) * (2 * G * NB + 4) * sizeof(double); }

// ------------------------------------------------------------------ band extraction / padding
// bandS[i][d] = T_b(i, i-d), d = 0..NB  (stride CHD_BAND_LD);  bandW: the same with room for the bulges
// (2 NB sub-diagonals, stride CHD_CHASE_LD), zero filled.
__global__ void band_extract_kernel(const double* __restrict__ A, size_t strideA, int ldk, int n,
                                    double* __restrict__ bandS, double* __restrict__ bandW) {
  const int t = blockIdx.y;
  const double* At = A + (size_t)t * strideA;
  double* bs = bandS + (size_t)t * n * CHD_BAND_LD;
  double* bw = bandW ? bandW + (size_t)t * n * CHD_CHASE_LD : nullptr;
  const size_t total = (size_t)n * CHD_CHASE_LD;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / CHD_CHASE_LD), d = (int)(e % CHD_CHASE_LD);
    double v = 0.0;
    if (d <= NB && i - d >= 0) v = At[(size_t)i * ldk + (i - d)];
    if (bw) bw[e] = v;
    if (d < CHD_BAND_LD) bs[(size_t)i * CHD_BAND_LD + d] = v;
  }
}

__global__ void zero_pad_kernel(double* __restrict__ A, size_t strideA, int ldk, int n) {
  double* At = A + (size_t)blockIdx.y * strideA;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int j = n; j < ldk; ++j) At[(size_t)i * ldk + j] = 0.0;
}

// ------------------------------------------------------------------ back-transformation  out = -beta0 Q x
// Q = H_{-1} Q_0 Q_1 ...: block reflectors applied last to first, x <- x - V (T (V^T x)).  One cluster of
// BT_CS CTAs per matrix; every CTA keeps its row slice of x in shared memory, V is read in place (below the
// band of the factored matrix), the NB-vector V^T x is exchanged through DSMEM.
constexpr int BT_CS = 8;
constexpr int BT_THREADS = 512;

struct Back2Args {
  const double* A; size_t strideA; int ldk;
  const double* Tf; size_t tf_stride;   // slot p = panel p (slot 0 = H_{-1})
  const double* v0;     // T x n
  const double* tau0;   // T
  const double* x;      // T x n
  const double* beta0;  // T
  double* out; size_t out_stride;
  int n, npanel;        // npanel = number of QR panels (slots 1..npanel)
  int RPB;              // rows of x per CTA
};

__global__ void __cluster_dims__(BT_CS, 1, 1) __launch_bounds__(BT_THREADS) backtransform2_kernel(Back2Args a) {
  extern __shared__ __align__(16) double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int g = (int)cluster.block_rank();
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = BT_THREADS / 32;
  const int n = a.n, ldk = a.ldk, RPB = a.RPB;
  const double* A = a.A + (size_t)t * a.strideA;
  double* xch = smem;                       // 2 x BT_CS x NB
  double* red = xch + 2 * BT_CS * NB;       // NW x (NB + 1)
  double* cs = red + NW * (NB + 1);         // NB   V^T x
  double* ss = cs + NB;                     // NB   T (V^T x)
  double* Ts = ss + NB;                     // NB x (NB + 1)
  double* xs = Ts + NB * (NB + 1);          // RPB
  const int row_lo = g * RPB, row_hi = min(n, row_lo + RPB);
  for (int i = row_lo + tid; i < row_hi; i += BT_THREADS) xs[i - row_lo] = a.x[(size_t)t * n + i];
  __syncthreads();
  int par = 0;
  for (int p = a.npanel; p >= 1; --p) {
    const int r0 = p * NB, c0 = r0 - NB;
    const int m = n - r0, nref = min(NB, m - 1);
    const double* Tf = a.Tf + (size_t)t * a.tf_stride + (size_t)p * NB * NB;
    for (int e = tid; e < NB * NB; e += BT_THREADS) Ts[(e / NB) * (NB + 1) + (e % NB)] = Tf[e];
    const int lo = max(row_lo, r0);
    // ---- c = V^T x over the owned rows: lanes = reflector columns, 4 rows in flight per warp iteration
    // (four rows' loads in flight per warp; ONE accumulator, rows in increasing order: the sum is bit-identical to the
    // rolled loop, which spent its time in one exposed HBM latency per row)
    auto vmask = [&](double v, int pr) {
      if (pr < NB) return (lane >= nref || pr < lane) ? 0.0 : ((pr == lane) ? 1.0 : v);
      return (lane >= nref) ? 0.0 : v;
    };
    double acc = 0.0;
    int i = lo + wid;
    for (; i + 3 * NW < row_hi; i += 4 * NW) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = A[(size_t)(i + u * NW) * ldk + c0 + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fma(vmask(v[u], i + u * NW - r0), xs[i + u * NW - row_lo], acc);
    }
    for (; i < row_hi; i += NW) acc = fma(vmask(A[(size_t)i * ldk + c0 + lane], i - r0), xs[i - row_lo], acc);
    red[wid * (NB + 1) + lane] = acc;
    __syncthreads();
    if (tid < NB) {
      double s = 0.0;
      for (int w = 0; w < NW; ++w) s += red[w * (NB + 1) + tid];
      for (int dst = 0; dst < BT_CS; ++dst) {
        double* remote = cluster.map_shared_rank(xch + (par * BT_CS + g) * NB, dst);
        remote[tid] = s;
      }
    }
    cluster.sync();
    if (tid < NB) {
      double s = 0.0;
      for (int q = 0; q < BT_CS; ++q) s += xch[(par * BT_CS + q) * NB + tid];
      cs[tid] = s;
    }
    par ^= 1;
    __syncthreads();
    if (tid < NB) {
      double s = 0.0;
      for (int k = tid; k < NB; ++k) s = fma(Ts[tid * (NB + 1) + k], cs[k], s);
      ss[tid] = s;
    }
    __syncthreads();
    // ---- x -= V s
    i = lo + wid;
    for (; i + 3 * NW < row_hi; i += 4 * NW) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = A[(size_t)(i + u * NW) * ldk + c0 + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double d = warp_sum(vmask(v[u], i + u * NW - r0) * ss[lane]);
        if (lane == 0) xs[i + u * NW - row_lo] -= d;
      }
    }
    for (; i < row_hi; i += NW) {
      const double d = warp_sum(vmask(A[(size_t)i * ldk + c0 + lane], i - r0) * ss[lane]);
      if (lane == 0) xs[i - row_lo] -= d;
    }
    __syncthreads();
  }
  // ---- H_{-1}
  {
    const double* v0 = a.v0 + (size_t)t * n;
    double acc = 0.0;
    for (int i = row_lo + tid; i < row_hi; i += BT_THREADS) acc = fma(v0[i], xs[i - row_lo], acc);
    acc = warp_sum(acc);
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < NW; ++w) s += red[w];
      for (int dst = 0; dst < BT_CS; ++dst) {
        double* remote = cluster.map_shared_rank(xch + (par * BT_CS + g) * NB, dst);
        remote[0] = s;
      }
    }
    cluster.sync();
    double s = 0.0;
    for (int q = 0; q < BT_CS; ++q) s += xch[(par * BT_CS + q) * NB];
    const double f = a.tau0[t] * s, sc = -a.beta0[t];
    for (int i = row_lo + tid; i < row_hi; i += BT_THREADS)
      a.out[(size_t)t * a.out_stride + i] = sc * (xs[i - row_lo] - f * v0[i]);
  }
  cluster.sync();
}

// ------------------------------------------------------------------ host side
int band_npanels(int n) {
  // panels p = 1.. with r0 = p NB and at least two rows below the band
  int np = 0;
  while (n - (np + 1) * NB >= 2) ++np;
  return np;
}

size_t band_workspace(const Plan* pl, int T) {
  const size_t n = pl->n;
  const size_t nblk = (n + SY_BM - 1) / SY_BM;
  size_t b = 0;
  b += 2 * align_up(T * n * LDV * sizeof(double), 256);                      // V, W (KD panels side by side)
  b += SY_MAXSPLIT * align_up(T * n * NB * sizeof(double), 256);             // Y (split-K partials)
  b += align_up((size_t)T * (band_npanels(pl->n) + 1) * NB * NB * sizeof(double), 256);   // T factors
  b += align_up((size_t)T * nblk * SY_MAXSPLIT * NB * NB * sizeof(double), 256);   // Mpart
  b += align_up((size_t)T * nblk * NB * SY_ZW * sizeof(double), 256);        // Zpart
  b += align_up((size_t)T * nblk * 2 * NB * NB * sizeof(double), 256);       // Gpart
  b += 2 * align_up((size_t)T * NB * NB * sizeof(double), 256);              // Mred, TM
  b += 2 * align_up((size_t)T * NB * SY_ZW * sizeof(double), 256);           // Zred, TZ
  b += align_up((size_t)T * 2 * NB * NB * sizeof(double), 256);              // Gred
  b += align_up(T * n * sizeof(double), 256);                                // v0
  b += align_up(T * sizeof(double), 256);                                    // tau0
  b += align_up((size_t)T * n * CHD_BAND_LD * sizeof(double), 256);          // bandS
  b += align_up((size_t)T * n * CHD_CHASE_LD * sizeof(double), 256);         // bandW
  return b + 8192;
}

bool band_carve(const Plan* pl, int T, Carver& cv, BandBuffers& bb) {
  const size_t n = pl->n;
  const size_t nblk = (n + SY_BM - 1) / SY_BM;
  bb.V = cv.take<double>(T * n * LDV);
  bb.W = cv.take<double>(T * n * LDV);
  bb.Y = cv.take<double>((size_t)SY_MAXSPLIT * T * n * NB);
  bb.tf_stride = (size_t)(band_npanels(pl->n) + 1) * NB * NB;
  bb.Tf = cv.take<double>((size_t)T * bb.tf_stride);
  bb.Mpart = cv.take<double>((size_t)T * nblk * SY_MAXSPLIT * NB * NB);
  bb.Zpart = cv.take<double>((size_t)T * nblk * NB * SY_ZW);
  bb.Gpart = cv.take<double>((size_t)T * nblk * 2 * NB * NB);
  bb.Mred = cv.take<double>((size_t)T * NB * NB);
  bb.TM = cv.take<double>((size_t)T * NB * NB);
  bb.Zred = cv.take<double>((size_t)T * NB * SY_ZW);
  bb.TZ = cv.take<double>((size_t)T * NB * SY_ZW);
  bb.Gred = cv.take<double>((size_t)T * 2 * NB * NB);
  bb.v0 = cv.take<double>(T * n);
  bb.tau0 = cv.take<double>(T);
  bb.bandS = cv.take<double>((size_t)T * n * CHD_BAND_LD);
  bb.bandW = cv.take<double>((size_t)T * n * CHD_CHASE_LD);
  return cv.ok;
}

static int qr_cluster_size(const Plan* pl, int m, int& RP, int& RS, size_t& smem) {
  // fixed part of the kernel's shared memory (doubles)
  const size_t fixed = 2 * QR_MAXCS * QR_XW + (QR_THREADS / 32) * QR_XW + 2 * QR_XW + 2 * NB + 2 * NB * QR_S;
  const size_t avail = (pl->smem_optin - fixed * sizeof(double)) / (QR_S * sizeof(double));
  int max_cs = 8;
  {
    const char* e = getenv("CHD_QR_MAXCS");   // tuning / test switch: largest cluster the panel QR may use
    const int v = e ? atoi(e) : 8;
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) max_cs = v;
  }
  int cs = 1;
  while (cs < max_cs && (m + cs - 1) / cs > 384) cs *= 2;
  RP = (m + cs - 1) / cs;
  RS = (int)((size_t)RP < avail ? (size_t)RP : avail);
  smem = (fixed + (size_t)RS * QR_S) * sizeof(double);
  return cs;
}

// One slot's work between its reflectors (H_{-1} or a panel QR) and the group's trailing update:
//   Y = A22 V (stale inside a group) -> partial sums -> small matrices -> W (with the stale-matrix corrections for the
//   second panel of a group) and the Z-block update.  q = position of the slot in its group.
static int slot_launch(const Plan* pl, double* K, int ldk, int T, double* B, int ldb, int nz, int r0, int slot, int q,
                       const BandBuffers& bb, cudaStream_t st) {
  const int n = pl->n, m = n - r0;
  const size_t strideA = (size_t)n * ldk;
  const int nblk = (m + SY_BM - 1) / SY_BM;
  const bool has_z = B && nz > 0, has_g = q > 0;
  const double* Vq = bb.V + q * NB;
  // DMMA forms of the small products: double2 accesses to the Z block need even ldb / nz and at most 13 column tiles
  const bool dmma_small = pl->small_dmma && (!has_z || ((ldb & 1) == 0 && (nz & 1) == 0 && nz <= 104));
  int rc;
  // split-K so that the launch fills the machine (2 CTAs per SM) at least ~twice, while every CTA keeps a few chunks
  int S = 1;
  {
    const int slots = 2 * pl->sm_count;
    const int niter_min = (m + 31) / 32;
    while (S < SY_MAXSPLIT && nblk * T * S < 2 * slots && niter_min / (2 * S) >= 4) S *= 2;
  }
  {
    SymmArgs a;
    a.A = K; a.strideA = strideA; a.ldk = ldk; a.V = Vq; a.Y = bb.Y; a.Bq = B; a.ldb = ldb; a.nz = nz;
    a.Mpart = bb.Mpart; a.Zpart = bb.Zpart; a.n = n; a.r0 = r0; a.nblk = nblk; a.nsplit = S; a.T = T;
    a.Va = has_g ? bb.V : nullptr; a.Wa = has_g ? bb.W : nullptr; a.Gpart = has_g ? bb.Gpart : nullptr;
    static const int kc = getenv("CHD_SYMM_KC") ? atoi(getenv("CHD_SYMM_KC")) : 32;
    if ((rc = profile_mark_begin(PROF_SYMM, st))) return rc;
    if (pl->symm_tma) {
      if ((rc = symm_tma_launch(pl, a, T, S, st))) return rc;
    } else if (kc == 64)
      symm_kernel<64><<<dim3(nblk, T, S), 256, SY_NST * SymmCfg<64>::STAGE * sizeof(double), st>>>(a);
    else
      symm_kernel<32><<<dim3(nblk, T, S), 256, SY_NST * SymmCfg<32>::STAGE * sizeof(double), st>>>(a);
    if (!pl->symm_tma) CHD_LAUNCHED();
    if ((rc = profile_mark_end(PROF_SYMM, st))) return rc;
    if ((rc = profile_mark_begin(PROF_WFORM, st))) return rc;
    if (dmma_small) {
      if (has_z) {
        zpart_dmma_kernel<<<dim3(nblk, T), 256, ZP_SMEM, st>>>(a);
        CHD_LAUNCHED();
      }
      partials_dmma_kernel<<<dim3(nblk, T), 256, PD_SMEM, st>>>(a);
    } else {
      symm_partials_kernel<<<dim3(nblk, T, has_z ? 2 : 1), 256, SP_SMEM, st>>>(a);
    }
    CHD_LAUNCHED();
  }
  {
    ReduceArgs a;
    a.Mpart = bb.Mpart; a.Zpart = bb.Zpart; a.Gpart = bb.Gpart; a.Mred = bb.Mred; a.Zred = bb.Zred; a.Gred = bb.Gred;
    a.nblk = nblk; a.has_z = has_z ? 1 : 0; a.has_g = has_g ? 1 : 0;
    const int total = NB * NB + NB * SY_ZW + (has_g ? 2 * NB * NB : 0);
    reduce_kernel<<<dim3((total + 255) / 256, T), 256, 0, st>>>(a);
    CHD_LAUNCHED();
  }
  {
    SmallArgs a;
    a.Tf = bb.Tf + (size_t)slot * NB * NB; a.tf_stride = bb.tf_stride; a.Mred = bb.Mred; a.Zred = bb.Zred;
    a.Gred = bb.Gred; a.TM = bb.TM; a.TZ = bb.TZ; a.has_z = has_z ? 1 : 0; a.has_g = has_g ? 1 : 0;
    smallmat_kernel<<<T, 256, 0, st>>>(a);
    CHD_LAUNCHED();
  }
  {
    WformArgs a;
    a.V = Vq; a.Y = bb.Y; a.W = bb.W + q * NB; a.nsplit = 1; a.T = T;   // split partials already summed
    a.Tf = bb.Tf + (size_t)slot * NB * NB; a.tf_stride = bb.tf_stride; a.TM = bb.TM;
    a.Gred = has_g ? bb.Gred : nullptr; a.Va = bb.V; a.Wa = bb.W; a.n = n; a.r0 = r0;
    ZupdArgs z;
    z.V = Vq; z.TZ = bb.TZ; z.Bq = B; z.ldb = ldb; z.nz = nz; z.n = n; z.r0 = r0;
    if (dmma_small) {
      wform_dmma_kernel<<<dim3((m + WF_BM - 1) / WF_BM, T), 256, WZ_SMEM, st>>>(a);
      CHD_LAUNCHED();
      if (has_z) {
        zupd_dmma_kernel<<<dim3((m + WF_BM - 1) / WF_BM, T, (nz + 8 * ZU_T - 1) / (8 * ZU_T)), 256, 0, st>>>(z);
        CHD_LAUNCHED();
      }
    } else {
      wform_kernel<<<dim3((m + WF_BM - 1) / WF_BM, T), 256, WF_SMEM, st>>>(a);
      CHD_LAUNCHED();
      if (has_z) {
        zupdate_kernel<<<dim3((m + WF_BM - 1) / WF_BM, T), 256, ZU_SMEM, st>>>(z);
        CHD_LAUNCHED();
      }
    }
  }
  if ((rc = profile_mark_end(PROF_WFORM, st))) return rc;
  return CHD_OK;
}

// Trailing update of a group of G panels: A[r0:, r0:] -= sum_q V_q W_q^T + W_q V_q^T  (lower tiles)
static int group_update_launch(const Plan* pl, double* K, int ldk, int T, int r0, int G, const BandBuffers& bb,
                               cudaStream_t st) {
  const int n = pl->n, m = n - r0;
  Syr2kArgs a;
  a.A = K; a.strideA = (size_t)n * ldk; a.ldk = ldk; a.V = bb.V; a.W = bb.W; a.n = n; a.r0 = r0;
  int rc;
  if ((rc = profile_mark_begin(PROF_SYR2K, st))) return rc;
  if (pl->syr2k_tma) {
    if ((rc = syr2k_tma_launch(pl, K, ldk, T, r0, G, bb.V, bb.W, st))) return rc;
    return profile_mark_end(PROF_SYR2K, st);
  }
  if (G == 1) {
    const int ntile = (m + 63) / 64;
    const dim3 grid((ntile + 7) / 8, (m + S3_BM - 1) / S3_BM, T);
    syr2k_row_kernel<64, 8, 1><<<grid, 256, syr2k_smem<64, 1>(), st>>>(a);
  } else {
    const int ntile = (m + 31) / 32;
    const dim3 grid((ntile + 15) / 16, (m + S3_BM - 1) / S3_BM, T);
    syr2k_row_kernel<32, 16, 2><<<grid, 256, syr2k_smem<32, 2>(), st>>>(a);
  }
  CHD_LAUNCHED();
  if ((rc = profile_mark_end(PROF_SYR2K, st))) return rc;
  return CHD_OK;
}

int band_device_init(Plan* pl) {
  CHD_CUDA(cudaFuncSetAttribute(symm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(SY_NST * SymmCfg<32>::STAGE * sizeof(double))));
  CHD_CUDA(cudaFuncSetAttribute(symm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(SY_NST * SymmCfg<64>::STAGE * sizeof(double))));
  CHD_CUDA(cudaFuncSetAttribute(wform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM));
  CHD_CUDA(cudaFuncSetAttribute(symm_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM));
  CHD_CUDA(cudaFuncSetAttribute(zupdate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZU_SMEM));
  CHD_CUDA(cudaFuncSetAttribute(panel_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PU_SMEM));
  CHD_CUDA(cudaFuncSetAttribute(wform_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WZ_SMEM));
  CHD_CUDA(cudaFuncSetAttribute(partials_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD_SMEM));
  CHD_CUDA(cudaFuncSetAttribute(zpart_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZP_SMEM));
  pl->small_dmma = getenv("CHD_SMALL_DMMA") ? atoi(getenv("CHD_SMALL_DMMA")) : 1;
  CHD_CUDA(cudaFuncSetAttribute(syr2k_row_kernel<64, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)syr2k_smem<64, 1>()));
  CHD_CUDA(cudaFuncSetAttribute(syr2k_row_kernel<32, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)syr2k_smem<32, 2>()));
  CHD_CUDA(cudaFuncSetAttribute(panel_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_optin));
  CHD_CUDA(cudaFuncSetAttribute(panel_qr_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CHD_CUDA(cudaFuncSetAttribute(panel_qr_reg_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CHD_CUDA(cudaFuncSetAttribute(panel_qr_reg_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CHD_CUDA(cudaFuncSetAttribute(panel_qr_reg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)RG_SMEM2));
  CHD_CUDA(cudaFuncSetAttribute(backtransform2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)pl->smem_optin));
  // panels per deferred-update group (CHD_KD=1: trailing update after every panel)
  pl->band_kd = KD;
  if (const char* e = getenv("CHD_KD")) {
    const int v = atoi(e);
    if (v >= 1 && v <= KD) pl->band_kd = v;
  }
  // largest cluster the register-resident panel QR may use: 16 (non-portable) when the device can co-schedule it
  const char* e = getenv("CHD_QR_REG");     // tuning / test switch: 0 disables the register-resident kernel
  if (e && atoi(e) == 0) {
    pl->qr_reg_max_cs = 0;
    pl->qr_two = 0;
  } else {
    pl->qr_reg_max_cs = 8;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(16, 1); cfg.blockDim = dim3(RG_THREADS); cfg.dynamicSmemBytes = 0;
    int nclusters = 0;
    cfg.dynamicSmemBytes = RG_SMEM2;          // the two-row variant is the more demanding one
    if (cudaOccupancyMaxActiveClusters(&nclusters, panel_qr_reg_kernel<true>, &cfg) == cudaSuccess && nclusters > 0)
      pl->qr_reg_max_cs = 16;
    else
      (void)cudaGetLastError();
    pl->qr_two = getenv("CHD_QR_TWO") ? atoi(getenv("CHD_QR_TWO")) : 1;
    if (const char* c = getenv("CHD_QR_REG_MAXCS")) {   // test switch: smaller clusters reach the two-row variant early
      const int v = atoi(c);
      if ((v == 1 || v == 2 || v == 4 || v == 8) && v < pl->qr_reg_max_cs) pl->qr_reg_max_cs = v;
    }
  }
  return CHD_OK;
}

int band_reduce_launch(const Plan* pl, double* K, int ldk, const double* ga, int T, double* B, int ldb, int nz,
                       double* beta0, const BandBuffers& bb, cudaStream_t st) {
  const int n = pl->n;
  if (nz > SY_ZW || (ldk & 1)) return CHD_EINVAL;
  const int reg_max_cs = pl->qr_reg_max_cs;
  const int kd = pl->band_kd;
  const size_t strideA = (size_t)n * ldk;
  if (ldk > n) {
    zero_pad_kernel<<<dim3((n + 255) / 256, T), 256, 0, st>>>(K, strideA, ldk, n);
    CHD_LAUNCHED();
  }
  const int np = band_npanels(n);
  int rc;
  // slots: 0 = H_{-1} (r0 = 0), p >= 1 = QR panel p (columns c0 = (p-1) NB .., rows r0 = p NB ..); groups of kd
  // consecutive slots share one trailing update
  for (int p = 0; p <= np; ++p) {
    const int q = p % kd;
    const int r0 = p * NB, c0 = r0 - NB, m = n - r0;
    if (p == 0) {
      Refl0Args a;
      a.ga = ga; a.V = bb.V; a.v0 = bb.v0; a.Tf = bb.Tf; a.tf_stride = bb.tf_stride; a.tau0 = bb.tau0;
      a.beta0 = beta0; a.n = n;
      refl0_kernel<<<T, 256, 0, st>>>(a);
      CHD_LAUNCHED();
    } else {
      if ((rc = profile_mark_begin(PROF_QR, st))) return rc;
      if (q > 0) {
        // bring this panel's block column up to date with the pending reflectors of its group
        PupdArgs u;
        u.A = K; u.strideA = strideA; u.ldk = ldk; u.Va = bb.V; u.Wa = bb.W; u.n = n; u.c0 = c0;
        if (pl->small_dmma) pupd_dmma_kernel<<<dim3((n - c0 + PU_BM - 1) / PU_BM, T), 256, PUD_SMEM, st>>>(u);
        else panel_update_kernel<<<dim3((n - c0 + PU_BM - 1) / PU_BM, T), 256, PU_SMEM, st>>>(u);
        CHD_LAUNCHED();
      }
      QrArgs a;
      a.A = K; a.strideA = strideA; a.ldk = ldk; a.V = bb.V + q * NB; a.Tf = bb.Tf + (size_t)p * NB * NB;
      a.tf_stride = bb.tf_stride; a.n = n; a.r0 = r0; a.c0 = c0;
      size_t smem;
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cfg.stream = st;
      int rcs = 1;
      while (rcs < reg_max_cs && (m + rcs - 1) / rcs > RG_THREADS) rcs *= 2;
      // One CTA per SM and a latency-bound column loop: when the one-row clusters of the T matrices need more than
      // one wave, half as many CTAs with two rows per thread (second row in shared memory, ~1.3x per column) win
      // (n = 8192, 64 matrices: 4.19 -> 3.06 ms per matrix).  CHD_QR_TWO: 0 only when needed, 1 auto, 2 always.
      if (rcs > 1 && (m + rcs - 1) / rcs <= RG_THREADS &&
          (pl->qr_two == 2 || (pl->qr_two == 1 && (long)T * rcs > pl->sm_count)))
        rcs /= 2;
      if (reg_max_cs > 0 && (m + rcs - 1) / rcs <= 2 * RG_THREADS) {
        // one row per thread in registers; a second one in shared memory when the cluster is too small for that
        const bool two = (m + rcs - 1) / rcs > RG_THREADS;
        a.RP = (m + rcs - 1) / rcs; a.RS = 0;
        cfg.gridDim = dim3(rcs, T);
        cfg.blockDim = dim3(RG_THREADS);
        cfg.dynamicSmemBytes = two ? RG_SMEM2 : 0;
        at[0].val.clusterDim.x = rcs;
        if (two) CHD_CUDA(cudaLaunchKernelEx(&cfg, panel_qr_reg_kernel<true>, a));
        else CHD_CUDA(cudaLaunchKernelEx(&cfg, panel_qr_reg_kernel<false>, a));
      } else {
        const int cs = qr_cluster_size(pl, m, a.RP, a.RS, smem);
        cfg.gridDim = dim3(cs, T);
        cfg.blockDim = dim3(QR_THREADS);
        cfg.dynamicSmemBytes = smem;
        at[0].val.clusterDim.x = cs;
        CHD_CUDA(cudaLaunchKernelEx(&cfg, panel_qr_kernel, a));
      }
      g_launches.fetch_add(1);
      if ((rc = profile_mark_end(PROF_QR, st))) return rc;
    }
    if ((rc = slot_launch(pl, K, ldk, T, B, ldb, nz, r0, p, q, bb, st))) return rc;
    if (q == kd - 1 || p == np)
      if ((rc = group_update_launch(pl, K, ldk, T, r0, q + 1, bb, st))) return rc;
  }
  band_extract_kernel<<<dim3(min((n * CHD_CHASE_LD + 255) / 256, 1024), T), 256, 0, st>>>(K, strideA, ldk, n,
                                                                                             bb.bandS, bb.bandW);
  CHD_LAUNCHED();
  return CHD_OK;
}

int backtransform2_launch(const Plan* pl, const double* K, int ldk, const BandBuffers& bb, const double* x,
                          const double* beta0, double* out, size_t out_stride, int T, cudaStream_t st) {
  Back2Args a;
  a.A = K; a.strideA = (size_t)pl->n * ldk; a.ldk = ldk; a.Tf = bb.Tf; a.tf_stride = bb.tf_stride; a.v0 = bb.v0;
  a.tau0 = bb.tau0; a.x = x; a.beta0 = beta0; a.out = out; a.out_stride = out_stride; a.n = pl->n;
  a.npanel = band_npanels(pl->n);
  a.RPB = (pl->n + BT_CS - 1) / BT_CS;
  const size_t smem = (2 * BT_CS * NB + (BT_THREADS / 32) * (NB + 1) + 2 * NB + NB * (NB + 1) + a.RPB) * sizeof(double);
  int rc;
  if ((rc = profile_mark_begin(PROF_BACK, st))) return rc;
  backtransform2_kernel<<<dim3(BT_CS, T), BT_THREADS, smem, st>>>(a);
  CHD_LAUNCHED();
  if ((rc = profile_mark_end(PROF_BACK, st))) return rc;
  return CHD_OK;
}

}  // namespace chd
