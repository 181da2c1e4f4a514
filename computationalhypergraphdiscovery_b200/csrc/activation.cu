// Per-cluster activations  act_c = yb^T Infl_c yb / energy  for every candidate ancestor cluster of
// every target in one pass, + argmin + row removal.  Replaces make_activation_function and the
// first two lines of `step` (helper_functions.py:40-134, 166-169) of the reference.
//
// All three kernels reduce to a few contractions, with
//   g_d   = x_d . yb,   Gm[d][d'] = sum_i x_id yb_i x_id'   (active variables only),
//   lin_c  = sum_{d in c} g_d^2
//   quad_c = 2 alpha lin_c + sum_{d in c} sum_{d' active} Gm[d][d']^2 * (d' in c ? 1 : 2)
//   gau_c  = sum_ij yb_i yb_j P_ij (1 - 1/prod_{d in c}(1 + E_d,ij)),  P = prod_{d active}(1 + E_d)
//   linear:    s lin_c / energy            (helper_functions.py:69-77 == kernels.py:104-106)
//   quadratic: s quad_c / energy           (helper_functions.py:92-109 == kernels.py:175-185)
//   gaussian:  (qs quad_c + s gau_c)/energy (kernels.py:270-278)
// which equal the reference's expressions term by term (the Hadamard-product quadratic forms
// yb^T (L_a o L_b) yb are sums of squares of Gm entries).
#include "internal.cuh"

namespace chd {

// w[t][d] = w0[t][d] && alive[t][cluster_of[d]]
__global__ void active_vars_kernel(const uint8_t* __restrict__ w0, const uint8_t* __restrict__ alive,
                                   const int32_t* __restrict__ cluster_of, int N, int C, uint8_t* __restrict__ w) {
  const int t = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < N) w[(size_t)t * N + d] = (w0[(size_t)t * N + d] && alive[(size_t)t * C + cluster_of[d]]) ? 1 : 0;
}

int active_vars_launch(const uint8_t* w0, const uint8_t* alive, const int32_t* cluster_of, int N, int C, int T,
                       uint8_t* w, cudaStream_t st) {
  active_vars_kernel<<<dim3((N + 127) / 128, T), 128, 0, st>>>(w0, alive, cluster_of, N, C, w);
  CHD_LAUNCHED();
  return CHD_OK;
}

__global__ void removed_vars_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int count,
                                    uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (a[i] && !b[i]) ? 1 : 0;
}

int removed_vars_launch(const uint8_t* w_prev, const uint8_t* w_new, int count, uint8_t* wrem, cudaStream_t st) {
  removed_vars_kernel<<<(count + 255) / 256, 256, 0, st>>>(w_prev, w_new, count, wrem);
  CHD_LAUNCHED();
  return CHD_OK;
}

// gdot[t][d] = sum_i X[i][d] yb[t][i];  energy[t] = -sum_i ga[t][i] yb[t][i]
__global__ void __launch_bounds__(256) gdot_kernel(const double* __restrict__ X, int n, int N,
                                                   const double* __restrict__ yb, const double* __restrict__ ga,
                                                   double* __restrict__ gdot, double* __restrict__ energy) {
  __shared__ double red[8][33];
  const int t = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  const double* y = yb + (size_t)t * n;
  double s = 0.0;
  if (d < N)
    for (int i = wid; i < n; i += 8) s += X[(size_t)i * N + d] * y[i];
  red[wid][lane] = s;
  __syncthreads();
  if (wid == 0 && d < N) {
    double r = 0.0;
    for (int q = 0; q < 8; ++q) r += red[q][lane];
    gdot[(size_t)t * N + d] = r;
  }
  if (blockIdx.x == 0) {
    __shared__ double sred[40];
    double e = 0.0;
    const double* g = ga + (size_t)t * n;
    for (int i = threadIdx.x; i < n; i += 256) e += g[i] * y[i];
    e = block_sum(e, sred);
    if (threadIdx.x == 0) energy[t] = -e;
  }
}

// Gm[t][d][d'] = sum_i X[i][d] yb_i X[i][d'], 32 x 32 tile per block (all variables; masked later).
__global__ void __launch_bounds__(256) gmat_kernel(const double* __restrict__ X, int n, int N,
                                                   const double* __restrict__ yb, double* __restrict__ Gm) {
  __shared__ double sa[32][33], sb[32][33];
  const int t = blockIdx.z;
  const int d0 = blockIdx.y * 32, e0 = blockIdx.x * 32;
  if (e0 > d0) return;  // symmetric: lower tiles only, mirrored on write
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7
  const double* y = yb + (size_t)t * n;
  double acc[4] = {0, 0, 0, 0};
  for (int i0 = 0; i0 < n; i0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty * 4 + r;
      double xa = 0.0, xb = 0.0;
      if (i < n) {
        const double yi = y[i];
        if (d0 + tx < N) xa = X[(size_t)i * N + d0 + tx] * yi;
        if (e0 + tx < N) xb = X[(size_t)i * N + e0 + tx];
      }
      sa[ty * 4 + r][tx] = xa;
      sb[ty * 4 + r][tx] = xb;
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const double b = sb[i][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += sa[i][ty * 4 + r] * b;
    }
    __syncthreads();
  }
  double* G = Gm + (size_t)t * N * N;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = d0 + ty * 4 + r, e = e0 + tx;
    if (d < N && e < N) {
      G[(size_t)d * N + e] = acc[r];
      // mirror only off-diagonal tiles: inside a diagonal tile (d, e) and (e, d) are both computed (with different
      // rounding), and a mirrored store would race with the other thread's own store
      if (d0 != e0) G[(size_t)e * N + d] = acc[r];
    }
  }
}

// rowq[t][d] = sum_{d' active} Gm[d][d']^2 * (same cluster ? 1 : 2)    (one warp per variable d)
__global__ void __launch_bounds__(256) quad_rows_kernel(const double* __restrict__ Gm, int N,
                                                        const uint8_t* __restrict__ w,
                                                        const int32_t* __restrict__ cluster_of,
                                                        double* __restrict__ rowq) {
  const int t = blockIdx.y, lane = threadIdx.x & 31;
  const int d = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (d >= N) return;
  const uint8_t* wt = w + (size_t)t * N;
  double s = 0.0;
  if (wt[d]) {
    const int cd = cluster_of[d];
    const double* row = Gm + ((size_t)t * N + d) * N;
    for (int e = lane; e < N; e += 32)
      if (wt[e]) {
        const double v = row[e];
        s += v * v * ((cluster_of[e] == cd) ? 1.0 : 2.0);
      }
  }
  s = warp_sum(s);
  if (lane == 0) rowq[(size_t)t * N + d] = s;
}

// Gaussian pair sums: partial[t][blk][c] = sum over the block's (i, j) tiles of
//   sym * yb_i yb_j P_ij (1 - 1/F_c,ij).  XT is N x n (variables major).  Members of cluster c are
// mem_idx[mem_off[c] .. mem_off[c+1]).
constexpr int GA_TILE = 64;
constexpr int GA_BLOCKS = 296;  // blocks of the pair-sum kernels = per-target stride of their partial sums
__global__ void __launch_bounds__(256) gauss_pairs_kernel(const double* __restrict__ XT, int n, int N, int C,
                                                          const double* __restrict__ yb,
                                                          const uint8_t* __restrict__ w,
                                                          const int32_t* __restrict__ mem_off,
                                                          const int32_t* __restrict__ mem_idx, double inv2l2,
                                                          double* __restrict__ partial, int nblk,
                                                          const double* __restrict__ Pc, int ldk,
                                                          const uint8_t* __restrict__ only_flagged) {
  extern __shared__ double sacc[];  // C x 8 warps
  const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (only_flagged && !only_flagged[t]) return;   // this target is served by gauss_shared_kernel
  const int tx = tid & 15, ty = tid >> 4;
  const uint8_t* wt = w + (size_t)t * N;
  const double* y = yb + (size_t)t * n;
  __shared__ double etab[64];
  exp_table_load(etab);
  for (int q = tid; q < C * 8; q += 256) sacc[q] = 0.0;
  __syncthreads();
  const int tb = (n + GA_TILE - 1) / GA_TILE;
  const int ntiles = tb * (tb + 1) / 2;
  for (int tile = blockIdx.x; tile < ntiles; tile += nblk) {
    int bi = (int)((sqrt(8.0 * tile + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= tile) ++bi;
    while (bi * (bi + 1) / 2 > tile) --bi;
    const int bj = tile - bi * (bi + 1) / 2;
    const int i0 = bi * GA_TILE + ty * 4, j0 = bj * GA_TILE + tx * 4;
    const double sym = (bi == bj) ? 1.0 : 2.0;
    double P[4][4], yy[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        P[a][b] = 1.0;
        const int i = i0 + a, j = j0 + b;
        yy[a][b] = (i < n && j < n) ? sym * y[i] * y[j] : 0.0;
      }
    if (Pc) {   // cached product map of the current mask: no first pass
      const double* Pm = Pc + (size_t)t * n * ldk;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (i0 + a < n && j0 + b < n) P[a][b] = Pm[(size_t)(i0 + a) * ldk + j0 + b];
    }
    for (int d = 0; d < N && !Pc; ++d) {
      if (!wt[d]) continue;
      const double* xr = XT + (size_t)d * n;
      double xi[4], xj[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        xi[a] = (i0 + a < n) ? __ldg(xr + i0 + a) : 0.0;
        xj[a] = (j0 + a < n) ? __ldg(xr + j0 + a) : 0.0;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double df = xi[a] - xj[b];
          P[a][b] *= 1.0 + exp_nonpos(-(df * df) * inv2l2, etab);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) P[a][b] *= yy[a][b];
    for (int c = 0; c < C; ++c) {
      double F[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) F[a][b] = 1.0;
      bool any = false;
      for (int q = mem_off[c]; q < mem_off[c + 1]; ++q) {
        const int d = mem_idx[q];
        if (!wt[d]) continue;
        any = true;
        const double* xr = XT + (size_t)d * n;
        double xi[4], xj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          xi[a] = (i0 + a < n) ? __ldg(xr + i0 + a) : 0.0;
          xj[a] = (j0 + a < n) ? __ldg(xr + j0 + a) : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double df = xi[a] - xj[b];
            F[a][b] *= 1.0 + exp_nonpos(-(df * df) * inv2l2, etab);
          }
      }
      if (!any) continue;
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s += P[a][b] * (1.0 - rcp_ge1(F[a][b]));
      s = warp_sum(s);
      if (lane == 0) sacc[c * 8 + wid] += s;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += sacc[c * 8 + q];
    partial[((size_t)t * GA_BLOCKS + blockIdx.x) * C + c] = s;
  }
}

// ---- Gaussian pair sums, SHARED across the targets of a batch.
// The per-pair factor  R_c,ij = 1 - 1/prod_{d in c}(1 + E_d,ij)  does not depend on the target (as long as every
// member of cluster c is active for it), only  M^t_ij = sym yb^t_i yb^t_j P^t_ij  does.  So for a batch of T targets
//     gau[t][c] = sum_(ij) M^t_ij R_c,ij
// is ONE GEMM (T x pairs) x (pairs x C): every exp / reciprocal is evaluated once per (pair, cluster) instead of once
// per (pair, cluster, target), and the contraction over the pairs runs on the FP64 tensor cores (DMMA m8n8k4; the
// R values are computed by each lane directly in its B-fragment position, M comes from shared memory).  Needs the
// cached product maps P^t (chd_prune_chain's P_state).  Targets for which some alive cluster has only part of its
// members active (multi-variable clusters containing the target itself or an impossible edge) are flagged and served
// by gauss_pairs_kernel instead.
constexpr int GS_IB = 8, GS_JB = 16;          // pair tile: 8 rows x 16 columns = 128 pairs = 32 k-steps
// Two shapes of the contraction (TG targets per group = DMMA row blocks, CW clusters per warp = DMMA column blocks, TH
// threads; the producer geometry needs TG = TH / 8):
//   <32, 64, 256>  batches of up to 32 targets: 8 warps x 64 clusters = 512 clusters per pass
//   <64, 16, 512>  larger batches: all 64 targets share one evaluation of every factor (half the exps of two 32-target
//                  groups) and 16 warps hide the latency of the exp chains (ncu of the first shape at 64 targets: 2 warps
//                  per scheduler, issue slots 26 % busy, DMMA 17 %, FP64 11 %: latency bound); 256 clusters per pass
template <int TG> struct GsCfg {
  static constexpr int MP = TG + 8;           // shared-memory pitch of M[pair][target] (== 8 mod 32: conflict free)
  static constexpr size_t SMEM = (size_t)2 * GS_IB * GS_JB * MP * sizeof(double);
};

// clist = clusters that are alive with at least one active member for SOME target of the batch (compact, in
// increasing order), nc = their number; flagged[t] = 1 if some alive cluster of target t is only partly active.
__global__ void __launch_bounds__(256) cluster_union_kernel(const uint8_t* __restrict__ w, int T, int N, int C,
                                                            const int32_t* __restrict__ mem_off,
                                                            const int32_t* __restrict__ mem_idx,
                                                            int32_t* __restrict__ clist, int32_t* __restrict__ nc,
                                                            uint8_t* __restrict__ flagged) {
  extern __shared__ int32_t s_any[];   // C
  const int tid = threadIdx.x;
  for (int t = tid; t < T; t += 256) flagged[t] = 0;
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    int any = 0;
    const int lo = mem_off[c], hi = mem_off[c + 1];
    for (int t = 0; t < T; ++t) {
      int act = 0;
      for (int q = lo; q < hi; ++q) act += w[(size_t)t * N + mem_idx[q]] ? 1 : 0;
      if (act > 0) {
        any = 1;
        if (act < hi - lo) flagged[t] = 1;      // benign race: every writer stores 1
      }
    }
    s_any[c] = any;
  }
  __syncthreads();
  if (tid == 0) {
    int k = 0;
    for (int c = 0; c < C; ++c)
      if (s_any[c]) clist[k++] = c;
    *nc = k;
  }
}

template <int TG, int CW, int TH>
__global__ void __launch_bounds__(TH, 1) gauss_shared_kernel(const double* __restrict__ X, int n, int N, int C, int T,
                                                              const double* __restrict__ yb,
                                                              const int32_t* __restrict__ mem_off,
                                                              const int32_t* __restrict__ mem_idx,
                                                              const int32_t* __restrict__ clist,
                                                              const int32_t* __restrict__ ncp, double inv2l2,
                                                              double* __restrict__ partial, int nblk,
                                                              const double* __restrict__ Pc, int ldk) {
  __shared__ double etab[64];
  extern __shared__ __align__(16) double gs_smem[];
  static_assert(TG == TH / 8 && TG % 8 == 0 && CW % 8 == 0, "producer geometry");
  constexpr int GS_MP = GsCfg<TG>::MP, MB = TG / 8, NBK = CW / 8, NW = TH / 32;
  constexpr int MS_SZ = GS_IB * GS_JB * GS_MP;                    // M[pair][target], double buffered
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int tg0 = blockIdx.y * TG;                // first target of this group
  const int nc = *ncp;
  exp_table_load(etab);
  const int ib_n = (n + GS_IB - 1) / GS_IB, jb_n = (n + GS_JB - 1) / GS_JB;
  // producer geometry for M: thread -> (target tp, row ip), 16 columns
  const int tp = tid >> 3, ip = tid & 7;
  const bool tp_ok = tg0 + tp < T;
  const double* yt = yb + (size_t)(tp_ok ? tg0 + tp : 0) * n;
  const double* Pt = Pc + (size_t)(tp_ok ? tg0 + tp : 0) * n * ldk;
  // tiles of the lower triangle (block row bi of 8 rows: column blocks 0 .. (8 bi + 7) / 16), dealt round robin
  auto fill = [&](int bi, int bj, double* dst) {
    const int i = bi * GS_IB + ip, j0 = bj * GS_JB;
    double pv[GS_JB];
    const bool row_ok = tp_ok && i < n;
    const double yi = row_ok ? yt[i] : 0.0;
#pragma unroll
    for (int b = 0; b < GS_JB; b += 2) {
      double2 v = make_double2(0.0, 0.0);
      if (row_ok && j0 + b <= i) {
        if (j0 + b + 1 < ldk) v = *reinterpret_cast<const double2*>(Pt + (size_t)i * ldk + j0 + b);
        else v.x = Pt[(size_t)i * ldk + j0 + b];
      }
      pv[b] = v.x; pv[b + 1] = v.y;
    }
#pragma unroll
    for (int b = 0; b < GS_JB; ++b) {
      const int j = j0 + b;
      double m = 0.0;
      if (row_ok && j <= i) m = ((j == i) ? 1.0 : 2.0) * yi * yt[j] * pv[b];
      dst[(ip * GS_JB + b) * GS_MP + tp] = m;
    }
  };
  for (int cpass = 0; cpass * (NW * CW) < nc; ++cpass) {
    const int cb = cpass * (NW * CW) + wid * CW;   // first compact cluster index of this warp
    // this lane's B-fragment clusters: compact index cb + 8 nb + fr
    int cl_lo[NBK], cl_hi[NBK];
#pragma unroll
    for (int nb = 0; nb < NBK; ++nb) {
      const int k = cb + 8 * nb + fr;
      if (k < nc) { const int c = clist[k]; cl_lo[nb] = mem_off[c]; cl_hi[nb] = mem_off[c + 1]; }
      else { cl_lo[nb] = 0; cl_hi[nb] = 0; }
    }
    double acc[MB][NBK][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int nb = 0; nb < NBK; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    // walk this CTA's tiles
    long long tile = blockIdx.x;
    int bi = 0;
    long long row_start = 0;          // index of the first tile of block row bi
    auto locate = [&](long long tl) {   // advance (bi, row_start) to the block row holding tile tl
      while (bi < ib_n) {
        const long long cnt = min((long long)jb_n, (long long)((bi * GS_IB + GS_IB - 1) / GS_JB + 1));
        if (tl < row_start + cnt) return true;
        row_start += cnt;
        ++bi;
      }
      return false;
    };
    int buf = 0;
    bool have = locate(tile);
    if (have) fill(bi, (int)(tile - row_start), gs_smem);
    __syncthreads();
    while (have) {
      const int cbi = bi, cbj = (int)(tile - row_start);
      // prefetch the next tile's M into the other buffer
      const long long ntile = tile + nblk;
      const bool nhave = locate(ntile);
      if (nhave) fill(bi, (int)(ntile - row_start), gs_smem + (buf ^ 1) * MS_SZ);
      const double* M = gs_smem + buf * MS_SZ;
      const int i0 = cbi * GS_IB, j0 = cbj * GS_JB;
      if (cb < nc) {
#pragma unroll 1
        for (int q = 0; q < GS_IB * GS_JB / 4; ++q) {
          const int il = q >> 2, jl = 4 * (q & 3) + fk;
          const int i = min(i0 + il, n - 1), j = min(j0 + jl, n - 1);
          const double* xi = X + (size_t)i * N;
          const double* xj = X + (size_t)j * N;
          double bf[NBK];
#pragma unroll
          for (int nb = 0; nb < NBK; ++nb) {
            double F = 1.0;
            for (int m = cl_lo[nb]; m < cl_hi[nb]; ++m) {
              const int d = mem_idx[m];
              const double df = __ldg(xi + d) - __ldg(xj + d);
              F *= 1.0 + exp_nonpos(-(df * df) * inv2l2, etab);
            }
            bf[nb] = 1.0 - rcp_ge1(F);
          }
          const double* mp = M + (4 * q + fk) * GS_MP + fr;
          double af[MB];
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) af[mb] = mp[8 * mb];
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int nb = 0; nb < NBK; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
        }
      }
      __syncthreads();
      tile = ntile;
      have = nhave;
      buf ^= 1;
    }
    // partial sums of this CTA: rows = targets 8 mb + fr, columns = compact clusters cb + 8 nb + 2 fk + {0, 1}
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      const int t = tg0 + 8 * mb + fr;
      if (t >= T) continue;
#pragma unroll
      for (int nb = 0; nb < NBK; ++nb)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = cb + 8 * nb + 2 * fk + h;
          if (k < nc) partial[((size_t)t * GA_BLOCKS + blockIdx.x) * C + clist[k]] = acc[mb][nb][h];
        }
    }
    __syncthreads();
  }
}

// act, argmin and row removal; one block per target.
__global__ void __launch_bounds__(256) act_finalize_kernel(chd_kernel_desc kd, int N, int C,
                                                           const uint8_t* __restrict__ w,
                                                           const int32_t* __restrict__ mem_off,
                                                           const int32_t* __restrict__ mem_idx,
                                                           const double* __restrict__ gdot,
                                                           const double* __restrict__ rowq,
                                                           const double* __restrict__ gpart, int nblk,
                                                           int nblk_shared, const uint8_t* __restrict__ flagged,
                                                           const double* __restrict__ energy,
                                                           uint8_t* __restrict__ alive, double* __restrict__ act,
                                                           size_t act_stride, int prune, int32_t* __restrict__ pruned,
                                                           size_t pruned_stride) {
  __shared__ double s_val[256];
  __shared__ int s_idx[256];
  const int t = blockIdx.x, tid = threadIdx.x;
  const uint8_t* wt = w + (size_t)t * N;
  const double en = energy[t];
  // Gaussian partials: from gauss_pairs_kernel (nblk blocks) or, for targets it serves, from gauss_shared_kernel
  const int nb_t = (nblk_shared > 0 && !(flagged && flagged[t])) ? nblk_shared : nblk;
  double bv = 0.0; int bi = -1; bool bnan = false;
  for (int c = tid; c < C; c += 256) {
    double lin = 0.0, qr = 0.0;
    bool any = false;
    for (int q = mem_off[c]; q < mem_off[c + 1]; ++q) {
      const int d = mem_idx[q];
      if (!wt[d]) continue;
      any = true;
      const double g = gdot[(size_t)t * N + d];
      lin += g * g;
      if (rowq) qr += rowq[(size_t)t * N + d];
    }
    double v;
    if (!any) {
      v = 2.0;
    } else if (kd.kind == CHD_KERNEL_LINEAR) {
      v = kd.scale * lin / en;
    } else if (kd.kind == CHD_KERNEL_QUADRATIC) {
      v = kd.scale * (2.0 * kd.alpha * lin + qr) / en;
    } else {
      double gs = 0.0;
      for (int b = 0; b < nb_t; ++b) gs += gpart[((size_t)t * GA_BLOCKS + b) * C + c];
      v = (kd.quad_scale * (2.0 * kd.alpha * lin + qr) + kd.scale * gs) / en;
    }
    act[(size_t)t * act_stride + c] = v;
    const bool vn = isnan(v);
    if (bi < 0) { bv = v; bi = c; bnan = vn; }
    else if (!bnan && (vn || v < bv)) { bv = v; bi = c; bnan = vn; }
  }
  if (!prune) return;
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
  if (tid == 0) {
    const int j = s_idx[0];
    pruned[(size_t)t * pruned_stride] = j;
    alive[(size_t)t * C + j] = 0;
  }
}

__global__ void transpose_kernel(const double* __restrict__ X, int n, int N, double* __restrict__ XT) {
  __shared__ double tile[32][33];
  const int i0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, d = d0 + tx;
    tile[r][tx] = (i < n && d < N) ? X[(size_t)i * N + d] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int d = d0 + r, i = i0 + tx;
    if (d < N && i < n) XT[(size_t)d * n + i] = tile[tx][r];
  }
}

// cluster membership in CSR form from cluster_of: one thread per cluster counts and fills its members in increasing
// variable order (stable), thread 0 does the exclusive prefix over the C counts (the former single-thread version
// cost C x N iterations = 13 ms per prune step at C = N = 1154)
__global__ void __launch_bounds__(1024) members_kernel(const int32_t* __restrict__ cluster_of, int N, int C,
                                                        int32_t* __restrict__ mem_off, int32_t* __restrict__ mem_idx) {
  const int tid = threadIdx.x;
  for (int c = tid; c < C; c += blockDim.x) {
    int cnt = 0;
    for (int d = 0; d < N; ++d) cnt += (cluster_of[d] == c) ? 1 : 0;
    mem_off[c + 1] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    mem_off[0] = 0;
    for (int c = 0; c < C; ++c) mem_off[c + 1] += mem_off[c];
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    int pos = mem_off[c];
    for (int d = 0; d < N; ++d)
      if (cluster_of[d] == c) mem_idx[pos++] = d;
  }
}

size_t act_workspace(const Plan* pl, int T) {
  const size_t n = pl->n, N = pl->N, C = pl->C;
  size_t b = 0;
  b += align_up(T * N, 256);                                   // w
  b += align_up((C + 1) * sizeof(int32_t), 256) + align_up(N * sizeof(int32_t), 256);
  b += 2 * align_up(T * N * sizeof(double), 256);              // gdot, rowq
  b += align_up(T * sizeof(double), 256);                      // energy
  b += align_up(T * N * N * sizeof(double), 256);              // Gm
  b += align_up(n * N * sizeof(double), 256);                  // XT
  b += align_up((size_t)T * GA_BLOCKS * C * sizeof(double), 256);
  b += align_up((C + 1) * sizeof(int32_t), 256) + align_up(T, 256);   // clist + nc, flagged
  return b + 4096;
}

int act_launch(const Plan* pl, const chd_kernel_desc* kd, const double* X, const int32_t* cluster_of,
               const uint8_t* w0, uint8_t* alive, const double* yb, const double* ga, int T, double* act,
               size_t act_stride, int prune, int32_t* pruned, size_t pruned_stride, void* ws, size_t ws_bytes,
               cudaStream_t st, const double* P, int ldk) {
  const int n = pl->n, N = pl->N, C = pl->C;
  Carver cv(ws, ws_bytes);
  uint8_t* w = cv.take<uint8_t>((size_t)T * N);
  int32_t* mem_off = cv.take<int32_t>(C + 1);
  int32_t* mem_idx = cv.take<int32_t>(N);
  double* gdot = cv.take<double>((size_t)T * N);
  double* rowq = cv.take<double>((size_t)T * N);
  double* energy = cv.take<double>(T);
  double* Gm = cv.take<double>((size_t)T * N * N);
  double* XT = cv.take<double>((size_t)n * N);
  double* gpart = cv.take<double>((size_t)T * GA_BLOCKS * C);
  int32_t* clist = cv.take<int32_t>(C + 1);
  uint8_t* flagged = cv.take<uint8_t>(T);
  if (!cv.ok) {
    set_error("chd_activations: workspace too small (%zu needed, %zu given)", cv.off, ws_bytes);
    return CHD_ENOSPC;
  }
  members_kernel<<<1, 1024, 0, st>>>(cluster_of, N, C, mem_off, mem_idx);
  CHD_LAUNCHED();
  active_vars_kernel<<<dim3((N + 127) / 128, T), 128, 0, st>>>(w0, alive, cluster_of, N, C, w);
  CHD_LAUNCHED();
  gdot_kernel<<<dim3((N + 31) / 32, T), 256, 0, st>>>(X, n, N, yb, ga, gdot, energy);
  CHD_LAUNCHED();
  const bool quad = kd->kind != CHD_KERNEL_LINEAR;
  int nblk = 0, nblk_shared = 0;
  if (quad) {
    const int nt = (N + 31) / 32;
    gmat_kernel<<<dim3(nt, nt, T), 256, 0, st>>>(X, n, N, yb, Gm);
    CHD_LAUNCHED();
    quad_rows_kernel<<<dim3((N + 7) / 8, T), 256, 0, st>>>(Gm, N, w, cluster_of, rowq);
    CHD_LAUNCHED();
  }
  if (kd->kind == CHD_KERNEL_GAUSSIAN) {
    transpose_kernel<<<dim3((n + 31) / 32, (N + 31) / 32), 256, 0, st>>>(X, n, N, XT);
    CHD_LAUNCHED();
    const int tb = (n + GA_TILE - 1) / GA_TILE;
    const int ntiles = tb * (tb + 1) / 2;
    nblk = ntiles < GA_BLOCKS ? ntiles : GA_BLOCKS;
    const size_t smem = (size_t)C * 8 * sizeof(double);
    const bool shared = P && pl->gauss_shared;
    if (shared) {
      // one evaluation of every per-pair factor for the whole batch + a DMMA contraction over the pairs; the targets
      // it cannot serve (partly active clusters) are flagged and go through gauss_pairs_kernel below
      cluster_union_kernel<<<1, 256, (size_t)C * sizeof(int32_t), st>>>(w, T, N, C, mem_off, mem_idx, clist, clist + C,
                                                                        flagged);
      CHD_LAUNCHED();
      nblk_shared = pl->sm_count < GA_BLOCKS ? pl->sm_count : GA_BLOCKS;
      if (T > 32 && pl->gauss_shared != 2)
        gauss_shared_kernel<64, 16, 512><<<dim3(nblk_shared, (T + 63) / 64), 512, GsCfg<64>::SMEM, st>>>(
            X, n, N, C, T, yb, mem_off, mem_idx, clist, clist + C, 1.0 / (2.0 * kd->l * kd->l), gpart, nblk_shared, P,
            ldk);
      else
        gauss_shared_kernel<32, 64, 256><<<dim3(nblk_shared, (T + 31) / 32), 256, GsCfg<32>::SMEM, st>>>(
            X, n, N, C, T, yb, mem_off, mem_idx, clist, clist + C, 1.0 / (2.0 * kd->l * kd->l), gpart, nblk_shared, P,
            ldk);
      CHD_LAUNCHED();
    }
    gauss_pairs_kernel<<<dim3(nblk, T), 256, smem, st>>>(XT, n, N, C, yb, w, mem_off, mem_idx,
                                                         1.0 / (2.0 * kd->l * kd->l), gpart, nblk, P, ldk,
                                                         shared ? flagged : nullptr);
    CHD_LAUNCHED();
  }
  act_finalize_kernel<<<T, 256, 0, st>>>(*kd, N, C, w, mem_off, mem_idx, gdot, quad ? rowq : nullptr, gpart, nblk,
                                         nblk_shared, flagged, energy, alive, act, act_stride, prune, pruned,
                                         pruned_stride);
  CHD_LAUNCHED();
  return CHD_OK;
}

int act_device_init(Plan* pl) {
  pl->gauss_shared = getenv("CHD_GAUSS_SHARED") ? atoi(getenv("CHD_GAUSS_SHARED")) : 1;
  // CHD_GAUSS_SHARED: 1 (default) shared factors, shape by batch size; 2: always the 32-target shape; 0: per-target kernel
  CHD_CUDA(cudaFuncSetAttribute(gauss_shared_kernel<32, 64, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)GsCfg<32>::SMEM));
  CHD_CUDA(cudaFuncSetAttribute(gauss_shared_kernel<64, 16, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)GsCfg<64>::SMEM));
  // the kernel also has static shared memory: the opt-in limit covers static + dynamic
  cudaFuncAttributes fa;
  CHD_CUDA(cudaFuncGetAttributes(&fa, gauss_pairs_kernel));
  CHD_CUDA(cudaFuncSetAttribute(gauss_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(pl->smem_optin - fa.sharedSizeBytes)));
  return CHD_OK;
}

}  // namespace chd

extern "C" int chd_activations(const chd_plan* plan, const chd_kernel_desc* kd, const double* X,
                               const int32_t* cluster_of, const uint8_t* w0, uint8_t* alive, const double* yb,
                               const double* ga, int T, double* act, int prune, int32_t* pruned, void* workspace,
                               size_t workspace_bytes, void* stream) {
  const chd::Plan* pl = (const chd::Plan*)plan;
  if (!pl || !kd || !X || !cluster_of || !w0 || !alive || !yb || !ga || !act || T <= 0) return CHD_EINVAL;
  if (prune && !pruned) return CHD_EINVAL;
  return chd::act_launch(pl, kd, X, cluster_of, w0, alive, yb, ga, T, act, (size_t)pl->C, prune, pruned, 1, workspace,
                         workspace_bytes, (cudaStream_t)stream, nullptr, 0);
}
