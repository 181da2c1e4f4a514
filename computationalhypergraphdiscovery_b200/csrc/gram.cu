// Mode-kernel (Gram) matrix assembly.
//   compact_kernel : per-target list of active variable indices (gather instead of the
//                    reference's multiply-by-mask; adds exact zeros only).
//   gather_kernel  : Xc[t] = X[:, active]  (n x ldp, zero padded to a multiple of GK = 32 columns)
//   gram_kernel    : L = Xq_c Xc^T on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA.8x8x4) with the
//                    polynomial / product-of-RBF map fused into the k-loop + epilogue:
//                      linear    1 + s L                              (Modes/kernels.py:99-102)
//                      quadratic s (alpha + L)^2 + 1 - alpha^2 s      (Modes/kernels.py:169-173)
//                      gaussian  s prod_d (1 + exp(-(x_d-y_d)^2/2l^2)) + quad - 1   (Modes/kernels.py:251-257)
//                    The n x n x N `exps` tensor of the reference (kernels.py:307-309) is never formed.
//   DOWNDATE mode   : the linear core of the full variable set G = X X^T does not depend on the target.  When a
//                    target has fewer REMOVED than active variables (always, in the first half of a prune chain),
//                    its core is L = G - X_r X_r^T with the removed columns X_r gathered instead of the active ones:
//                    2 n^2 r flop (r = 1 in stage 1, r = steps so far in a chain) instead of 2 n^2 p.  G is computed
//                    once per chd_gram / chd_prune_chain call into the caller's workspace.
#include "internal.cuh"

namespace chd {

__global__ void compact_kernel(const uint8_t* __restrict__ w, int N, int ldi, int32_t* __restrict__ idx,
                               int32_t* __restrict__ pcount, int invert = 0) {
  // one warp per target; invert: list the variables with w == 0 instead
  const int t = blockIdx.x;
  const int lane = threadIdx.x;
  const uint8_t* wt = w + (size_t)t * N;
  int32_t* it = idx + (size_t)t * ldi;
  int base = 0;
  for (int d0 = 0; d0 < N; d0 += 32) {
    int d = d0 + lane;
    bool on = (d < N) && ((wt[d] != 0) != (invert != 0));
    unsigned m = __ballot_sync(0xffffffffu, on);
    if (on) it[base + __popc(m & ((1u << lane) - 1u))] = d;
    base += __popc(m);
  }
  if (lane == 0) pcount[t] = base;
}

// per target: pick the shorter of the active / removed variable lists (down[t] = 1: removed list, L = G - X_r X_r^T)
__global__ void choose_list_kernel(const int32_t* __restrict__ idx_act, const int32_t* __restrict__ cnt_act,
                                   const int32_t* __restrict__ idx_rem, const int32_t* __restrict__ cnt_rem, int ldi,
                                   int allow_down, int32_t* __restrict__ idx_sel, int32_t* __restrict__ cnt_sel,
                                   int32_t* __restrict__ down) {
  const int t = blockIdx.x;
  const int pa = cnt_act[t], pr = cnt_rem[t];
  const bool dn = allow_down && pr < pa;
  const int32_t* src = (dn ? idx_rem : idx_act) + (size_t)t * ldi;
  const int cnt = dn ? pr : pa;
  for (int q = threadIdx.x; q < cnt; q += blockDim.x) idx_sel[(size_t)t * ldi + q] = src[q];
  if (threadIdx.x == 0) { cnt_sel[t] = cnt; down[t] = dn ? 1 : 0; }
}

__global__ void gather_kernel(const double* __restrict__ X, int n, int N, const int32_t* __restrict__ idx, int ldi,
                              const int32_t* __restrict__ pcount, double* __restrict__ Xc, int ldp) {
  const int t = blockIdx.y;
  const int p = pcount[t];
  const int32_t* it = idx + (size_t)t * ldi;
  double* out = Xc + (size_t)t * n * ldp;
  const size_t total = (size_t)n * ldp;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int i = (int)(e / ldp), q = (int)(e - (size_t)i * ldp);
    out[e] = (q < p) ? X[(size_t)i * N + it[q]] : 0.0;
  }
}

constexpr int CHD_KERNEL_RAW = 99;   // internal: K = L (the linear core itself), used for the shared G = X X^T
constexpr int GT = 64;        // CTA tile (GT x GT)
constexpr int GK = 32;        // k chunk
constexpr int GS = GK + 4;    // smem row stride (== 4 mod 8 -> conflict-free DMMA fragment loads)

// A = rows of Xq_c (nq x ldp), B = rows of Xc (n x ldp).  K[t] is nq x ldk.
// SYMM: Xq == X, only tiles with bi >= bj are computed and mirrored.
template <bool SYMM>
__global__ void __launch_bounds__(256, 2) gram_kernel(const double* __restrict__ XcA, const double* __restrict__ XcB,
                                                   int nq, int n, int ldp, const int32_t* __restrict__ pcount,
                                                   double* __restrict__ K, int ldk, chd_kernel_desc kd,
                                                   double* __restrict__ P, int pmode,
                                                   const double* __restrict__ Xr, const int32_t* __restrict__ rcount,
                                                   const double* __restrict__ Gfull,
                                                   const int32_t* __restrict__ down) {
  // Gaussian product-map cache P = prod_{d active}(1 + E_d):
  //   pmode 0: no cache;  1: compute the product and store it in P (K may be null: product only);
  //   2: P <- P / prod_{d removed}(1 + E_d) with the removed variables gathered in Xr (n x ldp), no full product.
  __shared__ __align__(16) double As[GT * GS];
  __shared__ __align__(16) double Bs[GT * GS];
  __shared__ double etab[64];
  exp_table_load(etab);     // visible after the first __syncthreads of the k-loop / the divisor loop
  const int t = blockIdx.y;
  int bi, bj;
  if (SYMM) {
    // linear tile id -> (bi >= bj)
    int id = blockIdx.x;
    bi = (int)((sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= id) ++bi;
    while (bi * (bi + 1) / 2 > id) --bi;
    bj = id - bi * (bi + 1) / 2;
  } else {
    const int tiles_j = (n + GT - 1) / GT;
    bi = blockIdx.x / tiles_j;
    bj = blockIdx.x - bi * tiles_j;
  }
  const int p = pcount[t];
  const bool dn = Gfull && down && down[t];      // gathered columns are the REMOVED ones: L = G - acc
  const double* A = XcA + (size_t)t * nq * ldp;
  const double* B = XcB + (size_t)t * n * ldp;
  double* Kt = K + (size_t)t * nq * ldk;
  const int row0 = bi * GT, col0 = bj * GT;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wm = wid >> 2, wn = wid & 3;  // 2 x 4 warps, warp tile 32 x 16
  const int fr = lane >> 2, fk = lane & 3;

  double acc[4][2][2];
  double prod[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      acc[a][b][0] = acc[a][b][1] = 0.0;
      prod[a][b][0] = prod[a][b][1] = 1.0;
    }
  const bool gauss = kd.kind == CHD_KERNEL_GAUSSIAN;
  const double inv2l2 = gauss ? 1.0 / (2.0 * kd.l * kd.l) : 0.0;

  const int kend = (p + GK - 1) / GK * GK;
  // tiles of 64 rows x 32 k: 1024 double2 per operand, 4 per thread; the next chunk is fetched into registers
  // while the current one is on the tensor cores
  double2 pa[4], pb[4];
  auto gload = [&](const double* __restrict__ SA, const double* __restrict__ SB, int k0) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int id = tid + s * 256, r = id >> 4, c2 = id & 15;
      pa[s] = (row0 + r < nq) ? *reinterpret_cast<const double2*>(SA + (size_t)(row0 + r) * ldp + k0 + 2 * c2)
                              : make_double2(0.0, 0.0);
      pb[s] = (col0 + r < n) ? *reinterpret_cast<const double2*>(SB + (size_t)(col0 + r) * ldp + k0 + 2 * c2)
                             : make_double2(0.0, 0.0);
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int id = tid + s * 256, r = id >> 4, c2 = id & 15;
      *reinterpret_cast<double2*>(&As[r * GS + 2 * c2]) = pa[s];
      *reinterpret_cast<double2*>(&Bs[r * GS + 2 * c2]) = pb[s];
    }
  };
  if (kend > 0) gload(A, B, 0);
  for (int k0 = 0; k0 < kend; k0 += GK) {
    sstore();
    __syncthreads();
    if (k0 + GK < kend) gload(A, B, k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; kk += 4) {
      double af[4], bf[2];
#pragma unroll
      for (int a = 0; a < 4; ++a) af[a] = As[(wm * 32 + a * 8 + fr) * GS + kk + fk];
#pragma unroll
      for (int b = 0; b < 2; ++b) bf[b] = Bs[(wn * 16 + b * 8 + fr) * GS + kk + fk];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    if (gauss && pmode != 2 && !dn) {
      const int dmax = min(GK, p - k0);
      for (int dd = 0; dd < dmax; ++dd) {
        double xr[4], xc[2][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) xr[a] = As[(wm * 32 + a * 8 + fr) * GS + dd];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          xc[b][0] = Bs[(wn * 16 + b * 8 + 2 * fk) * GS + dd];
          xc[b][1] = Bs[(wn * 16 + b * 8 + 2 * fk + 1) * GS + dd];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              double df = xr[a] - xc[b][c];
              prod[a][b][c] *= 1.0 + exp_nonpos(-(df * df) * inv2l2, etab);
            }
      }
    }
    __syncthreads();
  }

  if (gauss && pmode == 2) {
    // divisor F = prod over the removed variables (usually one), same tile staging as above
    const int pr = rcount[t];
    const double* R = Xr + (size_t)t * n * ldp;
    for (int k0 = 0; k0 < pr; k0 += GK) {
      gload(R, R, k0);
      sstore();
      __syncthreads();
      const int dmax = min(GK, pr - k0);
      for (int dd = 0; dd < dmax; ++dd) {
        double xr[4], xc[2][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) xr[a] = As[(wm * 32 + a * 8 + fr) * GS + dd];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          xc[b][0] = Bs[(wn * 16 + b * 8 + 2 * fk) * GS + dd];
          xc[b][1] = Bs[(wn * 16 + b * 8 + 2 * fk + 1) * GS + dd];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              double df = xr[a] - xc[b][c];
              prod[a][b][c] *= 1.0 + exp_nonpos(-(df * df) * inv2l2, etab);
            }
      }
      __syncthreads();
    }
  }
  double* Pt = P ? P + (size_t)t * nq * ldk : nullptr;

  // epilogue
  const double s = kd.scale, al = kd.alpha, qs = kd.quad_scale;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = row0 + wm * 32 + a * 8 + fr;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int j = col0 + wn * 16 + b * 8 + 2 * fk;
      double v[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double L = acc[a][b][c];
        if (dn) L = ((i < nq) && (j + c < n)) ? Gfull[(size_t)i * ldk + j + c] - L : 0.0;
        double r;
        if (kd.kind == CHD_KERNEL_RAW) {
          r = L;
        } else if (kd.kind == CHD_KERNEL_LINEAR) {
          r = 1.0 + s * L;
        } else if (kd.kind == CHD_KERNEL_QUADRATIC) {
          const double u = al + L;
          r = s * (u * u) + (1.0 - al * al * s);
        } else {
          const double u = al + L;
          const double q = qs * (u * u) + (1.0 - al * al * qs);
          double pv = prod[a][b][c];
          if (pmode == 2) {
            const bool in = (i < nq) && (j + c < n);
            pv = in ? Pt[(size_t)i * ldk + j + c] / pv : 1.0;
          }
          prod[a][b][c] = pv;
          r = (s * pv + q) - 1.0;
        }
        v[c] = r;
      }
      if (i < nq && Pt && gauss) {
        if (j + 1 < n) {
          *reinterpret_cast<double2*>(Pt + (size_t)i * ldk + j) = make_double2(prod[a][b][0], prod[a][b][1]);
        } else if (j < n) {
          Pt[(size_t)i * ldk + j] = prod[a][b][0];
        }
        if (SYMM && bi != bj) {
          if (j < n) Pt[(size_t)j * ldk + i] = prod[a][b][0];
          if (j + 1 < n) Pt[(size_t)(j + 1) * ldk + i] = prod[a][b][1];
        }
      }
      if (i < nq && K) {
        if (j + 1 < n) {
          *reinterpret_cast<double2*>(Kt + (size_t)i * ldk + j) = make_double2(v[0], v[1]);
        } else if (j < n) {
          Kt[(size_t)i * ldk + j] = v[0];
        }
        if (SYMM && bi != bj) {
          if (j < n) Kt[(size_t)j * ldk + i] = v[0];
          if (j + 1 < n) Kt[(size_t)(j + 1) * ldk + i] = v[1];
        }
      }
    }
  }
}

// out[t][q] = sum_j K[t][q][j] * (-yb[t][j])   (predict, _GraphDiscoveryMain.py:973)
__global__ void predict_gemv_kernel(const double* __restrict__ K, int nq, int n, int ldk,
                                    const double* __restrict__ yb, double* __restrict__ out) {
  const int t = blockIdx.y;
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * warps + wid;
  if (q >= nq) return;
  const double* row = K + ((size_t)t * nq + q) * ldk;
  const double* y = yb + (size_t)t * n;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += row[j] * (-y[j]);
  s = warp_sum(s);
  if (lane == 0) out[(size_t)t * nq + q] = s;
}

int gram_workspace(const Plan* pl, int T, int nq, size_t* bytes) {
  const int ldp = (int)align_up(pl->N, GK);
  size_t b = 0;
  b += align_up((size_t)T * pl->N * sizeof(int32_t), 256);
  b += align_up((size_t)T * sizeof(int32_t), 256);
  b += 2 * align_up((size_t)T * pl->n * ldp * sizeof(double), 256);   // Xc and the removed-variable gather
  b += align_up((size_t)T * pl->N * sizeof(int32_t), 256) + align_up((size_t)T * sizeof(int32_t), 256);
  b += 2 * (align_up((size_t)T * pl->N * sizeof(int32_t), 256) + align_up((size_t)T * sizeof(int32_t), 256));  // removed / chosen lists
  b += align_up((size_t)T * sizeof(int32_t), 256);                                                            // down flags
  if (nq >= 0) b += align_up((size_t)T * nq * ldp * sizeof(double), 256);
  *bytes = b + 1024;
  return ldp;
}

int gram_launch(const Plan* pl, const chd_kernel_desc* kd, const double* X, const double* Xq, int nq,
                const uint8_t* w, int T, double* K, int ldk, void* ws, size_t ws_bytes, cudaStream_t st,
                double* P, int pmode, const uint8_t* wrem, const double* Gfull) {
  const int n = pl->n, N = pl->N;
  const bool symm = (Xq == X) && (nq == n);
  Carver cv(ws, ws_bytes);
  const int ldp = (int)align_up(N, GK);
  int32_t* idx = cv.take<int32_t>((size_t)T * N);
  int32_t* pc = cv.take<int32_t>(T);
  double* Xc = cv.take<double>((size_t)T * n * ldp);
  double* Xr = cv.take<double>((size_t)T * n * ldp);
  int32_t* idx2 = cv.take<int32_t>((size_t)T * N);
  int32_t* pc2 = cv.take<int32_t>(T);
  int32_t* idx_rem = cv.take<int32_t>((size_t)T * N);
  int32_t* cnt_rem = cv.take<int32_t>(T);
  int32_t* idx_sel = cv.take<int32_t>((size_t)T * N);
  int32_t* cnt_sel = cv.take<int32_t>(T);
  int32_t* down = cv.take<int32_t>(T);
  double* Xqc = symm ? Xc : cv.take<double>((size_t)T * nq * ldp);
  if (kd->kind != CHD_KERNEL_GAUSSIAN || !P) pmode = 0;
  // downdate from the shared full-set core: symmetric train Gram only; a Gaussian full product (no cached map, or a
  // rebuild of it) needs the active columns staged, so it keeps the direct form
  if (!symm || (kd->kind == CHD_KERNEL_GAUSSIAN && pmode != 2)) Gfull = nullptr;
  if (pmode == 2 && (!symm || !wrem)) return CHD_EINVAL;
  if (!cv.ok) {
    set_error("chd_gram: workspace too small (%zu needed, %zu given)", cv.off, ws_bytes);
    return CHD_ENOSPC;
  }
  compact_kernel<<<T, 32, 0, st>>>(w, N, N, idx, pc);
  CHD_LAUNCHED();
  if (Gfull) {
    compact_kernel<<<T, 32, 0, st>>>(w, N, N, idx_rem, cnt_rem, 1);
    CHD_LAUNCHED();
    choose_list_kernel<<<T, 128, 0, st>>>(idx, pc, idx_rem, cnt_rem, N, 1, idx_sel, cnt_sel, down);
    CHD_LAUNCHED();
    idx = idx_sel;
    pc = cnt_sel;
  }
  {
    size_t total = (size_t)n * ldp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    gather_kernel<<<dim3(blocks, T), 256, 0, st>>>(X, n, N, idx, N, pc, Xc, ldp);
    CHD_LAUNCHED();
    if (!symm) {
      total = (size_t)nq * ldp;
      blocks = (int)((total + 255) / 256);
      if (blocks > 1184) blocks = 1184;
      gather_kernel<<<dim3(blocks, T), 256, 0, st>>>(Xq, nq, N, idx, N, pc, Xqc, ldp);
      CHD_LAUNCHED();
    }
  }
  if (pmode == 2) {
    compact_kernel<<<T, 32, 0, st>>>(wrem, N, N, idx2, pc2);
    CHD_LAUNCHED();
    size_t total = (size_t)n * ldp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    gather_kernel<<<dim3(blocks, T), 256, 0, st>>>(X, n, N, idx2, N, pc2, Xr, ldp);
    CHD_LAUNCHED();
  }
  if (symm) {
    const int tb = (n + GT - 1) / GT;
    gram_kernel<true><<<dim3(tb * (tb + 1) / 2, T), 256, 0, st>>>(Xc, Xc, n, n, ldp, pc, K, ldk, *kd, P, pmode, Xr,
                                                                 pc2, Gfull, Gfull ? down : nullptr);
  } else {
    const int ti = (nq + GT - 1) / GT, tj = (n + GT - 1) / GT;
    gram_kernel<false><<<dim3(ti * tj, T), 256, 0, st>>>(Xqc, Xc, nq, n, ldp, pc, K, ldk, *kd, nullptr, 0, Xr, pc2,
                                                         nullptr, nullptr);
  }
  CHD_LAUNCHED();
  return CHD_OK;
}

// G = X X^T over ALL N variables (raw linear core, n x ldk), for the downdate mode; `ones` = N bytes of scratch
int gram_full_core_launch(const Plan* pl, const double* X, double* G, int ldk, uint8_t* ones, void* ws,
                          size_t ws_bytes, cudaStream_t st) {
  CHD_CUDA(cudaMemsetAsync(ones, 1, (size_t)pl->N, st));
  chd_kernel_desc raw = {};
  raw.kind = CHD_KERNEL_RAW;
  raw.scale = 1.0;
  return gram_launch(pl, &raw, X, X, pl->n, ones, 1, G, ldk, ws, ws_bytes, st, nullptr, 0, nullptr, nullptr);
}

// bytes of the shared core + its mask, carved in front of a Gram workspace
size_t gram_full_core_bytes(const Plan* pl) {
  return align_up((size_t)pl->n * align_up(pl->n, 2) * sizeof(double), 256) + align_up((size_t)pl->N, 256);
}

}  // namespace chd

extern "C" int chd_gram(const chd_plan* plan, const chd_kernel_desc* kd, const double* X, const double* Xq, int nq,
                        const uint8_t* w, int T, double* K, int ldk, void* workspace, size_t workspace_bytes,
                        void* stream) {
  const chd::Plan* pl = (const chd::Plan*)plan;
  if (!pl || !kd || !X || !Xq || !w || !K || T <= 0 || nq <= 0 || ldk < pl->n || (ldk & 1)) return CHD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  // train Gram of a batch (stage 1: every target lacks only itself and its impossible edges): one shared core
  // G = X X^T, then a rank-r downdate per target
  const double* G = nullptr;
  char* ws = (char*)workspace;
  size_t left = workspace_bytes;
  if (pl->gram_downdate && Xq == X && nq == pl->n && T >= 4 && ldk == (int)chd::align_up(pl->n, 2)) {
    const size_t gb = chd::gram_full_core_bytes(pl);
    if (left > gb) {
      double* Gw = (double*)ws;
      uint8_t* ones = (uint8_t*)(ws + chd::align_up((size_t)pl->n * ldk * sizeof(double), 256));
      ws += gb; left -= gb;
      const int rc = chd::gram_full_core_launch(pl, X, Gw, ldk, ones, ws, left, st);
      if (rc) return rc;
      G = Gw;
    }
  }
  return chd::gram_launch(pl, kd, X, Xq, nq, w, T, K, ldk, ws, left, st, nullptr, 0, nullptr, G);
}

extern "C" int chd_predict(const chd_plan* plan, const chd_kernel_desc* kd, const double* X, const double* Xq,
                           int nq, const uint8_t* w, const double* yb, int T, double* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  const chd::Plan* pl = (const chd::Plan*)plan;
  if (!pl || !kd || !X || !Xq || !w || !yb || !out || T <= 0 || nq <= 0) return CHD_EINVAL;
  const int n = pl->n;
  const int ldk = (int)chd::align_up(n, 2);
  chd::Carver cv(workspace, workspace_bytes);
  double* K = cv.take<double>((size_t)T * nq * ldk);
  if (!cv.ok) return CHD_ENOSPC;
  size_t used = chd::align_up(cv.off, 256);
  // cross-Gram always takes the non-symmetric path: pass distinct pointer semantics via nq/Xq
  int rc = chd::gram_launch(pl, kd, X, Xq, nq, w, T, K, ldk, (char*)workspace + used, workspace_bytes - used,
                            (cudaStream_t)stream, nullptr, 0, nullptr, nullptr);
  if (rc) return rc;
  const int warps = 8;
  chd::predict_gemv_kernel<<<dim3((nq + warps - 1) / warps, T), warps * 32, 0, (cudaStream_t)stream>>>(
      K, nq, n, ldk, yb, out);
  CHD_LAUNCHED();
  return CHD_OK;
}
