// C-ABI entry points: plan, workspace carving, the fused regression and the device-resident
// prune chain (include/chd_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "internal.cuh"

namespace chd {

std::atomic<unsigned long long> g_launches{0};
static char g_err[1024] = "";
static std::mutex g_err_mu;

void set_error(const char* fmt, ...) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int ZLD = CHD_Z_SAMPLES;  // leading dimension of the Z-test block

struct RegressBuffers {
  uint32_t* subkeys;
  double *B, *d, *e, *beta0, *loss, *eta, *uu, *x;
  TridiagBuffers tb;
  // two-stage path
  BandBuffers bb;
  int* prog;
  double *Lg, *gu;
};

static size_t regress_workspace(const Plan* pl, int T) {
  const size_t n = pl->n;
  size_t b = 0;
  b += align_up(T * 2 * sizeof(uint32_t), 256);
  b += align_up(T * n * ZLD * sizeof(double), 256);
  b += 5 * align_up(T * n * sizeof(double), 256);
  b += align_up(T * sizeof(double), 256);
  b += align_up((size_t)T * CHD_GRID_POINTS * sizeof(double), 256);
  // sized for either path (chd_plan_set_two_stage / chd_tridiag may be used on the same workspace)
  size_t two = band_workspace(pl, T) + chase_workspace(pl, T) +
               align_up((size_t)T * n * CHD_BAND_LD * sizeof(double), 256) + align_up(T * sizeof(double), 256);
  size_t one = tridiag_workspace(pl, T);
  return b + (two > one ? two : one) + 8192;
}

static bool regress_carve(const Plan* pl, int T, Carver& cv, RegressBuffers& rb) {
  const size_t n = pl->n;
  rb.subkeys = cv.take<uint32_t>((size_t)T * 2);
  rb.B = cv.take<double>(T * n * ZLD);
  rb.d = cv.take<double>(T * n);
  rb.e = cv.take<double>(T * n);
  rb.eta = cv.take<double>(T * n);
  rb.uu = cv.take<double>(T * n);
  rb.x = cv.take<double>(T * n);
  rb.beta0 = cv.take<double>(T);
  rb.loss = cv.take<double>((size_t)T * CHD_GRID_POINTS);
  if (pl->two_stage) {
    const bool ok = band_carve(pl, T, cv, rb.bb);
    rb.prog = cv.take<int>(chase_prog_ints(pl, T));
    rb.Lg = cv.take<double>((size_t)T * n * CHD_BAND_LD);
    rb.gu = cv.take<double>(T);
    return ok && cv.ok;
  }
  return tridiag_carve(pl, T, cv, rb.tb) && cv.ok;
}

// key split -> Z samples -> tridiagonalisation (+ Q^T S) -> gamma / solve / Z-test -> yb = -beta0 Q x
static int regress_core(const Plan* pl, double* K, int ldk, const double* ga, int T, const double* grid,
                        int gamma_min_mode, double gamma_min_base, double* gamma_min, const uint32_t* keys,
                        uint32_t* keys_out, int want_lambda_min, double* lambda_min, double* yb, size_t yb_stride,
                        double* noise, double* z_low, double* z_high, double* gamma, size_t out_stride,
                        int32_t* status, const RegressBuffers& rb, cudaStream_t st) {
  int rc;
  if ((rc = chd_split(keys, T, keys_out, rb.subkeys, st))) return rc;
  if ((rc = chd_normal(rb.subkeys, T, pl->n, CHD_Z_SAMPLES, rb.B, ZLD, st))) return rc;
  if (pl->two_stage) {
    // dense -> band (DMMA GEMMs) -> tridiagonal (values only); gamma search on (d, e); solve + Z-test in band
    // coordinates; yb = -beta0 Q x through the block reflectors
    if ((rc = band_reduce_launch(pl, K, ldk, ga, T, rb.B, ZLD, CHD_Z_SAMPLES, rb.beta0, rb.bb, st))) return rc;
    if ((rc = chase_launch(pl, rb.bb.bandW, rb.prog, rb.d, rb.e, T, st))) return rc;
    if ((rc = gamma_launch(pl, rb.d, rb.e, rb.beta0, grid, rb.loss, nullptr, ZLD, 0, gamma_min_mode, gamma_min_base,
                           gamma_min, want_lambda_min, lambda_min, rb.eta, rb.uu, rb.x, noise, z_low, z_high, gamma,
                           out_stride, status, T, st, rb.gu)))
      return rc;
    if ((rc = band_solve_launch(pl, rb.bb.bandS, rb.Lg, rb.gu, rb.B, ZLD, CHD_Z_SAMPLES, rb.x, noise, z_low, z_high,
                                out_stride, T, st)))
      return rc;
    return backtransform2_launch(pl, K, ldk, rb.bb, rb.x, rb.beta0, yb, yb_stride, T, st);
  }
  if ((rc = tridiag_launch(pl, K, ldk, ga, T, rb.B, ZLD, CHD_Z_SAMPLES, rb.d, rb.e, rb.beta0, rb.tb, st))) return rc;
  if ((rc = gamma_launch(pl, rb.d, rb.e, rb.beta0, grid, rb.loss, rb.B, ZLD, CHD_Z_SAMPLES, gamma_min_mode,
                         gamma_min_base, gamma_min, want_lambda_min, lambda_min, rb.eta, rb.uu, rb.x, noise, z_low,
                         z_high, gamma, out_stride, status, T, st)))
    return rc;
  return backtransform_launch(pl, K, ldk, rb.tb, rb.x, rb.beta0, yb, yb_stride, T, st);
}

__global__ void refresh_gamma_min_kernel(const double* __restrict__ lambda_min, double* __restrict__ gamma_min,
                                         int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) {
    const double l = lambda_min[t];
    gamma_min[t] = (l + 1e-9 < 0.0) ? -2.0 * l : 1e-9;   // _GraphDiscoveryMain.py:628-631
  }
}

__global__ void dmma_peak_kernel(int iters, double* sink) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
}

// occupancy probe: `CH` independent accumulator chains per warp (what a GEMM warp tile of CH 8x8 blocks offers)
template <int CH>
__global__ void dmma_chain_kernel(int iters, double* sink) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8 / CH; ++r)
#pragma unroll
      for (int i = 0; i < CH; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
}

// mode 0: DMMA only, 1: DFMA only, 2: both interleaved (8 DFMA per DMMA = equal flops): do the FP64 tensor and the
// FP64 vector pipes add up?
__global__ void fp64_mix_kernel(int iters, int mode, double* sink) {
  double c[8][2];
  double f[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
    if (mode != 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    if (mode != 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
  if (s == 123.456) sink[0] = s;
}

}  // namespace chd

using namespace chd;

extern "C" const char* chd_version(void) { return "chd_b200 0.1.0 (sm_100a)"; }
extern "C" const char* chd_last_error(void) { return g_err; }
extern "C" unsigned long long chd_launch_count(void) { return g_launches.load(); }

extern "C" int chd_plan_create(chd_plan** plan, int n, int N, int C, int device, int ctas_per_matrix) {
  if (!plan || n < 2 || N < 1 || C < 1) return CHD_EINVAL;
  cudaDeviceProp prop;
  CHD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("chd_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return CHD_EINVAL;
  }
  Plan* pl = new Plan();
  pl->n = n; pl->N = N; pl->C = C; pl->device = device;
  pl->sm_count = prop.multiProcessorCount;
  pl->smem_optin = prop.sharedMemPerBlockOptin;
  pl->nb = 32;
  pl->rows_per_chunk = 16;
  pl->two_stage = getenv("CHD_TWO_STAGE") ? atoi(getenv("CHD_TWO_STAGE")) : 1;
  int G = ctas_per_matrix;
  if (G <= 0) {
    if (n <= 256) G = 2;
    else if (n <= 640) G = 4;
    else if (n <= 1536) G = 8;
    else if (n <= 3072) G = 18;
    else G = 37;
  }
  if (G > pl->sm_count) G = pl->sm_count;
  pl->G = G;
  // panel width 32 if the CTA's panel slice fits in shared memory, else 16, else more CTAs per matrix
  // Preferred: 256-thread CTAs, two per SM (of different matrices), so one streams while the other sits in a
  // barrier.  Large n does not leave room for two CTAs' shared memory: one 512-thread CTA per SM then.
  const size_t per_cta = (prop.sharedMemPerMultiprocessor - 2 * prop.reservedSharedMemPerBlock) / 2;
  auto fits = [&](size_t budget) {
    pl->nb = 32;
    if (tridiag_smem_bytes(pl) <= budget) return true;
    pl->nb = 16;
    return tridiag_smem_bytes(pl) <= budget;
  };
  const char* force = getenv("CHD_TD_THREADS");
  pl->bulk = getenv("CHD_TD_BULK") ? atoi(getenv("CHD_TD_BULK")) : 0;
  pl->threads = 256;
  bool ok = (force ? atoi(force) == 256 : n <= 6144) && fits(per_cta < pl->smem_optin ? per_cta : pl->smem_optin);
  if (!ok) {
    pl->threads = 512;
    if (ctas_per_matrix <= 0 && n > 4096) pl->G = 37;
    while (!(ok = fits(pl->smem_optin)) && pl->G < pl->sm_count) pl->G += 1;
  }
  if (!ok) {
    set_error("n=%d does not fit the tridiagonalisation's shared memory even with %d CTAs", n, pl->G);
    delete pl;
    return CHD_EINVAL;
  }
  {
    // spread the CTA slots evenly over the matrices of a wave
    const int slots = (pl->threads == 256 ? 2 : 1) * pl->sm_count;
    pl->wave = slots / pl->G;
    if (pl->wave < 1) pl->wave = 1;
    if (ctas_per_matrix <= 0) {
      const int Gfull = slots / pl->wave;
      if (Gfull > pl->G && Gfull <= pl->sm_count) pl->G = Gfull;
      (void)fits(pl->threads == 256 ? (per_cta < pl->smem_optin ? per_cta : pl->smem_optin) : pl->smem_optin);
    }
  }
  // function attributes and occupancy figures are per device: set them on the plan's device, whatever the
  // caller's current device is
  {
    int cur = 0;
    CHD_CUDA(cudaGetDevice(&cur));
    if (cur != device) CHD_CUDA(cudaSetDevice(device));
    int rc = band_device_init(pl);
    if (!rc) rc = band_tma_device_init(pl);
    if (!rc) rc = chase_device_init(pl);
    if (!rc) rc = tridiag_device_init(pl);
    if (!rc) rc = act_device_init(pl);
    pl->gram_downdate = getenv("CHD_GRAM_DOWNDATE") ? atoi(getenv("CHD_GRAM_DOWNDATE")) : 1;
    if (cur != device) cudaSetDevice(cur);
    if (rc) {
      delete pl;
      return rc;
    }
  }
  *plan = (chd_plan*)pl;
  return CHD_OK;
}

extern "C" void chd_plan_destroy(chd_plan* plan) { delete (Plan*)plan; }

extern "C" int chd_plan_info(const chd_plan* plan, int* ctas_per_matrix, int* matrices_per_wave, int* sm_count) {
  const Plan* pl = (const Plan*)plan;
  if (!pl) return CHD_EINVAL;
  if (ctas_per_matrix) *ctas_per_matrix = pl->G;
  if (matrices_per_wave) *matrices_per_wave = pl->wave;
  if (sm_count) *sm_count = pl->sm_count;
  return CHD_OK;
}

extern "C" int chd_plan_tuning(const chd_plan* plan, int* panel_width, int* threads_per_cta, int* bulk_symv) {
  const Plan* pl = (const Plan*)plan;
  if (!pl) return CHD_EINVAL;
  if (panel_width) *panel_width = pl->nb;
  if (threads_per_cta) *threads_per_cta = pl->threads;
  if (bulk_symv) *bulk_symv = pl->bulk;
  return CHD_OK;
}

extern "C" size_t chd_workspace_bytes(const chd_plan* plan, int T) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || T <= 0) return 0;
  const size_t n = pl->n;
  const size_t ldk = align_up(n, 2);
  size_t gb = 0;
  gram_workspace(pl, T, -1, &gb);
  size_t b = regress_workspace(pl, T);
  b += align_up(T * n * ldk * sizeof(double), 256);  // K work matrices (prune chain)
  b += 3 * align_up((size_t)T * pl->N, 256);         // active / previous / removed variable masks
  b += gb + act_workspace(pl, T);
  b += gram_full_core_bytes(pl);                     // shared linear core G = X X^T (downdate mode of the Gram)
  return b + 8192;
}

extern "C" int chd_regress(const chd_plan* plan, double* K, int ldk, const double* ga, int T,
                           const double* gamma_grid, int gamma_min_mode, double gamma_min_base, double* gamma_min,
                           const uint32_t* keys, double* yb, double* noise, double* z_low, double* z_high,
                           double* gamma, uint32_t* keys_out, double* lambda_min, int32_t* status, void* workspace,
                           size_t workspace_bytes, void* stream) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || !K || !ga || !gamma_grid || !gamma_min || !keys || !yb || !noise || !z_low || !z_high || !gamma ||
      !keys_out || !status || T <= 0 || ldk < pl->n || (ldk & 1))
    return CHD_EINVAL;
  Carver cv(workspace, workspace_bytes);
  RegressBuffers rb;
  if (!regress_carve(pl, T, cv, rb)) {
    set_error("chd_regress: workspace too small (%zu needed, %zu given)", cv.off, workspace_bytes);
    return CHD_ENOSPC;
  }
  return regress_core(pl, K, ldk, ga, T, gamma_grid, gamma_min_mode, gamma_min_base, gamma_min, keys, keys_out,
                      lambda_min != nullptr, lambda_min, yb, (size_t)pl->n, noise, z_low, z_high, gamma, 1, status, rb,
                      (cudaStream_t)stream);
}

extern "C" int chd_tridiag(const chd_plan* plan, double* K, int ldk, const double* ga, int T, double* B, int ldb,
                           int nb_cols, double* d, double* e, double* beta0, void* workspace,
                           size_t workspace_bytes, void* stream) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || !K || !ga || !d || !e || !beta0 || T <= 0 || ldk < pl->n || (ldk & 1)) return CHD_EINVAL;
  Carver cv(workspace, workspace_bytes);
  TridiagBuffers tb;
  if (!tridiag_carve(pl, T, cv, tb)) return CHD_ENOSPC;
  return tridiag_launch(pl, K, ldk, ga, T, B, ldb, B ? nb_cols : 0, d, e, beta0, tb, (cudaStream_t)stream);
}

extern "C" int chd_prune_chain(const chd_plan* plan, const chd_kernel_desc* kd, const double* X,
                               const int32_t* cluster_of, const uint8_t* w0, uint8_t* alive, const double* ga,
                               double* yb, uint32_t* keys, const double* gamma_grid, double* lambda_min,
                               double* gamma_min, int T, int first_step, int steps, int32_t* pruned, double* noise,
                               double* z_low, double* z_high, double* gamma, double* act, double* yb_chain,
                               double* P_state, int* p_valid, int32_t* status, void* workspace,
                               size_t workspace_bytes, void* stream) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || !kd || !X || !cluster_of || !w0 || !alive || !ga || !yb || !keys || !gamma_grid || !lambda_min ||
      !gamma_min || T <= 0 || steps < 0 || !pruned || !noise || !z_low || !z_high || !gamma || !act || !status)
    return CHD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = pl->n, N = pl->N, C = pl->C;
  const int ldk = (int)align_up(n, 2);
  Carver cv(workspace, workspace_bytes);
  RegressBuffers rb;
  bool ok = regress_carve(pl, T, cv, rb);
  double* K = cv.take<double>((size_t)T * n * ldk);
  uint8_t* w = cv.take<uint8_t>((size_t)T * N);
  uint8_t* wprev = cv.take<uint8_t>((size_t)T * N);
  uint8_t* wrem = cv.take<uint8_t>((size_t)T * N);
  size_t gb = 0;
  gram_workspace(pl, T, -1, &gb);
  char* gws = cv.take<char>(gb);
  const size_t ab = act_workspace(pl, T);
  char* aws = cv.take<char>(ab);
  double* Gfull = cv.take<double>((size_t)n * ldk);
  uint8_t* ones = cv.take<uint8_t>((size_t)N);
  if (!ok || !cv.ok) {
    set_error("chd_prune_chain: workspace too small (%zu needed, %zu given)", cv.off, workspace_bytes);
    return CHD_ENOSPC;
  }
  if (ldk > n) CHD_CUDA(cudaMemsetAsync(K, 0, (size_t)T * n * ldk * sizeof(double), st));
  int rc;
  // shared linear core of the full variable set: every step's Gram is a rank-r downdate of it while a target has
  // fewer removed than active variables (gram.cu); one extra Gram per call, amortised over T targets x steps
  const double* Gf = nullptr;
  if (pl->gram_downdate && steps > 0 && (long long)T * steps >= 4) {
    if ((rc = gram_full_core_launch(pl, X, Gfull, ldk, ones, gws, gb, st))) return rc;
    Gf = Gfull;
  }
  for (int s = 0; s < steps; ++s) {
    const long long step = (long long)first_step + s;
    const long long p = (long long)N - step - 1;
    auto is_refresh = [&](long long stp) {
      const long long pp = (long long)N - stp - 1;
      return (stp % 50 == 0) || (pp * (pp + 1) == 2LL * N);
    };
    (void)p;
    if (is_refresh(step)) {
      refresh_gamma_min_kernel<<<(T + 127) / 128, 128, 0, st>>>(lambda_min, gamma_min, T);
      CHD_LAUNCHED();
    }
    // Gaussian mode with a caller-provided product-map cache: P holds prod_{d active}(1 + E_d) of the current
    // mask; it is rebuilt from scratch when invalid and at the reference's own refresh cadence (every 50 steps),
    // otherwise updated by one division per element after the prune.
    const bool use_p = (kd->kind == CHD_KERNEL_GAUSSIAN) && P_state && p_valid;
    if (use_p) {
      if ((rc = active_vars_launch(w0, alive, cluster_of, N, C, T, wprev, st))) return rc;
      if (!*p_valid || step % 50 == 0) {
        if ((rc = gram_launch(pl, kd, X, X, n, wprev, T, nullptr, ldk, gws, gb, st, P_state, 1, nullptr))) return rc;
        *p_valid = 1;
      }
    }
    if ((rc = act_launch(pl, kd, X, cluster_of, w0, alive, yb, ga, T, act + (size_t)s * C, (size_t)steps * C, 1,
                         pruned + s, (size_t)steps, aws, ab, st, use_p ? P_state : nullptr, ldk)))
      return rc;
    if ((rc = active_vars_launch(w0, alive, cluster_of, N, C, T, w, st))) return rc;
    if (use_p) {
      if ((rc = removed_vars_launch(wprev, w, T * N, wrem, st))) return rc;
      if ((rc = gram_launch(pl, kd, X, X, n, w, T, K, ldk, gws, gb, st, P_state, 2, wrem, Gf))) return rc;
    } else {
      if ((rc = gram_launch(pl, kd, X, X, n, w, T, K, ldk, gws, gb, st, nullptr, 0, nullptr, Gf))) return rc;
    }
    if ((rc = regress_core(pl, K, ldk, ga, T, gamma_grid, 0, 0.0, gamma_min, keys, keys, is_refresh(step + 1),
                           lambda_min, yb, (size_t)n, noise + s, z_low + s, z_high + s, gamma + s, (size_t)steps,
                           status, rb, st)))
      return rc;
    if (yb_chain)
      CHD_CUDA(cudaMemcpy2DAsync(yb_chain + (size_t)s * n, (size_t)steps * n * sizeof(double), yb,
                                 (size_t)n * sizeof(double), (size_t)n * sizeof(double), T,
                                 cudaMemcpyDeviceToDevice, st));
  }
  return CHD_OK;
}

// ---- two-stage building blocks (inspection / test entry points)
struct BandWs {
  BandBuffers bb;
  int* prog;
  double *Lg, *dd, *ee;
};
static bool band_ws_carve(const Plan* pl, int T, void* ws, size_t bytes, BandWs& w) {
  Carver cv(ws, bytes);
  const bool ok = band_carve(pl, T, cv, w.bb);
  w.prog = cv.take<int>(chase_prog_ints(pl, T));
  w.Lg = cv.take<double>((size_t)T * pl->n * CHD_BAND_LD);
  return ok && cv.ok;
}

extern "C" size_t chd_band_workspace_bytes(const chd_plan* plan, int T) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || T <= 0) return 0;
  return band_workspace(pl, T) + chase_workspace(pl, T) +
         align_up((size_t)T * pl->n * CHD_BAND_LD * sizeof(double), 256) + 8192;
}

extern "C" int chd_band_reduce(const chd_plan* plan, double* K, int ldk, const double* ga, int T, double* B, int ldb,
                               int nb_cols, double* band, double* beta0, void* workspace, size_t workspace_bytes,
                               void* stream) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || !K || !ga || !band || !beta0 || T <= 0 || ldk < pl->n || (ldk & 1)) return CHD_EINVAL;
  BandWs w;
  if (!band_ws_carve(pl, T, workspace, workspace_bytes, w)) return CHD_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = band_reduce_launch(pl, K, ldk, ga, T, B, ldb, B ? nb_cols : 0, beta0, w.bb, st);
  if (rc) return rc;
  CHD_CUDA(cudaMemcpyAsync(band, w.bb.bandS, (size_t)T * pl->n * CHD_BAND_LD * sizeof(double),
                           cudaMemcpyDeviceToDevice, st));
  return CHD_OK;
}

extern "C" int chd_band_backtransform(const chd_plan* plan, const double* K, int ldk, const double* x,
                                      const double* beta0, double* out, int T, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || !K || !x || !beta0 || !out || T <= 0) return CHD_EINVAL;
  BandWs w;
  if (!band_ws_carve(pl, T, workspace, workspace_bytes, w)) return CHD_ENOSPC;
  return backtransform2_launch(pl, K, ldk, w.bb, x, beta0, out, (size_t)pl->n, T, (cudaStream_t)stream);
}

extern "C" int chd_band_chase(const chd_plan* plan, const double* band, int T, double* d, double* e, void* workspace,
                              size_t workspace_bytes, void* stream) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || !band || !d || !e || T <= 0) return CHD_EINVAL;
  BandWs w;
  if (!band_ws_carve(pl, T, workspace, workspace_bytes, w)) return CHD_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = pl->n;
  CHD_CUDA(cudaMemsetAsync(w.bb.bandW, 0, (size_t)T * n * CHD_CHASE_LD * sizeof(double), st));
  CHD_CUDA(cudaMemcpy2DAsync(w.bb.bandW, CHD_CHASE_LD * sizeof(double), band, CHD_BAND_LD * sizeof(double),
                             (CHD_BAND + 1) * sizeof(double), (size_t)T * n, cudaMemcpyDeviceToDevice, st));
  return chase_launch(pl, w.bb.bandW, w.prog, d, e, T, st);
}

extern "C" int chd_band_solve(const chd_plan* plan, const double* band, const double* gamma_used, double* Bq, int ldb,
                              int nz, int T, double* x, double* noise, double* z_low, double* z_high, void* workspace,
                              size_t workspace_bytes, void* stream) {
  const Plan* pl = (const Plan*)plan;
  if (!pl || !band || !gamma_used || !x || !noise || !z_low || !z_high || T <= 0) return CHD_EINVAL;
  BandWs w;
  if (!band_ws_carve(pl, T, workspace, workspace_bytes, w)) return CHD_ENOSPC;
  return band_solve_launch(pl, band, w.Lg, gamma_used, Bq, ldb, nz, x, noise, z_low, z_high, 1, T,
                           (cudaStream_t)stream);
}

extern "C" int chd_plan_set_two_stage(chd_plan* plan, int on) {
  Plan* pl = (Plan*)plan;
  if (!pl) return CHD_EINVAL;
  pl->two_stage = on ? 1 : 0;
  return CHD_OK;
}

extern "C" int chd_profile_enable(int on) { return profile_enable(on); }
extern "C" int chd_profile_read(double* total_ms, int* launches) { return profile_read(total_ms, launches); }
extern "C" int chd_profile_read_cat(int category, double* total_ms, int* launches) {
  return profile_read_cat(category, total_ms, launches);
}
extern "C" int chd_plan_two_stage(const chd_plan* plan) { return plan ? ((const Plan*)plan)->two_stage : -1; }

extern "C" int chd_fp64_probe(int iters, int mode, double* tflops) {
  if (iters <= 0 || !tflops) return CHD_EINVAL;
  int dev = 0;
  CHD_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CHD_CUDA(cudaGetDeviceProperties(&prop, dev));
  double* sink;
  CHD_CUDA(cudaMalloc(&sink, 8));
  const int blocks = prop.multiProcessorCount * 4, threads = 256;
  fp64_mix_kernel<<<blocks, threads>>>(iters / 10 + 1, mode, sink);
  cudaEvent_t e0, e1;
  CHD_CUDA(cudaEventCreate(&e0));
  CHD_CUDA(cudaEventCreate(&e1));
  CHD_CUDA(cudaEventRecord(e0));
  fp64_mix_kernel<<<blocks, threads>>>(iters, mode, sink);
  CHD_CUDA(cudaEventRecord(e1));
  CHD_CUDA(cudaEventSynchronize(e1));
  g_launches.fetch_add(2);
  float ms = 0;
  CHD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  double per_iter = 0.0;
  if (mode != 1) per_iter += 8.0 * 512.0;
  if (mode != 0) per_iter += 64.0 * 64.0;
  *tflops = (double)blocks * (threads / 32) * (double)iters * per_iter / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return CHD_OK;
}

extern "C" int chd_dmma_occupancy_probe(int iters, int warps_per_sm, int chains, double* tflops) {
  if (iters <= 0 || !tflops || warps_per_sm < 1 || warps_per_sm > 32) return CHD_EINVAL;
  int dev = 0;
  CHD_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CHD_CUDA(cudaGetDeviceProperties(&prop, dev));
  double* sink;
  CHD_CUDA(cudaMalloc(&sink, 8));
  const int blocks = prop.multiProcessorCount, threads = 32 * warps_per_sm;
  auto launch = [&](int it) {
    if (chains == 2) dmma_chain_kernel<2><<<blocks, threads>>>(it, sink);
    else if (chains == 4) dmma_chain_kernel<4><<<blocks, threads>>>(it, sink);
    else dmma_chain_kernel<8><<<blocks, threads>>>(it, sink);
  };
  launch(iters / 10 + 1);
  cudaEvent_t e0, e1;
  CHD_CUDA(cudaEventCreate(&e0));
  CHD_CUDA(cudaEventCreate(&e1));
  CHD_CUDA(cudaEventRecord(e0));
  launch(iters);
  CHD_CUDA(cudaEventRecord(e1));
  CHD_CUDA(cudaEventSynchronize(e1));
  g_launches.fetch_add(2);
  float ms = 0;
  CHD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *tflops = (double)blocks * warps_per_sm * (double)iters * 8.0 * 512.0 / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return CHD_OK;
}

extern "C" int chd_dmma_peak(int iters, double* tflops) {
  if (iters <= 0 || !tflops) return CHD_EINVAL;
  int dev = 0;
  CHD_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CHD_CUDA(cudaGetDeviceProperties(&prop, dev));
  double* sink;
  CHD_CUDA(cudaMalloc(&sink, 8));
  const int blocks = prop.multiProcessorCount * 4, threads = 256;
  dmma_peak_kernel<<<blocks, threads>>>(iters / 10 + 1, sink);  // warm-up
  cudaEvent_t e0, e1;
  CHD_CUDA(cudaEventCreate(&e0));
  CHD_CUDA(cudaEventCreate(&e1));
  CHD_CUDA(cudaEventRecord(e0));
  dmma_peak_kernel<<<blocks, threads>>>(iters, sink);
  CHD_CUDA(cudaEventRecord(e1));
  CHD_CUDA(cudaEventSynchronize(e1));
  g_launches.fetch_add(2);
  float ms = 0;
  CHD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  const double flops = (double)blocks * (threads / 32) * (double)iters * 8.0 * 512.0;
  *tflops = flops / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return CHD_OK;
}
