/*
 * chd_b200.h -- C ABI of the B200-native CHD hot path (libchd_b200.so).
 *
 * The reference (TheoBourdais/ComputationalHypergraphDiscovery v2.1.0) has no FFI: its
 * boundary for this path is the table of jit/vmap closures built by
 * GraphDiscovery.prepare_functions (src/ComputationalHypergraphDiscovery/_GraphDiscoveryMain.py:225-270,
 * listed in UNPICKLABLE_ATTRS, :48-55).  Each entry point below replaces one slot of that
 * table (cited per function) and is what a ctypes binding in the reference would call
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors on the host
 *     side); the library never allocates or frees device memory;
 *   - FP64 everywhere, row-major, contiguous; keys are uint32[T][2] (JAX threefry keys);
 *   - work is enqueued on the caller's cudaStream_t (passed as void*), no host sync
 *     unless stated;
 *   - every function returns 0 on success, a negative CHD_E* code otherwise, never throws;
 *   - no global state: everything hangs off the chd_plan the caller creates per device.
 */
#ifndef CHD_B200_H
#define CHD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHD_OK 0
#define CHD_EINVAL (-1)   /* bad argument */
#define CHD_ECUDA (-2)    /* CUDA runtime error (see chd_last_error) */
#define CHD_ENOSPC (-3)   /* workspace too small */

#define CHD_KERNEL_LINEAR 0
#define CHD_KERNEL_QUADRATIC 1
#define CHD_KERNEL_GAUSSIAN 2

#define CHD_Z_SAMPLES 100      /* interpolatory.py:98 */
#define CHD_GRID_POINTS 10000  /* interpolatory.py:129 */

/* One "mode kernel" as configured by ModeKernel.setup (Modes/kernels.py:108-123,187-209,293-317):
 *   linear    K = 1 + scale * L                        L = (X*w) X^T
 *   quadratic K = scale*(alpha+L)^2 + 1 - alpha^2*scale
 *   gaussian  K = scale*prod_d(1+w_d*exp(-(x_d-y_d)^2/(2 l^2))) + [quad_scale*(alpha+L)^2 + 1 - alpha^2*quad_scale] - 1 */
typedef struct chd_kernel_desc {
  int32_t kind;       /* CHD_KERNEL_* */
  int32_t reserved;
  double scale;
  double alpha;       /* 0.5 * scales["linear"] / quadratic scale */
  double quad_scale;  /* gaussian only: scale of its internal quadratic part */
  double l;           /* gaussian only: length scale */
} chd_kernel_desc;

typedef struct chd_plan chd_plan;

/* Library / build info. */
const char* chd_version(void);
const char* chd_last_error(void);
/* Number of kernels of this library launched since load (bench.py's gpu_launches). */
unsigned long long chd_launch_count(void);

/* Plan: problem sizes + device properties + tuning.  n samples, N variables, C clusters.
 * ctas_per_matrix <= 0 lets the library choose.  Host-only object. */
int chd_plan_create(chd_plan** plan, int n, int N, int C, int device, int ctas_per_matrix);
void chd_plan_destroy(chd_plan* plan);
/* Bytes of scratch the batched calls below need for T targets in flight. */
size_t chd_workspace_bytes(const chd_plan* plan, int T);
int chd_plan_info(const chd_plan* plan, int* ctas_per_matrix, int* matrices_per_wave, int* sm_count);
/* Tuning of the tridiagonalisation kernel: panel width, threads per CTA (256: two CTAs per SM, 512: one),
 * symv data path (1: cp.async.bulk tile ring, 0: register loads). */
int chd_plan_tuning(const chd_plan* plan, int* panel_width, int* threads_per_cta, int* bulk_symv);

/* jax.random.normal(key, (rows, cols)) for T keys (interpolatory.py:99):
 * out[t][r][c], leading dimension ld >= cols. */
int chd_normal(const uint32_t* keys, int T, int rows, int cols, double* out, int ld, void* stream);
/* jax.random.split(key) for T keys: new_key[t] = row 0, subkey[t] = row 1 (interpolatory.py:47). */
int chd_split(const uint32_t* keys, int T, uint32_t* new_keys, uint32_t* subkeys, void* stream);

/* Slot `vmaped_kernel[k]` (_GraphDiscoveryMain.py:251-255; Modes/kernels.py:99-102,169-173,251-257):
 * K[t] = kernel(Xq, X, w[t]) for T variable masks w (T x N, uint8 0/1).
 * X is n x N row-major (train); Xq is nq x N (pass X and nq=n for the symmetric train Gram).
 * K is T x nq x ldk. */
int chd_gram(const chd_plan* plan, const chd_kernel_desc* kd, const double* X, const double* Xq, int nq,
             const uint8_t* w, int T, double* K, int ldk, void* workspace, size_t workspace_bytes, void* stream);

/* Slot `*_regression_find_gamma` (_GraphDiscoveryMain.py:239-250; interpolatory.py:24-135):
 * for each of T Gram matrices (n x ldk, DESTROYED) and targets ga (T x n):
 *   key,subkey = split(key); gamma = find_gamma(...); yb, noise = solve; Z = Z_test(subkey)
 * gamma_min_mode 0: gamma_min[t] is used as given.
 * gamma_min_mode 1: gamma_min[t] = (lambda_min + base < 0) ? -2*lambda_min : base with
 *                   base = gamma_min_base (stage-1 rule, _GraphDiscoveryMain.py:476-479); written back.
 * Outputs (each T-long unless noted): yb (T x n), noise, z_low, z_high, gamma (raw, un-clamped),
 * keys_out (T x 2), lambda_min (smallest eigenvalue of K), status (int32 bit flags, see CHD_ST_*).
 * gamma_grid: the CHD_GRID_POINTS values of jnp.logspace(-6, 6, 10000) (interpolatory.py:129), uploaded once by
 * the host so that a returned grid value is bit-identical to the host's. */
#define CHD_ST_BFGS_SUCCESS 1
#define CHD_ST_USED_MEDIAN 2
#define CHD_ST_GRID_EDGE 4
int chd_regress(const chd_plan* plan, double* K, int ldk, const double* ga, int T, const double* gamma_grid,
                int gamma_min_mode, double gamma_min_base, double* gamma_min, const uint32_t* keys, double* yb,
                double* noise,
                double* z_low, double* z_high, double* gamma, uint32_t* keys_out, double* lambda_min,
                int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Activations of helper_functions.py:40-134 for T targets:
 *   act[t][c] = yb^T Infl_c yb / energy, 2.0 for empty rows.
 * cluster_of (N int32) maps a variable to its cluster; w0 (T x N uint8) is the initial variable
 * mask (target column and impossible edges removed); alive (T x C uint8) marks rows not yet pruned.
 * If prune != 0 also does the argmin + row removal of helper_functions.py:168-169:
 * pruned[t] = argmin index, alive[t][argmin] = 0. */
int chd_activations(const chd_plan* plan, const chd_kernel_desc* kd, const double* X, const int32_t* cluster_of,
                    const uint8_t* w0, uint8_t* alive, const double* yb, const double* ga, int T, double* act,
                    int prune, int32_t* pruned, void* workspace, size_t workspace_bytes, void* stream);

/* Slot `ancestor_finding_step_funcs[k]` driven by prune_ancestors (_GraphDiscoveryMain.py:559-701;
 * helper_functions.py:166-176): runs `steps` prune steps for T targets that chose kernel `kd`,
 * device-resident (no host round trip per step).  State in/out: alive (T x C), yb (T x n), keys (T x 2),
 * lambda_min (T; smallest eigenvalue of the current Gram, from the previous regression).
 * first_step is the reference's `step` index of the first step run here (gamma_min is refreshed from
 * lambda_min when step % 50 == 0 or p(p+1)/2 == N, p = N-step-1, _GraphDiscoveryMain.py:622-631);
 * gamma_min (T) carries the value between calls.
 * Chain outputs, [t][s] for s in [0, steps): pruned (int32), noise, z_low, z_high, gamma,
 * act (T x steps x C), yb_chain (T x steps x n, may be NULL).
 * P_state (T x n x ldk, may be NULL) + p_valid (HOST int, in/out): optional cache of the Gaussian product map
 * prod_{d active}(1 + E_d) carried between calls; when given, a prune step costs one division per element instead
 * of a full product (rebuilt when *p_valid == 0 and every 50 steps). */
int chd_prune_chain(const chd_plan* plan, const chd_kernel_desc* kd, const double* X, const int32_t* cluster_of,
                    const uint8_t* w0, uint8_t* alive, const double* ga, double* yb, uint32_t* keys,
                    const double* gamma_grid, double* lambda_min, double* gamma_min, int T, int first_step,
                    int steps, int32_t* pruned,
                    double* noise, double* z_low, double* z_high, double* gamma, double* act, double* yb_chain,
                    double* P_state, int* p_valid, int32_t* status, void* workspace, size_t workspace_bytes,
                    void* stream);

/* predict (_GraphDiscoveryMain.py:934-979): out[t][q] = sum_j K(Xq, X, w[t])[q][j] * (-yb[t][j]). */
int chd_predict(const chd_plan* plan, const chd_kernel_desc* kd, const double* X, const double* Xq, int nq,
                const uint8_t* w, const double* yb, int T, double* out, void* workspace, size_t workspace_bytes,
                void* stream);

/* Building blocks exposed for parity tests / profiling (same kernels the calls above launch). */
/* Householder tridiagonalisation of [0 ga^T; ga K]: d (T x n), e (T x n; e[n-1] unused), beta0 (T);
 * B (T x n x ldb, may be NULL, nb_cols columns) is overwritten with Q^T B. Reflectors stay in K. */
int chd_tridiag(const chd_plan* plan, double* K, int ldk, const double* ga, int T, double* B, int ldb, int nb_cols,
                double* d, double* e, double* beta0, void* workspace, size_t workspace_bytes, void* stream);
/* Two-stage reduction building blocks (inspection / test entry points -- chd_regress and
 * chd_prune_chain run the same kernels when the plan is in two-stage mode).  Replace, together, the eigh of
 * interpolatory.py:48.  CHD_BAND = 32 is the half-width of the band form; band arrays are T x n x CHD_BAND_LD with
 * band[t][i][d] = T_b(i, i-d), d = 0..CHD_BAND.
 *   chd_band_reduce         K (T x n x ldk, lower triangle read, destroyed: holds the block reflectors afterwards),
 *                           ga (T x n), B (T x n x ldb or NULL: B <- Q^T B)  ->  band, beta0 (T): Q^T ga = beta0 e_0
 *   chd_band_backtransform  out (T x n) = -beta0 Q x, with K / workspace as left by chd_band_reduce
 *   chd_band_chase          band -> tridiagonal d, e (T x n each, e[n-1] = 0), values only (bulge chasing)
 *   chd_band_solve          x = (T_b + gamma_used I)^-1 e_0, noise, and the Z-test bounds from Bq = Q^T S */
#define CHD_BAND 32
#define CHD_BAND_LD 34
size_t chd_band_workspace_bytes(const chd_plan* plan, int T);
int chd_band_reduce(const chd_plan* plan, double* K, int ldk, const double* ga, int T, double* B, int ldb, int nb_cols,
                    double* band, double* beta0, void* workspace, size_t workspace_bytes, void* stream);
int chd_band_backtransform(const chd_plan* plan, const double* K, int ldk, const double* x, const double* beta0,
                           double* out, int T, void* workspace, size_t workspace_bytes, void* stream);
int chd_band_chase(const chd_plan* plan, const double* band, int T, double* d, double* e, void* workspace,
                   size_t workspace_bytes, void* stream);
int chd_band_solve(const chd_plan* plan, const double* band, const double* gamma_used, double* Bq, int ldb, int nz,
                   int T, double* x, double* noise, double* z_low, double* z_high, void* workspace,
                   size_t workspace_bytes, void* stream);
/* Selects the regression path of chd_regress / chd_prune_chain: 1 = two-stage (default), 0 = one-stage
 * Householder tridiagonalisation (chd_tridiag).  Workspace sizes depend on it: call before chd_workspace_bytes. */
int chd_plan_set_two_stage(chd_plan* plan, int on);
/* Optional timing of the dominant kernel: when enabled, every tridiagonalisation launch is bracketed by
 * cudaEvents on its own stream; chd_profile_read synchronises those events and returns the summed duration
 * (ms) and the number of launches since the last chd_profile_enable(1). */
int chd_profile_enable(int on);
int chd_profile_read(double* total_ms, int* launches);
/* Per-kernel variant: category 0 = one-stage tridiagonalisation, 1 = symm (Y = A22 V), 2 = syr2k (trailing update),
 * 3 = panel QR, 4 = bulge chasing, 5 = W formation, 6 = band solve + Z-test, 7 = back-transformation; -1 = all. */
int chd_profile_read_cat(int category, double* total_ms, int* launches);
/* 1 if the plan runs the two-stage reduction (default; env CHD_TWO_STAGE=0 or chd_plan_set_two_stage to change). */
int chd_plan_two_stage(const chd_plan* plan);
/* FP64 tensor-core (DMMA) throughput probe: runs `iters` back-to-back m8n8k4 chains per warp on
 * the whole GPU and returns the measured TFLOP/s (host-synchronising; used for the roofline peak). */
int chd_dmma_peak(int iters, double* tflops);
/* Same probe at a GEMM kernel's occupancy: one CTA of `warps_per_sm` warps per SM, `chains` (2, 4 or 8) independent
 * accumulators per warp -- the DMMA rate a kernel with that many resident warps and that warp tile can reach at best. */
int chd_dmma_occupancy_probe(int iters, int warps_per_sm, int chains, double* tflops);
/* FP64 pipe probe: mode 0 = DMMA only, 1 = DFMA only, 2 = both interleaved at equal flops; returns TFLOP/s. */
int chd_fp64_probe(int iters, int mode, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CHD_B200_H */
