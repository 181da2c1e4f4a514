// Internal cross-translation-unit declarations of libchd_b200.
#pragma once
#include "common.cuh"

namespace chd {

struct TridiagBuffers {
  double *tau, *v0, *u, *rsum, *Vp, *Wp, *colpart, *part, *pgram;
  unsigned* bar;
};

size_t tridiag_smem_bytes(const Plan* pl);
size_t tridiag_workspace(const Plan* pl, int T);
bool tridiag_carve(const Plan* pl, int T, Carver& cv, TridiagBuffers& tb);
int tridiag_launch(const Plan* pl, double* K, int ldk, const double* ga, int T, double* B, int ldb, int nbc,
                   double* d, double* e, double* beta0, const TridiagBuffers& tb, cudaStream_t st);
int backtransform_launch(const Plan* pl, const double* K, int ldk, const TridiagBuffers& tb, const double* x,
                         const double* beta0, double* out, size_t out_stride, int T, cudaStream_t st);
int gamma_launch(const Plan* pl, const double* d, const double* e, const double* beta0, const double* grid,
                 double* loss, double* B, int ldb, int nz, int gamma_min_mode, double gamma_min_base,
                 double* gamma_min, int want_lambda_min, double* lambda_min, double* eta, double* uu, double* x,
                 double* noise, double* z_low, double* z_high, double* gamma, size_t out_stride, int32_t* status,
                 int T, cudaStream_t st, double* gu_out = nullptr);
// ---- two-stage reduction (band.cu, chase.cu, bandsolve.cu)
struct BandBuffers {
  double *V, *Y, *W, *Tf, *Mpart, *Zpart, *Gpart, *Mred, *Zred, *Gred, *TM, *TZ, *v0, *tau0, *bandS, *bandW;
  size_t tf_stride;
};
int band_npanels(int n);
size_t band_workspace(const Plan* pl, int T);
bool band_carve(const Plan* pl, int T, Carver& cv, BandBuffers& bb);
int band_reduce_launch(const Plan* pl, double* K, int ldk, const double* ga, int T, double* B, int ldb, int nz,
                       double* beta0, const BandBuffers& bb, cudaStream_t st);
int backtransform2_launch(const Plan* pl, const double* K, int ldk, const BandBuffers& bb, const double* x,
                          const double* beta0, double* out, size_t out_stride, int T, cudaStream_t st);
size_t chase_workspace(const Plan* pl, int T);
size_t chase_prog_ints(const Plan* pl, int T);
// per-device set-up (function attributes, occupancy queries), called by chd_plan_create on the plan's device
int chase_device_init(Plan* pl);
int band_device_init(Plan* pl);
int band_tma_device_init(Plan* pl);
// TMA-fed trailing update (band_tma.cu): A[r0:, r0:] -= sum_q V_q W_q^T + W_q V_q^T over the G panels of a group
int syr2k_tma_launch(const Plan* pl, double* K, int ldk, int T, int r0, int G, const double* V, const double* W,
                     cudaStream_t st);
int tridiag_device_init(Plan* pl);
int act_device_init(Plan* pl);
int chase_launch(const Plan* pl, double* bandW, int* prog, double* d, double* e, int T, cudaStream_t st);
int band_solve_launch(const Plan* pl, const double* bandS, double* Lg, const double* gu, double* Bq, int ldb, int nz,
                      double* x, double* noise, double* z_low, double* z_high, size_t out_stride, int T,
                      cudaStream_t st);
int gram_workspace(const Plan* pl, int T, int nq, size_t* bytes);
int gram_launch(const Plan* pl, const chd_kernel_desc* kd, const double* X, const double* Xq, int nq,
                const uint8_t* w, int T, double* K, int ldk, void* ws, size_t ws_bytes, cudaStream_t st,
                double* P, int pmode, const uint8_t* wrem, const double* Gfull = nullptr);
// shared linear core of the full variable set (downdate mode of the Gram, gram.cu)
int gram_full_core_launch(const Plan* pl, const double* X, double* G, int ldk, uint8_t* ones, void* ws,
                          size_t ws_bytes, cudaStream_t st);
size_t gram_full_core_bytes(const Plan* pl);
size_t act_workspace(const Plan* pl, int T);
int act_launch(const Plan* pl, const chd_kernel_desc* kd, const double* X, const int32_t* cluster_of,
               const uint8_t* w0, uint8_t* alive, const double* yb, const double* ga, int T, double* act,
               size_t act_stride, int prune, int32_t* pruned, size_t pruned_stride, void* ws, size_t ws_bytes,
               cudaStream_t st, const double* P, int ldk);
int removed_vars_launch(const uint8_t* w_prev, const uint8_t* w_new, int count, uint8_t* wrem, cudaStream_t st);
enum ProfCat { PROF_TRIDIAG = 0, PROF_SYMM = 1, PROF_SYR2K = 2, PROF_QR = 3, PROF_CHASE = 4, PROF_WFORM = 5,
               PROF_SOLVE = 6, PROF_BACK = 7 };
int profile_enable(int on);
bool profile_on();
int profile_mark_begin(int cat, cudaStream_t st);
int profile_mark_end(int cat, cudaStream_t st);
int profile_read(double* total_ms, int* launches);
int profile_read_cat(int cat, double* total_ms, int* launches);
int active_vars_launch(const uint8_t* w0, const uint8_t* alive, const int32_t* cluster_of, int N, int C, int T,
                       uint8_t* w, cudaStream_t st);

}  // namespace chd
