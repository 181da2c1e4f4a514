// Shared device/host helpers for libchd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/chd_b200.h"

namespace chd {

extern std::atomic<unsigned long long> g_launches;
void set_error(const char* fmt, ...);

#define CHD_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      chd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CHD_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

#define CHD_LAUNCHED()                          \
  do {                                          \
    chd::g_launches.fetch_add(1);               \
    CHD_CUDA(cudaGetLastError());               \
  } while (0)

#define CHD_CHASE_LD 66    /* row stride of the bulge-chasing work band (2 CHD_BAND sub-diagonals) */

struct Plan {
  int n, N, C, device;
  int sm_count;
  int G;          // CTAs cooperating on one matrix in the tridiagonalisation
  int wave;       // matrices per cooperative launch (G * wave <= sm_count)
  int nb;         // panel width
  int threads;    // threads per CTA of the tridiagonalisation (256: two CTAs per SM, 512: one)
  int bulk;       // symv data path: 1 = cp.async.bulk tile ring, 0 = register loads
  int rows_per_chunk;
  int two_stage;  // 1: dense -> band -> tridiagonal (band.cu / chase.cu / bandsolve.cu), 0: one-stage tridiag.cu
  size_t smem_optin;
  // per-device results of the *_device_init calls of chd_plan_create (function attributes are per device)
  int chase_variant;       // 1: CTA per sweep (default), 0: warp per sweep
  int chase_ctas_per_sm;   // resident bulge-chasing CTAs per SM
  int chase_wave;          // cap on matrices per bulge-chasing launch (0: none)
  size_t chase_l2_budget;  // bytes of work bands one bulge-chasing launch may touch
  int qr_reg_max_cs;       // largest cluster of the register-resident panel QR (0: disabled)
  int qr_two;              // two-rows-per-thread panel QR: 0 only when the cluster is too small, 1 auto, 2 always
  int small_dmma;          // 1: per-panel small products (V^T Y, V^T B, W, Z block) on the FP64 tensor cores
  int gram_downdate;       // 1: Gram cores as rank-r downdates of the shared G = X X^T when r < p
  int gauss_shared;        // 1: Gaussian activations of a batch share the per-pair factors (gauss_shared_kernel)
  int band_kd;             // panels per deferred trailing update of the dense -> band stage (1 or 2)
  int symm_tma_ctas;       // CTAs per SM of the TMA symm kernel (2: two 2-stage rings, 1: one 4-stage ring)
  int symm_tma;            // 1: TMA + mbarrier warp-specialised Y = A22 V with the paired block schedule
  int syr2k_tma_ctas;      // CTAs per SM of the TMA trailing update (2: 64-row blocks, 1: 128-row blocks)
  int syr2k_tma;           // 1: TMA + mbarrier warp-specialised trailing update (band_tma.cu), 0: cp.async kernel
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves a caller-provided workspace.
struct Carver {
  char* base;
  size_t off, cap;
  bool ok;
  Carver(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes), ok(true) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base + off);
    off += count * sizeof(T);
    if (off > cap) ok = false;
    return r;
  }
};

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ldcg2(const double2* p) { return __ldcg(p); }
__device__ __forceinline__ void stcg(double* p, double v) { __stcg(p, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; result valid in every thread. `scratch` holds >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? scratch[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// FP64 tensor-core MMA, m8n8k4 (SASS: DMMA.8x8x4).  a: A[row=lane/4][k=lane%4],
// b: B[k=lane%4][col=lane/4], c: C[row=lane/4][col=2*(lane%4)+{0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- exp(x) for x <= 0 (the RBF factors of the Gaussian mode): x = (64 q + j) ln2/64 + r, |r| <= ln2/128,
// exp(x) = 2^q * 2^(j/64) * (1 + r + ... + r^5/120).  ~17 instructions instead of libdevice's ~40; max error
// < 2 ulp (the truncation term r^6/720 is below 4e-17).  `tab` = 64 correctly rounded values of 2^(j/64) in shared
// memory (exp_table_load).  Arguments below -700 are clamped (result 1e-304 instead of 0).
static __device__ const double g_exp2_tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

__device__ __forceinline__ void exp_table_load(double* tab) {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) tab[i] = g_exp2_tab[i];
}

__device__ __forceinline__ double exp_nonpos(double x, const double* __restrict__ tab) {
  x = fmax(x, -700.0);      // exp(-700) = 1e-304 stands in for everything smaller (added to 1 by every caller)
  const double t = fma(x, 0x1.71547652b82fep+6, 0x1.8p+52);     // x * 64/ln2, rounded to an integer in the low bits
  const int k = __double2loint(t);
  const double kd = t - 0x1.8p+52;
  double r = fma(kd, -0x1.62e42fee00000p-7, x);                 // ln2/64 in two parts (hi has 21 trailing zero bits)
  r = fma(kd, -0x1.a39ef35793c76p-39, r);
  double p = fma(r, 0x1.1111111111111p-7, 0x1.5555555555555p-5);
  p = fma(p, r, 0x1.5555555555555p-3);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int q = k >> 6;
  const double s = tab[k & 63] * p;
  return __hiloint2double(__double2hiint(s) + (q << 20), __double2loint(s));   // q >= -1010: stays normal
}

// 1/f for f >= 1 (finite): hardware reciprocal seed (rcp.approx.ftz.f64, ~20 bits) + two Newton steps (-> 2^-80):
// 5 instructions instead of the ~12 of the IEEE-rounded __drcp_rn; error < 1 ulp.
__device__ __forceinline__ double rcp_ge1(double f) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(f));
  double e = fma(-f, r, 1.0);
  r = fma(r, e, r);
  e = fma(-f, r, 1.0);
  return fma(r, e, r);
}

}  // namespace chd
