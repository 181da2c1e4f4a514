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

}  // namespace chd
