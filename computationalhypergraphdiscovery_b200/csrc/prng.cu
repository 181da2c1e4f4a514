// JAX-compatible threefry2x32 PRNG on the device: random.split and random.normal (float64).
// Replaces jax.random.split / jax.random.normal as used at interpolatory.py:47,99 of the
// reference (jax==0.4.28, default non-partitionable threefry; algorithm restated in
// oracle/jaxprng.py and SURVEY.md App. B.1).
#include "common.cuh"

namespace chd {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += ks0;
  x1 += ks1;
#define CHD_TF_ROUND(r) \
  x0 += x1;             \
  x1 = rotl32(x1, r);   \
  x1 ^= x0;
  CHD_TF_ROUND(13) CHD_TF_ROUND(15) CHD_TF_ROUND(26) CHD_TF_ROUND(6)
  x0 += ks1; x1 += ks2 + 1u;
  CHD_TF_ROUND(17) CHD_TF_ROUND(29) CHD_TF_ROUND(16) CHD_TF_ROUND(24)
  x0 += ks2; x1 += ks0 + 2u;
  CHD_TF_ROUND(13) CHD_TF_ROUND(15) CHD_TF_ROUND(26) CHD_TF_ROUND(6)
  x0 += ks0; x1 += ks1 + 3u;
  CHD_TF_ROUND(17) CHD_TF_ROUND(29) CHD_TF_ROUND(16) CHD_TF_ROUND(24)
  x0 += ks1; x1 += ks2 + 4u;
  CHD_TF_ROUND(13) CHD_TF_ROUND(15) CHD_TF_ROUND(26) CHD_TF_ROUND(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef CHD_TF_ROUND
}

// split(key, 2): threefry over counts [0,1,2,3] -> lanes (0,2),(1,3); rows (o0[0],o0[1]),(o1[0],o1[1]).
__global__ void split_kernel(const uint32_t* __restrict__ keys, int T, uint32_t* __restrict__ new_keys,
                             uint32_t* __restrict__ subkeys) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint32_t k0 = keys[2 * t], k1 = keys[2 * t + 1];
  uint32_t a0 = 0, a1 = 2, b0 = 1, b1 = 3;
  threefry2x32(k0, k1, a0, a1);
  threefry2x32(k0, k1, b0, b1);
  // out = concat([a0, b0], [a1, b1]) reshaped (2, 2)
  new_keys[2 * t] = a0;
  new_keys[2 * t + 1] = b0;
  subkeys[2 * t] = a1;
  subkeys[2 * t + 1] = b1;
}

// normal(key, (rows, cols)) float64: element e = r*cols + c uses counter pair (e, size + e);
// bits = (o0 << 32) | o1; u = bitcast((bits >> 12) | 0x3FF0..) - 1; x = max(lo, u*(hi-lo)+lo);
// out = sqrt(2) * erfinv(x).
__global__ void normal_kernel(const uint32_t* __restrict__ keys, int rows, int cols, double* __restrict__ out,
                              int ld) {
  const int t = blockIdx.y;
  const uint32_t k0 = keys[2 * t], k1 = keys[2 * t + 1];
  const uint32_t size = (uint32_t)rows * (uint32_t)cols;
  const double lo = -0.99999999999999988897769753748;  // nextafter(-1, 0)
  const double span = __dsub_rn(1.0, lo);
  double* o = out + (size_t)t * rows * ld;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < size; e += gridDim.x * blockDim.x) {
    uint32_t x0 = e, x1 = size + e;
    threefry2x32(k0, k1, x0, x1);
    unsigned long long bits = ((unsigned long long)x0 << 32) | (unsigned long long)x1;
    double u = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ull)) - 1.0;
    double x = fmax(lo, __dadd_rn(__dmul_rn(u, span), lo));
    double v = __dmul_rn(1.4142135623730951, erfinv(x));
    uint32_t r = e / (uint32_t)cols, c = e - r * (uint32_t)cols;
    o[(size_t)r * ld + c] = v;
  }
}

}  // namespace chd

extern "C" int chd_split(const uint32_t* keys, int T, uint32_t* new_keys, uint32_t* subkeys, void* stream) {
  if (T <= 0) return CHD_OK;
  chd::split_kernel<<<(T + 127) / 128, 128, 0, (cudaStream_t)stream>>>(keys, T, new_keys, subkeys);
  CHD_LAUNCHED();
  return CHD_OK;
}

extern "C" int chd_normal(const uint32_t* keys, int T, int rows, int cols, double* out, int ld, void* stream) {
  if (T <= 0 || rows <= 0 || cols <= 0) return CHD_OK;
  if (ld < cols) return CHD_EINVAL;
  long long size = (long long)rows * cols;
  int blocks = (int)((size + 255) / 256);
  if (blocks > 1184) blocks = 1184;  // 8 x 148 SMs
  dim3 grid(blocks, T);
  chd::normal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(keys, rows, cols, out, ld);
  CHD_LAUNCHED();
  return CHD_OK;
}
