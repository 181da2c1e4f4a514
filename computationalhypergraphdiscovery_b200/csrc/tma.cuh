// TMA (cp.async.bulk.tensor) + mbarrier helpers for the sm_100a GEMM kernels of the dense -> band stage.
// Tiles move global -> shared memory through tensor maps (SASS: UTMALDG) and are handed from the producer warp to the
// DMMA consumer warps through mbarrier full / empty pairs (SASS: SYNCS) -- no CTA-wide barrier in the main loop.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace chd {
namespace tma {

// ---- host: tensor maps.  cuTensorMapEncodeTiled is fetched through the runtime (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// FP64 tensor of T matrices of `rows` x `cols` elements (row stride `ld` elements), boxes of box_rows x box_cols
// elements, out-of-bounds elements read as zero.  box_cols = 16 (128 B rows) is written with SWIZZLE_128B, box_cols = 8
// (64 B rows) unswizzled.  Coordinates: {column, row, matrix}.
inline int make_map_f64(CUtensorMap* map, const double* base, uint64_t cols, uint64_t rows, uint64_t T, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols = 16) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return CHD_ECUDA;
  }
  const cuuint64_t dims[3] = {cols, rows, T};
  const cuuint64_t strides[2] = {ld * sizeof(double), rows * ld * sizeof(double)};   // dims 1, 2 (bytes)
  const cuuint32_t box[3] = {box_cols, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code %d (cols %llu rows %llu T %llu ld %llu)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)T, (unsigned long long)ld);
    return CHD_ECUDA;
  }
  return CHD_OK;
}

// ---- device
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// one box {col, row, matrix} -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void load_box(void* dst, const CUtensorMap* m, int col, int row, int mat, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(col), "r"(row), "r"(mat)
      : "memory");
}

// Byte offset of element (row, col) inside a stack of 16-column boxes written with SWIZZLE_128B: every row is 128 B,
// the 16-byte chunk index is XORed with (row mod 8) (the pattern repeats every 1024 B, so box stacks must start on
// 1024-byte boundaries).  col in [0, 16).
__host__ __device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t col) {
  return row * 128u + ((((col >> 1) ^ (row & 7u)) << 4) | ((col & 1u) << 3));
}

}  // namespace tma
}  // namespace chd
