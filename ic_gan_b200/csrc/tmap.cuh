// Host-side TMA tensor-map encoding (driver entry point looked up at run time, no -lcuda) and the MN-major UMMA shared-
// memory descriptor, shared by the batched GEMM (tc_gemm.cu) and the fused non-local block kernels (tc_attn.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace icgan {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled is a driver-API call and wants a context current on the calling thread.  A thread that has so
// far only hit PyTorch's caching allocator (an autograd worker whose first kernel is one of ours) has none: one runtime
// call binds the device's primary context to it.
inline void bind_primary_context() {
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

inline EncodeTiledFn encode_fn() {
  bind_primary_context();
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D bf16 map: dims (inner, mid, batch) with element strides (1, ld, batch_stride); 128-byte swizzle.
inline int make_map3(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t mid, uint64_t batch, uint64_t ld,
              uint64_t batch_stride, uint32_t box_inner, uint32_t box_mid) {
  EncodeTiledFn fn = encode_fn();
  ICGAN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {inner, mid, batch};
  cuuint64_t str[2] = {ld * 2, (batch > 1 ? batch_stride : ld * mid) * 2};
  cuuint32_t box[3] = {box_inner, box_mid, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, str, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ICGAN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for dims %llu x %llu x %llu ld %llu",
                static_cast<int>(r), (unsigned long long)inner, (unsigned long long)mid, (unsigned long long)batch,
                (unsigned long long)ld);
  return 0;
}

__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

}  // namespace icgan
