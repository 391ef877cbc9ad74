// Host-side construction of TMA tensor maps.  cuTensorMapEncodeTiled is a driver API; it is resolved through
// cudaGetDriverEntryPoint at first use so that libspecgpu.so has no link-time dependency on libcuda.
#pragma once
#include "common.cuh"

#if defined(SPECGPU_EMULATE)
namespace specgpu {
struct TensorMap {   // placeholder: the emulation build has no TMA engine, kernels take their plain-store path
  unsigned char opaque[128];
};
}  // namespace specgpu
#else
#include <cuda.h>

namespace specgpu {

using TensorMap = CUtensorMap;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 tensor [d2][d1][d0] (d0 contiguous) with row pitch ld1 and plane pitch ld2 (elements); box = [box2][box1][box0] (box2 defaults to 1).
// swizzle_bytes in {0, 32, 64, 128}: shared-memory swizzle span (box0 * 4 bytes must not exceed it when non-zero).
// Returns false when the layout cannot be described (alignment) -- callers then keep their plain-store path.
inline bool make_tensor_map_f32_3d(TensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                                   uint64_t ld2, uint32_t box0, uint32_t box1, int swizzle_bytes, uint32_t box2 = 1) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld1 * 4) % 16 != 0 || (ld2 * 4) % 16 != 0) return false;
  if (box0 == 0 || box1 == 0 || box2 == 0 || box0 > 256 || box1 > 256 || box2 > 256 || (box0 * 4) % 16 != 0) return false;
  if (d0 == 0 || d1 == 0 || d2 == 0 || d0 >= (1ull << 32) || d1 >= (1ull << 32) || d2 >= (1ull << 32)) return false;
  if (ld1 * 4 >= (1ull << 40) || ld2 * 4 >= (1ull << 40)) return false;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {ld1 * 4, ld2 * 4};
  const cuuint32_t box[3] = {box0, box1, box2};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace specgpu
#endif
