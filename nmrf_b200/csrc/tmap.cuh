// Host-side helper: 2-D TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point -- the library does
// not link libcuda) over row-major fp32 matrices, 128-byte hardware swizzle = tc::swz(), out-of-bounds elements read as zero.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

namespace nmrf {

// matrix [rows, cols] fp32, `row_stride` floats between rows; box [box_rows, box_cols] with box_cols * 4 == 128 bytes
inline bool encode_tmap_2d(CUtensorMap* map, const float* base, long long rows, int cols, long long row_stride, int box_rows, int box_cols) {
  static const PFN_cuTensorMapEncodeTiled encode = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
  }();
  if (!encode || rows <= 0 || cols <= 0 || box_rows > 256 || box_cols * 4 != 128) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride * 4) % 16 != 0) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)row_stride * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace nmrf
