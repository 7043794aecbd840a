// Shared helpers for libnmrf_b200 (sm_100a).  No torch, no allocation, no host sync.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/nmrf_b200.h"

namespace nmrf {

constexpr int kEmbed = 128;     // NMP.PROP_EMBED_DIM / INFER_EMBED_DIM (default.py:44-45)
constexpr int kHeads = 4;       // NMP.*_N_HEADS (default.py:50-51)
constexpr int kHeadDim = 32;
constexpr int kQkv = 3 * kEmbed;
constexpr int kMaxK = 8;        // proposals per pixel supported by the small-K kernels

void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> NMRF_ERR_CUDA + message
void count_launch(int n = 1);

#define NMRF_REQUIRE(cond, ...)                       \
  do {                                                \
    if (!(cond)) {                                    \
      nmrf::set_error(__VA_ARGS__);                   \
      return NMRF_ERR_BAD_ARG;                        \
    }                                                 \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// fourier_coord_embed (NMP.py:35-51), N_freqs = 15, logscale: 15 sin, 15 cos, raw.
// Arguments reach ~3e4 rad: accurate sinf/cosf only (never __sinf).  Writes 32 floats
// (last one zero) so rows stay 16-byte aligned.
__device__ __forceinline__ void fourier32(float coord, float normalizer, float* __restrict__ dst, int lane) {
  // called by a full warp: lane i<15 -> sin, 15<=i<30 -> cos, 30 -> raw, 31 -> 0
  const float c = coord * normalizer;
  float v;
  if (lane < 15) v = sinf(c * exp2f((float)lane));
  else if (lane < 30) v = cosf(c * exp2f((float)(lane - 15)));
  else if (lane == 30) v = c;
  else v = 0.f;
  dst[lane] = v;
}

}  // namespace nmrf
