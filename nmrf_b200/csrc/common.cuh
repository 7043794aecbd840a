// Shared helpers for libnmrf_b200 (sm_100a).  No torch, no allocation, no host sync.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <atomic>
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

// Per-device launch configuration.  cudaFuncAttributeMaxDynamicSharedMemorySize and the SM count are properties of a
// (kernel, device) pair: a process that drives several GPUs must configure each of them (one static flag per kernel would
// leave every device but the first with the 48 KB default and fail the launch).  Races are benign (idempotent calls).
constexpr int kMaxDevices = 64;
struct PerDevice { std::atomic<int> v[kMaxDevices]; };      // zero-initialised statics
inline int current_device() { int d = 0; cudaGetDevice(&d); return d; }
int num_sms();                                                // SM count of the current device (api.cu)
template <typename Kernel>
inline void ensure_dynamic_smem(Kernel kernel, int bytes, PerDevice& done) {
  const int d = current_device();
  if (d >= 0 && d < kMaxDevices && done.v[d].load(std::memory_order_relaxed) >= bytes) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (d >= 0 && d < kMaxDevices) done.v[d].store(bytes, std::memory_order_relaxed);
}

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

// Same encoding from an EXTENDED-precision coordinate (label = hi + lo, see nmrf_prop_head_tail): the argument c * 2^i is
// formed and reduced in double, so the top frequencies (c * 2^14 ~ 2e4 rad, where one fp32 ulp of the label is ~1.5e-3 rad)
// are as accurate as the coordinate itself.  30 double sin/cos per token: noise next to the projections that consume them.
__device__ __forceinline__ void fourier32_ext(double coord, double normalizer, float* __restrict__ dst, int lane) {
  const double c = coord * normalizer;
  float v;
  if (lane < 15) v = (float)sin(c * (double)(1 << lane));
  else if (lane < 30) v = (float)cos(c * (double)(1 << (lane - 15)));
  else if (lane == 30) v = (float)c;
  else v = 0.f;
  dst[lane] = v;
}

}  // namespace nmrf
