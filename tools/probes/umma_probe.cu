// Issue/throughput probe for tcgen05.mma kind::tf32 on sm_100a (run under gpurun):
//   cycles per 128xNx8 MMA for A in TMEM (TS) / A in shared memory (SS), N = 64..256, alone and with other warps hammering
//   shared memory (LDS/STS) the way the GEMM's producers and epilogue do.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nmrf_b200/csrc -o tools/probes/umma_probe tools/probes/umma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

namespace nmrf { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } void count_launch(int) {} }
using namespace nmrf::tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
               : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFF));
  return pred != 0;
}
__device__ __forceinline__ uint32_t idesc_n(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// mode: 0 = TS (A in TMEM), 1 = SS.  noise: 0 none, 1 = warps 1..3 do LDS.128 loops, 2 = STS.128 loops, 3 = 12 extra warps LDS
template <int ELECT, int MODE>
__global__ void __launch_bounds__(512, 1) probe(int n, int iters, int noise, long long* out) {
  constexpr int mode = MODE;
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<float4*>(base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  __shared__ volatile int stop;
  if (tid == 0) stop = 0;
  __syncthreads();
  if (warp == 0 && (ELECT ? elect_one() : tid == 0)) {
    const uint32_t idesc = idesc_n(n);
    // B tiles: 3 slots x (hi, lo) of [256 rows x 32 k] = 32 KB each at base + slot*64KB... keep to 5 x 32 KB
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int slot = it % 2;
      const uint64_t dBh = make_desc(smem_u32(base + slot * 65536)), dBl = make_desc(smem_u32(base + slot * 65536 + 32768));
      const uint64_t dAh = make_desc(smem_u32(base + 131072)), dAl = make_desc(smem_u32(base + 131072 + 16384));
      const uint32_t tAh = tmem + 256 + (it & 1) * 64, tAl = tAh + 32;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);
        if (mode == 0) {
          umma_tf32_ta(tmem, tAl + ks * 8, dBh + adv, idesc, 1u);
          umma_tf32_ta(tmem, tAh + ks * 8, dBl + adv, idesc, 1u);
          umma_tf32_ta(tmem, tAh + ks * 8, dBh + adv, idesc, 1u);
        } else {
          umma_tf32(tmem, dAl + adv, dBh + adv, idesc, 1u);
          umma_tf32(tmem, dAh + adv, dBl + adv, idesc, 1u);
          umma_tf32(tmem, dAh + adv, dBh + adv, idesc, 1u);
        }
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    stop = 1;
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 0) {
  } else if (noise && warp >= 1 && (noise == 3 ? warp < 13 : warp < 4)) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* p = reinterpret_cast<float4*>(base + 131072 + 32768) + tid;     // scratch region not used as an operand: 24 KB
    int guard = 0;
    while (!stop && guard < (1 << 22)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (noise == 2) p[(j * 128) % 1024] = acc;
        else { const float4 v = p[(j * 128) % 1024]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      }
      ++guard;
    }
    if (acc.x == 12345.f) out[7] = guard;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  const int dyn = 200 * 1024;
  cudaFuncSetAttribute(probe<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  cudaFuncSetAttribute(probe<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  cudaFuncSetAttribute(probe<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  cudaFuncSetAttribute(probe<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  const int iters = 200;
  for (int elect = 0; elect < 2; ++elect)
  for (int noise = 0; noise < 4; noise += 3)
    for (int mode = 0; mode < 2; ++mode)
      for (int n : {64, 128, 256}) {
        if (n == 256 && mode == 0 && false) continue;
        if (elect == 0 && mode == 0) probe<0, 0><<<148, 512, dyn>>>(n, iters, noise, out);
        if (elect == 0 && mode == 1) probe<0, 1><<<148, 512, dyn>>>(n, iters, noise, out);
        if (elect == 1 && mode == 0) probe<1, 0><<<148, 512, dyn>>>(n, iters, noise, out);
        if (elect == 1 && mode == 1) probe<1, 1><<<148, 512, dyn>>>(n, iters, noise, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("elect %d noise %d %s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)%s\n", elect, noise, mode ? "SS" : "TS", n,
               (double)h[0] / (iters * 12), (double)h[1] / (iters * 12), n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
