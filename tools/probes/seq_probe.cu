// Cost of {mbarrier.try_wait on a completed barrier; __syncwarp; elect.sync} when it follows tcgen05 work in the same warp.
// modes: 0 nothing before; 1 a tcgen05.commit before; 2 12 MMAs + commit before; 3 as 2 plus tcgen05.fence::after_thread_sync
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nmrf_b200/csrc -o tools/probes/seq_probe tools/probes/seq_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
namespace nmrf { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } void count_launch(int) {} }
using namespace nmrf::tc;
template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ uint64_t bar[4], done_bar;
  __shared__ uint32_t tmem_base;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<float4*>(base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    long long t_issue = 0, t_wait = 0, t_sync = 0, t_elect = 0;
    const uint32_t idesc = make_idesc(128);
    const uint64_t dB = make_desc(smem_u32(base));
    for (int it = 0; it < 100; ++it) {
      const long long ta = clock64();
      if (elect_one()) {
        if (MODE >= 3) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (MODE >= 2) {
#pragma unroll
          for (int k = 0; k < 12; ++k) umma_tf32_ta(tmem, tmem + 256 + (k & 3) * 8, dB + (uint64_t)((k & 3) * 2), idesc, 1u);
        }
        if (MODE >= 1) umma_commit(&bar[it & 3]);
      }
      __syncwarp();
      const long long t0 = clock64();
      if (lane == 0) { while (!mbar_try(smem_u32(&done_bar), 0)) {} }
      const long long t1 = clock64();
      __syncwarp();
      const long long t2 = clock64();
      const bool e = elect_one();
      const long long t3 = clock64();
      if (e) { t_issue += t0 - ta; t_wait += t1 - t0; t_sync += t2 - t1; t_elect += t3 - t2; }
      __syncwarp();
    }
    if (lane == 0 && blockIdx.x == 0) { out[0] = t_issue / 100; out[1] = t_wait / 100; out[2] = t_sync / 100; out[3] = t_elect / 100; }
  }
  for (int i = 0; i < 1000; ++i) __nanosleep(100);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int MODE> void run(const char* name, long long* out) {
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  probe<MODE><<<148, 128, 80 * 1024>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  printf("%-34s issue-section %4lld | try_wait %4lld  syncwarp %4lld  elect %4lld   %s\n", name, h[0], h[1], h[2], h[3], e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  run<0>("nothing before", out);
  run<1>("commit before", out);
  run<2>("12 MMAs + commit before", out);
  run<3>("fence + 12 MMAs + commit before", out);
  return 0;
}
