// How long does the issuing thread spend in tcgen05.commit?  (elected lane, sm_100a)
//   mode 0: loop of {commit}                     mode 1: loop of {12 x MMA 128x128x8 TS, commit}
//   mode 2: loop of {12 x MMA, commit, commit}   mode 3: loop of {12 x MMA} with one commit at the end
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nmrf_b200/csrc -o tools/probes/commit_probe tools/probes/commit_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
namespace nmrf { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } void count_launch(int) {} }
using namespace nmrf::tc;
template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<float4*>(base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128);
      const uint64_t dB = make_desc(smem_u32(base));
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (MODE >= 1) {
#pragma unroll
          for (int k = 0; k < 12; ++k) umma_tf32_ta(tmem, tmem + 256 + (k & 3) * 8, dB + (uint64_t)((k & 3) * 2), idesc, 1u);
        }
        if (MODE <= 2) umma_commit(&bar[it & 3]);
        if (MODE == 2) umma_commit(&bar[(it + 1) & 3]);
      }
      t1 = clock64();
      umma_commit(&bar[0]);
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncwarp();
  }
  // give the async arrivals time to land before teardown
  for (int i = 0; i < 2000; ++i) __nanosleep(100);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int MODE> void run(const char* name, long long* out) {
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 100;
  probe<MODE><<<148, 128, 80 * 1024>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%-40s %.1f cycles per iteration (%s)\n", name, (double)h / iters, cudaGetErrorString(e));
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  run<0>("commit only", out);
  run<1>("12 MMA (768 cyc of work) + commit", out);
  run<2>("12 MMA + 2 commits", out);
  run<3>("12 MMA, no commit", out);
  return 0;
}
