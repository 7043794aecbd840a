// Wake-up latency of an mbarrier wait that is NOT complete at the first poll: warp 1 waits (try_wait loop / test_wait loop /
// try_wait with a suspend-time hint), warp 0 arrives after a delay; both stamp clock64 (same SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nmrf_b200/csrc -o tools/probes/wake_probe tools/probes/wake_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
namespace nmrf { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } void count_launch(int) {} }
using namespace nmrf::tc;
__device__ __forceinline__ bool test_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool try_wait_hint(uint32_t addr, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity), "r"(ns) : "memory");
  return done != 0;
}
template <int MODE>
__global__ void probe(int delay_ns, long long* out) {
  __shared__ uint64_t bar;
  __shared__ long long t_arrive[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long long sum = 0, mx = 0;
  for (int it = 0; it < 64; ++it) {
    if (warp == 0 && lane == 0) {
      __nanosleep(delay_ns);
      t_arrive[it] = clock64();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t addr = smem_u32(&bar), par = it & 1;
        if (MODE == 0) { while (!mbar_try(addr, par)) {} }
        if (MODE == 1) { while (!test_wait(addr, par)) {} }
        if (MODE == 2) { while (!try_wait_hint(addr, par, 20)) {} }
        const long long t = clock64();
        const long long d = t - *(volatile long long*)&t_arrive[it];
        sum += d; mx = d > mx ? d : mx;
      }
      __syncwarp();
    }
    __syncthreads();
  }
  if (warp == 1 && lane == 0 && blockIdx.x == 0) { out[0] = sum / 64; out[1] = mx; }
}
template <int MODE> void run(const char* name, long long* out) {
  for (int delay : {0, 500, 2000, 10000}) {
    probe<MODE><<<148, 64>>>(delay, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-28s arrive after %5d ns: observed %5lld cycles after the arrive (max %5lld)  %s\n", name, delay, h[0], h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  run<0>("try_wait loop", out);
  run<1>("test_wait loop", out);
  run<2>("try_wait, 20 ns hint", out);
  return 0;
}
