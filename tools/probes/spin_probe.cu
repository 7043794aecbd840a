// Does a set of warps spinning on mbarrier.try_wait slow down another warp's try_wait / __syncwarp / elect.sync?
// 576-thread CTA per SM; warp 8 runs {try_wait on a COMPLETED barrier; __syncwarp; elect} 200 times and reports cycles per step;
// NSPIN other warps (one lane each polling, 31 lanes parked at __syncwarp) spin on a barrier that never completes, with
// optional __nanosleep back-off.    build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nmrf_b200/csrc -o tools/probes/spin_probe tools/probes/spin_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
namespace nmrf { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } void count_launch(int) {} }
using namespace nmrf::tc;
__global__ void __launch_bounds__(576, 1) probe(int nspin, int sleep_ns, int all_lanes, long long* out) {
  __shared__ uint64_t done_bar, never_bar;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&done_bar, 1); mbar_init(&never_bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");   // phase 0 complete
  __syncthreads();
  if (warp == 8) {
    long long t_wait = 0, t_sync = 0, t_elect = 0;
    for (int it = 0; it < 200; ++it) {
      const long long a = clock64();
      if (lane == 0) { while (!mbar_try(smem_u32(&done_bar), 0)) {} }
      const long long b = clock64();
      __syncwarp();
      const long long c = clock64();
      const bool e = elect_one();
      const long long d = clock64();
      if (e) { t_wait += b - a; t_sync += c - b; t_elect += d - c; }
      __syncwarp();
    }
    if (lane == 0) { stop = 1; if (blockIdx.x == 0) { out[0] = t_wait / 200; out[1] = t_sync / 200; out[2] = t_elect / 200; } }
  } else if (warp < nspin + (warp > 8 ? 1 : 0) && warp != 8) {
    if (lane == 0 || all_lanes) {
      int guard = 0;
      while (!stop && guard++ < (1 << 20)) {
        if (mbar_try(smem_u32(&never_bar), 0)) break;
        if (sleep_ns) __nanosleep(sleep_ns);
      }
    }
    __syncwarp();
  }
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  for (int all = 0; all < 2; ++all)
    for (int sleep_ns : {0, 200})
      for (int nspin : {0, 4, 8, 16}) {
        probe<<<148, 576>>>(nspin, sleep_ns, all, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[3]; cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
        printf("spinners %2d (%s, sleep %3d ns): try_wait %4lld  syncwarp %4lld  elect %4lld cycles   %s\n", nspin, all ? "32 lanes" : "1 lane  ", sleep_ns, h[0], h[1], h[2],
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
