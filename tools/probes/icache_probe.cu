// Instruction-cache pressure test: warp 8 repeatedly runs a short sequence {try_wait on a completed mbarrier; __syncwarp; elect}
// spread over a few code lines, sleeping ~1 us between repetitions, while the other 17 warps stream through a large
// straight-line code body (BODY_KB of FMAs).  Reports cycles per step for warp 8.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nmrf_b200/csrc -DBODY_REPS=.. -o ... tools/probes/icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
namespace nmrf { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } void count_launch(int) {} }
using namespace nmrf::tc;
#define F1 asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b));
#define F8 F1 F1 F1 F1 F1 F1 F1 F1
#define F64 F8 F8 F8 F8 F8 F8 F8 F8
#define F512 F64 F64 F64 F64 F64 F64 F64 F64      // 512 instructions = 8 KB
template <int KB8>
__global__ void __launch_bounds__(576, 1) probe(int nbusy, long long* out, float a, float b) {
  __shared__ uint64_t done_bar;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&done_bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
  __syncthreads();
  if (warp == 8) {
    long long t_wait = 0, t_sync = 0, t_elect = 0;
    for (int it = 0; it < 100; ++it) {
      __nanosleep(1000);
      const long long t0 = clock64();
      if (lane == 0) { while (!mbar_try(smem_u32(&done_bar), 0)) {} }
      const long long t1 = clock64();
      __syncwarp();
      const long long t2 = clock64();
      const bool e = elect_one();
      const long long t3 = clock64();
      if (e) { t_wait += t1 - t0; t_sync += t2 - t1; t_elect += t3 - t2; }
      __syncwarp();
    }
    if (lane == 0) { stop = 1; if (blockIdx.x == 0) { out[0] = t_wait / 100; out[1] = t_sync / 100; out[2] = t_elect / 100; } }
  } else if (warp < nbusy) {
    float x = (float)tid;
    int guard = 0;
    while (!stop && guard++ < (1 << 16)) {
      if (KB8 >= 1) { F512 }
      if (KB8 >= 2) { F512 }
      if (KB8 >= 3) { F512 }
      if (KB8 >= 4) { F512 }
      if (KB8 >= 5) { F512 }
      if (KB8 >= 6) { F512 }
      if (KB8 >= 8) { F512 F512 }
      if (KB8 >= 12) { F512 F512 F512 F512 }
    }
    if (x == 12345.f) out[5] = 1;
  }
}
template <int KB8> void run(long long* out) {
  for (int nbusy : {0, 8, 18}) {
    probe<KB8><<<148, 576>>>(nbusy, out, 1.0001f, 0.5f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    printf("busy code %3d KB, busy warps %2d: try_wait %4lld  syncwarp %4lld  elect %4lld cycles  %s\n", KB8 * 8, nbusy, h[0], h[1], h[2],
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  run<1>(out); run<2>(out); run<3>(out); run<4>(out); run<6>(out); run<8>(out); run<12>(out);
  return 0;
}
