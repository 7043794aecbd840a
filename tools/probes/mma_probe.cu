// Throughput probe: mma.sync.m16n8k8 tf32 (legacy tensor path) on sm_100a, warps/SM swept.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 4);
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int iters = 20000;
    k<<<148, warps * 32>>>(out, 10);
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    cudaEventRecord(s); k<<<148, warps * 32>>>(out, iters); cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    const double mmas = 148.0 * warps * iters * 8;
    printf("warps/SM %2d: %.3f ms, %.2f TFLOP/s tf32 (m16n8k8), %.2f cycles/mma/SM at 1.9GHz\n", warps, ms,
           mmas * 2048 / ms / 1e9, ms * 1e-3 * 1.9e9 / (warps * iters * 8.0));
  }
  return 0;
}
