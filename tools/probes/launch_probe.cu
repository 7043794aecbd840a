// Fixed per-launch cost of a 148-CTA x 576-thread kernel with ~200 KB dynamic shared memory, back to back in a stream and in a
// CUDA graph: (0) empty, (1) + TMEM alloc/dealloc of 512 columns + mbarrier init + __syncthreads, (2) same with PDL
// (programmatic stream serialization + griddepcontrol).   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/launch_probe tools/probes/launch_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(576, 1) k(float* out) {
  extern __shared__ uint8_t dsm[];
  __shared__ uint32_t tmem_base;
  __shared__ uint64_t bars[16];
  if (MODE == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (MODE >= 1) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[i])), "r"(1));
    __syncthreads();
    if (MODE == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) out[blockIdx.x] = (float)tmem_base + dsm[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  } else if (threadIdx.x == 0 && out == nullptr) out[0] = dsm[0];
}
template <int MODE>
void run(float* out, int dyn, bool pdl, cudaStream_t st, int n) {
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = dyn; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k<MODE>, out);
  }
}
template <int MODE>
void bench(const char* name, float* out, int dyn, bool pdl) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  const int n = 200;
  run<MODE>(out, dyn, pdl, st, 20); cudaStreamSynchronize(st);
  cudaEventRecord(s, st); run<MODE>(out, dyn, pdl, st, n); cudaEventRecord(e, st); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal); run<MODE>(out, dyn, pdl, st, n); cudaStreamEndCapture(st, &g);
  cudaError_t err = cudaGraphInstantiate(&ge, g, 0);
  float msg = -1.f;
  if (err == cudaSuccess) {
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEventRecord(s, st); cudaGraphLaunch(ge, st); cudaEventRecord(e, st); cudaEventSynchronize(e);
    cudaEventElapsedTime(&msg, s, e);
  }
  printf("%-44s smem %3d KB: stream %.2f us/launch, graph %.2f us/launch (%s)\n", name, dyn >> 10, ms * 1e3 / n, msg * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float* out; cudaMalloc(&out, 4096);
  for (int dyn : {16 << 10, 200 << 10}) {
    bench<0>("empty", out, dyn, false);
    bench<1>("tmem alloc + barriers", out, dyn, false);
    bench<1>("tmem alloc + barriers, PDL attr only", out, dyn, true);
    bench<2>("tmem alloc + barriers, PDL + griddepcontrol", out, dyn, true);
  }
  return 0;
}
