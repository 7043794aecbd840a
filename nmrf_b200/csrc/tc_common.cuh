// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels (sm_100a inline PTX).
#pragma once
#include "common.cuh"

namespace nmrf {
namespace tc {

constexpr uint32_t SPIN_LIMIT = 1u << 26;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp (elect.sync).  tcgen05.mma / tcgen05.commit / cp.async.bulk are uniform-datapath
// instructions: under `if (lane == 0)` ptxas cannot prove a single active lane and wraps each of them in an
// ELECT ... BRA.U.ANY loop, which caps the issue rate at ~93 cycles per MMA (tools/probes/umma_probe.cu: a 128x128x8
// tf32 MMA then takes 93 cycles instead of its 64-cycle floor, a 128x32x8 one 93 instead of 16).  With the elect.sync
// predicate the instructions are emitted back to back and run at the hardware floor.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// round-to-nearest (ties away) to TF32 with two integer ops.  cvt.rna.tf32.f32 compiles to 4 ALU instructions (it also
// preserves inf/nan, which never occur here); the hi/lo split runs for every A element of every GEMM, so this matters.
__device__ __forceinline__ float rna_tf32_fast(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// Error-compensated 3xTF32 operand split: x = hi + lo, hi = rn_tf32(x), lo = rn_tf32(x - hi).  x - hi is exact in fp32; the
// tensor core reads only the upper 19 bits of an operand (it TRUNCATES), so an unrounded lo loses up to 2^-10 |lo| ~ 2^-21 |x|
// per product -- four times the 2^-23 |x| of a rounded lo and enough to raise the rate of flipped argmax / median decisions
// downstream (truncation split EPE 4.2e-3 px vs 7.6e-6 px for this one).
// Because of that truncation the rounding needs only its ADD: bits + 0x1000 carries into bit 13 exactly when the discarded
// low bits are >= half an ulp, and the tensor core drops whatever remains below bit 13 itself (bit-identical results with
// and without the mask: tools/dump_disp.py A/B; -DNMRF_LO_MASK builds the masked form).  One integer op per element.
__device__ __forceinline__ float lo_tf32(float x, float hi) {
#if defined(NMRF_LO_TRUNC)
  return x - hi;
#elif defined(NMRF_LO_MASK)
  return rna_tf32_fast(x - hi);
#else
  return __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
#endif
}

// cycle-stamp tracing of CTA 0 (tools/gemm_trace.py, tools/mlp_trace.py): compiled in only with -DNMRF_TRACE (make TRACE=1)
#ifdef NMRF_TRACE
#define NMRF_TRACE_STAMP(tp, idx) do { if ((tp) && (idx) < 4096) (tp)[(idx)] = clock64(); } while (0)
#else
#define NMRF_TRACE_STAMP(tp, idx) do { } while (0)
#endif

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (uint32_t spin = 0; spin < SPIN_LIMIT; ++spin)
    if (mbar_try(addr, parity)) return;
  __trap();
}
// Whole-warp wait (every lane must call it, converged): ALL lanes poll.  An earlier version let one lane poll and parked the
// other 31 at __syncwarp to keep the spin out of the MIO queue; measured inside the GEMM kernels (tools/mlp_trace.py), the
// reconvergence of that lane-divergent wait and the elect.sync / warp-collective instruction after it cost 600-750 cycles
// EACH per use (10 cycles in isolation), i.e. ~1400 cycles per weight unit on the MMA warp's critical path.  A uniform
// poll keeps the warp converged; try_wait suspends in hardware between polls, so the instruction rate stays moderate.
// `sleep_ns` > 0 backs off between polls: for waits with slack (an epilogue waiting for a whole tile of MMAs).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, uint32_t sleep_ns = 0) {
  const uint32_t addr = smem_u32(bar);
  bool ok = false;
#pragma unroll 1
  for (uint32_t spin = 0; spin < SPIN_LIMIT && !ok; ++spin) {
    ok = mbar_try(addr, parity);
    if (!ok && sleep_ns) __nanosleep(sleep_ns);
  }
  if (!ok) __trap();
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (=1, unused for swizzled K-major), [32,46) SBO >> 4
//   (= 1024 B between 8-row groups), [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format TF32 (=2 at bits 7,10),
// both K-major, N>>3 at [17,23), M>>4 at [24,29); M is always 128 here
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand in tensor memory (lane = row, one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (warp-collective: lanes 32*(warp%4)..+32)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  return v;
}

// GELU with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7 in exact arithmetic; measured in fp32: GELU abs error
// 4.7e-7, the same as 0.5x(1+erff(x/sqrt2)) evaluated in fp32): ~14 instructions instead of erff's ~35, which made the
// fc1 epilogue the longest phase of the MLP GEMMs.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-p, e, 1.f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float act_fast(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return gelu_fast(v);
  return v;
}

// byte offset of the 16-byte chunk (row r, chunk c of 8) inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }


}  // namespace tc
}  // namespace nmrf
