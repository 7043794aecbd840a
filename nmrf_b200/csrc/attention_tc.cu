// A6 on the tensor cores: cross-shaped stripe attention + LePE (NMP.py:429-505) with tcgen05.
//
// One CTA = 128 query tokens of one (stripe, head); 128 threads, thread r owns query row r (= TMEM lane r).
// Keys/values are streamed in chunks of 32 tokens with an online softmax (flash style):
//     S  = (s Q) K_c^T      UMMA 128x32x32, 3xTF32 (hi/lo RN split of both operands)      -> TMEM cols [0,32)
//     p  = exp(S - m)       registers (tcgen05.ld, one row per thread), mask NMP.py:203-208, running max / sum
//     O += P V_c            UMMA 128x32x32, 3xTF32, P staged through swizzled smem as the A operand, V_c^T as B
//                           -> TMEM cols [32,64), added into the register accumulator with the softmax rescale
// All operand tiles are K-major SWIZZLE_128B with 128-byte rows (32 fp32): Q [128 rows], K_c [32 keys],
// V_c^T [32 dims x 32 keys], P [128 rows x 32 keys].  80 KB of shared memory and 64 TMEM columns per CTA, so
// two CTAs share an SM and overlap each other's load / MMA / softmax phases.  Exactness: same 3xTF32 scheme as
// the GEMM; exp, max, sum and the LePE epilogue are fp32.
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

// softmax in base 2: Q is scaled by 32^-0.5 * log2(e) once, so p = 2^(S - m) (one MUFU.EX2, <= 2 ulp) == e^(s - m)
constexpr float kScaleLog2e = 0.17677669529663687f * 1.4426950408889634f;
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr int AT_THREADS = 128;
constexpr int AT_KC = 32;                         // keys per chunk
constexpr int AT_QTILE = 128 * 128;               // bytes of a [128 x 32 fp32] tile
constexpr int AT_KTILE = 32 * 128;                // bytes of a [32 x 32 fp32] tile
constexpr int AT_DYN = 2 * AT_QTILE + 4 * AT_KTILE + 2 * AT_QTILE + 1024;

struct AtSmem {
  uint64_t bar_s, bar_o;
  uint32_t tmem_base;
  float gv[3][32];          // LePE taps (prev, centre, next) of this head's 32 channels
};

__device__ __forceinline__ void split4(const float4 v, float4& h, float4& l) {
  h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
  l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
}

__global__ void __launch_bounds__(AT_THREADS)
stripe_attention_tc_kernel(const float* __restrict__ qkv, int B, int h, int w, int K,
                           const float* __restrict__ get_v0, const float* __restrict__ get_v1,
                           float* __restrict__ out) {
  const int head = blockIdx.z;
  const bool vertical = head < 2;
  const int nstripes = vertical ? B * w : B * h;
  const int sid = blockIdx.y;
  if (sid >= nstripes) return;
  const int L = vertical ? h : w;
  const int Lk = L * K;
  const int q0 = blockIdx.x * 128;
  if (q0 >= Lk) return;

  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ AtSmem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  uint8_t* sQh = base;
  uint8_t* sQl = sQh + AT_QTILE;
  uint8_t* sKh = sQl + AT_QTILE;
  uint8_t* sKl = sKh + AT_KTILE;
  uint8_t* sVh = sKl + AT_KTILE;
  uint8_t* sVl = sVh + AT_KTILE;
  uint8_t* sPh = sVl + AT_KTILE;
  uint8_t* sPl = sPh + AT_QTILE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = vertical ? sid / w : sid / h;
  const int fixed = vertical ? sid % w : sid % h;
  auto token_row = [&](int t) -> size_t {
    const int l = t / K, n = t - l * K;
    const int y = vertical ? l : fixed, x = vertical ? fixed : l;
    return ((size_t)(b * h + y) * w + x) * K + n;
  };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&sm.bar_s, 1);
    mbar_init(&sm.bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 96) {
    const float* gv = vertical ? get_v0 : get_v1;             // [64,1,3,3]
    const int tap = tid >> 5, ch = (head & 1) * 32 + (tid & 31);
    const int idx = vertical ? (tap == 0 ? 1 : tap == 1 ? 4 : 7) : (tap == 0 ? 3 : tap == 1 ? 4 : 5);
    sm.gv[tap][tid & 31] = gv[ch * 9 + idx];
  }

  // ---- Q: this thread's row, scaled, split, swizzled --------------------------------------------------
  const int ti = q0 + tid;
  const bool valid = ti < Lk;
  const size_t my_row = valid ? token_row(ti) : 0;
  {
    const float* src = qkv + my_row * kQkv + head * 32;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) v = *reinterpret_cast<const float4*>(src + c * 4);
      v.x *= kScaleLog2e; v.y *= kScaleLog2e; v.z *= kScaleLog2e; v.w *= kScaleLog2e;
      float4 hi, lo;
      split4(v, hi, lo);
      const uint32_t so = swz(tid, c);
      *reinterpret_cast<float4*>(sQh + so) = hi;
      *reinterpret_cast<float4*>(sQl + so) = lo;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  const uint32_t t_lane = ((uint32_t)(warp * 32)) << 16;      // this warp's TMEM lane quarter

  const int pix_lo = (ti / K) * K, pix_hi = pix_lo + K;        // tokens of this row's own pixel
  float m = -INFINITY, l = 0.f;
  float o[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) o[j] = 0.f;

  // K/V prefetch registers: items i = tid + 128 e  ->  key = i / 8, 16-byte chunk = i % 8
  float4 kr[2], vr[2];
  auto fetch_kv = [&](int c) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = tid + 128 * e;
      const int tj = c * AT_KC + (i >> 3);
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (tj < Lk) {
        const float* src = qkv + token_row(tj) * kQkv + head * 32 + (i & 7) * 4;
        kk = *reinterpret_cast<const float4*>(src + 128);
        vv = *reinterpret_cast<const float4*>(src + 256);
      }
      kr[e] = kk; vr[e] = vv;
    }
  };
  const int nchunks = (Lk + AT_KC - 1) / AT_KC;
  const uint32_t idesc = make_idesc(32);
  const uint64_t dQh = make_desc(smem_u32(sQh)), dQl = make_desc(smem_u32(sQl));
  const uint64_t dKh = make_desc(smem_u32(sKh)), dKl = make_desc(smem_u32(sKl));
  const uint64_t dVh = make_desc(smem_u32(sVh)), dVl = make_desc(smem_u32(sVl));
  const uint64_t dPh = make_desc(smem_u32(sPh)), dPl = make_desc(smem_u32(sPl));

  // shared-memory offsets of this thread's staging items (independent of the chunk)
  uint32_t k_off[2], v_off[2][4], p_off[8];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = tid + 128 * e, key = i >> 3, c8 = i & 7;
    k_off[e] = swz(key, c8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = c8 * 4 + j;
      v_off[e][j] = (uint32_t)(d * 128 + (((key >> 2) ^ (d & 7)) << 4) + (key & 3) * 4);
    }
  }
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) p_off[c8] = swz(tid, c8);

  fetch_kv(0);
  for (int c = 0; c < nchunks; ++c) {
    // ---- stage K_c (rows = keys) and V_c^T (rows = dims, columns = keys); previous chunk's MMAs are complete ----
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float4 hi, lo;
      split4(kr[e], hi, lo);
      *reinterpret_cast<float4*>(sKh + k_off[e]) = hi;
      *reinterpret_cast<float4*>(sKl + k_off[e]) = lo;
      split4(vr[e], hi, lo);
      const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<float*>(sVh + v_off[e][j]) = hv[j];
        *reinterpret_cast<float*>(sVl + v_off[e][j]) = lv[j];
      }
    }
    if (c + 1 < nchunks) fetch_kv(c + 1);                       // in flight during this chunk's MMAs and softmax
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);
        umma_tf32(tmem, dQl + adv, dKh + adv, idesc, ks > 0 ? 1u : 0u);
        umma_tf32(tmem, dQh + adv, dKl + adv, idesc, 1u);
        umma_tf32(tmem, dQh + adv, dKh + adv, idesc, 1u);
      }
      umma_commit(&sm.bar_s);
    }
    __syncwarp();
    mbar_wait(&sm.bar_s, c & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float s[32];
    tmem_ld32(tmem + t_lane, s);
    // ---- mask + online softmax on this thread's row ---------------------------------------------------------------
    // only chunks that reach past the stripe's end or touch this row's own pixel need the mask (NMP.py:203-208)
    const int t_lo = c * AT_KC;
    if (t_lo + AT_KC > Lk || (t_lo < pix_hi && t_lo + AT_KC > pix_lo)) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int tj = t_lo + j;
        const bool masked = (tj >= Lk) || (tj >= pix_lo && tj < pix_hi && tj != ti);
        s[j] = masked ? -INFINITY : s[j];
      }
    }
    float cmax = s[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, s[j]);
    const float mnew = fmaxf(m, cmax);
    const float msafe = (mnew == -INFINITY) ? 0.f : mnew;        // a fully masked prefix keeps p = 0, scale = 1
    const float scale = ex2(m - msafe);
    float psum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s[j] = ex2(s[j] - msafe);
      psum += s[j];
    }
    l = l * scale + psum;
    m = mnew;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float4 hi, lo;
      split4(make_float4(s[c8 * 4], s[c8 * 4 + 1], s[c8 * 4 + 2], s[c8 * 4 + 3]), hi, lo);
      *reinterpret_cast<float4*>(sPh + p_off[c8]) = hi;
      *reinterpret_cast<float4*>(sPl + p_off[c8]) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);
        umma_tf32(tmem + 32, dPl + adv, dVh + adv, idesc, ks > 0 ? 1u : 0u);
        umma_tf32(tmem + 32, dPh + adv, dVl + adv, idesc, 1u);
        umma_tf32(tmem + 32, dPh + adv, dVh + adv, idesc, 1u);
      }
      umma_commit(&sm.bar_o);
    }
    __syncwarp();
    mbar_wait(&sm.bar_o, c & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float pv[32];
    tmem_ld32(tmem + t_lane + 32, pv);
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = fmaf(o[j], scale, pv[j]);
  }

  // ---- epilogue: normalise + LePE (NMP.py:433-449), see attention.cu for the derivation ------------------------
  if (valid) {
    const float inv = 1.f / l;
    const int li = ti / K;
    float prev[32], next[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { prev[j] = 0.f; next[j] = 0.f; }
    for (int n = 0; n < K; ++n) {
      if (li > 0) {
        const float4* p = reinterpret_cast<const float4*>(qkv + token_row((li - 1) * K + n) * kQkv + 256 + head * 32);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) { const float4 v = p[c8]; prev[c8 * 4] += v.x; prev[c8 * 4 + 1] += v.y; prev[c8 * 4 + 2] += v.z; prev[c8 * 4 + 3] += v.w; }
      }
      if (li + 1 < L) {
        const float4* p = reinterpret_cast<const float4*>(qkv + token_row((li + 1) * K + n) * kQkv + 256 + head * 32);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) { const float4 v = p[c8]; next[c8 * 4] += v.x; next[c8 * 4 + 1] += v.y; next[c8 * 4 + 2] += v.z; next[c8 * 4 + 3] += v.w; }
      }
    }
    const float4* own = reinterpret_cast<const float4*>(qkv + my_row * kQkv + 256 + head * 32);
    float4* dst = reinterpret_cast<float4*>(out + my_row * kEmbed + head * 32);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      const float4 v = own[c8];
      const float vv[4] = {v.x, v.y, v.z, v.w};
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = c8 * 4 + j;
        float x = o[d] * inv;
        x = fmaf(sm.gv[1][d], vv[j], x);
        x = fmaf(sm.gv[0][d], prev[d], x);
        x = fmaf(sm.gv[2][d], next[d], x);
        r[j] = x;
      }
      dst[c8] = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
  }
}

}  // namespace

int stripe_attention_tc(const float* qkv, int B, int h, int w, int K, const float* get_v0, const float* get_v1,
                        float* out, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(stripe_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_DYN);
    configured = true;
  }
  const int Lmax = (h > w ? h : w) * K;
  const int smax = B * (h > w ? h : w);
  dim3 grid((Lmax + 127) / 128, smax, kHeads);
  stripe_attention_tc_kernel<<<grid, AT_THREADS, AT_DYN, stream>>>(qkv, B, h, w, K, get_v0, get_v1, out);
  count_launch();
  return check_launch("stripe_attention_tc");
}

}  // namespace nmrf
