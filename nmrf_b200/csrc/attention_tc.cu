// A6 on the tensor cores: cross-shaped stripe attention + LePE (NMP.py:429-505) with tcgen05.
//
// One CTA = 128 query tokens of one (stripe, head); 128 threads, thread r owns query row r (= TMEM lane r).
// Keys/values are streamed in chunks of 64 tokens with an online softmax (flash style):
//     S  = (s Q) K_c^T      UMMA 128x64x32, 3xTF32 (hi/lo RN split of both operands), Q and K_c from shared memory -> TMEM [0,64)
//     p  = exp(S - m)       registers (tcgen05.ld, one row per thread), mask NMP.py:203-208, running max / sum
//     O += P V_c            UMMA 128x32x64, 3xTF32; P goes back to TENSOR MEMORY (tcgen05.st, hi at [128,192), lo at [192,256))
//                           and is the A operand from there, V_c^T (two [32 dims x 32 keys] K-major tiles) is B
//                           -> TMEM [64,96), added into the register accumulator with the softmax rescale
// Shared-memory tiles are K-major SWIZZLE_128B with 128-byte rows (32 fp32): Q [128 rows], K_c [64 keys], V_c^T 2 x [32 dims x
// 32 keys].  65 KB of shared memory and 256 TMEM columns per CTA, so two CTAs share an SM and overlap each other's
// load / MMA / softmax phases.  (v1 used 32-key chunks and staged P through shared memory: twice the round trips per key and
// 32 KB of st.shared + a proxy fence per chunk.)  Exactness: same 3xTF32 scheme as the GEMM; exp, max, sum
// and the LePE epilogue are fp32.
#include <stdlib.h>
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
constexpr int AT_QTILE = 128 * 128;               // bytes of a [128 x 32 fp32] tile
// KC keys per chunk (32 or 64): K_c is a [KC x 32] tile, V_c^T KC/32 tiles of [32 dims x 32 keys];
// TMEM columns: S [0,KC)  O [KC,KC+32)  P_hi [2KC,3KC)  P_lo [3KC,4KC)  -> 4 KC columns, so 512 / (4 KC) CTAs fit an SM
template <int KC> struct AtCfg {
  static constexpr int KTILE = KC * 128;
  static constexpr int DYN = 2 * AT_QTILE + 4 * KTILE + 1024;
  static constexpr int COL_O = KC, COL_PH = 2 * KC, COL_PL = 3 * KC, TMEM_COLS = 4 * KC;
  static constexpr int CTAS = 512 / TMEM_COLS;
};

struct AtSmem {
  uint64_t bar_s, bar_o;
  uint32_t tmem_base;
  float gv[3][32];          // LePE taps (prev, centre, next) of this head's 32 channels
};

// hi = x rounded to TF32 (nearest, two integer ops), lo = x - hi (exact) rounded the same way (tc_common.cuh: lo_tf32; the
// tensor core would truncate an unrounded lo): 5 instructions per element instead of 9 with two cvt.rna
__device__ __forceinline__ void split4(const float4 v, float4& h, float4& l) {
  h.x = rna_tf32_fast(v.x); h.y = rna_tf32_fast(v.y); h.z = rna_tf32_fast(v.z); h.w = rna_tf32_fast(v.w);
  l.x = lo_tf32(v.x, h.x); l.y = lo_tf32(v.y, h.y); l.z = lo_tf32(v.z, h.z); l.w = lo_tf32(v.w, h.w);
}
__device__ __forceinline__ uint32_t idesc_n(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int AT_KC>
__global__ void __launch_bounds__(AT_THREADS, AtCfg<AT_KC>::CTAS)
stripe_attention_tc_kernel(const float* __restrict__ qkv, int B, int h, int w, int K,
                           const float* __restrict__ get_v0, const float* __restrict__ get_v1,
                           float* __restrict__ out) {
  constexpr int AT_KTILE = AtCfg<AT_KC>::KTILE, AT_COL_O = AtCfg<AT_KC>::COL_O, AT_COL_PH = AtCfg<AT_KC>::COL_PH,
                AT_COL_PL = AtCfg<AT_KC>::COL_PL, AT_TMEM_COLS = AtCfg<AT_KC>::TMEM_COLS;
  constexpr int NKI = AT_KC / 16;                 // K items (16-byte loads) per thread and chunk
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
  uint8_t* sVh = sKl + AT_KTILE;          // tile j (keys 32j..32j+31) at + j * 4096
  uint8_t* sVl = sVh + AT_KTILE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = vertical ? sid / w : sid / h;
  const int fixed = vertical ? sid % w : sid % h;
  auto token_row = [&](int t) -> size_t {
    const int l = t / K, n = t - l * K;
    const int y = vertical ? l : fixed, x = vertical ? fixed : l;
    return ((size_t)(b * h + y) * w + x) * K + n;
  };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(AT_TMEM_COLS));
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

  // K prefetch: items i = tid + 128 e (e < 4) -> key = i / 8, 16-byte chunk = i % 8 (8 lanes = one token's 128 B).
  // V prefetch: this thread owns the 4x4 block (keys 4 kg .. +3) x (dims 4 dg .. +3): four 16-byte loads, transposed in
  // registers so that V_c^T is written with 16-byte stores; kg varies fastest over the lanes of a quarter-warp, which makes
  // the swizzled stores bank-conflict free.
  const int kg = (tid & 7) + 8 * (tid >> 6), dg = (tid >> 3) & 7;
  const bool v_owner = kg * 4 < AT_KC;                         // with 32-key chunks only 64 threads stage V
  // token -> row of qkv without a division per load: (l, n) of every item advance by a constant per chunk
  const int adv_l = AT_KC / K, adv_n = AT_KC % K;
  int kl[NKI], kn[NKI], vl[4], vn[4];
#pragma unroll
  for (int e = 0; e < NKI; ++e) { const int t = (tid + 128 * e) >> 3; kl[e] = t / K; kn[e] = t % K; }
#pragma unroll
  for (int e = 0; e < 4; ++e) { const int t = kg * 4 + e; vl[e] = t / K; vn[e] = t % K; }
  auto row_of = [&](int l_, int n_) -> size_t {
    const int y = vertical ? l_ : fixed, x = vertical ? fixed : l_;
    return ((size_t)(b * h + y) * w + x) * K + n_;
  };
  float4 kr[NKI], vr[4];
  auto fetch_kv = [&]() {                                       // next chunk (call order = chunk order)
#pragma unroll
    for (int e = 0; e < NKI; ++e) {
      kr[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kl[e] < L) kr[e] = *reinterpret_cast<const float4*>(qkv + row_of(kl[e], kn[e]) * kQkv + head * 32 + (tid & 7) * 4 + 128);
      kl[e] += adv_l; kn[e] += adv_n;
      if (kn[e] >= K) { kn[e] -= K; ++kl[e]; }
    }
    if (v_owner) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        vr[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vl[e] < L) vr[e] = *reinterpret_cast<const float4*>(qkv + row_of(vl[e], vn[e]) * kQkv + head * 32 + dg * 4 + 256);
        vl[e] += adv_l; vn[e] += adv_n;
        if (vn[e] >= K) { vn[e] -= K; ++vl[e]; }
      }
    }
  };
  const int nchunks = (Lk + AT_KC - 1) / AT_KC;
  const uint32_t idesc_s = idesc_n(AT_KC), idesc_o = idesc_n(32);
  const uint64_t dQh = make_desc(smem_u32(sQh)), dQl = make_desc(smem_u32(sQl));
  const uint64_t dKh = make_desc(smem_u32(sKh)), dKl = make_desc(smem_u32(sKl));
  const uint64_t dVh = make_desc(smem_u32(sVh)), dVl = make_desc(smem_u32(sVl));

  // shared-memory offsets of this thread's staging items (independent of the chunk)
  uint32_t k_off[NKI], v_off[4];
#pragma unroll
  for (int e = 0; e < NKI; ++e) {
    const int i = tid + 128 * e;
    k_off[e] = swz(i >> 3, i & 7);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int d = dg * 4 + e;                                  // row of V_c^T written by store e
    v_off[e] = (uint32_t)((kg >> 3) * 4096 + d * 128 + (((kg & 7) ^ (d & 7)) << 4));
  }

  fetch_kv();
  for (int c = 0; c < nchunks; ++c) {
    // ---- stage K_c (rows = keys) and V_c^T (rows = dims, columns = keys); previous chunk's MMAs are complete ----
#pragma unroll
    for (int e = 0; e < NKI; ++e) {
      float4 hi, lo;
      split4(kr[e], hi, lo);
      *reinterpret_cast<float4*>(sKh + k_off[e]) = hi;
      *reinterpret_cast<float4*>(sKl + k_off[e]) = lo;
    }
    if (v_owner) {
      const float vt[4][4] = {{vr[0].x, vr[1].x, vr[2].x, vr[3].x}, {vr[0].y, vr[1].y, vr[2].y, vr[3].y},
                              {vr[0].z, vr[1].z, vr[2].z, vr[3].z}, {vr[0].w, vr[1].w, vr[2].w, vr[3].w}};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float4 hi, lo;
        split4(make_float4(vt[e][0], vt[e][1], vt[e][2], vt[e][3]), hi, lo);
        *reinterpret_cast<float4*>(sVh + v_off[e]) = hi;
        *reinterpret_cast<float4*>(sVl + v_off[e]) = lo;
      }
    }
    if (c + 1 < nchunks) fetch_kv();                            // in flight during this chunk's MMAs and softmax
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // fresh accumulator: the small lo.hi / hi.lo products first, then hi.hi -- the tensor core's round-toward-zero accumulator
      // update (gemm_tc6.cu) then only bites on the four (eight for PV) main products
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_tf32(tmem, dQl + (uint64_t)(ks * 2), dKh + (uint64_t)(ks * 2), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_tf32(tmem, dQh + (uint64_t)(ks * 2), dKl + (uint64_t)(ks * 2), idesc_s, 1u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_tf32(tmem, dQh + (uint64_t)(ks * 2), dKh + (uint64_t)(ks * 2), idesc_s, 1u);
      umma_commit(&sm.bar_s);
    }
    __syncwarp();
    mbar_wait(&sm.bar_s, c & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float s[AT_KC];
    tmem_ld32(tmem + t_lane, s);
    if (AT_KC == 64) tmem_ld32(tmem + t_lane + 32, s + (AT_KC == 64 ? 32 : 0));
    // ---- mask + online softmax on this thread's row ---------------------------------------------------------------
    // only chunks that reach past the stripe's end or touch this row's own pixel need the mask (NMP.py:203-208)
    const int t_lo = c * AT_KC;
    if (t_lo + AT_KC > Lk || (t_lo < pix_hi && t_lo + AT_KC > pix_lo)) {
#pragma unroll
      for (int j = 0; j < AT_KC; ++j) {
        const int tj = t_lo + j;
        const bool masked = (tj >= Lk) || (tj >= pix_lo && tj < pix_hi && tj != ti);
        s[j] = masked ? -INFINITY : s[j];
      }
    }
    // four independent chains for the max and the sum (a single chain is 64 dependent operations per chunk)
    float cm[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
    for (int j = 4; j < AT_KC; ++j) cm[j & 3] = fmaxf(cm[j & 3], s[j]);
    const float mnew = fmaxf(m, fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])));
    const float msafe = (mnew == -INFINITY) ? 0.f : mnew;        // a fully masked prefix keeps p = 0, scale = 1
    const float scale = ex2(m - msafe);
    float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < AT_KC; ++j) {
      s[j] = ex2(s[j] - msafe);
      ps[j & 3] += s[j];
    }
    l = l * scale + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
    m = mnew;
    // ---- P -> tensor memory (A operand of the PV product): hi at [128,192), lo at [192,256) of this row's lane -------
#pragma unroll
    for (int g16 = 0; g16 < AT_KC / 16; ++g16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = s[g16 * 16 + j];
        const float hv = rna_tf32_fast(x);
        hi[j] = __float_as_uint(hv);
        lo[j] = __float_as_uint(lo_tf32(x, hv));
      }
      tmem_st16(tmem + t_lane + (uint32_t)(AT_COL_PH + g16 * 16), hi);
      tmem_st16(tmem + t_lane + (uint32_t)(AT_COL_PL + g16 * 16), lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // keys 8 ks .. +7: V_c^T tile ks / 4, 32-byte step ks % 4 inside its swizzle rows; P columns 8 ks .. +7.  Small products first.
#pragma unroll
      for (int ks = 0; ks < AT_KC / 8; ++ks)
        umma_tf32_ta(tmem + AT_COL_O, tmem + AT_COL_PL + ks * 8, dVh + (uint64_t)((ks >> 2) * (4096 >> 4) + (ks & 3) * 2), idesc_o, ks > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < AT_KC / 8; ++ks)
        umma_tf32_ta(tmem + AT_COL_O, tmem + AT_COL_PH + ks * 8, dVl + (uint64_t)((ks >> 2) * (4096 >> 4) + (ks & 3) * 2), idesc_o, 1u);
#pragma unroll
      for (int ks = 0; ks < AT_KC / 8; ++ks)
        umma_tf32_ta(tmem + AT_COL_O, tmem + AT_COL_PH + ks * 8, dVh + (uint64_t)((ks >> 2) * (4096 >> 4) + (ks & 3) * 2), idesc_o, 1u);
      umma_commit(&sm.bar_o);
    }
    __syncwarp();
    mbar_wait(&sm.bar_o, c & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float pv[32];
    tmem_ld32(tmem + t_lane + AT_COL_O, pv);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AT_TMEM_COLS));
  }
}

}  // namespace

namespace {
template <int KC>
int launch_stripe(const float* qkv, int B, int h, int w, int K, const float* get_v0, const float* get_v1, float* out, cudaStream_t stream) {
  static PerDevice configured;
  ensure_dynamic_smem(stripe_attention_tc_kernel<KC>, AtCfg<KC>::DYN, configured);
  const int Lmax = (h > w ? h : w) * K;
  const int smax = B * (h > w ? h : w);
  dim3 grid((Lmax + 127) / 128, smax, kHeads);
  stripe_attention_tc_kernel<KC><<<grid, AT_THREADS, AtCfg<KC>::DYN, stream>>>(qkv, B, h, w, K, get_v0, get_v1, out);
  count_launch();
  return check_launch("stripe_attention_tc");
}
}  // namespace

int stripe_attention_tc(const float* qkv, int B, int h, int w, int K, const float* get_v0, const float* get_v1,
                        float* out, cudaStream_t stream) {
  // default: 64-key chunks, two CTAs per SM (127 us per launch at 68x120x4); NMRF_B200_STRIPE_KC=32: 32-key chunks, four
  // CTAs per SM (135 us)
  static const int kc = [] { const char* e = getenv("NMRF_B200_STRIPE_KC"); return (e && atoi(e) == 32) ? 32 : 64; }();
  return kc == 64 ? launch_stripe<64>(qkv, B, h, w, K, get_v0, get_v1, out, stream)
                  : launch_stripe<32>(qkv, B, h, w, K, get_v0, get_v1, out, stream);
}

}  // namespace nmrf
