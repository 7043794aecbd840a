// Disparity Proposal Network front end:
//   A1 cost volume (submodule.py:4-23) + A2 seed extraction (DPN.py:115-125) in ONE kernel,
//   A3/A4 gather (NMP.py:618-634, 35-51), A7 tail (DPN.py:131-132).
#include "common.cuh"

namespace nmrf {
namespace {

#ifndef CV_TX
#define CV_TX 16
#endif
constexpr int TX = CV_TX;       // pixels of one image row per CTA
constexpr int CV_THREADS = 256;
constexpr int CV_WARPS = CV_THREADS / 32;
#ifndef CV_MINB
#define CV_MINB 3
#endif

__device__ __forceinline__ uint32_t cv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Row layout of every conv1d operand in shared memory: value d at position d + 4, zeros elsewhere (the padding of the k = 5
// convolution), row stride DP floats with DP / 4 odd -- so that the float4 stores of 32 threads (32 different rows) hit 32
// different bank groups, and a chunk of four outputs d0 .. d0+3 reads its eight taps (positions d0+2 .. d0+9) as 8 + 16 + 8 bytes.
__host__ __device__ __forceinline__ int cv_row_stride(int D) {
  int dp = (D + 10 + 3) & ~3;
  if (((dp >> 2) & 1) == 0) dp += 4;
  return dp;
}

// conv1d (kernel 5, padding 2) over D for one (pixel, output channel): `in` = the pixel's CI input rows, 5 CI weights in
// registers.  Four outputs per step from three aligned loads per input row (20 FMAs per 3 loads; the round-1 kernel read one
// weight from shared memory per FMA).  Chunks d0 = 4 * c for c in [c_begin, c_end).
template <int CI>
__device__ __forceinline__ void conv5_row(const float* __restrict__ in, int DP, int c_begin, int c_end, const float (&w)[CI][5], float bias,
                                          bool relu, float* __restrict__ out /* row */) {
  for (int cch = c_begin; cch < c_end; ++cch) {
    const int d0 = cch * 4;
    float acc[4] = {bias, bias, bias, bias};
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) {
      const float* r = in + ci * DP + d0;
      const float2 a = *reinterpret_cast<const float2*>(r + 2);
      const float4 b = *reinterpret_cast<const float4*>(r + 4);
      const float2 c = *reinterpret_cast<const float2*>(r + 8);
      const float v[8] = {a.x, a.y, b.x, b.y, b.z, b.w, c.x, c.y};          // positions d0+2 .. d0+9 = taps d0-2 .. d0+5
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(w[ci][k], v[j + k], acc[j]);
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    // positions 4 + d0 .. + 3: 16-byte aligned.  Outputs past D land in the row's zero tail: the caller re-zeroes it (D % 4 != 0)
    *reinterpret_cast<float4*>(out + 4 + d0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// A1 for one pixel, one warp: lane = (group g = lane % G, dq = lane / G) computes whole dot products -- output (g, d) for
// d = dq, dq + 32 / G, ... -- over the C / G channels of its group, 16 bytes of f1 and f2 per step.  No cross-lane reduction
// (the first version of this round split the CHANNELS across lanes: three to four xor-shuffle stages per four disparities
// plus index arithmetic for the scattered group sums were half of the kernel's 32 M instructions).  Lane r = lane % 8
// starts at chunk r of its group and wraps: the eight lanes of a quarter-warp, whose rows / groups are whole multiples of
// 128 B apart, read eight different bank groups.
__device__ __forceinline__ void corr_pixel(const float* __restrict__ s_f1, const float* __restrict__ s_f2, float* __restrict__ s_cv,
                                           int px, int x, int C, int G, int D, int DP, int lane) {
  const int lg = 31 - __clz(G);                 // G is a power of two
  const int g = lane & (G - 1), dq = lane >> lg, dpw = 32 >> lg;
  const int cg = C / G, nch4 = cg >> 2;         // channels / 16-byte chunks per group (>= 4: C % 128 == 0, G <= 8)
  const int r = lane & (nch4 < 8 ? 3 : 7);
  const float inv = 1.f / (float)cg;
  const float4* a4 = reinterpret_cast<const float4*>(s_f1 + px * C + g * cg);
  for (int d0 = 0; d0 < D; d0 += dpw) {
    const int d = d0 + dq, dc = min(d, D - 1);
    const float4* b4 = reinterpret_cast<const float4*>(s_f2 + (size_t)(px + (D - 1) - dc) * C + g * cg);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = r;
#pragma unroll 4
    for (int c = 0; c < nch4; ++c) {
      const float4 va = a4[i], vb = b4[i];
      s0 = fmaf(va.x, vb.x, s0); s1 = fmaf(va.y, vb.y, s1); s2 = fmaf(va.z, vb.z, s2); s3 = fmaf(va.w, vb.w, s3);
      if (++i == nch4) i = 0;
    }
    if (d < D) s_cv[(px * G + g) * DP + 4 + d] = (x >= d) ? ((s0 + s1) + (s2 + s3)) * inv : 0.f;
  }
}

// One CTA = TX consecutive pixels of one 1/8-res row (B*h*ceil(w/TX) CTAs: 544 at 68x120, 3 760 for a KITTI batch of 8; three
// CTAs share an SM).  HBM sees every byte about once and nothing between the two feature maps and the outputs leaves the SM:
//   stage   the f1 tile (TX pixels) and the f2 tile with its D-1 halo are each ONE contiguous run of an NHWC row: two TMA
//           bulk copies (cp.async.bulk -> mbarrier), no per-thread loads;
//   A1      group-wise correlation: warp = pixel, lane = (group, disparity): whole dot products per lane, no shuffles
//           (corr_pixel); the [TX,G,D] slab is assembled in shared memory and leaves as ONE contiguous run of cost_volume
//           (16-byte coalesced stores), not as scattered 4-byte stores;
//   A2      conv1d G->8->16->1: thread = (pixel, output channel), its 5 CI weights in registers, four d per step (conv5_row);
//           softmax, 1-D NMS and top-K: WARP per pixel, lane = d, shuffles (the round-1 kernel ran them on 32 threads of
//           the CTA, serially over D).
//   smem: f1 [TX][C], f2 [TX+D-1][C]  (dead after A1, re-used for h1 [TX][8][DP], h2 [TX][16][DP]), cv [TX][G][DP] (conv
//         input, zero halos), logits [TX][D], conv weights: 70 KB at C = 256, D = 24
__global__ void __launch_bounds__(CV_THREADS, CV_MINB)
cost_volume_topk_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int ntiles,
                        int h, int w, int C, int G, int D, int K, float eps,
                        nmrf_seed_weights wt,
                        float* __restrict__ cost_volume, float* __restrict__ prob_out,
                        int64_t* __restrict__ seeds) {
  extern __shared__ __align__(128) float smem[];
  const int DP = cv_row_stride(D);
  const int nch = (D + 3) >> 2;              // chunks of four disparities
  const int featN = (2 * TX + D - 1) * C, hidN = TX * 24 * DP;
  float* s_f1 = smem;                        // TX*C
  float* s_f2 = s_f1 + TX * C;               // (TX+D-1)*C
  float* s_h1 = smem;                        // TX*8*DP   (aliases the feature staging)
  float* s_h2 = s_h1 + TX * 8 * DP;          // TX*16*DP
  float* s_cv = smem + (featN > hidN ? featN : hidN);   // TX*G*DP
  float* s_lg = s_cv + TX * G * DP;          // TX*D logits
  float* s_w = s_lg + ((TX * D + 3) & ~3);   // conv weights: 8*G*5 + 8 + 16*8*5 + 16 + 16*5 + 1
  __shared__ __align__(8) unsigned long long bar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_x = (w + TX - 1) / TX;

  // ---- stage: TMA bulk copies of the two feature tiles (one elected thread), weights by everybody meanwhile ---------------
  const uint32_t bar_a = cv_smem_u32(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // tile t -> TX pixels of row b*h + y; issued by thread 0 as soon as the staging area of the previous tile is dead
  auto issue_tile = [&](int t) {
    const int x0_ = (t % tiles_x) * TX, ntx_ = min(TX, w - x0_);
    const size_t rb = (size_t)(t / tiles_x) * w;
    const int xs = max(x0_ - (D - 1), 0);                       // first f2 pixel that exists
    const uint32_t bytes1 = (uint32_t)ntx_ * C * 4, bytes2 = (uint32_t)(x0_ + ntx_ - xs) * C * 4;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes1 + bytes2) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(cv_smem_u32(s_f1)), "l"(f1 + (rb + x0_) * C), "r"(bytes1), "r"(bar_a) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(cv_smem_u32(s_f2 + (size_t)(xs - (x0_ - (D - 1))) * C)), "l"(f2 + (rb + xs) * C), "r"(bytes2), "r"(bar_a) : "memory");
  };
  if (tid == 0 && (int)blockIdx.x < ntiles) issue_tile(blockIdx.x);
  const int nw0 = 8 * G * 5, nw1 = 16 * 8 * 5, nw2 = 16 * 5;
  float* sw0 = s_w; float* sb0 = sw0 + nw0; float* sw1 = sb0 + 8; float* sb1 = sw1 + nw1;
  float* sw2 = sb1 + 16; float* sb2 = sw2 + nw2;
  for (int i = tid; i < nw0; i += CV_THREADS) sw0[i] = wt.w0[i];
  for (int i = tid; i < nw1; i += CV_THREADS) sw1[i] = wt.w1[i];
  for (int i = tid; i < nw2; i += CV_THREADS) sw2[i] = wt.w2[i];
  if (tid < 8) sb0[tid] = wt.b0[tid];
  if (tid < 16) sb1[tid] = wt.b1[tid];
  if (tid == 0) sb2[0] = wt.b2[0];

  // persistent over tiles (grid = 3 CTAs per SM): no partial last wave (544 tiles on 444 slots cost two full rounds), the conv
  // weights are staged once, and the next tile's TMA copies fly under this tile's softmax / NMS / top-K
  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
  const int tile = t % tiles_x;
  const int by = t / tiles_x;                // b*h + y
  const int x0 = tile * TX;
  const int ntx = min(TX, w - x0);           // pixels of this tile inside the row
  const size_t row_base = (size_t)by * w;    // pixel index of (b,y,0)
  for (int i = tid; i < TX * G * DP; i += CV_THREADS) s_cv[i] = 0.f;   // conv halos (and the columns of pixels past the row end)
  {                                                                   // wait for the tiles (every thread observes the barrier)
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar_a), "r"(it & 1) : "memory");
  }
  __syncthreads();

  // ---- A1: group-wise correlation, warp = pixel (TX / 8 pixels per warp).  Pixels left of the image (x < d) read whatever the
  //      halo holds: their value is SELECTED to 0.
  for (int px = warp; px < ntx; px += CV_WARPS) corr_pixel(s_f1, s_f2, s_cv, px, x0 + px, C, G, D, DP, lane);
  __syncthreads();
  // the tile's [ntx, G, D] slab is one contiguous run of cost_volume: coalesced 16-byte stores, consecutive threads ->
  // consecutive addresses (a (pixel, group) row of D floats sits at position 4 of its shared-memory row)
  {
    float* dst = cost_volume + (row_base + x0) * G * D;
    if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      const int q4 = D >> 2, n4 = ntx * G * q4;
      for (int i = tid; i < n4; i += CV_THREADS) {
        const int r = i / q4, q = i - r * q4;
        reinterpret_cast<float4*>(dst)[i] = *reinterpret_cast<const float4*>(s_cv + r * DP + 4 + q * 4);
      }
    } else {
      for (int i = tid; i < ntx * G * D; i += CV_THREADS) dst[i] = s_cv[(i / D) * DP + 4 + i % D];
    }
  }
  // features are dead: zero the halos of h1 | h2 (value d lives at d + 4; everything else in a row must read as zero padding)
  auto zero_halo = [&](float* rows, int nrows, bool left) {
    for (int r = tid; r < nrows; r += CV_THREADS) {
      float* row = rows + r * DP;
      if (left) *reinterpret_cast<float4*>(row) = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = D + 4; i < DP; ++i) row[i] = 0.f;
    }
  };
  zero_halo(s_h1, TX * 24, true);
  __syncthreads();

  // ---- A2: conv1d G->8 (k5) + ReLU: thread = (pixel, output channel, half of the disparity chunks) ------------------------
  {
    constexpr int NPART = CV_THREADS / (TX * 8);                      // TX * 8 (pixel, channel) items, NPART threads each
    const int item = tid & (TX * 8 - 1), part = tid / (TX * 8);
    const int px = item >> 3, co = item & 7;
    const int cb = part * nch / NPART, ce = (part + 1) * nch / NPART;
    float wr[8][5];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci)
#pragma unroll
      for (int k = 0; k < 5; ++k) wr[ci][k] = ci < G ? sw0[(co * G + ci) * 5 + k] : 0.f;
    const float* in = s_cv + px * G * DP;
    float* out = s_h1 + (px * 8 + co) * DP;
    switch (G) {                                   // the input-channel count is a loop bound of fully unrolled code
      case 1: conv5_row<1>(in, DP, cb, ce, reinterpret_cast<const float(&)[1][5]>(wr), sb0[co], true, out); break;
      case 2: conv5_row<2>(in, DP, cb, ce, reinterpret_cast<const float(&)[2][5]>(wr), sb0[co], true, out); break;
      case 4: conv5_row<4>(in, DP, cb, ce, reinterpret_cast<const float(&)[4][5]>(wr), sb0[co], true, out); break;
      default: conv5_row<8>(in, DP, cb, ce, wr, sb0[co], true, out); break;
    }
  }
  __syncthreads();
  if (D & 3) { zero_halo(s_h1, TX * 8, false); __syncthreads(); }      // a chunk wrote past D into the zero tail
  // ---- conv1d 8->16 (k5) + ReLU: thread = (pixel, output channel, part of the disparity chunks) -----------------------------
  {
    constexpr int NPART = CV_THREADS / (TX * 16);
    const int item = tid & (TX * 16 - 1), part = tid / (TX * 16);
    const int px = item >> 4, co = item & 15;
    float wr[8][5];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci)
#pragma unroll
      for (int k = 0; k < 5; ++k) wr[ci][k] = sw1[(co * 8 + ci) * 5 + k];
    conv5_row<8>(s_h1 + px * 8 * DP, DP, part * nch / NPART, (part + 1) * nch / NPART, wr, sb1[co], true, s_h2 + (px * 16 + co) * DP);
  }
  __syncthreads();
  if (D & 3) { zero_halo(s_h2, TX * 16, false); __syncthreads(); }
  // ---- conv1d 16->1 (k5) -> logits: thread = (pixel, four disparities) ---------------------------------------------------
  for (int item = tid; item < TX * nch; item += CV_THREADS) {
    const int px = item / nch, d0 = (item % nch) * 4;
    float acc[4] = {sb2[0], sb2[0], sb2[0], sb2[0]};
    const float* in = s_h2 + px * 16 * DP + d0;
    for (int ci = 0; ci < 16; ++ci) {
      const float* r = in + ci * DP;
      const float2 a = *reinterpret_cast<const float2*>(r + 2);
      const float4 b = *reinterpret_cast<const float4*>(r + 4);
      const float2 c = *reinterpret_cast<const float2*>(r + 8);
      const float v[8] = {a.x, a.y, b.x, b.y, b.z, b.w, c.x, c.y};
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float wk = sw2[ci * 5 + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(wk, v[j + k], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (d0 + j < D) s_lg[px * D + d0 + j] = acc[j];
  }
  __syncthreads();
  // h1 | h2 (aliasing the feature staging) are dead: the next tile's copies may land while this tile finishes
  if (tid == 0 && t + (int)gridDim.x < ntiles) issue_tile(t + gridDim.x);

  // ---- softmax over D, 1-D NMS, top-K: warp = pixel, lane holds d = lane + 32 j (D <= 128) -----------------------------------
  for (int px = warp; px < ntx; px += CV_WARPS) {
    const size_t pix = row_base + x0 + px;
    const float* lg = s_lg + px * D;
    float v[4];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = lane + 32 * j;
      v[j] = d < D ? lg[d] : -INFINITY;
      m = fmaxf(m, v[j]);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = lane + 32 * j;
      v[j] = d < D ? expf(v[j] - m) : 0.f;
      sum += v[j];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = lane + 32 * j;
      v[j] = v[j] / sum;
      if (d < D) prob_out[pix * D + d] = v[j];
    }
    // NMS (max_pool1d k3 s1 p1, -inf padding): suppressed := eps; neighbours are compared on the un-suppressed values
    float nv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = lane + 32 * j;
      float up = __shfl_up_sync(0xffffffffu, v[j], 1);        // d - 1
      float dn = __shfl_down_sync(0xffffffffu, v[j], 1);      // d + 1
      const float prev_chunk_last = __shfl_sync(0xffffffffu, j > 0 ? v[j > 0 ? j - 1 : 0] : 0.f, 31);
      const float next_chunk_first = __shfl_sync(0xffffffffu, j < 3 ? v[j < 3 ? j + 1 : 3] : 0.f, 0);
      if (lane == 0) up = j > 0 ? prev_chunk_last : -INFINITY;
      if (lane == 31) dn = j < 3 ? next_chunk_first : -INFINITY;
      if (d - 1 < 0) up = -INFINITY;
      if (d + 1 >= D) dn = -INFINITY;
      const float mx = fmaxf(fmaxf(up, v[j]), dn);
      nv[j] = d < D ? ((v[j] != mx && v[j] > eps) ? eps : v[j]) : -INFINITY;
    }
    // top-K: value descending, index ascending among equals
    for (int k = 0; k < K; ++k) {
      float best = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = lane + 32 * j;
        if (d < D && nv[j] > best) { best = nv[j]; bi = d; }     // ascending d inside a lane: strict > keeps the first
      }
      // warp arg-max in two redux instructions: probabilities are >= 0, so their bit patterns order like the values
      // (key 0 = nothing left in this lane); among the lanes holding the maximum the smallest index wins
      const unsigned key = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
      const unsigned top = __reduce_max_sync(0xffffffffu, key);
      bi = (int)__reduce_min_sync(0xffffffffu, key == top ? (unsigned)bi : 0xffffffffu);
      if (lane == 0) seeds[pix * K + k] = bi;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lane + 32 * j == bi) nv[j] = -INFINITY;
    }
  }
  }   // tile loop (the next iteration's first __syncthreads orders this tile's reads of s_lg before its rewrite)
}

// A3/A4 gather: one warp per token.
__global__ void prop_gather_kernel(const float* __restrict__ cv, const int64_t* __restrict__ seeds,
                                   int T, int G, int D, int K, double normalizer, int extended,
                                   float* __restrict__ cost36, int ld_cost, float* __restrict__ enc32) {
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= T) return;
  const int p = t / K;
  const int s = (int)seeds[t];
  for (int i = lane; i < ld_cost; i += 32) {
    float v = 0.f;
    if (i < G * 9) {
      const int g = i / 9, o = i % 9 - 4;
      const int d = min(max(s + o, 0), D - 1);
      v = cv[((size_t)p * G + g) * D + d];
    }
    cost36[(size_t)t * ld_cost + i] = v;
  }
  // extended: the seed is an integer, its encoding is evaluated in double (exact coordinate, exact power-of-two frequencies)
  if (extended) fourier32_ext((double)s, normalizer, enc32 + (size_t)t * 32, lane);
  else fourier32((float)s, (float)normalizer, enc32 + (size_t)t * 32, lane);
}

// A7 tail: one warp per token: labels = relu(dot(hidden, w) + b + seed)
__global__ void prop_head_tail_kernel(const float* __restrict__ hidden, const float* __restrict__ w,
                                      const float* __restrict__ b, const int64_t* __restrict__ seeds,
                                      int T, float* __restrict__ labels, float* __restrict__ labels_lo) {
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= T) return;
  const float4 hv = *reinterpret_cast<const float4*>(hidden + (size_t)t * 128 + lane * 4);
  const float4 wv = *reinterpret_cast<const float4*>(w + lane * 4);
  float s = hv.x * wv.x;
  s = fmaf(hv.y, wv.y, s); s = fmaf(hv.z, wv.z, s); s = fmaf(hv.w, wv.w, s);
  s = warp_sum(s);
  if (lane == 0) {
    if (labels_lo) {
      // extended label: the integer seed (up to D-1) plus a small fp32 correction is summed in double and carried as
      // hi + lo, so that ulp(label) ~ 2e-6 px does not become ~1.5e-3 rad in the 2^14 Fourier frequency downstream
      const double v = fmax((double)(s + b[0]) + (double)seeds[t], 0.0);
      const float hi = (float)v;
      labels[t] = hi;
      labels_lo[t] = (float)(v - (double)hi);
    } else {
      labels[t] = fmaxf(s + b[0] + (float)seeds[t], 0.f);
    }
  }
}

}  // namespace

int cost_volume_topk(const float* f1, const float* f2, int B, int h, int w, int C, int G, int D, int K,
                     float eps, const nmrf_seed_weights* wt, float* cost_volume, float* prob,
                     int64_t* seeds, cudaStream_t stream) {
  NMRF_REQUIRE(f1 && f2 && wt && cost_volume && prob && seeds, "cost_volume_topk: null pointer");
  NMRF_REQUIRE(C % 128 == 0 && C <= 512, "cost_volume_topk: C=%d must be a multiple of 128 (<=512)", C);
  NMRF_REQUIRE(G == 1 || G == 2 || G == 4 || G == 8, "cost_volume_topk: cost_group=%d unsupported", G);
  NMRF_REQUIRE(D >= 1 && D <= 128 && K >= 1 && K <= D, "cost_volume_topk: D=%d K=%d unsupported", D, K);
  NMRF_REQUIRE((reinterpret_cast<uintptr_t>(f1) & 15) == 0 && (reinterpret_cast<uintptr_t>(f2) & 15) == 0,
               "cost_volume_topk: feature maps must be 16-byte aligned (TMA bulk copies)");
  const int DP = cv_row_stride(D);
  const size_t featN = (size_t)(2 * TX + D - 1) * C, hidN = (size_t)TX * 24 * DP;
  const size_t smem = sizeof(float) * ((featN > hidN ? featN : hidN) + (size_t)TX * G * DP +
                                       (((size_t)TX * D + 3) & ~(size_t)3) + 8 * G * 5 + 8 + 16 * 8 * 5 + 16 + 16 * 5 + 4);
  NMRF_REQUIRE(smem <= 227 * 1024, "cost_volume_topk: C=%d D=%d needs %zu B of shared memory", C, D, smem);
  static PerDevice configured;
  ensure_dynamic_smem(cost_volume_topk_kernel, (int)smem, configured);
  const int tiles_x = (w + TX - 1) / TX, ntiles = B * h * tiles_x;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  const int grid = ntiles < per_sm * nmrf::num_sms() ? ntiles : per_sm * nmrf::num_sms();
  cost_volume_topk_kernel<<<grid, CV_THREADS, smem, stream>>>(f1, f2, ntiles, h, w, C, G, D, K, eps, *wt,
                                                             cost_volume, prob, seeds);
  count_launch();
  return check_launch("cost_volume_topk");
}

int prop_gather(const float* cv, const int64_t* seeds, int P, int G, int D, int K, double normalizer, int extended,
                float* cost36, int ld_cost, float* enc32, cudaStream_t stream) {
  NMRF_REQUIRE(cv && seeds && cost36 && enc32, "prop_gather: null pointer");
  NMRF_REQUIRE(ld_cost >= G * 9, "prop_gather: ld_cost=%d < %d", ld_cost, G * 9);
  const int T = P * K;
  const int threads = 256, blocks = (T * 32 + threads - 1) / threads;
  prop_gather_kernel<<<blocks, threads, 0, stream>>>(cv, seeds, T, G, D, K, normalizer, extended, cost36, ld_cost, enc32);
  count_launch();
  return check_launch("prop_gather");
}

int prop_head_tail(const float* hidden, const float* w, const float* b, const int64_t* seeds, int T,
                   float* labels, float* labels_lo, cudaStream_t stream) {
  NMRF_REQUIRE(hidden && w && b && seeds && labels, "prop_head_tail: null pointer");
  const int threads = 256, blocks = (T * 32 + threads - 1) / threads;
  prop_head_tail_kernel<<<blocks, threads, 0, stream>>>(hidden, w, b, seeds, T, labels, labels_lo);
  count_launch();
  return check_launch("prop_head_tail");
}

}  // namespace nmrf
