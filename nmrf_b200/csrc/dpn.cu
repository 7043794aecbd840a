// Disparity Proposal Network front end:
//   A1 cost volume (submodule.py:4-23) + A2 seed extraction (DPN.py:115-125) in ONE kernel,
//   A3/A4 gather (NMP.py:618-634, 35-51), A7 tail (DPN.py:131-132).
#include "common.cuh"

namespace nmrf {
namespace {

constexpr int TX = 32;          // pixels of one image row per CTA
constexpr int CV_THREADS = 256;

// One CTA = TX consecutive pixels of one 1/8-res row.  Both feature rows are staged in shared
// memory once (f2 with a D-1 halo to the left), so HBM sees each feature byte ~once; the cost
// slab [TX,G,D], the three tiny conv1d layers, softmax, NMS and top-K never leave the SM.
//   smem: f1 [TX][C+4], f2 [TX+D-1][C+4]  (dead after A1, re-used for h1 [TX][8][D+4], h2 [TX][16][D+4]),
//         cv [TX][G][D+4], logits [TX][D], conv weights
__global__ void __launch_bounds__(CV_THREADS)
cost_volume_topk_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                        int h, int w, int C, int G, int D, int K, float eps,
                        nmrf_seed_weights wt,
                        float* __restrict__ cost_volume, float* __restrict__ prob_out,
                        int64_t* __restrict__ seeds) {
  extern __shared__ __align__(16) float smem[];
  const int CS = C + 4;                      // padded channel stride (bank spread for float4 rows)
  const int DP = D + 4;                      // conv halo of 2 on both sides
  const int featN = (2 * TX + D - 1) * CS, hidN = TX * 24 * DP;
  float* s_f1 = smem;                        // TX*CS
  float* s_f2 = s_f1 + TX * CS;              // (TX+D-1)*CS
  float* s_h1 = smem;                        // TX*8*DP   (aliases the feature staging)
  float* s_h2 = s_h1 + TX * 8 * DP;          // TX*16*DP
  float* s_cv = smem + (featN > hidN ? featN : hidN);   // TX*G*DP
  float* s_lg = s_cv + TX * G * DP;          // TX*D logits -> prob
  float* s_w = s_lg + TX * D;                // conv weights: 8*G*5 + 8 + 16*8*5 + 16 + 16*5 + 1

  const int tid = threadIdx.x;
  const int tiles_x = (w + TX - 1) / TX;
  const int tile = blockIdx.x % tiles_x;
  const int by = blockIdx.x / tiles_x;       // b*h + y
  const int x0 = tile * TX;
  const size_t row_base = (size_t)by * w;    // pixel index of (b,y,0)

  // ---- stage weights and features -----------------------------------------------------------
  const int nw0 = 8 * G * 5, nw1 = 16 * 8 * 5, nw2 = 16 * 5;
  float* sw0 = s_w; float* sb0 = sw0 + nw0; float* sw1 = sb0 + 8; float* sb1 = sw1 + nw1;
  float* sw2 = sb1 + 16; float* sb2 = sw2 + nw2;
  for (int i = tid; i < nw0; i += CV_THREADS) sw0[i] = wt.w0[i];
  for (int i = tid; i < nw1; i += CV_THREADS) sw1[i] = wt.w1[i];
  for (int i = tid; i < nw2; i += CV_THREADS) sw2[i] = wt.w2[i];
  if (tid < 8) sb0[tid] = wt.b0[tid];
  if (tid < 16) sb1[tid] = wt.b1[tid];
  if (tid == 0) sb2[0] = wt.b2[0];

  const int c4 = C / 4;
  for (int i = tid; i < TX * c4; i += CV_THREADS) {
    const int px = i / c4, c = (i % c4) * 4;
    const int x = x0 + px;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x < w) v = __ldg(reinterpret_cast<const float4*>(f1 + (row_base + x) * C + c));
    *reinterpret_cast<float4*>(s_f1 + px * CS + c) = v;
  }
  for (int i = tid; i < (TX + D - 1) * c4; i += CV_THREADS) {
    const int px = i / c4, c = (i % c4) * 4;
    const int x = x0 - (D - 1) + px;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x >= 0 && x < w) v = __ldg(reinterpret_cast<const float4*>(f2 + (row_base + x) * C + c));
    *reinterpret_cast<float4*>(s_f2 + px * CS + c) = v;
  }
  for (int i = tid; i < TX * G * DP; i += CV_THREADS) s_cv[i] = 0.f;   // conv halo
  __syncthreads();

  // ---- A1: group-wise correlation.  One warp per (pixel, d): lanes span channels, a group is
  //      C/G consecutive channels = (32/G) lanes when C/32 channels sit in each lane.
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int cpl = C / 32;                       // channels per lane (8 for C=256, 4 for C=128)
    const int lanes_per_group = 32 / G;
    const float inv = 1.f / (float)(C / G);
    for (int item = warp; item < TX * D; item += CV_THREADS / 32) {
      const int px = item / D, d = item % D;
      const int x = x0 + px;
      const float* a = s_f1 + px * CS + lane * cpl;
      const float* b = s_f2 + (px + (D - 1) - d) * CS + lane * cpl;
      float s = 0.f;
      for (int c = 0; c < cpl; c += 4) {
        const float4 va = *reinterpret_cast<const float4*>(a + c);
        const float4 vb = *reinterpret_cast<const float4*>(b + c);
        s = fmaf(va.x, vb.x, s); s = fmaf(va.y, vb.y, s);
        s = fmaf(va.z, vb.z, s); s = fmaf(va.w, vb.w, s);
      }
      for (int o = lanes_per_group >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((lane % lanes_per_group) == 0 && x < w) {
        const int g = lane / lanes_per_group;
        const float v = (x >= d) ? s * inv : 0.f;
        s_cv[(px * G + g) * DP + 2 + d] = v;
        cost_volume[((row_base + x) * G + g) * D + d] = v;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < TX * 24 * DP; i += CV_THREADS) s_h1[i] = 0.f;   // features are dead: zero h1|h2 (+halos)
  __syncthreads();

  // ---- A2: conv1d 4->8 (k5) + ReLU ------------------------------------------------------------
  for (int item = tid; item < TX * D; item += CV_THREADS) {
    const int px = item / D, d = item % D;
    float o[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) o[co] = sb0[co];
    for (int ci = 0; ci < G; ++ci) {
      const float* src = s_cv + (px * G + ci) * DP + d;   // taps d-2..d+2 live at +0..+4
      float t[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) t[k] = src[k];
#pragma unroll
      for (int co = 0; co < 8; ++co)
#pragma unroll
        for (int k = 0; k < 5; ++k) o[co] = fmaf(sw0[(co * G + ci) * 5 + k], t[k], o[co]);
    }
#pragma unroll
    for (int co = 0; co < 8; ++co) s_h1[(px * 8 + co) * DP + 2 + d] = fmaxf(o[co], 0.f);
  }
  __syncthreads();
  // ---- conv1d 8->16 (k5) + ReLU ---------------------------------------------------------------
  for (int item = tid; item < TX * D; item += CV_THREADS) {
    const int px = item / D, d = item % D;
    float o[16];
#pragma unroll
    for (int co = 0; co < 16; ++co) o[co] = sb1[co];
    for (int ci = 0; ci < 8; ++ci) {
      const float* src = s_h1 + (px * 8 + ci) * DP + d;
      float t[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) t[k] = src[k];
#pragma unroll
      for (int co = 0; co < 16; ++co)
#pragma unroll
        for (int k = 0; k < 5; ++k) o[co] = fmaf(sw1[(co * 8 + ci) * 5 + k], t[k], o[co]);
    }
#pragma unroll
    for (int co = 0; co < 16; ++co) s_h2[(px * 16 + co) * DP + 2 + d] = fmaxf(o[co], 0.f);
  }
  __syncthreads();
  // ---- conv1d 16->1 (k5) -> logits --------------------------------------------------------------
  for (int item = tid; item < TX * D; item += CV_THREADS) {
    const int px = item / D, d = item % D;
    float o = sb2[0];
    for (int ci = 0; ci < 16; ++ci) {
      const float* src = s_h2 + (px * 16 + ci) * DP + d;
#pragma unroll
      for (int k = 0; k < 5; ++k) o = fmaf(sw2[ci * 5 + k], src[k], o);
    }
    s_lg[px * D + d] = o;
  }
  __syncthreads();

  // ---- softmax over D, 1-D NMS, top-K: one thread per pixel (D is small) ------------------------
  if (tid < TX && x0 + tid < w) {
    float* p = s_lg + tid * D;
    const size_t pix = row_base + x0 + tid;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, p[d]);
    float sum = 0.f;
    for (int d = 0; d < D; ++d) { const float e = expf(p[d] - m); p[d] = e; sum += e; }
    for (int d = 0; d < D; ++d) { const float v = p[d] / sum; p[d] = v; prob_out[pix * D + d] = v; }
    // NMS (max_pool1d k3 s1 p1, -inf padding): suppressed := eps. Done out-of-place via a
    // rolling window so neighbours are compared on the un-suppressed values.
    float prev = -INFINITY, cur = p[0];
    for (int d = 0; d < D; ++d) {
      const float nxt = (d + 1 < D) ? p[d + 1] : -INFINITY;
      const float mx = fmaxf(fmaxf(prev, cur), nxt);
      const float v = (cur != mx && cur > eps) ? eps : cur;
      p[d] = v;
      prev = cur; cur = nxt;
    }
    // top-K: value descending, index ascending among equals
    for (int k = 0; k < K; ++k) {
      float best = -INFINITY; int bi = 0;
      for (int d = 0; d < D; ++d) if (p[d] > best) { best = p[d]; bi = d; }
      seeds[pix * K + k] = bi;
      p[bi] = -INFINITY;
    }
  }
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
  const int CS = C + 4, DP = D + 4;
  const size_t featN = (size_t)(2 * TX + D - 1) * CS, hidN = (size_t)TX * 24 * DP;
  const size_t smem = sizeof(float) * ((featN > hidN ? featN : hidN) + (size_t)TX * G * DP + (size_t)TX * D +
                                       8 * G * 5 + 8 + 16 * 8 * 5 + 16 + 16 * 5 + 4);
  NMRF_REQUIRE(smem <= 227 * 1024, "cost_volume_topk: C=%d D=%d needs %zu B of shared memory", C, D, smem);
  static PerDevice configured;
  ensure_dynamic_smem(cost_volume_topk_kernel, (int)smem, configured);
  const int tiles_x = (w + TX - 1) / TX;
  cost_volume_topk_kernel<<<B * h * tiles_x, CV_THREADS, smem, stream>>>(f1, f2, h, w, C, G, D, K, eps, *wt,
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
