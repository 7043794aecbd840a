// Memory-bound stages around the message-passing stacks:
//   A8  warp_corr_embed  (NMP.py:682-720,735-743 / 839-846)
//   A9  zero_pad_rows    (NMP.py:756-759: label_rep is zero-padded AFTER the ffn)
//   A12 select_median    (NMRF.py:218-232)
//   A13 refine_tail      (NMRF.py:238-245,250-251)
#include "common.cuh"

namespace nmrf {
namespace {

// One warp per token of the PADDED grid.  NHWC maps: the 256-ch group-wise map gives each lane
// its own correlation group (8 consecutive channels = two float4), the 64-ch map a float2.
__global__ void warp_corr_embed_kernel(const float* __restrict__ f1_cc, const float* __restrict__ f2_cc,
                                       const float* __restrict__ f1_gw, const float* __restrict__ f2_gw,
                                       const float* __restrict__ labels, const float* __restrict__ labels_lo,
                                       int B, int h, int w, int K,
                                       int Hp, int Wp, int top, int left, double normalizer,
                                       float* __restrict__ feat, float* __restrict__ enc) {
  const int lane = threadIdx.x & 31;
  const long long tp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long Tp = (long long)B * Hp * Wp * K;
  if (tp >= Tp) return;
  const int n = (int)(tp % K);
  const long long pp = tp / K;
  const int xp = (int)(pp % Wp), yp = (int)((pp / Wp) % Hp), b = (int)(pp / ((long long)Wp * Hp));
  const int y = yp - top, x = xp - left;
  float* frow = feat + tp * 160;
  float* erow = enc + tp * 32;
  if (y < 0 || y >= h || x < 0 || x >= w) {
    for (int i = lane; i < 160; i += 32) frow[i] = 0.f;
    erow[lane] = 0.f;
    return;
  }
  const size_t pix = ((size_t)b * h + y) * w + x;
  const float d = labels[pix * K + n];
  float a;
  int x0;
  if (labels_lo) {                                     // extended label hi + lo: sample position and blend weight in double
    const double xr = (double)x - ((double)d + (double)labels_lo[pix * K + n]);
    const double xf = floor(xr);
    a = (float)(xr - xf);
    x0 = (int)xf;
  } else {
    const float xr = (float)x - d;                     // NMP.py:699-702
    const float xf = floorf(xr);
    a = xr - xf;
    x0 = (int)xf;
  }
  const int x1 = x0 + 1;
  const bool ok0 = x0 >= 0 && x0 <= w - 1, ok1 = x1 >= 0 && x1 <= w - 1;
  const size_t rowpix = ((size_t)b * h + y) * w;
  const float w0 = 1.f - a, w1 = a;

  // 256-channel group-wise correlation: lane = group (NMP.py:716-719, cost_group 32)
  {
    const float4* p1 = reinterpret_cast<const float4*>(f1_gw + pix * 256 + lane * 8);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a0 = p1[0], a1 = p1[1];
    float4 t00 = z, t01 = z, t10 = z, t11 = z;
    if (ok0) { const float4* q = reinterpret_cast<const float4*>(f2_gw + (rowpix + x0) * 256 + lane * 8); t00 = q[0]; t01 = q[1]; }
    if (ok1) { const float4* q = reinterpret_cast<const float4*>(f2_gw + (rowpix + x1) * 256 + lane * 8); t10 = q[0]; t11 = q[1]; }
    float s = 0.f;
    s = fmaf(a0.x, t00.x * w0 + t10.x * w1, s); s = fmaf(a0.y, t00.y * w0 + t10.y * w1, s);
    s = fmaf(a0.z, t00.z * w0 + t10.z * w1, s); s = fmaf(a0.w, t00.w * w0 + t10.w * w1, s);
    s = fmaf(a1.x, t01.x * w0 + t11.x * w1, s); s = fmaf(a1.y, t01.y * w0 + t11.y * w1, s);
    s = fmaf(a1.z, t01.z * w0 + t11.z * w1, s); s = fmaf(a1.w, t01.w * w0 + t11.w * w1, s);
    frow[128 + lane] = s * 0.125f;
  }
  // 64-channel concat features
  {
    const float2 c1 = *reinterpret_cast<const float2*>(f1_cc + pix * 64 + lane * 2);
    float2 t0 = make_float2(0.f, 0.f), t1 = t0;
    if (ok0) t0 = *reinterpret_cast<const float2*>(f2_cc + (rowpix + x0) * 64 + lane * 2);
    if (ok1) t1 = *reinterpret_cast<const float2*>(f2_cc + (rowpix + x1) * 64 + lane * 2);
    *reinterpret_cast<float2*>(frow + lane * 2) = c1;
    *reinterpret_cast<float2*>(frow + 64 + lane * 2) = make_float2(t0.x * w0 + t1.x * w1, t0.y * w0 + t1.y * w1);
  }
  if (labels_lo) fourier32_ext((double)d + (double)labels_lo[pix * K + n], normalizer, erow, lane);
  else fourier32(d, (float)normalizer, erow, lane);
}

__global__ void zero_pad_rows_kernel(float* __restrict__ x, int B, int h, int w, int K, int Hp, int Wp,
                                     int top, int left) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 per thread
  const long long total = (long long)B * Hp * Wp * K * 32;
  if (i >= total) return;
  const long long pp = (i >> 5) / K;
  const int xp = (int)(pp % Wp), yp = (int)((pp / Wp) % Hp);
  const int y = yp - top, xx = xp - left;
  if (y < 0 || y >= h || xx < 0 || xx >= w)
    reinterpret_cast<float4*>(x)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// One thread per 1/4-resolution pixel: 16 full-resolution sub-pixels, argmax over K, x2, lower median.
template <typename T>     // T = float (plain) or double (extended labels: label = hi + lo, disp_curr written as hi + lo)
__global__ void select_median_kernel(const float* __restrict__ delta, const float* __restrict__ score,
                                     const float* __restrict__ labels, const float* __restrict__ labels_lo,
                                     int B, int h, int w, int K,
                                     int Hp, int Wp, int top, int left, float* __restrict__ disp_curr,
                                     float* __restrict__ disp_curr_lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int h4 = 2 * h, w4 = 2 * w;
  if (i >= (long long)B * h4 * w4) return;
  const int X = (int)(i % w4), Y = (int)((i / w4) % h4), b = (int)(i / ((long long)w4 * h4));
  const int y = Y >> 1, x = X >> 1, u0 = (Y & 1) * 4, v0 = (X & 1) * 4;
  const size_t pix = ((size_t)b * h + y) * w + x;
  const size_t prow = (((size_t)b * Hp + y + top) * Wp + x + left) * K;
  T val[16];
  float best[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) { val[e] = (T)0; best[e] = -INFINITY; }
  for (int n = 0; n < K; ++n) {
    T lab = (T)labels[pix * K + n];
    if (sizeof(T) == 8) lab += (T)labels_lo[pix * K + n];
    const float* dr = delta + (prow + n) * 64;
    const float* sr = score + (prow + n) * 64;
#pragma unroll
    for (int du = 0; du < 4; ++du) {
      const float4 d4 = *reinterpret_cast<const float4*>(dr + (u0 + du) * 8 + v0);
      const float4 s4 = *reinterpret_cast<const float4*>(sr + (u0 + du) * 8 + v0);
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
      const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (sv[e] > best[du * 4 + e]) {              // strict: first maximum wins (torch.max)
          best[du * 4 + e] = sv[e];
          const T c = lab + (T)dv[e];
          val[du * 4 + e] = (c > (T)0 ? c : (T)0) * (T)2;
        }
      }
    }
  }
  // lower median of 16 = 8th smallest (torch.median): selection by rank counting
  T med = val[0];
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    int less = 0, eq = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) { less += val[c] < val[a]; eq += val[c] == val[a]; }
    if (less <= 7 && 7 < less + eq) med = val[a];
  }
  const float hi = (float)med;
  disp_curr[i] = hi;
  if (sizeof(T) == 8) disp_curr_lo[i] = (float)(med - (T)hi);
}

__global__ void refine_tail_kernel(const float* __restrict__ delta, const float* __restrict__ disp_curr,
                                   const float* __restrict__ disp_curr_lo,
                                   int B, int h4, int w4, int Hp4, int Wp4, int top, int left, int H, int W,
                                   float* __restrict__ disp_pred, float* __restrict__ disp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * h4 * w4) return;
  const int X = (int)(i % w4), Y = (int)((i / w4) % h4), b = (int)(i / ((long long)w4 * h4));
  const size_t row = ((size_t)b * Hp4 + Y + top) * Wp4 + X + left;
  const float base = disp_curr[i];
  const double base_d = disp_curr_lo ? (double)base + (double)disp_curr_lo[i] : 0.0;
  const int Hf = 4 * h4, Wf = 4 * w4;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float4 d4 = *reinterpret_cast<const float4*>(delta + row * 16 + u * 4);
    float4 o;
    if (disp_curr_lo) {
      o.x = (float)fmax(base_d + (double)d4.x, 0.0); o.y = (float)fmax(base_d + (double)d4.y, 0.0);
      o.z = (float)fmax(base_d + (double)d4.z, 0.0); o.w = (float)fmax(base_d + (double)d4.w, 0.0);
    } else {
      o.x = fmaxf(base + d4.x, 0.f); o.y = fmaxf(base + d4.y, 0.f);
      o.z = fmaxf(base + d4.z, 0.f); o.w = fmaxf(base + d4.w, 0.f);
    }
    const int yy = 4 * Y + u, xx = 4 * X;
    *reinterpret_cast<float4*>(disp_pred + ((size_t)b * Hf + yy) * Wf + xx) = o;
    if (yy < H) {
      float* dst = disp + ((size_t)b * H + yy) * W;
      if (xx + 0 < W) dst[xx + 0] = o.x * 4.f;
      if (xx + 1 < W) dst[xx + 1] = o.y * 4.f;
      if (xx + 2 < W) dst[xx + 2] = o.z * 4.f;
      if (xx + 3 < W) dst[xx + 3] = o.w * 4.f;
    }
  }
}

}  // namespace

int warp_corr_embed(const float* f1_cc, const float* f2_cc, const float* f1_gw, const float* f2_gw,
                    const float* labels, const float* labels_lo, int B, int h, int w, int K, int Hp, int Wp, int top, int left,
                    double normalizer, float* feat160, float* enc32, cudaStream_t stream) {
  NMRF_REQUIRE(f1_cc && f2_cc && f1_gw && f2_gw && labels && feat160 && enc32, "warp_corr_embed: null pointer");
  NMRF_REQUIRE(Hp >= h + top && Wp >= w + left && top >= 0 && left >= 0, "warp_corr_embed: bad padding");
  const long long Tp = (long long)B * Hp * Wp * K;
  const int threads = 256;
  const long long blocks = (Tp * 32 + threads - 1) / threads;
  warp_corr_embed_kernel<<<(unsigned)blocks, threads, 0, stream>>>(f1_cc, f2_cc, f1_gw, f2_gw, labels, labels_lo, B, h, w, K, Hp, Wp,
                                                                   top, left, normalizer, feat160, enc32);
  count_launch();
  return check_launch("warp_corr_embed");
}

int zero_pad_rows(float* x, int B, int h, int w, int K, int Hp, int Wp, int top, int left, cudaStream_t stream) {
  NMRF_REQUIRE(x, "zero_pad_rows: null pointer");
  if (Hp == h && Wp == w) return NMRF_OK;
  const long long total = (long long)B * Hp * Wp * K * 32;
  const int threads = 256;
  zero_pad_rows_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(x, B, h, w, K, Hp, Wp, top, left);
  count_launch();
  return check_launch("zero_pad_rows");
}

int select_median(const float* delta, const float* score, const float* labels, const float* labels_lo, int B, int h, int w, int K,
                  int Hp, int Wp, int top, int left, float* disp_curr, float* disp_curr_lo, cudaStream_t stream) {
  NMRF_REQUIRE(delta && score && labels && disp_curr, "select_median: null pointer");
  NMRF_REQUIRE((labels_lo == nullptr) == (disp_curr_lo == nullptr), "select_median: labels_lo and disp_curr_lo go together");
  const long long total = (long long)B * 4 * h * w;
  const int threads = 128;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (labels_lo)
    select_median_kernel<double><<<blocks, threads, 0, stream>>>(delta, score, labels, labels_lo, B, h, w, K, Hp, Wp, top, left,
                                                                 disp_curr, disp_curr_lo);
  else
    select_median_kernel<float><<<blocks, threads, 0, stream>>>(delta, score, labels, nullptr, B, h, w, K, Hp, Wp, top, left,
                                                                disp_curr, nullptr);
  count_launch();
  return check_launch("select_median");
}

int refine_tail(const float* delta, const float* disp_curr, const float* disp_curr_lo, int B, int h4, int w4, int Hp4, int Wp4, int top, int left,
                int H, int W, float* disp_pred, float* disp, cudaStream_t stream) {
  NMRF_REQUIRE(delta && disp_curr && disp_pred && disp, "refine_tail: null pointer");
  NMRF_REQUIRE(H <= 4 * h4 && W <= 4 * w4, "refine_tail: output %dx%d larger than padded %dx%d", H, W, 4 * h4, 4 * w4);
  const long long total = (long long)B * h4 * w4;
  const int threads = 128;
  refine_tail_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(delta, disp_curr, disp_curr_lo, B, h4, w4, Hp4, Wp4,
                                                                                         top, left, H, W, disp_pred, disp);
  count_launch();
  return check_launch("refine_tail");
}

}  // namespace nmrf

// ---- N4: the step after the path in every pipeline of the reference -----------------------------------------------------
//   disparity metrics (DispEvaluator.process, nmrf/utils/evaluation.py:345-359) reduced on the device, and the KITTI 16-bit
//   disparity encoding (writeDispKITTI, nmrf/utils/frame_utils.py:237-239)
namespace nmrf {
namespace {

constexpr int kMaxThres = 8;
struct MetricThres { float t[kMaxThres]; int n; };

// per image: acc[0] = #valid, acc[1] = sum |pr - gt|, acc[2] = #D1 outliers ((e > 3) & (e / gt > 0.05)), acc[3 + i] = #(e > t_i);
// all over the valid pixels (valid_gt & gt < max_disp, or gt < max_disp when only_valid is off).  One block-level reduction
// per 4096 pixels, one double atomicAdd per block and statistic: the sums are exact counts / double sums.
__global__ void __launch_bounds__(256)
disp_metrics_kernel(const float* __restrict__ pr, const float* __restrict__ gt, const uint8_t* __restrict__ valid_gt,
                    long long HW, float max_disp, MetricThres th, double* __restrict__ acc) {
  const int b = blockIdx.y, nstat = 3 + th.n;
  const float* p = pr + (size_t)b * HW;
  const float* g = gt + (size_t)b * HW;
  const uint8_t* v = valid_gt ? valid_gt + (size_t)b * HW : nullptr;
  double s[3 + kMaxThres];
#pragma unroll
  for (int i = 0; i < 3 + kMaxThres; ++i) s[i] = 0.0;
  const long long base = (long long)blockIdx.x * 4096;
  for (int it = 0; it < 16; ++it) {
    const long long i = base + it * 256 + threadIdx.x;
    if (i >= HW) break;
    const float gi = g[i];
    const bool ok = (gi < max_disp) && (!v || v[i] != 0);
    if (!ok) continue;
    const float e = fabsf(p[i] - gi);
    s[0] += 1.0;
    s[1] += (double)e;
    s[2] += (e > 3.f && __fdiv_rn(e, gi) > 0.05f) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < kMaxThres; ++k)
      if (k < th.n) s[3 + k] += (e > th.t[k]) ? 1.0 : 0.0;
  }
  __shared__ double red[8][3 + kMaxThres];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < nstat; ++k) {
    double x = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < nstat) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) x += red[w][threadIdx.x];
    if (x != 0.0) atomicAdd(acc + (size_t)b * nstat + threadIdx.x, x);
  }
}

// out = (uint16) round_half_even(disp * 256): numpy's `np.round(disp * 256).astype(np.uint16)` for in-range values; like
// numpy on x86 the conversion goes through int32, i.e. values beyond 65535 wrap modulo 2^16 (KITTI disparities are < 256 px)
__global__ void disp_to_kitti_u16_kernel(const float* __restrict__ disp, long long n, uint16_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (uint16_t)(int)rintf(disp[i] * 256.f);
}

}  // namespace

int disp_metrics(const float* pr, const float* gt, const uint8_t* valid_gt, int B, long long HW, float max_disp,
                 const float* thresholds_host, int n_thres, double* acc, cudaStream_t stream) {
  NMRF_REQUIRE(pr && gt && acc && B > 0 && HW > 0, "disp_metrics: bad arguments");
  NMRF_REQUIRE(n_thres >= 0 && n_thres <= kMaxThres && (n_thres == 0 || thresholds_host), "disp_metrics: %d thresholds (max %d)", n_thres, kMaxThres);
  MetricThres th;
  th.n = n_thres;
  for (int i = 0; i < kMaxThres; ++i) th.t[i] = i < n_thres ? thresholds_host[i] : 0.f;
  dim3 grid((unsigned)((HW + 4095) / 4096), B);
  disp_metrics_kernel<<<grid, 256, 0, stream>>>(pr, gt, valid_gt, HW, max_disp, th, acc);
  count_launch();
  return check_launch("disp_metrics");
}

int disp_to_kitti_u16(const float* disp, long long n, uint16_t* out, cudaStream_t stream) {
  NMRF_REQUIRE(disp && out && n >= 0, "disp_to_kitti_u16: bad arguments");
  if (n == 0) return NMRF_OK;
  disp_to_kitti_u16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(disp, n, out);
  count_launch();
  return check_launch("disp_to_kitti_u16");
}

}  // namespace nmrf
