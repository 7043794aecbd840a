// N2 ("next"): element-wise glue of the convolutional feature extractor and the conv heads, NHWC.
// The convolutions are nmrf_conv2d (gemm_tc6.cu, CONV mode); fused here is everything between two convolutions, which in
// torch took 5 of the encoder's 7.8 ms (NCHW<->NHWC copies around InstanceNorm, batch_norm statistics / transform, ReLU,
// residual add):
//   instnorm_stats : per (sample, channel) sum and sum of squares over H*W   (reference: nn.InstanceNorm2d, backbone.py:28-41)
//   instnorm_apply : y = relu?(IN(x)) [+ r | + IN(r)] -> relu?
//   image_prep     : N3 prologue (replicate pad, normalisation, left/right batching, the stem's zero border)
//   avgpool2_split : feat@1/8 = avg_pool2d(feat@1/4) (backbone.py:96-98)
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using tc::rna_tf32_fast;

constexpr int ST_PIX = 16;     // pixels per thread in the statistics pass (fp32 partial sums stay short; enough CTAs to fill HBM)

// grid (chunks, N); block = (C/4) * ppb threads: thread -> (pixel lane, 4 channels)
__global__ void instnorm_stats_kernel(const float* __restrict__ x, int HW, int C, double* __restrict__ stats) {
  extern __shared__ double red[];                  // [ppb][C][2]
  const int c4n = C >> 2, ppb = blockDim.x / c4n;
  const int pl = threadIdx.x / c4n, c4 = threadIdx.x % c4n;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * ppb * ST_PIX;
  const float* base = x + ((size_t)n * HW) * C + c4 * 4;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int k = 0; k < ST_PIX; ++k) {
    const long long p = p0 + (long long)k * ppb + pl;
    if (p < HW) {
      const float4 v = *reinterpret_cast<const float4*>(base + (size_t)p * C);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[((size_t)pl * C + c4 * 4 + j) * 2 + 0] = (double)s[j];
    red[((size_t)pl * C + c4 * 4 + j) * 2 + 1] = (double)q[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
    double a = 0.0;
    for (int l = 0; l < ppb; ++l) a += red[(size_t)l * C * 2 + i];
    atomicAdd(stats + (size_t)n * C * 2 + i, a);
  }
}

__device__ __forceinline__ void mean_rstd(const double* st, int HW, float& mean, float& rstd) {
  const double m = st[0] / HW;
  double var = st[1] / HW - m * m;                 // biased variance, like InstanceNorm
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + 1e-5));
}

struct ApplyArgs {
  const float* x; const double* xs; const float* r; const double* rs;
  float* plain;
  int HW, C, relu_inner, relu_outer;
  long long per_sample4;                            // HW*C/4
};
constexpr int AP_ITEMS = 4;                         // float4 items per thread

// grid (chunks, N): a CTA works inside ONE sample, so the fp64 mean / rstd of the sample's channels are computed once per
// CTA into shared memory (they used to be recomputed per thread and channel: two fp64 divisions and a square root for every
// four outputs made this kernel compute-bound at 3 TB/s)
__global__ void instnorm_apply_kernel(const ApplyArgs a) {
  extern __shared__ float ms[];                     // mean_x[C] rstd_x[C] mean_r[C] rstd_r[C]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float m = 0.f, rs = 1.f;
    if (a.xs) mean_rstd(a.xs + ((size_t)n * a.C + c) * 2, a.HW, m, rs);
    ms[c] = m; ms[a.C + c] = rs;
    m = 0.f; rs = 1.f;
    if (a.rs) mean_rstd(a.rs + ((size_t)n * a.C + c) * 2, a.HW, m, rs);
    ms[2 * a.C + c] = m; ms[3 * a.C + c] = rs;
  }
  __syncthreads();
  const int c4n = a.C >> 2;
  const long long base = (long long)n * a.per_sample4;
#pragma unroll
  for (int it = 0; it < AP_ITEMS; ++it) {
    const long long li = ((long long)blockIdx.x * AP_ITEMS + it) * blockDim.x + threadIdx.x;
    if (li >= a.per_sample4) break;
    const int c = (int)(li % c4n) * 4;
    const long long pix = (base + li) / c4n;        // n*HW + p
    float4 v = *reinterpret_cast<const float4*>(a.x + pix * a.C + c);
    float o[4] = {v.x, v.y, v.z, v.w};
    if (a.xs) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (o[j] - ms[c + j]) * ms[a.C + c + j];
    }
    if (a.relu_inner) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    if (a.r) {
      const float4 rv = *reinterpret_cast<const float4*>(a.r + pix * a.C + c);
      float rr[4] = {rv.x, rv.y, rv.z, rv.w};
      if (a.rs) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = (rr[j] - ms[2 * a.C + c + j]) * ms[3 * a.C + c + j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += rr[j];
    }
    if (a.relu_outer) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    if (a.plain) *reinterpret_cast<float4*>(a.plain + pix * a.C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// left / right images -> [2B, Hp+6, Wp+8, 4]: 2 (x / 255) - 1 (backbone.py:86) as RGB0 pixels inside the stem's zero border;
// thread = one pixel of the output [2B, Hp+6, Wp+8] (both images, zero border included); the replicate padding of InputPadder
// (frame_utils.py:273-275, right / bottom) is a clamp of the source coordinate, and the source may be NCHW or NHWC (strides)
__global__ void image_prep_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int H, int W, int Hp, int Wp,
                                  long long sb, long long sc, long long sy, long long sx, long long per_img, long long total,
                                  float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Wb = Wp + 8, Hb = Hp + 6;
  const long long j = i < per_img ? i : i - per_img;
  const int x = (int)(j % Wb) - 3, y = (int)((j / Wb) % Hb) - 3;
  const long long b = j / ((long long)Wb * Hb);
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (x >= 0 && x < Wp && y >= 0 && y < Hp) {
    const float* src = (i < per_img ? img1 : img2) + b * sb + (long long)min(y, H - 1) * sy + (long long)min(x, W - 1) * sx;
    o.x = 2.f * __fdiv_rn(src[0], 255.f) - 1.f;
    o.y = 2.f * __fdiv_rn(src[sc], 255.f) - 1.f;
    o.z = 2.f * __fdiv_rn(src[2 * sc], 255.f) - 1.f;
  }
  reinterpret_cast<float4*>(out)[i] = o;
}

// 2x2 average pool of an NHWC map (backbone.py:96-98) -> plain halves (first N/2 samples to out_a, the rest to out_b: the hot
// path's f1_8 / f2_8) and the [hi | lo | hi] operand of the heads' 3x3 convolution; thread = 4 channels of one output pixel
__global__ void avgpool2_split_kernel(const float* __restrict__ x, int N, int h, int w, int C, float* __restrict__ out_a,
                                      float* __restrict__ out_b, float* __restrict__ out_all) {
  const int c4n = C >> 2, ho = h >> 1, wo = w >> 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * ho * wo * c4n;
  if (i >= total) return;
  const int c = (int)(i % c4n) * 4;
  long long p = i / c4n;
  const int xo = (int)(p % wo); p /= wo;
  const int yo = (int)(p % ho);
  const int n = (int)(p / ho);
  const float* s = x + (((size_t)n * h + 2 * yo) * w + 2 * xo) * C + c;
  const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + C);
  const float4 e = *reinterpret_cast<const float4*>(s + (size_t)w * C), f = *reinterpret_cast<const float4*>(s + (size_t)w * C + C);
  float o[4] = {((a.x + b.x) + (e.x + f.x)) * 0.25f, ((a.y + b.y) + (e.y + f.y)) * 0.25f, ((a.z + b.z) + (e.z + f.z)) * 0.25f,
                ((a.w + b.w) + (e.w + f.w)) * 0.25f};
  const size_t pix = ((size_t)n * ho + yo) * wo + xo;
  const int half = N >> 1;
  float* plain = n < half ? out_a + pix * C + c : out_b + (pix - (size_t)half * ho * wo) * C + c;
  *reinterpret_cast<float4*>(plain) = make_float4(o[0], o[1], o[2], o[3]);
  if (out_all) *reinterpret_cast<float4*>(out_all + pix * C + c) = make_float4(o[0], o[1], o[2], o[3]);
}
}  // namespace

int image_prep(const float* img1, const float* img2, int B, int H, int W, int Hp, int Wp, long long sb, long long sc,
               long long sy, long long sx, float* out, cudaStream_t stream) {
  NMRF_REQUIRE(img1 && img2 && out && B > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W, "image_prep: bad arguments");
  const long long per = (long long)B * (Hp + 6) * (Wp + 8), total = 2 * per;
  image_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(img1, img2, H, W, Hp, Wp, sb, sc, sy, sx, per, total, out);
  count_launch();
  return check_launch("image_prep");
}

int avgpool2_split(const float* x, int N, int h, int w, int C, float* out_a, float* out_b, float* out_all, cudaStream_t stream) {
  NMRF_REQUIRE(x && out_a && out_b, "avgpool2_split: null pointer");
  NMRF_REQUIRE(N > 0 && N % 2 == 0 && h >= 2 && w >= 2 && C % 4 == 0, "avgpool2_split: N=%d h=%d w=%d C=%d", N, h, w, C);
  const long long total = (long long)N * (h / 2) * (w / 2) * (C / 4);
  avgpool2_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, N, h, w, C, out_a, out_b, out_all);
  count_launch();
  return check_launch("avgpool2_split");
}

int instnorm_stats(const float* x, int N, int HW, int C, double* stats, cudaStream_t stream) {
  NMRF_REQUIRE(x && stats && N > 0 && HW > 0, "instnorm_stats: bad arguments");
  NMRF_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024, "instnorm_stats: C=%d must be a multiple of 4 (<= 1024)", C);
  const int c4n = C / 4;
  int ppb = 256 / c4n;
  if (ppb < 1) ppb = 1;
  const int threads = ppb * c4n;
  const int chunks = (HW + ppb * ST_PIX - 1) / (ppb * ST_PIX);
  const size_t smem = (size_t)ppb * C * 2 * sizeof(double);
  NMRF_REQUIRE(smem <= 48 * 1024, "instnorm_stats: C=%d needs %zu B of shared memory", C, smem);
  instnorm_stats_kernel<<<dim3(chunks, N), threads, smem, stream>>>(x, HW, C, stats);
  count_launch();
  return check_launch("instnorm_stats");
}

int instnorm_apply(const float* x, const double* x_stats, const float* r, const double* r_stats, int N, int HW, int C,
                   int relu_inner, int relu_outer, float* out_plain, cudaStream_t stream) {
  NMRF_REQUIRE(x && N > 0 && HW > 0 && out_plain, "instnorm_apply: bad arguments");
  NMRF_REQUIRE(C % 4 == 0, "instnorm_apply: C=%d must be a multiple of 4", C);
  NMRF_REQUIRE(r || !r_stats, "instnorm_apply: r_stats without r");
  ApplyArgs a;
  a.x = x; a.xs = x_stats; a.r = r; a.rs = r_stats; a.plain = out_plain;
  a.HW = HW; a.C = C; a.relu_inner = relu_inner; a.relu_outer = relu_outer;
  NMRF_REQUIRE(C <= 2048, "instnorm_apply: C=%d too large", C);
  a.per_sample4 = (long long)HW * (C / 4);
  const unsigned chunks = (unsigned)((a.per_sample4 + 256 * AP_ITEMS - 1) / (256 * AP_ITEMS));
  instnorm_apply_kernel<<<dim3(chunks, N), 256, 4 * C * sizeof(float), stream>>>(a);
  count_launch();
  return check_launch("instnorm_apply");
}

}  // namespace nmrf
