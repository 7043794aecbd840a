// A14: multi-scale deformable attention, forward only.
// Replaces ms_deformable_im2col_gpu_kernel (ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299): the
// reference spends one thread per output float and issues 4-byte gathers; here one thread owns a
// whole head vector (Dh channels = Dh/4 float4 per bilinear corner), reads loc/weight once, and the
// warp's outputs form one contiguous 32*Dh*4-byte segment (coalesced 128-bit stores).
#include "common.cuh"

namespace nmrf {
namespace {

constexpr int kMaxLevels = 8;
struct MsdaLevels { int H[kMaxLevels]; int W[kMaxLevels]; int start[kMaxLevels]; };

template <int DH4>   // Dh / 4
__global__ void __launch_bounds__(256)
msda_forward_kernel(const float* __restrict__ value, const MsdaLevels lv_host,
                    const int64_t* __restrict__ dev_shapes, const int64_t* __restrict__ dev_start,
                    const float* __restrict__ loc, const float* __restrict__ attn,
                    long long total, int S, int M, int L, int Lq, int P, float* __restrict__ out) {
  // level geometry: by value from the host, or (reference convention, cuh:274-277) read on the device
  __shared__ MsdaLevels lv;
  if (threadIdx.x < L) {
    const int l = threadIdx.x;
    lv.H[l] = dev_shapes ? (int)dev_shapes[2 * l] : lv_host.H[l];
    lv.W[l] = dev_shapes ? (int)dev_shapes[2 * l + 1] : lv_host.W[l];
    lv.start[l] = dev_start ? (int)dev_start[l] : lv_host.start[l];
  }
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, q, m)
  if (idx >= total) return;
  const int m = (int)(idx % M);
  const long long bq = idx / M;
  const int b = (int)(bq / Lq);
  constexpr int Dh = DH4 * 4;
  float4 acc[DH4];
#pragma unroll
  for (int c = 0; c < DH4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* wptr = attn + idx * L * P;
  const float2* lptr = reinterpret_cast<const float2*>(loc) + idx * L * P;
  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const float* vbase = value + (((size_t)b * S + lv.start[l]) * M + m) * Dh;
    const long long stride_x = (long long)M * Dh, stride_y = (long long)W * M * Dh;
    for (int pt = 0; pt < P; ++pt) {
      const float2 xy = lptr[l * P + pt];
      const float wt = wptr[l * P + pt];
      const float h_im = xy.y * H - 0.5f;        // cuh:285-286
      const float w_im = xy.x * W - 0.5f;
      if (!(h_im > -1.f && w_im > -1.f && h_im < H && w_im < W)) continue;   // cuh:288
      const float hf = floorf(h_im), wf = floorf(w_im);
      const int h0 = (int)hf, w0 = (int)wf, h1 = h0 + 1, w1 = w0 + 1;
      const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
      const float c00 = hh * hw * wt, c01 = hh * lw * wt, c10 = lh * hw * wt, c11 = lh * lw * wt;
      const bool y0 = h0 >= 0, y1 = h1 <= H - 1, x0 = w0 >= 0, x1 = w1 <= W - 1;
      const float* p00 = vbase + h0 * stride_y + w0 * stride_x;
#pragma unroll
      for (int c = 0; c < DH4; ++c) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v00 = (y0 && x0) ? __ldg(reinterpret_cast<const float4*>(p00) + c) : z;
        const float4 v01 = (y0 && x1) ? __ldg(reinterpret_cast<const float4*>(p00 + stride_x) + c) : z;
        const float4 v10 = (y1 && x0) ? __ldg(reinterpret_cast<const float4*>(p00 + stride_y) + c) : z;
        const float4 v11 = (y1 && x1) ? __ldg(reinterpret_cast<const float4*>(p00 + stride_y + stride_x) + c) : z;
        acc[c].x += c00 * v00.x + c01 * v01.x + c10 * v10.x + c11 * v11.x;
        acc[c].y += c00 * v00.y + c01 * v01.y + c10 * v10.y + c11 * v11.y;
        acc[c].z += c00 * v00.z + c01 * v01.z + c10 * v10.z + c11 * v11.z;
        acc[c].w += c00 * v00.w + c01 * v01.w + c10 * v10.w + c11 * v11.w;
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(out + idx * Dh);
#pragma unroll
  for (int c = 0; c < DH4; ++c) o[c] = acc[c];
}

// Generic head dims (Dh not a multiple of 4, e.g. the reference's own toy test, ops/test.py:16): one
// thread per output scalar, same arithmetic.
__global__ void __launch_bounds__(256)
msda_forward_scalar_kernel(const float* __restrict__ value, const MsdaLevels lv_host,
                           const int64_t* __restrict__ dev_shapes, const int64_t* __restrict__ dev_start,
                           const float* __restrict__ loc, const float* __restrict__ attn,
                           long long total, int S, int M, int Dh, int L, int Lq, int P, float* __restrict__ out) {
  __shared__ MsdaLevels lv;
  if (threadIdx.x < L) {
    const int l = threadIdx.x;
    lv.H[l] = dev_shapes ? (int)dev_shapes[2 * l] : lv_host.H[l];
    lv.W[l] = dev_shapes ? (int)dev_shapes[2 * l + 1] : lv_host.W[l];
    lv.start[l] = dev_start ? (int)dev_start[l] : lv_host.start[l];
  }
  __syncthreads();
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (b, q, m, c)
  if (o >= total * Dh) return;
  const int c = (int)(o % Dh);
  const long long idx = o / Dh;
  const int m = (int)(idx % M);
  const int b = (int)((idx / M) / Lq);
  float acc = 0.f;
  for (int l = 0; l < L; ++l) {
    const int H = lv.H[l], W = lv.W[l];
    const float* vbase = value + (((size_t)b * S + lv.start[l]) * M + m) * Dh + c;
    const long long stride_x = (long long)M * Dh, stride_y = (long long)W * M * Dh;
    for (int pt = 0; pt < P; ++pt) {
      const float lx = loc[(idx * L * P + l * P + pt) * 2], ly = loc[(idx * L * P + l * P + pt) * 2 + 1];
      const float wt = attn[idx * L * P + l * P + pt];
      const float h_im = ly * H - 0.5f, w_im = lx * W - 0.5f;
      if (!(h_im > -1.f && w_im > -1.f && h_im < H && w_im < W)) continue;
      const float hf = floorf(h_im), wf = floorf(w_im);
      const int h0 = (int)hf, w0 = (int)wf, h1 = h0 + 1, w1 = w0 + 1;
      const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
      const float* p00 = vbase + h0 * stride_y + w0 * stride_x;
      const float v00 = (h0 >= 0 && w0 >= 0) ? p00[0] : 0.f;
      const float v01 = (h0 >= 0 && w1 <= W - 1) ? p00[stride_x] : 0.f;
      const float v10 = (h1 <= H - 1 && w0 >= 0) ? p00[stride_y] : 0.f;
      const float v11 = (h1 <= H - 1 && w1 <= W - 1) ? p00[stride_y + stride_x] : 0.f;
      acc += (hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11) * wt;
    }
  }
  out[o] = acc;
}

}  // namespace

int ms_deform_attn_forward(const float* value, const int64_t* shapes, const int64_t* level_start,
                           const float* loc, const float* attn, int N, int S, int M, int Dh, int L, int Lq, int P,
                           float* out, bool shapes_on_device, cudaStream_t stream) {
  NMRF_REQUIRE(value && shapes && level_start && loc && attn && out, "ms_deform_attn_forward: null pointer");
  NMRF_REQUIRE(L >= 1 && L <= kMaxLevels, "ms_deform_attn_forward: n_levels=%d unsupported (max %d)", L, kMaxLevels);
  NMRF_REQUIRE(Dh >= 1, "ms_deform_attn_forward: head dim %d", Dh);
  MsdaLevels lv = {};
  const int64_t* dshapes = shapes_on_device ? shapes : nullptr;
  const int64_t* dstart = shapes_on_device ? level_start : nullptr;
  if (!shapes_on_device) {
    long long sum = 0;
    for (int l = 0; l < L; ++l) {
      lv.H[l] = (int)shapes[2 * l];
      lv.W[l] = (int)shapes[2 * l + 1];
      lv.start[l] = (int)level_start[l];
      sum += (long long)lv.H[l] * lv.W[l];
    }
    NMRF_REQUIRE(sum <= S, "ms_deform_attn_forward: levels cover %lld > S=%d positions", sum, S);
  }
  const long long total = (long long)N * Lq * M;
  if (total == 0) return NMRF_OK;
  const int threads = 256;
  if (Dh % 4 != 0 || Dh > 64) {
    const unsigned sblocks = (unsigned)((total * Dh + threads - 1) / threads);
    msda_forward_scalar_kernel<<<sblocks, threads, 0, stream>>>(value, lv, dshapes, dstart, loc, attn, total, S, M, Dh, L,
                                                               Lq, P, out);
    count_launch();
    return check_launch("ms_deform_attn_forward(scalar)");
  }
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
#define NMRF_MSDA_CASE(D4)                                                                             \
  case D4:                                                                                             \
    msda_forward_kernel<D4><<<blocks, threads, 0, stream>>>(value, lv, dshapes, dstart, loc, attn, total, S, M, L, Lq, P, out); \
    break;
  switch (Dh / 4) {
    NMRF_MSDA_CASE(1) NMRF_MSDA_CASE(2) NMRF_MSDA_CASE(3) NMRF_MSDA_CASE(4) NMRF_MSDA_CASE(5) NMRF_MSDA_CASE(6)
    NMRF_MSDA_CASE(7) NMRF_MSDA_CASE(8) NMRF_MSDA_CASE(9) NMRF_MSDA_CASE(10) NMRF_MSDA_CASE(11) NMRF_MSDA_CASE(12)
    NMRF_MSDA_CASE(13) NMRF_MSDA_CASE(14) NMRF_MSDA_CASE(15) NMRF_MSDA_CASE(16)
  }
#undef NMRF_MSDA_CASE
  count_launch();
  return check_launch("ms_deform_attn_forward");
}

}  // namespace nmrf
