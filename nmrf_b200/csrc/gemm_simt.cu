// Fused token GEMM, fp32 FMA path (exact-fp32 arithmetic; the 1e-3 EPE bar rules out
// single-pass TF32/BF16, see DESIGN.md "precision").
//
//   Y[r,n] = act( sum_k A[r,k] W[n,k] + bias[n] ) (+ R[r,n]),
//   A[r,:] = concat( LayerNorm?(X[r,0:Kx]), E[r/ediv, 0:Ke] )
//
// 128x128 CTA tile, BK=16, 256 threads, 8x8 register tile per thread (split 4+4 in both
// directions so shared-memory reads are conflict-free float4s), register-prefetch double
// buffering.  LayerNorm statistics are computed once per CTA for its 128 rows.
#include "common.cuh"

namespace nmrf {

namespace {
constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PADM = 4;

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  return v;
}

__global__ void __launch_bounds__(NT) token_gemm_kernel(const nmrf_gemm_args a) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Ws[2][BK][BN + PADM];
  __shared__ float s_mean[BM], s_rstd[BM];

  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ktot = a.Kx + a.Ke;
  const int nk = (Ktot + BK - 1) / BK;
  const bool ln = a.ln_gamma != nullptr;

  if (ln) {  // Kx == 128: one float4 per lane
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = 0; i < BM / 8; ++i) {
      const int lr = warp * (BM / 8) + i;
      const int r = row0 + lr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < a.rows) v = *reinterpret_cast<const float4*>(a.X + (size_t)r * a.ldx + lane * 4);
      const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / 128.f);
      const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
      const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f);
      if (lane == 0) {
        s_mean[lr] = mean;
        s_rstd[lr] = 1.f / sqrtf(var + 1e-5f);
      }
    }
    __syncthreads();
  }

  // loader mapping: each thread moves 8 consecutive k of one row (A) and of one n (W)
  const int l_row = tid >> 1;
  const int l_k = (tid & 1) * 8;
  const int g_row = row0 + l_row;
  const int g_n = n0 + l_row;
  const bool row_ok = g_row < a.rows;
  const bool n_ok = g_n < a.N;
  const float* xrow = a.X + (size_t)(row_ok ? g_row : 0) * a.ldx;
  const float* erow = a.E ? a.E + (size_t)((row_ok ? g_row : 0) / a.ediv) * a.lde : nullptr;
  const float* wrow = a.W + (size_t)(n_ok ? g_n : 0) * a.ldw;
  const float mean = ln ? s_mean[l_row] : 0.f;
  const float rstd = ln ? s_rstd[l_row] : 1.f;

  float4 ra[2], rw[2];
  auto load = [&](int kc) {
    const int k = kc * BK + l_k;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = k + h * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok && kk < Ktot) {
        if (kk < a.Kx) {
          v = *reinterpret_cast<const float4*>(xrow + kk);
          if (ln) {
            const float4 g = *reinterpret_cast<const float4*>(a.ln_gamma + kk);
            const float4 b = *reinterpret_cast<const float4*>(a.ln_beta + kk);
            v.x = (v.x - mean) * rstd * g.x + b.x;
            v.y = (v.y - mean) * rstd * g.y + b.y;
            v.z = (v.z - mean) * rstd * g.z + b.z;
            v.w = (v.w - mean) * rstd * g.w + b.w;
          }
        } else {
          v = *reinterpret_cast<const float4*>(erow + (kk - a.Kx));
        }
      }
      ra[h] = v;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n_ok && kk < Ktot) w = *reinterpret_cast<const float4*>(wrow + kk);
      rw[h] = w;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = l_k + h * 4;
      As[buf][kk + 0][l_row] = ra[h].x;
      As[buf][kk + 1][l_row] = ra[h].y;
      As[buf][kk + 2][l_row] = ra[h].z;
      As[buf][kk + 3][l_row] = ra[h].w;
      Ws[buf][kk + 0][l_row] = rw[h].x;
      Ws[buf][kk + 1][l_row] = rw[h].y;
      Ws[buf][kk + 2][l_row] = rw[h].z;
      Ws[buf][kk + 3][l_row] = rw[h].w;
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load(0);
  store(0);
  __syncthreads();
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nk) load(kc + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (kc + 1 < nk) {
      store(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: bias, activation, residual, float4 stores
#pragma unroll
  for (int jh = 0; jh < 2; ++jh) {
    const int n = n0 + jh * 64 + tx * 4;
    if (n >= a.N) continue;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bias = *reinterpret_cast<const float4*>(a.bias + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row0 + (i >> 2) * 64 + ty * 4 + (i & 3);
      if (r >= a.rows) continue;
      float4 v;
      v.x = act_fn(acc[i][jh * 4 + 0] + bias.x, a.act);
      v.y = act_fn(acc[i][jh * 4 + 1] + bias.y, a.act);
      v.z = act_fn(acc[i][jh * 4 + 2] + bias.z, a.act);
      v.w = act_fn(acc[i][jh * 4 + 3] + bias.w, a.act);
      if (a.R) {
        const float4 rr = *reinterpret_cast<const float4*>(a.R + (size_t)r * a.ldr + n);
        v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
      }
      *reinterpret_cast<float4*>(a.Y + (size_t)r * a.ldy + n) = v;
    }
  }
}
}  // namespace

int token_gemm_simt(const nmrf_gemm_args& a, cudaStream_t stream) {
  dim3 grid((a.rows + BM - 1) / BM, (a.N + BN - 1) / BN);
  token_gemm_kernel<<<grid, NT, 0, stream>>>(a);
  count_launch();
  return check_launch("token_gemm_simt");
}

}  // namespace nmrf
