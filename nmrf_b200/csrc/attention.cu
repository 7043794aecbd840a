// Attention cores of the neural-MRF message passing (fp32, exact softmax):
//   A10 proposal_attention : K x K attention among a pixel's proposals   (NMP.py:97-103)
//   A11 window_attention   : (shifted) window attention + contextual RPE (NMP.py:241-289)
//   A6  stripe_attention   : cross-shaped stripe attention + LePE         (NMP.py:429-505)
// None of them materialises an attention matrix in HBM (the reference writes
// [windows, heads, T, T] logits: 71-125 MB per layer at 540x960).
#include "common.cuh"

namespace nmrf {
namespace {

constexpr float kScale = 0.17677669529663687f;   // 32^-0.5 (NMP.py:79,163,412)

// ------------------------------------------------------------------------------------------------
// A10: one warp per pixel, lane = head-dim channel.
// ------------------------------------------------------------------------------------------------
__global__ void proposal_attention_kernel(const float* __restrict__ qkv, int P, int K, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p >= P) return;
  const float* base = qkv + (size_t)p * K * kQkv;
  for (int hd = 0; hd < kHeads; ++hd) {
    float q[kMaxK], k[kMaxK], v[kMaxK];
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) {
      if (i < K) {
        q[i] = base[i * kQkv + hd * 32 + lane];
        k[i] = base[i * kQkv + 128 + hd * 32 + lane];
        v[i] = base[i * kQkv + 256 + hd * 32 + lane];
      } else { q[i] = k[i] = v[i] = 0.f; }
    }
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) {
      if (i >= K) break;
      float s[kMaxK];
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < kMaxK; ++j) {
        if (j < K) { s[j] = warp_sum(q[i] * k[j]) * kScale; m = fmaxf(m, s[j]); }
      }
      float sum = 0.f, o = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxK; ++j) {
        if (j < K) { const float e = expf(s[j] - m); sum += e; o = fmaf(e, v[j], o); }
      }
      out[((size_t)p * K + i) * kEmbed + hd * 32 + lane] = o / sum;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// A11: window attention.  One CTA = one head of `wpc` consecutive windows (the head's slice of
// the relative-position table is staged once per CTA).  Token slot t=(pl,n), pl = ly*ws+lx.
//   logits[i,j] = s q_i.k_j + s q_i.Rk[rel(pi,pj)] + s k_j.Rq[rel(pi,pj)] + mask     (NMP.py:263-275)
//   out_i       = sum_j A_ij v_j + sum_pj (sum_n A_i,(pj,n)) Rv[rel(pi,pj)]          (NMP.py:282)
// The two RPE logit terms are evaluated per (token, pixel) pair once (QR, KR: T x P tables)
// instead of per (token, token) pair; the Rv term uses the per-pixel bucket sums of A.
// ------------------------------------------------------------------------------------------------
struct WinParams {
  const float* qkv; const float* table; float* out;
  int B, Hp, Wp, K, ws, shift, self_edge, nwy, nwx, nwin, wpc;
};

constexpr int WIN_THREADS = 256;
constexpr int WIN_WARPS = WIN_THREADS / 32;

__global__ void __launch_bounds__(WIN_THREADS) window_attention_kernel(const WinParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ws = p.ws, K = p.K;
  const int P = ws * ws, Tw = P * K, R = (2 * ws - 1) * (2 * ws - 1);
  const int TwP = Tw + 1;
  float* sRq = smem;                 // [R][32]  (pre-scaled)
  float* sRk = sRq + R * 32;         // [R][32]
  float* sRv = sRk + R * 32;         // [R][32]
  float* qs = sRv + R * 32;          // [Tw][32] (pre-scaled)
  float* kT = qs + Tw * 32;          // [32][Tw+1]
  float* vs = kT + 32 * TwP;         // [Tw][32]
  float* QR = vs + Tw * 32;          // [Tw][P]
  float* KR = QR + Tw * P;           // [Tw][P]
  float* rowbuf = KR + Tw * P;       // [WARPS][Tw]
  float* abuf = rowbuf + WIN_WARPS * Tw;   // [WARPS][P]
  int* tok_row = reinterpret_cast<int*>(abuf + WIN_WARPS * P);   // [Tw] global token row
  int* reg = tok_row + Tw;           // [P] Swin region id (rolled coordinates)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;

  for (int i = tid; i < R * 32; i += WIN_THREADS) {
    const int r = i >> 5, d = i & 31;
    const float* row = p.table + (size_t)r * kQkv + head * 96;
    sRq[i] = row[d] * kScale;
    sRk[i] = row[32 + d];
    sRv[i] = row[64 + d];
  }

  const int w_begin = blockIdx.x * p.wpc;
  const int w_end = min(w_begin + p.wpc, p.nwin);
  for (int win = w_begin; win < w_end; ++win) {
    const int b = win / (p.nwy * p.nwx);
    const int wy = (win / p.nwx) % p.nwy, wx = win % p.nwx;
    __syncthreads();   // previous window fully consumed (and tables visible on the first pass)
    for (int t = tid; t < Tw; t += WIN_THREADS) {
      const int pl = t / K, n = t % K;
      const int yr = wy * ws + pl / ws, xr = wx * ws + pl % ws;          // rolled coordinates
      const int y = (yr + p.shift) % p.Hp, x = (xr + p.shift) % p.Wp;    // NMP.py:249-250
      tok_row[t] = ((b * p.Hp + y) * p.Wp + x) * K + n;
      if (n == 0) {
        int r = 0;
        if (p.shift > 0) {                                               // NMP.py:221-232
          const int ry = (yr >= p.Hp - ws) + (yr >= p.Hp - p.shift);
          const int rx = (xr >= p.Wp - ws) + (xr >= p.Wp - p.shift);
          r = ry * 3 + rx;
        }
        reg[pl] = r;
      }
    }
    __syncthreads();
    for (int i = tid; i < Tw * 8; i += WIN_THREADS) {                    // float4 per thread
      const int t = i >> 3, c = (i & 7) * 4;
      const float* src = p.qkv + (size_t)tok_row[t] * kQkv + head * 32 + c;
      const float4 q = *reinterpret_cast<const float4*>(src);
      const float4 k = *reinterpret_cast<const float4*>(src + 128);
      const float4 v = *reinterpret_cast<const float4*>(src + 256);
      *reinterpret_cast<float4*>(qs + t * 32 + c) = make_float4(q.x * kScale, q.y * kScale, q.z * kScale, q.w * kScale);
      kT[(c + 0) * TwP + t] = k.x; kT[(c + 1) * TwP + t] = k.y;
      kT[(c + 2) * TwP + t] = k.z; kT[(c + 3) * TwP + t] = k.w;
      *reinterpret_cast<float4*>(vs + t * 32 + c) = v;
    }
    __syncthreads();
    // QR[t][pp] = (s q_t).Rk[rel(p_t,pp)]     KR[t][pp] = k_t.(s Rq[rel(pp,p_t)])
    for (int i = tid; i < Tw * P; i += WIN_THREADS) {
      const int t = i / P, pp = i % P;
      const int pt = t / K;
      const int dy = pt / ws - pp / ws, dx = pt % ws - pp % ws;
      const float* rk = sRk + ((dy + ws - 1) * (2 * ws - 1) + (dx + ws - 1)) * 32;
      const float* rq = sRq + ((-dy + ws - 1) * (2 * ws - 1) + (-dx + ws - 1)) * 32;
      float a = 0.f, c = 0.f;
#pragma unroll 8
      for (int d = 0; d < 32; ++d) {
        a = fmaf(qs[t * 32 + d], rk[d], a);
        c = fmaf(kT[d * TwP + t], rq[d], c);
      }
      QR[i] = a;
      KR[i] = c;
    }
    __syncthreads();

    float* myrow = rowbuf + warp * Tw;
    float* myab = abuf + warp * P;
    const int nchunk = (Tw + 31) / 32;
    for (int i = warp; i < Tw; i += WIN_WARPS) {
      const int pi = i / K;
      float qreg[32];
#pragma unroll
      for (int d = 0; d < 32; d += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(qs + i * 32 + d);
        qreg[d] = t4.x; qreg[d + 1] = t4.y; qreg[d + 2] = t4.z; qreg[d + 3] = t4.w;
      }
      float m = -INFINITY;
      for (int c = 0; c < nchunk; ++c) {
        const int j = c * 32 + lane;
        float s = -INFINITY;
        if (j < Tw) {
          const int pj = j / K;
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < 32; ++d) acc = fmaf(qreg[d], kT[d * TwP + j], acc);
          acc += QR[i * P + pj] + KR[j * P + pi];
          const bool masked = (reg[pi] != reg[pj]) || (p.self_edge && pi == pj && i != j);
          s = masked ? -INFINITY : acc;
          myrow[j] = s;
        }
        m = fmaxf(m, s);
      }
      m = warp_max(m);
      float sum = 0.f;
      __syncwarp();
      for (int c = 0; c < nchunk; ++c) {
        const int j = c * 32 + lane;
        if (j < Tw) { const float e = expf(myrow[j] - m); myrow[j] = e; sum += e; }
      }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      __syncwarp();
      for (int pp = lane; pp < P; pp += 32) {       // per-pixel bucket sums (un-normalised)
        float a = 0.f;
        for (int n = 0; n < K; ++n) a += myrow[pp * K + n];
        myab[pp] = a;
      }
      __syncwarp();
      float o = 0.f;
      for (int j = 0; j < Tw; ++j) o = fmaf(myrow[j], vs[j * 32 + lane], o);
      const int yi = pi / ws, xi = pi % ws;
      for (int pp = 0; pp < P; ++pp) {
        const int r = (yi - pp / ws + ws - 1) * (2 * ws - 1) + (xi - pp % ws + ws - 1);
        o = fmaf(myab[pp], sRv[r * 32 + lane], o);
      }
      p.out[(size_t)tok_row[i] * kEmbed + head * 32 + lane] = o * inv;
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// A6: stripe attention, flash-style streaming over K/V tiles with an online softmax.
// grid = (query blocks, stripes, 4 heads); heads 0,1: image columns, heads 2,3: image rows.
// ------------------------------------------------------------------------------------------------
constexpr int ST_THREADS = 256;
constexpr int ST_WARPS = 8;
constexpr int ST_RPW = 4;                 // query rows per warp
constexpr int ST_QB = ST_WARPS * ST_RPW;  // 32 query rows per CTA
constexpr int ST_KT = 64;                 // K/V tile

__global__ void __launch_bounds__(ST_THREADS)
stripe_attention_kernel(const float* __restrict__ qkv, int B, int h, int w, int K,
                        const float* __restrict__ get_v0, const float* __restrict__ get_v1,
                        float* __restrict__ out) {
  __shared__ __align__(16) float qs[ST_QB][32];
  __shared__ __align__(16) float kT[32][ST_KT + 1];
  __shared__ __align__(16) float vs[ST_KT][32];
  __shared__ __align__(16) float pbuf[ST_WARPS][ST_RPW][ST_KT];

  const int head = blockIdx.z;
  const bool vertical = head < 2;
  const int nstripes = vertical ? B * w : B * h;
  const int sid = blockIdx.y;
  if (sid >= nstripes) return;
  const int L = vertical ? h : w;           // pixels along the stripe
  const int Lk = L * K;                     // tokens in the stripe
  const int q0 = blockIdx.x * ST_QB;
  if (q0 >= Lk) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // token (l, n) of this stripe -> global token row
  const int b = vertical ? sid / w : sid / h;
  const int fixed = vertical ? sid % w : sid % h;
  auto token_row = [&](int t) -> size_t {
    const int l = t / K, n = t % K;
    const int y = vertical ? l : fixed, x = vertical ? fixed : l;
    return ((size_t)(b * h + y) * w + x) * K + n;
  };

  for (int i = tid; i < ST_QB * 8; i += ST_THREADS) {
    const int r = i >> 3, c = (i & 7) * 4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Lk) {
      q = *reinterpret_cast<const float4*>(qkv + token_row(q0 + r) * kQkv + head * 32 + c);
      q.x *= kScale; q.y *= kScale; q.z *= kScale; q.w *= kScale;
    }
    *reinterpret_cast<float4*>(&qs[r][c]) = q;
  }

  float m[ST_RPW], l[ST_RPW], acc[ST_RPW];
#pragma unroll
  for (int r = 0; r < ST_RPW; ++r) { m[r] = -INFINITY; l[r] = 0.f; acc[r] = 0.f; }

  for (int t0 = 0; t0 < Lk; t0 += ST_KT) {
    __syncthreads();
    for (int i = tid; i < ST_KT * 8; i += ST_THREADS) {
      const int t = i >> 3, c = (i & 7) * 4;
      float4 k = make_float4(0.f, 0.f, 0.f, 0.f), v = k;
      if (t0 + t < Lk) {
        const float* src = qkv + token_row(t0 + t) * kQkv + head * 32 + c;
        k = *reinterpret_cast<const float4*>(src + 128);
        v = *reinterpret_cast<const float4*>(src + 256);
      }
      kT[c][t] = k.x; kT[c + 1][t] = k.y; kT[c + 2][t] = k.z; kT[c + 3][t] = k.w;
      *reinterpret_cast<float4*>(&vs[t][c]) = v;
    }
    __syncthreads();

    float s[ST_RPW][2];
#pragma unroll
    for (int r = 0; r < ST_RPW; ++r) s[r][0] = s[r][1] = 0.f;
#pragma unroll
    for (int dq = 0; dq < 8; ++dq) {
      float k0[4], k1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { k0[e] = kT[dq * 4 + e][lane]; k1[e] = kT[dq * 4 + e][lane + 32]; }
#pragma unroll
      for (int r = 0; r < ST_RPW; ++r) {
        const float4 q4 = *reinterpret_cast<const float4*>(&qs[warp * ST_RPW + r][dq * 4]);
        s[r][0] = fmaf(q4.x, k0[0], s[r][0]); s[r][0] = fmaf(q4.y, k0[1], s[r][0]);
        s[r][0] = fmaf(q4.z, k0[2], s[r][0]); s[r][0] = fmaf(q4.w, k0[3], s[r][0]);
        s[r][1] = fmaf(q4.x, k1[0], s[r][1]); s[r][1] = fmaf(q4.y, k1[1], s[r][1]);
        s[r][1] = fmaf(q4.z, k1[2], s[r][1]); s[r][1] = fmaf(q4.w, k1[3], s[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < ST_RPW; ++r) {
      const int ti = q0 + warp * ST_RPW + r;
      float tmax = -INFINITY;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int tj = t0 + lane + e * 32;
        const bool masked = (tj >= Lk) || (tj / K == ti / K && tj != ti);   // NMP.py:203-208
        if (masked) s[r][e] = -INFINITY;
        tmax = fmaxf(tmax, s[r][e]);
      }
      tmax = warp_max(tmax);
      const float mnew = fmaxf(m[r], tmax);
      const float scale = (mnew == -INFINITY) ? 1.f : expf(m[r] - mnew);
      const float p0 = (mnew == -INFINITY) ? 0.f : expf(s[r][0] - mnew);
      const float p1 = (mnew == -INFINITY) ? 0.f : expf(s[r][1] - mnew);
      l[r] = l[r] * scale + warp_sum(p0 + p1);
      acc[r] *= scale;
      m[r] = mnew;
      pbuf[warp][r][lane] = p0;
      pbuf[warp][r][lane + 32] = p1;
    }
    __syncwarp();
#pragma unroll 4
    for (int j = 0; j < ST_KT; j += 4) {
      const float v0 = vs[j][lane], v1 = vs[j + 1][lane], v2 = vs[j + 2][lane], v3 = vs[j + 3][lane];
#pragma unroll
      for (int r = 0; r < ST_RPW; ++r) {
        const float4 p4 = *reinterpret_cast<const float4*>(&pbuf[warp][r][j]);
        acc[r] = fmaf(p4.x, v0, acc[r]); acc[r] = fmaf(p4.y, v1, acc[r]);
        acc[r] = fmaf(p4.z, v2, acc[r]); acc[r] = fmaf(p4.w, v3, acc[r]);
      }
    }
  }

  // epilogue: normalise + LePE (NMP.py:433-449).  With 1-pixel-wide stripes only the kernel's
  // centre column (vertical) / centre row (horizontal) is live:
  //   lepe[(l,n)] = w_c V[(l,n)] + sum_n' ( w_prev V[(l-1,n')] + w_next V[(l+1,n')] )
  const float* gv = vertical ? get_v0 : get_v1;            // [64,1,3,3]
  const int ch = (head & 1) * 32 + lane;
  const float w_prev = vertical ? gv[ch * 9 + 1] : gv[ch * 9 + 3];
  const float w_c = gv[ch * 9 + 4];
  const float w_next = vertical ? gv[ch * 9 + 7] : gv[ch * 9 + 5];
  const int vcol = 256 + head * 32 + lane;
#pragma unroll
  for (int r = 0; r < ST_RPW; ++r) {
    const int ti = q0 + warp * ST_RPW + r;
    if (ti >= Lk) continue;
    const int li = ti / K;
    float o = acc[r] / l[r];
    o = fmaf(w_c, qkv[token_row(ti) * kQkv + vcol], o);
    float prev = 0.f, next = 0.f;
    for (int n = 0; n < K; ++n) {
      if (li > 0) prev += qkv[token_row((li - 1) * K + n) * kQkv + vcol];
      if (li + 1 < L) next += qkv[token_row((li + 1) * K + n) * kQkv + vcol];
    }
    o = fmaf(w_prev, prev, o);
    o = fmaf(w_next, next, o);
    out[token_row(ti) * kEmbed + head * 32 + lane] = o;
  }
}

}  // namespace

int proposal_attention(const float* qkv, int P, int K, float* out, cudaStream_t stream) {
  NMRF_REQUIRE(qkv && out, "proposal_attention: null pointer");
  NMRF_REQUIRE(K >= 1 && K <= kMaxK, "proposal_attention: K=%d unsupported (max %d)", K, kMaxK);
  const int threads = 256, blocks = (int)(((size_t)P * 32 + threads - 1) / threads);
  proposal_attention_kernel<<<blocks, threads, 0, stream>>>(qkv, P, K, out);
  count_launch();
  return check_launch("proposal_attention");
}

int window_attention(const float* qkv, const float* table, int B, int Hp, int Wp, int K, int ws, int shift,
                     int self_edge, float* out, cudaStream_t stream) {
  NMRF_REQUIRE(qkv && table && out, "window_attention: null pointer");
  NMRF_REQUIRE(ws >= 1 && Hp % ws == 0 && Wp % ws == 0, "window_attention: grid %dx%d not a multiple of ws=%d", Hp, Wp, ws);
  NMRF_REQUIRE(shift >= 0 && shift < ws, "window_attention: shift=%d", shift);
  WinParams p;
  p.qkv = qkv; p.table = table; p.out = out;
  p.B = B; p.Hp = Hp; p.Wp = Wp; p.K = K; p.ws = ws; p.shift = shift; p.self_edge = self_edge;
  p.nwy = Hp / ws; p.nwx = Wp / ws; p.nwin = B * p.nwy * p.nwx;
  const int P = ws * ws, Tw = P * K, R = (2 * ws - 1) * (2 * ws - 1);
  const size_t smem = sizeof(float) * ((size_t)3 * R * 32 + (size_t)2 * Tw * 32 + (size_t)32 * (Tw + 1) + (size_t)2 * Tw * P +
                                       (size_t)WIN_WARPS * (Tw + P)) + sizeof(int) * (size_t)(Tw + P);
  NMRF_REQUIRE(smem <= 227 * 1024, "window_attention: ws=%d K=%d needs %zu B of shared memory", ws, K, smem);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  // small windows: several windows per CTA so the table staging is amortised and the grid is ~2 waves
  int wpc = 1;
  if (Tw <= 32) wpc = 8; else if (Tw <= 64) wpc = 4;
  p.wpc = wpc;
  dim3 grid((p.nwin + wpc - 1) / wpc, kHeads);
  window_attention_kernel<<<grid, WIN_THREADS, smem, stream>>>(p);
  count_launch();
  return check_launch("window_attention");
}

int stripe_attention(const float* qkv, int B, int h, int w, int K, const float* get_v0, const float* get_v1,
                     float* out, cudaStream_t stream) {
  NMRF_REQUIRE(qkv && get_v0 && get_v1 && out, "stripe_attention: null pointer");
  const int Lmax = (h > w ? h : w) * K;
  const int smax = B * (h > w ? h : w);
  NMRF_REQUIRE(smax <= 65535, "stripe_attention: %d stripes exceed grid.y", smax);
  dim3 grid((Lmax + ST_QB - 1) / ST_QB, smax, kHeads);
  stripe_attention_kernel<<<grid, ST_THREADS, 0, stream>>>(qkv, B, h, w, K, get_v0, get_v1, out);
  count_launch();
  return check_launch("stripe_attention");
}

}  // namespace nmrf
