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
// A10: one warp per pixel; lane = (head, 4 of its 32 channels): the four heads run side by side, every load is a float4 (a
// warp instruction covers the 512 contiguous bytes of one of q / k / v of a token) and a q.k dot product needs three shuffle
// stages (the earlier lane = channel mapping walked the heads one after the other with five-stage reductions: 320 dependent
// shuffles per pixel, 2 TB/s).
// ------------------------------------------------------------------------------------------------
__global__ void proposal_attention_kernel(const float* __restrict__ qkv, int P, int K, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p >= P) return;
  const float* base = qkv + (size_t)p * K * kQkv + lane * 4;        // head lane / 8, channels 4 (lane % 8) .. + 3
  float4 q[kMaxK], k[kMaxK], v[kMaxK];
#pragma unroll
  for (int i = 0; i < kMaxK; ++i) {
    if (i < K) {
      q[i] = *reinterpret_cast<const float4*>(base + i * kQkv);
      k[i] = *reinterpret_cast<const float4*>(base + i * kQkv + 128);
      v[i] = *reinterpret_cast<const float4*>(base + i * kQkv + 256);
    } else { q[i] = k[i] = v[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
  }
#pragma unroll
  for (int i = 0; i < kMaxK; ++i) {
    if (i >= K) break;
    float s[kMaxK];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) {
      if (j < K) {
        float d = (q[i].x * k[j].x + q[i].y * k[j].y) + (q[i].z * k[j].z + q[i].w * k[j].w);
        d += __shfl_xor_sync(0xffffffffu, d, 1); d += __shfl_xor_sync(0xffffffffu, d, 2); d += __shfl_xor_sync(0xffffffffu, d, 4);
        s[j] = d * kScale;
        m = fmaxf(m, s[j]);
      }
    }
    float sum = 0.f;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) {
      if (j < K) {
        const float e = expf(s[j] - m);
        sum += e;
        o.x = fmaf(e, v[j].x, o.x); o.y = fmaf(e, v[j].y, o.y); o.z = fmaf(e, v[j].z, o.z); o.w = fmaf(e, v[j].w, o.w);
      }
    }
    *reinterpret_cast<float4*>(out + ((size_t)p * K + i) * kEmbed + lane * 4) = make_float4(o.x / sum, o.y / sum, o.z / sum, o.w / sum);
  }
}

// ------------------------------------------------------------------------------------------------
// A11: window attention.  One CTA = one head of `wpc` windows processed TOGETHER (the head's slice of
// the relative-position table is staged once per CTA; small windows fill the CTA).  Token slot
// t=(pl,n), pl = ly*ws+lx.
//   logits[i,j] = s q_i.k_j + s q_i.Rk[rel(pi,pj)] + s k_j.Rq[rel(pi,pj)] + mask     (NMP.py:263-275)
//   out_i       = sum_j A_ij v_j + sum_pj (sum_n A_i,(pj,n)) Rv[rel(pi,pj)]          (NMP.py:282)
// The two RPE logit terms are evaluated per (token, pixel) pair once (QR, KR: T x P tables)
// instead of per (token, token) pair; the Rv term uses the per-pixel bucket sums of A.
// Register blocking: a warp owns R query rows at a time and each lane NCH key columns, so the QK^T
// inner loop issues R*NCH independent FMAs per (R/2 + NCH) shared-memory loads, and the A.V loop
// R FMAs per (1 + R/2) loads.
// ------------------------------------------------------------------------------------------------
struct WinParams {
  const float* qkv; const float* table; float* out;
  int B, Hp, Wp, K, ws, shift, self_edge, nwy, nwx, nwin, wpc;
};

constexpr int WIN_THREADS = 256;
constexpr int WIN_WARPS = WIN_THREADS / 32;

template <int R, int NCH>
__global__ void __launch_bounds__(WIN_THREADS) window_attention_kernel(const WinParams p) {
  extern __shared__ __align__(16) float smem[];
  const int ws = p.ws, K = p.K;
  const int P = ws * ws, Tw = P * K, NR = (2 * ws - 1) * (2 * ws - 1);
  const int TT = p.wpc * Tw;           // token slots of all windows of this CTA
  const int TP = TT + 2;               // even row stride: 8-byte aligned float2 broadcasts, conflict-free columns
  constexpr int RS = 33;               // table row stride: rows are read by lanes at a fixed d -> odd stride, no bank conflicts
  float* sRq = smem;                   // [NR][33]  (pre-scaled)
  float* sRk = sRq + NR * RS;          // [NR][33]
  float* sRv = sRk + NR * RS;          // [NR][33]
  float* qT = sRv + NR * RS + (NR & 1);   // [32][TP]  (pre-scaled); keep 8-byte alignment
  float* kT = qT + 32 * TP;            // [32][TP]
  float* vs = kT + 32 * TP;            // [TT][32]
  float* QR = vs + TT * 32;            // [TT][P]
  float* KR = QR + TT * P;             // [TT][P]
  float* pbuf = KR + TT * P;           // [WARPS][Tw][R]   un-normalised probabilities, transposed
  float* abuf = pbuf + WIN_WARPS * Tw * R;   // [WARPS][P][R]  per-pixel bucket sums
  int* tok_row = reinterpret_cast<int*>(abuf + WIN_WARPS * P * R);   // [TT] global token row
  int* reg = tok_row + TT;             // [wpc*P] Swin region id (rolled coordinates)
  int* pixof = reg + p.wpc * P;        // [Tw]   pixel of a window-local token slot (t / K) -- no divisions in the hot loops
  int* relof = pixof + Tw;             // [P*P]  relative-position row (x33) of the pixel pair (pi, pj)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;
  const int w_begin = blockIdx.x * p.wpc;
  const int nw = min(p.wpc, p.nwin - w_begin);      // windows actually present

  for (int i = tid; i < NR * 32; i += WIN_THREADS) {
    const int r = i >> 5, d = i & 31;
    const float* row = p.table + (size_t)r * kQkv + head * 96;
    sRq[r * RS + d] = row[d] * kScale;
    sRk[r * RS + d] = row[32 + d];
    sRv[r * RS + d] = row[64 + d];
  }
  for (int t = tid; t < Tw; t += WIN_THREADS) pixof[t] = t / K;
  for (int i = tid; i < P * P; i += WIN_THREADS) {
    const int pi = i / P, pj = i % P;
    relof[i] = ((pi / ws - pj / ws + ws - 1) * (2 * ws - 1) + (pi % ws - pj % ws + ws - 1)) * RS;
  }
  for (int t = tid; t < TT; t += WIN_THREADS) {
    const int wl = t / Tw, tl = t % Tw;
    const int win = min(w_begin + wl, p.nwin - 1);   // slots of absent windows alias the last one (never stored)
    const int b = win / (p.nwy * p.nwx);
    const int wy = (win / p.nwx) % p.nwy, wx = win % p.nwx;
    const int pl = tl / K, n = tl % K;
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
      reg[wl * P + pl] = r;
    }
  }
  __syncthreads();
  for (int i = tid; i < TT * 8; i += WIN_THREADS) {                    // one float4 of q, k, v per thread
    const int t = i >> 3, c = (i & 7) * 4;
    const float* src = p.qkv + (size_t)tok_row[t] * kQkv + head * 32 + c;
    const float4 q = *reinterpret_cast<const float4*>(src);
    const float4 k = *reinterpret_cast<const float4*>(src + 128);
    const float4 v = *reinterpret_cast<const float4*>(src + 256);
    qT[(c + 0) * TP + t] = q.x * kScale; qT[(c + 1) * TP + t] = q.y * kScale;
    qT[(c + 2) * TP + t] = q.z * kScale; qT[(c + 3) * TP + t] = q.w * kScale;
    kT[(c + 0) * TP + t] = k.x; kT[(c + 1) * TP + t] = k.y;
    kT[(c + 2) * TP + t] = k.z; kT[(c + 3) * TP + t] = k.w;
    *reinterpret_cast<float4*>(vs + t * 32 + c) = v;
  }
  __syncthreads();
  // QR[t][pp] = (s q_t).Rk[rel(p_t,pp)]     KR[t][pp] = k_t.(s Rq[rel(pp,p_t)])
  for (int i = tid; i < TT * P; i += WIN_THREADS) {
    const int t = i / P, pp = i - t * P;
    const int pt = pixof[t % Tw];
    const float* rk = sRk + relof[pt * P + pp];
    const float* rq = sRq + relof[pp * P + pt];
    float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll 8
    for (int d = 0; d < 32; d += 2) {
      a0 = fmaf(qT[d * TP + t], rk[d], a0);
      a1 = fmaf(qT[(d + 1) * TP + t], rk[d + 1], a1);
      c0 = fmaf(kT[d * TP + t], rq[d], c0);
      c1 = fmaf(kT[(d + 1) * TP + t], rq[d + 1], c1);
    }
    QR[i] = a0 + a1;
    KR[i] = c0 + c1;
  }
  __syncthreads();

  float* myp = pbuf + warp * Tw * R;
  float* myab = abuf + warp * P * R;
  const int groups_per_win = Tw / R;
  const int ngroups = nw * groups_per_win;
  for (int g = warp; g < ngroups; g += WIN_WARPS) {
    const int wl = g / groups_per_win;
    const int i0 = wl * Tw + (g % groups_per_win) * R;        // first query row (CTA slot index)
    const int c0 = wl * Tw;                                    // first key column of this window
    const int* wreg = reg + wl * P;
    float acc[R][NCH];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NCH; ++c) acc[r][c] = 0.f;
#pragma unroll 4
    for (int d = 0; d < 32; ++d) {
      float qv[R], kv[NCH];
      if constexpr (R % 2 == 0) {
#pragma unroll
        for (int r = 0; r < R; r += 2) {
          const float2 t2 = *reinterpret_cast<const float2*>(qT + d * TP + i0 + r);
          qv[r] = t2.x; qv[r + 1] = t2.y;
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) qv[r] = qT[d * TP + i0 + r];
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int j = c * 32 + lane;
        kv[c] = (j < Tw) ? kT[d * TP + c0 + j] : 0.f;
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc[r][c] = fmaf(qv[r], kv[c], acc[r][c]);
    }
    // RPE terms, masks, softmax (row-wise over the window's Tw columns)
    float inv[R];
#pragma unroll
    int pjc[NCH];                      // pixel of this lane's key columns
#pragma unroll
    for (int c = 0; c < NCH; ++c) pjc[c] = pixof[min(c * 32 + lane, Tw - 1)];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = i0 + r, il = i - c0, pi = pixof[il];
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int j = c * 32 + lane;
        float s = -INFINITY;
        if (j < Tw) {
          const int pj = pjc[c];
          s = acc[r][c] + QR[i * P + pj] + KR[(c0 + j) * P + pi];
          const bool masked = (wreg[pi] != wreg[pj]) || (p.self_edge && pi == pj && il != j);
          if (masked) s = -INFINITY;
        }
        acc[r][c] = s;
        m = fmaxf(m, s);
      }
      m = warp_max(m);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int j = c * 32 + lane;
        if (j < Tw) {
          const float e = expf(acc[r][c] - m);
          myp[j * R + r] = e;
          sum += e;
        }
      }
      inv[r] = 1.f / warp_sum(sum);
    }
    __syncwarp();
    for (int pp = lane; pp < P; pp += 32) {          // per-pixel bucket sums of the un-normalised A
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float a = 0.f;
        for (int n = 0; n < K; ++n) a += myp[(pp * K + n) * R + r];
        myab[pp * R + r] = a;
      }
    }
    __syncwarp();
    float o[R];
#pragma unroll
    for (int r = 0; r < R; ++r) o[r] = 0.f;
#pragma unroll 4
    for (int j = 0; j < Tw; ++j) {
      const float v = vs[(c0 + j) * 32 + lane];
      if constexpr (R % 2 == 0) {
#pragma unroll
        for (int r = 0; r < R; r += 2) {
          const float2 p2 = *reinterpret_cast<const float2*>(myp + j * R + r);
          o[r] = fmaf(p2.x, v, o[r]); o[r + 1] = fmaf(p2.y, v, o[r + 1]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = fmaf(myp[j * R + r], v, o[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int* rel = relof + pixof[i0 + r - c0] * P;
      float acc_rv = o[r];
#pragma unroll 4
      for (int pp = 0; pp < P; ++pp) acc_rv = fmaf(myab[pp * R + r], sRv[rel[pp] + lane], acc_rv);
      p.out[(size_t)tok_row[i0 + r] * kEmbed + head * 32 + lane] = acc_rv * inv[r];
    }
    __syncwarp();
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
  int pix_i[ST_RPW];                   // pixel (along the stripe) of this warp's query rows
#pragma unroll
  for (int r = 0; r < ST_RPW; ++r) {
    m[r] = -INFINITY; l[r] = 0.f; acc[r] = 0.f;
    pix_i[r] = (q0 + warp * ST_RPW + r) / K;
  }

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
    const int pix_j0 = (t0 + lane) / K, pix_j1 = (t0 + lane + 32) / K;   // once per tile, not per row
#pragma unroll
    for (int r = 0; r < ST_RPW; ++r) {
      const int ti = q0 + warp * ST_RPW + r;
      float tmax = -INFINITY;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int tj = t0 + lane + e * 32;
        const bool masked = (tj >= Lk) || ((e ? pix_j1 : pix_j0) == pix_i[r] && tj != ti);   // NMP.py:203-208
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

template <int R, int NCH>
static int launch_window(const WinParams& p, size_t smem, dim3 grid, cudaStream_t stream) {
  static PerDevice configured;
  ensure_dynamic_smem(window_attention_kernel<R, NCH>, (int)smem, configured);
  window_attention_kernel<R, NCH><<<grid, WIN_THREADS, smem, stream>>>(p);
  count_launch();
  return check_launch("window_attention");
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
  const int P = ws * ws, Tw = P * K, NR = (2 * ws - 1) * (2 * ws - 1);
  NMRF_REQUIRE(Tw <= 256, "window_attention: %d tokens per window exceed 256", Tw);
  const int R = (Tw % 6 == 0) ? 6 : (Tw % 4 == 0) ? 4 : (Tw % 2 == 0) ? 2 : 1;
  const int nch = (Tw + 31) / 32;
  // small windows: several windows per CTA so the CTA's 8 warps all have row groups
  int wpc = 1;
  while (wpc < 8 && (wpc * Tw) / R < 2 * WIN_WARPS && 2 * wpc * Tw <= 256) wpc *= 2;
  p.wpc = wpc;
  const size_t TT = (size_t)wpc * Tw;
  const size_t smem = sizeof(float) * ((size_t)3 * NR * 33 + (NR & 1) + (size_t)2 * 32 * (TT + 2) + TT * 32 + 2 * TT * P +
                                       (size_t)WIN_WARPS * (Tw + P) * R) + sizeof(int) * (TT + (size_t)wpc * P + Tw + (size_t)P * P);
  NMRF_REQUIRE(smem <= 227 * 1024, "window_attention: ws=%d K=%d needs %zu B of shared memory", ws, K, smem);
  dim3 grid((p.nwin + wpc - 1) / wpc, kHeads);
#define NMRF_WIN(RR, NN) if (R == RR && nch <= NN) return launch_window<RR, NN>(p, smem, grid, stream)
  NMRF_WIN(6, 2); NMRF_WIN(6, 3); NMRF_WIN(6, 4); NMRF_WIN(6, 5);
  NMRF_WIN(4, 1); NMRF_WIN(4, 2); NMRF_WIN(4, 4);
  NMRF_WIN(2, 1); NMRF_WIN(2, 2); NMRF_WIN(2, 4); NMRF_WIN(2, 8);
  NMRF_WIN(6, 8); NMRF_WIN(4, 8);
#undef NMRF_WIN
  return launch_window<1, 8>(p, smem, grid, stream);
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
