// A11 window attention on the warp-level tensor path (mma.sync m16n8k8, 3xTF32), registers only.
//
//   logits[i,j] = s q_i.k_j + s q_i.Rk[rel(pi,pj)] + s k_j.Rq[rel(pi,pj)] + mask     (reference NMP.py:263-275)
//   out_i       = sum_j A_ij v_j + sum_pj (sum_n A_i,(pj,n)) Rv[rel(pi,pj)]          (reference NMP.py:282)
//
// Why mma.sync and not tcgen05 here: a window has Tw = 144 tokens (ws=6, K=4) -- 128+16 for an M=128 UMMA --, the
// softmax needs two gathered per-pixel terms added to every logit (which wants the logits in registers, next to the
// index arithmetic), and S + P_hi + P_lo of one window already fill the 512 TMEM columns, which pins one CTA of five
// busy warps per SM.  The tensor work is small (0.6k mma per 16-row block); what the SIMT kernel (attention.cu) lost its
// time on was the fp32 FMA inner loops at 12% occupancy.  With m16n8k8 fragments:
//   * one warp owns a 16-row block of a window for all phases, two CTAs (18 warps) per SM;
//   * token slots are ordered so that a 16-row block is a BH x BW pixel patch (2x2 pixels x 4 proposals, or the whole
//     4x4 window at K=1): the relative positions a block needs from the RPE table are a 7x7 sub-block (49 rows),
//     so the two RPE logit tables are  Q_blk . Rk_sub^T  and  K_blk . Rq_sub^T  (7 n-tiles each, instead of the
//     gathered 36-wide dot products of the SIMT kernel), scattered from the C fragments into [token][pixel] tables;
//   * S = Q_blk . K^T stays in C fragments (18 n-tiles x 4 registers); masks, the two gathered RPE terms, max, exp2 and
//     the row sum are done in place; a quad shuffle gives the per-pixel bucket sums of A;
//   * the C fragment of P is reused as the A fragment of P.V with the k index permuted (key 2t, 2t+1 <-> fragment
//     column t, t+4) and V's B fragments loaded with the same permutation: no shuffles, no shared-memory round trip;
//   * the Rv term is one more small mma (bucket sums gathered into a [16 x 49] A operand, Rv_sub as B) accumulated
//     into the same O fragments.
// 3xTF32: every product is  a_lo.b_hi + a_hi.b_lo + a_hi.b_hi  with round-to-nearest splits (see tc_common.cuh).
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace nmrf {
namespace {
using tc::rna_tf32_fast;

constexpr float kScaleM = 0.17677669529663687f;   // 32^-0.5 (NMP.py:79,163,412)
constexpr int LD = 36;                             // row stride of the [rows][32] operand tiles: conflict-free fragment loads

struct WinMmaParams {
  const float* qkv; const float* table; float* out;
  int B, Hp, Wp, shift, self_edge, nwy, nwx, nwin;
};

// hi = x rounded to nearest tf32; lo = x - hi (exact) rounded to nearest tf32 as well: mma.sync reads only the upper 19 bits
// of an operand, i.e. it would TRUNCATE an unrounded lo (tc_common.cuh: lo_tf32)
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  const float h = rna_tf32_fast(x);
  hi = __float_as_uint(h);
  lo = __float_as_uint(tc::lo_tf32(x, h));
}
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a . b with both operands given in fp32: small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0, float b1) {
  uint32_t b0h, b0l, b1h, b1l;
  split(b0, b0h, b0l); split(b1, b1h, b1l);
  mma8(c, al, b0h, b1h);
  mma8(c, ah, b0l, b1l);
  mma8(c, ah, b0h, b1h);
}

// Long accumulations (P.V over 18 + 7 k-steps): mma.sync updates its accumulator with round-toward-zero like tcgen05 (relative
// bias -1.6e-8 per mma into the same registers, tools/precision_probe.py), so the kernel keeps the small lo.hi / hi.lo
// products of all k-steps in a separate accumulator (where truncation is relative to THEIR small magnitude) and adds it to
// the hi.hi accumulator once.

template <int WS, int K>
struct Geo {
  static constexpr int P = WS * WS, Tw = P * K, NB = Tw / 16, NT = Tw / 8;
  static constexpr int PPB = 16 / K;                         // pixels of a 16-row block
  static constexpr int BW = (PPB == 4) ? 2 : 4, BH = PPB / BW;   // K=4: 2x2 patch; K=1: 4x4
  static constexpr int NA = WS + BH - 1, NBX = WS + BW - 1;   // relative-position sub-block (rows x cols)
  static constexpr int NSUB = NA * NBX, NSUBT = (NSUB + 7) / 8;
  static constexpr int NRW = 2 * WS - 1, NR = NRW * NRW;
  static constexpr int PS = P + 1;                            // row stride of the [token][pixel] tables
  static_assert(Tw % 16 == 0 && WS % BH == 0 && WS % BW == 0 && 16 % K == 0 && (PPB == 4 || PPB == 16), "unsupported window geometry");
  // block -> pixel origin; slot (window-local token index in patch order) -> raster pixel
  __device__ static __forceinline__ int by(int blk) { return (blk / (WS / BW)) * BH; }
  __device__ static __forceinline__ int bx(int blk) { return (blk % (WS / BW)) * BW; }
  __device__ static __forceinline__ int pix_of_slot(int s) {
    const int blk = s >> 4, rp = (s & 15) / K;
    return (by(blk) + rp / BW) * WS + bx(blk) + rp % BW;
  }
};

template <int WS, int K, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * Geo<WS, K>::NB * 32, MINB)
window_attention_mma_kernel(const WinMmaParams p) {
  using G = Geo<WS, K>;
  constexpr int P = G::P, Tw = G::Tw, NB = G::NB, NT = G::NT, PS = G::PS, NR = G::NR, NRW = G::NRW;
  constexpr int BW = G::BW, BH = G::BH, NBX = G::NBX, NSUB = G::NSUB, NSUBT = G::NSUBT;
  constexpr int NWARP = WPC * NB, NTHREADS = NWARP * 32;
  constexpr int TAB = 2 * NR * LD, VSZ = WPC * Tw * LD;
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;                                  // [WPC*Tw][LD]
  float* KRs = Ks + WPC * Tw * LD;                   // [WPC*Tw][PS]   KR[j][pi] = k_j . s Rq[rel(pi,pj)]
  float* QRab = KRs + WPC * Tw * PS;                 // [NWARP][16][PS] QR[i][pj] = s q_i . Rk[rel(pi,pj)], then bucket sums of A
  float* uni = QRab + NWARP * 16 * PS;               // phase 1: Rk [NR][LD], Rq [NR][LD] (pre-scaled); afterwards V [WPC*Tw][LD]
  float* Rk = uni;
  float* Rq = uni + NR * LD;
  float* Vs = uni;
  int* tok_row = reinterpret_cast<int*>(uni + (TAB > VSZ ? TAB : VSZ));   // [WPC*Tw] global token row of a slot
  int* reg = tok_row + WPC * Tw;                     // [WPC*P]  Swin region id of a pixel (rolled coordinates)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y;
  const int w_begin = blockIdx.x * WPC;
  const int wl = warp / NB, blk = warp % NB;
  const bool live = w_begin + wl < p.nwin;           // windows past the end alias the last one and store nothing

  // ---- slot -> token row, pixel -> region ----
  for (int s = tid; s < WPC * Tw; s += NTHREADS) {
    const int wi = s / Tw, sl = s % Tw;
    const int win = min(w_begin + wi, p.nwin - 1);
    const int b = win / (p.nwy * p.nwx);
    const int wy = (win / p.nwx) % p.nwy, wx = win % p.nwx;
    const int pl = G::pix_of_slot(sl), n = sl % K;
    const int yr = wy * WS + pl / WS, xr = wx * WS + pl % WS;           // rolled coordinates
    const int y = (yr + p.shift) % p.Hp, x = (xr + p.shift) % p.Wp;     // NMP.py:249-250
    tok_row[s] = ((b * p.Hp + y) * p.Wp + x) * K + n;
    if (n == 0) {
      int r = 0;
      if (p.shift > 0) {                                                // NMP.py:221-232
        const int ry = (yr >= p.Hp - WS) + (yr >= p.Hp - p.shift);
        const int rx = (xr >= p.Wp - WS) + (xr >= p.Wp - p.shift);
        r = ry * 3 + rx;
      }
      reg[wi * P + pl] = r;
    }
  }
  for (int i = tid; i < NR * 8; i += NTHREADS) {                        // RPE table slices of this head
    const int r = i >> 3, c = (i & 7) * 4;
    const float* row = p.table + (size_t)r * kQkv + head * 96 + c;
    float4 q4 = *reinterpret_cast<const float4*>(row);
    const float4 k4 = *reinterpret_cast<const float4*>(row + 32);
    q4.x *= kScaleM; q4.y *= kScaleM; q4.z *= kScaleM; q4.w *= kScaleM;
    *reinterpret_cast<float4*>(Rq + r * LD + c) = q4;
    *reinterpret_cast<float4*>(Rk + r * LD + c) = k4;
  }
  __syncthreads();
  for (int i = tid; i < WPC * Tw * 8; i += NTHREADS) {                  // K rows
    const int s = i >> 3, c = (i & 7) * 4;
    *reinterpret_cast<float4*>(Ks + s * LD + c) =
        *reinterpret_cast<const float4*>(p.qkv + (size_t)tok_row[s] * kQkv + 128 + head * 32 + c);
  }
  // this warp's 16 query rows as A fragments (pre-scaled): q[ks*4 + {0: (g, t), 1: (g+8, t), 2: (g, t+4), 3: (g+8, t+4)}]
  const int slot0 = wl * Tw + blk * 16;
  float q[16];
  {
    const float* q0 = p.qkv + (size_t)tok_row[slot0 + g] * kQkv + head * 32;
    const float* q1 = p.qkv + (size_t)tok_row[slot0 + g + 8] * kQkv + head * 32;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      q[ks * 4 + 0] = q0[ks * 8 + t] * kScaleM;     q[ks * 4 + 1] = q1[ks * 8 + t] * kScaleM;
      q[ks * 4 + 2] = q0[ks * 8 + t + 4] * kScaleM; q[ks * 4 + 3] = q1[ks * 8 + t + 4] * kScaleM;
    }
  }
  __syncthreads();

  const int by = G::by(blk), bx = G::bx(blk);
  float* myqr = QRab + warp * 16 * PS;
  // ---- phase 1: the two RPE logit tables of this block ----
  {
    float cq[NSUBT][4], ck[NSUBT][4];
#pragma unroll
    for (int nt = 0; nt < NSUBT; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) { cq[nt][c] = 0.f; ck[nt][c] = 0.f; }
    // B rows: sub-block column n = a * NBX + b  ->  table row  (oy + a) * NRW + ox + b   (columns past NSUB: clamped, unused)
    int rq_row[NSUBT], rk_row[NSUBT];
#pragma unroll
    for (int nt = 0; nt < NSUBT; ++nt) {
      const int n = min(nt * 8 + g, NSUB - 1), a = n / NBX, b = n % NBX;
      rk_row[nt] = ((by + a) * NRW + bx + b) * LD;                               // QR: rel(pi, pp), pi in the block
      rq_row[nt] = ((WS - BH - by + a) * NRW + (WS - BW - bx + b)) * LD;        // KR: rel(pi, pj), pj in the block
    }
    const float* krow0 = Ks + (slot0 + g) * LD;
    const float* krow1 = Ks + (slot0 + g + 8) * LD;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t qh[4], ql[4], kh[4], kl[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) split(q[ks * 4 + c], qh[c], ql[c]);
      split(krow0[ks * 8 + t], kh[0], kl[0]);     split(krow1[ks * 8 + t], kh[1], kl[1]);
      split(krow0[ks * 8 + t + 4], kh[2], kl[2]); split(krow1[ks * 8 + t + 4], kh[3], kl[3]);
#pragma unroll
      for (int nt = 0; nt < NSUBT; ++nt) {
        mma3(cq[nt], qh, ql, Rk[rk_row[nt] + ks * 8 + t], Rk[rk_row[nt] + ks * 8 + t + 4]);
        mma3(ck[nt], kh, kl, Rq[rq_row[nt] + ks * 8 + t], Rq[rq_row[nt] + ks * 8 + t + 4]);
      }
    }
    // scatter: C element (row, n) -> pixel of the window, if it exists.  n = 8 nt + 2t + (c & 1) -> (a, b) = (n / NBX, n % NBX),
    // advanced incrementally over nt
    int a_[2], b_[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) { a_[e] = (2 * t + e) / NBX; b_[e] = (2 * t + e) % NBX; }
    const int ry_[2] = {(g / K) / BW, ((g + 8) / K) / BW}, rx_[2] = {(g / K) % BW, ((g + 8) / K) % BW};
#pragma unroll
    for (int nt = 0; nt < NSUBT; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int row = g + 8 * (c >> 1), a = a_[c & 1], b = b_[c & 1], ry = ry_[c >> 1], rx = rx_[c >> 1];
        if (a < G::NA) {
          const int yp = ry + WS - 1 - a, xp = rx + WS - 1 - b;                  // QR[i][pp]
          if ((unsigned)yp < (unsigned)WS && (unsigned)xp < (unsigned)WS) myqr[row * PS + yp * WS + xp] = cq[nt][c];
          const int yi = a - BH + 1 + ry, xi = b - BW + 1 + rx;                  // KR[j][pi]
          if ((unsigned)yi < (unsigned)WS && (unsigned)xi < (unsigned)WS) KRs[(slot0 + row) * PS + yi * WS + xi] = ck[nt][c];
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {          // n += 8
        b_[e] += 8 % NBX; a_[e] += 8 / NBX;
        if (b_[e] >= NBX) { b_[e] -= NBX; ++a_[e]; }
      }
    }
  }
  __syncthreads();                         // KR of every block is complete; Rk / Rq are dead
  for (int i = tid; i < WPC * Tw * 8; i += NTHREADS) {                  // V rows into the table's space, asynchronously
    const int s = i >> 3, c = (i & 7) * 4;
    const float* src = p.qkv + (size_t)tok_row[s] * kQkv + 256 + head * 32 + c;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(Vs + s * LD + c)), "l"(src));
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // ---- phase 2: S = Q_blk . K^T in C fragments ----
  float sc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) sc[nt][c] = 0.f;
  {
    const float* kb = Ks + (wl * Tw + g) * LD + t;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t qh[4], ql[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) split(q[ks * 4 + c], qh[c], ql[c]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma3(sc[nt], qh, ql, kb[nt * 8 * LD + ks * 8], kb[nt * 8 * LD + ks * 8 + 4]);
    }
  }
  // ---- RPE terms, masks, softmax (rows g and g+8 of the block; this thread: key columns 8 nt + 2t, +1) ----
  const int* wreg = reg + wl * P;
  const int il0 = blk * 16 + g, il1 = il0 + 8;                       // window-local query slots
  const int pi0 = G::pix_of_slot(il0), pi1 = G::pix_of_slot(il1);
  const int ri0 = wreg[pi0], ri1 = wreg[pi1];
  const float* kr = KRs + wl * Tw * PS;
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = nt * 8 + 2 * t + (c & 1);
      const int pj = G::pix_of_slot(j);
      const int row = g + 8 * (c >> 1), il = (c >> 1) ? il1 : il0, pi = (c >> 1) ? pi1 : pi0, ri = (c >> 1) ? ri1 : ri0;
      float s = sc[nt][c] + myqr[row * PS + pj] + kr[j * PS + pi];
      const bool masked = (ri != wreg[pj]) || (p.self_edge && pi == pj && il != j);
      if (masked) s = -INFINITY;
      sc[nt][c] = s;
      if (c >> 1) m1 = fmaxf(m1, s); else m0 = fmaxf(m0, s);
    }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  __syncwarp();                            // every lane has read its QR values: the buffer now takes the bucket sums
  float sum0 = 0.f, sum1 = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float e[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float x = (sc[nt][c] - ((c >> 1) ? m1 : m0)) * kLog2e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[c]) : "f"(x));
      sc[nt][c] = e[c];
    }
    sum0 += e[0] + e[1]; sum1 += e[2] + e[3];
    // per-pixel bucket sums of the un-normalised A (NMP.py:282): the K proposals of a pixel are adjacent key slots
    if constexpr (K == 1) {
      myqr[g * PS + G::pix_of_slot(nt * 8 + 2 * t)] = e[0];       myqr[g * PS + G::pix_of_slot(nt * 8 + 2 * t + 1)] = e[1];
      myqr[(g + 8) * PS + G::pix_of_slot(nt * 8 + 2 * t)] = e[2]; myqr[(g + 8) * PS + G::pix_of_slot(nt * 8 + 2 * t + 1)] = e[3];
    } else {
      float b0 = e[0] + e[1], b1 = e[2] + e[3];
#pragma unroll
      for (int o = 1; o < K / 2; o <<= 1) { b0 += __shfl_xor_sync(0xffffffffu, b0, o); b1 += __shfl_xor_sync(0xffffffffu, b1, o); }
      if ((t & (K / 2 - 1)) == 0) {
        const int pj = G::pix_of_slot(nt * 8 + 2 * t);
        myqr[g * PS + pj] = b0; myqr[(g + 8) * PS + pj] = b1;
      }
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

  // ---- phase 3: O = P . V  (+ bucket sums . Rv_sub) ----
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();                         // V has landed for everyone (and the bucket sums are visible warp-wide)
  float oc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) oc[nt][c] = 0.f;
  // P . V  and  (bucket sums) . Rv_sub  over 18 + 7 k-steps: the small lo.hi / hi.lo products go to their OWN accumulator (os),
  // the hi.hi products to oc, and the two are added once at the end (see mma_small)
  float os[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) os[nt][c] = 0.f;
  {
    // A fragment = C fragment of P with the k index permuted: fragment column t <-> key 2t, column t+4 <-> key 2t+1
    const float* vb = Vs + (wl * Tw + 2 * t) * LD + g;
#pragma unroll
    for (int ks = 0; ks < NT; ++ks) {
      uint32_t ph[4], pl[4];
      split(sc[ks][0], ph[0], pl[0]); split(sc[ks][2], ph[1], pl[1]);
      split(sc[ks][1], ph[2], pl[2]); split(sc[ks][3], ph[3], pl[3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t b0h, b0l, b1h, b1l;
        split(vb[ks * 8 * LD + nt * 8], b0h, b0l); split(vb[ks * 8 * LD + LD + nt * 8], b1h, b1l);
        mma8(os[nt], pl, b0h, b1h);
        mma8(os[nt], ph, b0l, b1l);
        mma8(oc[nt], ph, b0h, b1h);
      }
    }
  }
  {
    // A'[row][n] = bucket sum of the pixel at relative position n of the sub-block (0 outside the window); B = Rv_sub
    const float* rv = p.table + head * 96 + 64;
#pragma unroll
    for (int ks = 0; ks < NSUBT; ++ks) {
      float av[4];
      int rrow[2];
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        const int n = ks * 8 + t + 4 * hcol, nc = min(n, NSUB - 1), a = nc / NBX, b = nc % NBX;
        rrow[hcol] = (by + a) * NRW + bx + b;
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int row = g + 8 * hrow, rp = row / K, yp = rp / BW + WS - 1 - a, xp = rp % BW + WS - 1 - b;
          const bool ok = n < NSUB && yp >= 0 && yp < WS && xp >= 0 && xp < WS;
          av[hcol * 2 + hrow] = ok ? myqr[row * PS + yp * WS + xp] : 0.f;
        }
      }
      uint32_t ah[4], al[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) split(av[c], ah[c], al[c]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t b0h, b0l, b1h, b1l;
        split(__ldg(rv + (size_t)rrow[0] * kQkv + nt * 8 + g), b0h, b0l); split(__ldg(rv + (size_t)rrow[1] * kQkv + nt * 8 + g), b1h, b1l);
        mma8(os[nt], al, b0h, b1h);
        mma8(os[nt], ah, b0l, b1l);
        mma8(oc[nt], ah, b0h, b1h);
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) oc[nt][c] += os[nt][c];
  if (live) {
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    float* o0 = p.out + (size_t)tok_row[slot0 + g] * kEmbed + head * 32 + 2 * t;
    float* o1 = p.out + (size_t)tok_row[slot0 + g + 8] * kEmbed + head * 32 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<float2*>(o0 + nt * 8) = make_float2(oc[nt][0] * inv0, oc[nt][1] * inv0);
      *reinterpret_cast<float2*>(o1 + nt * 8) = make_float2(oc[nt][2] * inv1, oc[nt][3] * inv1);
    }
  }
}

template <int WS, int K, int WPC, int MINB>
int launch_mma(const WinMmaParams& p, cudaStream_t stream) {
  using G = Geo<WS, K>;
  constexpr int NWARP = WPC * G::NB;
  constexpr int TAB = 2 * G::NR * LD, VSZ = WPC * G::Tw * LD;
  constexpr size_t smem = sizeof(float) * ((size_t)WPC * G::Tw * LD + (size_t)WPC * G::Tw * G::PS + (size_t)NWARP * 16 * G::PS +
                                           (TAB > VSZ ? TAB : VSZ)) + sizeof(int) * (WPC * G::Tw + WPC * G::P);
  static PerDevice configured;
  ensure_dynamic_smem(window_attention_mma_kernel<WS, K, WPC, MINB>, (int)smem, configured);
  dim3 grid((p.nwin + WPC - 1) / WPC, kHeads);
  window_attention_mma_kernel<WS, K, WPC, MINB><<<grid, NWARP * 32, smem, stream>>>(p);
  count_launch();
  return check_launch("window_attention_mma");
}

}  // namespace

bool window_attention_mma_supported(int K, int ws) { return (ws == 6 && K == 4) || (ws == 4 && K == 1); }

int window_attention_mma(const float* qkv, const float* table, int B, int Hp, int Wp, int K, int ws, int shift,
                         int self_edge, float* out, cudaStream_t stream) {
  WinMmaParams p;
  p.qkv = qkv; p.table = table; p.out = out;
  p.B = B; p.Hp = Hp; p.Wp = Wp; p.shift = shift; p.self_edge = self_edge;
  p.nwy = Hp / ws; p.nwx = Wp / ws; p.nwin = B * p.nwy * p.nwx;
  static int minb = 0;
  if (!minb) { const char* e = getenv("NMRF_B200_WIN_MINB"); minb = e ? atoi(e) : 2; }
  if (ws == 6 && K == 4) return minb == 1 ? launch_mma<6, 4, 1, 1>(p, stream) : launch_mma<6, 4, 1, 2>(p, stream);
  if (ws == 4 && K == 1) return minb == 1 ? launch_mma<4, 1, 8, 1>(p, stream) : launch_mma<4, 1, 8, 2>(p, stream);
  set_error("window_attention_mma: unsupported geometry ws=%d K=%d", ws, K);
  return NMRF_ERR_BAD_ARG;
}

}  // namespace nmrf
