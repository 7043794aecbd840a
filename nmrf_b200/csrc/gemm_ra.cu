// Fused token GEMM on tcgen05 with the A operand RESIDENT in tensor memory for a whole row block ("ra").
//
// Why (measured on gemm_tc6.cu, tools/gemm_trace.py + profiles/r2_ncu_token_gemm_qkv.txt): the 128x128-tile kernel
// re-produces the LayerNorm'ed, hi/lo-split A tile once per 128-column chunk (three times for a qkv projection), issues
// 5.3k warp-instructions per 32-wide k-block (43 % of all issue slots; tensor pipe 28 %) and hands every tile to the epilogue
// through ~10k cycles of drain + store.  For K <= 192 -- every nn.Linear of the hot path -- the whole A row block fits in
// tensor memory: K/32 k-blocks x (32 hi + 32 lo columns) <= 384 of the 512 columns.  So:
//
//   row block  128 rows; the producers write A (hi | lo, all k-blocks) into TMEM ONCE, then every 64-column chunk of the
//              output is  A . W_chunk^T  against weight tiles streamed through an 8-slot ring (16 KB per k-block: the
//              [64 n x 32 k] hi and lo halves of the tile images of nmrf_pack_weight_tiles).
//   accuracy   as gemm_tc6: groups of <= 3 k-blocks get a fresh 64-column accumulator stage (ring of (512 - 64 K/32) / 64
//              stages, at most 4), inside a k-block the small lo.hi / hi.lo products go first, the epilogue adds a chunk's
//              groups in fp32 registers (the tensor core's accumulator update rounds toward zero).
//   warps      0-7 producers (cp.async raw ring -> LayerNorm from nmrf_gemm_args.ln_stats -> hi/lo -> tcgen05.st),
//              8 and 18 MMA issuers (alternate accumulator groups: a 64-column unit is only 384 tensor cycles, less than what
//              one issuer spends between units on barrier polls, elect and descriptors -- the tcgen05 queue holds 2-3 MMAs --
//              so a single issuer left the tensor pipe idle half of the time), 9-16 epilogue (32 columns x 32 rows each:
//              TMEM -> registers -> smem transpose -> coalesced bias / activation / residual / store), 17 TMA (weights).
//   LayerNorm  needs the handed-over statistics (ln_stats); without them the caller falls back to gemm_tc6.
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int RA_BM = 128, RA_BN = 64, RA_BK = 32, RA_NB = 8, RA_RAW = 3, RA_MAXKB = 6, RA_MAXS = 4;
constexpr int RA_BTILE = 2 * RA_BN * RA_BK * 4;     // 16 KB: hi half (8 KB) + lo half (8 KB) of a [64 n x 32 k] weight tile
constexpr int RA_RAWTILE = RA_BM * RA_BK * 4;       // 16 KB raw A k-block
constexpr int RA_PROD = 256, RA_MMA_WARP = 8, RA_EPI_WARP0 = 9, RA_EPI_WARPS = 8, RA_TMA_WARP = 17, RA_MMA2_WARP = 18;
constexpr int RA_BLOCK = (RA_MMA2_WARP + 1) * 32;   // 608
constexpr int RA_RAW_BAR = 5;
constexpr int RA_STAGE_FLOATS = 32 * 36;
constexpr int RA_DYN = RA_NB * RA_BTILE + RA_RAW * RA_RAWTILE + RA_EPI_WARPS * RA_STAGE_FLOATS * 4 + 1024;

struct RASmem {
  uint64_t full_b[RA_NB];        // weight tile landed (expect_tx 16 KB)
  uint64_t done_b[RA_NB];        // the MMAs that read the slot are complete (commit)
  uint64_t a_full[RA_MAXKB];     // A(kb) of the current row block is in TMEM (8 producer-warp arrivals)
  uint64_t a_free;               // every MMA of the row block is complete: A may be overwritten (one commit per issuer)
  uint64_t acc_full[RA_MAXS];    // accumulator stage holds a finished group (commit)
  uint64_t acc_empty[RA_MAXS];   // ... and the epilogue has read it (256 arrivals)
  uint32_t tmem_base;
  alignas(16) float gamma[128];
  alignas(16) float beta[128];
};

__device__ __forceinline__ void ra_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifdef NMRF_TRACE
__device__ long long* g_trace_ra = nullptr;
#endif
#define ra_trace(cond, idx) NMRF_TRACE_STAMP((cond) ? tp : nullptr, idx)

template <int ACT, bool LN>
__global__ void __launch_bounds__(RA_BLOCK, 1)
token_gemm_ra_kernel(const nmrf_gemm_args a, int n_rb) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ RASmem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  auto sB = [&](int i) { return base + i * RA_BTILE; };                       // hi half at +0, lo half at +8 KB
  auto sRaw = [&](int i) { return base + RA_NB * RA_BTILE + i * RA_RAWTILE; };
  float* stage_base = reinterpret_cast<float*>(base + RA_NB * RA_BTILE + RA_RAW * RA_RAWTILE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef NMRF_TRACE
  // cycle stamps of CTA 0 (tools/ra_trace.py): producer thread 0 [0,256), MMA issuers [256,768) / [768,1280), TMA lane [1280,1536),
  // epilogue warp 0 lane 0 [1536,2048)
  long long* const tp = blockIdx.x == 0 && lane == 0 ? g_trace_ra : nullptr;
#endif
  const int Ktot = a.Kx + a.Ke;
  const int nkb = (Ktot + RA_BK - 1) / RA_BK;               // <= RA_MAXKB (host check)
  const int nch = (a.N + RA_BN - 1) / RA_BN;                // 64-column chunks of the output
  const int G = nkb <= 3 ? nkb : (nkb == 4 ? 2 : 3);        // k-blocks per accumulator group
  const int ngrp = (nkb + G - 1) / G;
  int S = (512 - 64 * nkb) / 64;                            // accumulator stages
  if (S > RA_MAXS) S = RA_MAXS;
  const uint32_t acc_col0 = 64u * nkb;
  const int tstep = gridDim.x;
  // (A per-CTA rotation of the chunk order -- so that the CTAs of the grid do not stream the same 16 KB piece of the weight
  // image in lock step -- measured the same: 37.5 vs 35.7 us for the qkv projection.  NMRF_RA_ROT builds it.)
#ifdef NMRF_RA_ROT
  const int rot = (int)(blockIdx.x % (unsigned)nch);
#else
  const int rot = 0;
#endif
  auto chunk_of = [&](int c) { const int cc = c + rot; return cc >= nch ? cc - nch : cc; };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < RA_NB; ++i) { mbar_init(&sm.full_b[i], 1); mbar_init(&sm.done_b[i], 1); }
    for (int i = 0; i < RA_MAXKB; ++i) mbar_init(&sm.a_full[i], 8);
    mbar_init(&sm.a_free, 2);
    for (int i = 0; i < RA_MAXS; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], RA_EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (LN && tid < 128) { sm.gamma[tid] = a.ln_gamma[tid]; sm.beta[tid] = a.ln_beta[tid]; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp < 8) {
    // =============================================== producers ===============================================
    // warp w may touch TMEM lanes 32*(w%4)..+32 only: thread -> row 32*(w%4)+lane, k-columns 16*(w/4)..+16 of a k-block
    const int a_row = (warp & 3) * 32 + lane, a_c0 = (warp >> 2) * 4;
    const uint32_t a_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    // raw-A ring streaming across row blocks, two k-blocks ahead (as gemm_tc6.cu)
    const int f_c = tid & 7, f_r = tid >> 3;
    int f_t = blockIdx.x, f_kb = 0;
    const float* f_x[4];
    const float* f_e[4];
    uint32_t f_ok = 0;
    auto fetch_tile = [&]() {
      const int row0 = f_t * RA_BM;
      f_ok = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int grow = row0 + f_r + 32 * j;
        const bool ok = grow < a.rows;
        f_ok |= (ok ? 1u : 0u) << j;
        const int gr = ok ? grow : 0;
        f_x[j] = a.X + (size_t)gr * a.ldx + f_c * 4;
        f_e[j] = a.E ? a.E + (size_t)(gr / a.ediv) * a.lde + f_c * 4 - a.Kx : a.X;
      }
    };
    if (f_t < n_rb) fetch_tile();
    auto fetch_next = [&](uint32_t stage) {
      if (f_t < n_rb) {
        const uint32_t dst = smem_u32(sRaw(stage));
        const int k0 = f_kb * RA_BK;
        const bool in_x = k0 + f_c * 4 < a.Kx, in_k = k0 + f_c * 4 < Ktot;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = in_k && ((f_ok >> j) & 1u);
          const float* src = (in_x ? f_x[j] : f_e[j]) + k0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + swz(f_r + 32 * j, f_c)), "l"(ok ? src : a.X), "r"(ok ? 16 : 0));
        }
        if (++f_kb == nkb) { f_kb = 0; f_t += tstep; if (f_t < n_rb) fetch_tile(); }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch_next(0); fetch_next(1);
    uint32_t unit = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_rb; t += tstep, ++it) {
      const int row0 = t * RA_BM;
      const bool row_ok = row0 + a_row < a.rows;
      float mean = 0.f, rstd = 1.f;
      if (LN && row_ok) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(a.ln_stats) + row0 + a_row);
        mean = st.x; rstd = st.y;
      }
      for (int kb = 0; kb < nkb; ++kb, ++unit) {
        ra_trace(warp == 0 && unit < 64, unit * 4 + 0);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(RA_RAW_BAR), "r"(RA_PROD) : "memory");
        fetch_next((unit + 2) % RA_RAW);
        ra_trace(warp == 0 && unit < 64, unit * 4 + 1);
        const uint8_t* raw = sRaw(unit % RA_RAW);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = a_c0 + cc;
          const int kk = kb * RA_BK + c * 4;
          float4 v = *reinterpret_cast<const float4*>(raw + swz(a_row, c));
          if (LN && row_ok && kk < a.Kx) {
            const float4 g = *reinterpret_cast<const float4*>(sm.gamma + kk);
            const float4 b = *reinterpret_cast<const float4*>(sm.beta + kk);
            v.x = (v.x - mean) * rstd * g.x + b.x; v.y = (v.y - mean) * rstd * g.y + b.y;
            v.z = (v.z - mean) * rstd * g.z + b.z; v.w = (v.w - mean) * rstd * g.w + b.w;
          }
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float h = rna_tf32_fast(vv[j]);
            hi[cc * 4 + j] = __float_as_uint(h);
            lo[cc * 4 + j] = __float_as_uint(lo_tf32(vv[j], h));
          }
        }
        // A is single-buffered over row blocks: every MMA of the previous row block must be complete before its columns
        // are overwritten (the raw loads and the split above already ran under those MMAs)
        ra_trace(warp == 0 && unit < 64, unit * 4 + 2);
        if (kb == 0 && it > 0) mbar_wait_warp(&sm.a_free, (it - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem + a_lane + (uint32_t)(kb * 64 + a_c0 * 4);
        tmem_st16(ta, hi);
        tmem_st16(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) ra_arrive(&sm.a_full[kb]);
        ra_trace(warp == 0 && unit < 64, unit * 4 + 3);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == RA_MMA_WARP || warp == RA_MMA2_WARP) {
    // =============================================== MMA issuers ===============================================
    // both walk the same (chunk, k-block) sequence; issuer i issues the groups with grp % 2 == i.  tcgen05.commit tracks the
    // MMAs of the committing thread: done_b / acc_full are committed by the group's issuer, a_free by both.
    const uint32_t me = warp == RA_MMA_WARP ? 0u : 1u;
    uint32_t u = 0, grp = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_rb; t += tstep, ++it) {
      uint32_t seen = 0;      // k-blocks of A this issuer has already waited for in this row block
      for (int ci = 0; ci < nch; ++ci) {
        const int c = chunk_of(ci);
        const uint32_t idesc = make_idesc(min(RA_BN, a.N - c * RA_BN));
        for (int kb = 0; kb < nkb; ++kb, ++u) {
          const int slot = u % RA_NB;
          const int st = grp % S;
          const bool first = kb % G == 0, last = (kb % G == G - 1) || kb == nkb - 1;
          if ((grp & 1u) != me) { if (last) ++grp; continue; }
          ra_trace(u < 128, 256 + me * 512 + u * 4 + 0);
          if (first && grp >= (uint32_t)S) mbar_wait_warp(&sm.acc_empty[st], ((grp / S) - 1) & 1);   // the epilogue has read group - S
          // A(kb) of this row block is in TMEM.  Each issuer checks every k-block once per row block: with an odd number of
          // groups per chunk the k-blocks an issuer meets in chunk 0 are not the ones it meets later.
          if (!((seen >> kb) & 1u)) { mbar_wait_warp(&sm.a_full[kb], it & 1); seen |= 1u << kb; }
          ra_trace(u < 128, 256 + me * 512 + u * 4 + 1);
          mbar_wait_warp(&sm.full_b[slot], (u / RA_NB) & 1);                                        // the weight tile has landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          ra_trace(u < 128, 256 + me * 512 + u * 4 + 2);
          if (elect_one()) {
            const uint32_t bslot = smem_u32(sB(slot));
            const uint64_t dBh = make_desc(bslot), dBl = make_desc(bslot + RA_BTILE / 2);
            const uint32_t d = tmem + acc_col0 + (uint32_t)(st * RA_BN);
            const uint32_t tAh = tmem + (uint32_t)(kb * 64), tAl = tAh + 32;
            // small products first (fresh accumulator at the start of a group)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_tf32_ta(d, tAl + ks * 8, dBh + (uint64_t)(ks * 2), idesc, (ks > 0 || !first) ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_tf32_ta(d, tAh + ks * 8, dBl + (uint64_t)(ks * 2), idesc, 1u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_tf32_ta(d, tAh + ks * 8, dBh + (uint64_t)(ks * 2), idesc, 1u);
            umma_commit(&sm.done_b[slot]);
            if (last) umma_commit(&sm.acc_full[st]);
          }
          __syncwarp();
          ra_trace(u < 128, 256 + me * 512 + u * 4 + 3);
          if (last) ++grp;
        }
      }
      if (elect_one()) umma_commit(&sm.a_free);      // this issuer's MMAs of the row block (possibly none)
      __syncwarp();
    }
  } else if (warp == RA_TMA_WARP) {
    // =============================================== weight tiles (TMA) ===============================================
    if (elect_one()) {
      uint32_t u = 0;
      for (int t = blockIdx.x; t < n_rb; t += tstep) {
        for (int ci = 0; ci < nch; ++ci) {
          const int c = chunk_of(ci);
          for (int kb = 0; kb < nkb; ++kb, ++u) {
            const int slot = u % RA_NB;
            ra_trace(u < 128, 1280 + u * 2 + 0);
            if (u >= RA_NB) mbar_wait(&sm.done_b[slot], ((u - RA_NB) / RA_NB) & 1);
            ra_trace(u < 128, 1280 + u * 2 + 1);
            // tile image (n-chunk c / 2, k-block kb): 128 rows x 128 B; rows 64 (c & 1) .. + 63 are 8 contiguous KB of it
            const size_t toff = ((size_t)(c >> 1) * nkb + kb) * 4096 + (size_t)(c & 1) * 2048;
            const uint32_t bar = smem_u32(&sm.full_b[slot]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(RA_BTILE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sB(slot))), "l"(a.Wt_hi + toff), "r"(RA_BTILE / 2), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sB(slot)) + RA_BTILE / 2), "l"(a.Wt_lo + toff), "r"(RA_BTILE / 2), "r"(bar) : "memory");
          }
        }
      }
    }
  } else {
    // =============================================== epilogue ===============================================
    const int e = warp - RA_EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = e >> 2;                   // columns [32 half, 32 half + 32) of the 64-column chunk
    float* stage = stage_base + e * RA_STAGE_FLOATS;
    const int srow = lane >> 3, scol = (lane & 7) * 4;
    uint32_t grp = 0;
    for (int t = blockIdx.x; t < n_rb; t += tstep) {
      const int row0 = t * RA_BM;
      for (int ci = 0; ci < nch; ++ci) {
        const int c = chunk_of(ci);
        const int ncols = min(RA_BN, a.N - c * RA_BN);
        const bool mine = half * 32 < ncols;   // a chunk narrower than 33 columns has nothing for the second warp of a quarter
        float acc[32];
        for (int gi = 0; gi < ngrp; ++gi, ++grp) {
          const int st = grp % S;
          ra_trace(e == 0 && grp < 128, 1536 + grp * 4 + 0);
          mbar_wait_warp(&sm.acc_full[st], (grp / S) & 1, 32);
          ra_trace(e == 0 && grp < 128, 1536 + grp * 4 + 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (mine) {
#pragma unroll
            for (int h16 = 0; h16 < 2; ++h16) {
              float v[16];
              tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + acc_col0 + (uint32_t)(st * RA_BN + half * 32 + h16 * 16), v);
              if (gi == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[h16 * 16 + j] = v[j];
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[h16 * 16 + j] += v[j];
              }
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          ra_arrive(&sm.acc_empty[st]);
          ra_trace(e == 0 && grp < 128, 1536 + grp * 4 + 2);
        }
        if (mine) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stage + lane * 36 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
          __syncwarp();
          const int n = c * RA_BN + half * 32 + scol;
          if (n < a.N) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.bias) b = *reinterpret_cast<const float4*>(a.bias + n);
            float4 rr[8];                       // residual first (R may alias Y: read-before-write by the same thread)
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const int r = row0 + q * 32 + i8 * 4 + srow;
              rr[i8] = (a.R && r < a.rows) ? __ldcg(reinterpret_cast<const float4*>(a.R + (size_t)r * a.ldr + n))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const int lr = i8 * 4 + srow;
              const int r = row0 + q * 32 + lr;
              if (r < a.rows) {
                float4 o = *reinterpret_cast<const float4*>(stage + lr * 36 + scol);
                o.x = act_fast(o.x + b.x, ACT) + rr[i8].x; o.y = act_fast(o.y + b.y, ACT) + rr[i8].y;
                o.z = act_fast(o.z + b.z, ACT) + rr[i8].z; o.w = act_fast(o.w + b.w, ACT) + rr[i8].w;
                *reinterpret_cast<float4*>(a.Y + (size_t)r * a.ldy + n) = o;
              }
            }
          }
          __syncwarp();
        }
        ra_trace(e == 0 && grp - 1 < 128, 1536 + (grp - 1) * 4 + 3);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

template <int ACT, bool LN>
void launch_ra(const nmrf_gemm_args& a, int n_rb, int grid, cudaStream_t stream) {
  static PerDevice configured;
  ensure_dynamic_smem(token_gemm_ra_kernel<ACT, LN>, RA_DYN, configured);
  token_gemm_ra_kernel<ACT, LN><<<grid, RA_BLOCK, RA_DYN, stream>>>(a, n_rb);
}

}  // namespace

#ifdef NMRF_TRACE
int ra_set_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_trace_ra, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? NMRF_OK : NMRF_ERR_CUDA;
}
#endif

// K <= 192, tile images given, LayerNorm (if any) with handed-over statistics
bool token_gemm_ra_supported(const nmrf_gemm_args& a) {
  const int nkb = (a.Kx + a.Ke + RA_BK - 1) / RA_BK;
  return nkb >= 1 && nkb <= RA_MAXKB && a.Wt_hi && a.Wt_lo && (a.ln_gamma == nullptr || a.ln_stats != nullptr);
}

int token_gemm_ra(const nmrf_gemm_args& a, cudaStream_t stream) {
  const int num_sms = nmrf::num_sms();
  const int n_rb = (a.rows + RA_BM - 1) / RA_BM;
  const int grid = n_rb < num_sms ? n_rb : num_sms;
  const bool ln = a.ln_gamma != nullptr;
  switch (a.act * 2 + (ln ? 1 : 0)) {
    case 0: launch_ra<0, false>(a, n_rb, grid, stream); break;
    case 1: launch_ra<0, true>(a, n_rb, grid, stream); break;
    case 2: launch_ra<1, false>(a, n_rb, grid, stream); break;
    case 3: launch_ra<1, true>(a, n_rb, grid, stream); break;
    case 4: launch_ra<2, false>(a, n_rb, grid, stream); break;
    case 5: launch_ra<2, true>(a, n_rb, grid, stream); break;
    default: set_error("token_gemm: unknown activation %d", a.act); return NMRF_ERR_BAD_ARG;
  }
  count_launch();
  return check_launch("token_gemm_ra");
}

}  // namespace nmrf
