// Fused message-passing tail on tcgen05: one kernel per block for everything that is row-local after the attention
//     x1 = [att | x] . [Wproj | I]^T + b_proj          (NMP.py:358-359,570-571: proj + residual; the residual rides the
//                                                       tensor core as an identity block of the weight)
//     x  = x1 + fc2( GELU( fc1( LN2(x1) ) ) )           (NMP.py:362-363,572-573; timm Mlp)
// instead of three token GEMMs.  The ablation of the stand-alone GEMMs (tools/gemm_bench.py) showed ~14 us of fill/drain
// latency per launch and the [T,512] hidden activation (4x the token state, written by fc1 and re-read by fc2) as the
// dominant costs; here a CTA takes a 128-row tile through the whole chain and the only HBM/L2 traffic per tile is
// att + x in, x out and the weight stream.
//
//   tensor memory (512 columns):  [0,128) acc0: x1, later x1 + fc2(...)   [128,256) LN2(x1) hi   [256,384) LN2(x1) lo
//                                 [384,512) phase 1: two A buffers (hi|lo of a 32-wide k-block); phase 3: two fc1 accumulators
//                                 of 64 hidden columns each
//   shared memory:                weight ring 3 x 32 KB (every "unit" of the weight stream is a 16 KB hi image + a 16 KB lo
//                                 image and 768 tensor cycles), GELU'd hidden chunk hi/lo [128 x 64] (64 KB, SS-mode A operand
//                                 of fc2; doubles as the store staging of the final epilogue), raw-A ring 3 x 16 KB
//   unit stream per tile:         P1(0..n1-1)   F1(0) F1(1) F2(0) F1(2) F2(1) ... F1(7) F2(6) F2(7), two units each:
//                                 P1(j): k-block j of [Wproj | I] (128 x 32);  F1(c): fc1 rows 64c..64c+63 (64 x 128, two
//                                 k-block pairs);  F2(c): fc2 columns 64c..64c+63 (128 x 64, two k-blocks).
//                                 nmrf_b200/hotpath.py packs the stream in exactly this order (pack_mlp_stream).
//   warps                         0-7 producers (phase-1 A operand: raw ring -> hi/lo split -> tcgen05.st, as gemm_tc6),
//                                 8 MMA issuer, 9-16 LN / GELU / store warps (thread = row = TMEM lane), 17 TMA.
// Arithmetic: 3xTF32 with RN hi / exact lo split as in gemm_tc6; LayerNorm two-pass in fp32 from the fp32
// accumulator; GELU as gemm_tc6.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int M_BM = 128, M_BK = 32, M_NB = 3, M_RAW = 3;
constexpr int M_TILE = M_BM * M_BK * 4;            // 16 KB image
constexpr int M_UNIT = 2 * M_TILE;                 // 32 KB: hi image + lo image
constexpr int M_HID = 512, M_CH = 64, M_NCH = M_HID / M_CH;      // fc1 width, hidden chunk, chunks
constexpr int M_PROD = 256, M_MMA_WARP = 8, M_EPI_WARP0 = 9, M_EPI_WARPS = 8, M_TMA_WARP = 17;
constexpr int M_BLOCK = (M_TMA_WARP + 1) * 32;     // 576
constexpr int M_HANDOFF = M_PROD + 32;
constexpr int M_RAW_BAR = 5;
constexpr int M_COL_ALN_HI = 128, M_COL_ALN_LO = 256, M_COL_X = 384;      // TMEM columns
constexpr int M_OFF_H = M_NB * M_UNIT;             // 96 KB
constexpr int M_OFF_RAW = M_OFF_H + 4 * M_TILE;    // + 64 KB
constexpr int M_DYN = M_OFF_RAW + M_RAW * M_TILE + 1024;

struct MSmem {
  uint64_t done[M_NB];        // MMAs of the unit that used weight slot s are complete
  uint64_t full_b[M_NB];      // weight unit landed (expect_tx 32 KB)
  uint64_t p1_full;           // acc0 = x1 - b_proj is complete (commit)
  uint64_t aln_full;          // LN2(x1) hi/lo are in TMEM (8 warp arrivals)
  uint64_t acc1_full[2];      // fc1 chunk accumulator complete (commit)
  uint64_t acc1_empty[2];     // ... drained (8 warp arrivals)
  uint64_t h_full;            // hidden chunk hi/lo in shared memory (8 warp arrivals)
  uint64_t h_free;            // fc2 MMAs of the chunk complete (commit)
  uint64_t acc0_final;        // all MMAs of the tile complete (commit)
  uint64_t acc0_empty;        // final epilogue has read acc0 (8 warp arrivals)
  uint32_t tmem_base;
  alignas(16) float gamma[128];
  alignas(16) float beta[128];
  alignas(16) float bmid[128];
  alignas(16) float bout[128];
  alignas(16) float b1[M_HID];
};

__device__ __forceinline__ void mbar_arrive_m(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// optional cycle trace of CTA 0 (debug tooling, nmrf_debug_set_trace): MMA warp: unit g -> [g*4 + {0 before weight wait, 1 after,
// 2 after issue}], 512 units max; LN/GELU warp 9: 2048 + chunk*8 + {0 before acc1_full, 1 after, 2 after GELU, 3 after h_free,
// 4 after stores}; 3968 + tile*8 + {0 p1_full seen, 1 LN done, 2 acc0_final seen, 3 stored}
__device__ long long* g_trace_m = nullptr;
__device__ __forceinline__ void mtrace(long long* tp, int idx) {
  if (tp && idx < 4096) tp[idx] = clock64();
}
__device__ __forceinline__ uint32_t idesc_n(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, float* v) { tmem_ld32(taddr, v); }

// One hidden chunk on one of the 16 GELU workers (all producer and LN/store warps: 4 per TMEM lane quarter, worker j takes
// accumulator columns 16 j .. 16 j + 15 of its row): h = GELU(fc1 + b1) -> hi / lo -> the swizzled A-operand tiles of fc2.
// The fc1 accumulator is released right after the TMEM load; the hidden buffer is written once the fc2 MMAs of the previous
// chunk have read it.
__device__ __forceinline__ void gelu_worker(MSmem& sm, uint32_t tmem_lane, uint8_t* sH, int row, int j, uint32_t gc, const float* b1,
                                            int lane, long long* tp) {
  const int b = gc & 1;
  mtrace(tp, 2048 + gc * 8 + 0);
  mbar_wait_warp(&sm.acc1_full[b], (gc >> 1) & 1);
  mtrace(tp, 2048 + gc * 8 + 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float v[16];
  tmem_ld16(tmem_lane + (uint32_t)(M_COL_X + b * 64 + j * 16), v);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive_m(&sm.acc1_empty[b]);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i] + b1[i]);
  mtrace(tp, 2048 + gc * 8 + 2);
  if (gc >= 1) mbar_wait_warp(&sm.h_free, (gc - 1) & 1);
  mtrace(tp, 2048 + gc * 8 + 3);
  uint8_t* hi_t = sH + (j >> 1) * M_TILE;
  uint8_t* lo_t = hi_t + 2 * M_TILE;
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    float4 h, l;
    h.x = rna_tf32_fast(v[c4 * 4]); h.y = rna_tf32_fast(v[c4 * 4 + 1]); h.z = rna_tf32_fast(v[c4 * 4 + 2]); h.w = rna_tf32_fast(v[c4 * 4 + 3]);
    l.x = v[c4 * 4] - h.x; l.y = v[c4 * 4 + 1] - h.y; l.z = v[c4 * 4 + 2] - h.z; l.w = v[c4 * 4 + 3] - h.w;
    const uint32_t so = swz(row, (j & 1) * 4 + c4);
    *reinterpret_cast<float4*>(hi_t + so) = h;
    *reinterpret_cast<float4*>(lo_t + so) = l;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive_m(&sm.h_full);
  mtrace(tp, 2048 + gc * 8 + 4);
}

__global__ void __launch_bounds__(M_BLOCK, 1)
mlp_chain_kernel(const nmrf_mlp_args a, int ntiles, int dbg) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ MSmem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  auto sW = [&](int slot) { return base + slot * M_UNIT; };            // hi image at +0, lo image at +16 KB
  uint8_t* sH = base + M_OFF_H;                                        // hi q0, hi q1, lo q0, lo q1 (16 KB each)
  auto sRaw = [&](int i) { return base + M_OFF_RAW + i * M_TILE; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* const tp = (blockIdx.x == 0 && (tid == M_MMA_WARP * 32 || tid == M_EPI_WARP0 * 32)) ? g_trace_m : nullptr;
  const int Ktot = a.Kx + a.Ke;
  const int n1 = Ktot / M_BK;                      // phase-1 units per tile
  const int upt = n1 + 4 * M_NCH;                  // units per tile
  const int tstep = gridDim.x;
  // Every CTA walks the same weight stream; in lockstep all 148 of them would pull the same 32 KB out of the same few L2
  // slices at the same time (measured: 7-8k cycles per bulk copy).  Each CTA therefore starts at its own rotation of the
  // k-block order (phase 1) and of the hidden-chunk order (phase 3); both are sums, so only the fp32 summation order moves.
  const int rot = blockIdx.x & 7;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < M_NB; ++i) { mbar_init(&sm.done[i], 1); mbar_init(&sm.full_b[i], 1); }
    mbar_init(&sm.p1_full, 1); mbar_init(&sm.aln_full, M_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.acc1_full[i], 1); mbar_init(&sm.acc1_empty[i], M_EPI_WARPS + 8); }
    mbar_init(&sm.h_full, M_EPI_WARPS + 8); mbar_init(&sm.h_free, 1);
    mbar_init(&sm.acc0_final, 1); mbar_init(&sm.acc0_empty, M_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 128) { sm.gamma[tid] = a.ln_gamma[tid]; sm.beta[tid] = a.ln_beta[tid]; sm.bmid[tid] = a.bias_mid[tid]; sm.bout[tid] = a.bias_out[tid]; }
  if (tid < M_HID) sm.b1[tid] = a.b1[tid];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp < 8) {
    // =============================================== producers (phase 1) ===============================================
    const int a_row = (warp & 3) * 32 + lane, a_c0 = (warp >> 2) * 4;
    const uint32_t a_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    const int f_c = tid & 7, f_r = tid >> 3;
    int f_t = blockIdx.x, f_kb = 0;
    const float* f_x[4];
    const float* f_e[4];
    uint32_t f_ok = 0;
    auto fetch_tile = [&]() {
      const int row0 = f_t * M_BM;
      f_ok = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int grow = row0 + f_r + 32 * j;
        const bool ok = grow < a.rows;
        f_ok |= (ok ? 1u : 0u) << j;
        const int gr = ok ? grow : 0;
        f_x[j] = a.X + (size_t)gr * a.ldx + f_c * 4;
        f_e[j] = a.E ? a.E + (size_t)gr * a.lde + f_c * 4 - a.Kx : a.X;
      }
    };
    if (f_t < ntiles) fetch_tile();
    auto fetch_next = [&](uint32_t stage) {
      if (f_t < ntiles) {
        const uint32_t dst = smem_u32(sRaw(stage));
        const int k0 = ((f_kb + rot) % n1) * M_BK;
        const bool in_x = k0 + f_c * 4 < a.Kx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = (f_ok >> j) & 1u;
          const float* src = (in_x ? f_x[j] : f_e[j]) + k0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + swz(f_r + 32 * j, f_c)), "l"(ok ? src : a.X), "r"(ok ? 16 : 0));
        }
        if (++f_kb == n1) { f_kb = 0; f_t += tstep; if (f_t < ntiles) fetch_tile(); }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch_next(0); fetch_next(1);
    uint32_t pu = 0;          // phase-1 units produced so far by this CTA: ring stage pu % 3, A buffer pu & 1, hand-off barrier 1 + pu % 3
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
      for (int kb = 0; kb < n1; ++kb, ++pu) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(M_RAW_BAR), "r"(M_PROD) : "memory");
        // the stage consumed one unit ago is free again (every producer is past its reads): re-arm it two k-blocks ahead
        fetch_next((pu + 2) % M_RAW);
        const uint8_t* raw = sRaw(pu % M_RAW);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 v = *reinterpret_cast<const float4*>(raw + swz(a_row, a_c0 + cc));
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float h = rna_tf32_fast(vv[j]);
            hi[cc * 4 + j] = __float_as_uint(h);
            lo[cc * 4 + j] = __float_as_uint(vv[j] - h);
          }
        }
        // A buffer pu & 1 (TMEM columns of the fc1 accumulators): free once the MMAs of the phase-1 unit two back are
        // complete; for the first two units of a tile, once ALL MMAs of the previous tile are (acc0_final)
        if (kb >= 2) {
          const uint32_t g = (uint32_t)it * upt + kb - 2;
          mbar_wait_warp(&sm.done[g % M_NB], (g / M_NB) & 1);
        } else if (it > 0) {
          mbar_wait_warp(&sm.acc0_final, (it - 1) & 1, (dbg & 2) ? 512 : 0);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem + a_lane + (uint32_t)(M_COL_X + (pu & 1) * 64 + a_c0 * 4);
        tmem_st16(ta, hi);
        tmem_st16(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + pu % 3), "r"(M_HANDOFF) : "memory");
      }
      // phase 3: the producers are GELU workers 0 and 1 of their lane quarter
      for (int c = 0; c < M_NCH; ++c)
        gelu_worker(sm, tmem + a_lane, sH, a_row, warp >> 2, (uint32_t)(it * M_NCH + c),
                    sm.b1 + ((c + rot) & 7) * M_CH + (warp >> 2) * 16, lane, nullptr);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == M_MMA_WARP) {
    // =============================================== MMA issuer ===============================================
    const uint32_t idesc128 = idesc_n(128), idesc64 = idesc_n(64);
    const uint32_t acc0 = tmem;
    const uint64_t dHh0 = make_desc(smem_u32(sH)), dHl0 = make_desc(smem_u32(sH + 2 * M_TILE));
    uint32_t g = 0;           // global unit counter (weight ring)
    uint32_t pu = 0;          // phase-1 unit counter (A buffers / hand-off barriers)
    uint32_t gc = 0;          // global hidden-chunk counter
    auto wait_all = [&](uint64_t* bar, uint32_t parity) {
      if (dbg & 4096) { const uint32_t addr = smem_u32(bar); while (!mbar_try(addr, parity)) {} } else mbar_wait_warp(bar, parity);
    };
    auto wait_b = [&](uint32_t unit) { wait_all(&sm.full_b[unit % M_NB], (unit / M_NB) & 1); };
    // one F1 unit: 8 k-steps of the k-block pair p against 64 fc1 rows; A = LN2(x1) from TMEM
    auto issue_f1 = [&](uint32_t c_local, uint32_t cg, int p) {
      mtrace(tp, g * 4 + 0);
      if (dbg & 512) {          // fine-grained stamps (debug): 1024 + g*8 + {0 start, 1 after try_wait, 2 after syncwarp, 3 in elected block}
        mtrace(tp, 1024 + g * 8 + 0);
        if (lane == 0) { mbar_wait(&sm.full_b[g % M_NB], (g / M_NB) & 1); mtrace(tp, 1024 + g * 8 + 1); }
        __syncwarp();
        mtrace(tp, 1024 + g * 8 + 2);
      } else
      wait_b(g);
      mtrace(tp, g * 4 + 1);
      if (elect_one()) {
        if (dbg & 512) mtrace(tp, 1024 + g * 8 + 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg & 512) mtrace(tp, 1024 + g * 8 + 4);
        const uint32_t bslot = smem_u32(sW(g % M_NB));
        const uint32_t d = tmem + (uint32_t)(M_COL_X + (cg & 1) * 64);
#pragma unroll
        for (int kk = 0; kk < ((dbg & 8) ? 0 : 8); ++kk) {
          const uint64_t dBh = make_desc(bslot + (kk >> 2) * (M_TILE / 2)) + (uint64_t)((kk & 3) * 2);
          const uint64_t dBl = make_desc(bslot + M_TILE + (kk >> 2) * (M_TILE / 2)) + (uint64_t)((kk & 3) * 2);
          const uint32_t kcol = (uint32_t)(p * 64 + kk * 8);
          umma_tf32_ta(d, tmem + M_COL_ALN_LO + kcol, dBh, idesc64, (p > 0 || kk > 0) ? 1u : 0u);
          umma_tf32_ta(d, tmem + M_COL_ALN_HI + kcol, dBl, idesc64, 1u);
          umma_tf32_ta(d, tmem + M_COL_ALN_HI + kcol, dBh, idesc64, 1u);
        }
        if (dbg & 512) mtrace(tp, 1024 + g * 8 + 5);
        umma_commit(&sm.done[g % M_NB]);
        if (p == 1) umma_commit(&sm.acc1_full[cg & 1]);
        if (dbg & 512) mtrace(tp, 1024 + g * 8 + 6);
      }
      __syncwarp();
      mtrace(tp, g * 4 + 2);
      ++g;
      (void)c_local;
    };
    // one F2 unit: 4 k-steps of hidden k-block q against the 128 fc2 rows; A = GELU'd hidden chunk from shared memory
    auto issue_f2 = [&](int q, bool last_of_chunk, bool last_of_tile) {
      mtrace(tp, g * 4 + 0);
      wait_b(g);
      mtrace(tp, g * 4 + 1);
      if (elect_one()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bslot = smem_u32(sW(g % M_NB));
        const uint64_t dBh = make_desc(bslot), dBl = make_desc(bslot + M_TILE);
        const uint64_t dAh = dHh0 + (uint64_t)(q * (M_TILE >> 4)), dAl = dHl0 + (uint64_t)(q * (M_TILE >> 4));
#pragma unroll
        for (int ks = 0; ks < ((dbg & 4) ? 0 : 4); ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);
          umma_tf32(acc0, dAl + adv, dBh + adv, idesc128, 1u);
          umma_tf32(acc0, dAh + adv, dBl + adv, idesc128, 1u);
          umma_tf32(acc0, dAh + adv, dBh + adv, idesc128, 1u);
        }
        umma_commit(&sm.done[g % M_NB]);
        if (last_of_chunk) umma_commit(&sm.h_free);
        if (last_of_tile) umma_commit(&sm.acc0_final);
      }
      __syncwarp();
      mtrace(tp, g * 4 + 2);
      ++g;
    };
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
      if (it > 0) wait_all(&sm.acc0_empty, (it - 1) & 1);          // previous tile's x has been read out of acc0
      // ---- phase 1: acc0 = [att | x] . [Wproj | I]^T
      for (int kb = 0; kb < n1; ++kb, ++pu, ++g) {
        mtrace(tp, g * 4 + 0);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + pu % 3), "r"(M_HANDOFF) : "memory");
        mtrace(tp, g * 4 + 3);
        wait_b(g);
        mtrace(tp, g * 4 + 1);
        if (elect_one()) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t bslot = smem_u32(sW(g % M_NB));
          const uint64_t dBh = make_desc(bslot), dBl = make_desc(bslot + M_TILE);
          const uint32_t tAh = tmem + (uint32_t)(M_COL_X + (pu & 1) * 64), tAl = tAh + 32;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);
            umma_tf32_ta(acc0, tAl + ks * 8, dBh + adv, idesc128, (kb > 0 || ks > 0) ? 1u : 0u);
            umma_tf32_ta(acc0, tAh + ks * 8, dBl + adv, idesc128, 1u);
            umma_tf32_ta(acc0, tAh + ks * 8, dBh + adv, idesc128, 1u);
          }
          umma_commit(&sm.done[g % M_NB]);
          if (kb == n1 - 1) umma_commit(&sm.p1_full);
        }
        __syncwarp();
        mtrace(tp, g * 4 + 2);
      }
      // ---- phase 3: F1(0) F1(1) F2(0) F1(2) F2(1) ... F1(7) F2(6) F2(7)
      wait_all(&sm.aln_full, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = 0; c <= M_NCH; ++c) {
        if (c < M_NCH) {
          const uint32_t cg = gc + c;                                    // use index of accumulator cg & 1 is cg >> 1
          if (cg >= 2) {
            wait_all(&sm.acc1_empty[cg & 1], ((cg >> 1) - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          issue_f1(c, cg, 0);
          issue_f1(c, cg, 1);
        }
        if (c >= 1) {
          const uint32_t cg = gc + c - 1;
          wait_all(&sm.h_full, cg & 1);
          issue_f2(0, false, false);
          issue_f2(1, true, c == M_NCH);
        }
      }
      gc += M_NCH;
    }
  } else if (warp == M_TMA_WARP) {
    // =============================================== weight stream (TMA) ===============================================
    if (elect_one()) {
      uint32_t g = 0;
      for (int t = blockIdx.x; t < ntiles; t += tstep) {
        for (int u = 0; u < upt; ++u, ++g) {
          const int slot = g % M_NB;
          if (g >= M_NB) mbar_wait(&sm.done[slot], ((g - M_NB) / M_NB) & 1);
          const uint32_t bar = smem_u32(&sm.full_b[slot]);
          if ((dbg & 256) && g >= M_NB) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); continue; }
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(M_UNIT) : "memory");
          // source unit: stream layout is P1(0..n1-1), F1(c,p) at n1 + 2c + p, F2(c,q) at n1 + 16 + 2c + q; issue order of
          // phase 3 in blocks of two units: F1(0) F1(1) F2(0) F1(2) F2(1) ... F1(7) F2(6) F2(7), chunk = (c + rot) % 8
          int su;
          if (u < n1) {
            su = (u + rot) % n1;
          } else {
            const int s3 = u - n1, blk = s3 >> 1, pq = s3 & 1;
            const bool is_f2 = blk >= 2 && (blk == 15 || (blk & 1) == 0);
            const int cs = is_f2 ? (blk == 15 ? 7 : blk / 2 - 1) : (blk == 0 ? 0 : (blk + 1) / 2);
            su = n1 + (is_f2 ? 16 : 0) + ((cs + rot) & 7) * 2 + pq;
          }
          // several smaller bulk copies per unit: one copy keeps only a few L2 requests in flight
          const int nsplit = (dbg & 32) ? 1 : (dbg & 64) ? 2 : (dbg & 128) ? 16 : 8;
          const uint32_t piece = M_UNIT / nsplit;
          for (int i = 0; i < nsplit; ++i)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sW(slot)) + i * piece), "l"(a.Wstream + (size_t)su * (M_UNIT / 4) + (size_t)i * (piece / 4)), "r"(piece), "r"(bar) : "memory");
        }
      }
    }
  } else {
    // =============================================== LN / GELU / store warps ===============================================
    const int e = warp - M_EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = e >> 2;                   // two warps per quarter: columns [64 half, 64 half + 64) / hidden k-block `half`
    const int row = q * 32 + lane;             // row of the tile owned by this thread
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    // this warp's 4 KB of the hidden buffer: rows q*32..+32 of hi image `half`; also its store staging
    uint8_t* my_h_hi = sH + half * M_TILE;
    uint8_t* my_h_lo = sH + (2 + half) * M_TILE;
    float* stage = reinterpret_cast<float*>(my_h_hi + q * 32 * 128);
    const int srow = lane >> 3, sc8 = lane & 7;
    uint32_t gc = 0;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
      const int row0 = t * M_BM;
      // ---- LN2 of x1 = acc0 + b_proj, thread = row, two-pass (like torch); this warp normalises columns 64 half..+64 ----
      mbar_wait_warp(&sm.p1_full, it & 1);
      mtrace(tp, 3968 + it * 8 + 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[32];
      float s = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        tmem_ld32(tmem + t_lane + (uint32_t)(ch * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) s += v[j] + sm.bmid[ch * 32 + j];
      }
      const float mean = s * (1.f / 128.f);
      float qq = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        tmem_ld32(tmem + t_lane + (uint32_t)(ch * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = (v[j] + sm.bmid[ch * 32 + j]) - mean; qq = fmaf(d, d, qq); }
      }
      const float rstd = 1.f / sqrtf(qq * (1.f / 128.f) + 1e-5f);
#pragma unroll 1
      for (int ch = 2 * half; ch < 2 * half + 2; ++ch) {
        tmem_ld32(tmem + t_lane + (uint32_t)(ch * 32), v);
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int k = ch * 32 + j;
          const float y = ((v[j] + sm.bmid[k]) - mean) * rstd * sm.gamma[k] + sm.beta[k];
          const float h = rna_tf32_fast(y);
          hi[j] = __float_as_uint(h);
          lo[j] = __float_as_uint(y - h);
        }
        tmem_st16(tmem + t_lane + (uint32_t)(M_COL_ALN_HI + ch * 32), hi);
        tmem_st16(tmem + t_lane + (uint32_t)(M_COL_ALN_HI + ch * 32 + 16), hi + 16);
        tmem_st16(tmem + t_lane + (uint32_t)(M_COL_ALN_LO + ch * 32), lo);
        tmem_st16(tmem + t_lane + (uint32_t)(M_COL_ALN_LO + ch * 32 + 16), lo + 16);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_m(&sm.aln_full);
      mtrace(tp, 3968 + it * 8 + 1);
      // ---- hidden chunks: GELU(fc1 + b1) -> hi/lo -> shared memory (A operand of fc2); workers 2 and 3 of the quarter ----
      for (int c = 0; c < M_NCH; ++c, ++gc)
        gelu_worker(sm, tmem + t_lane, sH, row, 2 + half, gc, sm.b1 + ((c + rot) & 7) * M_CH + (2 + half) * 16, lane, tp);
      // ---- final: x = acc0 + (b_proj + b_fc2), columns 64 half..+64 of this warp's 32 rows, coalesced through the staging ----
      mbar_wait_warp(&sm.acc0_final, it & 1);
      mtrace(tp, 3968 + it * 8 + 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float w[32];
      tmem_ld32(tmem + t_lane + (uint32_t)(half * 64), v);
      tmem_ld32(tmem + t_lane + (uint32_t)(half * 64 + 32), w);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_m(&sm.acc0_empty);
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const float* src = ch == 0 ? v : w;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8)
          *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(stage) + swz(lane, c8)) =
              make_float4(src[c8 * 4], src[c8 * 4 + 1], src[c8 * 4 + 2], src[c8 * 4 + 3]);
        __syncwarp();
        const int n = half * 64 + ch * 32 + sc8 * 4;
        const float4 bo = *reinterpret_cast<const float4*>(sm.bout + n);
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const int lr = i8 * 4 + srow;
          const int r = row0 + q * 32 + lr;
          if (r < a.rows) {
            float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<uint8_t*>(stage) + swz(lr, sc8));
            o.x += bo.x; o.y += bo.y; o.z += bo.z; o.w += bo.w;
            *reinterpret_cast<float4*>(a.Y + (size_t)r * a.ldy + n) = o;
          }
        }
        __syncwarp();
      }
      mtrace(tp, 3968 + it * 8 + 3);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

}  // namespace

int mlp_set_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_trace_m, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? NMRF_OK : NMRF_ERR_CUDA;
}

int mlp_chain(const nmrf_mlp_args& a, cudaStream_t stream) {
  static int num_sms = 0;
  static bool configured = false;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!configured) {
    cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, M_DYN);
    configured = true;
  }
  const int ntiles = (a.rows + M_BM - 1) / M_BM;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  static const int dbg = [] { const char* e = getenv("NMRF_B200_DBG"); return e ? atoi(e) : 0; }();
  mlp_chain_kernel<<<grid, M_BLOCK, M_DYN, stream>>>(a, ntiles, dbg);
  count_launch();
  return check_launch("mlp_chain");
}

}  // namespace nmrf
