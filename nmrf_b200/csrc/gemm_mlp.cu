// Fused message-passing tail on tcgen05: one kernel per block for everything that is row-local after the attention
//     x1 = x + (att . Wproj^T + b_proj)                (NMP.py:358-359,570-571: proj + residual)
//     x  = x1 + (fc2( GELU( fc1( LN2(x1) ) ) ) + b_fc2) (NMP.py:362-363,572-573; timm Mlp)
// The residual stream lives in fp32 REGISTERS of the worker threads (thread = row, 32 columns each), never in a tensor-memory
// accumulator: the tensor core's accumulator update rounds toward zero (measured, tools/precision_probe.py: a relative bias of
// -1.6e-8 per MMA accumulated into the same columns, i.e. -7.7e-7 for K = 128 and -3e-6 for K = 512, against +-1e-10 for fp32
// FMAs), and every such truncation is relative to the magnitude IN the accumulator.  With x preloaded there (round 1) the 240
// MMAs of a tile shaved 7e-6 |x| off the residual stream per block -- 35 times the error of the fp32 reference arithmetic and
// the main source of flipped argmax / median decisions downstream.  Now the accumulators only ever hold the (small) updates,
// fc2's accumulator is drained into the registers every four hidden chunks, and inside a fresh accumulator the small
// lo.hi / hi.lo products are issued before the hi.hi ones; all sums that involve x are fp32 round-to-nearest adds.
// instead of three token GEMMs.  The ablation of the stand-alone GEMMs (tools/gemm_bench.py) showed ~14 us of fill/drain
// latency per launch and the [T,512] hidden activation (4x the token state, written by fc1 and re-read by fc2) as the
// dominant costs; here a CTA takes a 128-row tile through the whole chain and the only HBM/L2 traffic per tile is
// att + x in, x out and the weight stream.
//
//   tensor memory (512 columns):  [0,128) acc0: x1, later x1 + fc2(...)   [128,256) LN2(x1) hi   [256,384) LN2(x1) lo
//                                 [384,512) phase 1: two A buffers (hi|lo of a 32-wide k-block); phase 3: FOUR fc1 accumulators
//                                 of 32 hidden columns each
//   shared memory:                weight ring 3 x 32 KB (every "unit" of the weight stream is a 16 KB hi image + a 16 KB lo
//                                 image and 768 tensor cycles), TWO GELU'd hidden chunks hi/lo [128 x 32] (2 x 32 KB, SS-mode A
//                                 operand of fc2; also the store staging of the final epilogue), raw-A ring 3 x 16 KB
//   hidden chunks of 32:          fc1 runs three chunks ahead of fc2 (four accumulators), the hidden tile is double buffered, so
//                                 the chain  fc1(c) -> GELU(c) -> fc2(c)  of one chunk overlaps the tensor work of its neighbours
//                                 (with 64-wide chunks and one hidden buffer GELU + stores + fc2 were serialised: 5.4k cycles
//                                 per 3.1k cycles of tensor work, tools/mlp_trace.py)
//   unit stream per tile:         P1(0..n1-1)   F1(0) F1(1) F1(2)   [F1(c+3) F2(c)] c = 0..12   F2(13) F2(14) F2(15)
//                                 P1(j): k-block j of [Wproj | I] (128 x 32);  F1(c): fc1 rows 32c..32c+31 (32 x 128, four
//                                 k-block sub-images);  F2(c): fc2 columns 32c..32c+31 (128 x 32).
//                                 nmrf_b200/ops.py: pack_mlp_stream lays the stream out as P1 | F1(0..15) | F2(0..15).
//   warps                         0-7 producers (phase-1 A operand: raw ring -> hi/lo split -> tcgen05.st, as gemm_tc6) and GELU
//                                 workers, 9-16 LN / GELU / store warps (thread = row = TMEM lane), 17 TMA, and TWO MMA issuers:
//                                 warp 8 issues phase 1 and the fc2 units, warp 18 the fc1 units.  The tcgen05 issue queue holds
//                                 only 2-3 MMAs (tools/probes/seq_probe.cu: 12 MMAs block the issuing thread for 600 of their 768
//                                 cycles), so whatever one issuer spends between units (two barrier polls, elect, descriptors:
//                                 ~500 cycles) is tensor idle time unless the other issuer's unit is executing meanwhile.
// Arithmetic: 3xTF32 with RN hi / RN lo split as in gemm_tc6; LayerNorm two-pass in fp32 on x1; GELU as gemm_tc6.
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int M_BM = 128, M_BK = 32, M_NB = 3, M_RAW = 3;
constexpr int M_TILE = M_BM * M_BK * 4;            // 16 KB image
constexpr int M_UNIT = 2 * M_TILE;                 // 32 KB: hi image + lo image
constexpr int M_HID = 512, M_CH = 32, M_NCH = M_HID / M_CH;      // fc1 width, hidden chunk, chunks (16)
constexpr int M_AHEAD = 3;                         // fc1 runs this many chunks ahead of fc2
constexpr int M_PROD = 256, M_MMA_WARP = 8, M_EPI_WARP0 = 9, M_EPI_WARPS = 8, M_TMA_WARP = 17, M_MMA2_WARP = 18;
constexpr int M_WORKERS = M_EPI_WARPS + 8;         // GELU workers (warps)
constexpr int M_BLOCK = (M_MMA2_WARP + 1) * 32;    // 608
constexpr int M_HANDOFF = M_PROD + 32;
constexpr int M_RAW_BAR = 5;
constexpr int M_COL_ALN_HI = 128, M_COL_ALN_LO = 256, M_COL_X = 384;      // TMEM columns
constexpr int M_ABUF = 4;                          // phase-1 A buffers (hi|lo of a k-block, 64 columns): [384,512) and [128,256)
__device__ __forceinline__ uint32_t abuf_col(uint32_t b) { return b < 2 ? M_COL_X + 64 * b : M_COL_ALN_HI + 64 * (b - 2); }
constexpr int M_OFF_H = M_NB * M_UNIT;             // 96 KB
constexpr int M_OFF_RAW = M_OFF_H + 4 * M_TILE;    // + 64 KB: hidden buffer b = [hi 16 KB | lo 16 KB] at M_OFF_H + b * 32 KB
constexpr int M_DYN = M_OFF_RAW + M_RAW * M_TILE + 1024;

struct MSmem {
  uint64_t done[M_NB];        // MMAs of the unit that used weight slot s are complete
  uint64_t full_b[M_NB];      // weight unit landed (expect_tx 32 KB)
  uint64_t p1_full;           // acc0 = x1 - b_proj is complete (commit)
  uint64_t aln_full;          // LN2(x1) hi/lo are in TMEM (16 warp arrivals)
  uint64_t acc1_full[4];      // fc1 chunk accumulator complete (commit)
  uint64_t acc1_empty[4];     // ... drained (16 warp arrivals)
  uint64_t h_full[2];         // hidden chunk hi/lo in shared memory (16 warp arrivals)
  uint64_t h_free[2];         // fc2 MMAs of the chunk complete (commit)
  uint64_t acc0_final;        // all MMAs of the tile complete (commit)
  uint64_t acc0_empty;        // final epilogue has read acc0 (16 warp arrivals)
  uint64_t part_full;         // fc2 units 4d .. 4d+3 complete: acc0 holds their partial sum (commit; d = 0, 1, 2)
  uint64_t part_drained;      // ... and every worker has added it to its registers (16 warp arrivals): acc0 may be overwritten
  uint32_t tmem_base;
  alignas(16) float gamma[128];
  alignas(16) float beta[128];
  alignas(16) float bmid[128];
  alignas(16) float bout[128];      // fc2 bias
  alignas(16) float b1[M_HID];
  float red[2][4][M_BM];      // LayerNorm partial sums: [pass][worker of the quarter][row]
};

__device__ __forceinline__ void mbar_arrive_m(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// optional cycle trace of CTA 0 (debug tooling, nmrf_debug_set_trace; tools/mlp_trace.py): MMA warp: unit g -> g*4 + {0 before the
// weight wait, 1 after, 2 after issue, 3 after the hand-off barrier (phase 1)}; GELU warp 9: 2048 + chunk*8 + {0 before acc1_full,
// 1 after, 2 after GELU, 3 after h_free, 4 after stores}; 3968 + tile*8 + {0 p1_full seen, 1 LN done, 2 acc0_final seen, 3 stored}
#ifdef NMRF_TRACE
__device__ long long* g_trace_m = nullptr;
#endif
#define mtrace(tp, idx) NMRF_TRACE_STAMP(tp, idx)
__device__ __forceinline__ uint32_t idesc_n(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// x1 = x + (acc0 + b_proj) and LN2(x1) on the 16 worker warps: thread = row (TMEM lane), worker j of the lane quarter owns
// columns 32 j .. 32 j + 31; on entry v holds the thread's slice of the residual x (zeros if there is none), on exit x1 -- it
// stays in registers until the tile is stored.  Two-pass like torch (mean, then squared deviations); the four workers of a
// quarter exchange their partial sums through shared memory (named barrier 8 + quarter, 128 threads).  Result: hi / lo of
// the normalised row as the A operand of fc1 in TMEM.
__device__ __forceinline__ void ln_worker(MSmem& sm, uint32_t tmem_lane, int q, int row, int j, int lane, float (&v)[32]) {
  {
    float u[32];
    tmem_ld32(tmem_lane + (uint32_t)(j * 32), u);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += u[i] + sm.bmid[j * 32 + i];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  sm.red[0][j][row] = s;
  asm volatile("bar.sync %0, %1;" ::"r"(8 + q), "r"(128) : "memory");
  const float mean = ((sm.red[0][0][row] + sm.red[0][1][row]) + (sm.red[0][2][row] + sm.red[0][3][row])) * (1.f / 128.f);
  float qq = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; qq = fmaf(d, d, qq); }
  sm.red[1][j][row] = qq;
  asm volatile("bar.sync %0, %1;" ::"r"(8 + q), "r"(128) : "memory");
  const float var = ((sm.red[1][0][row] + sm.red[1][1][row]) + (sm.red[1][2][row] + sm.red[1][3][row])) * (1.f / 128.f);
  const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int half16 = 0; half16 < 2; ++half16) {       // 16 columns at a time: x1 stays live, so keep the temporaries small
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = j * 32 + half16 * 16 + i;
      const float y = (v[half16 * 16 + i] - mean) * rstd * sm.gamma[k] + sm.beta[k];
      const float h = rna_tf32_fast(y);
      hi[i] = __float_as_uint(h);
      lo[i] = __float_as_uint(lo_tf32(y, h));
    }
    tmem_st16(tmem_lane + (uint32_t)(M_COL_ALN_HI + j * 32 + half16 * 16), hi);
    tmem_st16(tmem_lane + (uint32_t)(M_COL_ALN_LO + j * 32 + half16 * 16), lo);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive_m(&sm.aln_full);
}

// One hidden chunk (32 columns) on one of the 16 GELU workers (all producer and LN/store warps: 4 per TMEM lane quarter, worker j
// takes accumulator columns 8 j .. 8 j + 7 of its row): h = GELU(fc1 + b1) -> hi / lo -> the swizzled A-operand tile of fc2.
// The fc1 accumulator is released right after the TMEM load; the hidden buffer gc & 1 is written once the fc2 MMAs of chunk
// gc - 2 have read it.
__device__ __forceinline__ void gelu_worker(MSmem& sm, uint32_t tmem_lane, uint8_t* sH, int row, int j, uint32_t gc, const float* b1,
                                            int lane, long long* tp) {
  const int ab = gc & 3, hb = gc & 1;
  mtrace(tp, 2048 + gc * 8 + 0);
  mbar_wait_warp(&sm.acc1_full[ab], (gc >> 2) & 1);
  mtrace(tp, 2048 + gc * 8 + 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float v[8];
  tmem_ld8(tmem_lane + (uint32_t)(M_COL_X + ab * M_CH + j * 8), v);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive_m(&sm.acc1_empty[ab]);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = gelu_fast(v[i] + b1[i]);
  mtrace(tp, 2048 + gc * 8 + 2);
  if (gc >= 2) mbar_wait_warp(&sm.h_free[hb], ((gc >> 1) - 1) & 1);
  mtrace(tp, 2048 + gc * 8 + 3);
  uint8_t* hi_t = sH + hb * M_UNIT;
  uint8_t* lo_t = hi_t + M_TILE;
#pragma unroll
  for (int c2 = 0; c2 < 2; ++c2) {
    float4 h, l;
    h.x = rna_tf32_fast(v[c2 * 4]); h.y = rna_tf32_fast(v[c2 * 4 + 1]); h.z = rna_tf32_fast(v[c2 * 4 + 2]); h.w = rna_tf32_fast(v[c2 * 4 + 3]);
    l.x = lo_tf32(v[c2 * 4], h.x); l.y = lo_tf32(v[c2 * 4 + 1], h.y); l.z = lo_tf32(v[c2 * 4 + 2], h.z); l.w = lo_tf32(v[c2 * 4 + 3], h.w);
    const uint32_t so = swz(row, j * 2 + c2);
    *reinterpret_cast<float4*>(hi_t + so) = h;
    *reinterpret_cast<float4*>(lo_t + so) = l;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive_m(&sm.h_full[hb]);
  mtrace(tp, 2048 + gc * 8 + 4);
}

// this thread's 32-column slice of the residual row (zeros without a residual / past the last row).  Plain (L1-cached) loads:
// thread = row, so one instruction touches 32 different lines and the next one the neighbouring 16 bytes of the same lines
__device__ __forceinline__ void load_residual(const nmrf_mlp_args& a, int grow, int j, float (&v)[32]) {
  if (a.e_identity && grow < a.rows) {
    const float4* p = reinterpret_cast<const float4*>(a.E + (size_t)grow * a.lde + j * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) { const float4 t = p[c]; v[c * 4] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w; }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
  }
}
// v += this thread's 32 columns of acc0 (fp32 round-to-nearest adds, outside the tensor core)
__device__ __forceinline__ void add_acc0(uint32_t tmem_lane, int j, float (&v)[32]) {
  float u[32];
  tmem_ld32(tmem_lane + (uint32_t)(j * 32), u);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] += u[i];
}

// Everything a worker thread does for one tile after phase 1: x1 and LayerNorm, the 16 hidden chunks, the three drains of
// fc2's accumulator, the final sum and the store.  v: residual slice on entry.
__device__ __forceinline__ void worker_tile(MSmem& sm, const nmrf_mlp_args& a, uint32_t tmem_lane, uint8_t* sH, int q, int row, int j,
                                            int lane, int it, int grow, int rot, float (&v)[32], long long* tp) {
  mbar_wait_warp(&sm.p1_full, it & 1);
  mtrace(tp, 3968 + it * 8 + 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  ln_worker(sm, tmem_lane, q, row, j, lane, v);
  mtrace(tp, 3968 + it * 8 + 1);
#pragma unroll 1
  for (int c = 0; c < M_NCH; ++c) {
    gelu_worker(sm, tmem_lane, sH, row, j, (uint32_t)(it * M_NCH + c), sm.b1 + ((c + rot) & (M_NCH - 1)) * M_CH + j * 8, lane, tp);
    if (c >= 4 && (c & 3) == 0) {
      // fc2 units c-4 .. c-1 are complete (they were issued while this chunk went through GELU): move their partial sum out of
      // the tensor-core accumulator; the issuer restarts acc0 with unit c once all 16 warps have done so
      const uint32_t d = (uint32_t)it * 3 + (uint32_t)(c >> 2) - 1;
      mbar_wait_warp(&sm.part_full, d & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      add_acc0(tmem_lane, j, v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_m(&sm.part_drained);
    }
  }
  // ---- final: y = x1 + (fc2 partial of units 12..15 + b_fc2), straight from the registers (thread = row: 128 contiguous bytes)
  mbar_wait_warp(&sm.acc0_final, it & 1);
  mtrace(tp, 3968 + it * 8 + 2);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    float u[32];
    tmem_ld32(tmem_lane + (uint32_t)(j * 32), u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive_m(&sm.acc0_empty);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += u[i] + sm.bout[j * 32 + i];
  }
  if (grow < a.rows) {
    float4* dst = reinterpret_cast<float4*>(a.Y + (size_t)grow * a.ldy + j * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
  }
  if (a.out_stats) {
    // (mean, rstd) of the output row for the NEXT block's LayerNorm (nmrf_gemm_args.ln_stats): the row is in the registers of
    // the quarter's four workers right now -- two-pass like ln_worker, through the same exchange buffer and named barrier
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    sm.red[0][j][row] = s;
    asm volatile("bar.sync %0, %1;" ::"r"(8 + q), "r"(128) : "memory");
    const float mean = ((sm.red[0][0][row] + sm.red[0][1][row]) + (sm.red[0][2][row] + sm.red[0][3][row])) * (1.f / 128.f);
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; qq = fmaf(d, d, qq); }
    sm.red[1][j][row] = qq;
    asm volatile("bar.sync %0, %1;" ::"r"(8 + q), "r"(128) : "memory");
    if (j == 0 && grow < a.rows) {
      const float var = ((sm.red[1][0][row] + sm.red[1][1][row]) + (sm.red[1][2][row] + sm.red[1][3][row])) * (1.f / 128.f);
      reinterpret_cast<float2*>(a.out_stats)[grow] = make_float2(mean, 1.f / sqrtf(var + 1e-5f));
    }
  }
  mtrace(tp, 3968 + it * 8 + 3);
}

// 608 threads = 19 warps = up to 5 warps per register-file partition (16 384 registers): at most 96 registers per thread
__global__ void __launch_bounds__(M_BLOCK, 1)
mlp_chain_kernel(const nmrf_mlp_args a, int ntiles) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ MSmem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  auto sW = [&](int slot) { return base + slot * M_UNIT; };            // hi image at +0, lo image at +16 KB
  uint8_t* sH = base + M_OFF_H;                                        // two hidden buffers: [hi | lo] each
  auto sRaw = [&](int i) { return base + M_OFF_RAW + i * M_TILE; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef NMRF_TRACE
  long long* const tp = (blockIdx.x == 0 && (tid == M_MMA_WARP * 32 || tid == M_MMA2_WARP * 32 || tid == M_EPI_WARP0 * 32)) ? g_trace_m : nullptr;
#else
  long long* const tp = nullptr;
#endif
  // phase 1 per tile: n1 weight units against the k-blocks of X (and of E when it is an ordinary concatenated operand; a
  // RESIDUAL E -- e_identity -- never enters the tensor core: the workers add it in registers)
  const int n1 = (a.e_identity ? a.Kx : a.Kx + a.Ke) / M_BK;
  const int upt = n1 + 2 * M_NCH;                  // weight units per tile
  const int tstep = gridDim.x;
  // every CTA starts at its own rotation of the k-block order (phase 1) and of the hidden-chunk order (phase 3), so that the
  // 148 CTAs do not all pull the same 32 KB of the weight stream at the same time; both are sums: only the fp32 summation
  // order depends on the CTA
  const int rot = blockIdx.x & (M_NCH - 1);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < M_NB; ++i) { mbar_init(&sm.done[i], 1); mbar_init(&sm.full_b[i], 1); }
    mbar_init(&sm.p1_full, 1); mbar_init(&sm.aln_full, M_WORKERS);
    for (int i = 0; i < 4; ++i) { mbar_init(&sm.acc1_full[i], 1); mbar_init(&sm.acc1_empty[i], M_WORKERS); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.h_full[i], M_WORKERS); mbar_init(&sm.h_free[i], 1); }
    mbar_init(&sm.acc0_final, 1); mbar_init(&sm.acc0_empty, M_WORKERS);
    mbar_init(&sm.part_full, 1); mbar_init(&sm.part_drained, M_WORKERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 128) { sm.gamma[tid] = a.ln_gamma[tid]; sm.beta[tid] = a.ln_beta[tid]; sm.bmid[tid] = a.bias_mid[tid]; sm.bout[tid] = a.bias_out[tid]; }
  if (tid < M_HID) sm.b1[tid] = a.b1[tid];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp < 8) {
    // =============================================== producers (phase 1) + workers 0, 1 ===============================================
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
        f_e[j] = (a.E && !a.e_identity) ? a.E + (size_t)gr * a.lde + f_c * 4 - a.Kx : a.X;
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
    uint32_t pu = 0;          // phase-1 units produced so far by this CTA: ring stage pu % 3, A buffer j % 4, hand-off barrier 1 + pu % 3
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
      for (int j = 0; j < n1; ++j, ++pu) {
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
          for (int e = 0; e < 4; ++e) {
            const float h = rna_tf32_fast(vv[e]);
            hi[cc * 4 + e] = __float_as_uint(h);
            lo[cc * 4 + e] = __float_as_uint(lo_tf32(vv[e], h));
          }
        }
        // TMEM columns written below were last touched by the previous tile: its MMAs (A buffers alias the fc1 accumulators
        // and LN2(x1)) and its final read of acc0 must be done
        if (it > 0 && j == 0) {
          mbar_wait_warp(&sm.acc0_final, (it - 1) & 1);
          mbar_wait_warp(&sm.acc0_empty, (it - 1) & 1);
        }
        // A buffer j % 4 (TMEM columns of the fc1 accumulators and of LN2(x1), both idle in phase 1): free once the MMAs of
        // the weight unit four back are complete.  The producers wait for the unit THREE back: the `done` barriers rotate
        // with the three weight slots, and a wait further back could miss its phase (the barrier would be two ahead).
        if (j >= M_NB) {
          const uint32_t g = (uint32_t)it * upt + j - M_NB;
          mbar_wait_warp(&sm.done[g % M_NB], (g / M_NB) & 1);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem + a_lane + abuf_col(j % M_ABUF) + (uint32_t)(a_c0 * 4);
        tmem_st16(ta, hi);
        tmem_st16(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + pu % 3), "r"(M_HANDOFF) : "memory");
      }
      // LayerNorm and phase 3: the producers are workers 0 and 1 of their lane quarter
      float v[32];
      load_residual(a, t * M_BM + a_row, warp >> 2, v);
      worker_tile(sm, a, tmem + a_lane, sH, warp & 3, a_row, warp >> 2, lane, it, t * M_BM + a_row, rot, v, nullptr);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == M_MMA_WARP || warp == M_MMA2_WARP) {
    // =============================================== MMA issuers ===============================================
    const uint32_t idesc128 = idesc_n(128), idesc32 = idesc_n(32);
    const uint32_t acc0 = tmem;
    uint32_t g = 0;           // global unit counter (weight ring)
    uint32_t pu = 0;          // phase-1 unit counter (A buffers / hand-off barriers)
    uint32_t gc = 0;          // global hidden-chunk counter of the tile's first chunk
    auto wait_b = [&](uint32_t unit) { mbar_wait_warp(&sm.full_b[unit % M_NB], (unit / M_NB) & 1); };
    // F1 unit: fc1 rows of one hidden chunk (N = 32) over the whole K = 128 into a FRESH accumulator; A = LN2(x1) from TMEM.
    // The small products first: while the accumulator is small its round-toward-zero update loses nothing that matters.
    auto issue_f1 = [&](uint32_t cg) {
      if (cg >= 4) {                                                       // accumulator cg & 3 drained (chunk cg - 4)
        mbar_wait_warp(&sm.acc1_empty[cg & 3], ((cg >> 2) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      mtrace(tp, g * 4 + 0);
      wait_b(g);
      mtrace(tp, g * 4 + 1);
      if (elect_one()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bslot = smem_u32(sW(g % M_NB));
        const uint32_t d = tmem + (uint32_t)(M_COL_X + (cg & 3) * M_CH);
        // k-block kk / 4: a [32 n x 32 k] sub-image (4 KB) of the unit's hi / lo image; 32-byte step kk % 4 inside its rows
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
          umma_tf32_ta(d, tmem + M_COL_ALN_LO + kk * 8, make_desc(bslot + (kk >> 2) * 4096) + (uint64_t)((kk & 3) * 2), idesc32, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
          umma_tf32_ta(d, tmem + M_COL_ALN_HI + kk * 8, make_desc(bslot + M_TILE + (kk >> 2) * 4096) + (uint64_t)((kk & 3) * 2), idesc32, 1u);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
          umma_tf32_ta(d, tmem + M_COL_ALN_HI + kk * 8, make_desc(bslot + (kk >> 2) * 4096) + (uint64_t)((kk & 3) * 2), idesc32, 1u);
        umma_commit(&sm.done[g % M_NB]);
        umma_commit(&sm.acc1_full[cg & 3]);
      }
      __syncwarp();
      mtrace(tp, g * 4 + 2);
      ++g;
    };
    // F2 unit: the 128 fc2 rows against one hidden chunk (K = 32); A = GELU'd hidden chunk from shared memory.  acc0 is restarted
    // with units 0, 4, 8, 12 of a tile (the workers have moved the previous partial sum into their registers) and handed to
    // the workers after units 3, 7, 11 (part_full) and 15 (acc0_final).
    auto issue_f2 = [&](uint32_t cg, int c, uint32_t it) {
      const bool fresh = (c & 3) == 0;
      mbar_wait_warp(&sm.h_full[cg & 1], (cg >> 1) & 1);
      if (fresh && c > 0) {
        const uint32_t d = it * 3 + (uint32_t)(c >> 2) - 1;
        mbar_wait_warp(&sm.part_drained, d & 1);
      }
      mtrace(tp, g * 4 + 0);
      wait_b(g);
      mtrace(tp, g * 4 + 1);
      if (elect_one()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bslot = smem_u32(sW(g % M_NB));
        const uint64_t dBh = make_desc(bslot), dBl = make_desc(bslot + M_TILE);
        const uint32_t hbuf = smem_u32(sH + (cg & 1) * M_UNIT);
        const uint64_t dAh = make_desc(hbuf), dAl = make_desc(hbuf + M_TILE);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_tf32(acc0, dAl + (uint64_t)(ks * 2), dBh + (uint64_t)(ks * 2), idesc128, (fresh && ks == 0) ? 0u : 1u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_tf32(acc0, dAh + (uint64_t)(ks * 2), dBl + (uint64_t)(ks * 2), idesc128, 1u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_tf32(acc0, dAh + (uint64_t)(ks * 2), dBh + (uint64_t)(ks * 2), idesc128, 1u);
        umma_commit(&sm.done[g % M_NB]);
        umma_commit(&sm.h_free[cg & 1]);
        if ((c & 3) == 3 && c < M_NCH - 1) umma_commit(&sm.part_full);
        if (c == M_NCH - 1) umma_commit(&sm.acc0_final);
      }
      __syncwarp();
      mtrace(tp, g * 4 + 2);
      ++g;
    };
    // position of a phase-3 unit in the tile's stream: F1(c) and F2(c) interleave as F1(0) F1(1) F1(2) [F1(c+3) F2(c)] ... F2(13..15)
    auto pos_f1 = [&](int c) { return n1 + (c < M_AHEAD ? c : M_AHEAD + 2 * (c - M_AHEAD)); };
    auto pos_f2 = [&](int c) { return n1 + (c <= M_NCH - 1 - M_AHEAD ? M_AHEAD + 2 * c + 1 : c + M_NCH); };
    int it = 0;
    if (warp == M_MMA_WARP) {
      for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
        if (it > 0) mbar_wait_warp(&sm.acc0_empty, (it - 1) & 1);         // previous tile's last partial sum has been read out of acc0
        // ---- phase 1: acc0 = [X | E] . W1cat^T  (the update only: no bias, no residual)
        g = (uint32_t)it * upt;
        for (int j = 0; j < n1; ++j, ++pu) {
          mtrace(tp, g * 4 + 0);
          asm volatile("bar.sync %0, %1;" ::"r"(1 + pu % 3), "r"(M_HANDOFF) : "memory");
          mtrace(tp, g * 4 + 3);
          wait_b(g);
          mtrace(tp, g * 4 + 1);
          if (elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bslot = smem_u32(sW(g % M_NB));
            const uint64_t dBh = make_desc(bslot), dBl = make_desc(bslot + M_TILE);
            const uint32_t tAh = tmem + abuf_col(j % M_ABUF), tAl = tAh + 32;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_tf32_ta(acc0, tAl + ks * 8, dBh + (uint64_t)(ks * 2), idesc128, (j > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_tf32_ta(acc0, tAh + ks * 8, dBl + (uint64_t)(ks * 2), idesc128, 1u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_tf32_ta(acc0, tAh + ks * 8, dBh + (uint64_t)(ks * 2), idesc128, 1u);
            umma_commit(&sm.done[g % M_NB]);
            if (j == n1 - 1) umma_commit(&sm.p1_full);
          }
          __syncwarp();
          mtrace(tp, g * 4 + 2);
          ++g;
        }
        // ---- phase 3, fc2 half: F2(0..15), each as soon as its hidden chunk is in shared memory
        for (int c = 0; c < M_NCH; ++c) {
          g = (uint32_t)it * upt + pos_f2(c);
          issue_f2(gc + c, c, (uint32_t)it);
        }
        gc += M_NCH;
      }
    } else {
      // ---- phase 3, fc1 half (warp 18): F1(0..15), up to four accumulators ahead of the GELU workers
      for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
        mbar_wait_warp(&sm.aln_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < M_NCH; ++c) {
          g = (uint32_t)it * upt + pos_f1(c);
          issue_f1(gc + c);
        }
        gc += M_NCH;
      }
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
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(M_UNIT) : "memory");
          // source unit: the stream is laid out P1(0..n1-1) | F1(0..15) | F2(0..15); issue order of phase 3 as above, with the
          // chunk index rotated by `rot`
          int su;
          if (u < n1) {
            su = (u + rot) % n1;
          } else {
            const int s3 = u - n1;
            int cs;
            bool is_f2;
            if (s3 < M_AHEAD) { is_f2 = false; cs = s3; }
            else if (s3 < M_AHEAD + 2 * (M_NCH - M_AHEAD)) { const int j = s3 - M_AHEAD; is_f2 = j & 1; cs = is_f2 ? (j >> 1) : (j >> 1) + M_AHEAD; }
            else { is_f2 = true; cs = s3 - M_NCH; }
            su = n1 + (is_f2 ? M_NCH : 0) + ((cs + rot) & (M_NCH - 1));
          }
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(sW(slot))), "l"(a.Wstream + (size_t)su * (M_UNIT / 4)), "r"(M_UNIT), "r"(bar) : "memory");
        }
      }
    }
  } else {
    // =============================================== workers 2, 3 of every lane quarter ===============================================
    const int e = warp - M_EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int j = 2 + (e >> 2);                // worker index: columns [32 j, 32 j + 32)
    const int row = q * 32 + lane;             // row of the tile owned by this thread
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
      float v[32];
      load_residual(a, t * M_BM + row, j, v);  // in flight during phase 1
      worker_tile(sm, a, tmem + t_lane, sH, q, row, j, lane, it, t * M_BM + row, rot, v, tp);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

}  // namespace

#ifdef NMRF_TRACE
int mlp_set_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_trace_m, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? NMRF_OK : NMRF_ERR_CUDA;
}
#endif

int mlp_chain(const nmrf_mlp_args& a, cudaStream_t stream) {
  const int num_sms = nmrf::num_sms();
  static PerDevice configured;
  ensure_dynamic_smem(mlp_chain_kernel, M_DYN, configured);
  const int ntiles = (a.rows + M_BM - 1) / M_BM;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  mlp_chain_kernel<<<grid, M_BLOCK, M_DYN, stream>>>(a, ntiles);
  count_launch();
  return check_launch("mlp_chain");
}

}  // namespace nmrf
