// Fused token GEMM on tcgen05, warp-specialised, A operand in TENSOR MEMORY (v6).
//
// Why: a cycle trace of v5 (tools/gemm_trace.py, profiles/) showed the MMA warp needing ~2100 cycles to issue the 12
// UMMAs of a unit whose tensor work is 768 cycles: every SS-mode tf32 UMMA (128x128x8) reads 4 KB of A and 4 KB of B from
// shared memory, so the MMAs alone saturate the shared-memory pipe while the producers' st.shared / cp.async traffic
// competes with them.  Here the producers write the LayerNorm'ed, concatenated, hi/lo-split A operand straight into
// TMEM (tcgen05.st) and the UMMA takes A from TMEM: shared memory carries the weight tiles (B) and the raw A ring.
//
//   warps 0-7   producers   thread = (row, half of the 32-wide k-block): warp w owns TMEM lane quarter w%4 and k-columns
//                           16*(w/4)..+16.  The raw A tile [128 x 32] of a k-block comes into a 3-deep swizzled smem ring by
//                           COALESCED cp.async (8 lanes = one row's 128 B; a per-thread row gather costs 32 L1 wavefronts
//                           per instruction); the ring streams ACROSS tiles, two k-blocks ahead.  Per k-block: LDS ->
//                           LayerNorm -> hi/lo split in registers, THEN wait for the MMAs that still read the A buffer,
//                           then tcgen05.st (A_hi / A_lo -> TMEM) and hand over to the MMA warp.
//   warp  8     MMA issuer  per unit: 4 k-steps x {A_lo.B_hi, A_hi.B_lo, A_hi.B_hi}, A from TMEM, B from swizzled smem
//   warps 9-16  epilogue    TMEM -> registers -> smem transpose -> coalesced bias/activation/residual/store; while they
//                           wait they also compute the LayerNorm statistics of the tile three ahead (the producers used
//                           to stall ~10k cycles per tile on those loads)
//   warp  17    TMA         one lane streams the pre-swizzled weight tile images (2 x 16 KB cp.async.bulk per unit) into a
//                           4-slot ring, as soon as the MMAs that read a slot have completed (a fourth slot instead of a
//                           fourth raw-A stage: the MMA warp waited ~300 cycles per unit for weights with three)
//
//   tile        128 rows x 128 output columns, unit = one k-block of 32; A tiles double buffered (2 x 64 columns).
//               TMEM map: [0,384) three accumulator stages, [384,512) A: buffer b at 384 + 64 b, hi at +0, lo at +32.
//   accumulation every GROUP of G = 2..4 units (half a tile's k-blocks) gets a fresh accumulator stage (ring of three): at
//               most 12 G MMAs -- per k-block the eight small lo.hi / hi.lo products first, then the four hi.hi ones -- ever
//               accumulate into the same tensor-memory columns, and the epilogue warps add the groups' partial sums in fp32
//               registers (round-to-nearest).  The tensor core updates its accumulator with round-toward-zero
//               (tools/precision_probe.py: relative bias -1.6e-8 per MMA, -7.7e-7 for a K = 128 product accumulated in place,
//               5x the rms error of fp32 FMAs): summed in place, the projections were the noisiest arithmetic of the whole
//               forward and flipped argmax / median decisions downstream.  (One stage per unit measured the same accuracy but
//               let the MMA warp run only 3 units ahead of a tile's store phase: 84 instead of 35 us per launch.)
//   waiting     all 32 lanes of a warp poll an mbarrier (tc_common.cuh: mbar_wait_warp): the warp stays converged.
#include <cstdlib>
#include "common.cuh"
#include "tmap.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int G6_BM = 128, G6_BN = 128, G6_BK = 32, G6_NB = 4, G6_ACC = 3;
constexpr int G6_ACOL = G6_ACC * G6_BN;             // first TMEM column of the A buffers
constexpr int G6_TILE = G6_BM * G6_BK * 4;          // 16 KB operand tile
constexpr int G6_PROD = 256;                        // producer threads (warps 0-7)
constexpr int G6_MMA_WARP = 8;
constexpr int G6_EPI_WARP0 = 9, G6_EPI_WARPS = 8;
constexpr int G6_TMA_WARP = G6_EPI_WARP0 + G6_EPI_WARPS;       // 17
constexpr int G6_BLOCK = (G6_TMA_WARP + 1) * 32;               // 576
constexpr int G6_HANDOFF = G6_PROD + 32;            // named barriers 1..3: producers arrive, MMA warp syncs
constexpr int G6_RAW_BAR = 5;                       // named barrier of the producers: raw tile visible / consumed
constexpr int G6_STAGE_FLOATS = 32 * 36;            // per-epilogue-warp transpose tile
constexpr int G6_RAW = 3;                           // raw-A ring depth (k-blocks)
constexpr int G6_STATS = 4;                         // LayerNorm statistics buffers (tiles in flight: 3 ahead)
constexpr int G6_DYN = (2 * G6_NB + G6_RAW) * G6_TILE + G6_EPI_WARPS * G6_STAGE_FLOATS * 4 + 1024;   // B_hi[4] B_lo[4] raw[3] + epilogue staging

struct G6Smem {
  uint64_t done[G6_NB];       // MMAs of the unit that used B slot s are complete (tcgen05.commit)
  uint64_t full_b[G6_NB];     // the slot's two weight tiles have landed (TMA bulk copy, expect_tx 32 KB)
  uint64_t acc_full[G6_ACC];  // accumulator stage holds a finished tile (tcgen05.commit)
  uint64_t acc_empty[G6_ACC]; // epilogue has drained the stage (256 arrivals)
  uint64_t slab_full[2];      // SLAB mode: the raw slab has landed in the stage (TMA tensor copy, expect_tx 130 x 128 B)
  uint64_t slab_free[2];      // SLAB mode: every producer warp is past its reads of the stage (8 arrivals)
  uint64_t stats_full[G6_STATS];          // LayerNorm statistics of a tile are in mean/rstd (8 arrivals: one per epilogue warp)
  uint32_t tmem_base;
  float mean[G6_STATS][G6_BM], rstd[G6_STATS][G6_BM];
  alignas(16) float gamma[128];           // LayerNorm affine (Kx == 128), staged once; read as float4
  alignas(16) float beta[128];
};

// CONV mode: the same kernel as an implicit-GEMM convolution over an NHWC image (N1 / N2 of the scope table: the k x k
// convolutions of the feature extractor and of the conv heads; reference nmrf/models/backbone.py:48-98, NMRF.py:56-65,
// DPN.py:45-49).  GEMM row = output pixel (n, yo, xo); k-block kb = (tap, 32-channel block): tap = kb / cpb = ky * kw + kx,
// c0 = 32 (kb % cpb).  The raw A tile of a k-block is, per row, the 128 contiguous bytes at input pixel
// (yo * stride - pad + ky, xo * stride - pad + kx), channel c0 -- zero-filled (cp.async src-size 0) outside the image, which
// IS the convolution's zero padding.  Everything else (hi/lo split, MMAs, grouped accumulation, epilogue) is unchanged.
struct ConvGeom {
  int H, W;                  // input extent (bounds of the taps)
  int pix_stride, row_stride;   // floats between horizontally / vertically adjacent input pixels
  long long img_stride;      // floats between samples
  int Ho, Wo;                // output extent: rows = N * Ho * Wo
  int kw, stride, pad, cpb;  // kernel width, stride, padding, k-blocks (32 channels) per tap
  int slab;                  // 3x3, stride 1, pad 1 over a dense NHWC tensor: the three taps of a kernel row share one raw slab
};
// SLAB mode.  These kernels run at the L2 -> SM ceiling (~30 B / cycle / SM measured, LTS cap ~42), and a 3x3 convolution
// fetched tap by tap reads its input nine times.  With stride 1 / pad 1 the GEMM row index IS the flattened input pixel
// index, so row r of a tile reads, for tap (ky, kx), flattened pixel p0 + r + (ky - 1) W + (kx - 1): the three kx taps of a
// kernel row are the same 130 contiguous pixels shifted by one.  One slab [130 px x 32 ch] per (ky, channel block) serves
// three units (k-block order ky, channel block, kx); what the flattening gets wrong (row / image borders) is exactly the
// zero padding and is masked per row.  Two slab stages (17 KB each) live in the raw ring's 48 KB.  A slab is ONE TMA tensor
// copy (2-D tensor map over [pixels, channels], box 130 x 32, hardware 128-byte swizzle = swz(), rows outside the tensor
// zero-filled) issued by a lane of the TMA warp: the producers' own cp.async fetch -- address arithmetic, wait_group and a
// 256-thread barrier per unit -- measured 15-20 % of a 3x3 convolution (tools/conv_bench.py with the fetch compiled out).
constexpr int G6_SLAB_ROWS = G6_BM + 2, G6_SLAB_BYTES = 136 * 128;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// optional cycle trace of CTA 0 (debug builds only: make TRACE=1 + nmrf_debug_set_trace); slot layout in tools/gemm_trace.py
#ifdef NMRF_TRACE
__device__ long long* g_trace6 = nullptr;
#endif
#define trace(tp, idx) NMRF_TRACE_STAMP(tp, idx)

// LayerNorm statistics of 16 rows of a tile (Kx == 128) by one warp: 8 lanes per row (16 floats each, every load
// instruction covers four rows' contiguous 128 B), four rows per pass, all 16 loads of a lane in flight at once, two
// 3-stage shuffle reductions (two-pass variance, like the reference's LayerNorm).  A whole-warp-per-row version chained
// 160 dependent shuffles and took ~11k cycles per call.  Out of line: the warp-specialised kernel's instruction footprint
// matters (four roles share the instruction cache).
__device__ __noinline__ void tile_stats6(const float* __restrict__ X, int ldx, int rows, int row0, int e, float* mean, float* rstd) {
  const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
  float4 v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + e * 16 + i * 4 + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v[i][j] = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + j * 32 + sub * 4);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += (v[i][j].x + v[i][j].y) + (v[i][j].z + v[i][j].w);
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float mu = s * (1.f / 128.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dx = v[i][j].x - mu, dy = v[i][j].y - mu, dz = v[i][j].z - mu, dw = v[i][j].w - mu;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
    if (sub == 0) { const int lr = e * 16 + i * 4 + g; mean[lr] = mu; rstd[lr] = 1.f / sqrtf(q * (1.f / 128.f) + 1e-5f); }
  }
}

// tile t -> (row block, 128-column chunk): column-chunk-major, so concurrently running CTAs share a weight chunk (L2)
// 576 threads = 18 warps = up to 5 warps on one of the SM's four register-file partitions (16 384 registers each): at most
// 96 registers per thread, which __launch_bounds__ makes ptxas respect (112 "fits" 65 536 / 576 but fails to launch)
template <int ACT, bool LN, bool CONV>
__global__ void __launch_bounds__(G6_BLOCK, 1)
token_gemm_tc6_kernel(const nmrf_gemm_args a, int n_rb, int n_nc, const ConvGeom cg, const __grid_constant__ CUtensorMap amap) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ G6Smem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  auto sB_hi = [&](int i) { return base + i * G6_TILE; };
  auto sB_lo = [&](int i) { return base + (G6_NB + i) * G6_TILE; };
  auto sRaw = [&](int i) { return base + (2 * G6_NB + i) * G6_TILE; };
  float* stage_base = reinterpret_cast<float*>(base + (2 * G6_NB + G6_RAW) * G6_TILE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef NMRF_TRACE
  long long* const tp = (blockIdx.x == 0 && (tid == 0 || tid == G6_MMA_WARP * 32 || tid == G6_EPI_WARP0 * 32)) ? g_trace6 : nullptr;
  long long* const tp2 = (blockIdx.x == 0 && tid == 32) ? g_trace6 : nullptr;
#endif
  const int Ktot = a.Kx + a.Ke;
  const int nkb = (Ktot + G6_BK - 1) / G6_BK;
  const int ntiles = n_rb * n_nc;
  const int tstep = gridDim.x;
  // units per accumulator stage: half the tile's k-blocks, between 2 and 4; groups per tile
  const int G = min(4, max(2, (nkb + 1) / 2));
  const int ngrp = (nkb + G - 1) / G;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < G6_NB; ++i) { mbar_init(&sm.done[i], 1); mbar_init(&sm.full_b[i], 1); }
    for (int i = 0; i < G6_ACC; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], G6_EPI_WARPS * 32); }
    for (int i = 0; i < G6_STATS; ++i) mbar_init(&sm.stats_full[i], G6_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.slab_full[i], 1); mbar_init(&sm.slab_free[i], G6_PROD / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (LN && tid < 128) { sm.gamma[tid] = a.ln_gamma[tid]; sm.beta[tid] = a.ln_beta[tid]; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp < 8) {
    // =============================================== producers ===============================================
    // warp w may touch TMEM lanes 32*(w%4)..+32 only: thread -> row 32*(w%4)+lane, k-columns 16*(w/4)..+16 of the k-block
    const int a_row = (warp & 3) * 32 + lane, a_c0 = (warp >> 2) * 4;
    const uint32_t a_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    // Raw-A ring, streaming across tiles.  fetch(): k-block f_kb of tile f_t -> stage `stage`: 1024 16-byte chunks, 4 per
    // thread, lanes 8j..8j+7 cover one row's 128 B; chunk c of row r lands at position c ^ (r & 7) (conflict-free row reads
    // later); rows / columns outside the problem are zero-filled (src-size 0).  Past the last tile: an empty group, so the
    // wait_group accounting stays uniform (one group per unit).
    const int f_c = tid & 7, f_r = tid >> 3;
    int f_t = blockIdx.x, f_kb = 0;
    const float* f_x[4];       // this thread's four source rows of the fetch cursor's tile (X part, E part), refreshed per tile:
    const float* f_e[4];       // the index arithmetic (a modulo, a division by ediv) stays out of the per-unit path
    uint32_t f_ok = 0;
    int f_yx[4];               // CONV: (yo * stride - pad) << 16 | (xo * stride - pad) & 0xffff of the four rows
    int f_tap = 0, f_cb = 0;   // CONV: tap and channel block of the fetch cursor's k-block
    const bool slab = CONV && cg.slab;
    auto fetch_tile = [&]() {
      const int row0 = (f_t % n_rb) * G6_BM;
      f_ok = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int grow = row0 + f_r + 32 * j;
        const bool ok = grow < a.rows;
        f_ok |= (ok ? 1u : 0u) << j;
        const int gr = ok ? grow : 0;
        if (CONV) {
          const int per = cg.Ho * cg.Wo;
          const int n = gr / per, rem = gr - n * per;
          const int yo = rem / cg.Wo, xo = rem - yo * cg.Wo;
          const int yi = yo * cg.stride - cg.pad, xi = xo * cg.stride - cg.pad;
          f_yx[j] = (yi << 16) | (xi & 0xffff);
          f_x[j] = a.X + (long long)n * cg.img_stride + (long long)yi * cg.row_stride + (long long)xi * cg.pix_stride + f_c * 4;
          f_e[j] = a.X;
        } else {
          f_x[j] = a.X + (size_t)gr * a.ldx + f_c * 4;
          f_e[j] = a.E ? a.E + (size_t)(gr / a.ediv) * a.lde + f_c * 4 - a.Kx : a.X;
        }
      }
    };
    if (f_t < ntiles && !slab) fetch_tile();
    auto fetch_next = [&](uint32_t stage) {
      if (f_t < ntiles) {
        const uint32_t dst = smem_u32(sRaw(stage));
        if (CONV) {
          const int ky = f_tap / cg.kw, kx = f_tap - ky * cg.kw;
          const long long off = (long long)ky * cg.row_stride + (long long)kx * cg.pix_stride + f_cb * G6_BK;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int yi = (f_yx[j] >> 16) + ky, xi = (int)(short)(f_yx[j] & 0xffff) + kx;
            const bool ok = ((f_ok >> j) & 1u) && (unsigned)yi < (unsigned)cg.H && (unsigned)xi < (unsigned)cg.W;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + swz(f_r + 32 * j, f_c)), "l"(ok ? f_x[j] + off : a.X), "r"(ok ? 16 : 0));
          }
          if (++f_cb == cg.cpb) { f_cb = 0; ++f_tap; }
        } else {
          const int k0 = f_kb * G6_BK;
          const bool in_x = k0 + f_c * 4 < a.Kx, in_k = k0 + f_c * 4 < Ktot;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok = in_k && ((f_ok >> j) & 1u);
            const float* src = (in_x ? f_x[j] : f_e[j]) + k0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + swz(f_r + 32 * j, f_c)), "l"(ok ? src : a.X), "r"(ok ? 16 : 0));
          }
        }
        if (++f_kb == nkb) { f_kb = 0; f_tap = 0; f_cb = 0; f_t += tstep; if (f_t < ntiles) fetch_tile(); }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (!slab) { fetch_next(0); fetch_next(1); }
    uint32_t gslab = 0, tapmask = 0;   // SLAB: slabs consumed so far; bit ky * 3 + kx = this thread's row has that tap inside the image
    uint32_t unit = 0;         // == k-blocks produced so far by this CTA: ring stage unit % 4, A buffer unit & 1, B slot unit % 3
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
      const int row0 = (t % n_rb) * G6_BM;
      const bool row_ok = row0 + a_row < a.rows;
      float mean = 0.f, rstd = 1.f;
      if (LN) {
        if (a.ln_stats) {
          // the kernel that wrote X also wrote each row's (mean, rstd) (nmrf_mlp_chain / nmrf_row_stats): 8 bytes per row
          // instead of a statistics pass over the tile by the epilogue warps (which was the bottleneck of this kernel: ~13k
          // cycles per tile, once per 128-column chunk, against ~6k cycles of MMA work)
          if (row_ok) {
            const float2 st = __ldg(reinterpret_cast<const float2*>(a.ln_stats) + row0 + a_row);
            mean = st.x; rstd = st.y;
          }
        } else {
          mbar_wait_warp(&sm.stats_full[it % G6_STATS], (it / G6_STATS) & 1);
          mean = sm.mean[it % G6_STATS][a_row]; rstd = sm.rstd[it % G6_STATS][a_row];
        }
      }
      if (slab) {
        tapmask = 0;
        if (row_ok) {
          const int rem = (row0 + a_row) % (cg.H * cg.W);
          const int yo = rem / cg.W, xo = rem - yo * cg.W;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              if ((unsigned)(yo + ky - 1) < (unsigned)cg.H && (unsigned)(xo + kx - 1) < (unsigned)cg.W) tapmask |= 1u << (ky * 3 + kx);
        }
      }
      int s_kx = 0, s_cb = 0, s_ky = 0;   // SLAB: tap column, channel block and kernel row of the unit
      for (int kb = 0; kb < nkb; ++kb, ++unit) {
        const int slot = unit % G6_NB;
        trace(tp, unit * 8 + 0);
        const uint8_t* raw;
        int rrow = a_row;
        bool live = true;
        bool slab_done = false;        // SLAB: this unit is the last reader of its slab
        if (slab) {
          if (s_kx == 0) mbar_wait_warp(&sm.slab_full[gslab & 1], (gslab >> 1) & 1);   // the slab of this kernel row has landed (TMA)
          raw = sRaw(0) + (gslab & 1) * G6_SLAB_BYTES;
          rrow = a_row + s_kx;
          live = (tapmask >> (s_ky * 3 + s_kx)) & 1u;
          slab_done = s_kx == 2;
        } else {
          // this k-block's raw tile has landed for every producer thread (own copies: wait_group; others': barrier)
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          asm volatile("bar.sync %0, %1;" ::"r"(G6_RAW_BAR), "r"(G6_PROD) : "memory");
          // the stage consumed one unit ago is free again (every producer is past its reads): re-arm it two k-blocks ahead
          fetch_next((unit + 2) % G6_RAW);
          raw = sRaw(unit % G6_RAW);
        }
        trace(tp2, 1024 + unit * 8 + 0);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = a_c0 + cc;
          const int kk = kb * G6_BK + c * 4;
          float4 v = *reinterpret_cast<const float4*>(raw + swz(rrow, c));
          if (CONV && !live) v = make_float4(0.f, 0.f, 0.f, 0.f);
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
        if (slab) {
          if (slab_done) {               // the slab's three taps are in registers: hand the stage back to the TMA lane
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.slab_free[gslab & 1]);
            ++gslab; s_kx = 0;
            if (++s_cb == cg.cpb) { s_cb = 0; ++s_ky; }
          } else {
            ++s_kx;
          }
        }
        trace(tp2, 1024 + unit * 8 + 1);
        // the MMAs of unit-2 (and, cumulatively, all earlier ones) are complete: A buffer unit & 1 may be overwritten
        if (unit >= 2) {
          mbar_wait_warp(&sm.done[(unit - 2) % G6_NB], ((unit - 2) / G6_NB) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        trace(tp, unit * 8 + 1);
        trace(tp2, 1024 + unit * 8 + 2);
        const uint32_t ta = tmem + a_lane + (uint32_t)(G6_ACOL + (unit & 1) * 64 + a_c0 * 4);
        tmem_st16(ta, hi);
        tmem_st16(ta + 32, lo);
        trace(tp2, 1024 + unit * 8 + 3);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // A written with tcgen05.st -> ordered before the hand-off
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + slot), "r"(G6_HANDOFF) : "memory");
        trace(tp, unit * 8 + 2);
        trace(tp2, 1024 + unit * 8 + 4);
        trace(tp2, 1024 + unit * 8 + 5);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == G6_MMA_WARP) {
    // =============================================== MMA issuer ===============================================
    uint32_t unit = 0, grp = 0;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
      const int n_base = (t / n_rb) * G6_BN;
      const uint32_t idesc = make_idesc(min(G6_BN, a.N - n_base));
      for (int kb = 0; kb < nkb; ++kb, ++unit) {
        const int slot = unit % G6_NB;
        const int as = grp % G6_ACC;            // accumulator stage of this group of units
        const bool first = kb % G == 0, last = (kb % G == G - 1) || kb == nkb - 1;
        if (first && grp >= G6_ACC) mbar_wait_warp(&sm.acc_empty[as], ((grp / G6_ACC) - 1) & 1);   // the epilogue has drained group - 3
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        trace(tp, 2048 + unit * 4 + 0);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(G6_HANDOFF) : "memory");      // A of this unit is in TMEM
        trace(tp, 2048 + unit * 4 + 3);
        mbar_wait_warp(&sm.full_b[slot], (unit / G6_NB) & 1);                             // B tiles have landed
        trace(tp, 2048 + unit * 4 + 1);
        if (elect_one()) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dBh = make_desc(smem_u32(sB_hi(slot))), dBl = make_desc(smem_u32(sB_lo(slot)));
          const uint32_t d = tmem + (uint32_t)(as * G6_BN);
          const uint32_t tAh = tmem + (uint32_t)(G6_ACOL + (unit & 1) * 64), tAl = tAh + 32;
          // k-step ks: B +32 bytes inside the swizzle row, A +8 TMEM columns.  Small products first (fresh accumulator).
#pragma unroll
          for (int ks = 0; ks < G6_BK / 8; ++ks) umma_tf32_ta(d, tAl + ks * 8, dBh + (uint64_t)(ks * 2), idesc, (ks > 0 || !first) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < G6_BK / 8; ++ks) umma_tf32_ta(d, tAh + ks * 8, dBl + (uint64_t)(ks * 2), idesc, 1u);
#pragma unroll
          for (int ks = 0; ks < G6_BK / 8; ++ks) umma_tf32_ta(d, tAh + ks * 8, dBh + (uint64_t)(ks * 2), idesc, 1u);
          umma_commit(&sm.done[slot]);
          if (last) umma_commit(&sm.acc_full[as]);
        }
        __syncwarp();
        if (last) ++grp;
        trace(tp, 2048 + unit * 4 + 2);
      }
    }
  } else if (warp == G6_TMA_WARP) {
    // =============================================== weight tiles (TMA) ===============================================
    if (lane == 1 && CONV && cg.slab) {
      // raw slabs of the A operand: slab gs -> stage gs & 1, one ahead of the producers (they free a stage after its third tap)
      uint32_t gs = 0;
      const uint32_t dst0 = smem_u32(sRaw(0));
      for (int t = blockIdx.x; t < ntiles; t += tstep) {
        const int row0 = (t % n_rb) * G6_BM;
        for (int ky = 0; ky < 3; ++ky)
          for (int cb = 0; cb < cg.cpb; ++cb, ++gs) {
            const uint32_t stage = gs & 1u;
            if (gs >= 2) mbar_wait(&sm.slab_free[stage], ((gs >> 1) - 1) & 1);
            const uint32_t bar = smem_u32(&sm.slab_full[stage]);
            const int c0 = cb * G6_BK, c1 = row0 + (ky - 1) * cg.W - 1;      // (channel, flattened pixel) of the box origin
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(G6_SLAB_ROWS * 128) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst0 + stage * G6_SLAB_BYTES), "l"(reinterpret_cast<uint64_t>(&amap)), "r"(c0), "r"(c1), "r"(bar) : "memory");
          }
      }
    } else if (lane == 0) {
      uint32_t unit = 0;
      for (int t = blockIdx.x; t < ntiles; t += tstep) {
        const int nchunk = t / n_rb;
        // only the rows of the tile image this chunk multiplies (N - n_base < 128 for a last / only chunk, e.g. the 64-channel
        // convolutions): a tile image is row-major, 128 B per output column, swizzled inside 8-row atoms -> a prefix of whole
        // atoms is a valid smaller tile.  These kernels run at the L2 -> SM ceiling; the weight stream is 2/3 of their bytes.
        const uint32_t tbytes = (uint32_t)((min(G6_BN, a.N - nchunk * G6_BN) + 7) & ~7) * 128u;
        for (int kb = 0; kb < nkb; ++kb, ++unit) {
          const int slot = unit % G6_NB;
          if (unit >= G6_NB) mbar_wait(&sm.done[slot], ((unit - G6_NB) / G6_NB) & 1);   // MMAs that read this slot are complete
          int wkb = kb;
          if (CONV && cg.slab) {                     // unit order (ky, channel block, kx) -> packed order (ky, kx, channel block)
            const int sl = kb / 3, kx = kb - sl * 3, ky = sl / cg.cpb, cb = sl - ky * cg.cpb;
            wkb = (ky * 3 + kx) * cg.cpb + cb;
          }
          const size_t toff = ((size_t)nchunk * nkb + wkb) * (G6_TILE / 4);
          const uint32_t bar = smem_u32(&sm.full_b[slot]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * tbytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(sB_hi(slot))), "l"(a.Wt_hi + toff), "r"(tbytes), "r"(bar) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(sB_lo(slot))), "l"(a.Wt_lo + toff), "r"(tbytes), "r"(bar) : "memory");
        }
      }
    }
  } else {
    // =============================================== epilogue ===============================================
    const int e = warp - G6_EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = e >> 2;                   // two warps per quarter: even / odd 32-column chunks
    float* stage = stage_base + e * G6_STAGE_FLOATS;
    const int srow = lane >> 3, scol = (lane & 7) * 4;
    // LayerNorm statistics of local tile j -> buffer j % 4 (16 rows per warp), three tiles ahead of the drain
    auto stats = [&](int j) {
      const int tj = blockIdx.x + j * tstep;
      if (tj < ntiles) {
        tile_stats6(a.X, a.ldx, a.rows, (tj % n_rb) * G6_BM, e, sm.mean[j % G6_STATS], sm.rstd[j % G6_STATS]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.stats_full[j % G6_STATS]);
      }
    };
    const bool own_stats = LN && a.ln_stats == nullptr;
    if (own_stats) { stats(0); stats(1); stats(2); }
    int it = 0;
    uint32_t grp = 0;
    for (int t = blockIdx.x; t < ntiles; t += tstep, ++it) {
      const int row0 = (t % n_rb) * G6_BM, n_base = (t / n_rb) * G6_BN;
#ifdef NMRF_STATS_FIRST
      if (own_stats) stats(it + 3);
#endif
      trace(tp, 3584 + it * 4 + 0);
      const int nchunks = (min(G6_BN, a.N - n_base) + 31) / 32;
      // the tile's sum over its groups of k-blocks, in registers: this thread's row, column chunks half and half + 2 (32 columns each)
      float acc[2][32];
      for (int kb = 0; kb < ngrp; ++kb, ++grp) {
        const int as = grp % G6_ACC;
        mbar_wait_warp(&sm.acc_full[as], (grp / G6_ACC) & 1, kb == 0 ? 64 : 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int ch = half + 2 * ci;
          if (ch < nchunks) {
#pragma unroll
            for (int h16 = 0; h16 < 2; ++h16) {      // 16 columns at a time: the 64 accumulator registers stay live
              float v[16];
              tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * G6_BN + ch * 32 + h16 * 16), v);
              if (kb == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[ci][h16 * 16 + j] = v[j];
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[ci][h16 * 16 + j] += v[j];
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&sm.acc_empty[as]);
      }
      trace(tp, 3584 + it * 4 + 1);
      // LayerNorm statistics three tiles ahead: AFTER this tile's accumulator stages have been handed back (the MMA warp is never
      // more than three groups ahead of the drains).  Buffer (it+3) % 4 last held tile it-1, whose statistics the producers
      // read before its MMAs, which this warp drained an iteration ago.
#ifndef NMRF_STATS_FIRST
      if (own_stats) stats(it + 3);
#endif
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int ch = half + 2 * ci;
        if (ch >= nchunks) break;
        const float* v = acc[ci];
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int n = n_base + ch * 32 + scol;
        if (n < a.N) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias) b = *reinterpret_cast<const float4*>(a.bias + n);
#pragma unroll 1
          for (int g4 = 0; g4 < 2; ++g4) {      // four rows per lane at a time: the other chunk's 32 accumulator registers are live
            float4 rr[4];                       // residual first (R may alias Y: read-before-write by the same thread)
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const int r = row0 + q * 32 + (g4 * 4 + i4) * 4 + srow;
              rr[i4] = (a.R && r < a.rows) ? __ldcg(reinterpret_cast<const float4*>(a.R + (size_t)r * a.ldr + n))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const int lr = (g4 * 4 + i4) * 4 + srow;
              const int r = row0 + q * 32 + lr;
              if (r < a.rows) {
                float4 o = *reinterpret_cast<const float4*>(stage + lr * 36 + scol);
                o.x = act_fast(o.x + b.x, ACT) + rr[i4].x; o.y = act_fast(o.y + b.y, ACT) + rr[i4].y;
                o.z = act_fast(o.z + b.z, ACT) + rr[i4].z; o.w = act_fast(o.w + b.w, ACT) + rr[i4].w;
                *reinterpret_cast<float4*>(a.Y + (size_t)r * a.ldy + n) = o;
              }
            }
          }
        }
        __syncwarp();
      }
      trace(tp, 3584 + it * 4 + 2);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

}  // namespace

namespace {
// one thread per destination float of the tile images
__global__ void pack_weight_tiles_kernel(const float* __restrict__ w, int N, int K, int nkb, long long total,
                                         float* __restrict__ hi, float* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int within = (int)(i % 4096);               // float index inside the 16 KB tile image
  const long long tile = i / 4096;
  const int kb = (int)(tile % nkb), nc = (int)(tile / nkb);
  const int r = within >> 5, pchunk = (within & 31) >> 2, e = within & 3;
  const int c = pchunk ^ (r & 7);                   // logical 16-byte chunk stored at this physical position
  const int n = nc * 128 + r, k = kb * 32 + c * 4 + e;
  const float x = (n < N && k < K) ? w[(size_t)n * K + k] : 0.f;
  const float h = tc::rna_tf32(x);
  hi[i] = h;
  lo[i] = tc::rna_tf32(x - h);
}
}  // namespace

int pack_weight_tiles(const float* w, int N, int K, float* hi, float* lo, cudaStream_t stream) {
  NMRF_REQUIRE(w && hi && lo && N > 0 && K > 0, "pack_weight_tiles: bad arguments");
  const int nkb = (K + 31) / 32, nnc = (N + 127) / 128;
  const long long total = (long long)nkb * nnc * 4096;
  pack_weight_tiles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(w, N, K, nkb, total, hi, lo);
  count_launch();
  return check_launch("pack_weight_tiles");
}

#ifdef NMRF_TRACE
int gemm6_set_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_trace6, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? NMRF_OK : NMRF_ERR_CUDA;
}
#endif

namespace {
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = w[i], h = tc::rna_tf32(x);
  hi[i] = h;
  lo[i] = tc::rna_tf32(x - h);
}
}  // namespace

int split_tf32(const float* w, float* hi, float* lo, long long n, cudaStream_t stream) {
  NMRF_REQUIRE(w && hi && lo && n >= 0, "split_tf32: bad arguments");
  if (n == 0) return NMRF_OK;
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, hi, lo, n);
  count_launch();
  return check_launch("split_tf32");
}

namespace {
template <int ACT, bool LN, bool CONV = false>
void launch6(const nmrf_gemm_args& a, int n_rb, int n_nc, int grid, cudaStream_t stream, const ConvGeom& cg = ConvGeom(),
             const CUtensorMap& amap = CUtensorMap()) {
  static PerDevice configured;        // per instantiation and device
  ensure_dynamic_smem(token_gemm_tc6_kernel<ACT, LN, CONV>, G6_DYN, configured);
  token_gemm_tc6_kernel<ACT, LN, CONV><<<grid, G6_BLOCK, G6_DYN, stream>>>(a, n_rb, n_nc, cg, amap);
}

}  // namespace

int token_gemm_tc6(const nmrf_gemm_args& a, cudaStream_t stream) {
  const int num_sms = nmrf::num_sms();
  const int n_rb = (a.rows + G6_BM - 1) / G6_BM;
  const int n_nc = (a.N + G6_BN - 1) / G6_BN;
  const int ntiles = n_rb * n_nc;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  const bool ln = a.ln_gamma != nullptr;
  // activation and LayerNorm are compile-time: each instantiation carries only its own epilogue / statistics code
  switch (a.act * 2 + (ln ? 1 : 0)) {
    case 0: launch6<0, false>(a, n_rb, n_nc, grid, stream); break;
    case 1: launch6<0, true>(a, n_rb, n_nc, grid, stream); break;
    case 2: launch6<1, false>(a, n_rb, n_nc, grid, stream); break;
    case 3: launch6<1, true>(a, n_rb, n_nc, grid, stream); break;
    case 4: launch6<2, false>(a, n_rb, n_nc, grid, stream); break;
    case 5: launch6<2, true>(a, n_rb, n_nc, grid, stream); break;
    default: set_error("token_gemm: unknown activation %d", a.act); return NMRF_ERR_BAD_ARG;
  }
  count_launch();
  return check_launch("token_gemm_tc6");
}

namespace {
// (mean, rstd) of every 128-wide row: two-pass like torch's LayerNorm; 8 lanes per row, four rows per warp and step
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ X, int ldx, int rows, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = warp * 4; r0 < rows; r0 += nwarps * 4) {
    const int r = r0 + g;
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v[j] = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + j * 32 + sub * 4);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float mu = s * (1.f / 128.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dx = v[j].x - mu, dy = v[j].y - mu, dz = v[j].z - mu, dw = v[j].w - mu;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
    if (sub == 0 && r < rows) stats[r] = make_float2(mu, 1.f / sqrtf(q * (1.f / 128.f) + 1e-5f));
  }
}
}  // namespace

int row_stats(const float* X, int ldx, int rows, float* stats, cudaStream_t stream) {
  NMRF_REQUIRE(X && stats && rows >= 0 && ldx >= 128 && ldx % 4 == 0, "row_stats: bad arguments");
  if (rows == 0) return NMRF_OK;
  const int warps = (rows + 3) / 4;
  const int blocks = min((warps + 7) / 8, 8 * nmrf::num_sms());
  row_stats_kernel<<<blocks, 256, 0, stream>>>(X, ldx, rows, reinterpret_cast<float2*>(stats));
  count_launch();
  return check_launch("row_stats");
}

// N1 / N2: k x k convolution over NHWC as an implicit GEMM on the kernel above (CONV mode)
int conv2d_tc6(const nmrf_conv_args& c, cudaStream_t stream) {
  NMRF_REQUIRE(c.X && c.Wt_hi && c.Wt_lo && c.Y, "conv2d: null pointer");
  NMRF_REQUIRE(c.N > 0 && c.H > 0 && c.W > 0 && c.kh > 0 && c.kw > 0 && c.stride > 0 && c.pad >= 0, "conv2d: bad geometry");
  NMRF_REQUIRE(c.Cin > 0 && c.Cin % 32 == 0, "conv2d: Cin=%d must be a multiple of 32 (floats consumed per tap)", c.Cin);
  NMRF_REQUIRE(c.Cout > 0 && c.Cout % 16 == 0, "conv2d: Cout=%d must be a multiple of 16", c.Cout);
  NMRF_REQUIRE(c.pix_stride % 4 == 0 && c.row_stride % 4 == 0 && c.img_stride % 4 == 0, "conv2d: strides must keep 16-byte alignment");
  NMRF_REQUIRE(c.H < 32768 && c.W < 32768, "conv2d: image extent above 32767");
  const int Ho = c.Ho > 0 ? c.Ho : (c.H + 2 * c.pad - c.kh) / c.stride + 1, Wo = c.Wo > 0 ? c.Wo : (c.W + 2 * c.pad - c.kw) / c.stride + 1;
  NMRF_REQUIRE(Ho > 0 && Wo > 0 && (long long)c.N * Ho * Wo < (1ll << 31), "conv2d: bad output extent %dx%d", Ho, Wo);
  nmrf_gemm_args a = {};
  a.X = c.X; a.Kx = c.kh * c.kw * c.Cin; a.ldx = 0;
  a.ediv = 1;
  a.bias = c.bias;
  a.Y = c.Y; a.ldy = c.Cout;
  a.rows = c.N * Ho * Wo; a.N = c.Cout;
  a.Wt_hi = c.Wt_hi; a.Wt_lo = c.Wt_lo;
  ConvGeom g;
  g.H = c.H; g.W = c.W; g.pix_stride = c.pix_stride; g.row_stride = c.row_stride; g.img_stride = c.img_stride;
  g.Ho = Ho; g.Wo = Wo; g.kw = c.kw; g.stride = c.stride; g.pad = c.pad; g.cpb = c.Cin / 32;
  g.slab = c.kh == 3 && c.kw == 3 && c.stride == 1 && c.pad == 1 && Ho == c.H && Wo == c.W &&
           c.row_stride == (long long)c.W * c.pix_stride && c.img_stride == (long long)c.H * c.W * c.pix_stride;
  static const bool slab_off = [] { const char* e = getenv("NMRF_B200_CONV_SLAB"); return e && e[0] == '0'; }();   // A/B switch
  if (slab_off) g.slab = 0;
  CUtensorMap amap = {};
  // the NHWC activation seen as [pixels, channels] (pixel stride pix_stride floats); box = one slab.  No driver entry point:
  // tap-by-tap gather
  if (g.slab && !encode_tmap_2d(&amap, c.X, (long long)c.N * c.H * c.W, c.Cin, c.pix_stride, G6_SLAB_ROWS, G6_BK)) g.slab = 0;
  const int num_sms = nmrf::num_sms();
  const int n_rb = (a.rows + G6_BM - 1) / G6_BM, n_nc = (a.N + G6_BN - 1) / G6_BN;
  const int ntiles = n_rb * n_nc;
  launch6<0, false, true>(a, n_rb, n_nc, ntiles < num_sms ? ntiles : num_sms, stream, g, amap);
  count_launch();
  return check_launch("conv2d");
}

}  // namespace nmrf
