// Fused token GEMM on tcgen05, warp-specialised, A operand in TENSOR MEMORY (v6).
//
// Why: a cycle trace of v5 (tools/gemm_trace.py, profiles/) showed the MMA warp needing ~2100 cycles to issue the 12
// UMMAs of a unit whose tensor work is 768 cycles: every SS-mode tf32 UMMA (128x128x8) reads 4 KB of A and 4 KB of B from
// shared memory, so the MMAs alone saturate the shared-memory pipe while the producers' st.shared / cp.async traffic
// competes with them.  Here the producers write the LayerNorm'ed, concatenated, hi/lo-split A operand straight into
// TMEM (tcgen05.st) and the UMMA takes A from TMEM: shared memory only carries the weight tiles (B).
//
//   warps 0-7   producers   thread = (row, half of the 32-wide k-block): warp w owns TMEM lane quarter w%4 and k-columns
//                           16*(w/4)..+16; raw A values prefetched three k-blocks ahead, LayerNorm statistics of the next
//                           tile computed early; A_hi / A_lo -> TMEM (tcgen05.st 32x32b.x16), weight tiles by cp.async
//   warp  8     MMA issuer  per unit: 4 k-steps x {A_lo.B_hi, A_hi.B_lo, A_hi.B_hi}, A from TMEM, B from swizzled smem
//   warps 9-16  epilogue    as v5 (TMEM -> registers -> smem transpose -> coalesced bias/activation/residual/store)
//
//   tile        128 rows x 128 output columns, unit = one k-block of 32; three accumulator stages (3 x 128 TMEM columns)
//               so the epilogue of a tile overlaps the MMAs of the next two; A tiles double buffered (2 x 64 columns).
//               TMEM map: [0,384) accumulators, [384,512) A: buffer b at 384 + 64 b, hi at +0, lo at +32.
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int G6_BM = 128, G6_BN = 128, G6_BK = 32, G6_NPASS = 128, G6_NB = 3, G6_ACC = 3;
constexpr int G6_ACOL = G6_ACC * G6_BN;             // first TMEM column of the A buffers
constexpr int G6_TILE = G6_BM * G6_BK * 4;          // 16 KB operand tile
constexpr int G6_PROD = 256;                        // producer threads (warps 0-7)
constexpr int G6_MMA_WARP = 8;
constexpr int G6_EPI_WARP0 = 9, G6_EPI_WARPS = 8;
constexpr int G6_BLOCK = (G6_EPI_WARP0 + G6_EPI_WARPS) * 32;   // 544
constexpr int G6_STATS_BAR = 4;                     // named barrier of the producers (LayerNorm statistics hand-over)
constexpr int G6_HANDOFF = G6_PROD + 32;            // named-barrier population: producers arrive, MMA warp syncs
constexpr int G6_STAGE_FLOATS = 32 * 36;            // per-epilogue-warp transpose tile
constexpr int G6_DYN = 6 * G6_TILE + G6_EPI_WARPS * G6_STAGE_FLOATS * 4 + 1024;   // B_hi[3] B_lo[3] + epilogue staging

struct G6Smem {
  uint64_t done[G6_NB];       // MMAs of the unit that used B slot s are complete (tcgen05.commit)
  uint64_t full_b[G6_NB];     // the slot's two weight tiles have landed (TMA bulk copy, expect_tx 32 KB)
  uint64_t acc_full[G6_ACC];  // accumulator stage holds a finished tile (tcgen05.commit)
  uint64_t acc_empty[G6_ACC]; // epilogue has drained the stage (256 arrivals)
  uint32_t tmem_base;
  float mean[2][G6_BM], rstd[2][G6_BM];   // LayerNorm statistics, double buffered across tiles
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TileCoord { int row0, n_base, npass, nnc; };

// optional cycle trace of CTA 0 (debug tooling: nmrf_debug_set_trace); slot layout documented in tools/gemm_trace.py
__device__ long long* g_trace6 = nullptr;
__device__ __forceinline__ void trace(long long* tp, int idx) {
  if (tp && idx < 4096) tp[idx] = clock64();
}

__global__ void __launch_bounds__(G6_BLOCK, 1)
token_gemm_tc6_kernel(const nmrf_gemm_args a, const float* __restrict__ W_lo, int n_rb, int n_np) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ G6Smem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  auto sB_hi = [&](int i) { return base + i * G6_TILE; };
  auto sB_lo = [&](int i) { return base + (3 + i) * G6_TILE; };
  float* stage_base = reinterpret_cast<float*>(base + 6 * G6_TILE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* const tp = (blockIdx.x == 0 && (tid == 0 || tid == G6_MMA_WARP * 32 || tid == G6_EPI_WARP0 * 32)) ? g_trace6 : nullptr;
  const int Ktot = a.Kx + a.Ke;
  const int nkb = (Ktot + G6_BK - 1) / G6_BK;
  const int nkb_w = nkb;                 // k-blocks per row of tiles in the weight images (K padded to 32)
  const int ntiles = n_rb * n_np;
  const bool ln = a.ln_gamma != nullptr;
  auto coord = [&](int t) {
    TileCoord c;
    c.row0 = (t % n_rb) * G6_BM;           // pass-major order: with a persistent stride of gridDim.x every CTA gets the same
    c.n_base = (t / n_rb) * G6_NPASS;      // mix of wide (256-column) and narrow passes
    c.npass = min(G6_NPASS, a.N - c.n_base);
    c.nnc = (c.npass + G6_BN - 1) / G6_BN;
    return c;
  };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < G6_NB; ++i) { mbar_init(&sm.done[i], 1); mbar_init(&sm.full_b[i], 1); }
    for (int i = 0; i < G6_ACC; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], G6_EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp < 8) {
    // =============================================== producers ===============================================
    // warp w may touch TMEM lanes 32*(w%4)..+32 only: thread -> row 32*(w%4)+lane, k-columns 16*(w/4)..+16 of the k-block
    const int a_row = (warp & 3) * 32 + lane, a_c0 = (warp >> 2) * 4;
    const uint32_t a_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    // weight tiles of local unit `ut` -> slot: ONE thread issues two 16 KB TMA bulk copies of the pre-swizzled tile images
    auto load_B = [&](const TileCoord& tc_, int ut, int slot) {
      if (tid == 0) {
        const int kb = ut / tc_.nnc, nchunk = tc_.n_base / G6_BN + (ut % tc_.nnc);
        const size_t toff = ((size_t)nchunk * nkb_w + kb) * (G6_TILE / 4);
        const uint32_t bar = smem_u32(&sm.full_b[slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * G6_TILE) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sB_hi(slot))), "l"(a.Wt_hi + toff), "r"(G6_TILE), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sB_lo(slot))), "l"(a.Wt_lo + toff), "r"(G6_TILE), "r"(bar) : "memory");
      }
    };
    // LayerNorm statistics of a tile's 128 rows into buffer `par` (Kx == 128): one warp per 16 rows, coalesced
    auto tile_stats = [&](int row0, int par) {
      for (int i = 0; i < G6_BM / 8; ++i) {
        const int lr = warp * (G6_BM / 8) + i, r = row0 + lr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < a.rows) v = *reinterpret_cast<const float4*>(a.X + (size_t)r * a.ldx + lane * 4);
        const float mu = warp_sum(v.x + v.y + v.z + v.w) * (1.f / 128.f);
        const float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
        const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f);
        if (lane == 0) { sm.mean[par][lr] = mu; sm.rstd[par][lr] = 1.f / sqrtf(var + 1e-5f); }
      }
    };
    uint32_t unit = 0;
    uint32_t akb = 0;          // k-blocks produced so far by this CTA: A buffer = akb & 1 (alternates ACROSS tiles too, so the
                               // buffer being rewritten was last read two k-blocks -- at least two units -- ago)
    int par = 0;
    if ((int)blockIdx.x < ntiles) {
      load_B(coord(blockIdx.x), 0, 0);
      if (ln) tile_stats(coord(blockIdx.x).row0, 0);
    }
    asm volatile("bar.sync %0, %1;" ::"r"(G6_STATS_BAR), "r"(G6_PROD) : "memory");
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const TileCoord tc_ = coord(t);
      const bool has_next = t + (int)gridDim.x < ntiles;
      const int g_row = tc_.row0 + a_row;
      const bool row_ok = g_row < a.rows;
      const float* xrow = a.X + (size_t)(row_ok ? g_row : 0) * a.ldx;
      const float* erow = a.E ? a.E + (size_t)((row_ok ? g_row : 0) / a.ediv) * a.lde : nullptr;
      const float mean = ln ? sm.mean[par][a_row] : 0.f, rstd = ln ? sm.rstd[par][a_row] : 1.f;
      float4 ar0[4], ar1[4], ar2[4];          // raw A values of three k-blocks in flight (round-robin, no register moves)
      auto fetch_A = [&](int kb, float4 (&dst)[4]) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int kk = kb * G6_BK + (a_c0 + cc) * 4;
          const float* p = (kk < a.Kx) ? xrow + kk : erow + (kk - a.Kx);
          dst[cc] = (row_ok && kk < Ktot) ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      fetch_A(0, ar0);
      fetch_A(1, ar1);
      fetch_A(2, ar2);
      const int upt = nkb * tc_.nnc;
      const int stats_at = upt > 3 ? upt - 3 : 0;            // next tile's statistics go out while the last units are produced
      for (int ut = 0; ut < upt; ++ut, ++unit) {
        const int kb = ut / tc_.nnc, nc = ut - kb * tc_.nnc;
        const int slot = unit % G6_NB;
        // MMAs of unit-2 (and, cumulatively, all earlier ones) are complete: frees B slot (unit+1)%3 and the A buffer of
        // k-block akb-2.  unit-2 is the newest unit whose barrier phase is unambiguous (its slot is next used by unit+1).
        trace(tp, unit * 8 + 0);
        if (unit >= 2) mbar_wait(&sm.done[(unit - 2) % G6_NB], ((unit - 2) / G6_NB) & 1);
        trace(tp, unit * 8 + 1);
        const bool prefetch = (ut + 1 < upt) || has_next;
        if (prefetch) {
          if (ut + 1 < upt) load_B(tc_, ut + 1, (unit + 1) % G6_NB);
          else load_B(coord(t + gridDim.x), 0, (unit + 1) % G6_NB);
        }
        trace(tp, unit * 8 + 2);
        if (ut == stats_at && has_next && ln) tile_stats(coord(t + gridDim.x).row0, par ^ 1);
        if (nc == 0) {
          auto produce = [&](float4 (&buf)[4]) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int kk = kb * G6_BK + (a_c0 + cc) * 4;
              float4 v = buf[cc];
              if (ln && row_ok && kk < a.Kx) {
                const float4 g = *reinterpret_cast<const float4*>(a.ln_gamma + kk);
                const float4 b = *reinterpret_cast<const float4*>(a.ln_beta + kk);
                v.x = (v.x - mean) * rstd * g.x + b.x; v.y = (v.y - mean) * rstd * g.y + b.y;
                v.z = (v.z - mean) * rstd * g.z + b.z; v.w = (v.w - mean) * rstd * g.w + b.w;
              }
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float h = rna_tf32_fast(vv[j]);
                hi[cc * 4 + j] = __float_as_uint(h);
                lo[cc * 4 + j] = __float_as_uint(rna_tf32_fast(vv[j] - h));
              }
            }
            const uint32_t ta = tmem + a_lane + (uint32_t)(G6_ACOL + (akb & 1) * 64 + a_c0 * 4);
            tmem_st16(ta, hi);
            tmem_st16(ta + 32, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fetch_A(kb + 3, buf);
          };
          const int which = kb % 3;
          if (which == 0) produce(ar0); else if (which == 1) produce(ar1); else produce(ar2);
          ++akb;
        }
        trace(tp, unit * 8 + 3);
        trace(tp, unit * 8 + 4);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // A written with tcgen05.st -> ordered before the hand-off
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + slot), "r"(G6_HANDOFF) : "memory");
        trace(tp, unit * 8 + 5);
      }
      // the next tile's statistics (written by other warps) become visible to every producer
      asm volatile("bar.sync %0, %1;" ::"r"(G6_STATS_BAR), "r"(G6_PROD) : "memory");
      par ^= 1;
    }
  } else if (warp == G6_MMA_WARP) {
    // =============================================== MMA issuer ===============================================
    uint32_t unit = 0;
    uint32_t akb = 0, abuf = 0;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const TileCoord tc_ = coord(t);
      const int as = it % G6_ACC;
      if (it >= G6_ACC) mbar_wait(&sm.acc_empty[as], ((it / G6_ACC) - 1) & 1);   // epilogue of tile it-3 has drained the stage
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int upt = nkb * tc_.nnc;
      for (int ut = 0; ut < upt; ++ut, ++unit) {
        const int kb = ut / tc_.nnc, nc = ut - kb * tc_.nnc;
        const int slot = unit % G6_NB;
        if (nc == 0) abuf = (akb++) & 1;
        trace(tp, 2048 + unit * 4 + 0);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(G6_HANDOFF) : "memory");      // A of this unit is in TMEM
        mbar_wait(&sm.full_b[slot], (unit / G6_NB) & 1);                                  // B tiles have landed
        trace(tp, 2048 + unit * 4 + 1);
        if (lane == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int bn = min(G6_BN, tc_.npass - nc * G6_BN);
          const uint32_t idesc = make_idesc(bn);
          const uint64_t dBh = make_desc(smem_u32(sB_hi(slot))), dBl = make_desc(smem_u32(sB_lo(slot)));
          const uint32_t d = tmem + (uint32_t)(as * G6_BN);
          const uint32_t tAh = tmem + (uint32_t)(G6_ACOL + abuf * 64), tAl = tAh + 32;
#pragma unroll
          for (int ks = 0; ks < G6_BK / 8; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);               // B: +32 bytes inside the swizzle row; A: +8 TMEM columns
            umma_tf32_ta(d, tAl + ks * 8, dBh + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            umma_tf32_ta(d, tAh + ks * 8, dBl + adv, idesc, 1u);
            umma_tf32_ta(d, tAh + ks * 8, dBh + adv, idesc, 1u);
          }
          umma_commit(&sm.done[slot]);
          if (ut == upt - 1) umma_commit(&sm.acc_full[as]);
        }
        __syncwarp();
        trace(tp, 2048 + unit * 4 + 2);
      }
    }
  } else {
    // =============================================== epilogue ===============================================
    const int e = warp - G6_EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = e >> 2;                   // two warps per quarter: even / odd 32-column chunks
    float* stage = stage_base + e * G6_STAGE_FLOATS;
    const int srow = lane >> 3, scol = (lane & 7) * 4;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const TileCoord tc_ = coord(t);
      const int as = it % G6_ACC;
      trace(tp, 3584 + it * 4 + 0);
      mbar_wait(&sm.acc_full[as], (it / G6_ACC) & 1);
      trace(tp, 3584 + it * 4 + 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nchunks = (tc_.npass + 31) / 32;
      for (int ch = half; ch < nchunks; ch += 2) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * G6_BN + ch * 32), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int n = tc_.n_base + ch * 32 + scol;
        if (n < a.N) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias) b = *reinterpret_cast<const float4*>(a.bias + n);
          float4 rr[8];                         // residual first (R may alias Y: read-before-write by the same thread)
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = tc_.row0 + q * 32 + i8 * 4 + srow;
            rr[i8] = (a.R && r < a.rows) ? __ldcg(reinterpret_cast<const float4*>(a.R + (size_t)r * a.ldr + n))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int lr = i8 * 4 + srow;
            const int r = tc_.row0 + q * 32 + lr;
            if (r < a.rows) {
              float4 o = *reinterpret_cast<const float4*>(stage + lr * 36 + scol);
              o.x = act_fast(o.x + b.x, a.act) + rr[i8].x; o.y = act_fast(o.y + b.y, a.act) + rr[i8].y;
              o.z = act_fast(o.z + b.z, a.act) + rr[i8].z; o.w = act_fast(o.w + b.w, a.act) + rr[i8].w;
              *reinterpret_cast<float4*>(a.Y + (size_t)r * a.ldy + n) = o;
            }
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&sm.acc_empty[as]);
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

int gemm6_set_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_trace6, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? NMRF_OK : NMRF_ERR_CUDA;
}

int token_gemm_tc6(const nmrf_gemm_args& a, const float* W_lo, cudaStream_t stream) {
  static int num_sms = 0;
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(token_gemm_tc6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G6_DYN);
    configured = true;
  }
  const int n_rb = (a.rows + G6_BM - 1) / G6_BM;
  const int n_np = (a.N + G6_NPASS - 1) / G6_NPASS;
  const int ntiles = n_rb * n_np;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  token_gemm_tc6_kernel<<<grid, G6_BLOCK, G6_DYN, stream>>>(a, W_lo, n_rb, n_np);
  count_launch();
  return check_launch("token_gemm_tc6");
}

}  // namespace nmrf
