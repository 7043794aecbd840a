// Fused token GEMM on tcgen05, warp-specialised (v5).  Same arithmetic as gemm_tc.cu (error-compensated 3xTF32,
// LayerNorm / concat prologue, bias / ReLU / GELU / residual epilogue); different schedule:
//
//   warps 0-7   producers   two threads per token row: LayerNorm statistics (cooperative, coalesced, computed for the NEXT
//                           tile while this tile's last units are produced), raw A values prefetched three k-blocks
//                           ahead in registers, hi/lo split into the swizzled A tiles (st.shared), weight tiles by
//                           cp.async one unit ahead; hand-off to the MMA warp through a per-slot named barrier
//                           (bar.arrive), never waits for the issue
//   warp  8     MMA issuer  bar.sync on the slot, 12 tcgen05.mma (4 k-steps x {lo.hi, hi.lo, hi.hi}) per unit,
//                           tcgen05.commit -> slot mbarrier; after the tile's last unit commit -> acc_full[stage]
//   warps 9-16  epilogue    wait acc_full[stage]; tcgen05.ld 32x32b per warp (its TMEM lane quarter), transpose through a
//                           private smem tile so global traffic is coalesced (8 lanes = one row's 128 B), bias, activation,
//                           residual, store; mbarrier.arrive acc_empty[stage]
//
//   tile        128 token rows x (up to) 256 output columns; the fp32 accumulators of TWO tiles live in TMEM
//               (2 x 256 of the 512 columns), so the epilogue of tile i overlaps the MMAs of tile i+1.
//               N > 256 (qkv: 384, fc1: 512) is covered by several column passes over the same rows.
//   units       (k-block of 32, n-chunk of 128): A tiles double buffered per k-block, B tiles in a 3-slot ring.
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int G5_BM = 128, G5_BN = 128, G5_BK = 32, G5_NPASS = 256, G5_NB = 3;
constexpr int G5_TILE = G5_BM * G5_BK * 4;          // 16 KB operand tile
constexpr int G5_PROD = 256;                        // producer threads (warps 0-7)
constexpr int G5_MMA_WARP = 8;
constexpr int G5_EPI_WARP0 = 9, G5_EPI_WARPS = 8;
constexpr int G5_BLOCK = (G5_EPI_WARP0 + G5_EPI_WARPS) * 32;   // 544
constexpr int G5_STATS_BAR = 4;                     // named barrier of the producers (LayerNorm statistics hand-over)
constexpr int G5_HANDOFF = G5_PROD + 32;            // named-barrier population: producers arrive, MMA warp syncs
constexpr int G5_STAGE_FLOATS = 32 * 36;            // per-epilogue-warp transpose tile
constexpr int G5_DYN = 10 * G5_TILE + G5_EPI_WARPS * G5_STAGE_FLOATS * 4 + 1024;

struct G5Smem {
  uint64_t done[G5_NB];       // MMAs of the unit that used B slot s are complete (tcgen05.commit)
  uint64_t acc_full[2];       // accumulator stage holds a finished tile (tcgen05.commit)
  uint64_t acc_empty[2];      // epilogue has drained the stage (256 arrivals)
  uint32_t tmem_base;
  float mean[2][G5_BM], rstd[2][G5_BM];   // LayerNorm statistics, double buffered across tiles
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TileCoord { int row0, n_base, npass, nnc; };

// optional cycle trace of CTA 0 (debug tooling: nmrf_debug_set_trace); slot layout documented in tools/gemm_trace.py
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ void trace(long long* tp, int idx) {
  if (tp && idx < 4096) tp[idx] = clock64();
}

__global__ void __launch_bounds__(G5_BLOCK, 1)
token_gemm_tc5_kernel(const nmrf_gemm_args a, const float* __restrict__ W_lo, int n_rb, int n_np) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ G5Smem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  auto sA_hi = [&](int i) { return base + i * G5_TILE; };
  auto sA_lo = [&](int i) { return base + (2 + i) * G5_TILE; };
  auto sB_hi = [&](int i) { return base + (4 + i) * G5_TILE; };
  auto sB_lo = [&](int i) { return base + (7 + i) * G5_TILE; };
  float* stage_base = reinterpret_cast<float*>(base + 10 * G5_TILE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* const tp = (blockIdx.x == 0 && (tid == 0 || tid == G5_MMA_WARP * 32 || tid == G5_EPI_WARP0 * 32)) ? g_trace : nullptr;
  const int Ktot = a.Kx + a.Ke;
  const int nkb = (Ktot + G5_BK - 1) / G5_BK;
  const int ntiles = n_rb * n_np;
  const bool ln = a.ln_gamma != nullptr;
  auto coord = [&](int t) {
    TileCoord c;
    c.row0 = (t % n_rb) * G5_BM;           // pass-major order: with a persistent stride of gridDim.x every CTA gets the same
    c.n_base = (t / n_rb) * G5_NPASS;      // mix of wide (256-column) and narrow passes
    c.npass = min(G5_NPASS, a.N - c.n_base);
    c.nnc = (c.npass + G5_BN - 1) / G5_BN;
    return c;
  };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < G5_NB; ++i) mbar_init(&sm.done[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], G5_EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp < 8) {
    // =============================================== producers ===============================================
    const int a_row = tid >> 1, a_c0 = (tid & 1) * 4;       // thread -> (row, 16 consecutive k = 4 chunks of 16 B)
    auto load_B = [&](const TileCoord& tc_, int ut, int slot) {
      const int kb = ut / tc_.nnc, n0 = tc_.n_base + (ut % tc_.nnc) * G5_BN;
      const int bn = min(G5_BN, a.N - n0);
      for (int i = tid; i < bn * 8; i += G5_PROD) {
        const int r = i >> 3, c = i & 7;
        const size_t goff = (size_t)(n0 + r) * a.ldw + kb * G5_BK + c * 4;
        const uint32_t so = swz(r, c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB_hi(slot) + so)), "l"(a.W + goff));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB_lo(slot) + so)), "l"(W_lo + goff));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // LayerNorm statistics of a tile's 128 rows into buffer `par` (Kx == 128): one warp per 16 rows, coalesced
    auto tile_stats = [&](int row0, int par) {
      for (int i = 0; i < G5_BM / 8; ++i) {
        const int lr = warp * (G5_BM / 8) + i, r = row0 + lr;
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
    asm volatile("bar.sync %0, %1;" ::"r"(G5_STATS_BAR), "r"(G5_PROD) : "memory");
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
          const int kk = kb * G5_BK + (a_c0 + cc) * 4;
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
        const int slot = unit % G5_NB;
        // MMAs of unit-2 (and, cumulatively, all earlier ones) are complete: frees B slot (unit+1)%3 and the A buffer of
        // k-block akb-2.  unit-2 is the newest unit whose barrier phase is unambiguous (its slot is next used by unit+1).
        trace(tp, unit * 8 + 0);
        if (unit >= 2) mbar_wait(&sm.done[(unit - 2) % G5_NB], ((unit - 2) / G5_NB) & 1);
        trace(tp, unit * 8 + 1);
        const bool prefetch = (ut + 1 < upt) || has_next;
        if (prefetch) {
          if (ut + 1 < upt) load_B(tc_, ut + 1, (unit + 1) % G5_NB);
          else load_B(coord(t + gridDim.x), 0, (unit + 1) % G5_NB);
        }
        trace(tp, unit * 8 + 2);
        if (ut == stats_at && has_next && ln) tile_stats(coord(t + gridDim.x).row0, par ^ 1);
        if (nc == 0) {
          auto produce = [&](float4 (&buf)[4]) {
            uint8_t* dh = sA_hi(akb & 1);
            uint8_t* dl = sA_lo(akb & 1);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int c = a_c0 + cc;
              const int kk = kb * G5_BK + c * 4;
              float4 v = buf[cc];
              if (ln && row_ok && kk < a.Kx) {
                const float4 g = *reinterpret_cast<const float4*>(a.ln_gamma + kk);
                const float4 b = *reinterpret_cast<const float4*>(a.ln_beta + kk);
                v.x = (v.x - mean) * rstd * g.x + b.x; v.y = (v.y - mean) * rstd * g.y + b.y;
                v.z = (v.z - mean) * rstd * g.z + b.z; v.w = (v.w - mean) * rstd * g.w + b.w;
              }
              float4 h, l;
              h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
              l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
              const uint32_t so = swz(a_row, c);
              *reinterpret_cast<float4*>(dh + so) = h;
              *reinterpret_cast<float4*>(dl + so) = l;
            }
            fetch_A(kb + 3, buf);
          };
          const int which = kb % 3;
          if (which == 0) produce(ar0); else if (which == 1) produce(ar1); else produce(ar2);
          ++akb;
        }
        trace(tp, unit * 8 + 3);
        if (prefetch) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        trace(tp, unit * 8 + 4);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + slot), "r"(G5_HANDOFF) : "memory");
        trace(tp, unit * 8 + 5);
      }
      // the next tile's statistics (written by other warps) become visible to every producer
      asm volatile("bar.sync %0, %1;" ::"r"(G5_STATS_BAR), "r"(G5_PROD) : "memory");
      par ^= 1;
    }
  } else if (warp == G5_MMA_WARP) {
    // =============================================== MMA issuer ===============================================
    uint32_t unit = 0;
    uint32_t akb = 0, abuf = 0;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const TileCoord tc_ = coord(t);
      const int as = it & 1;
      if (it >= 2) mbar_wait(&sm.acc_empty[as], ((it >> 1) - 1) & 1);     // epilogue of tile it-2 has drained the stage
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int upt = nkb * tc_.nnc;
      for (int ut = 0; ut < upt; ++ut, ++unit) {
        const int kb = ut / tc_.nnc, nc = ut - kb * tc_.nnc;
        const int slot = unit % G5_NB;
        if (nc == 0) abuf = (akb++) & 1;
        trace(tp, 2048 + unit * 4 + 0);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(G5_HANDOFF) : "memory");
        trace(tp, 2048 + unit * 4 + 1);
        if (elect_one()) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int bn = min(G5_BN, tc_.npass - nc * G5_BN);
          const uint32_t idesc = make_idesc(bn);
          const uint64_t dAh = make_desc(smem_u32(sA_hi(abuf))), dAl = make_desc(smem_u32(sA_lo(abuf)));
          const uint64_t dBh = make_desc(smem_u32(sB_hi(slot))), dBl = make_desc(smem_u32(sB_lo(slot)));
          const uint32_t d = tmem + (uint32_t)(as * G5_NPASS + nc * G5_BN);
#pragma unroll
          for (int ks = 0; ks < G5_BK / 8; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);
            umma_tf32(d, dAl + adv, dBh + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            umma_tf32(d, dAh + adv, dBl + adv, idesc, 1u);
            umma_tf32(d, dAh + adv, dBh + adv, idesc, 1u);
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
    const int e = warp - G5_EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = e >> 2;                   // two warps per quarter: even / odd 32-column chunks
    float* stage = stage_base + e * G5_STAGE_FLOATS;
    const int srow = lane >> 3, scol = (lane & 7) * 4;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const TileCoord tc_ = coord(t);
      const int as = it & 1;
      trace(tp, 3584 + it * 4 + 0);
      mbar_wait(&sm.acc_full[as], (it >> 1) & 1);
      trace(tp, 3584 + it * 4 + 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nchunks = (tc_.npass + 31) / 32;
      for (int ch = half; ch < nchunks; ch += 2) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * G5_NPASS + ch * 32), v);
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

int gemm_set_trace(long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_trace, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? NMRF_OK : NMRF_ERR_CUDA;
}

int token_gemm_tc5(const nmrf_gemm_args& a, const float* W_lo, cudaStream_t stream) {
  static int num_sms = 0;
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(token_gemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G5_DYN);
    configured = true;
  }
  const int n_rb = (a.rows + G5_BM - 1) / G5_BM;
  const int n_np = (a.N + G5_NPASS - 1) / G5_NPASS;
  const int ntiles = n_rb * n_np;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  token_gemm_tc5_kernel<<<grid, G5_BLOCK, G5_DYN, stream>>>(a, W_lo, n_rb, n_np);
  count_launch();
  return check_launch("token_gemm_tc5");
}

}  // namespace nmrf
