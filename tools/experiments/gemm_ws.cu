// Fused token GEMM on tcgen05 with the WEIGHTS STATIONARY IN TENSOR MEMORY (K <= 192).
//
// gemm_tc6.cu streams a 32 KB weight tile per 32-wide k-block through every SM (plus the activations, re-read once per
// 128-column chunk, plus the LayerNorm statistics pass): ~75 KB per 768 tensor cycles, about twice what the L2 -> SM path
// delivers, so its MMA warp waits for operands half of the time.  Here the operand roles are swapped:
//     D^T[f, t] = sum_k W[f, k] * A[t, k]        M = 128 output features, N = 64 tokens, K = 8 per MMA
// the "A" operand of the MMA is the weight chunk of the CTA -- hi and lo parts, loaded ONCE into tensor memory (2 x 192
// columns) -- and the "B" operand is the activation tile, produced into shared memory by the producer warps
// (coalesced cp.async raw ring -> LayerNorm -> hi/lo split -> swizzled K-major [64 tokens x 32] tiles).  No weight traffic
// at all after the prologue; the accumulator holds Y transposed (lane = feature, column = token), which makes the bias a
// per-thread scalar and every store of a warp one 128-byte row segment of Y.
//
//   warps 0-7   producers    thread = (token row, two 16-byte chunks of the k-block); raw ring 4 x 8 KB, operand ring 4 x 16 KB
//   warp  8     MMA issuer   per unit (64 tokens x 32 k): 4 k-steps x {W_lo.A_hi, W_hi.A_lo, W_hi.A_hi}, 384 tensor cycles
//   warps 9-16  epilogue     LayerNorm statistics of the tile two ahead; TMEM -> registers -> bias / activation / residual ->
//                            Y[t][f] (warp = 32 features x 32 tokens)
//   TMEM map    [0,128) two accumulator stages of 64 token columns, [128,320) W_hi, [320,512) W_lo
//   tiles       CTA c owns feature chunk c % n_nc and the token tiles c / n_nc, + grid / n_nc, ...
// Arithmetic identical to gemm_tc6 (3xTF32, RN hi / exact lo; fp32 LayerNorm two-pass; same GELU).
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int W_BT = 64, W_BF = 128, W_BK = 32;     // tokens per tile, features per chunk, k-block
constexpr int W_RAW = 4, W_OPS = 4, W_ACC = 2, W_STATS = 4;
constexpr int W_RAWT = W_BT * W_BK * 4;              // 8 KB raw tile
constexpr int W_OPT = 2 * W_RAWT;                    // 16 KB operand stage: hi tile | lo tile
constexpr int W_PROD = 256, W_MMA_WARP = 8, W_EPI_WARP0 = 9, W_EPI_WARPS = 8;
constexpr int W_BLOCK = (W_EPI_WARP0 + W_EPI_WARPS) * 32;      // 544
constexpr int W_RAW_BAR = 5;
constexpr int W_COL_WH = 128, W_COL_WL = 320, W_KMAX = 192;
constexpr int W_DYN = W_RAW * W_RAWT + W_OPS * W_OPT + 1024;

struct WSmem {
  uint64_t op_full[W_OPS];     // operand stage written (8 producer-warp arrivals)
  uint64_t op_free[W_OPS];     // MMAs that read the stage are complete (commit)
  uint64_t acc_full[W_ACC];    // accumulator stage holds a finished tile (commit)
  uint64_t acc_empty[W_ACC];   // drained (8 epilogue-warp arrivals)
  uint64_t stats_full[W_STATS];
  uint32_t tmem_base;
  float mean[W_STATS][W_BT], rstd[W_STATS][W_BT];
  alignas(16) float gamma[128];
  alignas(16) float beta[128];
};

__device__ __forceinline__ void mbar_arrive_w(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t idesc_w(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// LayerNorm statistics (Kx == 128) of 8 rows of a 64-row tile by one warp: 8 lanes per row, four rows per pass
__device__ __noinline__ void tile_stats_w(const float* __restrict__ X, int ldx, int rows, int row0, int e, float* mean, float* rstd) {
  const int lane = threadIdx.x & 31, g = lane >> 3, sub = lane & 7;
  float4 v[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = row0 + e * 8 + i * 4 + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v[i][j] = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + j * 32 + sub * 4);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
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
    if (sub == 0) { const int lr = e * 8 + i * 4 + g; mean[lr] = mu; rstd[lr] = 1.f / sqrtf(q * (1.f / 128.f) + 1e-5f); }
  }
}

template <int ACT, bool LN>
__global__ void __launch_bounds__(W_BLOCK, 1)
token_gemm_ws_kernel(const nmrf_gemm_args a, int n_tt, int n_nc, int cpc) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ WSmem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  auto sRaw = [&](int i) { return base + i * W_RAWT; };
  auto sOp = [&](int i) { return base + W_RAW * W_RAWT + i * W_OPT; };      // hi tile at +0, lo tile at +8 KB

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ktot = a.Kx + a.Ke;
  const int nkb = (Ktot + W_BK - 1) / W_BK;
  const int nc = blockIdx.x % n_nc, tt0 = blockIdx.x / n_nc;               // feature chunk, first token tile; tiles step by cpc
  const int n_base = nc * W_BF;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < W_OPS; ++i) { mbar_init(&sm.op_full[i], 8); mbar_init(&sm.op_free[i], 1); }
    for (int i = 0; i < W_ACC; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], W_EPI_WARPS); }
    for (int i = 0; i < W_STATS; ++i) mbar_init(&sm.stats_full[i], W_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (LN && tid < 128) { sm.gamma[tid] = a.ln_gamma[tid]; sm.beta[tid] = a.ln_beta[tid]; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  // ---- prologue: this CTA's weight chunk (hi and lo, rows n_base .. +128, columns 0 .. 32 nkb) into tensor memory; thread =
  // feature row = TMEM lane (warps 9-12 cover the four lane quarters) ----
  if (warp >= W_EPI_WARP0 && warp < W_EPI_WARP0 + 4) {
    const int q = warp & 3, f = n_base + q * 32 + lane;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    for (int part = 0; part < 2; ++part) {
      const float* W = part == 0 ? a.W : a.W_lo;
      for (int kb = 0; kb < nkb; ++kb) {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (f < a.N) v = *reinterpret_cast<const float4*>(W + (size_t)f * a.ldw + kb * 32 + c * 4);
          r[c * 4] = __float_as_uint(v.x); r[c * 4 + 1] = __float_as_uint(v.y); r[c * 4 + 2] = __float_as_uint(v.z); r[c * 4 + 3] = __float_as_uint(v.w);
        }
        const uint32_t col = (uint32_t)((part == 0 ? W_COL_WH : W_COL_WL) + kb * 32);
        tmem_st16(tmem + t_lane + col, r);
        tmem_st16(tmem + t_lane + col + 16, r + 16);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp < 8) {
    // =============================================== producers ===============================================
    // fetch: 512 16-byte chunks per k-block, 2 per thread: chunk id = tid + 256 j -> row id / 8, chunk id % 8 (8 lanes = one row)
    const int f_c = tid & 7, f_r = tid >> 3;                       // rows f_r and f_r + 32
    int f_tt = tt0, f_kb = 0;
    const float* f_x[2];
    const float* f_e[2];
    uint32_t f_ok = 0;
    auto fetch_tile = [&]() {
      const int row0 = f_tt * W_BT;
      f_ok = 0;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int grow = row0 + f_r + 32 * j;
        const bool ok = grow < a.rows;
        f_ok |= (ok ? 1u : 0u) << j;
        const int gr = ok ? grow : 0;
        f_x[j] = a.X + (size_t)gr * a.ldx + f_c * 4;
        f_e[j] = a.E ? a.E + (size_t)(gr / a.ediv) * a.lde + f_c * 4 - a.Kx : a.X;
      }
    };
    if (f_tt < n_tt) fetch_tile();
    auto fetch_next = [&](uint32_t stage) {
      if (f_tt < n_tt) {
        const uint32_t dst = smem_u32(sRaw(stage));
        const int k0 = f_kb * W_BK;
        const bool in_x = k0 + f_c * 4 < a.Kx, in_k = k0 + f_c * 4 < Ktot;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bool ok = in_k && ((f_ok >> j) & 1u);
          const float* src = (in_x ? f_x[j] : f_e[j]) + k0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + swz(f_r + 32 * j, f_c)), "l"(ok ? src : a.X), "r"(ok ? 16 : 0));
        }
        if (++f_kb == nkb) { f_kb = 0; f_tt += cpc; if (f_tt < n_tt) fetch_tile(); }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch_next(0); fetch_next(1); fetch_next(2);
    // split: thread = (row tid / 4, chunks 2 (tid % 4) and + 1)
    const int p_row = tid >> 2, p_c = (tid & 3) * 2;
    uint32_t unit = 0;
    int it = 0;
    for (int tt = tt0; tt < n_tt; tt += cpc, ++it) {
      const bool row_ok = tt * W_BT + p_row < a.rows;
      float mean = 0.f, rstd = 1.f;
      if (LN) {
        mbar_wait_warp(&sm.stats_full[it % W_STATS], (it / W_STATS) & 1);
        mean = sm.mean[it % W_STATS][p_row]; rstd = sm.rstd[it % W_STATS][p_row];
      }
      for (int kb = 0; kb < nkb; ++kb, ++unit) {
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(W_RAW_BAR), "r"(W_PROD) : "memory");
        fetch_next((unit + 3) % W_RAW);                            // the stage consumed one unit ago is free again
        const uint8_t* raw = sRaw(unit % W_RAW);
        float4 h[2], l[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = p_c + j, kk = kb * W_BK + c * 4;
          float4 v = *reinterpret_cast<const float4*>(raw + swz(p_row, c));
          if (LN && row_ok && kk < a.Kx) {
            const float4 g = *reinterpret_cast<const float4*>(sm.gamma + kk);
            const float4 b = *reinterpret_cast<const float4*>(sm.beta + kk);
            v.x = (v.x - mean) * rstd * g.x + b.x; v.y = (v.y - mean) * rstd * g.y + b.y;
            v.z = (v.z - mean) * rstd * g.z + b.z; v.w = (v.w - mean) * rstd * g.w + b.w;
          }
          h[j].x = rna_tf32_fast(v.x); h[j].y = rna_tf32_fast(v.y); h[j].z = rna_tf32_fast(v.z); h[j].w = rna_tf32_fast(v.w);
          l[j].x = v.x - h[j].x; l[j].y = v.y - h[j].y; l[j].z = v.z - h[j].z; l[j].w = v.w - h[j].w;
        }
        const uint32_t st = unit % W_OPS;
        if (unit >= W_OPS) mbar_wait_warp(&sm.op_free[st], ((unit / W_OPS) - 1) & 1);     // MMAs of unit - 4 have read the stage
        uint8_t* op = sOp(st);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t so = swz(p_row, p_c + j);
          *reinterpret_cast<float4*>(op + so) = h[j];
          *reinterpret_cast<float4*>(op + W_RAWT + so) = l[j];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_w(&sm.op_full[st]);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == W_MMA_WARP) {
    // =============================================== MMA issuer ===============================================
    const uint32_t idesc = idesc_w(W_BT);
    uint32_t unit = 0;
    int it = 0;
    for (int tt = tt0; tt < n_tt; tt += cpc, ++it) {
      const int as = it % W_ACC;
      if (it >= W_ACC) {
        mbar_wait_warp(&sm.acc_empty[as], ((it / W_ACC) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int kb = 0; kb < nkb; ++kb, ++unit) {
        const uint32_t st = unit % W_OPS;
        mbar_wait_warp(&sm.op_full[st], (unit / W_OPS) & 1);
        if (elect_one()) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ob = smem_u32(sOp(st));
          const uint64_t dAh = make_desc(ob), dAl = make_desc(ob + W_RAWT);
          const uint32_t d = tmem + (uint32_t)(as * W_BT);
          const uint32_t wh = tmem + (uint32_t)(W_COL_WH + kb * 32), wl = tmem + (uint32_t)(W_COL_WL + kb * 32);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);
            umma_tf32_ta(d, wl + ks * 8, dAh + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            umma_tf32_ta(d, wh + ks * 8, dAl + adv, idesc, 1u);
            umma_tf32_ta(d, wh + ks * 8, dAh + adv, idesc, 1u);
          }
          umma_commit(&sm.op_free[st]);
          if (kb == nkb - 1) umma_commit(&sm.acc_full[as]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================================== epilogue ===============================================
    const int e = warp - W_EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quarter = features n_base + 32 q .. + 32
    const int half = e >> 2;                   // token columns 32 half .. + 32 of the tile
    const int f = n_base + q * 32 + lane;
    const bool f_ok = f < a.N;
    const float bias = (a.bias && f_ok) ? a.bias[f] : 0.f;
    auto stats = [&](int j) {
      const int tj = tt0 + j * cpc;
      if (tj < n_tt) {
        tile_stats_w(a.X, a.ldx, a.rows, tj * W_BT, e, sm.mean[j % W_STATS], sm.rstd[j % W_STATS]);
        __syncwarp();
        if (lane == 0) mbar_arrive_w(&sm.stats_full[j % W_STATS]);
      }
    };
    if (LN) { stats(0); stats(1); stats(2); }
    int it = 0;
    for (int tt = tt0; tt < n_tt; tt += cpc, ++it) {
      const int as = it % W_ACC;
      // statistics buffer (it + 3) % 4 last held tile it - 1, whose operands the producers have finished (its MMAs were drained here)
      if (LN) stats(it + 3);
      mbar_wait_warp(&sm.acc_full[as], (it / W_ACC) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * W_BT + half * 32), v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_w(&sm.acc_empty[as]);
      const int t_base = tt * W_BT + half * 32;
      if (f_ok) {
        if (a.R) {
          // residual first (R may alias Y): all loads of this thread before its stores
          float rr[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) rr[t] = (t_base + t < a.rows) ? __ldcg(a.R + (size_t)(t_base + t) * a.ldr + f) : 0.f;
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (t_base + t < a.rows) a.Y[(size_t)(t_base + t) * a.ldy + f] = act_fast(v[t] + bias, ACT) + rr[t];
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (t_base + t < a.rows) a.Y[(size_t)(t_base + t) * a.ldy + f] = act_fast(v[t] + bias, ACT);
        }
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
void launch_ws(const nmrf_gemm_args& a, int n_tt, int n_nc, int cpc, cudaStream_t stream) {
  static bool configured = false;     // per instantiation
  if (!configured) {
    cudaFuncSetAttribute(token_gemm_ws_kernel<ACT, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_DYN);
    configured = true;
  }
  token_gemm_ws_kernel<ACT, LN><<<n_nc * cpc, W_BLOCK, W_DYN, stream>>>(a, n_tt, n_nc, cpc);
}
}  // namespace

bool token_gemm_ws_supported(const nmrf_gemm_args& a) {
  const int kpad = ((a.Kx + a.Ke + 31) / 32) * 32;
  return a.W_lo != nullptr && kpad <= W_KMAX && a.ldw >= kpad && a.ldy % 1 == 0;
}

int token_gemm_ws(const nmrf_gemm_args& a, cudaStream_t stream) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_tt = (a.rows + W_BT - 1) / W_BT;
  const int n_nc = (a.N + W_BF - 1) / W_BF;
  int cpc = num_sms / n_nc;                       // CTAs per feature chunk
  if (cpc < 1) cpc = 1;
  if (cpc > n_tt) cpc = n_tt;
  const bool ln = a.ln_gamma != nullptr;
  switch (a.act * 2 + (ln ? 1 : 0)) {
    case 0: launch_ws<0, false>(a, n_tt, n_nc, cpc, stream); break;
    case 1: launch_ws<0, true>(a, n_tt, n_nc, cpc, stream); break;
    case 2: launch_ws<1, false>(a, n_tt, n_nc, cpc, stream); break;
    case 3: launch_ws<1, true>(a, n_tt, n_nc, cpc, stream); break;
    case 4: launch_ws<2, false>(a, n_tt, n_nc, cpc, stream); break;
    case 5: launch_ws<2, true>(a, n_tt, n_nc, cpc, stream); break;
    default: set_error("token_gemm: unknown activation %d", a.act); return NMRF_ERR_BAD_ARG;
  }
  count_launch();
  return check_launch("token_gemm_ws");
}

}  // namespace nmrf
