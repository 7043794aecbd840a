// Fused token GEMM on the 5th-generation tensor cores (tcgen05, accumulators in TMEM), exact to fp32
// level through error-compensated 3xTF32:
//     x = hi + lo,  hi = rna_tf32(x),  lo = rna_tf32(x - hi);      x.w ~= hi.hi + hi.lo + lo.hi
// (round-to-nearest splitting is essential: DESIGN.md §3 -- truncation splitting fails the EPE bar).
// The weights arrive pre-split (W_hi, W_lo: nmrf_split_tf32); the activations are LayerNorm'ed,
// concatenated and split on the fly by the producer threads while they stage the A operand.
//
// Structure (one persistent CTA per SM, 256 threads, 1 CTA/SM because of TMEM):
//   tile      : 128 token rows x all N (<=512) output columns; fp32 accumulators = N TMEM columns
//   operands  : K-major, SWIZZLE_128B, BK = 32 fp32 (=128 B) per k-block
//                 A_hi/A_lo [128 x 32]  double buffered  (2 x 32 KB)   filled by st.shared
//                 B_hi/B_lo [128 x 32]  triple buffered  (3 x 32 KB)   filled by cp.async, one unit ahead
//   schedule  : "units" u = (k-block, n-chunk of 128).  Iteration u: wait for the MMAs of unit u-2 (frees the
//               B buffer of unit u+1 and the A buffer about to be rewritten), start the cp.async of B(u+1),
//               produce A if the k-block changed, wait for B(u), fence to the async proxy, __syncthreads, then
//               ONE thread issues 12 tcgen05.mma (4 k-steps x 3 products) and commits to the unit's mbarrier.
//               MMA(u) is issued while MMA(u-1) is still running; the weight prefetch runs across tiles.
//   roles     : warps 0-7 produce (and run the epilogue); warp 8 only issues MMAs.  Producers do not wait for
//               the issue: they bar.arrive on the unit's named barrier (id 1 + unit%3) and go on filling; the MMA
//               warp bar.syncs on it.  (tcgen05.mma issue back-pressures on the tensor queue; issuing from a
//               producer thread stalled that warp's share of the next fill and everybody behind a __syncthreads.)
//   epilogue  : tcgen05.ld 32x32b (each warp its TMEM lane quarter) -> bias/act/residual -> global.
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int TC_THREADS = 256;             // producer / epilogue threads (warps 0-7)
constexpr int TC_BLOCK = TC_THREADS + 32;   // + warp 8: the MMA issuer
constexpr int TC_BM = 128;
constexpr int TC_BN = 128;                 // n-chunk per MMA
constexpr int TC_BK = 32;                  // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;   // 16 KB per operand tile

constexpr int TC_NB = 3;    // B buffers
struct TcSmem {
  // dynamic shared memory, 1024-byte aligned: A_hi[2] A_lo[2] B_hi[3] B_lo[3] (16 KB each)
  uint64_t bar[TC_NB];      // one per B buffer: completion of the MMAs of the unit that used it
  uint32_t tmem_base;
  float mean[2][TC_BM], rstd[2][TC_BM];   // double buffered: the next tile's statistics are computed before this tile's epilogue
};

__global__ void __launch_bounds__(TC_BLOCK, 1)
token_gemm_tc_kernel(const nmrf_gemm_args a, const float* __restrict__ W_lo, int ntiles, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ TcSmem sm;
  // carve the operand tiles (the dynamic window is 1024-byte aligned by the launch: see launcher)
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  // tile order: A_hi[0..1] A_lo[0..1] B_hi[0..2] B_lo[0..2]   (address arithmetic, no pointer arrays -> no local memory)
  auto sA_hi = [&](int i) { return base + i * TC_TILE_BYTES; };
  auto sA_lo = [&](int i) { return base + (2 + i) * TC_TILE_BYTES; };
  auto sB_hi = [&](int i) { return base + (4 + i) * TC_TILE_BYTES; };
  auto sB_lo = [&](int i) { return base + (7 + i) * TC_TILE_BYTES; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ktot = a.Kx + a.Ke;
  const int nkb = (Ktot + TC_BK - 1) / TC_BK;
  const int nnc = (a.N + TC_BN - 1) / TC_BN;
  const bool ln = a.ln_gamma != nullptr;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < TC_NB; ++i) mbar_init(&sm.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  uint32_t unit = 0;                 // global unit counter of this CTA (B buffer = unit % 3, use = unit / 3)
  uint32_t akb = 0, abuf = 0;        // global k-block counter: A buffer = akb & 1 alternates across tiles as well
  const int upt = nkb * nnc;         // units per tile

  // producer mapping for A: thread -> (row = tid/2, 16 consecutive k = 4 chunks of 16 B)
  const int a_row = tid >> 1, a_c0 = (tid & 1) * 4;

  // B tile of local unit `ut` (k-block ut / nnc, n-chunk ut % nnc) -> buffer b; 16-byte cp.async, swizzled
  auto load_B = [&](int ut, int b) {
    const int kb = ut / nnc, n0 = (ut % nnc) * TC_BN;
    const int bn = min(TC_BN, a.N - n0);
    for (int i = tid; i < bn * 8; i += TC_THREADS) {
      const int r = i >> 3, c = i & 7;
      const size_t goff = (size_t)(n0 + r) * a.ldw + kb * TC_BK + c * 4;
      const uint32_t so = swz(r, c);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB_hi(b) + so)), "l"(a.W + goff));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB_lo(b) + so)), "l"(W_lo + goff));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (blockIdx.x < ntiles && warp < 8) load_B(0, 0);        // prologue: the first unit's weights

  // per-tile producer state (set by begin_tile): row pointers of this thread, prefetched raw A values
  int row0 = 0;
  bool row_ok = false;
  const float* xrow = a.X;
  const float* erow = a.E;
  // raw (pre-LayerNorm) A values of this thread: 4 chunks of 4 floats per k-block, fetched THREE k-blocks ahead
  // (the rows come from HBM/L2: ~1-2 us latency vs ~0.4 us of MMA per unit); three buffers used round-robin
  float4 ar0[4], ar1[4], ar2[4];
  auto fetch_A = [&](int kb, float4 (&dst)[4]) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int kk = kb * TC_BK + (a_c0 + cc) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok && kk < Ktot) v = (kk < a.Kx) ? *reinterpret_cast<const float4*>(xrow + kk)
                                               : *reinterpret_cast<const float4*>(erow + (kk - a.Kx));
      dst[cc] = v;
    }
  };
  // start a tile: row pointers, first three k-blocks of A in flight, LayerNorm statistics into buffer `par`
  // (Kx == 128: one warp per 16 rows).  The caller provides the barrier before the statistics are read.
  auto begin_tile = [&](int tile, int par) {
    row0 = tile * TC_BM;
    const int g_row = row0 + a_row;
    row_ok = g_row < a.rows;
    xrow = a.X + (size_t)(row_ok ? g_row : 0) * a.ldx;
    erow = a.E ? a.E + (size_t)((row_ok ? g_row : 0) / a.ediv) * a.lde : nullptr;
    fetch_A(0, ar0);
    fetch_A(1, ar1);          // k-blocks beyond Ktot read nothing (zero fill)
    fetch_A(2, ar2);
    if (ln) {
      for (int i = 0; i < TC_BM / 8; ++i) {
        const int lr = warp * (TC_BM / 8) + i, r = row0 + lr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < a.rows) v = *reinterpret_cast<const float4*>(a.X + (size_t)r * a.ldx + lane * 4);
        const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / 128.f);
        const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
        const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f);
        if (lane == 0) { sm.mean[par][lr] = mean; sm.rstd[par][lr] = 1.f / sqrtf(var + 1e-5f); }
      }
    }
  };

  int par = 0;
  if (blockIdx.x < ntiles) {
    if (warp < 8) begin_tile(blockIdx.x, 0);
    __syncthreads();
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const bool has_next_tile = tile + (int)gridDim.x < ntiles;
    const int cur_row0 = row0;
    const float mean = (ln && warp < 8) ? sm.mean[par][a_row] : 0.f, rstd = (ln && warp < 8) ? sm.rstd[par][a_row] : 1.f;

    if (warp < 8) {   // ================= producers =================
    for (int ut = 0; ut < upt; ++ut, ++unit) {
      const int kb = ut / nnc, nc = ut % nnc;
      const int buf = unit % TC_NB;
      // MMAs of unit-2 (hence of every earlier unit) complete: frees B[(unit+1)%3] and the A buffer of k-block kb-2.
      // (unit-2 is the newest unit whose barrier phase is unambiguous: its buffer is next used by unit+1.)
      if (unit >= 2) mbar_wait(&sm.bar[(unit - 2) % TC_NB], ((unit - 2) / TC_NB) & 1);
      // ---- prefetch the next unit's weights (next tile starts again at k-block 0, n-chunk 0) ----------------
      const bool prefetch = (ut + 1 < upt) || has_next_tile;
      if (prefetch) load_B((ut + 1 < upt) ? ut + 1 : 0, (unit + 1) % TC_NB);
      // ---- A: produced once per k-block (nc == 0): LayerNorm, concat, split hi/lo from the prefetched registers;
      //      the buffer is immediately re-armed with k-block kb+3 (three buffers used round-robin: no register moves,
      //      so the loads stay in flight for three k-blocks of MMAs)
      if (nc == 0) {
        auto produce = [&](float4 (&buf)[4]) {
          uint8_t* dh = sA_hi(akb & 1);
          uint8_t* dl = sA_lo(akb & 1);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int c = a_c0 + cc;
            const int kk = kb * TC_BK + c * 4;
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
      // B(unit) has landed (only the group just committed for unit+1 may still be in flight)
      if (prefetch) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      asm volatile("bar.arrive %0, %1;" ::"r"(1 + (int)(unit % TC_NB)), "r"(TC_BLOCK) : "memory");   // hand unit to the MMA warp
    }
    } else {          // ================= MMA issuer (warp 8) ========
    for (int ut = 0; ut < upt; ++ut, ++unit) {
      const int kb = ut / nnc, nc = ut % nnc;
      const int buf = unit % TC_NB;
      if (nc == 0) abuf = (akb++) & 1;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + buf), "r"(TC_BLOCK) : "memory");      // unit's operands are in shared memory
      if (elect_one()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int n0 = nc * TC_BN;
        const int bn = min(TC_BN, a.N - n0);
        const uint32_t idesc = make_idesc(bn);
        const uint64_t dAh = make_desc(smem_u32(sA_hi(abuf))), dAl = make_desc(smem_u32(sA_lo(abuf)));
        const uint64_t dBh = make_desc(smem_u32(sB_hi(buf))), dBl = make_desc(smem_u32(sB_lo(buf)));
        const uint32_t d = tmem + (uint32_t)n0;
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);             // +32 bytes (>>4) inside the 128-byte swizzle row
          umma_tf32(d, dAl + adv, dBh + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);   // small terms first
          umma_tf32(d, dAh + adv, dBl + adv, idesc, 1u);
          umma_tf32(d, dAh + adv, dBh + adv, idesc, 1u);
        }
        umma_commit(&sm.bar[buf]);
      }
      __syncwarp();
    }
    }
    // ---- next tile: rows, A prefetch and LayerNorm statistics go out BEFORE this tile's epilogue, so their
    //      latency overlaps the last MMAs and the epilogue --------------------------------------------------
    if (warp < 8 && has_next_tile) begin_tile(tile + (int)gridDim.x, par ^ 1);
    par ^= 1;
    // ---- wait for the tile's last MMAs, then the epilogue (producer warps) ---------------------------
    if (warp < 8) {
      const uint32_t last = unit - 1;
      mbar_wait(&sm.bar[last % TC_NB], (last / TC_NB) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __syncwarp();
    }
    if (warp < 8) {
      // warp w reads TMEM lanes 32*(w%4)..+32 (its row quarter); warps 0-3 take even 32-column chunks, 4-7 odd.
      // The 32x32 block (lane = row) is transposed through a per-warp staging tile in the (now idle) A buffers so
      // that global traffic is coalesced: 8 lanes cover one row's 128 bytes, a warp instruction covers 4 rows.
      const int q = warp & 3, half = warp >> 2;
      float* stage = reinterpret_cast<float*>(base) + warp * (32 * 36);
      const int nchunks = (a.N + 31) / 32;
      const int srow = lane >> 3, scol = (lane & 7) * 4;
      for (int ch = half; ch < nchunks; ch += 2) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int n = ch * 32 + scol;
        if (n < a.N) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.bias) b = *reinterpret_cast<const float4*>(a.bias + n);
          // all residual loads first (R may alias Y: in-place residual; every element is read before it is written,
          // by the same thread -- hoisting keeps the 8 loads in flight instead of load->store->load serialisation)
          float4 rr[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = cur_row0 + q * 32 + it * 4 + srow;
            rr[it] = (a.R && r < a.rows) ? __ldcg(reinterpret_cast<const float4*>(a.R + (size_t)r * a.ldr + n))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int lr = it * 4 + srow;
            const int r = cur_row0 + q * 32 + lr;
            if (r < a.rows) {
              float4 o = *reinterpret_cast<const float4*>(stage + lr * 36 + scol);
              o.x = act_fn(o.x + b.x, a.act) + rr[it].x; o.y = act_fn(o.y + b.y, a.act) + rr[it].y;
              o.z = act_fn(o.z + b.z, a.act) + rr[it].z; o.w = act_fn(o.w + b.w, a.act) + rr[it].w;
              *reinterpret_cast<float4*>(a.Y + (size_t)r * a.ldy + n) = o;
            }
          }
        }
        __syncwarp();
      }
    }
    // TMEM (and the LN statistics) are reused by the next tile
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = w[i];
  const float h = rna_tf32(x);
  hi[i] = h;
  lo[i] = rna_tf32(x - h);
}

}  // namespace

// W = W_hi, W_lo pre-split; ldw must be a multiple of 32 and cover Kx+Ke rounded up to 32 (zero padded)
int token_gemm_tc(const nmrf_gemm_args& a, const float* W_lo, cudaStream_t stream) {
  static int num_sms = 0;
  static bool configured = false;
  const size_t dyn = 10 * TC_TILE_BYTES + 1024;
  if (!configured) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(token_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    configured = true;
  }
  const int ntiles = (a.rows + TC_BM - 1) / TC_BM;
  int cols = 32;
  while (cols < a.N) cols <<= 1;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  token_gemm_tc_kernel<<<grid, TC_BLOCK, dyn, stream>>>(a, W_lo, ntiles, cols);
  count_launch();
  return check_launch("token_gemm_tc");
}

int split_tf32(const float* w, float* hi, float* lo, long long n, cudaStream_t stream) {
  NMRF_REQUIRE(w && hi && lo && n >= 0, "split_tf32: bad arguments");
  if (n == 0) return NMRF_OK;
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, hi, lo, n);
  count_launch();
  return check_launch("split_tf32");
}

}  // namespace nmrf
