// EXPERIMENT, not built into libnmrf_b200.so: CTA-pair schedule of the block tail for the 64-wide-chunk weight stream of commit
// 0c94aff..; passed tests/test_gpu_stages.py::test_mlp_chain on B200 but ran 87 us against 78 us for the single-CTA kernel.
// nmrf_mlp_chain on CTA PAIRS (tcgen05 cta_group::2): the same fused block tail as gemm_mlp.cu
//     x1 = [att | x] . [Wproj | I]^T + b_proj ;  x = x1 + fc2(GELU(fc1(LN2(x1))))        (NMP.py:358-363, 570-573)
// but two CTAs of a cluster take a 256-token tile together.  Each CTA owns 128 rows (its A operands, accumulators, LayerNorm,
// GELU and stores are exactly those of gemm_mlp.cu); the B operand of every MMA is SPLIT: a CTA stages only its half of the
// weight rows (N/2) and the leader CTA issues one 256 x N x 8 MMA for both.  The single-CTA kernel streams 1.25 MB of weights
// per 128 tokens through every SM's L2 port (337 MB per launch, ~24 B/clk/SM sustained, 4000-cycle bulk-copy latencies) and is
// bound by that; here the weight bytes per SM are halved and the 96 KB ring holds six units instead of three.
//
//   per CTA, as gemm_mlp.cu:  TMEM [0,128) acc0, [128,384) LN2(x1) hi | lo, [384,512) phase-1 A buffers / fc1 accumulators;
//                             smem: weight ring 6 x 16 KB (this CTA's half of a unit: hi half | lo half), hidden chunk 64 KB,
//                             raw-A ring 3 x 16 KB;  warps 0-7 producers + GELU workers, 8 MMA, 9-16 LN / GELU / store, 17 TMA
//   cross-CTA protocol:       everything the MMA waits for (A operands written, hidden chunk written, accumulators drained,
//                             the peer's weight half landed) is an mbarrier in the LEADER's shared memory on which the warps
//                             of both CTAs arrive (release.cluster; the leader waits acquire.cluster); everything that follows
//                             an MMA (slot free, accumulator full, ...) is a tcgen05.commit multicast to the same barrier in
//                             both CTAs.  The peer's MMA warp only forwards "my half of unit g has landed".
#include "common.cuh"
#include "tc_common.cuh"

namespace nmrf {
namespace {
using namespace tc;

constexpr int P_BM = 128, P_BK = 32, P_NB = 6, P_RAW = 3;
constexpr int P_TILE = P_BM * P_BK * 4;            // 16 KB: a [128 x 32] fp32 image
constexpr int P_SLOT = P_TILE;                     // this CTA's half of a weight unit: 8 KB hi + 8 KB lo
constexpr int P_UNIT = 2 * P_TILE;                 // a unit of the weight stream in global memory (hi image + lo image)
constexpr int P_HID = 512, P_CH = 64, P_NCH = P_HID / P_CH;
constexpr int P_PROD = 256, P_MMA_WARP = 8, P_EPI_WARP0 = 9, P_EPI_WARPS = 8, P_TMA_WARP = 17;
constexpr int P_BLOCK = (P_TMA_WARP + 1) * 32;     // 576
constexpr int P_RAW_BAR = 5;
constexpr int P_COL_ALN_HI = 128, P_COL_ALN_LO = 256, P_COL_X = 384;
constexpr int P_OFF_H = P_NB * P_SLOT;             // 96 KB
constexpr int P_OFF_RAW = P_OFF_H + 4 * P_TILE;    // + 64 KB
constexpr int P_DYN = P_OFF_RAW + P_RAW * P_TILE + 1024;

struct P2Smem {
  uint64_t done[P_NB];        // both CTAs: MMAs of the unit that used weight slot s are complete (multicast commit)
  uint64_t full_b[P_NB];      // local: this CTA's half of the unit landed (expect_tx 16 KB)
  uint64_t peer_full[P_NB];   // leader: the peer's half landed (1 remote arrival)
  uint64_t a_full[2];         // leader: phase-1 A buffer written by the producers of both CTAs (16 warp arrivals)
  uint64_t p1_full;           // both: acc0 = x1 - b_proj complete (commit)
  uint64_t aln_full;          // leader: LN2(x1) in TMEM in both CTAs (16)
  uint64_t acc1_full[2];      // both: fc1 chunk accumulator complete (commit)
  uint64_t acc1_empty[2];     // leader: drained in both CTAs (32)
  uint64_t h_full;            // leader: hidden chunk in shared memory in both CTAs (32)
  uint64_t h_free;            // both: fc2 MMAs of the chunk complete (commit)
  uint64_t acc0_final;        // both: all MMAs of the tile complete (commit)
  uint64_t acc0_empty;        // leader: final epilogue has read acc0 in both CTAs (16)
  uint32_t tmem_base;
  alignas(16) float gamma[128];
  alignas(16) float beta[128];
  alignas(16) float bmid[128];
  alignas(16) float bout[128];
  alignas(16) float b1[P_HID];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release at cluster scope) on the LEADER CTA's copy of a barrier; valid from either CTA of the pair
__device__ __forceinline__ void arrive_leader(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// whole-warp wait (uniform poll) with cluster-scope acquire: for barriers the peer CTA arrives on
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  bool ok = false;
#pragma unroll 1
  for (uint32_t spin = 0; spin < SPIN_LIMIT && !ok; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    ok = done != 0;
  }
  if (!ok) __trap();
}
__device__ __forceinline__ uint32_t idesc2(int n) {   // M = 256 (two CTAs x 128 rows), N columns, tf32 in / f32 out, K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void umma2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all MMAs issued so far -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// one hidden chunk on one of the 16 GELU workers of a CTA (see gemm_mlp.cu: gelu_worker)
__device__ __forceinline__ void gelu_worker2(P2Smem& sm, uint32_t tmem_lane, uint8_t* sH, int row, int j, uint32_t gc, const float* b1, int lane) {
  const int b = gc & 1;
  mbar_wait_warp(&sm.acc1_full[b], (gc >> 1) & 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float v[16];
  tmem_ld16(tmem_lane + (uint32_t)(P_COL_X + b * 64 + j * 16), v);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (lane == 0) arrive_leader(&sm.acc1_empty[b]);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i] + b1[i]);
  if (gc >= 1) mbar_wait_warp(&sm.h_free, (gc - 1) & 1);
  uint8_t* hi_t = sH + (j >> 1) * P_TILE;
  uint8_t* lo_t = hi_t + 2 * P_TILE;
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
  if (lane == 0) arrive_leader(&sm.h_full);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_BLOCK, 1)
mlp_chain2_kernel(const nmrf_mlp_args a, int npt) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ P2Smem sm;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)dsm + 1023) & ~(uintptr_t)1023);
  auto sW = [&](int slot) { return base + slot * P_SLOT; };            // hi half at +0, lo half at +8 KB
  uint8_t* sH = base + P_OFF_H;
  auto sRaw = [&](int i) { return base + P_OFF_RAW + i * P_TILE; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int Ktot = a.Kx + a.Ke;
  const int n1 = Ktot / P_BK;
  const int upt = n1 + 4 * P_NCH;
  const int rot = pair & 7;                        // must be the same in both CTAs of the pair: they share the weight stream

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < P_NB; ++i) { mbar_init(&sm.done[i], 1); mbar_init(&sm.full_b[i], 1); mbar_init(&sm.peer_full[i], 1); }
    mbar_init(&sm.a_full[0], 16); mbar_init(&sm.a_full[1], 16);
    mbar_init(&sm.p1_full, 1); mbar_init(&sm.aln_full, 2 * P_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.acc1_full[i], 1); mbar_init(&sm.acc1_empty[i], 2 * (P_EPI_WARPS + 8)); }
    mbar_init(&sm.h_full, 2 * (P_EPI_WARPS + 8)); mbar_init(&sm.h_free, 1);
    mbar_init(&sm.acc0_final, 1); mbar_init(&sm.acc0_empty, 2 * P_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 128) { sm.gamma[tid] = a.ln_gamma[tid]; sm.beta[tid] = a.ln_beta[tid]; sm.bmid[tid] = a.bias_mid[tid]; sm.bout[tid] = a.bias_out[tid]; }
  if (tid < P_HID) sm.b1[tid] = a.b1[tid];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                              // both CTAs' barriers are initialised before any remote arrival / multicast
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp < 8) {
    // =============================================== producers (phase 1) + GELU workers 0, 1 ===============================================
    const int a_row = (warp & 3) * 32 + lane, a_c0 = (warp >> 2) * 4;
    const uint32_t a_lane = ((uint32_t)((warp & 3) * 32)) << 16;
    const int f_c = tid & 7, f_r = tid >> 3;
    int f_pt = pair, f_kb = 0;
    const float* f_x[4];
    const float* f_e[4];
    uint32_t f_ok = 0;
    auto fetch_tile = [&]() {
      const int row0 = (f_pt * 2 + (int)rank) * P_BM;
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
    if (f_pt < npt) fetch_tile();
    auto fetch_next = [&](uint32_t stage) {
      if (f_pt < npt) {
        const uint32_t dst = smem_u32(sRaw(stage));
        const int k0 = ((f_kb + rot) % n1) * P_BK;
        const bool in_x = k0 + f_c * 4 < a.Kx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = (f_ok >> j) & 1u;
          const float* src = (in_x ? f_x[j] : f_e[j]) + k0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + swz(f_r + 32 * j, f_c)), "l"(ok ? src : a.X), "r"(ok ? 16 : 0));
        }
        if (++f_kb == n1) { f_kb = 0; f_pt += npairs; if (f_pt < npt) fetch_tile(); }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch_next(0); fetch_next(1);
    uint32_t pu = 0;
    int it = 0;
    for (int pt = pair; pt < npt; pt += npairs, ++it) {
      for (int kb = 0; kb < n1; ++kb, ++pu) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(P_RAW_BAR), "r"(P_PROD) : "memory");
        fetch_next((pu + 2) % P_RAW);
        const uint8_t* raw = sRaw(pu % P_RAW);
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
        if (kb >= 2) {
          const uint32_t g = (uint32_t)it * upt + kb - 2;
          mbar_wait_warp(&sm.done[g % P_NB], (g / P_NB) & 1);
        } else if (it > 0) {
          mbar_wait_warp(&sm.acc0_final, (it - 1) & 1);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem + a_lane + (uint32_t)(P_COL_X + (pu & 1) * 64 + a_c0 * 4);
        tmem_st16(ta, hi);
        tmem_st16(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) arrive_leader(&sm.a_full[pu & 1]);
      }
      for (int c = 0; c < P_NCH; ++c)
        gelu_worker2(sm, tmem + a_lane, sH, a_row, warp >> 2, (uint32_t)(it * P_NCH + c),
                     sm.b1 + ((c + rot) & 7) * P_CH + (warp >> 2) * 16, lane);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == P_MMA_WARP) {
    if (leader) {
      // =============================================== MMA issuer (leader CTA, for both) ===============================================
      const uint32_t idesc128 = idesc2(128), idesc64 = idesc2(64);
      const uint32_t acc0 = tmem;
      const uint64_t dHh0 = make_desc(smem_u32(sH)), dHl0 = make_desc(smem_u32(sH + 2 * P_TILE));
      uint32_t g = 0, pu = 0, gc = 0;
      auto wait_b = [&](uint32_t unit) {
        mbar_wait_warp(&sm.full_b[unit % P_NB], (unit / P_NB) & 1);
        mbar_wait_cluster(&sm.peer_full[unit % P_NB], (unit / P_NB) & 1);
      };
      // F1 unit: 8 k-steps of k-block pair p against this chunk's 64 fc1 rows (32 per CTA); A = LN2(x1) from TMEM
      auto issue_f1 = [&](uint32_t cg, int p) {
        wait_b(g);
        if (elect_one()) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t bslot = smem_u32(sW(g % P_NB));
          const uint32_t d = tmem + (uint32_t)(P_COL_X + (cg & 1) * 64);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t dBh = make_desc(bslot + (kk >> 2) * 4096) + (uint64_t)((kk & 3) * 2);
            const uint64_t dBl = make_desc(bslot + 8192 + (kk >> 2) * 4096) + (uint64_t)((kk & 3) * 2);
            const uint32_t kcol = (uint32_t)(p * 64 + kk * 8);
            umma2_ts(d, tmem + P_COL_ALN_LO + kcol, dBh, idesc64, (p > 0 || kk > 0) ? 1u : 0u);
            umma2_ts(d, tmem + P_COL_ALN_HI + kcol, dBl, idesc64, 1u);
            umma2_ts(d, tmem + P_COL_ALN_HI + kcol, dBh, idesc64, 1u);
          }
          commit2(&sm.done[g % P_NB]);
          if (p == 1) commit2(&sm.acc1_full[cg & 1]);
        }
        __syncwarp();
        ++g;
      };
      // F2 unit: 4 k-steps of hidden k-block q against the 128 fc2 rows (64 per CTA); A = hidden chunk from shared memory
      auto issue_f2 = [&](int q, bool last_of_chunk, bool last_of_tile) {
        wait_b(g);
        if (elect_one()) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t bslot = smem_u32(sW(g % P_NB));
          const uint64_t dBh = make_desc(bslot), dBl = make_desc(bslot + 8192);
          const uint64_t dAh = dHh0 + (uint64_t)(q * (P_TILE >> 4)), dAl = dHl0 + (uint64_t)(q * (P_TILE >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);
            umma2_ss(acc0, dAl + adv, dBh + adv, idesc128, 1u);
            umma2_ss(acc0, dAh + adv, dBl + adv, idesc128, 1u);
            umma2_ss(acc0, dAh + adv, dBh + adv, idesc128, 1u);
          }
          commit2(&sm.done[g % P_NB]);
          if (last_of_chunk) commit2(&sm.h_free);
          if (last_of_tile) commit2(&sm.acc0_final);
        }
        __syncwarp();
        ++g;
      };
      int it = 0;
      for (int pt = pair; pt < npt; pt += npairs, ++it) {
        if (it > 0) mbar_wait_cluster(&sm.acc0_empty, (it - 1) & 1);
        for (int kb = 0; kb < n1; ++kb, ++pu, ++g) {
          mbar_wait_cluster(&sm.a_full[pu & 1], (pu >> 1) & 1);
          wait_b(g);
          if (elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bslot = smem_u32(sW(g % P_NB));
            const uint64_t dBh = make_desc(bslot), dBl = make_desc(bslot + 8192);
            const uint32_t tAh = tmem + (uint32_t)(P_COL_X + (pu & 1) * 64), tAl = tAh + 32;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 2);
              umma2_ts(acc0, tAl + ks * 8, dBh + adv, idesc128, (kb > 0 || ks > 0) ? 1u : 0u);
              umma2_ts(acc0, tAh + ks * 8, dBl + adv, idesc128, 1u);
              umma2_ts(acc0, tAh + ks * 8, dBh + adv, idesc128, 1u);
            }
            commit2(&sm.done[g % P_NB]);
            if (kb == n1 - 1) commit2(&sm.p1_full);
          }
          __syncwarp();
        }
        mbar_wait_cluster(&sm.aln_full, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c <= P_NCH; ++c) {
          if (c < P_NCH) {
            const uint32_t cg = gc + c;
            if (cg >= 2) {
              mbar_wait_cluster(&sm.acc1_empty[cg & 1], ((cg >> 1) - 1) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            issue_f1(cg, 0);
            issue_f1(cg, 1);
          }
          if (c >= 1) {
            mbar_wait_cluster(&sm.h_full, (gc + c - 1) & 1);
            issue_f2(0, false, false);
            issue_f2(1, true, c == P_NCH);
          }
        }
        gc += P_NCH;
      }
    } else {
      // peer CTA: forward "my half of unit g has landed" to the leader
      uint32_t g = 0;
      for (int pt = pair; pt < npt; pt += npairs)
        for (int u = 0; u < upt; ++u, ++g) {
          mbar_wait_warp(&sm.full_b[g % P_NB], (g / P_NB) & 1);
          if (lane == 0) arrive_leader(&sm.peer_full[g % P_NB]);
          __syncwarp();
        }
    }
  } else if (warp == P_TMA_WARP) {
    // =============================================== weight stream: this CTA's half of every unit ===============================================
    if (elect_one()) {
      uint32_t g = 0;
      for (int pt = pair; pt < npt; pt += npairs) {
        for (int u = 0; u < upt; ++u, ++g) {
          const int slot = g % P_NB;
          if (g >= P_NB) mbar_wait(&sm.done[slot], ((g - P_NB) / P_NB) & 1);
          const uint32_t bar = smem_u32(&sm.full_b[slot]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(P_SLOT) : "memory");
          int su;
          bool is_f1 = false;
          if (u < n1) {
            su = (u + rot) % n1;
          } else {
            const int s3 = u - n1, blk = s3 >> 1, pq = s3 & 1;
            const bool is_f2 = blk >= 2 && (blk == 15 || (blk & 1) == 0);
            const int cs = is_f2 ? (blk == 15 ? 7 : blk / 2 - 1) : (blk == 0 ? 0 : (blk + 1) / 2);
            su = n1 + (is_f2 ? 16 : 0) + ((cs + rot) & 7) * 2 + pq;
            is_f1 = !is_f2;
          }
          const float* src = a.Wstream + (size_t)su * (P_UNIT / 4);
          const uint32_t dst = smem_u32(sW(slot));
          if (!is_f1) {
            // [128 n x 32 k] images: rows 64 rank .. +64 of the hi image and of the lo image
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src + rank * 2048), "r"(8192), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + 8192), "l"(src + 4096 + rank * 2048), "r"(8192), "r"(bar) : "memory");
          } else {
            // two [64 n x 32 k] sub-images per part: rows 32 rank .. +32 of each
#pragma unroll
            for (int part = 0; part < 2; ++part)
#pragma unroll
              for (int sub = 0; sub < 2; ++sub)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst + part * 8192 + sub * 4096), "l"(src + part * 4096 + sub * 2048 + rank * 1024), "r"(4096), "r"(bar) : "memory");
          }
        }
      }
    }
  } else {
    // =============================================== LN / GELU / store warps ===============================================
    const int e = warp - P_EPI_WARP0;
    const int q = warp & 3;
    const int half = e >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = ((uint32_t)(q * 32)) << 16;
    uint8_t* stage = sH + half * P_TILE + q * 32 * 128;      // staging of the final store: rows of the hidden buffer nobody writes then
    const int srow = lane >> 3, sc8 = lane & 7;
    uint32_t gc = 0;
    int it = 0;
    for (int pt = pair; pt < npt; pt += npairs, ++it) {
      const int row0 = (pt * 2 + (int)rank) * P_BM;
      // ---- LN2 of x1 = acc0 + b_proj, thread = row, two-pass; this warp normalises columns 64 half .. +64 ----
      mbar_wait_warp(&sm.p1_full, it & 1);
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
        tmem_st16(tmem + t_lane + (uint32_t)(P_COL_ALN_HI + ch * 32), hi);
        tmem_st16(tmem + t_lane + (uint32_t)(P_COL_ALN_HI + ch * 32 + 16), hi + 16);
        tmem_st16(tmem + t_lane + (uint32_t)(P_COL_ALN_LO + ch * 32), lo);
        tmem_st16(tmem + t_lane + (uint32_t)(P_COL_ALN_LO + ch * 32 + 16), lo + 16);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) arrive_leader(&sm.aln_full);
      // ---- hidden chunks: GELU workers 2, 3 of the quarter ----
      for (int c = 0; c < P_NCH; ++c, ++gc)
        gelu_worker2(sm, tmem + t_lane, sH, row, 2 + half, gc, sm.b1 + ((c + rot) & 7) * P_CH + (2 + half) * 16, lane);
      // ---- final: x = acc0 + (b_proj + b_fc2) ----
      mbar_wait_warp(&sm.acc0_final, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float w[32];
      tmem_ld32(tmem + t_lane + (uint32_t)(half * 64), v);
      tmem_ld32(tmem + t_lane + (uint32_t)(half * 64 + 32), w);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) arrive_leader(&sm.acc0_empty);
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const float* src = ch == 0 ? v : w;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8)
          *reinterpret_cast<float4*>(stage + swz(lane, c8)) = make_float4(src[c8 * 4], src[c8 * 4 + 1], src[c8 * 4 + 2], src[c8 * 4 + 3]);
        __syncwarp();
        const int n = half * 64 + ch * 32 + sc8 * 4;
        const float4 bo = *reinterpret_cast<const float4*>(sm.bout + n);
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const int lr = i8 * 4 + srow;
          const int r = row0 + q * 32 + lr;
          if (r < a.rows) {
            float4 o = *reinterpret_cast<const float4*>(stage + swz(lr, sc8));
            o.x += bo.x; o.y += bo.y; o.z += bo.z; o.w += bo.w;
            *reinterpret_cast<float4*>(a.Y + (size_t)r * a.ldy + n) = o;
          }
        }
        __syncwarp();
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  __syncwarp();
  cluster_sync_all();                              // no CTA leaves while its peer may still arrive on its barriers / use its TMEM
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

}  // namespace

int mlp_chain2(const nmrf_mlp_args& a, cudaStream_t stream) {
  static int num_sms = 0;
  static bool configured = false;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!configured) {
    cudaFuncSetAttribute(mlp_chain2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_DYN);
    configured = true;
  }
  const int npt = (a.rows + 2 * P_BM - 1) / (2 * P_BM);       // 256-row pair tiles
  int pairs = num_sms / 2;
  if (pairs > npt) pairs = npt;
  mlp_chain2_kernel<<<2 * pairs, P_BLOCK, P_DYN, stream>>>(a, npt);
  count_launch();
  return check_launch("mlp_chain2");
}

}  // namespace nmrf
