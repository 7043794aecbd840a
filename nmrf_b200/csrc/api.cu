// extern "C" surface of libnmrf_b200.so (see include/nmrf_b200.h).  Argument validation,
// error text, launch counting; the kernels live in the sibling translation units.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include "common.cuh"

namespace nmrf {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int num_sms() {
  static PerDevice cache;
  const int d = current_device();
  if (d >= 0 && d < kMaxDevices) {
    const int c = cache.v[d].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
  if (n <= 0) n = 148;
  if (d >= 0 && d < kMaxDevices) cache.v[d].store(n, std::memory_order_relaxed);
  return n;
}
int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return NMRF_OK;
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return NMRF_ERR_CUDA;
}

int token_gemm_simt(const nmrf_gemm_args& a, cudaStream_t stream);
int token_gemm_tc6(const nmrf_gemm_args& a, cudaStream_t stream);
int token_gemm_ra(const nmrf_gemm_args& a, cudaStream_t stream);
bool token_gemm_ra_supported(const nmrf_gemm_args& a);
int conv2d_tc6(const nmrf_conv_args& c, cudaStream_t stream);
int row_stats(const float* X, int ldx, int rows, float* stats, cudaStream_t stream);
int gemm6_set_trace(long long* dev_ptr);
int mlp_chain(const nmrf_mlp_args& a, cudaStream_t stream);
int mlp_set_trace(long long* dev_ptr);
int ra_set_trace(long long* dev_ptr);
int pack_weight_tiles(const float* w, int N, int K, float* hi, float* lo, cudaStream_t stream);
int split_tf32(const float* w, float* hi, float* lo, long long n, cudaStream_t stream);
int cost_volume_topk(const float*, const float*, int, int, int, int, int, int, int, float, const nmrf_seed_weights*,
                     float*, float*, int64_t*, cudaStream_t);
int prop_gather(const float*, const int64_t*, int, int, int, int, double, int, float*, int, float*, cudaStream_t);
int prop_head_tail(const float*, const float*, const float*, const int64_t*, int, float*, float*, cudaStream_t);
int proposal_attention(const float*, int, int, float*, cudaStream_t);
int window_attention(const float*, const float*, int, int, int, int, int, int, int, float*, cudaStream_t);
int stripe_attention(const float*, int, int, int, int, const float*, const float*, float*, cudaStream_t);
int stripe_attention_tc(const float*, int, int, int, int, const float*, const float*, float*, cudaStream_t);
bool window_attention_mma_supported(int K, int ws);
int instnorm_stats(const float*, int, int, int, double*, cudaStream_t);
int instnorm_apply(const float*, const double*, const float*, const double*, int, int, int, int, int, float*, cudaStream_t);
int image_prep(const float*, const float*, int, int, int, int, int, long long, long long, long long, long long, float*, cudaStream_t);
int avgpool2_split(const float*, int, int, int, int, float*, float*, float*, cudaStream_t);
int window_attention_mma(const float*, const float*, int, int, int, int, int, int, int, float*, cudaStream_t);
int warp_corr_embed(const float*, const float*, const float*, const float*, const float*, const float*, int, int, int, int, int,
                    int, int, int, double, float*, float*, cudaStream_t);
int zero_pad_rows(float*, int, int, int, int, int, int, int, int, cudaStream_t);
int select_median(const float*, const float*, const float*, const float*, int, int, int, int, int, int, int, int, float*, float*,
                  cudaStream_t);
int refine_tail(const float*, const float*, const float*, int, int, int, int, int, int, int, int, int, float*, float*, cudaStream_t);
int disp_metrics(const float*, const float*, const uint8_t*, int, long long, float, const float*, int, double*, cudaStream_t);
int disp_to_kitti_u16(const float*, long long, uint16_t*, cudaStream_t);
int ms_deform_attn_forward(const float*, const int64_t*, const int64_t*, const float*, const float*, int, int, int, int,
                           int, int, int, float*, bool, cudaStream_t);

}  // namespace nmrf

using namespace nmrf;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
static_assert(sizeof(nmrf_gemm_args) == 144 && sizeof(nmrf_mlp_args) == 112, "ctypes mirrors in nmrf_b200/_lib.py assume these layouts");

// NMRF_B200_ATTN=simt selects the fp32-FMA attention kernels (default: tcgen05 3xTF32)
static std::atomic<int> g_attn_tc{-1};
static bool attn_on_tensor_cores() {
  int v = g_attn_tc.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("NMRF_B200_ATTN");
    v = !(e && (e[0] == 's' || e[0] == 'S'));
    g_attn_tc.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

extern "C" {

int nmrf_abi_version(void) { return NMRF_B200_ABI_VERSION; }
const char* nmrf_last_error(void) { return g_err; }
uint64_t nmrf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nmrf_token_gemm(const nmrf_gemm_args* a, void* stream) {
  NMRF_REQUIRE(a && a->X && a->Y && (a->W || a->Wt_hi), "token_gemm: null pointer");
  NMRF_REQUIRE(a->rows >= 0 && a->N > 0 && a->N % 4 == 0, "token_gemm: rows=%d N=%d (N must be a multiple of 4)", a->rows, a->N);
  NMRF_REQUIRE(a->Kx > 0 && a->Kx % 8 == 0 && a->Ke >= 0 && a->Ke % 8 == 0, "token_gemm: Kx=%d Ke=%d must be multiples of 8", a->Kx, a->Ke);
  NMRF_REQUIRE(a->ldx % 4 == 0 && a->ldy % 4 == 0 && (a->Wt_hi || (a->ldw % 4 == 0 && a->ldw >= a->Kx + a->Ke)), "token_gemm: bad leading dimension");
  NMRF_REQUIRE((a->Ke == 0) == (a->E == nullptr), "token_gemm: E/Ke mismatch");
  NMRF_REQUIRE(a->Ke == 0 || (a->ediv >= 1 && a->lde % 4 == 0), "token_gemm: bad ediv/lde");
  NMRF_REQUIRE((a->ln_gamma == nullptr) == (a->ln_beta == nullptr), "token_gemm: LayerNorm needs gamma and beta");
  NMRF_REQUIRE(a->ln_gamma == nullptr || a->Kx == 128, "token_gemm: LayerNorm prologue needs Kx == 128");
  NMRF_REQUIRE(a->R == nullptr || a->ldr % 4 == 0, "token_gemm: bad ldr");
  NMRF_REQUIRE(a->act >= 0 && a->act <= 2, "token_gemm: act=%d", a->act);
  if (a->rows == 0) return NMRF_OK;
  NMRF_REQUIRE((a->Wt_hi == nullptr) == (a->Wt_lo == nullptr), "token_gemm: Wt_hi and Wt_lo go together");
  if (a->Wt_hi) {
    NMRF_REQUIRE(a->N % 16 == 0 && a->N <= 512, "token_gemm(tc): N=%d must be a multiple of 16, <= 512", a->N);
    // K <= 192 (every nn.Linear of the hot path): A resident in tensor memory for the whole row block (gemm_ra.cu); longer
    // contractions stream A per 128-column tile (gemm_tc6.cu).  NMRF_B200_GEMM_RA=0 (development A/B) forces the latter.
    static const bool ra = [] { const char* e = getenv("NMRF_B200_GEMM_RA"); return !(e && e[0] == '0'); }();
    if (ra && token_gemm_ra_supported(*a)) return token_gemm_ra(*a, ST(stream));
    return token_gemm_tc6(*a, ST(stream));
  }
  return token_gemm_simt(*a, ST(stream));
}
int nmrf_mlp_chain(const nmrf_mlp_args* a, void* stream) {
  NMRF_REQUIRE(a && a->X && a->Wstream && a->Y && a->bias_mid && a->bias_out && a->ln_gamma && a->ln_beta && a->b1,
               "mlp_chain: null pointer");
  NMRF_REQUIRE(a->rows >= 0, "mlp_chain: rows=%d", a->rows);
  NMRF_REQUIRE(a->Kx > 0 && a->Kx % 32 == 0 && a->Ke >= 0 && a->Ke % 32 == 0 && a->Kx + a->Ke <= 512,
               "mlp_chain: Kx=%d Ke=%d must be multiples of 32, sum <= 512", a->Kx, a->Ke);
  NMRF_REQUIRE((a->Ke == 0) == (a->E == nullptr), "mlp_chain: E/Ke mismatch");
  NMRF_REQUIRE(!a->e_identity || a->Ke == 128, "mlp_chain: a residual E (e_identity) must have 128 columns, Ke=%d", a->Ke);
  NMRF_REQUIRE(a->ldx % 4 == 0 && a->ldy % 4 == 0 && (a->Ke == 0 || a->lde % 4 == 0), "mlp_chain: bad leading dimension");
  if (a->rows == 0) return NMRF_OK;
  return mlp_chain(*a, ST(stream));
}
int nmrf_row_stats(const float* X, int ldx, int rows, float* stats, void* stream) { return row_stats(X, ldx, rows, stats, ST(stream)); }
int nmrf_conv2d(const nmrf_conv_args* c, void* stream) {
  NMRF_REQUIRE(c != nullptr, "conv2d: null argument struct");
  return conv2d_tc6(*c, ST(stream));
}
int nmrf_set_attention_impl(int tensor_cores) {
  g_attn_tc.store(tensor_cores ? 1 : 0, std::memory_order_relaxed);
  return NMRF_OK;
}
int nmrf_pack_weight_tiles(const float* w, int N, int K, float* hi_tiles, float* lo_tiles, void* stream) {
  return pack_weight_tiles(w, N, K, hi_tiles, lo_tiles, ST(stream));
}
int nmrf_debug_set_trace(void* dev_i64_4096) {
#ifdef NMRF_TRACE
  mlp_set_trace(reinterpret_cast<long long*>(dev_i64_4096));
  ra_set_trace(reinterpret_cast<long long*>(dev_i64_4096));
  return gemm6_set_trace(reinterpret_cast<long long*>(dev_i64_4096));
#else
  if (dev_i64_4096 == nullptr) return NMRF_OK;
  set_error("nmrf_debug_set_trace: this build has no cycle tracing (rebuild with make TRACE=1)");
  return NMRF_ERR_UNSUPPORTED;
#endif
}
int nmrf_split_tf32(const float* w, float* hi, float* lo, int64_t n, void* stream) {
  return split_tf32(w, hi, lo, (long long)n, ST(stream));
}

int nmrf_cost_volume_topk(const float* f1, const float* f2, int B, int h, int w, int C, int G, int D, int K, float eps,
                          const nmrf_seed_weights* wt, float* cost_volume, float* prob, int64_t* seeds, void* stream) {
  return cost_volume_topk(f1, f2, B, h, w, C, G, D, K, eps, wt, cost_volume, prob, seeds, ST(stream));
}
int nmrf_prop_gather(const float* cv, const int64_t* seeds, int P, int G, int D, int K, double normalizer, int extended,
                     float* cost36, int ld_cost, float* enc32, void* stream) {
  return prop_gather(cv, seeds, P, G, D, K, normalizer, extended, cost36, ld_cost, enc32, ST(stream));
}
int nmrf_stripe_attention(const float* qkv, int B, int h, int w, int K, const float* gv0, const float* gv1, float* out,
                          void* stream) {
  if (attn_on_tensor_cores()) {
    NMRF_REQUIRE(qkv && gv0 && gv1 && out, "stripe_attention: null pointer");
    NMRF_REQUIRE(B * (h > w ? h : w) <= 65535, "stripe_attention: too many stripes for grid.y");
    NMRF_REQUIRE(K >= 1, "stripe_attention: K=%d", K);
    return stripe_attention_tc(qkv, B, h, w, K, gv0, gv1, out, ST(stream));
  }
  return stripe_attention(qkv, B, h, w, K, gv0, gv1, out, ST(stream));
}
int nmrf_prop_head_tail(const float* hidden, const float* w, const float* b, const int64_t* seeds, int T, float* labels,
                        float* labels_lo, void* stream) {
  return prop_head_tail(hidden, w, b, seeds, T, labels, labels_lo, ST(stream));
}
int nmrf_warp_corr_embed(const float* f1_cc, const float* f2_cc, const float* f1_gw, const float* f2_gw,
                         const float* labels, const float* labels_lo, int B, int h, int w, int K, int Hp, int Wp, int top,
                         int left, double normalizer, float* feat160, float* enc32, void* stream) {
  return warp_corr_embed(f1_cc, f2_cc, f1_gw, f2_gw, labels, labels_lo, B, h, w, K, Hp, Wp, top, left, normalizer, feat160,
                         enc32, ST(stream));
}
int nmrf_zero_pad_rows(float* x, int B, int h, int w, int K, int Hp, int Wp, int top, int left, void* stream) {
  return zero_pad_rows(x, B, h, w, K, Hp, Wp, top, left, ST(stream));
}
int nmrf_proposal_attention(const float* qkv, int P, int K, float* out, void* stream) {
  return proposal_attention(qkv, P, K, out, ST(stream));
}
int nmrf_window_attention(const float* qkv, const float* table, int B, int Hp, int Wp, int K, int ws, int shift,
                          int self_edge_mask, float* out, void* stream) {
  if (attn_on_tensor_cores() && window_attention_mma_supported(K, ws)) {
    NMRF_REQUIRE(qkv && table && out, "window_attention: null pointer");
    NMRF_REQUIRE(Hp % ws == 0 && Wp % ws == 0, "window_attention: grid %dx%d not a multiple of ws=%d", Hp, Wp, ws);
    NMRF_REQUIRE(shift >= 0 && shift < ws, "window_attention: shift=%d", shift);
    return window_attention_mma(qkv, table, B, Hp, Wp, K, ws, shift, self_edge_mask, out, ST(stream));
  }
  return window_attention(qkv, table, B, Hp, Wp, K, ws, shift, self_edge_mask, out, ST(stream));
}
int nmrf_instnorm_stats(const float* x, int N, int HW, int C, double* stats, void* stream) {
  return instnorm_stats(x, N, HW, C, stats, ST(stream));
}
int nmrf_instnorm_apply(const float* x, const double* x_stats, const float* r, const double* r_stats, int N, int HW, int C,
                        int relu_inner, int relu_outer, float* out, void* stream) {
  return instnorm_apply(x, x_stats, r, r_stats, N, HW, C, relu_inner, relu_outer, out, ST(stream));
}
int nmrf_image_prep(const float* img1, const float* img2, int B, int H, int W, int Hp, int Wp, int64_t stride_b, int64_t stride_c,
                    int64_t stride_y, int64_t stride_x, float* out_cat3, void* stream) {
  return image_prep(img1, img2, B, H, W, Hp, Wp, stride_b, stride_c, stride_y, stride_x, out_cat3, ST(stream));
}
int nmrf_avgpool2_split(const float* x, int N, int h, int w, int C, float* out_a, float* out_b, float* out_cat3, void* stream) {
  return avgpool2_split(x, N, h, w, C, out_a, out_b, out_cat3, ST(stream));
}
int nmrf_select_median(const float* delta, const float* score, const float* labels, const float* labels_lo, int B, int h,
                       int w, int K, int Hp, int Wp, int top, int left, float* disp_curr, float* disp_curr_lo, void* stream) {
  return select_median(delta, score, labels, labels_lo, B, h, w, K, Hp, Wp, top, left, disp_curr, disp_curr_lo, ST(stream));
}
int nmrf_refine_tail(const float* delta, const float* disp_curr, const float* disp_curr_lo, int B, int h4, int w4, int Hp4,
                     int Wp4, int top, int left, int H, int W, float* disp_pred, float* disp, void* stream) {
  return refine_tail(delta, disp_curr, disp_curr_lo, B, h4, w4, Hp4, Wp4, top, left, H, W, disp_pred, disp, ST(stream));
}
int nmrf_disp_metrics(const float* disp_pr, const float* disp_gt, const uint8_t* valid_gt, int B, int64_t HW, float max_disp,
                      const float* thresholds_host, int n_thres, double* acc, void* stream) {
  return disp_metrics(disp_pr, disp_gt, valid_gt, B, (long long)HW, max_disp, thresholds_host, n_thres, acc, ST(stream));
}
int nmrf_disp_to_kitti_u16(const float* disp, int64_t n, uint16_t* out, void* stream) {
  return disp_to_kitti_u16(disp, (long long)n, out, ST(stream));
}
int nmrf_ms_deform_attn_forward(const float* value, const int64_t* shapes, const int64_t* level_start, const float* loc,
                                const float* attn, int N, int S, int M, int Dh, int L, int Lq, int P, float* out,
                                void* stream) {
  return ms_deform_attn_forward(value, shapes, level_start, loc, attn, N, S, M, Dh, L, Lq, P, out, false, ST(stream));
}
int nmrf_ms_deform_attn_forward_dev(const float* value, const int64_t* shapes_dev, const int64_t* level_start_dev,
                                    const float* loc, const float* attn, int N, int S, int M, int Dh, int L, int Lq, int P,
                                    float* out, void* stream) {
  return ms_deform_attn_forward(value, shapes_dev, level_start_dev, loc, attn, N, S, M, Dh, L, Lq, P, out, true, ST(stream));
}

}  // extern "C"
