/*
 * nmrf_b200.h -- C-ABI of libnmrf_b200.so: the NMRF-Stereo inference hot path on B200 (sm_100a).
 *
 * Drop-in boundary for aeolusguan/NMRF (paths below are relative to that repository).
 * Every entry point
 *   - takes plain device pointers + sizes + a CUDA stream (passed as void*, = cudaStream_t),
 *   - never allocates, frees or synchronises (safe to capture in a CUDA graph),
 *   - returns 0 on success, non-zero NMRF_ERR_* otherwise; nmrf_last_error() gives the text.
 * All tensors are fp32 unless noted, contiguous, 16-byte aligned.  Feature maps are NHWC
 * (torch channels_last); token tensors are [rows, 128] row-major, row = ((b*H + y)*W + x)*K + n,
 * exactly the reference's '(b h w) n c' order (nmrf/models/NMP.py:347,360).
 *
 * Packed weights: the reference's nn.Linear weights ([out,in] row-major) are used as they are,
 * except that the input dimension is zero-padded to a multiple of 16 and q/k/v projections
 * that share an input are stacked along `out` (see nmrf_b200/hotpath.py: PackedWeights, which does
 * the packing with torch ops at load time; layouts are documented per struct below).
 *
 * Extended labels.  The disparity labels that travel between stages (`labels` after the propagation head, `disp_curr`
 * after the selection) feed Fourier features with frequencies up to 2^14 (NMP.py:42-48): one fp32 ulp of a label
 * (~2e-6 px) is ~1.5e-3 rad in the top frequency, the largest single rounding effect of the reference's fp32 forward and
 * a source of flipped argmax / median decisions (NMRF.py:228-231).  Every entry point that produces or consumes a label
 * therefore takes an optional `*_lo` pointer: label = hi + lo (two fp32 words, the sum formed in double).  `*_lo == NULL`
 * reproduces the reference's plain fp32 arithmetic; the `hi` word alone is always the fp32 value the reference returns.
 */
#ifndef NMRF_B200_H
#define NMRF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMRF_B200_ABI_VERSION 7

enum {
  NMRF_OK = 0,
  NMRF_ERR_BAD_ARG = 1,      /* null pointer / unsupported dimension */
  NMRF_ERR_UNSUPPORTED = 2,  /* valid in the reference but outside this build's limits */
  NMRF_ERR_CUDA = 3          /* cudaGetLastError() != cudaSuccess after a launch */
};

/* ---- library helpers ------------------------------------------------------------------- */
int nmrf_abi_version(void);
const char* nmrf_last_error(void);          /* thread-local, valid until the next call */
/* number of kernels this library has launched since load (bench.py's "gpu_launches") */
uint64_t nmrf_launch_count(void);

/* attention kernels: 1 = tcgen05 3xTF32 (default), 0 = fp32 FMA.  Process-wide; the environment variable
 * NMRF_B200_ATTN=simt selects 0 at first use.  (The GEMM path is chosen per call by nmrf_gemm_args.W_lo.) */
int nmrf_set_attention_impl(int tensor_cores);

/* ---- generic fused token GEMM -------------------------------------------------------------
 * Y[r, n] = act( sum_k A[r,k] * W[n,k] + bias[n] ) (+ R[r,n])
 * A[r, :] = concat( LN?(X[r, 0:Kx]), E[r / ediv, 0:Ke] ),   W is [N, ldw] with ldw >= Kx+Ke.
 * This is the dense part of every nn.Linear on the path (NMP.py:82,84,326,332,337,522,525,537,
 * 607-612,675; NMRF.py:82-83,105; DPN.py:65) with LayerNorm (eps 1e-5), concat, bias, ReLU/GELU
 * and the residual add fused in.  Kx, Ke, ldw, ldx, lde, ldy, ldr must be multiples of 4.
 */
typedef struct {
  const float* X; int ldx; int Kx;
  const float* E; int lde; int Ke; int ediv;   /* E may be NULL (Ke = 0) */
  const float* ln_gamma; const float* ln_beta; /* both NULL = no LayerNorm; else Kx must be 128 */
  const float* ln_stats;                       /* optional [rows, 2] = (mean, 1 / sqrt(var + 1e-5)) of every row of X, as written by
                                                  nmrf_mlp_chain (out_stats) or nmrf_row_stats; NULL: the kernel computes them */
  const float* W; int ldw;
  const float* bias;                           /* [N] or NULL */
  const float* R; int ldr;                     /* residual or NULL; may alias Y */
  float* Y; int ldy;
  int rows; int N;
  int act;                                     /* 0 none, 1 ReLU, 2 GELU (erf) */
  /* Tensor-core path (tcgen05 kind::tf32, error-compensated 3xTF32 with round-to-nearest hi AND lo operand parts,
   * fp32-level accuracy): the weight as hi / lo TILE IMAGES produced by nmrf_pack_weight_tiles --
   * [ceil(N/128)][ceil(K/32)] tiles of 128 rows x 32 fp32, each tile stored exactly as its SWIZZLE_128B shared-memory
   * image (16 KB), so a tile is fetched with one TMA bulk copy (cp.async.bulk).  N % 16 == 0, N <= 512; W may then be
   * NULL.  Wt_hi == Wt_lo == NULL selects the exact-fp32 FMA kernel on W. */
  const float* Wt_hi;
  const float* Wt_lo;
} nmrf_gemm_args;
int nmrf_token_gemm(const nmrf_gemm_args* a, void* stream);
/* LayerNorm statistics of every row of X [rows, 128] (two-pass, eps 1e-5): stats [rows, 2] = (mean, rstd) */
int nmrf_row_stats(const float* X, int ldx, int rows, float* stats, void* stream);
/* w [N,K] row-major (device) -> hi/lo tile images, K zero-padded to a multiple of 32, N to a multiple of 128:
 * hi_tiles / lo_tiles must hold ceil(N/128)*ceil(K/32)*4096 floats each. */
int nmrf_pack_weight_tiles(const float* w, int N, int K, float* hi_tiles, float* lo_tiles, void* stream);
/* ---- fused block tail: proj + residual + LayerNorm + Mlp in ONE launch ----------------------------
 *   x1 = E[r, 0:128] + (X[r, 0:Kx] . W1^T + bias_mid)                (e_identity = 1; SwinNMP / CSWinNMP: x + proj(attn),
 *                                                                    NMP.py:358-359, 570-571: E is the residual stream x; it is
 *                                                                    held in fp32 registers and added with round-to-nearest
 *                                                                    adds, never inside a tensor-core accumulator)
 *   x1 = concat(X[r, 0:Kx], E[r, 0:Ke]) . W1cat^T + bias_mid         (e_identity = 0: E is an ordinary concatenated operand)
 *   Y  = x1 + (fc2( GELU( fc1( LN(x1) ) ) ) + bias_out)              (x + Mlp(norm2(x)), NMP.py:362-363,572-573; timm Mlp 128->512->128;
 *                                                                    bias_out = the fc2 bias)
 * Wstream: n1 + 32 units of 8192 floats, n1 = Kx/32 (e_identity) or (Kx+Ke)/32 (16 KB hi image + 16 KB lo image, SWIZZLE_128B
 * shared-memory images of [128 x 32] fp32 tiles):  P1(0..n1-1), then F1(c) at n1 + c, then F2(c) at n1 + 16 + c (c = hidden chunk
 * of 32, 0..15; every CTA walks the k-blocks and the chunks starting from its own rotation, so the fp32 summation order differs
 * per tile):  P1(j)[n,k] = W1[n, 32j+k];  F1(c)[r,k] = Wfc1[32c + r%32, 32(r/32) + k];  F2(c)[n,k] = Wfc2[n, 32c+k]
 * (nmrf_b200/ops.py: pack_mlp_stream).  Kx and Ke must be multiples of 32, Kx + Ke <= 512; Y may alias E (every element
 * is read and later written by the same thread).  Same 3xTF32 arithmetic as nmrf_token_gemm; the fc2 accumulator is drained
 * into the registers every four hidden chunks (at most 48 MMAs ever accumulate into the same tensor-memory columns). */
typedef struct {
  const float* X; int ldx; int Kx;
  const float* E; int lde; int Ke;             /* may be NULL (Ke = 0) */
  const float* Wstream;
  const float* bias_mid;                       /* [128] */
  const float* ln_gamma; const float* ln_beta; /* [128] */
  const float* b1;                             /* [512] fc1 bias */
  const float* bias_out;                       /* [128] fc2 bias */
  float* Y; int ldy;
  int rows;
  float* out_stats;                            /* optional [rows, 2]: (mean, rstd) of every OUTPUT row (the next block's LayerNorm) */
  int e_identity;                              /* 1: E [rows,128] is the residual, added in fp32 registers (not part of the weight stream) */
} nmrf_mlp_args;
int nmrf_mlp_chain(const nmrf_mlp_args* a, void* stream);

/* debug tooling: device buffer of 4096 int64 that CTA 0 of the tensor-core GEMMs fills with clock64() stamps (NULL = off).
 * Only in builds made with `make TRACE=1`; release builds carry no tracing code and return NMRF_ERR_UNSUPPORTED. */
int nmrf_debug_set_trace(void* dev_i64_4096);
/* hi = rna_tf32(w), lo = rna_tf32(w - hi), elementwise over n floats (device pointers) */
int nmrf_split_tf32(const float* w, float* hi, float* lo, int64_t n, void* stream);

/* ---- A1+A2: cost volume + seed extraction ------------------------------------------------
 * replaces build_correlation_volume (nmrf/models/submodule.py:4-23) and DPN.forward step 1
 * (nmrf/models/DPN.py:115-125): conv1d 4->8->16->1 (k5,pad2)+ReLU, softmax over D, 1-D NMS
 * (max_pool1d k3), suppressed := eps, top-K.
 * Tie rule (reference: unspecified torch.topk order): value descending, then index ascending.
 */
typedef struct {
  const float* w0; const float* b0;   /* [8,G,5], [8]   dpn.mlp.0 */
  const float* w1; const float* b1;   /* [16,8,5], [16] dpn.mlp.2 */
  const float* w2; const float* b2;   /* [1,16,5], [1]  dpn.mlp.4 */
} nmrf_seed_weights;
int nmrf_cost_volume_topk(const float* f1_nhwc, const float* f2_nhwc,   /* [B,h,w,C] */
                          int B, int h, int w, int C, int G, int D, int K, float eps,
                          const nmrf_seed_weights* wt,
                          float* cost_volume,   /* [B*h*w, G, D] */
                          float* prob,          /* [B*h*w, D]    */
                          int64_t* seeds,       /* [B*h*w, K] int64, descending prob */
                          void* stream);

/* ---- A3+A4 gather part: sample_cost + Fourier embed ----------------------------------------
 * replaces Propagation.sample_cost (NMP.py:618-634) and fourier_coord_embed (NMP.py:35-51).
 * cost36 row layout: [g*9 + (o+4)] for o in -4..4, zero-padded to ld_cost (>= 36, mult of 16).
 * enc row layout: 15 sin, 15 cos, raw coordinate, then zero up to 32.
 */
int nmrf_prop_gather(const float* cost_volume, const int64_t* seeds, int P, int G, int D, int K,
                     double normalizer,
                     int extended,                /* 1: encoding evaluated in double (see "Extended labels"), 0: fp32 as the reference */
                     float* cost36, int ld_cost,  /* [P*K, ld_cost] */
                     float* enc32,                /* [P*K, 32] */
                     void* stream);

/* ---- A6: cross-shaped (stripe) attention with LePE ------------------------------------------
 * replaces CSWinAttention.forward + get_rpe with split_size 1 (NMP.py:429-505).
 * qkv: [B*h*w*K, 384] = q|k|v, 4 heads x 32.  Heads 0,1 attend inside image columns
 * (H_sp=h, W_sp=1), heads 2,3 inside image rows.  Self-edge mask NMP.py:203-208.
 * get_v0 / get_v1: attns.{0,1}.get_v.weight [64,1,3,3].
 */
int nmrf_stripe_attention(const float* qkv, int B, int h, int w, int K,
                          const float* get_v0, const float* get_v1,
                          float* out /* [B*h*w*K,128] */, void* stream);

/* ---- A7 tail: labels = relu(hidden . w + b + seed) ------------------------------------------
 * last layer of dpn.prop_head (DPN.py:65,131-132). hidden: [T,128] (after the two ReLU layers). */
int nmrf_prop_head_tail(const float* hidden, const float* w /*[128]*/, const float* b /*[1]*/,
                        const int64_t* seeds, int T, float* labels /*[T]*/, float* labels_lo /*[T] or NULL*/,
                        void* stream);

/* ---- A8: warp + group-wise correlation + Fourier embed ---------------------------------------
 * replaces Inference.sample_fmap/corr (NMP.py:682-720,735-743) and Refinement's K=1 case
 * (NMP.py:839-846).  Writes on the zero-padded token grid (NMP.py:745-762): Hp,Wp multiples of
 * the window size, top/left = centre-pad offsets; pad tokens get zeros.
 * feat row: [f1_cc(64) | warp f2_cc(64) | corr(32)], enc row as in nmrf_prop_gather.
 */
int nmrf_warp_corr_embed(const float* f1_cc, const float* f2_cc,   /* [B,h,w,64]  NHWC */
                         const float* f1_gw, const float* f2_gw,   /* [B,h,w,256] NHWC */
                         const float* labels,                      /* [B*h*w, K] */
                         const float* labels_lo,                   /* [B*h*w, K] or NULL */
                         int B, int h, int w, int K, int Hp, int Wp, int top, int left,
                         double normalizer,
                         float* feat160, float* enc32,             /* [B*Hp*Wp*K, 160 | 32] */
                         void* stream);
/* zero the rows of x[B*Hp*Wp*K, 128] that are padding (label_rep is padded AFTER ffn) */
int nmrf_zero_pad_rows(float* x, int B, int h, int w, int K, int Hp, int Wp, int top, int left,
                       void* stream);

/* ---- A10: attention among the K proposals of a pixel ------------------------------------------
 * core of BasicAttention.forward_pre (NMP.py:97-103). qkv [P*K,384] -> out [P*K,128]. K <= 8. */
int nmrf_proposal_attention(const float* qkv, int P, int K, float* out, void* stream);

/* ---- A11: (shifted-)window attention with contextual relative position encoding ---------------
 * replaces WindowAttention.forward (NMP.py:241-289) incl. masks (NMP.py:195-239, 802-826).
 * qkv [B*Hp*Wp*K, 384]; table = relative_position_enc_table [(2ws-1)^2, 384], column
 * h*96 + {0:32 Rq, 32:64 Rk, 64:96 Rv}.  shift in {0, ws/2}: done by indexing, no roll.
 * self_edge_mask: 1 for Inference (K proposals), 0 for Refinement (NMP.py:869).
 */
int nmrf_window_attention(const float* qkv, const float* table,
                          int B, int Hp, int Wp, int K, int ws, int shift, int self_edge_mask,
                          float* out /* [B*Hp*Wp*K,128] */, void* stream);

/* ---- A12: proposal selection ---------------------------------------------------------------------
 * replaces NMRF.forward select (NMRF.py:218-232): coarse = relu(label + delta); 8x8 un-shuffle;
 * argmax over K of score (first max on ties); x2; lower median of each 4x4 block -> disp_curr.
 * delta, score: [B*Hp*Wp*K, 64] on the PADDED token grid (rows of pad tokens are ignored);
 * labels [B*h*w, K].  disp_curr [B, 2h, 2w] (units: 1/4-res pixels).
 */
int nmrf_select_median(const float* delta, const float* score, const float* labels,
                       const float* labels_lo,     /* [B*h*w, K] or NULL */
                       int B, int h, int w, int K, int Hp, int Wp, int top, int left,
                       float* disp_curr,
                       float* disp_curr_lo,        /* [B, 2h, 2w]; NULL iff labels_lo is NULL */
                       void* stream);

/* ---- A13 tail: disp = relu(disp_curr + delta) un-shuffled -------------------------------------
 * replaces NMRF.py:238-245,250-251. delta [B*Hp4*Wp4, 16] on the padded 1/4 grid;
 * disp_pred [B, 4*h4, 4*w4] (1/4-res units), disp [B, H, W] = 4*disp_pred cropped (unpad).
 */
int nmrf_refine_tail(const float* delta, const float* disp_curr,
                     const float* disp_curr_lo,    /* [B, h4, w4] or NULL */
                     int B, int h4, int w4, int Hp4, int Wp4, int top, int left, int H, int W,
                     float* disp_pred, float* disp, void* stream);

/* ---- A14: multi-scale deformable attention forward ---------------------------------------------
 * replaces ms_deform_attn_forward (ops/src/vision.cpp:13-16, ops/src/ms_deform_attn.h:20-39,
 * ops/src/cuda/ms_deform_attn_cuda.cu:20-80, ms_deform_im2col_cuda.cuh:237-299).
 * value [N,S,M,Dh]; spatial_shapes [L,2] int64 (H,W) and level_start_index [L] int64 are HOST
 * pointers (read once per call); loc [N,Lq,M,L,P,2]; attn [N,Lq,M,L,P]; out [N,Lq,M*Dh].
 * Any Dh (vectorised kernel when Dh % 4 == 0, as in NMRF: Dh = 8); fp32 only (the reference also
 * dispatches fp64).
 */
int nmrf_ms_deform_attn_forward(const float* value, const int64_t* spatial_shapes_host,
                                const int64_t* level_start_index_host,
                                const float* sampling_loc, const float* attn_weight,
                                int N, int S, int M, int Dh, int L, int Lq, int P,
                                float* out, void* stream);
/* same, with spatial_shapes / level_start_index as DEVICE pointers (the reference's convention:
 * ms_deform_im2col_cuda.cuh:274-277 reads them in the kernel).  No host read => CUDA-graph safe. */
int nmrf_ms_deform_attn_forward_dev(const float* value, const int64_t* spatial_shapes_dev,
                                    const int64_t* level_start_index_dev,
                                    const float* sampling_loc, const float* attn_weight,
                                    int N, int S, int M, int Dh, int L, int Lq, int P,
                                    float* out, void* stream);

/* ---- N1 / N2 (next rows of the scope table): k x k convolution over an NHWC image on tcgen05 ------------------------------
 * replaces nn.Conv2d in the reference's Backbone and conv heads (nmrf/models/backbone.py:16-98, NMRF.py:56-65,
 * DPN.py:45-49): Y[n, yo, xo, :] = bias + sum_{ky, kx, c} X[n, yo*stride - pad + ky, xo*stride - pad + kx, c] * W[:, ky, kx, c]
 * (zero padding), as an implicit GEMM on the token-GEMM kernel: GEMM row = output pixel, k = (ky*kw + kx)*Cin + c, the A
 * tile of a (tap, 32-channel block) gathered with zero-filled cp.async straight from X -- no im2col buffer, no hi/lo copy of
 * the activations (the operand split happens in the producer warps).  Same error-compensated 3xTF32 arithmetic and the
 * same grouped accumulation as nmrf_token_gemm, so a K = 9 * 256 convolution is as accurate as an fp32 one.
 * X is addressed by strides (floats): img_stride between samples, row_stride between rows, pix_stride between pixels; the Cin
 * floats of a tap are contiguous, Cin % 32 == 0, all strides multiples of 4.  (The 7x7 / 3-channel stem runs on the same
 * kernel over a zero-bordered 4-channel image: one "tap" per kernel row = 8 pixels x 4 floats, see nmrf_image_prep.)
 * Wt_hi / Wt_lo: nmrf_pack_weight_tiles of the weight reshaped to [Cout, kh*kw*Cin] (tap-major).  Y: [N, Ho, Wo, Cout]
 * contiguous; Ho / Wo = 0 derive them from the usual formula.  Cout % 16 == 0. */
typedef struct {
  const float* X; int N; int H; int W;
  int64_t img_stride; int row_stride; int pix_stride;
  int Cin; int kh; int kw; int stride; int pad;
  const float* Wt_hi; const float* Wt_lo;
  const float* bias;                           /* [Cout] or NULL */
  float* Y; int Cout; int Ho; int Wo;
} nmrf_conv_args;
int nmrf_conv2d(const nmrf_conv_args* c, void* stream);

/* ---- N2 (next row of the scope table): element-wise glue of the convolutional feature extractor, NHWC ----
 * replaces nn.InstanceNorm2d + ReLU + residual add between the convolutions of the reference's Backbone /
 * conv heads (nmrf/models/backbone.py:13-45, NMRF.py:56-65, DPN.py:45-49).  All tensors NHWC ([N, H*W, C], C % 4 == 0).
 *   stats [N, C, 2] doubles, ZEROED by the caller: sum and sum of squares over H*W (biased variance, eps 1e-5).
 *   apply:  y = IN(x) if x_stats else x;  relu if relu_inner;  y += (IN(r) if r_stats else r) if r;  relu if relu_outer;  out = y
 */
int nmrf_instnorm_stats(const float* x, int N, int HW, int C, double* stats, void* stream);
int nmrf_instnorm_apply(const float* x, const double* x_stats, const float* r, const double* r_stats,
                        int N, int HW, int C, int relu_inner, int relu_outer, float* out, void* stream);
/* N3 prologue in one pass: left / right images [B,3,H,W], values 0..255, in ANY layout (element strides of the batch,
 * channel, row and column dimensions: NCHW = (3HW, HW, W, 1), channels_last = (3HW, 1, 3W, 3)) -> out [2B, Hp+6, Wp+8, 4]:
 * 2 (x / 255) - 1 (nmrf/models/backbone.py:86) of the image replicate-padded on the right / bottom to Hp x Wp (InputPadder
 * mode 'proposal', nmrf/utils/frame_utils.py:264-275), stored as RGB0 pixels inside a ZERO border of 3 pixels (left / top)
 * and 5 / 3 pixels (right / bottom) -- the zero padding of the 7x7 stride-2 stem convolution, materialised so that the stem is
 * one nmrf_conv2d with kh = 7, kw = 1, Cin = 32 (8 pixels x 4 floats per kernel row, 16-byte aligned), stride 2. */
int nmrf_image_prep(const float* img1, const float* img2, int B, int H, int W, int Hp, int Wp,
                    int64_t stride_b, int64_t stride_c, int64_t stride_y, int64_t stride_x, float* out_rgb0, void* stream);
/* 2x2 average pool of an NHWC map [N,h,w,C] (backbone.py:96-98; h, w even are used, odd tails dropped like avg_pool2d):
 * the result split by sample half (out_a: samples 0..N/2-1 = left images, out_b: the rest) and, optionally (out_all != NULL),
 * once more as one [N,h/2,w/2,C] tensor (the operand of the conv heads at 1/8 resolution) */
int nmrf_avgpool2_split(const float* x, int N, int h, int w, int C, float* out_a, float* out_b, float* out_all, void* stream);

/* ---- N4 (next row): what every pipeline of the reference does with the disparity ------------------------------------
 * Disparity metrics of DispEvaluator.process (nmrf/utils/evaluation.py:345-359), reduced on the device: per image b,
 *   acc[b][0] = number of valid pixels (gt < max_disp, and valid_gt != 0 when valid_gt is given: `only_valid`),
 *   acc[b][1] = sum of |pr - gt| over them,   acc[b][2] = number of D1 outliers  (e > 3) & (e / gt > 0.05),
 *   acc[b][3 + i] = number of pixels with e > thresholds[i]  ("bad tau").
 * acc: [B, 3 + n_thres] doubles, ZEROED by the caller (the kernel accumulates: calls over batches may share it);
 * thresholds is a HOST array (n_thres <= 8).  The reference's per-image means are acc[b][k] / acc[b][0]; an image without
 * valid pixels is skipped by the reference (NaN check, evaluation.py:349-350). */
int nmrf_disp_metrics(const float* disp_pr, const float* disp_gt, const uint8_t* valid_gt /* or NULL */, int B, int64_t HW,
                      float max_disp, const float* thresholds_host, int n_thres, double* acc, void* stream);
/* KITTI 16-bit disparity encoding of writeDispKITTI (nmrf/utils/frame_utils.py:237-239): out = uint16(round(disp * 256)),
 * round half to even like np.round; n values. */
int nmrf_disp_to_kitti_u16(const float* disp, int64_t n, uint16_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NMRF_B200_H */
