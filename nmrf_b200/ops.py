"""Tensor-level wrappers of the C entry points (PyTorch tensors in, PyTorch tensors out).

One function per `nmrf_*` symbol of include/nmrf_b200.h; each checks device/dtype/contiguity,
allocates the outputs and launches on the current stream.  `hotpath.HotPathPlan` bypasses these
(pre-built ctypes argument lists); they exist for users who want a single operator, and for the
stage-level parity tests.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, MlpArgs, SeedWeights, lib


def _chk(t, name, dtype=torch.float32):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (nmrf_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} tensor has to be contiguous")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """run `fn` with the device of its first CUDA tensor argument current: libnmrf_b200 launches on the current device
    and configures its kernels per device, so an operator called on `cuda:1` tensors must not launch on `cuda:0`"""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def _p(t):
    return None if t is None else t.data_ptr()


@_on_device
def split_tf32(W):
    """(hi, lo) with hi = rna_tf32(W), lo = rna_tf32(W - hi), elementwise (any shape)"""
    _chk(W, "W")
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    _lib.check(lib.nmrf_split_tf32(W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), _stream()), "split_tf32")
    return hi, lo


@_on_device
def pack_weight_tiles(W):
    """(hi_tiles, lo_tiles): the weight as SWIZZLE_128B tile images for the TMA bulk copies (see nmrf_b200.h)"""
    _chk(W, "W")
    N, K = W.shape
    ntile = ((N + 127) // 128) * ((K + 31) // 32)
    hi, lo = (torch.empty(ntile * 4096, device=W.device) for _ in range(2))
    _lib.check(lib.nmrf_pack_weight_tiles(W.data_ptr(), N, K, hi.data_ptr(), lo.data_ptr(), _stream()), "pack_weight_tiles")
    return hi, lo


@_on_device
def row_stats(X):
    """(mean, rstd) of every row of X [rows, 128] (LayerNorm statistics, eps 1e-5) -> [rows, 2]"""
    _chk(X, "X")
    st = torch.empty(X.shape[0], 2, device=X.device)
    _lib.check(lib.nmrf_row_stats(X.data_ptr(), X.stride(0), X.shape[0], st.data_ptr(), _stream()), "row_stats")
    return st


@_on_device
def token_gemm(X, W, *, E=None, ediv=1, ln=None, bias=None, R=None, act=0, out=None, Wt=None, ln_stats=None):
    """Y = act(concat(LN?(X), E[r // ediv]) @ W.T + bias) (+ R).  X [rows,Kx], E [*,Ke], W [N,>=Kx+Ke].
    With Wt = pack_weight_tiles(W) the tcgen05 3xTF32 kernel is used, otherwise the exact-fp32 FMA kernel."""
    for n, t in (("X", X), ("W", W), ("E", E), ("bias", bias), ("R", R)):
        _chk(t, n)
    rows, Kx = X.shape
    N = W.shape[0]
    Y = out if out is not None else torch.empty(rows, N, device=X.device, dtype=torch.float32)
    a = GemmArgs()
    a.X, a.ldx, a.Kx = X.data_ptr(), X.stride(0), Kx
    a.E, a.lde, a.Ke, a.ediv = _p(E), (E.stride(0) if E is not None else 0), (E.shape[1] if E is not None else 0), ediv
    a.ln_gamma, a.ln_beta = (_p(ln[0]), _p(ln[1])) if ln is not None else (None, None)
    a.ln_stats = _p(ln_stats)
    a.W, a.ldw = W.data_ptr(), W.stride(0)
    a.bias = _p(bias)
    a.R, a.ldr = _p(R), (R.stride(0) if R is not None else 0)
    a.Y, a.ldy = Y.data_ptr(), Y.stride(0)
    a.rows, a.N, a.act = rows, N, act
    a.Wt_hi, a.Wt_lo = (_p(Wt[0]), _p(Wt[1])) if Wt is not None else (None, None)
    _lib.check(lib.nmrf_token_gemm(ctypes.byref(a), _stream()), "token_gemm")
    return Y


@_on_device
def _tile_images(T):
    """T [U,128,32] fp32 (device) -> [U, 8192]: per tile the SWIZZLE_128B image of hi = rna_tf32(T) (4096 floats) followed
    by the image of lo = rna_tf32(T - hi).  Image: the 16-byte chunk c of row r sits at chunk position c ^ (r & 7)."""
    T = T.contiguous()
    hi, lo = torch.empty_like(T), torch.empty_like(T)
    _lib.check(lib.nmrf_split_tf32(T.data_ptr(), hi.data_ptr(), lo.data_ptr(), T.numel(), _stream()), "split_tf32")
    r = torch.arange(128, device=T.device)[:, None]
    src = torch.arange(8, device=T.device)[None, :] ^ (r & 7)          # out[r, p] = in[r, p ^ (r & 7)]
    img = lambda x: x.view(-1, 128, 8, 4)[:, r, src, :].reshape(-1, 4096)
    return torch.cat([img(hi), img(lo)], 1).contiguous()


def pack_mlp_stream(W1cat, Wfc1, Wfc2):
    """Weight stream of nmrf_mlp_chain (P1 units, then F1(0..15), then F2(0..15); see nmrf_b200.h): W1cat [128, K1] (K1 % 32 == 0; for a block tail
    [Wproj | I]), Wfc1 [512,128], Wfc2 [128,512] -> [K1/32 + 32, 8192] fp32."""
    for n, t in (("W1cat", W1cat), ("Wfc1", Wfc1), ("Wfc2", Wfc2)):
        _chk(t, n)
    assert W1cat.shape[0] == 128 and W1cat.shape[1] % 32 == 0 and Wfc1.shape == (512, 128) and Wfc2.shape == (128, 512)
    p1 = [W1cat[:, 32 * j:32 * j + 32] for j in range(W1cat.shape[1] // 32)]
    # F1(c): fc1 rows 32c..32c+31 against the whole K = 128, as four [32 n x 32 k] sub-images stacked along the rows
    f1 = [torch.cat([Wfc1[32 * c:32 * c + 32, 32 * kb:32 * kb + 32] for kb in range(4)], 0) for c in range(16)]
    # F2(c): the 128 fc2 rows against hidden columns 32c..32c+31
    f2 = [Wfc2[:, 32 * c:32 * c + 32] for c in range(16)]
    units = p1 + f1 + f2
    return _tile_images(torch.stack(units, 0))


@_on_device
def mlp_chain(X, wstream, bias_mid, ln, b1, bias_out, *, E=None, out=None, e_identity=False, out_stats=None):
    """Y = x1 + (fc2(GELU(fc1(LN(x1)))) + bias_out) (bias_out = b_fc2) with x1 = E + (X @ W1.T + bias_mid) (e_identity: E is the
    residual, kept in fp32 registers) or x1 = concat(X, E) @ W1cat.T + bias_mid; wstream from pack_mlp_stream(W1 or W1cat, ...).
    `out` may alias E."""
    for n, t in (("X", X), ("E", E), ("wstream", wstream), ("bias_mid", bias_mid), ("gamma", ln[0]), ("beta", ln[1]), ("b1", b1),
                 ("bias_out", bias_out)):
        _chk(t, n)
    rows, Kx = X.shape
    Ke = E.shape[1] if E is not None else 0
    n1 = (Kx if e_identity else Kx + Ke) // 32
    if wstream.shape != (n1 + 32, 8192):
        raise RuntimeError(f"wstream has shape {tuple(wstream.shape)}, expected {(n1 + 32, 8192)}")
    Y = out if out is not None else torch.empty(rows, 128, device=X.device, dtype=torch.float32)
    a = MlpArgs()
    a.X, a.ldx, a.Kx = X.data_ptr(), X.stride(0), Kx
    a.E, a.lde, a.Ke = _p(E), (E.stride(0) if E is not None else 0), Ke
    a.Wstream, a.bias_mid, a.ln_gamma, a.ln_beta = wstream.data_ptr(), bias_mid.data_ptr(), ln[0].data_ptr(), ln[1].data_ptr()
    a.b1, a.bias_out = b1.data_ptr(), bias_out.data_ptr()
    a.Y, a.ldy, a.rows, a.e_identity = Y.data_ptr(), Y.stride(0), rows, int(e_identity)
    a.out_stats = _p(out_stats)
    _lib.check(lib.nmrf_mlp_chain(ctypes.byref(a), _stream()), "mlp_chain")
    return Y


@_on_device
def cost_volume_topk(f1_nhwc, f2_nhwc, conv_w, D, K, G=4, eps=1e-3):
    """conv_w: dict w0,b0,w1,b1,w2,b2 (dpn.mlp.{0,2,4}).  -> cost_volume [P,G,D], prob [P,D], seeds [P,K] int64"""
    _chk(f1_nhwc, "f1"); _chk(f2_nhwc, "f2")
    for k, t in conv_w.items():
        _chk(t, k)
    B, h, w, C = f1_nhwc.shape
    P = B * h * w
    dev = f1_nhwc.device
    cv = torch.empty(P, G, D, device=dev)
    prob = torch.empty(P, D, device=dev)
    seeds = torch.empty(P, K, device=dev, dtype=torch.int64)
    sw = SeedWeights(*[conv_w[k].data_ptr() for k in ("w0", "b0", "w1", "b1", "w2", "b2")])
    _lib.check(lib.nmrf_cost_volume_topk(f1_nhwc.data_ptr(), f2_nhwc.data_ptr(), B, h, w, C, G, D, K, eps,
                                         ctypes.byref(sw), cv.data_ptr(), prob.data_ptr(), seeds.data_ptr(), _stream()),
               "cost_volume_topk")
    return cv, prob, seeds


@_on_device
def prop_gather(cost_volume, seeds, normalizer=3.14 / 64, ld_cost=48, extended=True):
    """extended: Fourier encoding of the (integer) seeds evaluated in double, else in fp32 like the reference"""
    _chk(cost_volume, "cost_volume"); _chk(seeds, "seeds", torch.int64)
    P, G, D = cost_volume.shape
    K = seeds.shape[1]
    cost = torch.empty(P * K, ld_cost, device=seeds.device)
    enc = torch.empty(P * K, 32, device=seeds.device)
    _lib.check(lib.nmrf_prop_gather(cost_volume.data_ptr(), seeds.data_ptr(), P, G, D, K, normalizer, int(extended),
                                    cost.data_ptr(), ld_cost, enc.data_ptr(), _stream()), "prop_gather")
    return cost, enc


@_on_device
def stripe_attention(qkv, B, h, w, K, get_v0, get_v1):
    _chk(qkv, "qkv"); _chk(get_v0, "get_v0"); _chk(get_v1, "get_v1")
    out = torch.empty(qkv.shape[0], 128, device=qkv.device)
    _lib.check(lib.nmrf_stripe_attention(qkv.data_ptr(), B, h, w, K, get_v0.data_ptr(), get_v1.data_ptr(),
                                         out.data_ptr(), _stream()), "stripe_attention")
    return out


@_on_device
def prop_head_tail(hidden, w, b, seeds, extended=False):
    """labels = relu(hidden . w + b + seed).  extended: returns (hi, lo) with label = hi + lo summed in double"""
    _chk(hidden, "hidden"); _chk(w, "w"); _chk(b, "b"); _chk(seeds, "seeds", torch.int64)
    T = hidden.shape[0]
    labels = torch.empty(T, device=hidden.device)
    lo = torch.empty(T, device=hidden.device) if extended else None
    _lib.check(lib.nmrf_prop_head_tail(hidden.data_ptr(), w.data_ptr(), b.data_ptr(), seeds.data_ptr(), T,
                                       labels.data_ptr(), _p(lo), _stream()), "prop_head_tail")
    return (labels, lo) if extended else labels


@_on_device
def warp_corr_embed(f1_cc, f2_cc, f1_gw, f2_gw, labels, K, Hp, Wp, top, left, normalizer, labels_lo=None):
    """NHWC maps [B,h,w,64|256]; labels [B*h*w,K] (+ optional low words) -> feat [B*Hp*Wp*K,160], enc [.,32] on the padded grid"""
    for n, t in (("f1_cc", f1_cc), ("f2_cc", f2_cc), ("f1_gw", f1_gw), ("f2_gw", f2_gw), ("labels", labels), ("labels_lo", labels_lo)):
        _chk(t, n)
    B, h, w, _ = f1_cc.shape
    Tp = B * Hp * Wp * K
    feat = torch.empty(Tp, 160, device=labels.device)
    enc = torch.empty(Tp, 32, device=labels.device)
    _lib.check(lib.nmrf_warp_corr_embed(f1_cc.data_ptr(), f2_cc.data_ptr(), f1_gw.data_ptr(), f2_gw.data_ptr(),
                                        labels.data_ptr(), _p(labels_lo), B, h, w, K, Hp, Wp, top, left, normalizer, feat.data_ptr(),
                                        enc.data_ptr(), _stream()), "warp_corr_embed")
    return feat, enc


@_on_device
def zero_pad_rows(x, B, h, w, K, Hp, Wp, top, left):
    _chk(x, "x")
    _lib.check(lib.nmrf_zero_pad_rows(x.data_ptr(), B, h, w, K, Hp, Wp, top, left, _stream()), "zero_pad_rows")
    return x


@_on_device
def proposal_attention(qkv, K):
    _chk(qkv, "qkv")
    out = torch.empty(qkv.shape[0], 128, device=qkv.device)
    _lib.check(lib.nmrf_proposal_attention(qkv.data_ptr(), qkv.shape[0] // K, K, out.data_ptr(), _stream()),
               "proposal_attention")
    return out


@_on_device
def window_attention(qkv, table, B, Hp, Wp, K, ws, shift, self_edge_mask):
    _chk(qkv, "qkv"); _chk(table, "table")
    out = torch.empty(qkv.shape[0], 128, device=qkv.device)
    _lib.check(lib.nmrf_window_attention(qkv.data_ptr(), table.data_ptr(), B, Hp, Wp, K, ws, shift,
                                         1 if self_edge_mask else 0, out.data_ptr(), _stream()), "window_attention")
    return out


@_on_device
def select_median(delta, score, labels, B, h, w, K, Hp, Wp, top, left, labels_lo=None):
    """-> disp_curr [B,2h,2w]; with labels_lo: (disp_curr, disp_curr_lo), the selection evaluated in double"""
    _chk(delta, "delta"); _chk(score, "score"); _chk(labels, "labels"); _chk(labels_lo, "labels_lo")
    out = torch.empty(B, 2 * h, 2 * w, device=delta.device)
    out_lo = torch.empty_like(out) if labels_lo is not None else None
    _lib.check(lib.nmrf_select_median(delta.data_ptr(), score.data_ptr(), labels.data_ptr(), _p(labels_lo), B, h, w, K, Hp, Wp,
                                      top, left, out.data_ptr(), _p(out_lo), _stream()), "select_median")
    return (out, out_lo) if labels_lo is not None else out


@_on_device
def refine_tail(delta, disp_curr, Hp4, Wp4, top, left, H, W, disp_curr_lo=None):
    _chk(delta, "delta"); _chk(disp_curr, "disp_curr"); _chk(disp_curr_lo, "disp_curr_lo")
    B, h4, w4 = disp_curr.shape
    disp_pred = torch.empty(B, 4 * h4, 4 * w4, device=delta.device)
    disp = torch.empty(B, H, W, device=delta.device)
    _lib.check(lib.nmrf_refine_tail(delta.data_ptr(), disp_curr.data_ptr(), _p(disp_curr_lo), B, h4, w4, Hp4, Wp4, top, left, H, W,
                                    disp_pred.data_ptr(), disp.data_ptr(), _stream()), "refine_tail")
    return disp_pred, disp
