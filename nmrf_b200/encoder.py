"""Fused NHWC execution of the reference's ResNet-style feature extractor and conv heads.

Same math as `backbone.Backbone` + the three conv3x3 -> InstanceNorm -> ReLU -> conv1x1 heads (reference
nmrf/models/backbone.py:13-98, NMRF.py:56-65, DPN.py:45-49), organised for the GPU instead of for autograd:

* activations stay NHWC end to end (the hot path consumes NHWC): no NCHW<->NHWC copies around InstanceNorm;
* every k x k convolution is ONE cuDNN call on the channel-concatenated operand [x_hi | x_lo | x_hi] against
  [w_hi | w_hi | w_lo] -- error-compensated 3xTF32 with the three products accumulated inside the kernel;
* everything between two convolutions (InstanceNorm, ReLU, residual add -- optionally through the shortcut's own
  InstanceNorm --, the hi/lo split of the next operand) is one `nmrf_instnorm_stats` + one `nmrf_instnorm_apply`;
* the 1x1 convolutions without a norm after them (backbone.conv2, the heads' projections) are token GEMMs on the
  tcgen05 kernel, written straight into the hot path's input buffers; the heads' 3x3 convolutions share one call.

Biases of convolutions that feed an InstanceNorm are dropped: the norm subtracts them again (the reference adds and
removes them; the difference is below fp32 rounding of the normalised value).
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import GemmArgs, lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _split(t):
    t = t.contiguous()
    hi, lo = torch.empty_like(t), torch.empty_like(t)
    _lib.check(lib.nmrf_split_tf32(t.data_ptr(), hi.data_ptr(), lo.data_ptr(), t.numel(), _stream()), "split_tf32")
    return hi, lo


def _cat3_weight(w):
    """[Cout, Cin, kh, kw] -> [Cout, 3 Cin, kh, kw] = [w_hi | w_hi | w_lo], channels_last"""
    hi, lo = _split(w.detach().float())
    return torch.cat((hi, hi, lo), 1).contiguous(memory_format=torch.channels_last)


class _Gemm:
    """pre-packed 1x1 convolution as a token GEMM (tcgen05 3xTF32)"""

    def __init__(self, w2d, bias=None):
        w2d = w2d.detach().float().contiguous()
        self.N, self.K = w2d.shape
        ntile = ((self.N + 127) // 128) * ((self.K + 31) // 32)
        self.thi, self.tlo = (torch.empty(ntile * 4096, device=w2d.device) for _ in range(2))
        _lib.check(lib.nmrf_pack_weight_tiles(w2d.data_ptr(), self.N, self.K, self.thi.data_ptr(), self.tlo.data_ptr(),
                                              _stream()), "pack_weight_tiles")
        self.bias = bias.detach().float().contiguous() if bias is not None else None
        self._keep = w2d

    def __call__(self, x_ptr, ldx, rows, y_ptr, ldy):
        a = GemmArgs()
        a.X, a.ldx, a.Kx = x_ptr, ldx, self.K
        a.E, a.lde, a.Ke, a.ediv = None, 0, 0, 1
        a.ln_gamma, a.ln_beta = None, None
        a.W, a.ldw = self._keep.data_ptr(), self._keep.stride(0)
        a.bias = self.bias.data_ptr() if self.bias is not None else None
        a.R, a.ldr = None, 0
        a.Y, a.ldy = y_ptr, ldy
        a.rows, a.N, a.act = rows, self.N, 0
        a.Wt_hi, a.Wt_lo = self.thi.data_ptr(), self.tlo.data_ptr()
        _lib.check(lib.nmrf_token_gemm(ctypes.byref(a), _stream()), "token_gemm")


class FusedEncoder:
    """`FusedEncoder(model).run(img1, img2, plan)` fills the hot path's inputs (f1_8, f2_8, context, cc8, gw8, cc4, gw4)."""

    def __init__(self, model):
        bb = model.backbone
        self.w = {}
        cw = lambda conv: _cat3_weight(conv.weight)
        self.stem = cw(bb.conv1)
        self.blocks = []
        for layer in (bb.layer1, bb.layer2, bb.layer3):
            for blk in layer:
                self.blocks.append(dict(
                    w1=cw(blk.conv1), w2=cw(blk.conv2), stride=blk.conv1.stride[0], cout=blk.conv1.out_channels,
                    wd=cw(blk.downsample[0]) if blk.downsample is not None else None))
        self.out = _Gemm(bb.conv2.weight[:, :, 0, 0], bb.conv2.bias)
        heads8 = (model.concatconv, model.gw, model.dpn.proj)
        heads4 = (model.concatconv, model.gw)
        self.head3x3 = {8: _cat3_weight(torch.cat([h[0].weight for h in heads8], 0)),
                        4: _cat3_weight(torch.cat([h[0].weight for h in heads4], 0))}
        self.head1x1 = [_Gemm(h[3].weight[:, :, 0, 0]) for h in heads8]      # concatconv, gw, proj
        self.feat_dim = bb.conv2.out_channels
        self._ws = {}
        # cuDNN algorithm choice for the k x k convolutions: heuristics (default, deterministic) or autotuned.  Autotuning is
        # only switched on through `autotune()`, which VERIFIES that the tuned algorithms reproduce the heuristic ones'
        # features (a Winograd/FFT pick would not keep the [hi|lo|hi] products exact)
        self.cudnn_autotune = False
        self.autotune_report = None

    # ---- kernels ---------------------------------------------------------------------------------
    def _conv(self, cat3_nhwc, w, stride, pad):
        x = cat3_nhwc.permute(0, 3, 1, 2)                   # NCHW view of an NHWC buffer == channels_last
        with torch.backends.cudnn.flags(enabled=True, benchmark=self.cudnn_autotune, allow_tf32=True):
            y = F.conv2d(x, w, None, stride, pad)
        if not y.is_contiguous(memory_format=torch.channels_last):
            y = y.contiguous(memory_format=torch.channels_last)
        return y.permute(0, 2, 3, 1)                        # NHWC, contiguous

    @staticmethod
    def _stats(y, st):
        N, H, W, C = y.shape
        _lib.check(lib.nmrf_instnorm_stats(y.data_ptr(), N, H * W, C, st.data_ptr(), _stream()), "instnorm_stats")

    @staticmethod
    def _apply(y, st, r=None, rst=None, relu_inner=True, relu_outer=False, plain=None, cat3=None):
        N, H, W, C = y.shape
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(lib.nmrf_instnorm_apply(y.data_ptr(), p(st), p(r), p(rst), N, H * W, C, int(relu_inner), int(relu_outer),
                                           p(plain), p(cat3), _stream()), "instnorm_apply")

    def _workspace(self, N, H, W, dev):
        key = (N, H, W, str(dev))
        if key not in self._ws:
            new = lambda *s: torch.empty(*s, device=dev)
            h2, w2 = (H + 1) // 2, (W + 1) // 2
            h4, w4 = (h2 + 1) // 2, (w2 + 1) // 2
            h8, w8 = h4 // 2, w4 // 2
            n_norm = 1 + sum(3 if b["wd"] is not None else 2 for b in self.blocks) + 2
            self._ws[key] = dict(
                stats=torch.zeros(n_norm, N, 384, 2, dtype=torch.float64, device=dev),
                img=new(N, H, W, 9),
                # two ping-pong sets of (plain, cat3) per resolution, plus the inner cat3 of a block
                p2=[new(N, h2, w2, 64) for _ in range(2)], c2=[new(N, h2, w2, 192) for _ in range(3)],
                p4=[new(N, h4, w4, 128) for _ in range(2)], c4=[new(N, h4, w4, 384) for _ in range(3)],
                feat4=new(N, h4, w4, self.feat_dim), f4c=new(N, h4, w4, 3 * self.feat_dim),
                f8c=new(N, h8, w8, 3 * self.feat_dim),
                hd4=new(N, h4, w4, 256), hd8=new(N, h8, w8, 384), dims=(h2, w2, h4, w4, h8, w8))
        return self._ws[key]

    @torch.no_grad()
    def autotune(self, img1, img2, plan, tol=2e-6):
        """Let cuDNN benchmark its algorithms for this shape and keep them only if every tensor the hot path consumes agrees
        with the heuristic algorithms' result to `tol` (relative to the tensor's max; fp32 accumulation order is all that may
        differ).  Returns the report dict (also in self.autotune_report)."""
        outs = lambda: [t.clone() for t in (plan.f1_8, plan.f2_8, plan.context, *plan.cc8, *plan.gw8, *plan.cc4, *plan.gw4)]
        self.cudnn_autotune = False
        self.run(img1, img2, plan)
        ref = outs()
        self.cudnn_autotune = True
        self.run(img1, img2, plan)                             # first call benchmarks and caches the algorithms
        self.run(img1, img2, plan)
        worst = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)) for a, b in zip(outs(), ref))
        ok = worst <= tol
        self.cudnn_autotune = ok
        self.autotune_report = {"enabled": ok, "max_rel_diff_vs_heuristic": worst, "tol": tol}
        return self.autotune_report

    # ---- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, img1, img2, plan):
        B, _, H, W = img1.shape
        N = 2 * B
        ws = self._workspace(N, H, W, img1.device)
        h2, w2, h4, w4, h8, w8 = ws["dims"]
        stats = ws["stats"]
        stats.zero_()
        si = iter(range(stats.shape[0]))
        def st(C):
            # a [N, C, 2] view must be contiguous for the kernels: carve it from the flat per-layer slab
            i = next(si)
            return stats[i].reshape(-1)[: N * C * 2].view(N, C, 2)

        # backbone.py:86 normalisation, left/right batching and the hi/lo split of the stem operand in one pass; the images are
        # channels_last, i.e. already NHWC in memory
        i1 = img1.permute(0, 2, 3, 1).contiguous()
        i2 = img2.permute(0, 2, 3, 1).contiguous()
        _lib.check(lib.nmrf_image_prep(i1.data_ptr(), i2.data_ptr(), B, H, W, ws["img"].data_ptr(), _stream()), "image_prep")
        y = self._conv(ws["img"], self.stem, 2, 3)
        s = st(64); self._stats(y, s)
        P, C3 = ws["p2"], ws["c2"]
        cur = 0
        self._apply(y, s, plain=P[cur].view(N, h2, w2, -1), cat3=C3[cur])
        plain, cat3 = P[cur], C3[cur]
        for bi, blk in enumerate(self.blocks):
            cout = blk["cout"]
            if bi == 2:                                      # layer2 onwards lives at 1/4 resolution
                P, C3, cur = ws["p4"], ws["c4"], 1
            res_in = plain
            view = lambda buf, c: buf.reshape(-1)[: N * (h4 if bi >= 2 else h2) * (w4 if bi >= 2 else w2) * c].view(
                N, (h4 if bi >= 2 else h2), (w4 if bi >= 2 else w2), c)
            y1 = self._conv(cat3, blk["w1"], blk["stride"], 1)
            s1 = st(cout); self._stats(y1, s1)
            inner = view(C3[2], 3 * cout)
            self._apply(y1, s1, cat3=inner)
            y2 = self._conv(inner, blk["w2"], 1, 1)
            s2 = st(cout); self._stats(y2, s2)
            nxt_i = 1 - cur
            out_plain, out_cat3 = view(P[nxt_i], cout), view(C3[nxt_i], 3 * cout)
            if blk["wd"] is not None:
                z = self._conv(cat3, blk["wd"], blk["stride"], 0)
                sz = st(cout); self._stats(z, sz)
                self._apply(y2, s2, r=z, rst=sz, relu_outer=True, plain=out_plain, cat3=out_cat3)
            else:
                self._apply(y2, s2, r=res_in, relu_outer=True, plain=out_plain, cat3=out_cat3)
            plain, cat3, cur = out_plain, out_cat3, nxt_i
        # 1x1 output convolution (with bias) -> feat @1/4, NHWC; feat @1/8 = avg_pool2 (backbone.py:96-98)
        rows4 = N * h4 * w4
        self.out(plain.data_ptr(), plain.shape[-1], rows4, ws["feat4"].data_ptr(), self.feat_dim)
        _lib.check(lib.nmrf_avgpool2_split(ws["feat4"].data_ptr(), N, h4, w4, self.feat_dim, plan.f1_8.data_ptr(), plan.f2_8.data_ptr(),
                                           ws["f8c"].data_ptr(), _stream()), "avgpool2_split")
        # heads: one 3x3 convolution for all heads of a scale, InstanceNorm + ReLU, then the 1x1 projections as GEMMs
        for scale, feat, f3, hd, hh, ww, cc, gw in ((8, None, ws["f8c"], ws["hd8"], h8, w8, plan.cc8, plan.gw8),
                                                    (4, ws["feat4"], ws["f4c"], ws["hd4"], h4, w4, plan.cc4, plan.gw4)):
            rows = N * hh * ww
            if feat is not None:                             # the 1/8 operand was written by the pooling kernel
                _lib.check(lib.nmrf_split_cat3(feat.data_ptr(), rows, self.feat_dim, f3.data_ptr(), _stream()), "split_cat3")
            y = self._conv(f3, self.head3x3[scale], 1, 1)
            C = y.shape[-1]
            s = st(C); self._stats(y, s)
            self._apply(y, s, plain=hd)
            half = B * hh * ww
            for img in range(2):
                base = hd.data_ptr() + img * half * C * 4
                self.head1x1[0](base, C, half, cc[img].data_ptr(), 64)                 # concatconv
                self.head1x1[1](base + 128 * 4, C, half, gw[img].data_ptr(), 256)      # gw
            if scale == 8:
                self.head1x1[2](hd.data_ptr() + 256 * 4, C, half, plan.context.data_ptr(), 64)   # dpn.proj, left image
