"""Fused NHWC execution of the reference's ResNet-style feature extractor and conv heads.

Same math as `backbone.Backbone` + the three conv3x3 -> InstanceNorm -> ReLU -> conv1x1 heads (reference
nmrf/models/backbone.py:13-98, NMRF.py:56-65, DPN.py:45-49), organised for the GPU instead of for autograd -- and every
launch is a libnmrf_b200 kernel (no cuDNN, no ATen on this path):

* activations stay NHWC end to end (the hot path consumes NHWC): no NCHW<->NHWC copies around InstanceNorm;
* every k x k convolution is ONE `nmrf_conv2d`: an implicit GEMM on the tcgen05 token-GEMM kernel (error-compensated
  3xTF32, partial sums of at most four k-blocks per tensor-memory accumulator, combined in fp32 registers), reading the
  plain fp32 activation in place -- round 1 ran cuDNN on channel-concatenated [x_hi | x_lo | x_hi] copies, three times
  the activation traffic and, because a TF32 implicit GEMM accumulates K = 27 C products in place with round-toward-zero
  updates, 8x the error of an fp32 convolution (features 5.4e-6 rms from float64 against 6.9e-7 for the reference's fp32
  arithmetic) -- enough to flip argmax / median decisions downstream;
* everything between two convolutions (InstanceNorm, ReLU, residual add -- optionally through the shortcut's own
  InstanceNorm) is one `nmrf_instnorm_stats` + one `nmrf_instnorm_apply`;
* replicate padding, `2 x / 255 - 1`, left/right batching and the stem's zero border are one `nmrf_image_prep`; the 7x7
  stride-2 stem is the same convolution kernel over that 4-channel image (one tap per kernel row);
* the 1x1 convolutions without a norm after them (backbone.conv2, the heads' projections) are token GEMMs written straight
  into the hot path's input buffers; the heads' 3x3 convolutions of a scale share one call.

Biases of convolutions that feed an InstanceNorm are dropped: the norm subtracts them again (the reference adds and
removes them; the difference is below fp32 rounding of the normalised value).
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvArgs, GemmArgs, lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _tiles(w2d):
    """hi / lo tile images of an [N, K] weight (nmrf_pack_weight_tiles)"""
    w2d = w2d.detach().float().contiguous()
    N, K = w2d.shape
    ntile = ((N + 127) // 128) * ((K + 31) // 32)
    thi, tlo = (torch.empty(ntile * 4096, device=w2d.device) for _ in range(2))
    _lib.check(lib.nmrf_pack_weight_tiles(w2d.data_ptr(), N, K, thi.data_ptr(), tlo.data_ptr(), _stream()), "pack_weight_tiles")
    return thi, tlo, w2d


class _Conv:
    """pre-packed k x k convolution (nmrf_conv2d) over an NHWC activation [N, H, W, Cin]"""

    def __init__(self, weight, stride, pad):
        Cout, Cin, kh, kw = weight.shape
        assert Cin % 32 == 0 and Cout % 16 == 0, (Cin, Cout)
        self.Cout, self.Cin, self.kh, self.kw, self.stride, self.pad = Cout, Cin, kh, kw, stride, pad
        # tap-major: k = (ky * kw + kx) * Cin + c
        self.thi, self.tlo, self._keep = _tiles(weight.detach().float().permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin))

    def out_hw(self, H, W):
        return (H + 2 * self.pad - self.kh) // self.stride + 1, (W + 2 * self.pad - self.kw) // self.stride + 1

    def __call__(self, x, y):
        """x [N, H, W, Cin] contiguous -> y [N, Ho, Wo, Cout] (pre-allocated)"""
        N, H, W, C = x.shape
        a = ConvArgs()
        a.X, a.N, a.H, a.W = x.data_ptr(), N, H, W
        a.img_stride, a.row_stride, a.pix_stride = H * W * C, W * C, C
        a.Cin, a.kh, a.kw, a.stride, a.pad = self.Cin, self.kh, self.kw, self.stride, self.pad
        a.Wt_hi, a.Wt_lo, a.bias = self.thi.data_ptr(), self.tlo.data_ptr(), None
        a.Y, a.Cout, a.Ho, a.Wo = y.data_ptr(), self.Cout, 0, 0
        _lib.check(lib.nmrf_conv2d(ctypes.byref(a), _stream()), "conv2d")
        return y


class _Stem:
    """the 7x7 stride-2 pad-3 convolution of 3 channels (backbone.py:52) as nmrf_conv2d over nmrf_image_prep's zero-bordered
    RGB0 image [N, Hp+6, Wp+8, 4]: kernel row ky = one tap of 8 pixels x 4 floats (pixel 7 and channel 3 carry zero weights)"""

    def __init__(self, weight):
        Cout = weight.shape[0]
        w = weight.detach().float().new_zeros(Cout, 7, 8, 4)
        w[:, :, :7, :3] = weight.detach().float().permute(0, 2, 3, 1)             # [Cout, ky, kx, c]
        self.Cout = Cout
        self.thi, self.tlo, self._keep = _tiles(w.reshape(Cout, 7 * 32))

    def __call__(self, img, Hp, Wp, y):
        N = img.shape[0]
        Hb, Wb = Hp + 6, Wp + 8
        a = ConvArgs()
        a.X, a.N, a.H, a.W = img.data_ptr(), N, Hb, Wb
        a.img_stride, a.row_stride, a.pix_stride = Hb * Wb * 4, Wb * 4, 4
        a.Cin, a.kh, a.kw, a.stride, a.pad = 32, 7, 1, 2, 0
        a.Wt_hi, a.Wt_lo, a.bias = self.thi.data_ptr(), self.tlo.data_ptr(), None
        a.Y, a.Cout, a.Ho, a.Wo = y.data_ptr(), self.Cout, Hp // 2, Wp // 2
        _lib.check(lib.nmrf_conv2d(ctypes.byref(a), _stream()), "conv2d(stem)")
        return y


class _Gemm:
    """pre-packed 1x1 convolution as a token GEMM (tcgen05 3xTF32)"""

    def __init__(self, w2d, bias=None):
        self.N, self.K = w2d.shape
        self.thi, self.tlo, self._keep = _tiles(w2d)
        self.bias = bias.detach().float().contiguous() if bias is not None else None

    def __call__(self, x_ptr, ldx, rows, y_ptr, ldy):
        a = GemmArgs()
        a.X, a.ldx, a.Kx = x_ptr, ldx, self.K
        a.E, a.lde, a.Ke, a.ediv = None, 0, 0, 1
        a.ln_gamma, a.ln_beta, a.ln_stats = None, None, None
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
        self.stem = _Stem(bb.conv1.weight)
        self.blocks = []
        for layer in (bb.layer1, bb.layer2, bb.layer3):
            for blk in layer:
                self.blocks.append(dict(
                    c1=_Conv(blk.conv1.weight, blk.conv1.stride[0], 1), c2=_Conv(blk.conv2.weight, 1, 1),
                    cout=blk.conv1.out_channels,
                    cd=_Conv(blk.downsample[0].weight, blk.downsample[0].stride[0], 0) if blk.downsample is not None else None))
        self.out = _Gemm(bb.conv2.weight[:, :, 0, 0], bb.conv2.bias)
        heads8 = (model.concatconv, model.gw, model.dpn.proj)
        heads4 = (model.concatconv, model.gw)
        self.head3x3 = {8: _Conv(torch.cat([h[0].weight for h in heads8], 0), 1, 1),
                        4: _Conv(torch.cat([h[0].weight for h in heads4], 0), 1, 1)}
        self.head1x1 = [_Gemm(h[3].weight[:, :, 0, 0]) for h in heads8]      # concatconv, gw, proj
        self.feat_dim = bb.conv2.out_channels
        self._ws = {}

    # ---- kernels ---------------------------------------------------------------------------------
    @staticmethod
    def _stats(y, st):
        N, H, W, C = y.shape
        _lib.check(lib.nmrf_instnorm_stats(y.data_ptr(), N, H * W, C, st.data_ptr(), _stream()), "instnorm_stats")

    @staticmethod
    def _apply(y, st, r=None, rst=None, relu_inner=True, relu_outer=False, plain=None):
        N, H, W, C = y.shape
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(lib.nmrf_instnorm_apply(y.data_ptr(), p(st), p(r), p(rst), N, H * W, C, int(relu_inner), int(relu_outer),
                                           p(plain), _stream()), "instnorm_apply")

    def _workspace(self, N, H, W, dev):
        key = (N, H, W, str(dev))
        if key not in self._ws:
            new = lambda *s: torch.empty(*s, device=dev)
            h2, w2 = H // 2, W // 2
            h4, w4 = h2 // 2, w2 // 2
            h8, w8 = h4 // 2, w4 // 2
            n_norm = 1 + sum(3 if b["cd"] is not None else 2 for b in self.blocks) + 2
            big = N * h2 * w2 * 64                            # the largest activation (1/2 resolution, 64 channels)
            self._ws[key] = dict(
                stats=torch.zeros(n_norm, N, 384, 2, dtype=torch.float64, device=dev),
                img=new(N, H + 6, W + 8, 4),
                # raw convolution outputs (y1 / y2 / shortcut) and two ping-pong activations, all carved from flat slabs
                raw=[new(max(big, N * h4 * w4 * 384)) for _ in range(3)], act=[new(big) for _ in range(3)],
                feat4=new(N, h4, w4, self.feat_dim), feat8=new(N, h8, w8, self.feat_dim),
                hd4=new(N, h4, w4, 256), hd8=new(N, h8, w8, 384), dims=(h2, w2, h4, w4, h8, w8))
        return self._ws[key]

    # ---- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, img1, img2, plan, padded=None):
        """img1 / img2: [B,3,H,W] (any strides; both the same), UNPADDED; padded = (Hp, Wp) of InputPadder (default: no pad)"""
        B, _, Hi, Wi = img1.shape
        H, W = padded if padded is not None else (Hi, Wi)
        N = 2 * B
        if img2.stride() != img1.stride():
            img2 = img2.contiguous(); img1 = img1.contiguous()
        ws = self._workspace(N, H, W, img1.device)
        h2, w2, h4, w4, h8, w8 = ws["dims"]
        stats = ws["stats"]
        stats.zero_()
        si = iter(range(stats.shape[0]))

        def st(C):
            # a [N, C, 2] view must be contiguous for the kernels: carve it from the flat per-layer slab
            i = next(si)
            return stats[i].reshape(-1)[: N * C * 2].view(N, C, 2)

        view = lambda buf, h, w, c: buf[: N * h * w * c].view(N, h, w, c)
        raw, act = ws["raw"], ws["act"]
        # replicate padding (frame_utils.py:273-275), backbone.py:86 normalisation, left/right batching and the stem's zero
        # border in one pass, straight from the caller's layout (NCHW or channels_last: strides)
        sb, sc, sy, sx = img1.stride()
        _lib.check(lib.nmrf_image_prep(img1.data_ptr(), img2.data_ptr(), B, Hi, Wi, H, W, sb, sc, sy, sx, ws["img"].data_ptr(),
                                       _stream()), "image_prep")
        y = self.stem(ws["img"], H, W, view(raw[0], h2, w2, 64))
        s = st(64); self._stats(y, s)
        cur = 0
        x = view(act[cur], h2, w2, 64)
        self._apply(y, s, plain=x)
        hh, ww = h2, w2
        for blk in self.blocks:
            cout = blk["cout"]
            ho, wo = blk["c1"].out_hw(hh, ww)
            y1 = blk["c1"](x, view(raw[0], ho, wo, cout))
            s1 = st(cout); self._stats(y1, s1)
            inner = view(act[2], ho, wo, cout)
            self._apply(y1, s1, plain=inner)
            y2 = blk["c2"](inner, view(raw[1], ho, wo, cout))
            s2 = st(cout); self._stats(y2, s2)
            nxt = view(act[1 - cur], ho, wo, cout)
            if blk["cd"] is not None:
                z = blk["cd"](x, view(raw[2], ho, wo, cout))
                sz = st(cout); self._stats(z, sz)
                self._apply(y2, s2, r=z, rst=sz, relu_outer=True, plain=nxt)
            else:
                self._apply(y2, s2, r=x, relu_outer=True, plain=nxt)
            x, cur, hh, ww = nxt, 1 - cur, ho, wo
        # 1x1 output convolution (with bias) -> feat @1/4, NHWC; feat @1/8 = avg_pool2 (backbone.py:96-98)
        rows4 = N * h4 * w4
        self.out(x.data_ptr(), x.shape[-1], rows4, ws["feat4"].data_ptr(), self.feat_dim)
        _lib.check(lib.nmrf_avgpool2_split(ws["feat4"].data_ptr(), N, h4, w4, self.feat_dim, plan.f1_8.data_ptr(), plan.f2_8.data_ptr(),
                                           ws["feat8"].data_ptr(), _stream()), "avgpool2_split")
        # heads: one 3x3 convolution for all heads of a scale, InstanceNorm + ReLU, then the 1x1 projections as GEMMs
        for scale, feat, hd, hh, ww, cc, gw in ((8, ws["feat8"], ws["hd8"], h8, w8, plan.cc8, plan.gw8),
                                                (4, ws["feat4"], ws["hd4"], h4, w4, plan.cc4, plan.gw4)):
            conv = self.head3x3[scale]
            y = conv(feat, view(raw[0], hh, ww, conv.Cout))
            C = conv.Cout
            s = st(C); self._stats(y, s)
            self._apply(y, s, plain=hd)
            half = B * hh * ww
            for img in range(2):
                base = hd.data_ptr() + img * half * C * 4
                self.head1x1[0](base, C, half, cc[img].data_ptr(), 64)                 # concatconv
                self.head1x1[1](base + 128 * 4, C, half, gw[img].data_ptr(), 256)      # gw
            if scale == 8:
                self.head1x1[2](hd.data_ptr() + 256 * 4, C, half, plan.context.data_ptr(), 64)   # dpn.proj, left image
