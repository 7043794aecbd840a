"""Host-side driver of the CUDA hot path: weight packing, workspaces and the launch plan.

Everything here is plumbing: torch is used for device memory only.  All arithmetic of the
path happens in libnmrf_b200.so (see include/nmrf_b200.h); there is no PyTorch fallback.

A `HotPathPlan` is built once per (batch, padded image size) and holds
  * pre-allocated workspaces (so a forward allocates nothing and is CUDA-graph capturable),
  * a flat list of (C function, ctypes args) launches.
Stage order follows NMRF.forward (nmrf/models/NMRF.py:207-245 of the reference).
"""
import ctypes
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import GemmArgs, MlpArgs, SeedWeights, lib

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


def _use_tc_default():
    import os
    return os.environ.get("NMRF_B200_GEMM", "tc").lower() != "simt"


@dataclass
class HotPathConfig:
    """The hyper-parameters the kernels need (nmrf/config/default.py:37-61)."""
    max_disp: int = 320
    num_proposals: int = 4
    cost_group: int = 4
    window_size: int = 6
    refine_window_size: int = 4
    num_prop_layers: int = 5
    num_infer_layers: int = 5
    num_refine_layers: int = 5
    eps: float = 1e-3                      # DPN.py:51
    tensor_cores: bool = True              # tcgen05 3xTF32 GEMMs (False / NMRF_B200_GEMM=simt: exact-fp32 FMA kernel)
    # labels / disp_curr carried between stages as hi + lo fp32 pairs summed in double (include/nmrf_b200.h, "Extended
    # labels"); False reproduces the reference's plain fp32 label arithmetic
    extended_labels: bool = True


def _extended_labels_default():
    import os
    return os.environ.get("NMRF_B200_EXT_LABELS", "1") != "0"


def _use_mlp_chain_default():
    import os
    return os.environ.get("NMRF_B200_MLP", "1") != "0"


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def _pad_cols(w, to):
    """zero-pad the input dimension of an nn.Linear weight [out,in] to `to` columns"""
    out, k = w.shape
    if k == to:
        return _f32(w)
    p = w.new_zeros(out, to)
    p[:, :k] = w
    return _f32(p)


def center_pad(n, ws):
    """NMP.py:747-754: (padded size, leading pad)"""
    pad = (ws - n % ws) % ws
    return n + pad, pad // 2


class PackedWeights:
    """Kernel-ready views/copies of the reference state-dict tensors (Appendix C of SURVEY.md).

    Packing is layout only: q/k/v projections sharing an input are stacked along `out`, input
    dimensions are zero-padded to a multiple of 8 (159 -> 160, 36 -> 48)."""

    def __init__(self, sd, cfg: HotPathConfig):
        g = lambda k: _f32(sd[k])
        self.seed = {k: g(f"dpn.mlp.{i}.{n}") for k, (i, n) in
                     dict(w0=(0, "weight"), b0=(0, "bias"), w1=(2, "weight"), b1=(2, "bias"),
                          w2=(4, "weight"), b2=(4, "bias")).items()}
        p = "dpn.propagation"
        self.ce0_w, self.ce0_b = _pad_cols(sd[p + ".cost_encoder.0.weight"], 48), g(p + ".cost_encoder.0.bias")
        self.ce2_w, self.ce2_b = g(p + ".cost_encoder.2.weight"), g(p + ".cost_encoder.2.bias")
        self.pproj_w = _pad_cols(sd[p + ".proj.weight"], 160)
        self.prop_layers = []
        for i in range(cfg.num_prop_layers):
            q = f"{p}.layers.{i}.nmp"
            wv = _pad_cols(sd[q + ".v.weight"], 192)
            L = dict(
                qkv_w=_f32(torch.cat([sd[q + ".q.weight"], sd[q + ".k.weight"], wv], 0)),
                qkv_b=_f32(torch.cat([sd[q + ".q.bias"], sd[q + ".k.bias"], sd[q + ".v.bias"]], 0)),
                n1=(g(q + ".norm1.weight"), g(q + ".norm1.bias")),
                n2=(g(q + ".norm2.weight"), g(q + ".norm2.bias")),
                proj_w=g(q + ".proj.weight"), proj_b=g(q + ".proj.bias"),
                gv0=g(q + ".attns.0.get_v.weight"), gv1=g(q + ".attns.1.get_v.weight"),
                fc1_w=g(q + ".mlp.fc1.weight"), fc1_b=g(q + ".mlp.fc1.bias"),
                fc2_w=g(q + ".mlp.fc2.weight"), fc2_b=g(q + ".mlp.fc2.bias"))
            self.prop_layers.append(L)
        self.prop_norm = (g(p + ".norm.weight"), g(p + ".norm.bias"))
        self.prop_head = [(g(f"dpn.prop_head.layers.{i}.weight"), g(f"dpn.prop_head.layers.{i}.bias")) for i in range(3)]

        self.stacks = {}
        for name, n_layers, with_self in (("inference", cfg.num_infer_layers, True),
                                          ("refinement", cfg.num_refine_layers, False)):
            S = dict(ffn1_w=g(name + ".ffn.fc1.weight"), ffn1_b=g(name + ".ffn.fc1.bias"),
                     ffn2_w=g(name + ".ffn.fc2.weight"), ffn2_b=g(name + ".ffn.fc2.bias"),
                     norm=(g(name + ".norm.weight"), g(name + ".norm.bias")), layers=[])
            for i in range(n_layers):
                q = f"{name}.layers.{i}"
                L = {}
                if with_self:
                    s = q + ".self_nmp"
                    L.update(
                        s_qkv_w=_f32(torch.cat([_pad_cols(sd[s + ".q.weight"], 160), _pad_cols(sd[s + ".k.weight"], 160),
                                                _pad_cols(sd[s + ".v.weight"], 160)], 0)),
                        s_qkv_b=_f32(torch.cat([sd[s + ".q.bias"], sd[s + ".k.bias"], sd[s + ".v.bias"]], 0)),
                        s_n1=(g(s + ".norm1.weight"), g(s + ".norm1.bias")),
                        s_proj_w=g(s + ".proj.weight"), s_proj_b=g(s + ".proj.bias"))
                m = q + ".nmp"
                L.update(
                    qkv_w=_pad_cols(sd[m + ".qkv.weight"], 160), qkv_b=g(m + ".qkv.bias"),
                    n1=(g(m + ".norm1.weight"), g(m + ".norm1.bias")),
                    n2=(g(m + ".norm2.weight"), g(m + ".norm2.bias")),
                    table=g(m + ".attn.relative_position_enc_table"),
                    proj_w=g(m + ".proj.weight"), proj_b=g(m + ".proj.bias"),
                    fc1_w=g(m + ".mlp.fc1.weight"), fc1_b=g(m + ".mlp.fc1.bias"),
                    fc2_w=g(m + ".mlp.fc2.weight"), fc2_b=g(m + ".mlp.fc2.bias"))
                S["layers"].append(L)
            self.stacks[name] = S
        self.infer_head = [(g(f"infer_head.layers.{i}.weight"), g(f"infer_head.layers.{i}.bias")) for i in range(3)]
        self.score_head = (g("infer_score_head.weight"), g("infer_score_head.bias"))
        self.refine_head = [(g(f"refine_head.layers.{i}.weight"), g(f"refine_head.layers.{i}.bias")) for i in range(3)]


class _Launches:
    """Flat launch list; keeps every ctypes struct and tensor it references alive."""

    def __init__(self):
        self.calls = []
        self.cost = []          # algorithmic (flops, HBM bytes) per launch, for the roofline report
        self._keep = []
        self._splits = {}
        self._streams = {}
        self.tensor_cores = False
        self.mlp_chain = False
        import os
        self.residual_preload = os.environ.get("NMRF_B200_RESIDUAL", "preload").lower() != "identity"

    def add(self, fn, what, *args, flops=0.0, bytes=0.0):
        self.calls.append((fn, what, args))
        self.cost.append((flops, bytes))

    def keep(self, *objs):
        self._keep.extend(objs)

    def tiles(self, W):
        """hi / lo tile images (nmrf_pack_weight_tiles) of an [N, K] weight for the TMA bulk copies; cached per tensor"""
        key = W.data_ptr()
        if key not in self._splits:
            N, K = W.shape
            ntile = ((N + 127) // 128) * ((K + 31) // 32)
            thi, tlo = (torch.empty(ntile * 4096, device=W.device) for _ in range(2))
            Wc = W.contiguous()
            st = torch.cuda.current_stream(W.device).cuda_stream
            _lib.check(lib.nmrf_pack_weight_tiles(Wc.data_ptr(), N, K, thi.data_ptr(), tlo.data_ptr(), st), "pack_weight_tiles")
            self._splits[key] = (thi, tlo, W, Wc)
        return self._splits[key][:2]

    def gemm(self, what, X, W, Y, rows, N, *, Kx=None, E=None, Ke=0, ediv=1, ln=None, bias=None, R=None, act=ACT_NONE, stats=None):
        Wt_hi = Wt_lo = None
        if self.tensor_cores and N % 16 == 0 and N <= 512 and W.is_cuda:
            Wt_hi, Wt_lo = self.tiles(W)
        a = GemmArgs()
        a.X, a.ldx, a.Kx = X.data_ptr(), X.stride(0), (Kx if Kx is not None else X.shape[1])
        a.E, a.lde, a.Ke, a.ediv = (E.data_ptr() if E is not None else None), (E.stride(0) if E is not None else 0), Ke, ediv
        a.ln_gamma, a.ln_beta = (ln[0].data_ptr(), ln[1].data_ptr()) if ln is not None else (None, None)
        a.ln_stats = stats.data_ptr() if (stats is not None and ln is not None) else None
        a.W, a.ldw = W.data_ptr(), W.stride(0)
        a.bias = bias.data_ptr() if bias is not None else None
        a.R, a.ldr = (R.data_ptr(), R.stride(0)) if R is not None else (None, 0)
        a.Y, a.ldy = Y.data_ptr(), Y.stride(0)
        a.rows, a.N, a.act = rows, N, act
        a.Wt_hi = Wt_hi.data_ptr() if Wt_hi is not None else None
        a.Wt_lo = Wt_lo.data_ptr() if Wt_lo is not None else None
        self.keep(a, X, W, Wt_hi, Wt_lo, Y, E, ln, bias, R, stats)
        # algorithmic work: 2*MAC flops; activations read+written once (weights are L2-resident, excluded)
        flops = 2.0 * rows * N * (a.Kx + Ke)
        nbytes = 4.0 * (rows * a.Kx + (rows // max(ediv, 1)) * Ke + rows * N * (2 if R is not None else 1))
        self.add(lib.nmrf_token_gemm, what, ctypes.byref(a), flops=flops, bytes=nbytes)

    def row_stats(self, what, X, rows, stats):
        """(mean, rstd) of every row of X for the LayerNorm prologue of the GEMMs that read it (X written by a token GEMM; the
        fused block tail writes the statistics of its output itself)"""
        self.keep(X, stats)
        self.add(lib.nmrf_row_stats, what, X.data_ptr(), X.stride(0), rows, stats.data_ptr(), bytes=4.0 * rows * 130)

    def block_tail(self, what, att, x, rows, proj_w, proj_b, n2, fc1_w, fc1_b, fc2_w, fc2_b, stats=None):
        """x = x1 + Mlp(LN2(x1)), x1 = x + proj(att): ONE launch of nmrf_mlp_chain (SwinNMP / CSWinNMP tail, NMP.py:358-363,
        570-573).  The residual stream x stays in fp32 registers inside the kernel (gemm_mlp.cu: never in a tensor-core
        accumulator, whose round-toward-zero updates would bias it); NMRF_B200_RESIDUAL=identity (experiment) lets it ride the
        tensor core as an identity block appended to the proj weight instead."""
        from . import ops
        key = (proj_w.data_ptr(), fc1_w.data_ptr(), fc2_w.data_ptr())
        if key not in self._streams:
            if self.residual_preload:
                w1 = proj_w.contiguous()
            else:
                w1 = torch.cat([proj_w, torch.eye(128, device=proj_w.device, dtype=torch.float32)], 1).contiguous()
            ws = ops.pack_mlp_stream(w1, fc1_w.contiguous(), fc2_w.contiguous())
            self._streams[key] = (ws, fc2_b.contiguous())
        ws, bias_out = self._streams[key]
        a = MlpArgs()
        a.X, a.ldx, a.Kx = att.data_ptr(), att.stride(0), 128
        a.E, a.lde, a.Ke = x.data_ptr(), x.stride(0), 128
        a.Wstream, a.bias_mid, a.ln_gamma, a.ln_beta = ws.data_ptr(), proj_b.data_ptr(), n2[0].data_ptr(), n2[1].data_ptr()
        a.b1, a.bias_out = fc1_b.data_ptr(), bias_out.data_ptr()
        a.Y, a.ldy, a.rows, a.e_identity = x.data_ptr(), x.stride(0), rows, int(self.residual_preload)
        a.out_stats = stats.data_ptr() if stats is not None else None
        self.keep(a, att, x, ws, bias_out, proj_b, n2, fc1_b, stats)
        flops = 2.0 * rows * (128 * 128 + 2 * 128 * 512)
        self.add(lib.nmrf_mlp_chain, what, ctypes.byref(a), flops=flops, bytes=4.0 * rows * 128 * 3)

    def run(self, stream):
        for fn, what, args in self.calls:
            rc = fn(*args, stream)
            if rc != 0:
                _lib.check(rc, what)

    def run_timed(self, reps=3):
        """eager run with a CUDA-event pair around every launch (on the launching stream; call with the plan's device
        current).  Returns [(what, symbol, ms, flops, bytes)] with ms = best of `reps`."""
        cur = torch.cuda.current_stream()
        stream = cur.cuda_stream
        best = [float("inf")] * len(self.calls)
        for _ in range(reps):
            evs = []
            for fn, what, args in self.calls:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(cur)
                rc = fn(*args, stream)
                e.record(cur)
                if rc != 0:
                    _lib.check(rc, what)
                evs.append((s, e))
            torch.cuda.synchronize()
            best = [min(b, s.elapsed_time(e)) for b, (s, e) in zip(best, evs)]
        return [(what, fn.__name__, ms, c[0], c[1]) for (fn, what, _), ms, c in zip(self.calls, best, self.cost)]


class HotPathPlan:
    """Launch plan for one shape.  B pairs, 1/8 grid h8 x w8, 1/4 grid h4 x w4, output H x W."""

    def __init__(self, pw: PackedWeights, cfg: HotPathConfig, B, C, h8, w8, H, W, device):
        assert h8 * 8 >= H and w8 * 8 >= W
        self.cfg, self.B, self.C, self.h8, self.w8, self.H, self.W = cfg, B, C, h8, w8, H, W
        self.device = torch.device(device)
        K, G, D = cfg.num_proposals, cfg.cost_group, cfg.max_disp // 8
        h4, w4 = 2 * h8, 2 * w8
        P8, P4 = B * h8 * w8, B * h4 * w4
        T8 = P8 * K
        ws, rws = cfg.window_size, cfg.refine_window_size
        Hp8, top8 = center_pad(h8, ws)
        Wp8, left8 = center_pad(w8, ws)
        Hp4, top4 = center_pad(h4, rws)
        Wp4, left4 = center_pad(w4, rws)
        T8p, T4p = B * Hp8 * Wp8 * K, B * Hp4 * Wp4
        Tmax = max(T8, T8p, T4p)
        self.geom = dict(K=K, G=G, D=D, h4=h4, w4=w4, P8=P8, P4=P4, T8=T8, T8p=T8p, T4p=T4p,
                         Hp8=Hp8, Wp8=Wp8, top8=top8, left8=left8, Hp4=Hp4, Wp4=Wp4, top4=top4, left4=left4)
        new = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=device)
        # ---- inputs (NHWC, filled by the caller before run()) ------------------------------------
        self.f1_8, self.f2_8 = new(B, h8, w8, C), new(B, h8, w8, C)
        self.context = new(B, h8, w8, 64)
        self.cc8 = [new(B, h8, w8, 64), new(B, h8, w8, 64)]
        self.gw8 = [new(B, h8, w8, 256), new(B, h8, w8, 256)]
        self.cc4 = [new(B, h4, w4, 64), new(B, h4, w4, 64)]
        self.gw4 = [new(B, h4, w4, 256), new(B, h4, w4, 256)]
        # ---- outputs ----------------------------------------------------------------------------
        self.cost_volume, self.prob = new(P8, G, D), new(P8, D)
        self.seeds = new(P8, K, dtype=torch.int64)
        self.labels = new(P8, K)
        self.disp_curr = new(B, h4, w4)
        ext = bool(cfg.extended_labels) and _extended_labels_default()
        self.labels_lo = new(P8, K) if ext else None           # label = labels + labels_lo (summed in double by the kernels)
        self.disp_curr_lo = new(B, h4, w4) if ext else None
        self.disp_pred, self.disp = new(B, 4 * h4, 4 * w4), new(B, H, W)
        # ---- workspaces -------------------------------------------------------------------------
        self.x, self.att, self.h1, self.h2 = new(Tmax, 128), new(Tmax, 128), new(Tmax, 128), new(Tmax, 128)
        self.qkv, self.hid = new(Tmax, 384), new(Tmax, 512)
        self.feat, self.enc = new(Tmax, 160), new(Tmax, 32)
        self.xstats = new(Tmax, 2)                              # (mean, rstd) of every row of x: the next LayerNorm's statistics
        self.cost48 = new(T8, 48)
        self.delta, self.score = new(Tmax, 64), new(Tmax, 64)
        self.pw = pw
        self.launches = _Launches()
        self.launches.tensor_cores = bool(cfg.tensor_cores) and _use_tc_default()
        self.launches.mlp_chain = self.launches.tensor_cores and _use_mlp_chain_default()
        with torch.cuda.device(self.device):               # weight packing launches kernels: on the plan's device
            self._build(pw)

    # -------------------------------------------------------------------------------------------
    def _mlp_block(self, L_, T, n2, fc1_w, fc1_b, fc2_w, fc2_b, tag):
        L_.gemm(tag + ".fc1", self.x, fc1_w, self.hid, T, 512, ln=n2, bias=fc1_b, act=ACT_GELU)
        L_.gemm(tag + ".fc2", self.hid, fc2_w, self.x, T, 128, bias=fc2_b, R=self.x)

    def _block_tail(self, L_, T, w, tag):
        """proj + residual, then the Mlp block (NMP.py:358-363 / 570-573): one fused launch, or three token GEMMs"""
        if L_.mlp_chain:
            L_.block_tail(tag + ".tail", self.att, self.x, T, w["proj_w"], w["proj_b"], w["n2"], w["fc1_w"], w["fc1_b"],
                          w["fc2_w"], w["fc2_b"], stats=self.xstats)
        else:
            L_.gemm(tag + ".proj", self.att, w["proj_w"], self.x, T, 128, bias=w["proj_b"], R=self.x)
            self._mlp_block(L_, T, w["n2"], w["fc1_w"], w["fc1_b"], w["fc2_w"], w["fc2_b"], tag)
            L_.row_stats(tag + ".stats", self.x, T, self.xstats)

    def _build(self, pw):
        c, g, L_ = self.cfg, self.geom, self.launches
        B, C, h8, w8 = self.B, self.C, self.h8, self.w8
        K, G, D = g["K"], g["G"], g["D"]
        P8, T8 = g["P8"], g["T8"]
        ptr = lambda t: t.data_ptr()

        # A1+A2 -------------------------------------------------------------------------------------
        sw = SeedWeights(*[ptr(pw.seed[k]) for k in ("w0", "b0", "w1", "b1", "w2", "b2")])
        L_.keep(sw)
        L_.add(lib.nmrf_cost_volume_topk, "cost_volume_topk", ptr(self.f1_8), ptr(self.f2_8), B, h8, w8, C, G, D, K,
               c.eps, ctypes.byref(sw), ptr(self.cost_volume), ptr(self.prob), ptr(self.seeds),
               # algorithmic bytes (SURVEY.md §8(d), S1 + S2 fused): both feature maps in, cost volume + prob + int64 seeds out
               bytes=4.0 * (2 * P8 * C + P8 * G * D + P8 * D) + 8.0 * P8 * K,
               flops=2.0 * P8 * D * C + 2.0 * P8 * D * 5 * (G * 8 + 8 * 16 + 16))
        # A3+A4 -------------------------------------------------------------------------------------
        ext = int(self.labels_lo is not None)
        optr = lambda t: None if t is None else t.data_ptr()
        L_.add(lib.nmrf_prop_gather, "prop_gather", ptr(self.cost_volume), ptr(self.seeds), P8, G, D, K, 3.14 / 64, ext,
               ptr(self.cost48), 48, ptr(self.enc), bytes=4.0 * (P8 * G * D + T8 * (48 + 32)) + 8.0 * T8)
        L_.gemm("cost_encoder.0", self.cost48, pw.ce0_w, self.h1, T8, 128, bias=pw.ce0_b, act=ACT_GELU)
        L_.gemm("cost_encoder.2", self.h1, pw.ce2_w, self.h2, T8, 128, bias=pw.ce2_b)
        L_.gemm("propagation.proj", self.h2, pw.pproj_w, self.x, T8, 128, E=self.enc, Ke=32)
        L_.row_stats("propagation.stats", self.x, T8, self.xstats)
        # A5+A6 -------------------------------------------------------------------------------------
        ctx = self.context.view(P8, 64)
        for i, w in enumerate(pw.prop_layers):
            t = f"prop{i}"
            L_.gemm(t + ".qkv", self.x, w["qkv_w"], self.qkv, T8, 384, E=ctx, Ke=64, ediv=K, ln=w["n1"], bias=w["qkv_b"],
                    stats=self.xstats)
            # stripe attention (A6): 2 heads x (q.k^T + p.v) x 32 dims over full-height and full-width stripes; qkv in, att out
            L_.add(lib.nmrf_stripe_attention, t + ".stripe", ptr(self.qkv), B, h8, w8, K, ptr(w["gv0"]), ptr(w["gv1"]),
                   ptr(self.att), bytes=4.0 * T8 * (384 + 128),
                   flops=2.0 * B * (w8 * 2 * (h8 * K) ** 2 * 32 * 2 + h8 * 2 * (w8 * K) ** 2 * 32 * 2))
            self._block_tail(L_, T8, w, t)
        # A7 ----------------------------------------------------------------------------------------
        (w0, b0), (w1, b1), (w2, b2) = pw.prop_head
        L_.gemm("prop_head.0", self.x, w0, self.h1, T8, 128, ln=pw.prop_norm, bias=b0, act=ACT_RELU, stats=self.xstats)
        L_.gemm("prop_head.1", self.h1, w1, self.h2, T8, 128, bias=b1, act=ACT_RELU)
        L_.add(lib.nmrf_prop_head_tail, "prop_head.2", ptr(self.h2), ptr(w2), ptr(b2), ptr(self.seeds), T8, ptr(self.labels),
               optr(self.labels_lo), bytes=4.0 * T8 * 128 + 8.0 * T8 + 8.0 * T8)

        # A8-A12: inference @1/8 --------------------------------------------------------------------
        self._stack(L_, pw.stacks["inference"], "inference", self.labels, self.labels_lo, self.cc8, self.gw8, h8, w8, K,
                    g["Hp8"], g["Wp8"], g["top8"], g["left8"], c.window_size, 3.14 / 64, True)
        T8p = g["T8p"]
        (w0, b0), (w1, b1), (w2, b2) = pw.infer_head
        nrm = pw.stacks["inference"]["norm"]
        L_.gemm("infer_head.0", self.x, w0, self.h1, T8p, 128, ln=nrm, bias=b0, act=ACT_RELU, stats=self.xstats)
        L_.gemm("infer_head.1", self.h1, w1, self.h2, T8p, 128, bias=b1, act=ACT_RELU)
        L_.gemm("infer_head.2", self.h2, w2, self.delta, T8p, 64, bias=b2)
        # 0.25 * score (NMRF.py:220) does not change the argmax: the exact power-of-two scale is dropped
        L_.gemm("infer_score_head", self.x, pw.score_head[0], self.score, T8p, 64, ln=nrm, bias=pw.score_head[1], stats=self.xstats)
        L_.add(lib.nmrf_select_median, "select_median", ptr(self.delta), ptr(self.score), ptr(self.labels), optr(self.labels_lo),
               B, h8, w8, K, g["Hp8"], g["Wp8"], g["top8"], g["left8"], ptr(self.disp_curr), optr(self.disp_curr_lo),
               bytes=4.0 * (2 * T8 * 64 + 2 * T8 + 2 * g["P4"]))

        # A13: refinement @1/4 ----------------------------------------------------------------------
        h4, w4 = g["h4"], g["w4"]
        self._stack(L_, pw.stacks["refinement"], "refinement", self.disp_curr, self.disp_curr_lo, self.cc4, self.gw4, h4, w4, 1,
                    g["Hp4"], g["Wp4"], g["top4"], g["left4"], c.refine_window_size, 3.14 / 128, False)
        T4p = g["T4p"]
        (w0, b0), (w1, b1), (w2, b2) = pw.refine_head
        nrm = pw.stacks["refinement"]["norm"]
        L_.gemm("refine_head.0", self.x, w0, self.h1, T4p, 128, ln=nrm, bias=b0, act=ACT_RELU, stats=self.xstats)
        L_.gemm("refine_head.1", self.h1, w1, self.h2, T4p, 128, bias=b1, act=ACT_RELU)
        delta16 = self.delta.view(-1)[:T4p * 16].view(T4p, 16)      # dense [T4p,16] as nmrf_refine_tail expects
        L_.gemm("refine_head.2", self.h2, w2, delta16, T4p, 16, bias=b2)
        L_.add(lib.nmrf_refine_tail, "refine_tail", ptr(delta16), ptr(self.disp_curr), optr(self.disp_curr_lo), B, h4, w4, g["Hp4"], g["Wp4"],
               g["top4"], g["left4"], self.H, self.W, ptr(self.disp_pred), ptr(self.disp),
               bytes=4.0 * (g["P4"] * 16 + 2 * g["P4"] + 2 * B * 16 * h4 * w4))

    def _stack(self, L_, S, name, labels, labels_lo, cc, gw, h, w, K, Hp, Wp, top, left, ws, normalizer, with_self):
        B = self.B
        ptr = lambda t: t.data_ptr()
        Tp = B * Hp * Wp * K
        L_.add(lib.nmrf_warp_corr_embed, name + ".embed", ptr(cc[0]), ptr(cc[1]), ptr(gw[0]), ptr(gw[1]), ptr(labels),
               None if labels_lo is None else ptr(labels_lo), B, h, w, K, Hp, Wp, top, left, normalizer, ptr(self.feat),
               ptr(self.enc),
               # S6 / S9: the four feature maps (64 + 64 + 256 + 256 channels) once, labels, feat [Tp,160] + enc [Tp,32] out
               bytes=4.0 * (640 * B * h * w + 2 * B * h * w * K + Tp * 192))
        L_.gemm(name + ".ffn.fc1", self.feat, S["ffn1_w"], self.h1, Tp, 128, bias=S["ffn1_b"], act=ACT_GELU)
        L_.gemm(name + ".ffn.fc2", self.h1, S["ffn2_w"], self.x, Tp, 128, bias=S["ffn2_b"])
        if Hp != h or Wp != w:
            L_.add(lib.nmrf_zero_pad_rows, name + ".zero_pad", ptr(self.x), B, h, w, K, Hp, Wp, top, left,
                   bytes=4.0 * 128 * (Tp - B * h * w * K))
        L_.row_stats(name + ".stats", self.x, Tp, self.xstats)
        for i, wt in enumerate(S["layers"]):
            t = f"{name}{i}"
            shift = 0 if i % 2 == 0 else ws // 2                     # NMRF.py:72,96
            if with_self:
                L_.gemm(t + ".self.qkv", self.x, wt["s_qkv_w"], self.qkv, Tp, 384, E=self.enc, Ke=32, ln=wt["s_n1"],
                        bias=wt["s_qkv_b"], stats=self.xstats)
                L_.add(lib.nmrf_proposal_attention, t + ".self.attn", ptr(self.qkv), Tp // K, K, ptr(self.att),
                       bytes=4.0 * Tp * (384 + 128), flops=2.0 * (Tp // K) * 4 * K * K * 32 * 2)
                L_.gemm(t + ".self.proj", self.att, wt["s_proj_w"], self.x, Tp, 128, bias=wt["s_proj_b"], R=self.x)
                L_.row_stats(t + ".self.stats", self.x, Tp, self.xstats)
            L_.gemm(t + ".qkv", self.x, wt["qkv_w"], self.qkv, Tp, 384, E=self.enc, Ke=32, ln=wt["n1"], bias=wt["qkv_b"],
                    stats=self.xstats)
            # window attention (A11): 4 heads x 5 contractions of (ws^2 K)^2 x 32 per window (q.k, q.Rk, k.Rq, A.v, A.Rv; H4)
            L_.add(lib.nmrf_window_attention, t + ".window", ptr(self.qkv), ptr(wt["table"]), B, Hp, Wp, K, ws, shift,
                   1 if with_self else 0, ptr(self.att), bytes=4.0 * Tp * (384 + 128),
                   flops=2.0 * (B * (Hp // ws) * (Wp // ws)) * 4 * (ws * ws * K) ** 2 * 32 * 5)
            self._block_tail(L_, Tp, wt, t)

    # -------------------------------------------------------------------------------------------
    def run(self):
        """Launch the whole hot path on the current stream of the plan's device (inputs already copied in).  The device is made
        current for the launches: the library configures kernels per device and launches on the current one."""
        with torch.cuda.device(self.device):
            self.launches.run(torch.cuda.current_stream(self.device).cuda_stream)

    def run_with_taps(self):
        """Eager run that clones the stage-boundary tensors (same names as the oracle's taps).  Debug /
        parity tooling only: allocates."""
        with torch.cuda.device(self.device):
            return self._run_with_taps()

    def _run_with_taps(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        g, taps = self.geom, {}
        K = g["K"]
        for fn, what, args in self.launches.calls:
            rc = fn(*args, stream)
            if rc != 0:
                _lib.check(rc, what)
            if what == "cost_volume_topk":
                taps.update(cost_volume=self.cost_volume.clone(), prob=self.prob.clone(), seeds=self.seeds.clone())
            elif what == "propagation.proj":
                taps["prop_embed"] = self.x[:g["T8"]].reshape(-1, K, 128).clone()
            elif what.startswith("prop") and what.endswith((".fc2", ".tail")) and what[4].isdigit():
                taps[f"prop_layer{what[4:what.rindex('.')]}"] = self.x[:g["T8"]].reshape(-1, K, 128).clone()
            elif what == "prop_head.2":
                taps["labels"] = self.labels.clone()
            elif what.startswith("inference") and what.endswith((".fc2", ".tail")) and what[9].isdigit():
                taps[f"inference_layer{what[9:what.rindex('.')]}"] = self.x[:g["T8p"]].reshape(-1, K, 128).clone()
            elif what.startswith("refinement") and what.endswith((".fc2", ".tail")) and what[10].isdigit():
                taps[f"refinement_layer{what[10:what.rindex('.')]}"] = self.x[:g["T4p"]].reshape(-1, 1, 128).clone()
            elif what == "infer_score_head":
                taps.update(delta=self.delta[:g["T8p"]].clone(), score=self.score[:g["T8p"]].clone())
            elif what == "select_median":
                taps["disp_curr"] = self.disp_curr.clone()
            elif what == "refine_tail":
                taps.update(disp=self.disp.clone(), disp_pred=self.disp_pred.clone())
        return taps

    @property
    def num_launches(self):
        return len(self.launches.calls)
