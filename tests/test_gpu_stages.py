"""GPU: every C-ABI entry point against the CPU oracle on identical seeded inputs (stress weights, so
every term matters), plus the committed reference goldens.  Integer paths bit-exact (modulo
numerically undecidable near-ties, which are counted and bounded); fp32 paths to ~1e-5 relative.
"""
import numpy as np
import pytest
import torch

from helpers import T, build_product_model, golden, oracle_cfg
from oracle import nmrf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cuda(t):
    return t.to(DEV).contiguous()


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(1e-6, float(b.abs().max())))


@pytest.fixture(scope="module")
def ops():
    import nmrf_b200.ops as ops
    return ops


@pytest.fixture(scope="module")
def stress():
    _, sd = build_product_model(192, 4, (2, 2, 2), 11, "stress")
    return sd


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,Kx,Ke,ediv,N,ln,act,res", [
    (300, 128, 0, 1, 128, False, 0, False),
    (777, 128, 32, 1, 384, True, 0, False),
    (512, 128, 64, 4, 384, True, 0, False),
    (1000, 128, 0, 1, 512, True, 2, False),
    (1000, 512, 0, 1, 128, False, 0, True),
    (129, 48, 0, 1, 128, False, 2, False),
    (64, 160, 0, 1, 128, False, 2, False),
    (333, 128, 0, 1, 64, True, 1, False),
    (40, 128, 0, 1, 16, False, 0, False),
    (1, 128, 0, 1, 128, True, 1, True),
])
@pytest.mark.parametrize("tc", [False, True], ids=["fma", "tcgen05"])
def test_token_gemm(ops, rows, Kx, Ke, ediv, N, ln, act, res, tc):
    g = torch.Generator().manual_seed(rows + N)
    X = torch.randn(rows, Kx, generator=g)
    E = torch.randn((rows + ediv - 1) // ediv, Ke, generator=g) if Ke else None
    W = torch.randn(N, Kx + Ke, generator=g) / (Kx + Ke) ** 0.5
    b = torch.randn(N, generator=g)
    gam, bet = 1 + 0.1 * torch.randn(Kx, generator=g), 0.1 * torch.randn(Kx, generator=g)
    R = torch.randn(rows, N, generator=g) if res else None
    A = torch.nn.functional.layer_norm(X, (Kx,), gam, bet, 1e-5) if ln else X
    if Ke:
        A = torch.cat([A, E.repeat_interleave(ediv, 0)[:rows]], 1)
    ref = (A.double() @ W.double().T + b.double())
    ref = torch.relu(ref) if act == 1 else torch.nn.functional.gelu(ref) if act == 2 else ref
    if res:
        ref = ref + R.double()
    Wt = ops.pack_weight_tiles(cuda(W)) if tc else None
    out = ops.token_gemm(cuda(X), cuda(W), E=cuda(E) if Ke else None, ediv=ediv, Wt=Wt,
                         ln=(cuda(gam), cuda(bet)) if ln else None, bias=cuda(b), R=cuda(R) if res else None, act=act)
    assert rel_err(out, ref) <= (4e-6 if tc else 2e-6)


def test_token_gemm_tc_many_tiles_and_split_exactness(ops):
    """persistent schedule (more tiles than SMs), and the hi/lo split itself: hi+lo reproduces W to 2^-21 relative,
    hi and lo are TF32-representable (13 low mantissa bits zero)."""
    g = torch.Generator().manual_seed(77)
    W = torch.randn(384, 160, generator=g)
    hi, lo = ops.split_tf32(cuda(W))
    assert int((hi.view(torch.int32) & 0x1fff).abs().max()) == 0 and int((lo.view(torch.int32) & 0x1fff).abs().max()) == 0
    assert float(((hi + lo).cpu() - W).abs().max() / W.abs().max()) <= 2 ** -21
    rows = 128 * 400 + 37
    X = torch.randn(rows, 128, generator=g)
    E = torch.randn(rows, 32, generator=g)
    ref = torch.cat([X, E], 1).double() @ W.double().T
    out = ops.token_gemm(cuda(X), cuda(W), E=cuda(E), Wt=ops.pack_weight_tiles(cuda(W)))
    assert rel_err(out, ref) <= 4e-6
    out_fma = ops.token_gemm(cuda(X), cuda(W), E=cuda(E))           # without tile images: the exact-fp32 FMA kernel
    assert rel_err(out_fma, ref) <= 2e-6


def test_token_gemm_residual_in_place(ops):
    g = torch.Generator().manual_seed(5)
    X, W, Y = torch.randn(500, 128, generator=g), torch.randn(128, 128, generator=g) * 0.1, torch.randn(500, 128, generator=g)
    ref = Y.double() + X.double() @ W.double().T
    y = cuda(Y)
    ops.token_gemm(cuda(X), cuda(W), R=y, out=y)
    assert rel_err(y, ref) <= 2e-6


@pytest.mark.parametrize("rows,mode", [(128, "residual"), (1000, "residual"), (128 * 300 + 5, "residual"), (300, "concat"), (700, "mlp")])
def test_mlp_chain(ops, rows, mode):
    """fused block tail (proj + residual + LN2 + fc1 + GELU + fc2 + residual, NMP.py:358-363) against float64.
    residual: x preloaded into the accumulator; concat: x through an identity block of the weight; mlp: Mlp block only."""
    g = torch.Generator().manual_seed(rows)
    att, x = torch.randn(rows, 128, generator=g), 2.0 * torch.randn(rows, 128, generator=g) + 0.3
    Wp, bp = torch.randn(128, 128, generator=g) / 128 ** 0.5, 0.1 * torch.randn(128, generator=g)
    W1, b1 = torch.randn(512, 128, generator=g) / 128 ** 0.5, 0.1 * torch.randn(512, generator=g)
    W2, b2 = torch.randn(128, 512, generator=g) / 512 ** 0.5, 0.1 * torch.randn(128, generator=g)
    gam, bet = 1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g)
    if mode == "mlp":                                           # Mlp block only: an identity weight carries x into the accumulator
        x1 = x.double()
        W1cat, X, E, bmid, ident = torch.eye(128), x, None, torch.zeros(128), False
    else:
        x1 = x.double() + att.double() @ Wp.double().T + bp.double()
        W1cat, X, E, bmid, ident = (Wp, att, x, bp, True) if mode == "residual" else (torch.cat([Wp, torch.eye(128)], 1), att, x, bp, False)
    t = torch.nn.functional.layer_norm(x1, (128,), gam.double(), bet.double(), 1e-5)
    ref = x1 + torch.nn.functional.gelu(t @ W1.double().T + b1.double()) @ W2.double().T + b2.double()
    ws = ops.pack_mlp_stream(cuda(W1cat), cuda(W1), cuda(W2))
    xe = cuda(E) if E is not None else None
    args = (cuda(X), ws, cuda(bmid), (cuda(gam), cuda(bet)), cuda(b1), cuda(b2))
    out = ops.mlp_chain(*args, E=xe, e_identity=ident)
    assert rel_err(out, ref) <= 1e-5         # three chained 3xTF32 GEMMs (4e-6 each) + LN + GELU
    if mode == "residual":                                      # in place on the residual stream, as the hot path runs it
        ops.mlp_chain(*args, E=xe, out=xe, e_identity=True)
        assert rel_err(xe, ref) <= 1e-5


def test_row_stats_and_external_layernorm_statistics(ops):
    """nmrf_row_stats / nmrf_mlp_chain(out_stats) hand the next LayerNorm its (mean, rstd): the token GEMM that takes them
    (nmrf_gemm_args.ln_stats) must give what it gives when it computes the statistics itself"""
    g = torch.Generator().manual_seed(12)
    rows = 128 * 3 + 77
    X = 2.0 * torch.randn(rows, 128, generator=g) + 0.5
    st = ops.row_stats(cuda(X))
    mean, var = X.double().mean(1), X.double().var(1, unbiased=False)
    assert float((st[:, 0].cpu().double() - mean).abs().max()) <= 1e-6
    assert float((st[:, 1].cpu().double() * (var + 1e-5).sqrt() - 1).abs().max()) <= 2e-6
    W = torch.randn(384, 160, generator=g) / 160 ** 0.5
    E = torch.randn(rows, 32, generator=g)
    gam, bet = 1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g)
    kw = dict(E=cuda(E), ln=(cuda(gam), cuda(bet)), Wt=ops.pack_weight_tiles(cuda(W)))
    own = ops.token_gemm(cuda(X), cuda(W), **kw)
    ext = ops.token_gemm(cuda(X), cuda(W), ln_stats=st, **kw)
    ref = torch.cat([torch.nn.functional.layer_norm(X.double(), (128,), gam.double(), bet.double(), 1e-5), E.double()], 1) @ W.double().T
    assert rel_err(ext, ref) <= 4e-6 and rel_err(own, ref) <= 4e-6
    assert rel_err(ext, own) <= 1e-6
    # the fused block tail writes the statistics of ITS output
    att = torch.randn(rows, 128, generator=g)
    Wp, W1, W2 = (torch.randn(128, 128, generator=g) / 11.3, torch.randn(512, 128, generator=g) / 11.3, torch.randn(128, 512, generator=g) / 22.6)
    z128, z512, one = torch.zeros(128), torch.zeros(512), torch.ones(128)
    ws = ops.pack_mlp_stream(cuda(Wp), cuda(W1), cuda(W2))
    stats = torch.zeros(rows, 2, device=DEV)
    y = ops.mlp_chain(cuda(att), ws, cuda(z128), (cuda(one), cuda(z128)), cuda(z512), cuda(z128), E=cuda(X), e_identity=True, out_stats=stats)
    st2 = ops.row_stats(y)
    assert float((stats[:, 0] - st2[:, 0]).abs().max()) <= 1e-6 and float((stats[:, 1] / st2[:, 1] - 1).abs().max()) <= 2e-6


def test_token_gemm_rejects_bad_arguments(ops):
    with pytest.raises(RuntimeError, match="multiple"):
        ops.token_gemm(torch.zeros(8, 100, device=DEV), torch.zeros(128, 100, device=DEV))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.token_gemm(torch.zeros(8, 128), torch.zeros(128, 128))


# ---------------------------------------------------------------------------------------------
def _decidable(prob_nms, K, tol=1e-6):
    """rows whose top-K is numerically unambiguous: the K+1 largest values are pairwise separated by
    more than tol (relative), or tied EXACTLY (exact ties are resolved by the index rule)."""
    v, _ = torch.sort(prob_nms, dim=-1, descending=True)
    top = v[:, :K + 1]
    gap = (top[:, :-1] - top[:, 1:])
    ok = (gap == 0) | (gap > tol * top[:, :-1].abs().clamp_min(1e-12))
    return ok.all(-1)


@pytest.mark.parametrize("B,h,w,C,D,K", [(1, 7, 13, 256, 24, 4), (2, 9, 40, 256, 8, 2), (1, 5, 70, 128, 32, 4),
                                         (1, 3, 33, 256, 40, 3)])
def test_cost_volume_topk(ops, stress, B, h, w, C, D, K):
    g = torch.Generator().manual_seed(B * 100 + w)
    f1 = torch.randn(B, C, h, w, generator=g)
    f2 = torch.roll(f1, -3, -1) + 0.3 * torch.randn(B, C, h, w, generator=g)
    cv_o = O.cost_volume(f1, f2, D, 4)
    prob_o, pn_o, seeds_o = O.seed_extraction(stress, cv_o, K)
    cw = {k: cuda(stress[f"dpn.mlp.{i}.{n}"]) for k, (i, n) in dict(w0=(0, "weight"), b0=(0, "bias"), w1=(2, "weight"),
                                                                     b1=(2, "bias"), w2=(4, "weight"), b2=(4, "bias")).items()}
    cv, prob, seeds = ops.cost_volume_topk(cuda(f1.permute(0, 2, 3, 1)), cuda(f2.permute(0, 2, 3, 1)), cw, D, K)
    assert rel_err(cv, cv_o) <= 2e-6
    assert float((prob.cpu() - prob_o).abs().max()) <= 2e-6
    ok = _decidable(pn_o, K)
    assert ok.float().mean() > 0.9
    assert torch.equal(seeds.cpu()[ok], seeds_o[ok])                      # bit-exact where decidable
    # everywhere: the picked VALUES agree (a different pick can only be a numerically tied entry)
    assert float((pn_o.gather(1, seeds.cpu()) - pn_o.gather(1, seeds_o)).abs().max()) <= 1e-6


def test_cost_volume_topk_exact_ties_use_index_order(ops, stress):
    """all-zero features: flat softmax, everything suppressed to eps or tied -> seeds must be the
    canonical (value desc, index asc) pick, identical to the oracle."""
    f = torch.zeros(1, 256, 4, 20)
    cv_o = O.cost_volume(f, f, 16, 4)
    _, _, seeds_o = O.seed_extraction(stress, cv_o, 4)
    cw = {k: cuda(stress[f"dpn.mlp.{i}.{n}"]) for k, (i, n) in dict(w0=(0, "weight"), b0=(0, "bias"), w1=(2, "weight"),
                                                                     b1=(2, "bias"), w2=(4, "weight"), b2=(4, "bias")).items()}
    z = cuda(f.permute(0, 2, 3, 1))
    _, _, seeds = ops.cost_volume_topk(z, z, cw, 16, 4)
    assert torch.equal(seeds.cpu(), seeds_o)


def test_cost_volume_topk_reference_golden(ops):
    g = golden("stages")
    _, sd = build_product_model(192, 4, g["L"], int(g["weight_seed"]), "stress")
    cw = {k: cuda(sd[f"dpn.mlp.{i}.{n}"]) for k, (i, n) in dict(w0=(0, "weight"), b0=(0, "bias"), w1=(2, "weight"),
                                                                 b1=(2, "bias"), w2=(4, "weight"), b2=(4, "bias")).items()}
    cv, prob, seeds = ops.cost_volume_topk(cuda(T(g["f8a"]).permute(0, 2, 3, 1)), cuda(T(g["f8b"]).permute(0, 2, 3, 1)), cw, 24, 4)
    assert rel_err(cv, T(g["cost_volume"])) <= 5e-6
    assert float((prob.cpu() - T(g["prob"])).abs().max()) <= 5e-6
    _, pn, _ = O.seed_extraction(sd, T(g["cost_volume"]), 4)
    ok = _decidable(pn, 4)
    assert torch.equal(seeds.cpu()[ok], T(g["seeds"])[ok])


# ---------------------------------------------------------------------------------------------
def test_prop_gather(ops):
    g = torch.Generator().manual_seed(3)
    P, G, D, K = 200, 4, 24, 4
    cv = torch.randn(P, G, D, generator=g)
    seeds = torch.randint(0, D, (P, K), generator=g)
    cost, enc = ops.prop_gather(cuda(cv), cuda(seeds), extended=False)
    assert torch.equal(cost.cpu()[:, :36], O.sample_cost(cv, seeds).reshape(P * K, 36))     # pure gather: exact
    assert float(cost[:, 36:].abs().max()) == 0.0
    ref = O.fourier_embed(seeds.float().reshape(-1), 3.14 / 64)            # the reference's fp32 arithmetic
    assert float((enc.cpu()[:, :31] - ref).abs().max()) <= 1e-6
    assert float(enc[:, 31].abs().max()) == 0.0
    # extended: the encoding of the integer seed evaluated in double -> held to the float64 oracle (fp32 storage rounding only)
    cost2, enc2 = ops.prop_gather(cuda(cv), cuda(seeds), extended=True)
    assert torch.equal(cost2, cost)
    ref64 = O.fourier_embed(seeds.double().reshape(-1), 3.14 / 64)
    assert float((enc2.cpu()[:, :31].double() - ref64).abs().max()) <= 6e-8
    assert float((ref.double() - ref64).abs().max()) >= 1e-5              # ... which the fp32 evaluation misses by this much


def test_prop_head_tail(ops):
    g = torch.Generator().manual_seed(4)
    hid, w, b = torch.randn(333, 128, generator=g), torch.randn(128, generator=g), torch.randn(1, generator=g)
    seeds = torch.randint(0, 24, (333,), generator=g)
    ref = torch.relu(hid.double() @ w.double() + b.double() + seeds.double())
    out = ops.prop_head_tail(cuda(hid), cuda(w), cuda(b), cuda(seeds))
    assert rel_err(out, ref) <= 2e-6
    hi, lo = ops.prop_head_tail(cuda(hid), cuda(w), cuda(b), cuda(seeds), extended=True)
    assert torch.equal(hi, out) or rel_err(hi, ref) <= 2e-6                 # the hi word is the fp32 label
    # hi + lo carries the sum seed + (fp32 dot + bias) without rounding it to fp32 again
    s32 = (hid.double() @ w.double() + b.double()).float().double()         # what an exact fp32 dot would hold (~1e-6 off ours)
    ext = hi.cpu().double() + lo.cpu().double()
    assert float((ext - torch.relu(s32 + seeds.double())).abs().max()) <= 2e-5
    assert float(lo.abs().max()) <= 2e-6 * 24


@pytest.fixture(params=[1, 0], ids=["tcgen05", "fma"])
def attn_impl(request):
    from nmrf_b200 import _lib
    _lib.lib.nmrf_set_attention_impl(request.param)
    yield request.param
    _lib.lib.nmrf_set_attention_impl(1)


@pytest.mark.parametrize("B,h,w,K", [(1, 7, 13, 4), (2, 5, 9, 2), (1, 40, 3, 3), (1, 20, 70, 4), (1, 1, 1, 4), (1, 68, 120, 4)])
def test_stripe_attention(ops, stress, attn_impl, B, h, w, K):
    g = torch.Generator().manual_seed(h * w)
    qkv = torch.randn(B, h, w, K, 384, generator=g)
    qkv[..., :256] *= 2.0                                                   # peaky softmax
    p = "dpn.propagation.layers.0.nmp"
    q, k, v = qkv[..., :128], qkv[..., 128:256], qkv[..., 256:]
    x1 = O.stripe_attention(stress, p + ".attns.0", q[..., :64], k[..., :64], v[..., :64], True)
    x2 = O.stripe_attention(stress, p + ".attns.1", q[..., 64:], k[..., 64:], v[..., 64:], False)
    ref = torch.cat([x1, x2], -1).reshape(-1, 128)
    out = ops.stripe_attention(cuda(qkv.reshape(-1, 384)), B, h, w, K, cuda(stress[p + ".attns.0.get_v.weight"]),
                               cuda(stress[p + ".attns.1.get_v.weight"]))
    assert rel_err(out, ref) <= 1e-5


@pytest.mark.parametrize("P,K", [(100, 4), (33, 2), (17, 1), (9, 8)])
def test_proposal_attention(ops, P, K):
    g = torch.Generator().manual_seed(P)
    qkv = torch.randn(P, K, 384, generator=g) * 1.5
    hd = lambda z: z.reshape(P, K, 4, 32).permute(0, 2, 1, 3)
    q, k, v = hd(qkv[..., :128]), hd(qkv[..., 128:256]), hd(qkv[..., 256:])
    attn = torch.softmax((q.double() @ k.double().transpose(-2, -1)) * 32 ** -0.5, -1)
    ref = (attn @ v.double()).permute(0, 2, 1, 3).reshape(P * K, 128)
    out = ops.proposal_attention(cuda(qkv.reshape(-1, 384)), K)
    assert rel_err(out, ref) <= 1e-5


@pytest.mark.parametrize("B,Hp,Wp,K,ws,shift,self_edge", [
    (1, 12, 18, 4, 6, 0, True), (1, 12, 18, 4, 6, 3, True), (2, 6, 12, 2, 6, 3, True),
    (1, 8, 12, 1, 4, 0, False), (2, 8, 12, 1, 4, 2, False), (1, 16, 28, 1, 4, 2, False), (1, 6, 6, 3, 6, 3, True), (3, 18, 12, 4, 6, 3, True), (1, 36, 44, 1, 4, 2, True)])
def test_window_attention(ops, attn_impl, B, Hp, Wp, K, ws, shift, self_edge):
    g = torch.Generator().manual_seed(Hp * Wp + shift)
    qkv = torch.randn(B, Hp, Wp, K, 384, generator=g)
    table = 0.5 * torch.randn((2 * ws - 1) ** 2, 384, generator=g)
    ref = O.window_attention({"a.relative_position_enc_table": table}, "a", qkv, B, Hp, Wp, K, ws, shift, self_edge)
    out = ops.window_attention(cuda(qkv.reshape(-1, 384)), cuda(table), B, Hp, Wp, K, ws, shift, self_edge)
    assert rel_err(out, ref.reshape(-1, 128)) <= 1e-5


@pytest.mark.parametrize("B,h,w,K,ws,norm", [(1, 7, 13, 4, 6, 3.14 / 64), (2, 5, 9, 2, 6, 3.14 / 64), (1, 14, 26, 1, 4, 3.14 / 128),
                                             (1, 6, 12, 3, 6, 3.14 / 64)])
def test_warp_corr_embed(ops, B, h, w, K, ws, norm):
    from nmrf_b200.hotpath import center_pad
    g = torch.Generator().manual_seed(h + w)
    cc1, cc2 = torch.randn(B, 64, h, w, generator=g), torch.randn(B, 64, h, w, generator=g)
    gw1, gw2 = torch.randn(B, 256, h, w, generator=g), torch.randn(B, 256, h, w, generator=g)
    labels = torch.rand(B, h, w, K, generator=g) * (w + 4) - 2          # includes out-of-range and negative targets
    labels[0, 0, 0, 0] = 3.0                                            # an integral disparity (a == 0)
    Hp, top = center_pad(h, ws)
    Wp, left = center_pad(w, ws)
    nh = lambda t: cuda(t.permute(0, 2, 3, 1))
    feat, enc = ops.warp_corr_embed(nh(cc1), nh(cc2), nh(gw1), nh(gw2), cuda(labels.reshape(-1, K)), K, Hp, Wp, top, left, norm)
    wg = O.warp_sample(gw2, labels)
    corr = (gw1.permute(0, 2, 3, 1)[:, :, :, None, :] * wg).reshape(B, h, w, K, 32, 8).mean(-1)
    ref = torch.cat([cc1.permute(0, 2, 3, 1)[:, :, :, None, :].expand(B, h, w, K, 64), O.warp_sample(cc2, labels), corr], -1)
    featg = feat.cpu().reshape(B, Hp, Wp, K, 160)
    encg = enc.cpu().reshape(B, Hp, Wp, K, 32)
    inner = featg[:, top:top + h, left:left + w]
    assert rel_err(inner, ref) <= 2e-6
    ref_enc = O.fourier_embed(labels, norm)
    assert float((encg[:, top:top + h, left:left + w, :, :31] - ref_enc).abs().max()) <= 2e-6
    mask = torch.ones(B, Hp, Wp, dtype=torch.bool)
    mask[:, top:top + h, left:left + w] = False
    assert float(featg[mask].abs().max() if mask.any() else 0) == 0.0     # pad tokens are exactly zero
    assert float(encg[mask].abs().max() if mask.any() else 0) == 0.0
    x = torch.randn(B * Hp * Wp * K, 128, generator=g)
    xz = ops.zero_pad_rows(cuda(x), B, h, w, K, Hp, Wp, top, left).cpu().reshape(B, Hp, Wp, K, 128)
    assert torch.equal(xz[:, top:top + h, left:left + w], x.reshape(B, Hp, Wp, K, 128)[:, top:top + h, left:left + w])
    assert float(xz[mask].abs().max() if mask.any() else 0) == 0.0


@pytest.mark.parametrize("B,h,w,K,ws", [(1, 7, 13, 4, 6), (2, 6, 12, 2, 6), (1, 3, 5, 1, 6)])
def test_select_median(ops, stress, B, h, w, K, ws):
    from nmrf_b200.hotpath import center_pad
    g = torch.Generator().manual_seed(h * 7 + w)
    Hp, top = center_pad(h, ws)
    Wp, left = center_pad(w, ws)
    P = B * h * w
    delta, score = torch.randn(P, K, 64, generator=g) * 3, torch.randn(P, K, 64, generator=g)
    score[:, :, ::7] = 0.5                                              # exact score ties -> first index must win
    labels = torch.rand(P, K, generator=g) * 20
    coarse = torch.relu(labels[..., None] + delta)
    unshuf = lambda t: t.reshape(B, h, w, K, 8, 8).permute(0, 1, 4, 2, 5, 3).reshape(B, h * 8, w * 8, K)
    _, idx = torch.max(unshuf(score), -1, keepdim=True)
    d = torch.gather(unshuf(coarse), -1, idx).squeeze(-1) * 2
    ref = torch.median(d.reshape(B, h * 2, 4, w * 2, 4).permute(0, 1, 3, 2, 4).reshape(B, h * 2, w * 2, 16), -1)[0]
    pad = lambda t: torch.nn.functional.pad(t.reshape(B, h, w, K, 64), (0, 0, 0, 0, left, Wp - w - left, top, Hp - h - top),
                                            value=float("nan")).reshape(-1, 64)
    out = ops.select_median(cuda(pad(delta)), cuda(pad(score)), cuda(labels), B, h, w, K, Hp, Wp, top, left)
    assert torch.equal(out.cpu(), ref)                                  # selection + median: bit-exact


def test_refine_tail(ops):
    from nmrf_b200.hotpath import center_pad
    g = torch.Generator().manual_seed(9)
    B, h4, w4, H, W = 2, 10, 18, 37, 70
    Hp, top = center_pad(h4, 4)
    Wp, left = center_pad(w4, 4)
    delta = torch.randn(B, h4, w4, 16, generator=g) * 3
    dc = torch.rand(B, h4, w4, generator=g) * 30
    pred = torch.relu(dc[..., None] + delta).reshape(B, h4, w4, 4, 4).permute(0, 1, 3, 2, 4).reshape(B, h4 * 4, w4 * 4)
    dpad = torch.nn.functional.pad(delta, (0, 0, left, Wp - w4 - left, top, Hp - h4 - top), value=float("nan")).reshape(-1, 16)
    disp_pred, disp = ops.refine_tail(cuda(dpad), cuda(dc), Hp, Wp, top, left, H, W)
    assert torch.equal(disp_pred.cpu(), pred)
    assert torch.equal(disp.cpu(), (pred * 4)[:, :H, :W])


# ---- extended labels (label = hi + lo, include/nmrf_b200.h): held to FLOAT64 evaluations of the same formulas ----
def _split_hi_lo(x64):
    hi = x64.float()
    return hi, (x64 - hi.double()).float()


def test_warp_corr_embed_extended_labels(ops):
    from nmrf_b200.hotpath import center_pad
    g = torch.Generator().manual_seed(31)
    B, h, w, K, ws, norm = 1, 7, 40, 4, 6, 3.14 / 64
    cc1, cc2 = torch.randn(B, 64, h, w, generator=g), torch.randn(B, 64, h, w, generator=g)
    gw1, gw2 = torch.randn(B, 256, h, w, generator=g), torch.randn(B, 256, h, w, generator=g)
    lab64 = torch.rand(B, h, w, K, generator=g, dtype=torch.float64) * (w + 4) - 2
    hi, lo = _split_hi_lo(lab64)
    Hp, top = center_pad(h, ws)
    Wp, left = center_pad(w, ws)
    nh = lambda t: cuda(t.permute(0, 2, 3, 1))
    feat, enc = ops.warp_corr_embed(nh(cc1), nh(cc2), nh(gw1), nh(gw2), cuda(hi.reshape(-1, K)), K, Hp, Wp, top, left, norm,
                                    labels_lo=cuda(lo.reshape(-1, K)))
    d = lambda t: t.double()
    wg = O.warp_sample(d(gw2), lab64)
    corr = (d(gw1).permute(0, 2, 3, 1)[:, :, :, None, :] * wg).reshape(B, h, w, K, 32, 8).mean(-1)
    ref = torch.cat([d(cc1).permute(0, 2, 3, 1)[:, :, :, None, :].expand(B, h, w, K, 64), O.warp_sample(d(cc2), lab64), corr], -1)
    inner = feat.cpu().reshape(B, Hp, Wp, K, 160)[:, top:top + h, left:left + w]
    assert rel_err(inner, ref) <= 1e-6
    ref_enc = O.fourier_embed(lab64, norm)
    got = enc.cpu().reshape(B, Hp, Wp, K, 32)[:, top:top + h, left:left + w, :, :31]
    assert float((got.double() - ref_enc).abs().max()) <= 2.5e-7       # fp32 storage rounding only (values up to ~2.3), at every frequency
    # the plain fp32 path (labels_lo = NULL) cannot do that: one ulp of the label is ~1e-3 rad at 2^14
    _, enc32 = ops.warp_corr_embed(nh(cc1), nh(cc2), nh(gw1), nh(gw2), cuda(hi.reshape(-1, K)), K, Hp, Wp, top, left, norm)
    got32 = enc32.cpu().reshape(B, Hp, Wp, K, 32)[:, top:top + h, left:left + w, :, :31]
    assert float((got32.double() - ref_enc).abs().max()) >= 1e-4


def test_select_median_and_refine_tail_extended(ops):
    from nmrf_b200.hotpath import center_pad
    g = torch.Generator().manual_seed(77)
    B, h, w, K, ws = 1, 7, 13, 4, 6
    Hp, top = center_pad(h, ws)
    Wp, left = center_pad(w, ws)
    P = B * h * w
    delta, score = torch.randn(P, K, 64, generator=g) * 3, torch.randn(P, K, 64, generator=g)
    lab64 = torch.rand(P, K, generator=g, dtype=torch.float64) * 20
    hi, lo = _split_hi_lo(lab64)
    coarse = torch.relu(lab64[..., None] + delta.double())
    unshuf = lambda t: t.reshape(B, h, w, K, 8, 8).permute(0, 1, 4, 2, 5, 3).reshape(B, h * 8, w * 8, K)
    _, idx = torch.max(unshuf(score), -1, keepdim=True)
    d = torch.gather(unshuf(coarse), -1, idx).squeeze(-1) * 2
    ref = torch.median(d.reshape(B, h * 2, 4, w * 2, 4).permute(0, 1, 3, 2, 4).reshape(B, h * 2, w * 2, 16), -1)[0]
    pad = lambda t: torch.nn.functional.pad(t.reshape(B, h, w, K, 64), (0, 0, 0, 0, left, Wp - w - left, top, Hp - h - top),
                                            value=float("nan")).reshape(-1, 64)
    dc, dc_lo = ops.select_median(cuda(pad(delta)), cuda(pad(score)), cuda(hi), B, h, w, K, Hp, Wp, top, left, labels_lo=cuda(lo))
    ext = dc.cpu().double() + dc_lo.cpu().double()
    assert float((ext - ref).abs().max()) <= 1e-11                    # the double sum, carried as hi + lo (48 bits)
    assert torch.equal(dc.cpu(), ref.float())
    # refine tail on the extended disp_curr
    h4, w4, H, W = 2 * h, 2 * w, 8 * h - 3, 8 * w - 5
    Hp4, top4 = center_pad(h4, 4)
    Wp4, left4 = center_pad(w4, 4)
    dl = torch.randn(B, h4, w4, 16, generator=g) * 3
    pred = torch.relu(ref[..., None] + dl.double()).reshape(B, h4, w4, 4, 4).permute(0, 1, 3, 2, 4).reshape(B, h4 * 4, w4 * 4)
    dpad = torch.nn.functional.pad(dl, (0, 0, left4, Wp4 - w4 - left4, top4, Hp4 - h4 - top4), value=float("nan")).reshape(-1, 16)
    disp_pred, disp = ops.refine_tail(cuda(dpad), dc, Hp4, Wp4, top4, left4, H, W, disp_curr_lo=dc_lo)
    assert torch.equal(disp_pred.cpu(), pred.float())                   # correctly rounded double sums
    assert torch.equal(disp.cpu(), (pred.float() * 4)[:, :H, :W])


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["toy", "nmrf", "multi"])
@pytest.mark.parametrize("shapes_on_device", [True, False])
def test_msda_reference_golden(case, shapes_on_device):
    """the reference's own known-answer test (ops/test.py:53-75: rtol 1e-2, atol 1e-3); held to 1e-6 rel here."""
    import nmrf_b200.msda as msda
    g = golden("msda")
    shp, st = T(g[f"{case}_shapes"]), T(g[f"{case}_start"])
    if shapes_on_device:
        shp, st = shp.to(DEV), st.to(DEV)
    out = msda.MSDeformAttnFunction.apply(cuda(T(g[f"{case}_value"])), shp, st, cuda(T(g[f"{case}_loc"])),
                                          cuda(T(g[f"{case}_w"])), 64)
    ref = T(g[f"{case}_out"])
    assert np.allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-2, atol=1e-3)
    assert rel_err(out, ref) <= 2e-6


def test_msda_large_against_oracle():
    import nmrf_b200.msda as msda
    g = torch.Generator().manual_seed(21)
    N, M, Dh, Lq, P = 2, 8, 8, 5000, 4
    shp = torch.tensor([[50, 100]])
    st = torch.tensor([0])
    value = torch.randn(N, 5000, M, Dh, generator=g)
    loc = torch.rand(N, Lq, M, 1, P, 2, generator=g) * 1.2 - 0.1
    w = torch.softmax(torch.randn(N, Lq, M, P, generator=g), -1).reshape(N, Lq, M, 1, P)
    ref = O.ms_deform_attn(value, shp, st, loc, w)
    out = msda.ms_deform_attn_forward(cuda(value), shp.to(DEV), st.to(DEV), cuda(loc), cuda(w), 64)
    assert rel_err(out, ref) <= 1e-5


def test_msda_error_behaviour():
    """ms_deform_attn_cuda.cu:28-52: non-contiguous / non-CUDA inputs and bad im2col_step raise."""
    import nmrf_b200.msda as msda
    v = torch.zeros(3, 24, 2, 4, device=DEV)
    shp, st = torch.tensor([[6, 4]], device=DEV), torch.tensor([0], device=DEV)
    loc, w = torch.zeros(3, 5, 2, 1, 2, 2, device=DEV), torch.zeros(3, 5, 2, 1, 2, device=DEV)
    with pytest.raises(RuntimeError, match="contiguous"):
        msda.ms_deform_attn_forward(v.transpose(2, 3), shp, st, loc, w, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        msda.ms_deform_attn_forward(v.cpu(), shp, st, loc, w, 64)
    with pytest.raises(RuntimeError, match="im2col_step"):
        msda.ms_deform_attn_forward(v, shp, st, loc, w, 2)
    assert msda.ms_deform_attn_forward(v, shp, st, loc, w, 3).shape == (3, 5, 8)


# ---- N2: feature-extractor glue (instance norm statistics / fused apply) and the fused encoder ----
@pytest.mark.parametrize("N,H,W,C", [(2, 17, 23, 64), (1, 40, 31, 96), (2, 9, 12, 128), (1, 6, 7, 384)])
def test_instnorm_kernels(ops, N, H, W, C):
    import torch.nn.functional as F
    from nmrf_b200 import _lib
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(N, H, W, C, generator=g) * 3 + 1.5).cuda()
    r = torch.randn(N, H, W, C, generator=g).cuda()
    st = lambda: torch.zeros(N, C, 2, dtype=torch.float64, device="cuda")
    sx, sr = st(), st()
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib.nmrf_instnorm_stats(x.data_ptr(), N, H * W, C, sx.data_ptr(), s), "stats")
    _lib.check(_lib.lib.nmrf_instnorm_stats(r.data_ptr(), N, H * W, C, sr.data_ptr(), s), "stats")
    plain = torch.empty_like(x)
    _lib.check(_lib.lib.nmrf_instnorm_apply(x.data_ptr(), sx.data_ptr(), r.data_ptr(), sr.data_ptr(), N, H * W, C, 1, 1,
                                            plain.data_ptr(), s), "apply")
    nchw = lambda t: t.permute(0, 3, 1, 2).double()
    ref = F.relu(F.relu(F.instance_norm(nchw(x))) + F.instance_norm(nchw(r))).permute(0, 2, 3, 1)
    assert rel_err(plain, ref) <= 2e-6
    # plain residual, no norm on it, inner relu only
    _lib.check(_lib.lib.nmrf_instnorm_apply(x.data_ptr(), sx.data_ptr(), r.data_ptr(), None, N, H * W, C, 1, 0,
                                            plain.data_ptr(), s), "apply")
    ref = (F.relu(F.instance_norm(nchw(x))) + nchw(r)).permute(0, 2, 3, 1)
    assert rel_err(plain, ref) <= 2e-6


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride,pad", [
    (2, 17, 23, 64, 64, 3, 1, 1), (1, 20, 31, 64, 96, 3, 2, 1), (2, 9, 12, 128, 384, 3, 1, 1), (1, 12, 16, 64, 96, 1, 2, 0),
    (1, 6, 7, 256, 320, 3, 1, 1), (3, 5, 40, 96, 128, 1, 1, 0),
    # slab mode (3x3, stride 1) edge cases: three channel blocks per tap, one-pixel-wide / one-pixel-high images (every
    # horizontal / vertical neighbour is padding), and more row blocks than CTAs (slabs streamed across tiles of one CTA)
    (1, 30, 50, 96, 96, 3, 1, 1), (2, 40, 1, 64, 64, 3, 1, 1), (1, 1, 300, 64, 80, 3, 1, 1), (1, 160, 130, 64, 64, 3, 1, 1)])
def test_conv2d_against_float64(N, H, W, Cin, Cout, k, stride, pad):
    """nmrf_conv2d (implicit GEMM on tcgen05, 3xTF32, grouped accumulation) against torch's float64 convolution, next to the
    fp32 convolution of the reference arithmetic (torch CPU, ~2e-7 rms): at most 48 MMAs accumulate in place, which bounds the
    tensor core's round-toward-zero bias at ~8e-7 (measured 7.8e-7; cuDNN's TF32 path on split operands was at 5e-6)"""
    import torch.nn.functional as F
    from nmrf_b200.encoder import _Conv
    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cout * k * k)) ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, stride, pad)
    conv = _Conv(w.cuda(), stride, pad)
    Ho, Wo = conv.out_hw(H, W)
    assert (Ho, Wo) == tuple(ref.shape[-2:])
    y = torch.empty(N, Ho, Wo, Cout, device="cuda")
    conv(cuda(x.permute(0, 2, 3, 1)), y)
    rms = lambda a: float(((a.double().cpu() - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())
    ours, fp32 = rms(y.permute(0, 3, 1, 2)), rms(F.conv2d(x, w, None, stride, pad))
    assert rel_err(y.permute(0, 3, 1, 2), ref) <= 2e-6
    assert ours <= 1.0e-6 and fp32 <= 5e-7, (ours, fp32)


@pytest.mark.parametrize("B,H,W", [(1, 64, 96), (2, 40, 72), (1, 100, 180)])
def test_fused_encoder_against_float64_oracle(B, H, W):
    """encoder.FusedEncoder (image_prep incl. replicate padding, stem + residual stages + heads on nmrf_conv2d, InstanceNorm
    glue) against the oracle's feature extractor in float64; the fp32 oracle (reference arithmetic) is the yardstick"""
    from helpers import build_product_model
    from nmrf_b200.synthetic import synthetic_pair
    model, sd = build_product_model(64, 2, (1, 1, 1), 0, "reference")
    model = model.cuda()
    img1, img2 = synthetic_pair(B, H, W, 64, index=1)
    model.forward_device(img1.cuda(), img2.cuda())
    Hp, Wp = (H + 7) // 8 * 8, (W + 7) // 8 * 8
    plan = model.plan_for(B, 256, Hp // 8, Wp // 8, H, W)

    def feats(sd_, dt):
        both = torch.cat([O.pad_images(img1.to(dt), 8)[0], O.pad_images(img2.to(dt), 8)[0]], 0)
        f4, f8 = O.backbone_resnet(sd_, "backbone", both)
        out = {"f1_8": f8[:B], "f2_8": f8[B:], "context": O.conv_head(sd_, "dpn.proj", f8[:B])}
        for s, f in ((8, f8), (4, f4)):
            c, g_ = O.conv_head(sd_, "concatconv", f), O.conv_head(sd_, "gw", f)
            out.update({f"cc{s}0": c[:B], f"cc{s}1": c[B:], f"gw{s}0": g_[:B], f"gw{s}1": g_[B:]})
        return {k: v.permute(0, 2, 3, 1) for k, v in out.items()}
    t64, t32 = feats(O.to_float64(sd), torch.float64), feats(sd, torch.float32)
    got = {"f1_8": plan.f1_8, "f2_8": plan.f2_8, "context": plan.context}
    for s in (8, 4):
        for i in range(2):
            got[f"cc{s}{i}"] = getattr(plan, f"cc{s}")[i]
            got[f"gw{s}{i}"] = getattr(plan, f"gw{s}")[i]
    rms = lambda a, b: float(((a.double().cpu() - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt())
    for k in t64:
        ours, fp32 = rms(got[k], t64[k]), rms(t32[k], t64[k])
        assert rel_err(got[k], t64[k]) <= 1e-5, k
        assert ours <= max(3.0 * fp32, 2e-6), (k, ours, fp32)      # measured 1.4e-6 vs 6.6e-7 (round 1, cuDNN: 5.4e-6)


def test_module_path_for_foreign_encoders_agrees_with_fused_encoder():
    """the torch-module path (any encoder with the reference's return convention, convolutions through exactconv's cuDNN
    3xTF32 wrapper) fills the same buffers; agreement to that wrapper's accuracy"""
    from helpers import build_product_model
    from nmrf_b200.synthetic import synthetic_pair
    B, H, W = 1, 64, 96
    model, _ = build_product_model(64, 2, (1, 1, 1), 0, "reference")
    model = model.cuda()
    img1, img2 = (t.cuda() for t in synthetic_pair(B, H, W, 64, index=1))
    names = ("f1_8", "f2_8", "context")
    outs = {}
    for fused in (True, False):
        model.fused_encoder = fused
        model.forward_device(img1, img2)
        plan = model.plan_for(B, 256, H // 8, W // 8, H, W)
        outs[fused] = [getattr(plan, n).clone() for n in names] + [t.clone() for t in plan.cc8 + plan.gw8 + plan.cc4 + plan.gw4]
    for a, b in zip(outs[True], outs[False]):
        assert rel_err(a, b) <= 5e-5
