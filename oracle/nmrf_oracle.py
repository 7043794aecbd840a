"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the NMRF-Stereo inference hot path.

A functional, state-dict-driven restatement (torch CPU, fp32) of the reference
algorithm.  Every function cites the reference file:line it follows
(paths relative to aeolusguan/NMRF).  It is pinned against the real reference
by `oracle/make_golden.py` (run in the build container, where `/root/reference`
exists) and the committed fixtures under `tests/golden/`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.  The product path
(`nmrf_b200/`) never does.

Conventions shared with the CUDA path (documented deviations from the
reference are *canonicalisations of behaviour the reference leaves
implementation-defined*, nothing else):
  * top-K ties (DPN.py:121-125 overwrite suppressed entries with the same
    eps, torch.topk tie order is unspecified): value descending, then index
    ascending.
  * warp sampling (NMP.py:682-707) is evaluated directly at x - d (the
    reference's normalise -> grid_sample un-normalise round trip is the
    identity up to 1 ulp).
"""
from dataclasses import dataclass, field
import math

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    """Hyper-parameters of the path (nmrf/config/default.py:37-61)."""
    max_disp: int = 320          # DPN.MAX_DISP
    num_proposals: int = 4       # DPN.NUM_PROPOSALS
    cost_group: int = 4          # DPN.COST_GROUP
    num_prop_layers: int = 5
    num_infer_layers: int = 5
    num_refine_layers: int = 5
    window_size: int = 6         # NMP.WINDOW_SIZE
    refine_window_size: int = 4  # NMP.REFINE_WINDOW_SIZE
    n_heads: int = 4
    divis_by: int = 8            # DATASETS.DIVIS_BY
    backbone_prefix: str = "backbone"   # BACKBONE.COMPAT=True (NMRF.py:109-113)
    taps: dict = field(default_factory=dict, repr=False)  # stage taps, filled if not None


# --------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------
def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _ln(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _mlp_relu(sd, prefix, x, n=3):
    """NMP.py:54-66 `MLP`: Linear+ReLU ... Linear."""
    for i in range(n):
        x = _lin(sd, f"{prefix}.layers.{i}", x)
        if i < n - 1:
            x = F.relu(x)
    return x


def _timm_mlp(sd, prefix, x):
    """timm Mlp: fc2(GELU(fc1(x))) (call sites NMP.py:337,537,675)."""
    return _lin(sd, prefix + ".fc2", F.gelu(_lin(sd, prefix + ".fc1", x)))


def fourier_embed(c, normalizer):
    """NMP.py:35-51 with N_freqs=15, logscale: [sin(c' 2^i) (15), cos(c' 2^i) (15), c']."""
    freq = 2 ** torch.linspace(0, 14, 15, dtype=c.dtype, device=c.device)
    cs = (c * normalizer).unsqueeze(-1)
    f = cs * freq
    return torch.cat([f.sin(), f.cos(), cs], dim=-1)


# --------------------------------------------------------------------------------------
# features (NOT the hot path; needed so the oracle is a whole forward)
# --------------------------------------------------------------------------------------
def pad_images(img, divis_by):
    """frame_utils.py:259-275, mode='proposal': replicate-pad right/bottom."""
    ht, wd = img.shape[-2:]
    pad_ht = (((ht // divis_by) + 1) * divis_by - ht) % divis_by
    pad_wd = (((wd // divis_by) + 1) * divis_by - wd) % divis_by
    return F.pad(img, [0, pad_wd, 0, pad_ht], mode="replicate"), (pad_ht, pad_wd)


def _res_block(sd, p, x, stride):
    """backbone.py:16-45 ResidualBlock with InstanceNorm2d (no affine)."""
    y = F.relu(F.instance_norm(F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)))
    y = F.relu(F.instance_norm(F.conv2d(y, sd[p + ".conv2.weight"], None, 1, 1)))
    if (p + ".downsample.0.weight") in sd:
        x = F.instance_norm(F.conv2d(x, sd[p + ".downsample.0.weight"], sd[p + ".downsample.0.bias"], stride))
    return F.relu(x + y)


def backbone_resnet(sd, p, x):
    """backbone.py:85-98: returns [feat@1/4, feat@1/8]."""
    x = 2 * (x / 255.0) - 1.0
    x = F.relu(F.instance_norm(F.conv2d(x, sd[p + ".conv1.weight"], None, 2, 3)))
    x = _res_block(sd, p + ".layer1.0", x, 1)
    x = _res_block(sd, p + ".layer1.1", x, 1)
    x = _res_block(sd, p + ".layer2.0", x, 2)
    x = _res_block(sd, p + ".layer2.1", x, 1)
    x = _res_block(sd, p + ".layer3.0", x, 1)
    x = _res_block(sd, p + ".layer3.1", x, 1)
    x = F.conv2d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])
    return [x, F.avg_pool2d(x, 2, 2)]


def conv_head(sd, p, x):
    """NMRF.py:56-65 / DPN.py:45-49: conv3x3 -> InstanceNorm -> ReLU -> conv1x1 (no biases)."""
    x = F.relu(F.instance_norm(F.conv2d(x, sd[p + ".0.weight"], None, 1, 1)))
    return F.conv2d(x, sd[p + ".3.weight"])


# --------------------------------------------------------------------------------------
# A1  cost volume   (submodule.py:4-23)
# --------------------------------------------------------------------------------------
def cost_volume(f1, f2, D, G):
    """cv[p,g,d] = mean_{c in group g} f1[b,c,y,x] * f2[b,c,y,x-d]  (0 for x<d),
    returned pixel-major [B*h*w, G, D] as DPN.py:117 reshapes it."""
    B, C, h, w = f1.shape
    cv = f1.new_zeros(B, G, D, h, w)
    for d in range(min(D, w)):
        prod = f1[..., d:] * f2[..., : w - d]
        cv[:, :, d, :, d:] = prod.view(B, G, C // G, h, w - d).mean(2)
    return cv.permute(0, 3, 4, 1, 2).reshape(B * h * w, G, D).contiguous()


# --------------------------------------------------------------------------------------
# A2  seeds   (DPN.py:115-125)
# --------------------------------------------------------------------------------------
def seed_extraction(sd, cv, K, eps=1e-3):
    x = F.relu(F.conv1d(cv, sd["dpn.mlp.0.weight"], sd["dpn.mlp.0.bias"], padding=2))
    x = F.relu(F.conv1d(x, sd["dpn.mlp.2.weight"], sd["dpn.mlp.2.bias"], padding=2))
    cost = F.conv1d(x, sd["dpn.mlp.4.weight"], sd["dpn.mlp.4.bias"], padding=2).squeeze(-2)
    prob = F.softmax(cost, dim=-1)
    out = F.max_pool1d(prob.unsqueeze(-2), 3, 1, 1).squeeze(-2)
    nlm = (prob != out) & (prob > eps)
    prob_ = prob.clone()
    prob_[nlm] = eps
    # canonical top-K: value desc, index asc (stable sort keeps index order among ties)
    _, order = torch.sort(prob_, dim=-1, descending=True, stable=True)
    return prob, prob_, order[:, :K].contiguous()


# --------------------------------------------------------------------------------------
# A3-A7  label-seed propagation   (NMP.py:603-667, 401-600; DPN.py:128-132)
# --------------------------------------------------------------------------------------
def sample_cost(cv, seeds):
    """NMP.py:618-634: cost[p,n,g*9+(o+4)] = cv[p,g,clamp(seed+o,0,D-1)]."""
    P, G, D = cv.shape
    K = seeds.shape[1]
    off = torch.arange(-4, 5, device=seeds.device)
    idx = (seeds[..., None] + off).clamp(0, D - 1)                 # [P,K,9]
    g = cv[:, None, :, :].expand(P, K, G, D).gather(3, idx[:, :, None, :].expand(P, K, G, 9))
    return g.reshape(P, K, G * 9)


def stripe_attention(sd, prefix, q, k, v, vertical):
    """CSWinAttention.forward with split_size 1 (NMP.py:451-505).
    q,k,v: [B,h,w,K,64] (one channel half).  vertical=True: windows are image
    columns (H_sp=h,W_sp=1); else rows (H_sp=1,W_sp=w).  2 heads x 32."""
    B, h, w, K, _ = q.shape
    if vertical:
        perm = lambda t: t.permute(0, 2, 1, 3, 4)      # [B, w, h, K, 64]  windows=w, line=h
    else:
        perm = lambda t: t                             # [B, h, w, K, 64]  windows=h, line=w
    q, k, v = perm(q), perm(k), perm(v)
    Bq, Wn, L, _, _ = q.shape
    split = lambda t: t.reshape(B, Wn, L * K, 2, 32).permute(0, 1, 3, 2, 4)   # [B,Wn,2,L*K,32]
    qh, kh, vh = split(q) * (32 ** -0.5), split(k), split(v)
    attn = qh @ kh.transpose(-2, -1)
    pix = torch.arange(L, device=q.device).repeat_interleave(K)
    same = pix[:, None] == pix[None, :]
    mask = torch.zeros(L * K, L * K, dtype=q.dtype, device=q.device)
    mask[same] = float("-inf")
    mask.fill_diagonal_(0.0)                            # NMP.py:203-208
    attn = F.softmax(attn + mask, dim=-1)
    out = attn @ vh                                     # [B,Wn,2,L*K,32]
    # LePE (NMP.py:433-449): depth-wise 3x3 conv per (window, proposal) image, zero pad at
    # the window border, summed over proposals, minus the centre tap of the other proposals.
    wgt = sd[prefix + ".get_v.weight"]                  # [64,1,3,3]
    img = v.permute(0, 1, 3, 4, 2).reshape(B * Wn * K, 64, L)       # [(b win n), 64, L]
    img = img[..., None] if vertical else img[:, :, None, :]        # [., 64, L, 1] or [., 64, 1, L]
    conv = F.conv2d(img, wgt, None, 1, 1, 1, 64).reshape(B, Wn, K, 64, L)
    R = conv.sum(2, keepdim=True)                                     # [B,Wn,1,64,L]
    vimg = img.reshape(B, Wn, K, 64, L)
    others = vimg.sum(2, keepdim=True) - vimg
    lepe = R - wgt[:, 0, 1, 1][None, None, None, :, None] * others   # [B,Wn,K,64,L]
    lepe = lepe.permute(0, 1, 4, 2, 3)                                # [B,Wn,L,K,64]
    out = out.permute(0, 1, 3, 2, 4).reshape(B, Wn, L, K, 64) + lepe
    return out.permute(0, 2, 1, 3, 4) if vertical else out           # [B,h,w,K,64]


def cswin_layer(sd, p, x, context, B, h, w, K):
    """CSWinNMP.forward_pre (NMP.py:544-574). x: [P*K... as [P,K,128]], context [B,h,w,64]."""
    t = _ln(sd, p + ".norm1", x).reshape(B, h, w, K, 128)
    qk_in = torch.cat([t, context[:, :, :, None, :].expand(B, h, w, K, 64)], dim=-1)
    q, k, v = _lin(sd, p + ".q", qk_in), _lin(sd, p + ".k", qk_in), _lin(sd, p + ".v", t)
    x1 = stripe_attention(sd, p + ".attns.0", q[..., :64], k[..., :64], v[..., :64], True)
    x2 = stripe_attention(sd, p + ".attns.1", q[..., 64:], k[..., 64:], v[..., 64:], False)
    msg = torch.cat([x1, x2], dim=-1).reshape(x.shape)
    x = x + _lin(sd, p + ".proj", msg)
    return x + _timm_mlp(sd, p + ".mlp", _ln(sd, p + ".norm2", x))


def propagation(sd, cfg, cv, seeds, context, B, h, w):
    """Propagation.forward (NMP.py:636-667) + prop_head (DPN.py:131-132)."""
    K = seeds.shape[1]
    p = "dpn.propagation"
    cost = sample_cost(cv, seeds)
    cf = _lin(sd, p + ".cost_encoder.2", F.gelu(_lin(sd, p + ".cost_encoder.0", cost)))
    seeds_f = seeds.to(cv.dtype)
    enc = fourier_embed(seeds_f, 3.14 / 64)
    x = F.linear(torch.cat([cf, enc], dim=-1), sd[p + ".proj.weight"])
    if cfg.taps is not None:
        cfg.taps["prop_embed"] = x
    for i in range(cfg.num_prop_layers):
        x = cswin_layer(sd, f"{p}.layers.{i}.nmp", x, context, B, h, w, K)
        if cfg.taps is not None:
            cfg.taps[f"prop_layer{i}"] = x
    x = _ln(sd, p + ".norm", x)
    delta = _mlp_relu(sd, "dpn.prop_head", x).squeeze(-1)
    return F.relu(delta + seeds_f)


# --------------------------------------------------------------------------------------
# A8  warp + group correlation + ffn embed   (NMP.py:682-720, 735-743)
# --------------------------------------------------------------------------------------
def warp_sample(fmap, labels):
    """Bilinear sample of fmap [B,C,h,w] along x at (x - label), same row, zeros outside
    [0,w-1] (grid_sample bilinear/zeros/align_corners=True, NMP.py:695-706).
    labels [B,h,w,K] -> [B,h,w,K,C]."""
    B, C, h, w = fmap.shape
    xs = torch.arange(w, dtype=labels.dtype, device=labels.device).view(1, 1, w, 1)
    xr = xs - labels
    x0 = torch.floor(xr)
    a = xr - x0
    x0 = x0.long()
    x1 = x0 + 1
    fm = fmap.permute(0, 2, 3, 1)                                   # [B,h,w,C]

    def tap(xi):
        ok = ((xi >= 0) & (xi <= w - 1)).unsqueeze(-1)
        xi = xi.clamp(0, w - 1)
        g = fm[:, :, :, None, :].expand(B, h, w, xi.shape[-1], C).gather(
            2, xi[..., None].expand(B, h, w, xi.shape[-1], C))
        return g * ok

    return tap(x0) * (1 - a).unsqueeze(-1) + tap(x1) * a.unsqueeze(-1)


def warp_corr_embed(sd, p, labels, f1, f2, f1_gw, f2_gw):
    """labels [B,h,w,K] -> tokens [B,h,w,K,128] (Inference.forward NMP.py:735-741;
    Refinement.forward NMP.py:839-844 is the K=1 case)."""
    B, _, h, w = f1.shape
    K = labels.shape[-1]
    wg = warp_sample(f2_gw, labels)                                  # [B,h,w,K,256]
    f1g = f1_gw.permute(0, 2, 3, 1)[:, :, :, None, :]
    corr = (f1g * wg).reshape(B, h, w, K, 32, 8).mean(-1)            # 32 groups of 8 (NMP.py:716-719)
    wc = warp_sample(f2, labels)                                     # [B,h,w,K,64]
    f1c = f1.permute(0, 2, 3, 1)[:, :, :, None, :].expand(B, h, w, K, 64)
    return _timm_mlp(sd, p + ".ffn", torch.cat([f1c, wc, corr], dim=-1))


# --------------------------------------------------------------------------------------
# A10  intra-pixel attention   (BasicAttention.forward_pre NMP.py:90-108)
# --------------------------------------------------------------------------------------
def basic_attention(sd, p, x, enc):
    """x [P,K,128], enc [P,K,31]."""
    P, K, _ = x.shape
    t = _ln(sd, p + ".norm1", x)
    qk_in = torch.cat([t, enc], dim=-1)
    hd = lambda z: z.reshape(P, K, 4, 32).permute(0, 2, 1, 3)
    q, k, v = hd(_lin(sd, p + ".q", qk_in)), hd(_lin(sd, p + ".k", qk_in)), hd(_lin(sd, p + ".v", t))
    attn = F.softmax((q @ k.transpose(-2, -1)) * (32 ** -0.5), dim=-1)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(P, K, 128)
    return x + _lin(sd, p + ".proj", out)


# --------------------------------------------------------------------------------------
# A11  (shifted-)window attention with contextual RPE   (NMP.py:241-289, 343-364, 195-239)
# --------------------------------------------------------------------------------------
def _window_index(Hp, Wp, ws, shift, device=None):
    """Token-grid index of every window slot, shift done by indexing instead of roll
    (NMP.py:249-250: rolled[yr] = orig[(yr+shift) % Hp]); plus Swin region ids in rolled
    coordinates (NMP.py:221-232)."""
    yr = torch.arange(Hp, device=device)
    xr = torch.arange(Wp, device=device)
    yo, xo = (yr + shift) % Hp, (xr + shift) % Wp
    lin = (yo[:, None] * Wp + xo[None, :])                          # [Hp,Wp] rolled -> original linear idx
    win = lin.reshape(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    if shift > 0:
        band = lambda n: (torch.arange(n, device=device) >= n - ws).long() + (torch.arange(n, device=device) >= n - shift).long()
        reg = band(Hp)[:, None] * 3 + band(Wp)[None, :]
    else:
        reg = torch.zeros(Hp, Wp, dtype=torch.long, device=device)
    reg = reg.reshape(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    return win, reg


def window_attention(sd, p, qkv, B, Hp, Wp, K, ws, shift, self_edge_mask):
    """qkv [B,Hp,Wp,K,384] -> [B,Hp,Wp,K,128]."""
    nh = 4
    Pn = ws * ws
    dev = qkv.device
    win, reg = _window_index(Hp, Wp, ws, shift, dev)                # [Wn,Pn]
    Wn = win.shape[0]
    flat = qkv.reshape(B, Hp * Wp, K, 3, nh, 32)
    g = flat[:, win.reshape(-1)].reshape(B, Wn, Pn, K, 3, nh, 32)
    q, k, v = g[..., 0, :, :], g[..., 1, :, :], g[..., 2, :, :]     # [B,Wn,Pn,K,nh,32]
    table = sd[p + ".relative_position_enc_table"]                  # [(2ws-1)^2, 384]
    cy, cx = torch.meshgrid(torch.arange(ws, device=dev), torch.arange(ws, device=dev), indexing="ij")
    cy, cx = cy.reshape(-1), cx.reshape(-1)
    rel = (cy[:, None] - cy[None, :] + ws - 1) * (2 * ws - 1) + (cx[:, None] - cx[None, :] + ws - 1)
    rpe = table[rel.reshape(-1)].reshape(Pn, Pn, nh, 96)            # NMP.py:257-260
    Rq, Rk, Rv = rpe[..., 0:32], rpe[..., 32:64], rpe[..., 64:96]
    s = 32 ** -0.5
    qs = q * s
    qk = torch.einsum("bwpnhc,bwqmhc->bwhpnqm", qs, k)
    qr = torch.einsum("bwpnhc,pqhc->bwhpnq", qs, Rk)
    kr = torch.einsum("bwqmhc,pqhc->bwhpqm", k, Rq * s)
    logits = qk + qr[..., None] + kr[:, :, :, :, None, :, :]
    mask = torch.zeros(Wn, Pn, K, Pn, K, dtype=qkv.dtype, device=dev)
    if shift > 0:
        diff = reg[:, :, None] != reg[:, None, :]                   # [Wn,Pn,Pn]
        mask = mask.masked_fill(diff[:, :, None, :, None], float("-inf"))
    if self_edge_mask:
        eyeP = torch.eye(Pn, dtype=torch.bool, device=dev)[:, None, :, None]
        eyeK = torch.eye(K, dtype=torch.bool, device=dev)[None, :, None, :]
        mask = mask.masked_fill((eyeP & ~eyeK)[None], float("-inf"))
    logits = logits + mask[None, :, None]
    A = F.softmax(logits.reshape(B, Wn, nh, Pn, K, Pn * K), dim=-1).reshape(logits.shape)
    out = torch.einsum("bwhpnqm,bwqmhc->bwpnhc", A, v) + torch.einsum("bwhpnqm,pqhc->bwpnhc", A, Rv)
    res = qkv.new_zeros(B, Hp * Wp, K, 128)
    res[:, win.reshape(-1)] = out.reshape(B, Wn * Pn, K, 128)
    return res.reshape(B, Hp, Wp, K, 128)


def swin_layer(sd, p, x, enc, B, Hp, Wp, K, ws, shift, self_edge_mask):
    """SwinNMP.forward_pre (NMP.py:350-364). x [P,K,128], enc [P,K,31]."""
    t = _ln(sd, p + ".norm1", x)
    qkv = _lin(sd, p + ".qkv", torch.cat([t, enc], dim=-1)).reshape(B, Hp, Wp, K, 384)
    msg = window_attention(sd, p + ".attn", qkv, B, Hp, Wp, K, ws, shift, self_edge_mask)
    x = x + _lin(sd, p + ".proj", msg.reshape(x.shape))
    return x + _timm_mlp(sd, p + ".mlp", _ln(sd, p + ".norm2", x))


def _center_pad(t, B, h, w, ws):
    """NMP.py:745-762: zero pad token grid to a multiple of ws (top=pad//2)."""
    Hpad, Wpad = (ws - h % ws) % ws, (ws - w % ws) % ws
    top, left = Hpad // 2, Wpad // 2
    t = t.reshape(B, h, w, *t.shape[1:])
    t = F.pad(t, (0, 0, 0, 0, left, Wpad - left, top, Hpad - top))
    return t.reshape(B * (h + Hpad) * (w + Wpad), *t.shape[3:]), h + Hpad, w + Wpad, top, left


def mrf_stack(sd, cfg, p, labels, f1, f2, f1_gw, f2_gw, n_layers, ws, normalizer, with_self):
    """Inference.forward (NMP.py:722-798) / Refinement.forward (NMP.py:828-900).
    labels [B,h,w,K] -> [B*h*w, K, 128] (after crop + final LayerNorm)."""
    B, _, h, w = f1.shape
    K = labels.shape[-1]
    x = warp_corr_embed(sd, p, labels, f1, f2, f1_gw, f2_gw).reshape(B * h * w, K, 128)
    enc = fourier_embed(labels.reshape(B * h * w, K), normalizer)
    if cfg.taps is not None:
        cfg.taps[p + "_embed"] = x
    x, Hp, Wp, top, left = _center_pad(x, B, h, w, ws)
    enc = _center_pad(enc, B, h, w, ws)[0]
    for i in range(n_layers):
        shift = 0 if i % 2 == 0 else ws // 2                        # NMRF.py:72
        lp = f"{p}.layers.{i}"
        if with_self:
            x = basic_attention(sd, lp + ".self_nmp", x, enc)       # NMP.py:955
        x = swin_layer(sd, lp + ".nmp", x, enc, B, Hp, Wp, K, ws, shift, with_self)
        if cfg.taps is not None:
            cfg.taps[f"{p}_layer{i}"] = x
    x = x.reshape(B, Hp, Wp, K, 128)[:, top:top + h, left:left + w].reshape(B * h * w, K, 128)
    return _ln(sd, p + ".norm", x)


# --------------------------------------------------------------------------------------
# A12  heads + select   (NMRF.py:218-232)
# --------------------------------------------------------------------------------------
def infer_select(sd, tgt, labels, B, h, w):
    """tgt [P,K,128], labels [P,K] -> coarse [B,8h,8w,K], score, disp_curr [B,2h,2w]."""
    K = labels.shape[1]
    delta = _mlp_relu(sd, "infer_head", tgt)                         # [P,K,64]
    coarse = F.relu(labels[..., None] + delta)
    score = 0.25 * _lin(sd, "infer_score_head", tgt)
    unshuf = lambda t: t.reshape(B, h, w, K, 8, 8).permute(0, 1, 4, 2, 5, 3).reshape(B, h * 8, w * 8, K)
    coarse, score = unshuf(coarse), unshuf(score)
    _, idx = torch.max(score, dim=-1, keepdim=True)                  # first max on ties
    d = torch.gather(coarse, -1, idx).squeeze(-1) * 2
    d = d.reshape(B, h * 2, 4, w * 2, 4).permute(0, 1, 3, 2, 4).reshape(B, h * 2, w * 2, 16)
    disp_curr = torch.median(d, dim=-1)[0]                           # lower median
    return coarse, score, idx.squeeze(-1), disp_curr


# --------------------------------------------------------------------------------------
# whole forward   (NMRF.forward NMRF.py:189-262, eval mode)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def hot_path(sd, cfg, f1_list, f2_list):
    """Everything after the backbone.  f*_list = [feat@1/8, feat@1/4]."""
    taps = cfg.taps
    f1_8, f2_8 = f1_list[0], f2_list[0]
    B, C, h8, w8 = f1_8.shape
    K = cfg.num_proposals
    D = cfg.max_disp // 8
    cv = cost_volume(f1_8, f2_8, D, cfg.cost_group)
    prob, prob_nms, seeds = seed_extraction(sd, cv, K)
    context = conv_head(sd, "dpn.proj", f1_8).permute(0, 2, 3, 1)
    labels = propagation(sd, cfg, cv, seeds, context, B, h8, w8)     # [P8,K]
    if taps is not None:
        taps.update(cost_volume=cv, prob=prob, prob_nms=prob_nms, seeds=seeds, context=context, labels=labels)

    fc1, fc2 = conv_head(sd, "concatconv", f1_8), conv_head(sd, "concatconv", f2_8)
    fg1, fg2 = conv_head(sd, "gw", f1_8), conv_head(sd, "gw", f2_8)
    if taps is not None:
        taps.update(f8=(f1_8, f2_8), cc8=(fc1, fc2), gw8=(fg1, fg2))
    tgt = mrf_stack(sd, cfg, "inference", labels.reshape(B, h8, w8, K), fc1, fc2, fg1, fg2,
                    cfg.num_infer_layers, cfg.window_size, 3.14 / 64, True)
    coarse, score, sel, disp_curr = infer_select(sd, tgt, labels, B, h8, w8)
    if taps is not None:
        taps.update(infer_tgt=tgt, coarse=coarse, score=score, sel=sel, disp_curr=disp_curr)

    f1_4, f2_4 = f1_list[1], f2_list[1]
    fc1, fc2 = conv_head(sd, "concatconv", f1_4), conv_head(sd, "concatconv", f2_4)
    fg1, fg2 = conv_head(sd, "gw", f1_4), conv_head(sd, "gw", f2_4)
    h4, w4 = f1_4.shape[-2:]
    if taps is not None:
        taps.update(cc4=(fc1, fc2), gw4=(fg1, fg2))
    tgt = mrf_stack(sd, cfg, "refinement", disp_curr[..., None], fc1, fc2, fg1, fg2,
                    cfg.num_refine_layers, cfg.refine_window_size, 3.14 / 128, False)
    delta = _mlp_relu(sd, "refine_head", tgt.squeeze(1)).reshape(B, h4, w4, 4, 4)   # NMRF.py:238-242
    disp_pred = F.relu(disp_curr[..., None, None] + delta)
    disp_pred = disp_pred.permute(0, 1, 3, 2, 4).reshape(B, h4 * 4, w4 * 4)
    if taps is not None:
        taps.update(refine_tgt=tgt)
    return {
        "proposal": labels.reshape(B, -1, K),
        "prob": prob,
        "initial_proposal": seeds.to(prob.dtype).reshape(B, -1, K),
        "disp_pred": disp_pred,
        "disp_padded": disp_pred * 4,
    }


def to_float64(sd, device=None):
    """state-dict with every floating-point tensor cast to float64 (integer buffers untouched), optionally moved to `device`.
    The oracle is device-agnostic torch code: on a CUDA device (fp64, no TF32 involved) it serves the GPU tests as a fast
    float64 truth for stage-level comparisons."""
    return {k: (v.double() if v.dtype.is_floating_point else v).to(device or v.device) for k, v in sd.items()}


@torch.no_grad()
def forward(sd, cfg, img1, img2):
    """NMRF.forward(sample) on CPU: images [B,3,H,W] in 0..255 -> output dict.
    Arithmetic type = the state-dict's: fp32 restates the reference; a state-dict cast with
    `to_float64` gives the float64 "truth" the parity tests measure both implementations against."""
    H, W = img1.shape[-2:]
    w0 = sd[cfg.backbone_prefix + ".conv1.weight"]
    img1, _ = pad_images(img1.to(w0.device, w0.dtype), cfg.divis_by)
    img2, _ = pad_images(img2.to(w0.device, w0.dtype), cfg.divis_by)
    feats = backbone_resnet(sd, cfg.backbone_prefix, torch.cat([img1, img2], 0))
    f4a, f4b = feats[0].chunk(2, 0)
    f8a, f8b = feats[1].chunk(2, 0)
    out = hot_path(sd, cfg, [f8a, f4a], [f8b, f4b])
    out["disp"] = out.pop("disp_padded")[:, :H, :W]
    return out


# --------------------------------------------------------------------------------------
# A14  multi-scale deformable attention forward
#      (ops/src/cuda/ms_deform_im2col_cuda.cuh:33-84, 237-299)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """value [N,S,M,Dh]; shapes [L,2] (H,W); loc [N,Lq,M,L,P,2] (x,y in 0..1); w [N,Lq,M,L,P]
    -> [N,Lq,M*Dh].  h_im = loc_y*H - 0.5 (cuh:285-286); a sample contributes only if
    -1 < h_im < H and -1 < w_im < W (cuh:288); 4-tap zero-padded bilinear (cuh:38-83)."""
    N, S, M, Dh = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    out = value.new_zeros(N, Lq, M, Dh)
    for l in range(L):
        H, W = int(spatial_shapes[l, 0]), int(spatial_shapes[l, 1])
        st = int(level_start_index[l])
        val = value[:, st:st + H * W]                                 # [N,HW,M,Dh]
        loc = sampling_locations[:, :, :, l]                          # [N,Lq,M,P,2]
        w_im = loc[..., 0] * W - 0.5
        h_im = loc[..., 1] * H - 0.5
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h0, w0 = torch.floor(h_im), torch.floor(w_im)
        lh, lw = h_im - h0, w_im - w0
        h0, w0 = h0.long(), w0.long()
        acc = value.new_zeros(N, Lq, M, P, Dh)
        for dy, dx, wt in ((0, 0, (1 - lh) * (1 - lw)), (0, 1, (1 - lh) * lw),
                           (1, 0, lh * (1 - lw)), (1, 1, lh * lw)):
            hy, wx = h0 + dy, w0 + dx
            ok = inside & (hy >= 0) & (hy <= H - 1) & (wx >= 0) & (wx <= W - 1)
            lin = (hy.clamp(0, H - 1) * W + wx.clamp(0, W - 1))      # [N,Lq,M,P]
            src = val.permute(0, 2, 1, 3)                             # [N,M,HW,Dh]
            idx = lin.permute(0, 2, 1, 3).reshape(N, M, Lq * P, 1).expand(N, M, Lq * P, Dh)
            g = src.gather(2, idx).reshape(N, M, Lq, P, Dh).permute(0, 2, 1, 3, 4)
            acc = acc + g * (wt * ok)[..., None]
        out = out + (acc * attention_weights[:, :, :, l][..., None]).sum(3)
    return out.reshape(N, Lq, M * Dh)


# --------------------------------------------------------------------------------------
# N4  evaluation metrics and the KITTI writer's encoding
#     (nmrf/utils/evaluation.py:326-359, 398-407; nmrf/utils/frame_utils.py:237-239)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def disp_metrics(disp_pr, disp_gt, valid_gt, only_valid, max_disp, thres):
    """DispEvaluator.process + evaluate for a batch [B,H,W]: per-image epe / d1 / bad-t over the valid pixels (images
    without valid pixels skipped), then the mean over images; d1 and bad-t in percent."""
    epes, d1s, bads = [], [], {t: [] for t in thres}
    for pr, gt, vg in zip(disp_pr, disp_gt, valid_gt):
        valid = (vg & (gt < max_disp)) if only_valid else (gt < max_disp)
        epe = torch.abs(pr - gt).flatten()
        val = valid.flatten()
        if not bool(val.any()):
            continue
        epes.append(epe[val].mean().item())
        d1s.append(((epe[val] > 3) & (epe[val] / gt.flatten()[val] > 0.05)).float().mean().item())
        for t in thres:
            bads[t].append((epe > float(t))[val].float().mean().item())
    res = {"epe": torch.tensor(epes).mean().item(), "d1": torch.tensor(d1s).mean().item() * 100}
    for t in thres:
        res[f"bad {t}"] = torch.tensor(bads[t]).mean().item() * 100
    return res


def kitti_u16(disp):
    """writeDispKITTI's encoding: np.round(disp * 256).astype(np.uint16)"""
    import numpy as np
    return np.round(disp.cpu().numpy() * 256).astype(np.uint16)
