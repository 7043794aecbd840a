"""Seeded synthetic weights and stereo pairs (no datasets/checkpoints are available offline).

`synthetic_state_dict` draws every floating-point tensor of a model's state-dict from a CPU
`torch.Generator`, iterating keys in sorted order, so the result depends only on (keys, shapes,
seed, mode) -- not on module construction order -- and is identical in the build container
(where the reference consumes it) and on the GPU box.

mode "reference": the reference's init distributions (Linear normal std .02 -- the +-2 trunc never
    bites at that std --, zero biases, unit LayerNorms, kaiming convs, zero RPE tables) except
    `dpn.prop_head.layers.2` which is drawn like any other Linear (its zero init
    would hide the propagation stack).
mode "stress": O(1) activations everywhere (weights ~ 1.5/sqrt(fan_in), random biases, LayerNorm
    affine and RPE tables): every term of every kernel matters; used for stage-level parity.
"""
import torch


def synthetic_state_dict(template_sd, seed=0, mode="reference"):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(template_sd.keys()):
        v = template_sd[k]
        if (not v.dtype.is_floating_point) or v.numel() == 0:
            out[k] = v.detach().cpu().clone()
            continue
        shape = tuple(v.shape)
        parts = k.split(".")
        is_norm = len(parts) >= 2 and parts[-2].startswith("norm")
        if mode == "reference":
            if is_norm:
                t = torch.ones(shape) if k.endswith("weight") else torch.zeros(shape)
            elif k.endswith(".bias"):
                t = torch.zeros(shape)
            elif "relative_position_enc_table" in k:
                t = torch.zeros(shape)
            elif len(shape) >= 3:                                   # conv: kaiming_normal fan_out
                fan_out = shape[0]
                for s in shape[2:]:
                    fan_out *= s
                t = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
            else:
                t = torch.randn(shape, generator=g) * 0.02
        elif mode == "stress":
            if is_norm:
                t = 0.1 * torch.randn(shape, generator=g)
                if k.endswith("weight"):
                    t = t + 1
            elif k.endswith(".bias"):
                t = 0.1 * torch.randn(shape, generator=g)
            elif "relative_position_enc_table" in k:
                t = 0.5 * torch.randn(shape, generator=g)
            else:
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
                t = torch.randn(shape, generator=g) * (1.5 / fan_in ** 0.5)
        else:
            raise ValueError(mode)
        out[k] = t.to(v.dtype)
    return out


def state_dict_fingerprint(sd):
    """Order-independent digest used by the golden fixtures to detect RNG drift."""
    tot = 0.0
    for k in sorted(sd.keys()):
        v = sd[k]
        if v.dtype.is_floating_point and v.numel():
            tot += float(v.double().abs().sum()) * (1 + (len(k) % 7))
    return tot


def synthetic_pair(B, H, W, max_disp, index=0):
    """img1 ~ U(0,255); img2 = img1 shifted left by s px + N(0, 2^2) so the cost
    volume has real peaks (pure noise maximises top-K ties)."""
    g = torch.Generator().manual_seed(1000 + index)
    img1 = torch.rand(B, 3, H, W, generator=g) * 255
    s = 8 + ((7 * index) % max(max_disp // 2, 1))
    img2 = torch.roll(img1, -s, dims=-1) + torch.randn(B, 3, H, W, generator=g) * 2
    return img1, img2
