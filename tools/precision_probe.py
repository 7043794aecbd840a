"""GPU-side precision probe (run under gpurun): error of every dense / attention kernel against FLOAT64, next to the fp32
arithmetic of the reference (torch CPU fp32 = MKL sgemm, and torch CUDA fp32 with TF32 off), so that the noise each kernel
injects into the discrete decisions downstream can be compared with what the reference itself injects.

    python tools/precision_probe.py  ->  gpurun_out/precision_<tag>.json   (NMRF_B200_LIB selects another build)
rms = rms(err) / rms(value), max = max|err| / max|value|, bias = mean(err * sign(value)) / mean|value| (a systematic
shrink towards zero shows up here: round-toward-zero accumulation).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False


def stats(a, ref):
    a, ref = a.double().cpu(), ref.double().cpu()
    e = a - ref
    return {"rms": float((e ** 2).mean().sqrt() / (ref ** 2).mean().sqrt()), "max": float(e.abs().max() / ref.abs().max()),
            "bias": float((e * ref.sign()).mean() / ref.abs().mean())}


def main():
    import nmrf_b200.ops as ops
    from oracle import nmrf_oracle as O
    tag = sys.argv[1] if len(sys.argv) > 1 else "default"
    res = {}
    g = torch.Generator().manual_seed(0)
    for name, rows, K, N, scale in (("qkv_like", 16384, 160, 384, 0.02), ("proj_like", 16384, 128, 128, 0.02), ("fc2_like", 16384, 512, 128, 0.02),
                                    ("unit_scale", 16384, 128, 128, 1.0 / 128 ** 0.5)):
        X = torch.randn(rows, K, generator=g)
        if name == "fc2_like":
            X = torch.nn.functional.gelu(X)              # positive-skewed activations like the hidden layer
        W = torch.randn(N, K, generator=g) * scale
        ref = X.double() @ W.double().T
        Xc, Wc = X.cuda(), W.cuda()
        r = {}
        r["tcgen05_3xtf32"] = stats(ops.token_gemm(Xc, Wc, Wt=ops.pack_weight_tiles(Wc)), ref)
        r["fma_fp32"] = stats(ops.token_gemm(Xc, Wc), ref)
        r["torch_cpu_fp32"] = stats(X @ W.T, ref)
        r["torch_cuda_fp32"] = stats(Xc @ Wc.T, ref)
        res["gemm_" + name] = r
    # fused block tail
    rows = 16384
    att, x = torch.randn(rows, 128, generator=g), torch.randn(rows, 128, generator=g)
    Wp, W1, W2 = (torch.randn(128, 128, generator=g) * 0.02, torch.randn(512, 128, generator=g) * 0.02, torch.randn(128, 512, generator=g) * 0.02)
    z128, z512, one = torch.zeros(128), torch.zeros(512), torch.ones(128)

    def tail(att, x, Wp, W1, W2):
        x1 = x + att @ Wp.T
        t = torch.nn.functional.layer_norm(x1, (128,))
        return x1 + torch.nn.functional.gelu(t @ W1.T) @ W2.T
    ref = tail(att.double(), x.double(), Wp.double(), W1.double(), W2.double())
    ws = ops.pack_mlp_stream(Wp.cuda(), W1.cuda(), W2.cuda())
    out = ops.mlp_chain(att.cuda(), ws, z128.cuda(), (one.cuda(), z128.cuda()), z512.cuda(), z128.cuda(), E=x.cuda(), e_identity=True)
    # the update (what the block adds to the residual stream) is what carries the arithmetic error
    res["block_tail_update"] = {"tcgen05_3xtf32": stats(out.cpu().double() - x.double(), ref - x.double()),
                                "torch_cpu_fp32": stats(tail(att, x, Wp, W1, W2).double() - x.double(), ref - x.double()),
                                "torch_cuda_fp32": stats(tail(att.cuda(), x.cuda(), Wp.cuda(), W1.cuda(), W2.cuda()).cpu().double() - x.double(), ref - x.double())}
    # attention kernels against the float64 oracle functions
    from helpers import build_product_model
    _, sd = build_product_model(192, 4, (1, 1, 1), 0, "reference")
    B, h, w, K = 1, 34, 60, 4
    qkv = torch.randn(B, h, w, K, 384, generator=g)
    p = "dpn.propagation.layers.0.nmp"

    def stripe(qkv, sd):
        q, k, v = qkv[..., :128], qkv[..., 128:256], qkv[..., 256:]
        x1 = O.stripe_attention(sd, p + ".attns.0", q[..., :64], k[..., :64], v[..., :64], True)
        x2 = O.stripe_attention(sd, p + ".attns.1", q[..., 64:], k[..., 64:], v[..., 64:], False)
        return torch.cat([x1, x2], -1).reshape(-1, 128)
    ref = stripe(qkv.double(), O.to_float64(sd))
    r = {"torch_cpu_fp32": stats(stripe(qkv, sd), ref)}
    from nmrf_b200 import _lib
    for impl, nm in ((1, "tcgen05_3xtf32"), (0, "fma_fp32")):
        _lib.lib.nmrf_set_attention_impl(impl)
        r[nm] = stats(ops.stripe_attention(qkv.reshape(-1, 384).cuda(), B, h, w, K, sd[p + ".attns.0.get_v.weight"].cuda(),
                                           sd[p + ".attns.1.get_v.weight"].cuda()), ref)
    res["stripe_attention"] = r
    Hp, Wp_ = 36, 60
    qkv = torch.randn(B, Hp, Wp_, K, 384, generator=g)
    table = 0.02 * torch.randn(121, 384, generator=g)
    for shift in (0, 3):
        ref = O.window_attention({"a.relative_position_enc_table": table.double()}, "a", qkv.double(), B, Hp, Wp_, K, 6, shift, True).reshape(-1, 128)
        r = {"torch_cpu_fp32": stats(O.window_attention({"a.relative_position_enc_table": table}, "a", qkv, B, Hp, Wp_, K, 6, shift, True).reshape(-1, 128), ref)}
        for impl, nm in ((1, "mma_3xtf32"), (0, "fma_fp32")):
            _lib.lib.nmrf_set_attention_impl(impl)
            r[nm] = stats(ops.window_attention(qkv.reshape(-1, 384).cuda(), table.cuda(), B, Hp, Wp_, K, 6, shift, True), ref)
        res[f"window_attention_shift{shift}"] = r
    _lib.lib.nmrf_set_attention_impl(1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"precision_{tag}.json"), "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print(k)
        for who, s in v.items():
            print(f"   {who:18s} rms {s['rms']:.2e}  max {s['max']:.2e}  bias {s['bias']:+.2e}")


if __name__ == "__main__":
    main()
