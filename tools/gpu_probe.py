"""GPU-side probes (run under gpurun): per-launch timing of the hot path and the torch feature-extractor cost
under different cuDNN settings.  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

sys.argv = [sys.argv[0]]
import bench  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    dev = torch.device("cuda", 0)
    out = {}
    model, sd = bench.build_model(dev)
    w = bench.WORKLOAD
    from nmrf_b200.synthetic import synthetic_pair
    img1, img2 = (t.to(dev) for t in synthetic_pair(1, w["H"], w["W"], w["max_disp"], 0))
    model.forward_device(img1, img2)
    plan = next(iter(model._plans.values()))
    rows = plan.launches.run_timed(reps=5)
    out["launches"] = [dict(what=a, sym=b, ms=round(c, 4), gflop=round(d / 1e9, 3), mb=round(e / 1e6, 3)) for a, b, c, d, e in rows]
    out["hot_path_ms"] = sum(r[2] for r in rows)
    out["forward_eager_ms"] = timeit(lambda: model.forward_device(img1, img2))

    # torch part under different cuDNN settings
    x = torch.cat([img1, img2], 0)
    xp = torch.nn.functional.pad(x, [0, 0, 0, 4], mode="replicate")

    def torch_part(cl):
        xi = xp.contiguous(memory_format=torch.channels_last) if cl else xp.contiguous()
        feats = model.backbone(xi)[::-1]
        res = [model.dpn.proj(feats[0][:1])]
        for f in feats:
            res.append(model.concatconv(f)); res.append(model.gw(f))
        return res

    for cl in (True, False):
        for bm in (False, True):
            for tf32 in (False, True):
                with torch.backends.cudnn.flags(enabled=True, benchmark=bm, allow_tf32=tf32):
                    with torch.no_grad():
                        try:
                            ms = timeit(lambda: torch_part(cl), n=5, warm=3)
                        except Exception as ex:  # noqa
                            ms = str(ex)[:100]
                out[f"torch_part_ms[channels_last={cl},benchmark={bm},tf32={tf32}]"] = ms
    # backbone only vs heads only (NCHW, benchmark on, exact fp32)
    with torch.backends.cudnn.flags(enabled=True, benchmark=True, allow_tf32=False), torch.no_grad():
        xi = xp.contiguous()
        out["backbone_ms[nchw,bench]"] = timeit(lambda: model.backbone(xi), n=5)
        feats = model.backbone(xi)[::-1]
        out["heads8_ms[nchw,bench]"] = timeit(lambda: (model.concatconv(feats[0]), model.gw(feats[0])), n=5)
        out["heads4_ms[nchw,bench]"] = timeit(lambda: (model.concatconv(feats[1]), model.gw(feats[1])), n=5)
        # instance norm alone
        t = torch.randn(2, 128, 136, 240, device=dev)
        out["instancenorm_128x136x240_ms"] = timeit(lambda: torch.nn.functional.instance_norm(t), n=10)
        c = torch.nn.Conv2d(256, 128, 3, 1, 1, bias=False).to(dev)
        f = torch.randn(2, 256, 136, 240, device=dev)
        out["conv3x3_256to128_136x240_ms"] = timeit(lambda: c(f), n=10)
        out["conv3x3_gflop"] = 2 * 2 * 136 * 240 * 256 * 128 * 9 / 1e9
    # per-module time of the feature extractor (exact fp32), to find the slow cuDNN layers
    recs = []
    def pre(m, i):
        e = torch.cuda.Event(enable_timing=True); e.record(); m._t0 = e
    def post(name):
        def f(m, i, o):
            e = torch.cuda.Event(enable_timing=True); e.record()
            recs.append((name, type(m).__name__, tuple(i[0].shape), tuple(o.shape), m._t0, e))
        return f
    hs = []
    for name, m in model.backbone.named_modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.InstanceNorm2d)):
            hs.append(m.register_forward_pre_hook(pre)); hs.append(m.register_forward_hook(post(name)))
    with torch.backends.cudnn.flags(enabled=True, benchmark=False, allow_tf32=False), torch.no_grad():
        xi = xp.contiguous(memory_format=torch.channels_last)
        model.backbone(xi); recs.clear()
        model.backbone(xi)
    torch.cuda.synchronize()
    out["backbone_modules_ms"] = [dict(name=n, kind=k, inp=list(i), out=list(o), ms=round(a.elapsed_time(b), 4)) for n, k, i, o, a, b in recs]
    for h in hs:
        h.remove()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ("launches", "backbone_modules_ms")}, indent=1))
    for r in out["backbone_modules_ms"]:
        print(r)
    agg = {}
    for r in out["launches"]:
        key = r["what"].split(".")[-1] if r["sym"] == "nmrf_token_gemm" else r["sym"]
        stage = "prop" if r["what"].startswith(("prop", "cost")) else "inference" if r["what"].startswith("infer") else "refinement"
        a = agg.setdefault((stage, r["sym"], key), [0.0, 0, 0.0])
        a[0] += r["ms"]; a[1] += 1; a[2] += r["gflop"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(k, round(v[0], 3), "ms", v[1], "launches", round(v[2] / max(v[0], 1e-9), 1), "TFLOP/s")


if __name__ == "__main__":
    main()
