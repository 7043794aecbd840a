"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shims.py) on seeded inputs and weights.

Run in the build container (the reference does not exist on the GPU box):
    python oracle/make_golden.py

The only intervention on the reference is the top-K tie rule: `torch.topk` inside
nmrf.models.DPN is routed to a stable sort (value desc, index asc).  The reference leaves the
order among exact ties implementation-defined (DPN.py:121-125 overwrites suppressed entries with
one constant), so this selects ONE of its valid outputs; the native-topk seeds are stored as well
so the tests can check they differ from the canonical ones only inside tie groups.

Fixtures
  e2e_tiny.npz    reference-init weights, 1x96x160, D=8 K=2 L=2/2/2    whole-forward outputs
  e2e_small.npz   reference-init weights, 2x75x150, D=24 K=4 L=3/3/3   whole-forward outputs
                  (D > w8: the cost-volume edge case; both window pads; odd layer count)
  stages.npz      stress weights, 1x36x68, D=24 K=4 L=2/3/2            inputs+outputs of every stage
  msda.npz        ops/test.py's toy problem (seed 3) + NMRF-shaped cases  vs ms_deform_attn_core_pytorch
  truth_c1.npz    BASELINE config 2 (SceneFlow 1x540x960, D=24, K=4, L=8/8/8): FLOAT64 oracle disparity + decisions,
  truth_c1b.npz   the same at the checkpoint-compatible depth 5/5/5,
  truth_c2.npz    BASELINE config 3 (KITTI 8x375x1248, D=24, K=4, L=8/8/8)
                  and, as scalars, how far the REAL fp32 reference itself is from that float64 result (EPE, max, flipped
                  selections): the yardstick the GPU parity tests hold the CUDA path to (`python oracle/make_golden.py truth`)
Weights are NOT stored (MBs): they are regenerated from the seed by
nmrf_b200.synthetic.synthetic_state_dict; a fingerprint guards against RNG drift.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from nmrf_b200.synthetic import state_dict_fingerprint, synthetic_pair, synthetic_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _canon_topk(x, k, dim=-1):
    v, i = torch.sort(x, dim=dim, descending=True, stable=True)
    return v[..., :k], i[..., :k]


class _TorchProxy:
    """stands in for the `torch` global of nmrf.models.DPN; only topk is changed"""
    def __init__(self, canonical):
        self.canonical = canonical

    def __getattr__(self, name):
        if name == "topk" and self.canonical:
            return _canon_topk
        return getattr(torch, name)


def build(cfgd, mode, seed):
    model = ref_shims.build_reference_model(max_disp=cfgd["max_disp"], num_proposals=cfgd["K"],
                                            num_prop_layers=cfgd["L"][0], num_infer_layers=cfgd["L"][1],
                                            num_refine_layers=cfgd["L"][2])
    sd = synthetic_state_dict(model.state_dict(), seed=seed, mode=mode)
    model.load_state_dict(sd, strict=True)
    return model.eval(), sd


def run(model, img1, img2, canonical=True):
    import nmrf.models.DPN as DPNmod
    DPNmod.torch = _TorchProxy(canonical)
    try:
        with torch.no_grad():
            return model({"img1": img1.clone(), "img2": img2.clone()})
    finally:
        DPNmod.torch = torch


def npy(t):
    return t.detach().cpu().numpy()


def e2e_fixture(name, B, H, W, cfgd, index):
    model, sd = build(cfgd, "reference", seed=0)
    img1, img2 = synthetic_pair(B, H, W, cfgd["max_disp"], index)
    out = run(model, img1, img2, canonical=True)
    native = run(model, img1, img2, canonical=False)
    np.savez_compressed(
        os.path.join(OUT, name),
        img1=npy(img1), img2=npy(img2), max_disp=cfgd["max_disp"], K=cfgd["K"], L=np.array(cfgd["L"]),
        index=index, weight_seed=0, fingerprint=state_dict_fingerprint(sd),
        keys=np.array(sorted(sd.keys())), shapes=np.array([str(tuple(sd[k].shape)) for k in sorted(sd.keys())]),
        prob=npy(out["prob"]), initial_proposal=npy(out["initial_proposal"]), proposal=npy(out["proposal"]),
        disp=npy(out["disp"]), disp_pred=npy(out["disp_pred"]),
        initial_proposal_native=npy(native["initial_proposal"]))
    print(name, "disp range", float(out["disp"].min()), float(out["disp"].max()),
          "native-vs-canonical seed rows differing:",
          int((native["initial_proposal"] != out["initial_proposal"]).any(-1).sum()))


def stages_fixture():
    cfgd = dict(max_disp=192, K=4, L=(2, 3, 2))
    model, sd = build(cfgd, "stress", seed=7)
    img1, img2 = synthetic_pair(1, 36, 68, cfgd["max_disp"], 1)
    caps = {}

    def hook(name):
        def f(mod, inp, out):
            caps[name] = (inp, out)
        return f
    hs = [model.dpn.register_forward_hook(hook("dpn")), model.dpn.propagation.register_forward_hook(hook("propagation")),
          model.inference.register_forward_hook(hook("inference")), model.refinement.register_forward_hook(hook("refinement")),
          model.dpn.proj.register_forward_hook(hook("dpn_proj")), model.backbone.register_forward_hook(hook("backbone"))]
    out = run(model, img1, img2, canonical=True)
    for h in hs:
        h.remove()
    feats = caps["backbone"][1]                      # [f@1/4, f@1/8] for cat(left,right)
    f4a, f4b = feats[0].chunk(2, 0)
    f8a, f8b = feats[1].chunk(2, 0)
    cv_in, fmap1_list = caps["dpn"][0]               # cost volume [B,G,D,h,w]
    cv_pix, prob, seeds_f, labels = caps["dpn"][1]
    d = dict(
        img1=npy(img1), img2=npy(img2), max_disp=192, K=4, L=np.array(cfgd["L"]), weight_seed=7,
        fingerprint=state_dict_fingerprint(sd),
        f8a=npy(f8a), f8b=npy(f8b),
        cost_volume=npy(cv_pix), prob=npy(prob), seeds=npy(seeds_f).astype(np.int64), labels=npy(labels[-1]),
        context=npy(caps["dpn_proj"][1]),
        prop_in_seeds=npy(caps["propagation"][0][1]),
        prop_memory=npy(caps["propagation"][1][0][0]),
        disp=npy(out["disp"]), disp_pred=npy(out["disp_pred"]), proposal=npy(out["proposal"]),
    )
    for name, n in (("inference", cfgd["L"][1]), ("refinement", cfgd["L"][2])):
        inp, o = caps[name]
        short = "inf" if name == "inference" else "ref"
        d[f"{short}_labels"] = npy(inp[0])
        for k, t in zip(("cc1", "cc2", "gw1", "gw2"), inp[1:]):
            d[f"{short}_{k}"] = npy(t)
        d[f"{short}_out"] = npy(o[0])
    np.savez_compressed(os.path.join(OUT, "stages"), **d)
    print("stages: labels range", float(labels.min()), float(labels.max()), "disp", float(out["disp"].min()),
          float(out["disp"].max()))


def msda_fixture():
    ref_shims.install()
    from ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch
    d = {}
    cases = [("toy", 1, 2, 2, 2, [(6, 4), (3, 2)], 2),        # ops/test.py:16-23
             ("nmrf", 2, 8, 8, 200, [(12, 20)], 4),           # the neck's shape: 8 heads x 8, 1 level, 4 points
             ("multi", 1, 4, 16, 100, [(8, 10), (4, 5), (2, 3)], 3)]
    torch.manual_seed(3)                                      # ops/test.py:23
    for name, N, M, Dh, Lq, shapes, P in cases:
        shp = torch.as_tensor(shapes, dtype=torch.long)
        lsi = torch.cat((shp.new_zeros((1,)), shp.prod(1).cumsum(0)[:-1]))
        S, L = int(shp.prod(1).sum()), len(shapes)
        value = torch.rand(N, S, M, Dh) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        if name != "toy":
            loc = loc * 1.3 - 0.15                            # exercise the out-of-range skip and border taps
        w = torch.rand(N, Lq, M, L, P) + 1e-5
        w = w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)
        out = ms_deform_attn_core_pytorch(value, shp, loc, w)
        out64 = ms_deform_attn_core_pytorch(value.double(), shp, loc.double(), w.double())
        d.update({f"{name}_value": npy(value), f"{name}_shapes": npy(shp), f"{name}_start": npy(lsi),
                  f"{name}_loc": npy(loc), f"{name}_w": npy(w), f"{name}_out": npy(out)})
        if name == "toy":
            d["toy_out64"] = npy(out64)
    np.savez_compressed(os.path.join(OUT, "msda"), **d)
    print("msda fixtures:", [c[0] for c in cases])


def truth_fixture(name, B, H, W, cfgd, index):
    """float64 'truth' for a BASELINE config: the oracle run in float64 (state-dict and images cast; one pair at a time --
    pairs are independent, InstanceNorm/LayerNorm are per sample), next to the REAL reference in fp32 on the whole batch.
    The disparity is stored rounded to fp32 (its rounding, <= 6e-6 px, is far below the 1e-3 px bar)."""
    from oracle import nmrf_oracle as O
    model, sd = build(cfgd, "reference", seed=0)
    img1, img2 = synthetic_pair(B, H, W, cfgd["max_disp"], index)
    ref = run(model, img1, img2, canonical=True)                       # the real reference, fp32, whole batch
    sd64 = O.to_float64(sd)
    mk = lambda: O.OracleConfig(max_disp=cfgd["max_disp"], num_proposals=cfgd["K"], num_prop_layers=cfgd["L"][0],
                                num_infer_layers=cfgd["L"][1], num_refine_layers=cfgd["L"][2], taps={})
    disp64, sel64, seeds64, lab64, sel32 = [], [], [], [], []
    for b in range(B):
        c64 = mk()
        o64 = O.forward(sd64, c64, img1[b:b + 1], img2[b:b + 1])
        disp64.append(o64["disp"]); sel64.append(c64.taps["sel"]); seeds64.append(c64.taps["seeds"]); lab64.append(c64.taps["labels"])
        c32 = mk()
        O.forward(sd, c32, img1[b:b + 1], img2[b:b + 1])
        sel32.append(c32.taps["sel"])
        del c64, c32
    disp64, sel64, seeds64 = torch.cat(disp64), torch.cat(sel64), torch.cat(seeds64)
    lab64, sel32 = torch.cat(lab64), torch.cat(sel32)
    d = (ref["disp"].double() - disp64).abs()
    per_pair = d.flatten(1).mean(1)
    seeds_ref = ref["initial_proposal"].reshape(-1, cfgd["K"]).long()
    np.savez_compressed(
        os.path.join(OUT, name),
        B=B, H=H, W=W, max_disp=cfgd["max_disp"], K=cfgd["K"], L=np.array(cfgd["L"]), index=index, weight_seed=0,
        fingerprint=state_dict_fingerprint(sd),
        disp64=npy(disp64.float()), sel64=npy(sel64).astype(np.uint8), seeds64=npy(seeds64).astype(np.uint8),
        proposal64=npy(lab64.float()),
        ref32_epe=float(d.mean()), ref32_max=float(d.max()), ref32_epe_per_pair=npy(per_pair),
        ref32_frac_gt_1e3=float((d > 1e-3).double().mean()),
        ref32_seed_rows_identical=float((seeds_ref == seeds64).all(-1).double().mean()),
        oracle32_selection_flips=int((sel32 != sel64).sum()), n_selections=int(sel64.numel()),
        ref32_proposal_epe=float((ref["proposal"].reshape(-1, cfgd["K"]).double() - lab64).abs().mean()),
        threads=torch.get_num_threads())
    print(name, "real fp32 reference vs float64 oracle: EPE %.3e max %.3f px, >1e-3 px: %.2e of pixels, seeds identical %.6f, "
          "fp32-oracle selection flips %d / %d" % (float(d.mean()), float(d.max()), float((d > 1e-3).double().mean()),
                                                   float((seeds_ref == seeds64).all(-1).double().mean()),
                                                   int((sel32 != sel64).sum()), sel64.numel()), flush=True)


TRUTH = {
    "truth_c1": (1, 540, 960, dict(max_disp=192, K=4, L=(8, 8, 8)), 0),
    "truth_c1b": (1, 540, 960, dict(max_disp=192, K=4, L=(5, 5, 5)), 0),
    "truth_c2": (8, 375, 1248, dict(max_disp=192, K=4, L=(8, 8, 8)), 0),
}


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "truth":
        for name in (sys.argv[2:] or list(TRUTH)):
            truth_fixture(name, *TRUTH[name])
        sys.exit(0)
    e2e_fixture("e2e_tiny", 1, 96, 160, dict(max_disp=64, K=2, L=(2, 2, 2)), 0)
    e2e_fixture("e2e_small", 2, 75, 150, dict(max_disp=192, K=4, L=(3, 3, 3)), 3)
    stages_fixture()
    msda_fixture()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
