"""Shared test utilities: golden fixtures, synthetic weights, oracle configs."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from nmrf_b200.synthetic import state_dict_fingerprint, synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def make_cfg(max_disp, K, L):
    import nmrf_b200
    cfg = nmrf_b200.get_cfg()
    cfg.DPN.MAX_DISP, cfg.DPN.NUM_PROPOSALS = int(max_disp), int(K)
    cfg.NMP.NUM_PROP_LAYERS, cfg.NMP.NUM_INFER_LAYERS, cfg.NMP.NUM_REFINE_LAYERS = (int(x) for x in L)
    return cfg


def build_product_model(max_disp, K, L, seed, mode):
    """nmrf_b200.NMRF (CPU, parameters from synthetic_state_dict). Returns (model, state_dict)."""
    import nmrf_b200
    model = nmrf_b200.build_model(make_cfg(max_disp, K, L)).eval()
    sd = synthetic_state_dict(model.state_dict(), seed=seed, mode=mode)
    model.load_state_dict(sd, strict=True)
    return model, sd


def oracle_cfg(max_disp, K, L, taps=True):
    from oracle import nmrf_oracle as O
    return O.OracleConfig(max_disp=int(max_disp), num_proposals=int(K), num_prop_layers=int(L[0]),
                          num_infer_layers=int(L[1]), num_refine_layers=int(L[2]), taps={} if taps else None)


def check_fingerprint(sd, expected):
    got = state_dict_fingerprint(sd)
    assert abs(got - float(expected)) <= 1e-6 * abs(float(expected)), (
        f"synthetic weights differ from the ones the golden fixture was generated with "
        f"(fingerprint {got} vs {float(expected)}): torch RNG drift?")


def epe(a, b):
    return float((a.double() - b.double()).abs().mean())


def _rel(a, b):
    """max |a-b| / max |b|"""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _rms(a, b):
    """rms(a-b) / rms(b): the noise level of a stage, insensitive to single outliers"""
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt().clamp_min(1e-30))


def truth_forward(sd, max_disp, K, L, img1, img2, device=None):
    """FLOAT64 truth: the oracle with state-dict and images cast to float64 (on `device`: the GPU in the GPU tests -- plain
    torch fp64 kernels as the checker, ~100x faster than the host for the big configs).  Returns (outputs, taps), on the CPU."""
    from oracle import nmrf_oracle as O
    cfg = oracle_cfg(max_disp, K, L)
    out = O.forward(O.to_float64(sd, device), cfg, img1, img2)
    one = lambda v: v.cpu() if torch.is_tensor(v) else tuple(one(x) for x in v) if isinstance(v, tuple) else v
    cpu = lambda d: {k: one(v) for k, v in d.items()}
    return cpu(out), cpu(cfg.taps)


def load_truth_features(plan, tt):
    """EXPERIMENT: overwrite the hot path's inputs with the float64 oracle's feature maps (rounded to fp32), so that the
    hot path's own arithmetic is measured without the encoder's"""
    nhwc = lambda t: t.permute(0, 2, 3, 1).float().contiguous()
    plan.f1_8.copy_(nhwc(tt["f8"][0])); plan.f2_8.copy_(nhwc(tt["f8"][1]))
    plan.context.copy_(tt["context"].float())
    for name in ("cc8", "gw8", "cc4", "gw4"):
        for i in range(2):
            getattr(plan, name)[i].copy_(nhwc(tt[name][i]))


def parity_metrics(model, sd, max_disp, K, L, img1, img2, truth_features=False):
    """The CUDA path AND the fp32 reference arithmetic (CPU oracle, fp32) against FLOAT64 truth on the same inputs.

    Why float64 truth: the reference's own fp32 forward is ~1e-3 px (EPE) away from exact arithmetic at these weights, because
    a handful of argmax-over-K / median decisions (NMRF.py:228-231) are numerically tied and ANY fp32 rounding pattern flips
    some of them (a flipped 4x4 median block moves 16 pixels by whole pixels).  Comparing two fp32 implementations with each
    other therefore measures the sum of both error processes; comparing each with the float64 result measures them one at a
    time, and the fp32 reference's own distance is the yardstick the CUDA path is held to (test_gpu_e2e.assert_parity).

    Returns (metrics, model outputs, truth outputs).  metrics["cuda"] / metrics["ref32"] hold, per implementation: EPE,
    max error, fraction of pixels > 1e-3 px, selection flips, seed agreement, label error; metrics["stage"] the relative
    error (max and rms) of every stage boundary for both."""
    from oracle import nmrf_oracle as O
    B, _, H, W = img1.shape
    out = model({"img1": img1, "img2": img2})              # builds the plan, fills its input buffers
    plan = model.plan_for(B, plan_C(model), *feat_hw(model, H, W), H, W)
    dev = model.device if model.device.type == "cuda" else None
    t_out, tt = truth_forward(sd, max_disp, K, L, img1, img2, dev)
    if truth_features:
        load_truth_features(plan, tt)
    taps = {k: v.cpu() for k, v in plan.run_with_taps().items()}
    if truth_features:
        out = dict(out, disp=taps["disp"], proposal=taps["labels"].reshape(B, -1, K))
    ocfg = oracle_cfg(max_disp, K, L)
    r_out = O.forward(sd, ocfg, img1, img2)                # the reference arithmetic: torch CPU fp32
    rt = ocfg.taps
    g, h8, w8 = plan.geom, plan.h8, plan.w8
    Hp8, Wp8, top, left = g["Hp8"], g["Wp8"], g["top8"], g["left8"]

    def cuda_sel():
        sc = taps["score"].reshape(B, Hp8, Wp8, K, 64)[:, top:top + h8, left:left + w8]
        sc = sc.reshape(B, h8, w8, K, 8, 8).permute(0, 1, 4, 2, 5, 3).reshape(B, h8 * 8, w8 * 8, K)
        return sc.argmax(-1)

    def side(disp, sel, seeds, labels, proposal):
        d = (disp.double().cpu() - t_out["disp"]).abs()
        same_seed = (seeds == tt["seeds"]).all(-1)
        pn = tt["prob_nms"]
        return {"EPE": float(d.mean()), "max_err_px": float(d.max()), "frac_px_err_gt_1e-3": float((d > 1e-3).double().mean()),
                "selection_flips": int((sel != tt["sel"]).sum()), "n_selections": int(sel.numel()),
                "seed_rows_identical": float(same_seed.double().mean()),
                "seed_value_gap_max": float((pn.gather(1, seeds) - pn.gather(1, tt["seeds"])).abs().max()),
                "labels_abs_err_max": float((labels.double() - tt["labels"])[same_seed].abs().max()),
                "proposal_EPE": float((proposal.double().cpu().reshape(-1, K) - tt["labels"]).abs().mean()),
                "prob_abs_err": None}, same_seed

    m = {}
    m["cuda"], same_c = side(out["disp"], cuda_sel(), taps["seeds"], taps["labels"], out["proposal"])
    m["ref32"], same_r = side(r_out["disp"], rt["sel"], rt["seeds"], rt["labels"], r_out["proposal"])
    m["cuda"]["prob_abs_err"] = float((taps["prob"].double() - tt["prob"]).abs().max())
    m["ref32"]["prob_abs_err"] = float((rt["prob"].double() - tt["prob"]).abs().max())
    m["EPE_cuda_vs_ref32"] = float((out["disp"].double().cpu() - r_out["disp"].double()).abs().mean())

    # ---- stage boundaries (tokens of pixels whose seeds agree with the truth) ------------------------------------------
    sd64 = O.to_float64(sd, dev)
    both64 = torch.cat([O.pad_images(img1.double(), 8)[0], O.pad_images(img2.double(), 8)[0]], 0).to(sd64["backbone.conv1.weight"].device)
    feats64 = [f.cpu() for f in O.backbone_resnet(sd64, "backbone", both64)]
    feats32 = O.backbone_resnet(sd, "backbone", torch.cat([O.pad_images(img1, 8)[0], O.pad_images(img2, 8)[0]], 0))
    f64 = feats64[1].chunk(2, 0)[0].permute(0, 2, 3, 1)
    stage = {"features@1/8": {"cuda": (plan.f1_8.cpu(), f64), "ref32": (feats32[1].chunk(2, 0)[0].permute(0, 2, 3, 1), f64)},
             "cost_volume": {"cuda": (taps["cost_volume"], tt["cost_volume"]), "ref32": (rt["cost_volume"], tt["cost_volume"])}}
    for k in ["prop_embed"] + [f"prop_layer{i}" for i in range(int(L[0]))]:
        stage[k] = {"cuda": (taps[k][same_c], tt[k][same_c]), "ref32": (rt[k][same_r], tt[k][same_r])}
    for i in range(int(L[1])):
        k = f"inference_layer{i}"
        stage[k] = {"cuda": (taps[k], tt[k]), "ref32": (rt[k], tt[k])}
    # refinement tokens sit downstream of the discrete selection: compared where disp_curr agrees with the truth to 1e-4 in a
    # 13x13 neighbourhood (the reach of the shifted 4x4 windows over the stack), i.e. away from flipped blocks
    Hp4, Wp4, t4, l4 = g["Hp4"], g["Wp4"], g["top4"], g["left4"]

    def calm(dc):
        bad = ((dc.double().cpu() - tt["disp_curr"]).abs() > 1e-4).float()
        return F.max_pool2d(bad[:, None], 13, 1, 6)[:, 0] == 0
    calm_c, calm_r = calm(taps["disp_curr"]), calm(rt["disp_curr"])
    crop = lambda t: t.reshape(B, Hp4, Wp4, 128)[:, t4:t4 + 2 * h8, l4:l4 + 2 * w8]
    for i in range(int(L[2])):
        k = f"refinement_layer{i}"
        stage[k] = {"cuda": (crop(taps[k])[calm_c], crop(tt[k])[calm_c]), "ref32": (crop(rt[k])[calm_r], crop(tt[k])[calm_r])}
    m["stage"] = {k: {who: {"max": _rel(a, b), "rms": _rms(a, b)} if a.numel() else None for who, (a, b) in v.items()}
                  for k, v in stage.items()}
    m["frac_px_calm"] = {"cuda": float(calm_c.double().mean()), "ref32": float(calm_r.double().mean())}
    return m, out, t_out


def plan_C(model):
    enc = model.backbone if model.compat else model.image_encoder
    return enc.output_dim


def feat_hw(model, H, W):
    d = model.divis_by
    Hp, Wp = H + (((H // d) + 1) * d - H) % d, W + (((W // d) + 1) * d - W) % d
    return Hp // 8, Wp // 8
