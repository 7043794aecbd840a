"""GPU-side parity report (run under gpurun): the CUDA path against the CPU oracle, stage by stage, with the
decomposition SURVEY.md H2 asks for -- arithmetic error vs discrete-decision flips.

    python tools/parity_report.py [--config small|half|full] [--gemm tc|simt]

Prints one JSON object and writes it to gpurun_out/parity_<config>_<gemm>.json.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    "small": dict(B=2, H=136, W=240, max_disp=192, K=4, L=(2, 3, 2)),
    "half": dict(B=1, H=270, W=480, max_disp=192, K=4, L=(5, 5, 5)),
    "full": dict(B=1, H=540, W=960, max_disp=192, K=4, L=(8, 8, 8)),
}


def report(cfgname, gemm):
    os.environ["NMRF_B200_GEMM"] = gemm
    import torch
    from helpers import build_product_model, oracle_cfg
    from nmrf_b200.synthetic import synthetic_pair
    from oracle import nmrf_oracle as O
    c = CONFIGS[cfgname]
    B, H, W, K, L = c["B"], c["H"], c["W"], c["K"], c["L"]
    model, sd = build_product_model(c["max_disp"], K, L, 0, "reference")
    model = model.cuda()
    img1, img2 = synthetic_pair(B, H, W, c["max_disp"], index=2)
    model({"img1": img1, "img2": img2})                      # builds the plan, fills its input buffers
    plan = next(iter(model._plans.values()))
    taps = {k: v.cpu() for k, v in plan.run_with_taps().items()}
    ocfg = oracle_cfg(c["max_disp"], K, L)
    ref = O.forward(sd, ocfg, img1, img2)
    ot = ocfg.taps
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))
    out = {"config": cfgname, "gemm": gemm, "shape": [B, H, W], "layers": list(L)}
    g = plan.geom
    h8, w8 = plan.h8, plan.w8
    # features as the kernels saw them vs the oracle's CPU features
    # (oracle does not tap the backbone output; recompute it)
    feats = O.backbone_resnet(sd, "backbone", torch.cat([O.pad_images(img1, 8)[0], O.pad_images(img2, 8)[0]], 0))
    f8 = feats[1].chunk(2, 0)[0].permute(0, 2, 3, 1)
    out["rel_err"] = {"features@1/8": rel(plan.f1_8.cpu(), f8)}
    out["rel_err"]["cost_volume"] = rel(taps["cost_volume"], ot["cost_volume"])
    out["abs_err_prob"] = float((taps["prob"] - ot["prob"]).abs().max())
    same_seed = (taps["seeds"] == ot["seeds"]).all(-1)
    out["seed_rows_identical"] = float(same_seed.float().mean())
    pn = ot["prob_nms"]
    out["seed_value_gap_max"] = float((pn.gather(1, taps["seeds"]) - pn.gather(1, ot["seeds"])).abs().max())
    for k in ["prop_embed"] + [f"prop_layer{i}" for i in range(L[0])]:
        out["rel_err"][k] = rel(taps[k][same_seed], ot[k][same_seed])
    lab_err = (taps["labels"] - ot["labels"]).abs()
    out["labels_abs_err_max_same_seed"] = float(lab_err[same_seed].max())
    for i in range(L[1]):
        out["rel_err"][f"inference_layer{i}"] = rel(taps[f"inference_layer{i}"], ot[f"inference_layer{i}"])
    # selection agreement at full resolution (NMRF.py:228)
    Hp8, Wp8, top, left = g["Hp8"], g["Wp8"], g["top8"], g["left8"]
    sc = taps["score"].reshape(B, Hp8, Wp8, K, 64)[:, top:top + h8, left:left + w8]
    sc = sc.reshape(B, h8, w8, K, 8, 8).permute(0, 1, 4, 2, 5, 3).reshape(B, h8 * 8, w8 * 8, K)
    sel = sc.argmax(-1)
    agree = sel == ot["sel"]
    out["selection_agreement"] = float(agree.float().mean())
    dc_err = (taps["disp_curr"] - ot["disp_curr"]).abs()
    blk = agree.reshape(B, 2 * h8, 4, 2 * w8, 4).all(2).all(-1)           # 4x4 median blocks whose 16 selections all agree
    out["median_blocks_all_agree"] = float(blk.float().mean())
    out["disp_curr_abs_err_max_on_agreeing_blocks"] = float(dc_err[blk].max())
    d = (taps["disp"] - ref["disp"]).abs()
    full = blk.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :H, :W]
    # a flipped block also perturbs its window neighbours through the refinement attention: dilate by one 4x4 window (16 px)
    import torch.nn.functional as F
    bad = F.max_pool2d((~blk).float()[:, None], 9, 1, 4)[:, 0] > 0          # +-4 blocks = the 4x4 window and its shifted partner
    clean = (~bad).repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :H, :W]
    out["EPE"] = float(d.mean())
    out["max_err_px"] = float(d.max())
    out["frac_px_err_gt_1e-3"] = float((d > 1e-3).float().mean())
    out["EPE_on_agreeing_blocks"] = float(d[full].mean())
    out["EPE_away_from_flips"] = float(d[clean].mean()) if clean.any() else None
    out["frac_px_away_from_flips"] = float(clean.float().mean())
    out["max_err_away_from_flips"] = float(d[clean].max()) if clean.any() else None
    return out


def oracle_self_conditioning(cfgname, eps=1e-6):
    """CPU-only control: the oracle against itself with relative noise `eps` on the images."""
    import torch
    from helpers import build_product_model, oracle_cfg
    from nmrf_b200.synthetic import synthetic_pair
    from oracle import nmrf_oracle as O
    c = CONFIGS[cfgname]
    _, sd = build_product_model(c["max_disp"], c["K"], c["L"], 0, "reference")
    img1, img2 = synthetic_pair(c["B"], c["H"], c["W"], c["max_disp"], index=2)
    a = O.forward(sd, oracle_cfg(c["max_disp"], c["K"], c["L"], taps=False), img1, img2)
    gen = torch.Generator().manual_seed(1)
    n1 = img1 * (1 + eps * torch.randn(img1.shape, generator=gen))
    n2 = img2 * (1 + eps * torch.randn(img2.shape, generator=gen))
    b = O.forward(sd, oracle_cfg(c["max_disp"], c["K"], c["L"], taps=False), n1, n2)
    d = (a["disp"] - b["disp"]).abs()
    return {"noise": eps, "EPE": float(d.mean()), "max_err_px": float(d.max())}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small")
    ap.add_argument("--gemm", default="tc")
    ap.add_argument("--control", action="store_true", help="also run the CPU oracle-vs-noisy-oracle control")
    args = ap.parse_args()
    r = report(args.config, args.gemm)
    if args.control:
        r["oracle_vs_noisy_oracle"] = [oracle_self_conditioning(args.config, e) for e in (1e-7, 1e-6)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"parity_{args.config}_{args.gemm}.json"), "w") as f:
        json.dump(r, f, indent=1)
    print(json.dumps(r, indent=1))
