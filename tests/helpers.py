"""Shared test utilities: golden fixtures, synthetic weights, oracle configs."""
import os

import numpy as np
import torch

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


def parity_metrics(model, sd, max_disp, K, L, img1, img2):
    """CUDA path vs CPU oracle on the same inputs, with the decomposition SURVEY.md H2 asks for: arithmetic
    error at every stage boundary, agreement of the discrete decisions (seeds, argmax-over-K selection), and
    the end-point error overall / away from flipped decisions.  Returns (metrics dict, model output dict)."""
    import torch.nn.functional as F
    from oracle import nmrf_oracle as O
    B, _, H, W = img1.shape
    out = model({"img1": img1, "img2": img2})              # builds the plan, fills its input buffers
    plan = model.plan_for(B, plan_C(model), *feat_hw(model, H, W), H, W)
    taps = {k: v.cpu() for k, v in plan.run_with_taps().items()}
    ocfg = oracle_cfg(max_disp, K, L)
    ref = O.forward(sd, ocfg, img1, img2)
    ot = ocfg.taps
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))
    g, h8, w8 = plan.geom, plan.h8, plan.w8
    m = {"rel_err": {}}
    feats = O.backbone_resnet(sd, "backbone", torch.cat([O.pad_images(img1, 8)[0], O.pad_images(img2, 8)[0]], 0))
    m["rel_err"]["features@1/8"] = rel(plan.f1_8.cpu(), feats[1].chunk(2, 0)[0].permute(0, 2, 3, 1))
    m["rel_err"]["cost_volume"] = rel(taps["cost_volume"], ot["cost_volume"])
    m["abs_err_prob"] = float((taps["prob"] - ot["prob"]).abs().max())
    same_seed = (taps["seeds"] == ot["seeds"]).all(-1)
    m["seed_rows_identical"] = float(same_seed.float().mean())
    pn = ot["prob_nms"]
    m["seed_value_gap_max"] = float((pn.gather(1, taps["seeds"]) - pn.gather(1, ot["seeds"])).abs().max())
    for k in ["prop_embed"] + [f"prop_layer{i}" for i in range(int(L[0]))]:
        m["rel_err"][k] = rel(taps[k][same_seed], ot[k][same_seed])
    m["labels_abs_err_max_same_seed"] = float((taps["labels"] - ot["labels"]).abs()[same_seed].max())
    for i in range(int(L[1])):
        m["rel_err"][f"inference_layer{i}"] = rel(taps[f"inference_layer{i}"], ot[f"inference_layer{i}"])
    Hp8, Wp8, top, left = g["Hp8"], g["Wp8"], g["top8"], g["left8"]
    sc = taps["score"].reshape(B, Hp8, Wp8, K, 64)[:, top:top + h8, left:left + w8]
    sc = sc.reshape(B, h8, w8, K, 8, 8).permute(0, 1, 4, 2, 5, 3).reshape(B, h8 * 8, w8 * 8, K)
    agree = sc.argmax(-1) == ot["sel"]
    m["selection_agreement"] = float(agree.float().mean())
    blk = agree.reshape(B, 2 * h8, 4, 2 * w8, 4).all(2).all(-1)            # 4x4 median blocks with all 16 selections agreeing
    pix_seed = same_seed.reshape(B, h8, w8).repeat_interleave(2, 1).repeat_interleave(2, 2)
    blk = blk & pix_seed
    m["median_blocks_all_agree"] = float(blk.float().mean())
    dc = (taps["disp_curr"] - ot["disp_curr"]).abs()
    m["disp_curr_abs_err_max_on_agreeing_blocks"] = float(dc[blk].max()) if blk.any() else None
    d = (out["disp"].cpu() - ref["disp"]).abs()
    # a flipped block perturbs its neighbours through the (shifted) 4x4-window refinement attention and, via the
    # 6x6-window inference attention, its 1/8-res neighbourhood: exclude +-6 blocks (24 px) around every flip
    bad = F.max_pool2d((~blk).float()[:, None], 13, 1, 6)[:, 0] > 0
    clean = (~bad).repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :H, :W]
    # refinement tokens (padded 1/4 grid) compared away from flips only: next to a flip they legitimately differ
    Hp4, Wp4, t4, l4 = g["Hp4"], g["Wp4"], g["top4"], g["left4"]
    for i in range(int(L[2])):
        a = taps[f"refinement_layer{i}"].reshape(B, Hp4, Wp4, 128)[:, t4:t4 + 2 * h8, l4:l4 + 2 * w8]
        b = ot[f"refinement_layer{i}"].reshape(B, Hp4, Wp4, 128)[:, t4:t4 + 2 * h8, l4:l4 + 2 * w8]
        m["rel_err"][f"refinement_layer{i}"] = rel(a[~bad], b[~bad]) if (~bad).any() else 0.0
    m["EPE"] = float(d.mean())
    m["max_err_px"] = float(d.max())
    m["frac_px_err_gt_1e-3"] = float((d > 1e-3).float().mean())
    m["frac_px_away_from_flips"] = float(clean.float().mean())
    m["EPE_away_from_flips"] = float(d[clean].mean()) if clean.any() else None
    m["max_err_away_from_flips"] = float(d[clean].max()) if clean.any() else None
    m["proposal_EPE"] = float((out["proposal"].cpu() - ref["proposal"]).abs().mean())
    return m, out, ref


def plan_C(model):
    enc = model.backbone if model.compat else model.image_encoder
    return enc.output_dim


def feat_hw(model, H, W):
    d = model.divis_by
    Hp, Wp = H + (((H // d) + 1) * d - H) % d, W + (((W // d) + 1) * d - W) % d
    return Hp // 8, Wp // 8
