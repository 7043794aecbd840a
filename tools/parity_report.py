"""GPU-side parity report (run under gpurun): the CUDA path and the fp32 reference arithmetic against FLOAT64 truth,
stage by stage, with the decomposition SURVEY.md H2 asks for -- arithmetic error vs discrete-decision flips.

    python tools/parity_report.py [--config small|half|c1|c1b|c2 ...] [--tag NAME]

`c1`, `c1b`, `c2` are the BASELINE configs; for those the committed fixtures tests/golden/truth_*.npz (float64 oracle +
the REAL reference's distance from it, written by oracle/make_golden.py) are reported next to the live numbers.
Dev toggles (A/B in one gpurun call): NMRF_B200_LIB=<path to another build of the library>, NMRF_B200_EXT_LABELS=0.
Prints one JSON object per config and writes gpurun_out/parity_<config>_<tag>.json.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TRUTH_FEATURES = False      # --truth-features: feed the hot path the float64 oracle's feature maps (isolates the encoder)

CONFIGS = {
    "small": dict(B=2, H=136, W=240, max_disp=192, K=4, L=(2, 3, 2), index=2),
    "half": dict(B=1, H=270, W=480, max_disp=192, K=4, L=(5, 5, 5), index=2),
    "c1": dict(B=1, H=540, W=960, max_disp=192, K=4, L=(8, 8, 8), index=0, truth="truth_c1"),
    "c1b": dict(B=1, H=540, W=960, max_disp=192, K=4, L=(5, 5, 5), index=0, truth="truth_c1b"),
    "c2": dict(B=8, H=375, W=1248, max_disp=192, K=4, L=(8, 8, 8), index=0, truth="truth_c2"),
}


def report(cfgname, pairs):
    import numpy as np
    import torch
    from helpers import build_product_model, golden, parity_metrics
    from nmrf_b200.synthetic import synthetic_pair
    c = CONFIGS[cfgname]
    model, sd = build_product_model(c["max_disp"], c["K"], c["L"], 0, "reference")
    model = model.cuda()
    out = {"config": cfgname, "shape": [c["B"], c["H"], c["W"]], "layers": list(c["L"]), "runs": []}
    for i in range(pairs):
        t0 = time.time()
        img1, img2 = synthetic_pair(c["B"], c["H"], c["W"], c["max_disp"], index=c["index"] + i)
        m, o, t = parity_metrics(model, sd, c["max_disp"], c["K"], c["L"], img1, img2, truth_features=TRUTH_FEATURES)
        m["index"] = c["index"] + i
        m["seconds"] = round(time.time() - t0, 1)
        if i == 0 and "truth" in c:
            g = golden(c["truth"])
            d = (o["disp"].double().cpu() - torch.from_numpy(g["disp64"]).double()).abs()
            m["fixture"] = {"EPE_cuda_vs_disp64": float(d.mean()), "max": float(d.max()),
                            "live_truth_vs_fixture_max": float((t["disp"] - torch.from_numpy(g["disp64"]).double()).abs().max()),
                            "real_reference_EPE": float(g["ref32_epe"]), "real_reference_max": float(g["ref32_max"]),
                            "oracle32_selection_flips": int(g["oracle32_selection_flips"]),
                            "bar": max(1e-3, 1.25 * float(g["ref32_epe"]))}
        out["runs"].append(m)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", nargs="+", default=["small"])
    ap.add_argument("--pairs", type=int, default=1, help="consecutive synthetic pair indices to report per config")
    ap.add_argument("--tag", default="default")
    ap.add_argument("--truth-features", action="store_true")
    args = ap.parse_args()
    TRUTH_FEATURES = args.truth_features
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for name in args.config:
        r = report(name, args.pairs)
        r["tag"] = args.tag
        r["env"] = {k: v for k, v in os.environ.items() if k.startswith("NMRF_B200_")}
        with open(os.path.join(ROOT, "gpurun_out", f"parity_{name}_{args.tag}.json"), "w") as f:
            json.dump(r, f, indent=1)
        brief = [{"index": m["index"], "cuda": m["cuda"], "ref32": m["ref32"], "fixture": m.get("fixture"),
                  "stage_rms_last": {k: v for k, v in m["stage"].items() if k.endswith(("7", "4")) or "feat" in k}} for m in r["runs"]]
        print(json.dumps({"config": name, "tag": args.tag, "runs": brief}, indent=1), flush=True)
