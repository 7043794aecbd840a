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
    from helpers import build_product_model, parity_metrics
    from nmrf_b200.synthetic import synthetic_pair
    c = CONFIGS[cfgname]
    model, sd = build_product_model(c["max_disp"], c["K"], c["L"], 0, "reference")
    model = model.cuda()
    img1, img2 = synthetic_pair(c["B"], c["H"], c["W"], c["max_disp"], index=2)
    m, _, _ = parity_metrics(model, sd, c["max_disp"], c["K"], c["L"], img1, img2)
    out = {"config": cfgname, "gemm": gemm, "shape": [c["B"], c["H"], c["W"]], "layers": list(c["L"])}
    out.update(m)
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
