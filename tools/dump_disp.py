"""Dump the disparity map (and the propagated proposals) of one seeded forward to gpurun_out/<tag>.pt, to compare two
builds of the library bit by bit:
    NMRF_B200_LIB=nmrf_b200/libA.so python tools/dump_disp.py a;  python tools/dump_disp.py b;  python tools/dump_disp.py cmp a b"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
out = os.path.join(ROOT, "gpurun_out")
if sys.argv[1] == "cmp":
    a, b = (torch.load(os.path.join(out, t + ".pt")) for t in sys.argv[2:4])
    for k in a:
        d = (a[k].double() - b[k].double()).abs()
        print(f"{k:18s} bit-identical: {torch.equal(a[k], b[k])}   max |diff| {float(d.max()):.3e}   differing elements {int((d > 0).sum())} / {d.numel()}")
    sys.exit(0)
from helpers import build_product_model
from nmrf_b200.synthetic import synthetic_pair
res = {}
for name, (B, H, W, L) in dict(c1=(1, 540, 960, (8, 8, 8)), small=(2, 136, 240, (2, 3, 2))).items():
    model, _ = build_product_model(192, 4, L, 0, "reference")
    model = model.cuda()
    i1, i2 = (t.cuda() for t in synthetic_pair(B, H, W, 192, 0))
    o = model.forward_device(i1, i2)
    res[name + "_disp"] = o["disp"].cpu()
    res[name + "_proposal"] = o["proposal"].cpu()
os.makedirs(out, exist_ok=True)
torch.save(res, os.path.join(out, sys.argv[1] + ".pt"))
print("saved", sys.argv[1], {k: tuple(v.shape) for k, v in res.items()})
