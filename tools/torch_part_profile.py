"""Kernel-time breakdown of the torch part of the forward (feature extractor + conv heads) at the bench workload (gpurun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import profile, ProfilerActivity
from helpers import build_product_model
from nmrf_b200.synthetic import synthetic_pair

model, sd = build_product_model(192, 4, (8, 8, 8), 0, "reference")
model = model.cuda().eval()
img1, img2 = synthetic_pair(1, 540, 960, 192, index=0)
img1, img2 = img1.cuda(), img2.cuda()
for _ in range(3):
    model.forward_device(img1, img2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.forward_device(img1, img2)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t > 0:
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time {tot/1e3:.2f} ms")
for t, c, k in rows[:40]:
    print(f"{t/1e3:8.3f} ms  x{c:4d}  {k[:110]}")
