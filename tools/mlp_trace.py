"""Cycle trace of CTA 0 of nmrf_mlp_chain (run under gpurun); stamp layout in csrc/gemm_mlp.cu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200 import _lib, ops
dev = "cuda"
g = torch.Generator().manual_seed(0)
rows = 34560
att, x = torch.randn(rows, 128, generator=g).to(dev), torch.randn(rows, 128, generator=g).to(dev)
Wp, W1, W2 = (torch.randn(128, 128, generator=g) / 11).to(dev), (torch.randn(512, 128, generator=g) / 11).to(dev), (torch.randn(128, 512, generator=g) / 22).to(dev)
PRELOAD = os.environ.get('NMRF_B200_RESIDUAL', 'preload') != 'identity'
ws = ops.pack_mlp_stream(Wp.contiguous() if PRELOAD else torch.cat([Wp, torch.eye(128, device=dev)], 1).contiguous(), W1, W2)
z, o, b1 = torch.zeros(128, device=dev), torch.ones(128, device=dev), torch.zeros(512, device=dev)
for _ in range(3):
    ops.mlp_chain(att, ws, z, (o, z), b1, z, E=x, out=x, e_identity=PRELOAD)
tr = torch.zeros(4096, dtype=torch.int64, device=dev)
_lib.check(_lib.lib.nmrf_debug_set_trace(tr.data_ptr()), "set_trace")
ops.mlp_chain(att, ws, z, (o, z), b1, z, E=x, out=x, e_identity=PRELOAD)
torch.cuda.synchronize()
_lib.check(_lib.lib.nmrf_debug_set_trace(None), "set_trace")
t = tr.cpu().tolist()
t0 = min(v for v in t if v > 0)
print("MMA warp, per unit: start | (bar.sync) weight-wait issue | gap")
for u in range(80):
    r = t[u * 4:u * 4 + 4]
    if not r[0]:
        break
    nxt = t[(u + 1) * 4] - r[2] if t[(u + 1) * 4] else None
    kind = "P1" if (u % 40) < 8 else "F"
    print(f"  u{u:02d} {kind} {r[0]-t0:7d} | {('%4d' % (r[3]-r[0])) if r[3] else '    '} {r[1]-(r[3] or r[0]):5d} {r[2]-r[1]:5d} | {nxt}")
print("GELU warp 9, per chunk: start | acc1_full wait, ld+GELU, h_free wait, stores")
for c in range(16):
    r = t[2048 + c * 8:2048 + c * 8 + 5]
    if not r[0]:
        break
    print(f"  c{c:02d} {r[0]-t0:7d} | {r[1]-r[0]:5d} {r[2]-r[1]:5d} {r[3]-r[2]:5d} {r[4]-r[3]:5d}")
for i in range(2):
    r = t[3968 + i * 8:3968 + i * 8 + 4]
    if r[0]:
        print(f"tile {i}: p1_full seen {r[0]-t0}, LN done +{r[1]-r[0]}, acc0_final seen {r[2]-t0}, stored +{r[3]-r[2]}")

print("fine stamps of F1 units (DBG bit 512): try_wait, syncwarp, elect, fence, desc+mma, commits")
for u in range(8, 40):
    r = t[1024 + u * 8:1024 + u * 8 + 7]
    if r[0] and r[6]:
        print(f"  u{u:02d} {r[1]-r[0]:5d} {r[2]-r[1]:5d} {r[3]-r[2]:5d} {r[4]-r[3]:5d} {r[5]-r[4]:5d} {r[6]-r[5]:5d}")
