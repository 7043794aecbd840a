"""ncu target: three launches of nmrf_mlp_chain at the bench shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200 import ops
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
torch.cuda.synchronize()
