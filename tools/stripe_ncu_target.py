"""ncu target: stripe attention at the bench shape (68 x 120 grid, K = 4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200 import ops
g = torch.Generator().manual_seed(0)
B, h, w, K = 1, 68, 120, 4
qkv = torch.randn(B * h * w * K, 384, generator=g).cuda()
gv0, gv1 = (0.2 * torch.randn(64, 1, 3, 3, generator=g)).cuda(), (0.2 * torch.randn(64, 1, 3, 3, generator=g)).cuda()
for _ in range(3):
    ops.stripe_attention(qkv, B, h, w, K, gv0, gv1)
torch.cuda.synchronize()
