"""Window / stripe attention kernel timings at the bench workload's shapes, SIMT vs tensor-core paths (run under gpurun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200 import _lib, ops

def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3

g = torch.Generator().manual_seed(0)
for name, (B, Hp, Wp, K, ws, shift, se) in {"infer 1/8": (1, 72, 120, 4, 6, 3, True), "refine 1/4": (1, 136, 240, 1, 4, 2, False)}.items():
    qkv = torch.randn(B * Hp * Wp * K, 384, generator=g).cuda()
    table = (0.5 * torch.randn((2 * ws - 1) ** 2, 384, generator=g)).cuda()
    res = {}
    for impl in (0, 1):
        _lib.check(_lib.lib.nmrf_set_attention_impl(impl), "impl")
        out = ops.window_attention(qkv, table, B, Hp, Wp, K, ws, shift, se)
        res[impl] = (timeit(lambda: ops.window_attention(qkv, table, B, Hp, Wp, K, ws, shift, se)), out)
    d = (res[0][1] - res[1][1]).abs().max().item() / res[0][1].abs().max().item()
    print(f"window {name}: simt {res[0][0]:.1f} us, mma {res[1][0]:.1f} us, max rel diff {d:.2e}")
_lib.check(_lib.lib.nmrf_set_attention_impl(1), "impl")

# stripe attention (propagation stack): 68 x 120 grid, K = 4
B, h, w, K = 1, 68, 120, 4
qkv = torch.randn(B * h * w * K, 384, generator=g).cuda()
gv0, gv1 = (0.2 * torch.randn(64, 1, 3, 3, generator=g)).cuda(), (0.2 * torch.randn(64, 1, 3, 3, generator=g)).cuda()
res = {}
for impl in (0, 1):
    _lib.check(_lib.lib.nmrf_set_attention_impl(impl), "impl")
    out = ops.stripe_attention(qkv, B, h, w, K, gv0, gv1)
    res[impl] = (timeit(lambda: ops.stripe_attention(qkv, B, h, w, K, gv0, gv1)), out)
d = (res[0][1] - res[1][1]).abs().max().item() / res[0][1].abs().max().item()
print(f"stripe 68x120 K=4: simt {res[0][0]:.1f} us, tcgen05 {res[1][0]:.1f} us, max rel diff {d:.2e}")
_lib.check(_lib.lib.nmrf_set_attention_impl(1), "impl")

# intra-pixel proposal attention (inference stack): 72 x 120 padded grid, K = 4
P, K = 72 * 120, 4
qkv = torch.randn(P * K, 384, generator=g).cuda()
print(f"proposal attention P={P} K={K}: {timeit(lambda: ops.proposal_attention(qkv, K)):.1f} us")
