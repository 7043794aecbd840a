"""Stand-alone timing of the four token-GEMM shapes of one inference layer (run under gpurun).
    python tools/gemm_bench.py [reps]        (NMRF_B200_DBG / NMRF_B200_GEMM_V select experiment paths)
Prints us per launch (CUDA events around `reps` back-to-back launches, L2-warm like the hot path's graph)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda"
g = torch.Generator().manual_seed(0)
res = []
for name, rows, Kx, Ke, N, ln, act, r in [("qkv", 34560, 128, 32, 384, True, 0, False), ("proj", 34560, 128, 0, 128, False, 0, True),
                                           ("fc1", 34560, 128, 0, 512, True, 2, False), ("fc2", 34560, 512, 0, 128, False, 0, True),
                                           ("pqkv", 32640, 128, 64, 384, True, 0, False)]:
    X = torch.randn(rows, Kx, generator=g).to(dev)
    E = torch.randn(rows, Ke, generator=g).to(dev) if Ke else None
    W = (torch.randn(N, Kx + Ke, generator=g) / (Kx + Ke) ** 0.5).to(dev)
    Wt = ops.pack_weight_tiles(W)
    b = torch.randn(N, generator=g).to(dev)
    gam, bet = torch.ones(Kx, device=dev), torch.zeros(Kx, device=dev)
    R = torch.randn(rows, N, generator=g).to(dev) if r else None
    Y = torch.empty(rows, N, device=dev)
    kw = dict(E=E, ln=(gam, bet) if ln else None, ln_stats=ops.row_stats(X) if ln else None, bias=b, R=R, act=act, Wt=Wt, out=Y)
    for _ in range(3):
        ops.token_gemm(X, W, **kw)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        ops.token_gemm(X, W, **kw)
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / reps
    units = ((rows + 127) // 128) * ((N + 127) // 128) * ((Kx + Ke + 31) // 32)
    res.append(f"{name} {us:6.1f}us ({us * 1.9e3 * 148 / units:5.0f} cyc/unit/SM)")
print(f"DBG={os.environ.get('NMRF_B200_DBG', '0'):>2s} V={os.environ.get('NMRF_B200_GEMM_V', '6')}:  " + "  ".join(res))

# fused block tail (nmrf_mlp_chain): proj + residual + LN2 + fc1 + GELU + fc2 in one launch
for rows in (34560, 32640):
    att, x = torch.randn(rows, 128, generator=g).to(dev), torch.randn(rows, 128, generator=g).to(dev)
    Wp, W1, W2 = (torch.randn(128, 128, generator=g) / 11).to(dev), (torch.randn(512, 128, generator=g) / 11).to(dev), (torch.randn(128, 512, generator=g) / 22).to(dev)
    PRELOAD = os.environ.get('NMRF_B200_RESIDUAL', 'preload') != 'identity'
    ws = ops.pack_mlp_stream(Wp.contiguous() if PRELOAD else torch.cat([Wp, torch.eye(128, device=dev)], 1).contiguous(), W1, W2)
    z, o, b1 = torch.zeros(128, device=dev), torch.ones(128, device=dev), torch.zeros(512, device=dev)
    for _ in range(3):
        ops.mlp_chain(att, ws, z, (o, z), b1, z, E=x, out=x, e_identity=PRELOAD)
    x.normal_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        ops.mlp_chain(att, ws, z, (o, z), b1, z, E=x, out=x, e_identity=PRELOAD)
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / reps
    units = ((rows + 127) // 128) * (36 if PRELOAD else 40)
    print(f"mlp_chain rows={rows}: {us:6.1f}us ({us * 1.9e3 * 148 / units:5.0f} cyc/unit/SM; replaces proj+fc1+fc2)")
