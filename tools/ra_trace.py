"""Cycle-level trace of CTA 0 of the resident-A token GEMM (gemm_ra.cu); needs a TRACE build:
    make -C nmrf_b200/csrc TRACE=1 BUILD=build_tr LIB=../libnmrf_b200_trace.so
    NMRF_B200_LIB=nmrf_b200/libnmrf_b200_trace.so python tools/ra_trace.py
Stamps (clock64, CTA 0): producer thread 0: u*4 + {0 unit start, 1 raw landed + next fetch issued, 2 split done, 3 A k-block handed over};
MMA issuer i: 256 + 512 i + u*4 + {0 unit start, 1 accumulator + A ready, 2 weights landed, 3 issued}; TMA lane: 1280 + u*2 + {0, 1 slot free};
epilogue warp 0: 1536 + g*4 + {0 wait start, 1 group complete, 2 drained, 3 chunk stored}."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200 import _lib, ops

dev = "cuda"
g = torch.Generator().manual_seed(0)
for name, rows, Kx, Ke, N, ln, act, res in [("qkv", 34560, 128, 32, 384, True, 0, False), ("proj", 34560, 128, 0, 128, False, 0, True)]:
    X = torch.randn(rows, Kx, generator=g).to(dev)
    E = torch.randn(rows, Ke, generator=g).to(dev) if Ke else None
    W = (torch.randn(N, Kx + Ke, generator=g) / (Kx + Ke) ** 0.5).to(dev)
    Wt = ops.pack_weight_tiles(W)
    b = torch.randn(N, generator=g).to(dev)
    gam, bet = torch.ones(Kx, device=dev), torch.zeros(Kx, device=dev)
    R = torch.randn(rows, N, generator=g).to(dev) if res else None
    kw = dict(E=E, ln=(gam, bet) if ln else None, ln_stats=ops.row_stats(X) if ln else None, bias=b, R=R, act=act, Wt=Wt)
    for _ in range(3):
        ops.token_gemm(X, W, **kw)
    tr = torch.zeros(4096, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib.nmrf_debug_set_trace(tr.data_ptr()), "set_trace")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.token_gemm(X, W, **kw); e.record(); torch.cuda.synchronize()
    _lib.check(_lib.lib.nmrf_debug_set_trace(None), "set_trace")
    t = tr.cpu().tolist()
    t0 = min(v for v in t if v > 0)
    rel = lambda i: (t[i] - t0) if t[i] else None
    print(f"== {name}: {s.elapsed_time(e) * 1e3:.1f} us (traced)")
    print(" producer thread 0 (unit: start | wait+fetch, split, a_free+st+handoff):")
    for u in range(12):
        r = [rel(u * 4 + k) for k in range(4)]
        if r[0] is None: break
        print(f"   u{u:02d} {r[0]:7d} | {r[1]-r[0]:5d} {r[2]-r[1]:5d} {r[3]-r[2]:5d}")
    for i in range(2):
        print(f" MMA issuer {i} (unit: start | acc/A wait, weight wait, issue):")
        n = 0
        for u in range(128):
            r = [rel(256 + 512 * i + u * 4 + k) for k in range(4)]
            if r[0] is None: continue
            print(f"   u{u:03d} {r[0]:7d} | {r[1]-r[0]:5d} {r[2]-r[1]:5d} {r[3]-r[2]:5d}")
            n += 1
            if n >= 34: break
    print(" TMA lane (unit: start | slot wait):")
    print("   ", [(rel(1280 + u * 2), rel(1280 + u * 2 + 1) - rel(1280 + u * 2)) for u in range(64) if rel(1280 + u * 2) is not None][:64])
    print(" epilogue warp 0 (group: wait start | wait, drain | stored):")
    for gi in range(28):
        r = [rel(1536 + gi * 4 + k) for k in range(4)]
        if r[0] is None: break
        print(f"   g{gi:02d} {r[0]:7d} | {r[1]-r[0]:6d} {r[2]-r[1]:5d} | {'' if r[3] is None else r[3]-r[2]}")
