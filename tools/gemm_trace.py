"""Cycle-level trace of CTA 0 of the warp-specialised tcgen05 GEMM (run under gpurun).
Stamps (clock64): producer thread 32: 1024 + unit*8 + {0 raw tile visible, 1 split done, 2 A buffer free (done-wait),
3 tcgen05.st issued, 4 handed off, 5 ring re-armed}; MMA lane 0: 2048 + unit*4 + {0 before bar.sync, 1 after, 2 after issue+commit};
epilogue warp 0 lane 0: 3584 + tile*4 + {0 before acc_full wait, 1 after, 2 after drain}."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200 import _lib, ops

dev = "cuda"
g = torch.Generator().manual_seed(0)
out = {}
for name, rows, Kx, Ke, N, ln, act, res in [("fc2", 34560, 512, 0, 128, False, 0, True), ("qkv", 34560, 128, 32, 384, True, 0, False),
                                              ("fc1", 34560, 128, 0, 512, True, 2, False), ("proj", 34560, 128, 0, 128, False, 0, True)]:
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
    prod = [[(t[u * 8 + k] - t0) if t[u * 8 + k] else None for k in range(3)] for u in range(40) if t[u * 8]]
    mma = [[(t[2048 + u * 4 + k] - t0) if t[2048 + u * 4 + k] else None for k in range(3)] for u in range(40) if t[2048 + u * 4]]
    epi = [[(t[3584 + i * 4 + k] - t0) if t[3584 + i * 4 + k] else None for k in range(3)] for i in range(8) if t[3584 + i * 4]]
    out[name] = dict(us=s.elapsed_time(e) * 1e3, producer=prod, mma=mma, epilogue=epi)
    print(name, f"{s.elapsed_time(e)*1e3:.1f} us")
    p2 = [[(t[1024 + u * 8 + k] - t0) if t[1024 + u * 8 + k] else None for k in range(6)] for u in range(24) if t[1024 + u * 8]]
    print("  producer thread 32: raw tile ready at | LDS+LN+split, done-wait, STTM issue, wait::st+handoff, ring re-arm | next unit's raw wait")
    for u, r in enumerate(p2[:12]):
        nxt = (p2[u + 1][0] - r[5]) if u + 1 < len(p2) else None
        print("   u%02d" % u, r[0], "|", r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], "|", nxt)
    print("  mma (before sync | bar.sync, full_b wait, issue | gap to next):")
    m4 = [[(t[2048 + u * 4 + k] - t0) for k in range(4)] for u in range(40) if t[2048 + u * 4]]
    for u, r in enumerate(m4[:28]):
        nxt = (m4[u + 1][0] - r[2]) if u + 1 < len(m4) else None
        print("   u%02d" % u, r[0], "|", r[3] - r[0], r[1] - r[3], r[2] - r[1], "|", nxt)
    print("  epilogue (start, wait, drain):", [(r[0], r[1] - r[0], r[2] - r[1]) for r in epi])
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gemm_trace.json"), "w"))
