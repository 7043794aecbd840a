"""Cycle-level trace of CTA 0 of nmrf_conv2d (the tcgen05 token-GEMM kernel in CONV mode); needs a TRACE build:
    make -C nmrf_b200/csrc TRACE=1 BUILD=build_tr LIB=../libnmrf_b200_trace.so
    NMRF_B200_LIB=nmrf_b200/libnmrf_b200_trace.so python tools/conv_trace.py
Stamps as in tools/gemm_trace.py."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200 import _lib
from nmrf_b200.encoder import _Conv

dev = "cuda"
g = torch.Generator().manual_seed(0)
out = {}
for name, N, H, W, Cin, Cout in [("l1_64", 2, 272, 480, 64, 64), ("l3_128", 2, 136, 240, 128, 128), ("head4_256", 2, 136, 240, 256, 256)]:
    x = torch.randn(N, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev)
    conv = _Conv(w, 1, 1)
    y = torch.empty(N, H, W, Cout, device=dev)
    for _ in range(3):
        conv(x, y)
    tr = torch.zeros(8192, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib.nmrf_debug_set_trace(tr.data_ptr()), "set_trace")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); conv(x, y); e.record(); torch.cuda.synchronize()
    _lib.check(_lib.lib.nmrf_debug_set_trace(None), "set_trace")
    t = tr.cpu().tolist()
    t0 = min(v for v in t if v > 0)
    print(name, f"{s.elapsed_time(e)*1e3:.1f} us")
    p1 = [[(t[u * 8 + k] - t0) if t[u * 8 + k] else None for k in range(3)] for u in range(40) if t[u * 8]]
    print("  producer thread 0: unit start | to done-wait passed, to hand-off | next start - handoff")
    for u, r in enumerate(p1[:30]):
        nxt = (p1[u + 1][0] - r[2]) if u + 1 < len(p1) else None
        print("   u%02d" % u, r[0], "|", r[1] - r[0], r[2] - r[1], "|", nxt)
    p2 = [[(t[1024 + u * 8 + k] - t0) if t[1024 + u * 8 + k] else None for k in range(6)] for u in range(40) if t[1024 + u * 8]]
    print("  producer thread 32: raw ready at | LDS+split, done-wait, STTM issue, wait::st+handoff |")
    for u, r in enumerate(p2[:30]):
        print("   u%02d" % u, r[0], "|", r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3])
    print("  mma (before sync | bar.sync, full_b wait, issue | gap to next):")
    m4 = [[(t[2048 + u * 4 + k] - t0) for k in range(4)] for u in range(60) if t[2048 + u * 4]]
    for u, r in enumerate(m4[:40]):
        nxt = (m4[u + 1][0] - r[2]) if u + 1 < len(m4) else None
        print("   u%02d" % u, r[0], "|", r[3] - r[0], r[1] - r[3], r[2] - r[1], "|", nxt)
    epi = [[(t[3584 + i * 4 + k] - t0) if t[3584 + i * 4 + k] else None for k in range(3)] for i in range(8) if t[3584 + i * 4]]
    print("  epilogue (start, wait+drain groups, store):", [(r[0], r[1] - r[0], r[2] - r[1]) for r in epi])
