"""CUDA-event timing of nmrf_conv2d at the encoder's shapes (run under gpurun):  python tools/conv_bench.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmrf_b200.encoder import _Conv

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda"
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, N, H, W, Cin, Cout, k, stride, pad in [("layer1 64->64 3x3", 2, 272, 480, 64, 64, 3, 1, 1), ("layer2 64->96 3x3 s2", 2, 272, 480, 64, 96, 3, 2, 1),
                                                 ("layer2 96->96 3x3", 2, 136, 240, 96, 96, 3, 1, 1), ("layer3 128->128 3x3", 2, 136, 240, 128, 128, 3, 1, 1),
                                                 ("head8 256->384 3x3", 2, 68, 120, 256, 384, 3, 1, 1), ("head4 256->256 3x3", 2, 136, 240, 256, 256, 3, 1, 1)]:
    x = torch.randn(N, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (k * k * Cin) ** 0.5).to(dev)
    conv = _Conv(w, stride, pad)
    Ho, Wo = conv.out_hw(H, W)
    y = torch.empty(N, Ho, Wo, Cout, device=dev)
    for _ in range(3):
        conv(x, y)
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); conv(x, y); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    fl = 2.0 * N * Ho * Wo * Cout * Cin * k * k
    units = ((N * Ho * Wo + 127) // 128) * ((Cout + 127) // 128) * (k * k * Cin // 32)
    print(f"{name:22s} {us:7.1f} us  {fl / us * 1e-6:6.1f} TFLOP/s   {us * 1e-6 * 1.965e9 * 148 / units:6.0f} cycles/unit/SM")
