"""GPU-side (run under gpurun): MSDeformAttn forward -- libnmrf_b200's kernel against the INCUMBENT, the reference's own
CUDA kernel (ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299) JIT-compiled for sm_100 from baseline/_ref/ops/src on this
box ("the bar is this kernel recompiled for sm_100"), at the shapes of BASELINE config 5 (Swin-T neck,
SURVEY.md App. B: N = 2 pairs-sides, Lq = (Hp/4)(Wp/4), 8 heads x 8 channels, 1 level, 4 points, value strides 4/8/16/32).

    python tools/msda_bench.py [--H 1024 --W 1504]   ->  gpurun_out/msda_bench.json
Reports per call: us (CUDA events on the launching stream, L2 flushed between launches), algorithmic GB/s
(4 N (64 S + 224 Lq) bytes) against MEASURED_PEAKS.json, max |difference| of the two kernels' outputs and of
each against the float64 oracle.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def build_incumbent():
    """the reference extension, unmodified sources, JIT-built with torch.utils.cpp_extension (ninja + nvcc on the box)"""
    src = os.path.join(ROOT, "baseline", "_ref", "ops", "src")
    if not os.path.isdir(src):
        return None, "baseline/_ref/ops/src not staged (python baseline/stage_reference.py in the build container)"
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    try:
        mod = load(name="MultiScaleDeformableAttention_ref",
                   sources=[os.path.join(src, "vision.cpp"), os.path.join(src, "cpu", "ms_deform_attn_cpu.cpp"),
                            os.path.join(src, "cuda", "ms_deform_attn_cuda.cu")],
                   extra_include_paths=[src], extra_cflags=["-DWITH_CUDA"],
                   extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                      "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"],
                   build_directory=None, verbose=False)
        return mod, None
    except Exception as e:          # deprecated ATen APIs may stop compiling one day: report, do not fake
        return None, f"reference extension failed to build: {str(e)[-400:]}"


def time_us(fn, flush, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=1024)
    ap.add_argument("--W", type=int, default=1504)
    ap.add_argument("--pairs", type=int, default=1)
    args = ap.parse_args()
    import nmrf_b200.msda as msda
    from oracle import nmrf_oracle as O
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    ref_mod, why = build_incumbent()
    dev = "cuda"
    flush = torch.empty(64 << 20, device=dev)
    N, M, Dh, P = 2 * args.pairs, 8, 8, 4
    Lq = (args.H // 4) * (args.W // 4)
    g = torch.Generator().manual_seed(0)
    out = {"shape": {"N": N, "Lq": Lq, "M": M, "Dh": Dh, "P": P, "image": [args.H, args.W]}, "incumbent": "built" if ref_mod else why,
           "peak_hbm_gbs": peaks["hbm_gbs"], "levels": []}
    for stride in (4, 8, 16, 32):
        h, w = args.H // stride, args.W // stride
        S = h * w
        shp = torch.tensor([[h, w]], dtype=torch.long, device=dev)
        st = torch.zeros(1, dtype=torch.long, device=dev)
        value = torch.randn(N, S, M, Dh, generator=g).to(dev)
        # sampling locations as the neck produces them: reference point of the query + small learned offsets
        ys, xs = torch.meshgrid(torch.linspace(0.5 / (args.H // 4), 1 - 0.5 / (args.H // 4), args.H // 4),
                                torch.linspace(0.5 / (args.W // 4), 1 - 0.5 / (args.W // 4), args.W // 4), indexing="ij")
        refp = torch.stack([xs, ys], -1).reshape(1, Lq, 1, 1, 1, 2)
        loc = (refp + 0.02 * torch.randn(N, Lq, M, 1, P, 2, generator=g)).to(dev).contiguous()
        w_ = torch.softmax(torch.randn(N, Lq, M, P, generator=g), -1).reshape(N, Lq, M, 1, P).to(dev).contiguous()
        ours = lambda: msda.ms_deform_attn_forward(value, shp, st, loc, w_, 64)
        o = ours()
        bytes_alg = 4.0 * N * (64 * S + 224 * Lq)
        row = {"value_stride": stride, "S": S, "algorithmic_MB": bytes_alg / 1e6}
        t = time_us(ours, flush)
        row["nmrf_b200"] = {"us": t, "GBs": bytes_alg / t / 1e3, "frac_hbm": bytes_alg / t / 1e3 / peaks["hbm_gbs"]}
        if ref_mod is not None:
            theirs = lambda: ref_mod.ms_deform_attn_forward(value, shp, st, loc, w_, 64)
            r = theirs()
            t2 = time_us(theirs, flush)
            row["reference_sm100"] = {"us": t2, "GBs": bytes_alg / t2 / 1e3, "frac_hbm": bytes_alg / t2 / 1e3 / peaks["hbm_gbs"]}
            row["speedup"] = t2 / t
            row["max_abs_diff_vs_reference_kernel"] = float((o - r).abs().max())
        if stride >= 16:                                    # float64 oracle on the device for the small levels
            o64 = O.ms_deform_attn(value.double(), shp, st, loc.double(), w_.double())
            row["max_abs_err_vs_float64"] = float((o.double() - o64).abs().max())
        out["levels"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "msda_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
