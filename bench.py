#!/usr/bin/env python
"""bench.py -- stereo pairs/s of the NMRF-Stereo inference path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c1|c1b|c2|c3|c4]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU)

Workload (config.workload), default c1 = BASELINE.json configs[1] -- SceneFlow 540x960 (padded 544x960), D_max=192 (D=24),
K=4 proposals, 8/8/8 layers ("8 iters"), batch 1 per GPU, fp32, synthetic images and seeded random-init
weights (no datasets/checkpoints offline).  A step = one forward over one batch.  --config selects the other BASELINE
configurations: c1b = c1 at the checkpoint-compatible depth 5/5/5; c2 = KITTI 375x1248, batch 8; c3 =
SceneFlow batch 8 per rank (run with --gpus 4: batch 32 over 4 GPUs); c4 = Swin-T encoder (the reference's own module from
baseline/_ref, with nmrf_b200.msda as its deformable-attention op), 1000x1500, D_max=256, one pair per rank (--gpus 8).

  value  : pairs/s of the WHOLE forward (= the reference's NMRF.forward: feature extractor + conv heads + hot path, every
           launch a libnmrf_b200 kernel, one CUDA graph) over images resident in HBM -- the same scope as the reference arm;
           CUDA-event time per step, L2 flushed between steps, max over ranks.  `hot_path` reports the §8(a) rows alone (cost
           volume ... disparity over resident feature maps) with the per-kernel breakdown.
  e2e    : the whole forward through the public streaming API (`GraphedNMRF.stream`) with pinned HOST images: H2D of both
           images and D2H of the disparity map of every step inside the timed region, overlapped with the neighbouring
           steps' compute (the number to hold against `--impl reference`).
  roofline: dominant kernel of the hot path (per-launch CUDA events, eager) against MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the reference's CPU forward on the host cores -- the UNMODIFIED reference itself
           (kind "reference") when its staged copy baseline/_ref is present (baseline/stage_reference.py; git-ignored, it
           travels to the GPU box), else its oracle port (kind "port"); the thread count is calibrated, rank 0 only,
           a bounded number of forwards.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    "c1": dict(name="sceneflow_540x960_D192_K4_L8", B=1, H=540, W=960, max_disp=192, K=4, L=(8, 8, 8),
               metric="stereo pairs/sec at 960x540 D192 K4 8-iter"),
    "c1b": dict(name="sceneflow_540x960_D192_K4_L5", B=1, H=540, W=960, max_disp=192, K=4, L=(5, 5, 5),
                metric="stereo pairs/sec at 960x540 D192 K4 5/5/5 layers (checkpoint-compatible depth)"),
    "c2": dict(name="kitti_375x1248_D192_K4_L8_batch8", B=8, H=375, W=1248, max_disp=192, K=4, L=(8, 8, 8),
               metric="stereo pairs/sec at KITTI 1248x375 D192 K4 8-iter, batch 8"),
    "c3": dict(name="sceneflow_540x960_D192_K4_L8_batch8_per_gpu", B=8, H=540, W=960, max_disp=192, K=4, L=(8, 8, 8),
               metric="stereo pairs/sec at 960x540 D192 K4 8-iter, batch 8 per GPU (batch 32 on 4 GPUs)"),
    "c4": dict(name="swint_1000x1500_D256_K4_L8", B=1, H=1000, W=1500, max_disp=256, K=4, L=(8, 8, 8), swin=True,
               metric="stereo pairs/sec at 1500x1000 D256 K4 8-iter, Swin-T encoder, 1 pair per GPU"),
}
WORKLOAD = WORKLOADS["c1"]
METRIC = WORKLOAD["metric"]
N_PAIRS = 4                 # distinct synthetic pairs rotated through the timed steps
FLUSH_BYTES = 256 << 20     # > 126 MB L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, tf=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def reference_tree_on_path():
    """the staged reference (baseline/_ref, written by baseline/stage_reference.py) for the Swin-T encoder of config c4"""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "nmrf")) and ref not in sys.path:
        sys.path.insert(0, ref)
    return os.path.isdir(os.path.join(ref, "nmrf"))


def build_model(device):
    import nmrf_b200
    from nmrf_b200.synthetic import synthetic_state_dict
    w = WORKLOAD
    cfg = nmrf_b200.get_cfg()
    cfg.DPN.MAX_DISP, cfg.DPN.NUM_PROPOSALS = w["max_disp"], w["K"]
    cfg.NMP.NUM_PROP_LAYERS, cfg.NMP.NUM_INFER_LAYERS, cfg.NMP.NUM_REFINE_LAYERS = w["L"]
    if w.get("swin"):                                  # configs/sceneflow_swint.yaml:3-7
        cfg.BACKBONE.MODEL_TYPE, cfg.BACKBONE.OUT_CHANNELS, cfg.BACKBONE.DROP_PATH, cfg.DATASETS.DIVIS_BY = "swin", 128, 0.0, 32
        cfg.BACKBONE.COMPAT = False
        reference_tree_on_path()
    model = nmrf_b200.build_model(cfg).eval()
    sd = synthetic_state_dict(model.state_dict(), 0, "reference")
    model.load_state_dict(sd)
    return (model.to(device) if device is not None else model), sd


def oracle_forward_fn(sd):
    from oracle import nmrf_oracle as O
    w = WORKLOAD
    cfg = O.OracleConfig(max_disp=w["max_disp"], num_proposals=w["K"], num_prop_layers=w["L"][0],
                         num_infer_layers=w["L"][1], num_refine_layers=w["L"][2], taps=None)
    return lambda a, b: O.forward(sd, cfg, a, b)


def pick_cpu_threads():
    """all host cores the reference's PyTorch-CPU forward can actually use: intra-op parallelism over the small
    per-window / per-stripe tensors of this model stops scaling (and then collapses) well below a 128-core host,
    so the thread count is calibrated once on a short forward and the best one is used and reported."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    if len(cands) == 1:
        return cands[0]
    from nmrf_b200.synthetic import synthetic_pair
    from oracle import nmrf_oracle as O
    import nmrf_b200
    cfg = nmrf_b200.get_cfg()
    cfg.DPN.MAX_DISP, cfg.NMP.NUM_PROP_LAYERS, cfg.NMP.NUM_INFER_LAYERS, cfg.NMP.NUM_REFINE_LAYERS = 192, 1, 1, 1
    from nmrf_b200.synthetic import synthetic_state_dict
    sd = synthetic_state_dict(nmrf_b200.build_model(cfg).state_dict(), 0, "reference")
    ocfg = O.OracleConfig(max_disp=192, num_proposals=4, num_prop_layers=1, num_infer_layers=1, num_refine_layers=1, taps=None)
    a, b = synthetic_pair(1, 272, 480, 192, 0)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        O.forward(sd, ocfg, a, b)
        t0 = time.perf_counter()
        O.forward(sd, ocfg, a, b)
        t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = c, t
    return best


def time_cpu(fn, pairs, steps, warmup):
    torch.set_num_threads(pick_cpu_threads())
    for i in range(warmup):
        fn(*pairs[i % len(pairs)])
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        fn(*pairs[i % len(pairs)])
        ts.append(time.perf_counter() - t0)
    return ts


def reference_forward_fn(sd):
    """The UNMODIFIED reference on the CPU through its own `NMRF.forward` (nmrf/models/NMRF.py:189-262), from the staged copy
    baseline/_ref (baseline/stage_reference.py; git-ignored, travels with gpurun).  None when the copy is absent or does not
    import here (then the oracle port is timed instead)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if WORKLOAD.get("swin") or not os.path.isdir(os.path.join(ref, "nmrf", "models")):
        return None
    try:
        os.environ["NMRF_REFERENCE_ROOT"] = ref
        from oracle import ref_shims
        w = WORKLOAD
        model = ref_shims.build_reference_model(max_disp=w["max_disp"], num_proposals=w["K"], num_prop_layers=w["L"][0],
                                                num_infer_layers=w["L"][1], num_refine_layers=w["L"][2])
        model.load_state_dict(sd)
        model.eval()
    except Exception as e:                       # noqa: BLE001  (report, fall back to the port)
        print(f"[bench] staged reference not usable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
        return None

    def fn(a, b):
        with torch.no_grad():
            return model({"img1": a, "img2": b})["disp"]
    return fn


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU forward on the host cores, rank 0 only: the staged reference itself when
    baseline/_ref is there, else its oracle port."""
    if rank != 0:
        return
    from nmrf_b200.synthetic import synthetic_pair
    w = WORKLOAD
    if w.get("swin"):
        print(json.dumps({"impl": "reference", "metric": METRIC, "config": {"workload": w["name"]},
                          "unavailable": "the Swin-T encoder's CPU forward needs timm weights/ops that are not in this image"}))
        return
    _, sd = build_model(None)
    pairs = [synthetic_pair(w["B"], w["H"], w["W"], w["max_disp"], i) for i in range(2)]
    fn = reference_forward_fn(sd)
    kind = "reference" if fn is not None else "port"
    what = ("the staged reference's NMRF.forward (baseline/_ref), torch CPU fp32" if fn is not None
            else "oracle port of NMRF.forward, torch CPU fp32")
    ts = time_cpu(fn or oracle_forward_fn(sd), pairs, args.steps, max(args.warmup, 1))
    total = sum(ts)
    val = w["B"] * len(ts) / total
    cores = torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "batch": w["B"], "parallelism": "cpu"},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": kind,
                         "sample": f"{len(ts)} whole forwards of {w['B']} pair(s) ({what})"},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def kernel_report(plan):
    """per-launch CUDA-event times of the hot path (eager), aggregated per C entry point"""
    rows = plan.launches.run_timed(reps=3)
    agg = {}
    for what, sym, ms, fl, by in rows:
        a = agg.setdefault(sym, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        a["ms"] += ms; a["flops"] += fl; a["bytes"] += by; a["launches"] += 1
    total = sum(a["ms"] for a in agg.values())
    return agg, total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c1", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global WORKLOAD, METRIC
    WORKLOAD = WORKLOADS[args.config]
    METRIC = WORKLOAD["metric"]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if WORKLOAD.get("swin"):
        missing = None if reference_tree_on_path() else "baseline/_ref (the staged reference tree: baseline/stage_reference.py) is absent"
        if missing:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "config": {"workload": WORKLOAD["name"]}, "unavailable": missing}))
            return

    from nmrf_b200 import _lib
    from nmrf_b200.runner import GraphedNMRF
    from nmrf_b200.sharding import gather_stats
    from nmrf_b200.synthetic import synthetic_pair
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False            # the 1e-3 EPE bar needs exact-fp32 convolutions too
    torch.backends.cuda.matmul.allow_tf32 = False
    w = WORKLOAD
    B, H, W, K, steps, warmup = w["B"], w["H"], w["W"], w["K"], args.steps, max(args.warmup, 3)
    model, sd = build_model(dev)
    runner = GraphedNMRF(model, B, H, W, graph_full=not w.get("swin"))
    plan = runner.plan
    # distinct pairs per rank (weak scaling: every rank processes its own pairs)
    host = [tuple(t.pin_memory() for t in synthetic_pair(B, H, W, w["max_disp"], rank * N_PAIRS + i)) for i in range(N_PAIRS)]
    devp = [(a.to(dev), b.to(dev)) for a, b in host]
    flush = torch.empty(FLUSH_BYTES // 4, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, prep_fn=None):
        """K steps, each bracketed by its own CUDA-event pair on the launching stream; L2 flushed before each"""
        marks = []
        barrier()
        for i in range(steps):
            if prep_fn is not None:
                prep_fn(i)
            flush.zero_()                               # evict L2 between timed steps
            s, e = ev(), ev()
            s.record(); step_fn(i); e.record()
            marks.append((s, e))
        barrier()
        return sum(s.elapsed_time(e) for s, e in marks) / 1e3

    def load(i):
        a, b = devp[i % N_PAIRS]
        runner.img1.copy_(a); runner.img2.copy_(b)

    for i in range(warmup):
        load(i); runner.replay(); runner.replay_hot_path()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    # ---- "value": one pass of the hot path (libnmrf_b200 kernels, one CUDA graph) over feature maps resident in HBM
    t_hot = timed(lambda i: runner.replay_hot_path())
    # ---- whole forward, device-resident images (torch feature extractor + hot path, one CUDA graph) -----------------
    t_dev = timed(lambda i: runner.replay(), load)
    # ---- "e2e": public streaming API (GraphedNMRF.stream): pinned HOST images in, pinned HOST disparity out, every step's H2D
    # and D2H inside the timed region (they overlap the neighbouring steps' compute on a copy stream); ONE event pair around
    # the whole loop, the L2 flushes included
    def e2e_loop(n):
        barrier()
        s, e = ev(), ev()
        s.record()
        got = 0
        for _ in runner.stream(host[i % N_PAIRS] for i in range(n)):
            got += 1
            flush.zero_()
        e.record()
        barrier()
        assert got == n
        return s.elapsed_time(e) / 1e3
    e2e_loop(2)
    t_e2e = e2e_loop(steps)
    clocks = sampler.stop()
    launches_per_step = plan.num_launches                 # hot path only
    n0 = _lib.launch_count()                              # the whole forward, counted by the library itself (eager call)
    model.forward_device(runner.img1, runner.img2)
    torch.cuda.synchronize(dev)
    fwd_launches = _lib.launch_count() - n0

    # ---- roofline of the dominant hot-path kernel (eager, per-launch events) ------------------------
    agg, hot_ms = kernel_report(plan)
    pk = peaks()
    dom = max(agg, key=lambda k: agg[k]["ms"])
    d = agg[dom]
    ai = d["flops"] / max(d["bytes"], 1.0)
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/r2_ncu_traffic.json; c1 only)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        traffic = tj.get(dom, {}).get("dram_bytes_per_launch") if args.config == "c1" else None
    except Exception:
        pass
    # the path's arithmetic is 3xTF32 (the 1e-3 EPE bar): its ceiling is 1/6 of the dense bf16 peak.  A kernel
    # whose algorithmic intensity puts its HBM roofline above that ceiling is tensor-bound.
    if ai * pk["hbm_gbs"] * 1e9 > pk["tf"] * 1e12 / 6.0:
        roof = {"kernel": dom, "bound": "tensor", "achieved": d["flops"] / (d["ms"] * 1e-3) / 1e12, "peak": pk["tf"],
                "unit": "TFLOP/s", "traffic": traffic,
                "note": f"achieved = algorithmic 2*MAC flops / per-launch CUDA-event time; peak = {pk['source']} bf16 cuBLAS burst; "
                        "the arithmetic is error-compensated 3xTF32 (fp32-accurate, required by the 1e-3 EPE bar): ceiling = peak/6",
                "frac_of_3xtf32_ceiling": d["flops"] / (d["ms"] * 1e-3) / 1e12 / (pk["tf"] / 6.0)}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["bytes"] / (d["ms"] * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "traffic": traffic, "note": f"peak = {pk['source']} copy bandwidth"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["launches"], roof["avg_launch_us"] = d["launches"], 1e3 * d["ms"] / d["launches"]
    roof["share_of_hot_path"] = d["ms"] / hot_ms
    roof["scope"] = ("dominant kernel of the hot path (SURVEY.md §8(a) rows), timed eagerly per launch; the feature extractor's "
                     "convolutions run on the same tcgen05 3xTF32 GEMM kernel (nmrf_conv2d: profiles/r2_conv_bench.txt)")
    kernels = {k: {"ms": round(v["ms"], 4), "share": round(v["ms"] / hot_ms, 4), "launches": v["launches"],
                   "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 3) if v["flops"] else None,
                   "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] else None} for k, v in agg.items()}

    # ---- max over ranks ------------------------------------------------------------------------------
    allv = gather_stats(torch.tensor([t_dev, t_e2e, float(B * steps), t_hot], dtype=torch.float64), device=dev)
    t_dev_max, t_e2e_max, pairs_total = float(allv[:, 0].max()), float(allv[:, 1].max()), float(allv[:, 2].sum())
    t_hot_max = float(allv[:, 3].max())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pairs = [synthetic_pair(B, H, W, w["max_disp"], i) for i in range(2)]
        sd_cpu = {k: v.cpu() for k, v in sd.items()}
        fn = reference_forward_fn(sd_cpu)
        ts = time_cpu(fn or oracle_forward_fn(sd_cpu), pairs, 3, 1)
        cpu = {"value": B * len(ts) / sum(ts), "unit": "pairs/s", "cores": torch.get_num_threads(),
               "kind": "reference" if fn is not None else "port",
               "sample": f"{len(ts)} whole forwards of {B} pair(s) after 1 warm-up (" +
                         ("the staged reference's NMRF.forward, baseline/_ref" if fn is not None else "oracle port of NMRF.forward") +
                         ", torch CPU fp32)"}
    if rank == 0:
        img_bytes = 2 * B * 3 * H * W * 4
        print(json.dumps({
            "metric": METRIC, "value": pairs_total / t_dev_max, "unit": "pairs/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * t_dev_max / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "batch_per_gpu": B, "parallelism": f"dp{world} (independent pairs per rank)",
                       "step": "one whole forward = the reference's NMRF.forward (feature extractor + conv heads + hot path "
                               f"A1-A13: {fwd_launches} libnmrf_b200 kernels" + (" + the reference's Swin-T encoder, eager" if w.get("swin") else
                                                                                 ", one CUDA graph") + ") over images resident in HBM",
                       "l2": "flushed between timed steps (256 MiB memset)", "cuda_graph": True,
                       "gemm": "tcgen05 3xTF32" if plan.launches.tensor_cores else "fp32 FMA"},
            "e2e": {"value": pairs_total / t_e2e_max, "unit": "pairs/s", "h2d_bytes_per_step": img_bytes,
                    "d2h_bytes_per_step": B * H * W * 4},
            "gpu_launches": fwd_launches * steps,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "encoder_ms_per_step": 1e3 * (t_dev_max - t_hot_max) / steps,
            "hot_path": {"value": pairs_total / t_hot_max, "unit": "pairs/s", "ms_per_step": 1e3 * t_hot_max / steps,
                         "includes": "SURVEY.md §8(a) A1-A13 only (cost volume ... disparity), one CUDA graph over feature maps resident "
                                     "in HBM", "ms_eager_sum": round(hot_ms, 3), "launches_per_step": launches_per_step,
                         "kernels": kernels},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
