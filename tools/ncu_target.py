"""Target process for ncu (run under gpurun, one GPU): the bench workload, eager (no CUDA graph), N forwards.
    ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 2
    ncu --set full --clock-control none --import-source on -k regex:token_gemm_tc -s 40 -c 3 -o gpurun_out/prof_gemm python tools/ncu_target.py 1
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
hot_only = len(sys.argv) > 2 and sys.argv[2] == "hot"
sys.argv = [sys.argv[0]]
import torch  # noqa: E402
import bench  # noqa: E402
from nmrf_b200.synthetic import synthetic_pair  # noqa: E402

dev = torch.device("cuda", 0)
model, sd = bench.build_model(dev)
w = bench.WORKLOAD
img1, img2 = (t.to(dev) for t in synthetic_pair(w["B"], w["H"], w["W"], w["max_disp"], 0))
model.forward_device(img1, img2)          # warm-up: builds the plan
torch.cuda.synchronize()
plan = next(iter(model._plans.values()))
torch.cuda.cudart().cudaProfilerStart()
for _ in range(n):
    if hot_only:
        plan.run()
    else:
        model.forward_device(img1, img2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", n, "forwards; hot-path launches per forward:", plan.num_launches)
