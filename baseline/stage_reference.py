"""Stage the UNMODIFIED reference tree into baseline/_ref/ (git-ignored, but it travels to the GPU box with gpurun).

    python baseline/stage_reference.py            (run in the build container, where /root/reference exists)

The reference is a plain Python package without setup.py / pyproject (pip cannot install it) plus one CUDA extension
(`ops/`) whose setup.py refuses to build without a GPU (ops/setup.py:42-43); so "installing" it means copying the package
directories as they are:
    nmrf/ ops/ configs/   ->   baseline/_ref/
On the GPU box they serve (a) the MSDA incumbent: the reference kernel JIT-compiled for sm_100 from baseline/_ref/ops/src
(tools/msda_bench.py), (b) the Swin-T + DeformNeck encoder of BASELINE config 5 (`nmrf.models.backbone.SwinAdaptor`), run
with `nmrf_b200.msda` as its MultiScaleDeformableAttention extension.  Nothing under baseline/_ref is part of the product
or of the git history.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("NMRF_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(os.path.join(SRC, "nmrf")):
        print(f"{SRC} not found: nothing staged")
        return 1
    os.makedirs(DST, exist_ok=True)
    for d in ("nmrf", "ops", "configs"):
        dst = os.path.join(DST, d)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, d), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "build", "*.egg-info"))
    print("staged", sorted(os.listdir(DST)), "->", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
