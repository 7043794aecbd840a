"""On-device evaluation and disparity writers: the step after the hot path in every pipeline of the reference
(SURVEY.md §8(f) N4).

`DispEvaluator` mirrors `nmrf.utils.evaluation.DispEvaluator` (reference nmrf/utils/evaluation.py:292-417): same
constructor arguments, `reset / process(inputs, outputs) / evaluate()`, same result dict ({'disp': {'epe', 'd1',
'bad <t>'...}}).  Differences in HOW, not WHAT:
  * the per-image statistics (EPE, D1, bad-tau) are reduced on the GPU by `nmrf_disp_metrics` (one launch per batch, no
    host synchronisation in `process`); the reference pulls one `.item()` per image and statistic;
  * across ranks ONE all_gather of a small fp64 vector (sum of per-image values + image count) replaces the reference's
    gloo `gather_object` of Python lists (nmrf/utils/dist_utils.py:142-171);
  * sums are accumulated in double (the reference's `.mean()` is an fp32 reduction).
`eval_prop=True` needs the superpixel-guided downsample operator, which does not exist in the reference snapshot
(SURVEY.md D1): it raises.

`write_disp_kitti` mirrors `frame_utils.writeDispKITTI` (nmrf/utils/frame_utils.py:237-239): uint16(round(disp * 256)) PNG;
the encoding runs on the device (`nmrf_disp_to_kitti_u16`), only the 16-bit image crosses PCIe.
"""
import ctypes
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import lib
from .sharding import gather_stats


def disp_metrics(disp_pr, disp_gt, valid=None, max_disp=float("inf"), thresholds=()):
    """[B,H,W] CUDA tensors -> [B, 3 + len(thresholds)] float64 (on the device): valid count, sum |e|, #D1, #bad per threshold."""
    for n, t in (("disp_pr", disp_pr), ("disp_gt", disp_gt)):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError(f"{n} must be a float32 CUDA tensor (nmrf_b200 has no CPU path)")
    if disp_pr.shape != disp_gt.shape:
        raise RuntimeError(f"shape mismatch {tuple(disp_pr.shape)} vs {tuple(disp_gt.shape)}")
    B = disp_pr.shape[0] if disp_pr.dim() == 3 else 1
    HW = disp_pr.numel() // B
    pr, gt = disp_pr.contiguous(), disp_gt.contiguous()
    v = None
    if valid is not None:
        v = valid.to(device=pr.device, dtype=torch.uint8).contiguous()
    th = (ctypes.c_float * max(len(thresholds), 1))(*[float(t) for t in thresholds])
    acc = torch.zeros(B, 3 + len(thresholds), dtype=torch.float64, device=pr.device)
    with torch.cuda.device(pr.device):
        rc = lib.nmrf_disp_metrics(pr.data_ptr(), gt.data_ptr(), None if v is None else v.data_ptr(), B, HW,
                                   float(min(max_disp, 3.0e38)), th, len(thresholds), acc.data_ptr(),
                                   torch.cuda.current_stream(pr.device).cuda_stream)
    _lib.check(rc, "disp_metrics")
    return acc


class DispEvaluator:
    """Evaluate disparity accuracy (reference nmrf/utils/evaluation.py:292-417)."""

    def __init__(self, thres, only_valid, max_disp=None, eval_prop=False, divis_by=8):
        if eval_prop:
            raise NotImplementedError("eval_prop needs frame_utils.downsample_disp (the superpixel-guided downsample), which is "
                                      "not defined anywhere in the reference snapshot (evaluation.py:363-366 is a dangling call)")
        self._max_disp = np.inf if max_disp is None else max_disp
        self._thres = list(thres) if thres is not None else []
        self._only_valid = only_valid
        self._divis_by = divis_by
        self.reset()

    def reset(self):
        self._acc = []          # device tensors [B, 3 + nt], one per processed batch

    def process(self, inputs, outputs):
        """inputs: {'disp': [B,H,W], 'valid': [B,H,W] bool, ...}; outputs: the model's dict ('disp' [B,H,W])."""
        disp_pr = outputs["disp"]
        disp_gt = inputs["disp"].to(disp_pr.device, torch.float32)
        assert disp_pr.shape == disp_gt.shape, (disp_pr.shape, disp_gt.shape)
        valid = inputs["valid"] if self._only_valid else None
        self._acc.append(disp_metrics(disp_pr, disp_gt, valid, self._max_disp, [float(t) for t in self._thres]))

    def _local_sums(self):
        """[1 + 2 + nt] float64: number of images with valid pixels, then the SUM over those images of the per-image epe, d1, bad-t"""
        nt = len(self._thres)
        if not self._acc:
            return torch.zeros(3 + nt, dtype=torch.float64)
        acc = torch.cat(self._acc).cpu()                         # the only device->host transfer
        ok = acc[:, 0] > 0                                       # images without valid pixels are skipped (evaluation.py:349-350)
        per_image = acc[ok, 1:] / acc[ok, :1]
        return torch.cat([ok.sum().double().reshape(1), per_image.sum(0)])

    def evaluate(self):
        loc = self._local_sums()
        dev = self._acc[0].device if self._acc else None
        allv = gather_stats(loc, device=dev if (dev is not None and torch.distributed.is_initialized()
                                                 and torch.distributed.get_backend() == "nccl") else None)
        tot = allv.sum(0)
        n = max(float(tot[0]), 1.0)
        res = OrderedDict(epe=float(tot[1]) / n, d1=100.0 * float(tot[2]) / n)
        for i, t in enumerate(self._thres):
            res[f"bad {t}"] = 100.0 * float(tot[3 + i]) / n
        return {"disp": res}


def disp_to_kitti_u16(disp):
    """[...] float32 CUDA disparity -> uint16 tensor of the same shape: round(disp * 256) (frame_utils.py:237-239)"""
    if not (disp.is_cuda and disp.dtype == torch.float32):
        raise RuntimeError("disp must be a float32 CUDA tensor")
    d = disp.contiguous()
    out = torch.empty(d.shape, dtype=torch.uint16, device=d.device)
    with torch.cuda.device(d.device):
        rc = lib.nmrf_disp_to_kitti_u16(d.data_ptr(), d.numel(), out.data_ptr(), torch.cuda.current_stream(d.device).cuda_stream)
    _lib.check(rc, "disp_to_kitti_u16")
    return out


def write_disp_kitti(filename, disp):
    """`frame_utils.writeDispKITTI(filename, disp)` for a [H,W] CUDA disparity map: 16-bit PNG of round(disp * 256)."""
    u16 = disp_to_kitti_u16(disp).cpu().numpy()
    try:
        import cv2
        if not cv2.imwrite(filename, u16):
            raise IOError(f"cv2.imwrite failed for {filename}")
    except ImportError:
        from PIL import Image
        Image.fromarray(u16).save(filename)
