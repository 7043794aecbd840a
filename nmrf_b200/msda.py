"""Drop-in for the reference's only native extension (operator boundary B2).

Replaces the pybind module `MultiScaleDeformableAttention` (ops/src/vision.cpp:13-16) and
`ops.functions.MSDeformAttnFunction` (ops/functions/ms_deform_attn_func.py:19-46) with the
sm_100a kernel behind `nmrf_ms_deform_attn_forward[_dev]` (include/nmrf_b200.h).

Usage inside the reference tree:
    import nmrf_b200.msda as msda; msda.install_as_reference_extension()
    # from now on `import MultiScaleDeformableAttention as MSDA` resolves to this module, so
    # ops/functions/ms_deform_attn_func.py and ops/modules/ms_deform_attn.py work unchanged.

Same argument meaning and error behaviour as the reference (ms_deform_attn_cuda.cu:28-52):
all tensors must be contiguous CUDA tensors; batch % min(batch, im2col_step) == 0.  Forward
only (inference hot path): backward raises.
"""
import ctypes
import sys

import torch
from torch.autograd import Function

from . import _lib
from ._lib import lib


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """value [N,S,M,Dh]; spatial_shapes [L,2] int64; level_start_index [L] int64;
    sampling_loc [N,Lq,M,L,P,2]; attn_weight [N,Lq,M,L,P] -> [N,Lq,M*Dh] (fp32)."""
    for name, t in (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)):
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    for name, t in (("value", value), ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)):
        _require(t.is_cuda, f"{name} must be a CUDA tensor")
        _require(t.dtype == torch.float32, f"{name} must be float32 (fp64 is not built; the reference casts to fp32 "
                                           "under autocast, ms_deform_attn_func.py:21)")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes / level_start_index must be int64")
    N, S, M, Dh = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    step = min(N, int(im2col_step))
    _require(step > 0 and N % step == 0, f"batch({N}) must divide im2col_step({step})")
    out = torch.empty(N, Lq, M * Dh, dtype=torch.float32, device=value.device)
    stream = torch.cuda.current_stream(value.device).cuda_stream
    with torch.cuda.device(value.device):
        if spatial_shapes.is_cuda and level_start_index.is_cuda:
            rc = lib.nmrf_ms_deform_attn_forward_dev(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), N, S, M, Dh, L, Lq, P, out.data_ptr(), stream)
        else:
            sh = spatial_shapes.cpu().contiguous()
            st = level_start_index.cpu().contiguous()
            rc = lib.nmrf_ms_deform_attn_forward(
                value.data_ptr(), ctypes.cast(sh.data_ptr(), ctypes.POINTER(ctypes.c_int64)),
                ctypes.cast(st.data_ptr(), ctypes.POINTER(ctypes.c_int64)), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), N, S, M, Dh, L, Lq, P, out.data_ptr(), stream)
    _lib.check(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_backward(*args, **kwargs):
    raise NotImplementedError("nmrf_b200 implements the inference hot path: MSDeformAttn backward "
                              "(ops/src/cuda/ms_deform_im2col_cuda.cuh:301-920) is out of scope")


class MSDeformAttnFunction(Function):
    """ops.functions.MSDeformAttnFunction: `.apply(value, shapes, level_start, loc, attn, im2col_step)`."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        return ms_deform_attn_forward(value.float(), value_spatial_shapes, value_level_start_index,
                                      sampling_locations.float(), attention_weights.float(), im2col_step)

    @staticmethod
    def backward(ctx, grad_output):
        ms_deform_attn_backward()


def install_as_reference_extension():
    """Make `import MultiScaleDeformableAttention` resolve to this module."""
    sys.modules["MultiScaleDeformableAttention"] = sys.modules[__name__]
