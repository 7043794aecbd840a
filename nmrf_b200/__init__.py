"""nmrf_b200: NMRF-Stereo's inference hot path as hand-written sm_100a CUDA kernels behind a C-ABI.

Public surface (mirrors the reference's own):
    nmrf_b200.NMRF, DPN, build_model, get_cfg      <- nmrf.models.NMRF / nmrf.models.build_model
    nmrf_b200.msda.MSDeformAttnFunction, ...        <- ops.functions / MultiScaleDeformableAttention
Importing the package loads nmrf_b200/libnmrf_b200.so and fails loudly if it is missing.
"""
from . import _lib  # noqa: F401  (loads the shared library or raises)
from .config import build_model, get_cfg  # noqa: F401
from .model import DPN, NMRF  # noqa: F401
from .backbone import Backbone  # noqa: F401

__all__ = ["NMRF", "DPN", "Backbone", "build_model", "get_cfg"]
