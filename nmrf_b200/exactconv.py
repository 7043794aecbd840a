"""fp32-accurate convolutions at tensor-core speed for the (out-of-path) torch feature extractor.

cuDNN's fp32 path with TF32 disabled runs at ~3 TFLOP/s on B200 for this network (62 ms for the encoder at
544x960), while its TF32 path moves the features by ~5e-4 relative, which breaks parity.
Same remedy as in the hot path: error-compensated 3xTF32.  Both operands are split with round-to-nearest
(`nmrf_split_tf32`: x = hi + lo, both exactly TF32-representable), so cuDNN's internal fp32->tf32 conversion
is the identity and     conv(x, w) ~= conv(hi, w_hi) + (conv(hi, w_lo) + conv(lo, w_hi))     (fp32 accumulate).
"""
import torch
import torch.nn.functional as F

from . import _lib
from ._lib import lib


def split_tf32(t):
    t = t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t.contiguous()
    hi, lo = torch.empty_like(t), torch.empty_like(t)
    _lib.check(lib.nmrf_split_tf32(t.data_ptr(), hi.data_ptr(), lo.data_ptr(), t.numel(),
                                   torch.cuda.current_stream().cuda_stream), "split_tf32")
    return hi, lo


class ExactConvCache:
    """pre-split weights per Conv2d module (invalidated together with the model's packed weights)"""

    def __init__(self):
        self._w = {}
        self.enabled = True

    def clear(self):
        self._w.clear()

    def weights(self, conv):
        k = id(conv)
        if k not in self._w:
            self._w[k] = split_tf32(conv.weight.detach())
        return self._w[k]

    def conv(self, conv, x):
        """drop-in for `conv(x)` (nn.Conv2d, groups=1), fp32-accurate, TF32 tensor cores"""
        a = dict(stride=conv.stride, padding=conv.padding, dilation=conv.dilation, groups=conv.groups)
        if not self.enabled or not x.is_cuda:
            return F.conv2d(x, conv.weight, conv.bias, **a)
        w_hi, w_lo = self.weights(conv)
        x_hi, x_lo = split_tf32(x)
        with torch.backends.cudnn.flags(enabled=True, benchmark=False, allow_tf32=True):
            y = F.conv2d(x_hi, w_lo, None, **a)
            y += F.conv2d(x_lo, w_hi, None, **a)
            y += F.conv2d(x_hi, w_hi, None, **a)
        if conv.bias is not None:
            y += conv.bias.view(1, -1, 1, 1)
        return y


def patch_convs(module, cache):
    """route every nn.Conv2d under `module` through `cache.conv` (forward only; parameters untouched)"""
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d) and m.groups == 1 and not hasattr(m, "_nmrf_exact"):
            m._nmrf_exact = True
            m.forward = (lambda x, _m=m: cache.conv(_m, x))
