"""ctypes binding of libnmrf_b200.so (C-ABI declared in include/nmrf_b200.h).

The product path has NO fallback: if the shared library is missing, import of this
module raises, and every entry point raises RuntimeError on a non-zero return code.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# NMRF_B200_LIB: development override (A/B of two builds of the SAME library, e.g. `make LO_TRUNC=1 LIB=...`); not a fallback
LIB_PATH = os.environ.get("NMRF_B200_LIB") or os.path.join(_HERE, "libnmrf_b200.so")
ABI_VERSION = 7


class GemmArgs(Structure):
    """mirror of nmrf_gemm_args"""
    _fields_ = [
        ("X", c_void_p), ("ldx", c_int), ("Kx", c_int),
        ("E", c_void_p), ("lde", c_int), ("Ke", c_int), ("ediv", c_int),
        ("ln_gamma", c_void_p), ("ln_beta", c_void_p),
        ("ln_stats", c_void_p),
        ("W", c_void_p), ("ldw", c_int),
        ("bias", c_void_p),
        ("R", c_void_p), ("ldr", c_int),
        ("Y", c_void_p), ("ldy", c_int),
        ("rows", c_int), ("N", c_int),
        ("act", c_int),
        ("Wt_hi", c_void_p), ("Wt_lo", c_void_p),
    ]


class MlpArgs(Structure):
    """mirror of nmrf_mlp_args"""
    _fields_ = [
        ("X", c_void_p), ("ldx", c_int), ("Kx", c_int),
        ("E", c_void_p), ("lde", c_int), ("Ke", c_int),
        ("Wstream", c_void_p),
        ("bias_mid", c_void_p),
        ("ln_gamma", c_void_p), ("ln_beta", c_void_p),
        ("b1", c_void_p),
        ("bias_out", c_void_p),
        ("Y", c_void_p), ("ldy", c_int),
        ("rows", c_int),
        ("out_stats", c_void_p),
        ("e_identity", c_int),
    ]


class ConvArgs(Structure):
    """mirror of nmrf_conv_args"""
    _fields_ = [
        ("X", c_void_p), ("N", c_int), ("H", c_int), ("W", c_int),
        ("img_stride", c_int64), ("row_stride", c_int), ("pix_stride", c_int),
        ("Cin", c_int), ("kh", c_int), ("kw", c_int), ("stride", c_int), ("pad", c_int),
        ("Wt_hi", c_void_p), ("Wt_lo", c_void_p),
        ("bias", c_void_p),
        ("Y", c_void_p), ("Cout", c_int), ("Ho", c_int), ("Wo", c_int),
    ]


class SeedWeights(Structure):
    """mirror of nmrf_seed_weights"""
    _fields_ = [("w0", c_void_p), ("b0", c_void_p), ("w1", c_void_p), ("b1", c_void_p),
                ("w2", c_void_p), ("b2", c_void_p)]


# name -> argtypes; every function returns int except the three helpers
_I, _F, _D, _P = c_int, c_float, c_double, c_void_p
SIGNATURES = {
    "nmrf_token_gemm": [POINTER(GemmArgs), _P],
    "nmrf_mlp_chain": [POINTER(MlpArgs), _P],
    "nmrf_conv2d": [POINTER(ConvArgs), _P],
    "nmrf_row_stats": [_P, _I, _I, _P, _P],
    "nmrf_split_tf32": [_P, _P, _P, c_int64, _P],
    "nmrf_instnorm_stats": [_P, _I, _I, _I, _P, _P],
    "nmrf_instnorm_apply": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "nmrf_image_prep": [_P, _P, _I, _I, _I, _I, _I, c_int64, c_int64, c_int64, c_int64, _P, _P],
    "nmrf_avgpool2_split": [_P, _I, _I, _I, _I, _P, _P, _P, _P],
    "nmrf_set_attention_impl": [_I],
    "nmrf_debug_set_trace": [_P],
    "nmrf_pack_weight_tiles": [_P, _I, _I, _P, _P, _P],
    "nmrf_cost_volume_topk": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, POINTER(SeedWeights), _P, _P, _P, _P],
    "nmrf_prop_gather": [_P, _P, _I, _I, _I, _I, _D, _I, _P, _I, _P, _P],
    "nmrf_stripe_attention": [_P, _I, _I, _I, _I, _P, _P, _P, _P],
    "nmrf_prop_head_tail": [_P, _P, _P, _P, _I, _P, _P, _P],
    "nmrf_warp_corr_embed": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _D, _P, _P, _P],
    "nmrf_zero_pad_rows": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "nmrf_proposal_attention": [_P, _I, _I, _P, _P],
    "nmrf_window_attention": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "nmrf_select_median": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "nmrf_refine_tail": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "nmrf_disp_metrics": [_P, _P, _P, _I, c_int64, _F, POINTER(c_float), _I, _P, _P],
    "nmrf_disp_to_kitti_u16": [_P, c_int64, _P, _P],
    "nmrf_ms_deform_attn_forward": [_P, POINTER(c_int64), POINTER(c_int64), _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "nmrf_ms_deform_attn_forward_dev": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
}
HELPERS = {"nmrf_abi_version": (c_int, []), "nmrf_last_error": (c_char_p, []), "nmrf_launch_count": (c_uint64, [])}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C nmrf_b200/csrc`. nmrf_b200 has no CPU/PyTorch fallback.")


class _Library:
    """The shared library, dlopen'ed at FIRST USE (its presence is checked at import, above).  Code that only needs the
    parameter containers or the synthetic generators -- bench.py's `--impl reference` arm, the CPU tests of the host logic --
    therefore never maps libnmrf_b200.so into its process; every compute entry point still fails loudly without it."""

    def __init__(self):
        self._cdll = None

    def _load(self):
        cdll = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(cdll, name)
            fn.argtypes = args
            fn.restype = c_int
        for name, (res, args) in HELPERS.items():
            fn = getattr(cdll, name)
            fn.argtypes = args
            fn.restype = res
        if cdll.nmrf_abi_version() != ABI_VERSION:
            raise ImportError(f"libnmrf_b200.so ABI {cdll.nmrf_abi_version()} != binding ABI {ABI_VERSION}: rebuild")
        self._cdll = cdll
        return cdll

    @property
    def loaded(self):
        return self._cdll is not None

    def __getattr__(self, name):
        cdll = self.__dict__.get("_cdll") or self._load()
        return getattr(cdll, name)


lib = _Library()


def check(rc, what):
    if rc != 0:
        msg = lib.nmrf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count():
    return int(lib.nmrf_launch_count())
