"""CPU: the C-ABI shared library loads and exports every symbol include/nmrf_b200.h declares, with the
argument counts the ctypes binding assumes (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nmrf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|uint64_t|const char\*)\s+(nmrf_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_header_symbols_are_exported():
    from nmrf_b200 import _lib
    decl = _declared()
    assert len(decl) >= 26
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/nmrf_b200.h but not exported"


def test_binding_matches_header():
    from nmrf_b200 import _lib
    decl = _declared()
    bound = dict(_lib.SIGNATURES)
    bound.update({k: v[1] for k, v in _lib.HELPERS.items()})
    assert set(bound) == set(decl), set(bound) ^ set(decl)
    for name, args in bound.items():
        assert len(args) == decl[name], f"{name}: binding has {len(args)} args, header {decl[name]}"


def test_helpers_work_without_a_gpu():
    from nmrf_b200 import _lib
    assert _lib.lib.nmrf_abi_version() == _lib.ABI_VERSION == 6
    assert isinstance(_lib.launch_count(), int)
    # argument validation happens before any CUDA call: a bad GEMM is rejected with a message
    a = _lib.GemmArgs()
    rc = _lib.lib.nmrf_token_gemm(ctypes.byref(a), None)
    assert rc == 1 and b"null pointer" in _lib.lib.nmrf_last_error()


def test_struct_layout_matches_c():
    """GemmArgs / SeedWeights mirror the C structs (sizes under the LP64 ABI with natural alignment)."""
    from nmrf_b200 import _lib
    assert ctypes.sizeof(_lib.SeedWeights) == 6 * 8
    # X,ldx,Kx | E,lde,Ke,ediv | g,b | W,ldw | bias | R,ldr | Y,ldy | rows,N,act
    assert ctypes.sizeof(_lib.GemmArgs) == 144
    # X,ldx,Kx | E,lde,Ke | Wstream | bias_mid | g,b | b1 | bias_out | Y,ldy,rows | e_identity (+pad)
    assert ctypes.sizeof(_lib.MlpArgs) == 104
