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
    assert _lib.lib.nmrf_abi_version() == _lib.ABI_VERSION == 7
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
    assert ctypes.sizeof(_lib.MlpArgs) == 112


def test_struct_field_offsets_match_a_c_compiler(tmp_path):
    """compile include/nmrf_b200.h as plain C (gcc) and compare offsetof() of every struct field with the ctypes mirrors"""
    import shutil
    import subprocess
    from nmrf_b200 import _lib
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no gcc")
    structs = {"nmrf_gemm_args": _lib.GemmArgs, "nmrf_mlp_args": _lib.MlpArgs, "nmrf_seed_weights": _lib.SeedWeights,
               "nmrf_conv_args": _lib.ConvArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "nmrf_b200.h")}"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.strip().splitlines():
        cname, fname, val = line.split()
        ct = structs[cname]
        want = ctypes.sizeof(ct) if fname == "size" else getattr(ct, fname).offset
        assert int(val) == want, f"{cname}.{fname}: C says {val}, ctypes {want}"
        seen += 1
    assert seen == sum(len(ct._fields_) + 1 for ct in structs.values())
