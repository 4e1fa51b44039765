"""The C-ABI library builds, loads and exports every symbol include/facemap_b200.h declares; the
ctypes table covers them all.  No compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "facemap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for must in ("fm_bin_motion", "fm_gemm_tn", "fm_syevd", "fm_row_means", "fm_gram_assemble",
                 "fm_topk_components", "fm_transpose", "fm_gather_roi", "fm_mean_update"):
        assert must in syms


def test_library_builds_and_exports_every_symbol():
    from facemap_b200 import _lib, build

    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(_lib.SIGNATURES) == set(_declared_symbols())
    assert _lib.load().fm_version() >= 100


def test_library_carries_blackwell_tensor_code():
    """SASS of the GEMM must contain tcgen05 MMA (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG)."""
    import shutil
    import subprocess

    from facemap_b200 import build

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert re.search(r"UTC\w*MMA", sass), "no tcgen05.mma in SASS"
    assert "LDTM" in sass and "UTMALDG" in sass


def test_no_device_fails_loudly():
    import torch

    from facemap_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.FacemapB200Error):
        _lib.require_device(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "facemap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_header_is_plain_c():
    """the boundary is a C ABI: the header must compile as C99 (no C++ types, `extern "C"` guarded)"""
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "hdr.c")
        with open(src, "w") as f:
            f.write('#include "facemap_b200.h"\nint main(void) { return fm_version(); }\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                               "-I", os.path.join(ROOT, "include"), src])


def test_ctypes_signatures_match_header_prototypes():
    """ctypes does not check a prototype against the library: compare `_lib.SIGNATURES` with the header argument by
    argument (count, and pointer / integer / floating-point class), so that a C-ABI change cannot silently shift the
    arguments of a Python call"""
    import ctypes as C
    import re

    from facemap_b200 import _lib

    text = open(os.path.join(ROOT, "include", "facemap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {m.group(2): m.group(3) for m in re.finditer(r"\b([A-Za-z_][\w\s\*]*?)\b(fm_\w+)\s*\(([^;{]*?)\)\s*;", text)}

    def kind_of_c(arg):
        arg = arg.strip()
        if "*" in arg:
            return "ptr"
        base = arg.rsplit(None, 1)[0] if " " in arg else arg
        if base in ("double", "float"):
            return base
        if "64" in base or base in ("size_t", "long long", "unsigned long long"):
            return "int64"
        return "int32"

    def kind_of_ctypes(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        if t is C.c_double:
            return "double"
        if t is C.c_float:
            return "float"
        return "int64" if C.sizeof(t) == 8 else "int32"

    checked = 0
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert name in protos, f"{name} is in SIGNATURES but not declared in the header"
        raw = protos[name].strip()
        cargs = [] if raw in ("", "void") else [a for a in raw.split(",")]
        assert len(cargs) == len(argtypes), f"{name}: header has {len(cargs)} arguments, SIGNATURES {len(argtypes)}"
        for i, (ca, ct) in enumerate(zip(cargs, argtypes)):
            assert kind_of_c(ca) == kind_of_ctypes(ct), f"{name} argument {i}: header `{ca.strip()}` vs ctypes {ct}"
        checked += 1
    assert checked >= 40


def test_integration_doc_bindings_match_the_abi():
    """the ctypes stubs INTEGRATION.md shows a maintainer must have the argument counts of the real entry points"""
    from facemap_b200 import _lib

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    found = re.findall(r"lib\.(fm_\w+)\.argtypes\s*=\s*\[([^\]]*)\]", text)
    assert len(found) >= 10
    for name, args in found:
        n = len([a for a in args.split(",") if a.strip()])
        assert name in _lib.SIGNATURES, name
        assert n == len(_lib.SIGNATURES[name][1]), f"INTEGRATION.md: {name} has {n} argtypes, the ABI {len(_lib.SIGNATURES[name][1])}"
