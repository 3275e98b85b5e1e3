"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/hipStateVec.h
declares, validates arguments like the reference, and fails loudly without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from rocquantum_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "hipStateVec.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rocsvx?[A-Z]\w*)\s*\(", txt)))


@pytest.mark.parametrize("prec", ["c64", "c128"])
def test_every_declared_symbol_is_exported(prec):
    lib = capi.load(prec)
    declared = _header_symbols()
    assert len([s for s in declared if not s.startswith("rocsvx")]) == 42      # the reference's 42 rocsv* entry points
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(capi.SYMBOLS) == declared                                     # ctypes table and header agree
    assert lib.rocsvxGetPrecisionBytes() == (4 if prec == "c64" else 8)


def test_reference_header_signatures_are_kept():
    """Every rocsv* prototype of the reference header appears with the same parameter list."""
    ref = "/root/reference/rocquantum/include/rocquantum/hipStateVec.h"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present")

    def protos(path):
        txt = open(path).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"//[^\n]*", "", txt)
        out = {}
        for m in re.finditer(r"(rocqStatus_t|void\*)\s+(rocsv[A-Z]\w*)\s*\(([^)]*)\)\s*;", txt):
            params = [re.sub(r"\s+", " ", p.strip()) for p in m.group(3).split(",")]
            types = [re.sub(r"\s*\w+$", "", p).replace(" *", "*") if p != "void" else "void" for p in params]
            out[m.group(2)] = (m.group(1), types)
        return out

    want, have = protos(ref), protos(os.path.join(ROOT, "include", "hipStateVec.h"))
    assert len(want) == 42
    for name, sig in want.items():
        assert name in have, name
        assert have[name] == sig, (name, have[name], sig)


def test_null_handle_is_invalid_value():
    lib = capi.load("c64")
    assert lib.rocsvCreate(None) == capi.INVALID_VALUE
    assert lib.rocsvDestroy(None) == capi.SUCCESS                   # hipStateVec.cpp:203-210
    assert lib.rocsvApplyH(None, None, 3, 0) == capi.INVALID_VALUE  # :280
    assert lib.rocsvFreeState(None) == capi.INVALID_VALUE
    assert lib.rocsvApplyMultiControlledX(None, None, 3, None, 0, 0) == capi.INVALID_VALUE
    assert lib.rocsvGetStateVectorFull(None, None, None) == capi.INVALID_VALUE


def test_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.load("c64")
    h = C.c_void_p()
    assert lib.rocsvCreate(C.byref(h)) == capi.HIP_ERROR            # no CPU fallback
    assert not h.value


def test_exchange_plan_validation():
    lib = capi.load("c64")
    n = C.c_size_t()
    assert lib.rocsvxDistPlanExchange(4, 3, 0, capi.uarr([3]), capi.uarr([4]), 1, None, 0, C.byref(n)) == capi.INVALID_VALUE
    assert lib.rocsvxDistPlanExchange(4, 2, 0, capi.uarr([4]), capi.uarr([4]), 1, None, 0, C.byref(n)) == capi.INVALID_VALUE
    assert lib.rocsvxDistPlanExchange(4, 2, 0, capi.uarr([3]), capi.uarr([4]), 1, None, 0, C.byref(n)) == capi.SUCCESS
    assert n.value == 1


def test_product_never_touches_the_oracle_or_the_test_interpreter():
    """oracle/ and tests/hostemu are test infrastructure: nothing under rocquantum_b200/ or include/ imports, includes or
    loads them (comments may cite them), and the engine libraries depend on CUDA + libc only."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    py = re.compile(r"^\s*(from|import)\s+(oracle|tests)\b", re.M)
    cc = re.compile(r"#\s*include\s*[<\"][^>\"]*(oracle|hostemu)|dlopen\([^)]*(oracle|emul)", re.M)
    for base in ("rocquantum_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, base)):
            if "_obj" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                txt = open(os.path.join(dirpath, f), errors="ignore").read() if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp")) else ""
                assert not (f.endswith(".py") and py.search(txt)), f
                assert not (not f.endswith(".py") and cc.search(txt)), f
    for prec in ("c64", "c128"):
        out = subprocess.run(["ldd", capi.lib_path(prec)], capture_output=True, text=True).stdout
        assert "oracle" not in out and "emul" not in out
        assert "libcudart" in out
