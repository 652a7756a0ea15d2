"""CPU: the C-ABI library exists, loads, and exports every symbol include/optb.h declares."""
import ctypes
import os
import re

from optable_b200 import _abi as A
from optable_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "optb.h")).read()
    return sorted(set(re.findall(r"\b(optb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_python_mirror_agree():
    assert sorted(A.EXPORTED_SYMBOLS) == _declared()
    src = open(os.path.join(ROOT, "include", "optb.h")).read()
    assert f"#define OPTB_ABI_VERSION {A.ABI_VERSION}" in src
    for name, val in (("OPTB_NI_STRIDE", A.NI_STRIDE), ("OPTB_NF_STRIDE", A.NF_STRIDE), ("OPTB_MON_STRIDE", A.MON_STRIDE),
                      ("OPTB_NF_CAPMAX", A.NF_CAPMAX), ("OPTB_POLY_HEADER", A.POLY_HEADER), ("OPTB_C_COUNT", A.C_COUNT)):
        assert re.search(rf"{name}\s*=\s*{val}\b", src), name


def test_library_exports_all_symbols():
    so = build.build_extension()
    L = ctypes.CDLL(so)
    for sym in _declared():
        assert hasattr(L, sym), sym
    L.optb_abi_version.restype = ctypes.c_int
    assert L.optb_abi_version() == A.ABI_VERSION


def test_struct_sizes():
    assert ctypes.sizeof(A.SceneDesc) == 6 * 4 + 8 + 6 * 8
    assert ctypes.sizeof(A.Rays) == 8 + 15 * 8 + 8
    assert ctypes.sizeof(A.Params) == 8 + 8 + 6 * 4
    assert ctypes.sizeof(A.Result) == 2 * 8 + (13 + 3 + 1 + 1 + 2 + 10) * 8 + 4 * 8
