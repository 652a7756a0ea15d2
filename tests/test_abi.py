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


def test_product_fails_loudly_without_gpu_or_library(monkeypatch, tmp_path):
    """No CPU fallback: without a CUDA device the engine raises; without the shared library the loader raises."""
    import pytest
    import torch

    from optable_b200 import backend

    if not torch.cuda.is_available():
        with pytest.raises(backend.BackendError):
            backend.Engine.get(0)
        import optable_b200 as ob

        table = ob.OpticalTable()
        table.add_components([ob.Mirror([0, 0, 0])])
        with pytest.raises(backend.BackendError):
            table.ray_tracing([ob.Ray([-1, 0, 0], [1, 0, 0])])
    monkeypatch.setattr(backend, "_lib", None)
    monkeypatch.setattr(backend, "_SO", str(tmp_path / "missing.so"))
    with pytest.raises(backend.BackendError):
        backend.lib()


def test_product_never_imports_the_oracle():
    """The checker is test infrastructure: nothing under optable_b200/ may reference it."""
    import glob

    for path in glob.glob(os.path.join(ROOT, "optable_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h")):
            text = open(path).read()
            assert "import oracle" not in text and "from oracle" not in text and "optb_oracle" not in text, path
