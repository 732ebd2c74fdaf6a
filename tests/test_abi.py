"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol the
header declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from ungar_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ungar_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ungar_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_for_sm_100a():
    path = build.build()
    assert os.path.exists(path)
    assert "arch=compute_100a,code=sm_100a" in " ".join(build.NVCC_FLAGS)


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    names = declared_symbols()
    assert names == sorted(_lib.SYMBOLS)
    for name in names:
        assert getattr(lib, name) is not None
    assert lib.ungar_b200_abi_version() == 1


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    handle = ctypes.c_void_p()
    assert lib.ungar_b200_model_create(None, ctypes.byref(handle)) == _lib.EINVAL
    bad = _lib.ModelDesc(7, 30, _lib.F64, 0, 100.0, 2e-5)
    assert lib.ungar_b200_model_create(ctypes.byref(bad), ctypes.byref(handle)) == _lib.EINVAL
    assert b"unknown model kind" in lib.ungar_b200_last_error()
    bad = _lib.ModelDesc(0, 1, _lib.F64, 0, 100.0, 2e-5)
    assert lib.ungar_b200_model_create(ctypes.byref(bad), ctypes.byref(handle)) == _lib.EINVAL
    bad = _lib.ModelDesc(0, 30, _lib.F64, 0, -1.0, 2e-5)
    assert lib.ungar_b200_model_create(ctypes.byref(bad), ctypes.byref(handle)) == _lib.EINVAL
    assert lib.ungar_b200_kkt_blocks(None, None, 1, 1, None, 1, 0, None) == _lib.EINVAL


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour on a box WITHOUT a GPU")
def test_no_cpu_fallback_without_a_device():
    """The product path must fail loudly, not fall back to a CPU evaluation."""
    lib = _lib.load()
    handle = ctypes.c_void_p()
    desc = _lib.ModelDesc(2, 30, _lib.F64, 0, 1.0, 1.0)
    assert lib.ungar_b200_model_create(ctypes.byref(desc), ctypes.byref(handle)) == _lib.ECUDA
    assert b"no CPU fallback" in lib.ungar_b200_last_error()
    import ungar_b200

    with pytest.raises(_lib.UngarB200Error):
        ungar_b200.Model("quadruped", 30)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under ungar_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ungar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and '#include "../../oracle' not in text, f
