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
    assert lib.ungar_b200_abi_version() == _lib.EXPECTED_ABI == 2


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


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument errors of the QP / line-search / SQP / tape / generic-QP entry points are reported before any device work."""
    import numpy as np

    lib = _lib.load()
    opts = _lib.SqpOptions()
    assert lib.ungar_b200_sqp_options_default(ctypes.byref(opts)) == _lib.OK
    # reference defaults: soft_sqp.hpp:44-50, backtracking_line_search.hpp:70-76, the 1e-6 literal of soft_sqp.hpp:103
    assert (opts.max_iterations, opts.constraint_violation_multiplier) == (10, 1.0)
    assert (opts.alpha_min, opts.theta_min, opts.theta_max, opts.eta) == (1e-4, 1e-6, 1e-2, 1e-4)
    assert (opts.gamma_phi, opts.gamma_theta, opts.gamma_alpha, opts.objective_tolerance) == (1e-6, 1e-6, 0.5, 1e-6)
    assert lib.ungar_b200_sqp_options_default(None) == _lib.EINVAL
    assert lib.ungar_b200_line_search(None, None, 1, 1, None, 1, ctypes.byref(opts), None, None, None) == _lib.EINVAL
    assert lib.ungar_b200_sqp_solve(None, None, 1, 1, ctypes.byref(opts), None, None, 0, None) == _lib.EINVAL
    assert lib.ungar_b200_qp_solve(None, None, 1, 1, None, 1, None, 0, None) == _lib.EINVAL
    assert lib.ungar_b200_jacobian_blocks(None, None, 1, 1, None, 1, 0, None) == _lib.EINVAL
    assert lib.ungar_b200_tape_forward_zero(None, None, 1, 1, None, 1, 0, None) == _lib.EINVAL
    assert lib.ungar_b200_tape_sparse_jacobian(None, None, 1, 1, None, 1, 0, None) == _lib.EINVAL
    assert lib.ungar_b200_tape_sparse_hessian(None, None, None, 1, 1, None, 1, 0, None) == _lib.EINVAL
    assert lib.ungar_b200_tape_info(None, None) == _lib.EINVAL
    h = ctypes.c_void_p()
    assert lib.ungar_b200_tape_create(None, 3, 1, None, None, 1, 0, ctypes.byref(h)) == _lib.EINVAL
    assert lib.ungar_b200_tape_create(None, 0, 0, None, None, 0, 0, None) == _lib.EINVAL
    # generic QP: sizes and null arrays; too large for the dense fallback
    one = np.zeros(2, dtype=np.int32)
    v = np.zeros(1)
    assert lib.ungar_b200_kkt_solve_csc(0, 0, one.ctypes.data, None, None, v.ctypes.data, None, None, None, None, 1e-9, 1e-9,
                                        v.ctypes.data, None, 0) == _lib.EINVAL
    assert lib.ungar_b200_kkt_solve_csc(20000, 0, one.ctypes.data, None, None, v.ctypes.data, None, None, None, None, 1e-9, 1e-9,
                                        v.ctypes.data, None, 0) == _lib.EUNSUPPORTED
    assert b"16384" in lib.ungar_b200_last_error()


def test_tape_handles_work_without_a_gpu_until_the_first_evaluation():
    """ungar_b200_tape_create analyses on the host only; evaluation fails loudly (ECUDA) when no device exists."""
    import numpy as np

    from ungar_b200 import autodiff as A

    f = A.MakeFunction(A.Blueprint(lambda v: [v[0] * v[1] + A.sin(v[2])], 3, 0, "tiny", A.ALL))
    assert f.JacobianSparsity()[1].tolist() == [0, 1, 2]
    assert list(zip(*[a.tolist() for a in f.HessianSparsity()])) == [(0, 1), (2, 2)]
    if not _has_gpu():
        with pytest.raises(_lib.UngarB200Error) as err:
            f(np.zeros(3))
        assert err.value.code == _lib.ECUDA and "no CPU fallback" in str(err.value)
