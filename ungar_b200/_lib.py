"""ctypes binding of the C ABI declared in include/ungar_b200.h (the only way Python reaches the kernels)."""
from __future__ import annotations

import ctypes
import os

from . import build as _build

c_i32, c_i64, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
c_i64_p = ctypes.POINTER(c_i64)

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, 1, 2, 3, 4
F32, F64 = 0, 1
OBJECTIVE, EQUALITIES, INEQUALITIES, SOFT_INEQUALITIES = 0, 1, 2, 3
MEM_DEVICE, MEM_HOST = 0, 1
RECORD_DENSE, RECORD_COMPACT = 0, 1
EXPECTED_ABI = 2  # include/ungar_b200.h UNGAR_B200_ABI_VERSION the signatures below were written for
SUMMARY_SIZE = 32
LINE_SEARCH_INFO_SIZE = 8
SQP_RUNNING, SQP_CONVERGED, SQP_LINE_SEARCH_FAILED = 0, 1, 2

# every symbol include/ungar_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "ungar_b200_model_create", "ungar_b200_model_destroy", "ungar_b200_function_info",
    "ungar_b200_jacobian_sparsity", "ungar_b200_hessian_sparsity", "ungar_b200_forward_zero",
    "ungar_b200_sparse_jacobian", "ungar_b200_sparse_hessian", "ungar_b200_kkt_layout_get",
    "ungar_b200_kkt_blocks", "ungar_b200_summaries", "ungar_b200_kkt_step", "ungar_b200_set_profiling",
    "ungar_b200_sweep_times", "ungar_b200_qp_solve", "ungar_b200_launch_count", "ungar_b200_last_error",
    "ungar_b200_abi_version", "ungar_b200_sqp_options_default", "ungar_b200_line_search", "ungar_b200_sqp_solve",
    "ungar_b200_tape_create", "ungar_b200_tape_destroy", "ungar_b200_tape_info", "ungar_b200_tape_jacobian_pattern",
    "ungar_b200_tape_hessian_pattern", "ungar_b200_tape_set_jacobian_elements", "ungar_b200_tape_set_hessian_elements",
    "ungar_b200_tape_forward_zero", "ungar_b200_tape_sparse_jacobian", "ungar_b200_tape_sparse_hessian", "ungar_b200_kkt_solve_csc",
    "ungar_b200_jacobian_blocks", "ungar_b200_kkt_compact_map", "ungar_b200_set_parameters", "ungar_b200_kkt_step_x", "ungar_b200_tape_special_info", "ungar_b200_tape_kernel_source", "ungar_b200_tape_special_wait",
]


class ModelDesc(ctypes.Structure):
    _fields_ = [("kind", c_i32), ("horizon", c_i32), ("dtype", c_i32), ("device", c_i32),
                ("barrier_stiffness", c_f64), ("barrier_epsilon", c_f64), ("record_format", c_i32), ("reserved", c_i32)]


class KktLayout(ctypes.Structure):
    _names = ["g", "A", "C", "h", "cost", "grad", "H", "HN", "Hc", "size", "nx", "nu", "nz", "horizon", "n_dec",
              "n_par", "m_eq", "m_ineq", "tri", "tri_terminal", "legs", "hc_per_node", "compact", "dense_size", "node_stride",
              "c_Cs", "c_Cp", "c_g", "c_q", "c_Hd", "c_Hb", "c_h", "c_AQ", "c_AP", "tail", "t_g0", "t_qN", "t_HN", "t_cost"]
    _fields_ = [(n, c_i64) for n in _names]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n in self._names}


class SqpOptions(ctypes.Structure):
    _fields_ = [("max_iterations", c_i32), ("reserved", c_i32), ("constraint_violation_multiplier", c_f64),
                ("alpha_min", c_f64), ("theta_min", c_f64), ("theta_max", c_f64), ("eta", c_f64), ("gamma_phi", c_f64),
                ("gamma_theta", c_f64), ("gamma_alpha", c_f64), ("objective_tolerance", c_f64)]


class UngarB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"ungar_b200 error {code}: {message}")
        self.code = code


_lib = None


def load() -> ctypes.CDLL:
    """Load (building if the sources are newer) the in-tree CUDA library.  Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path) or not _build.up_to_date():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on the box: use the prebuilt library if there is one
            if not os.path.exists(path):
                raise RuntimeError(f"ungar_b200: CUDA library missing and cannot be built: {exc}") from exc
            import warnings

            warnings.warn(f"ungar_b200: the CUDA library is older than its sources and the rebuild failed ({exc}); "
                          f"binding the stale library (its ABI version is checked below)")
    L = ctypes.CDLL(path)
    missing = [name for name in SYMBOLS if not hasattr(L, name)]
    if missing:
        raise RuntimeError(f"ungar_b200: {path} does not export {missing}: stale build, run `python -m ungar_b200.build --force`")
    L.ungar_b200_abi_version.restype = c_i32
    abi = int(L.ungar_b200_abi_version())
    if abi != EXPECTED_ABI:
        raise RuntimeError(f"ungar_b200: {path} implements ABI {abi}, these bindings were written for ABI {EXPECTED_ABI}: "
                           f"stale build, run `python -m ungar_b200.build --force`")
    L.ungar_b200_model_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.POINTER(c_vp)]
    L.ungar_b200_model_destroy.argtypes = [c_vp]
    L.ungar_b200_function_info.argtypes = [c_vp, c_i32, c_i64_p, c_i64_p, c_i64_p, c_i64_p, c_i64_p]
    for name in ("ungar_b200_jacobian_sparsity", "ungar_b200_hessian_sparsity"):
        getattr(L, name).argtypes = [c_vp, c_i32, ctypes.POINTER(c_i64_p), ctypes.POINTER(c_i64_p), c_i64_p]
    for name in ("ungar_b200_forward_zero", "ungar_b200_sparse_jacobian", "ungar_b200_sparse_hessian"):
        getattr(L, name).argtypes = [c_vp, c_i32, c_vp, c_i64, c_i64, c_vp, c_i64, c_i32, c_vp]
    L.ungar_b200_kkt_layout_get.argtypes = [c_vp, ctypes.POINTER(KktLayout)]
    L.ungar_b200_kkt_compact_map.argtypes = [c_vp, ctypes.POINTER(ctypes.POINTER(c_i32)), c_i64_p]
    L.ungar_b200_kkt_blocks.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i32, c_vp]
    L.ungar_b200_jacobian_blocks.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i32, c_vp]
    L.ungar_b200_summaries.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]
    L.ungar_b200_kkt_step.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i32, c_vp]
    L.ungar_b200_kkt_step_x.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i32, c_vp]
    L.ungar_b200_set_parameters.argtypes = [c_vp, c_vp, c_i64, c_i64, c_i32, c_vp]
    L.ungar_b200_qp_solve.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp]
    L.ungar_b200_sqp_options_default.argtypes = [ctypes.POINTER(SqpOptions)]
    L.ungar_b200_line_search.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, ctypes.POINTER(SqpOptions), c_vp, c_vp, c_vp]
    L.ungar_b200_sqp_solve.argtypes = [c_vp, c_vp, c_i64, c_i64, ctypes.POINTER(SqpOptions), c_vp, c_vp, c_i32, c_vp]
    L.ungar_b200_tape_create.argtypes = [c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_i32, ctypes.POINTER(c_vp)]
    L.ungar_b200_tape_destroy.argtypes = [c_vp]
    L.ungar_b200_tape_info.argtypes = [c_vp, c_i64_p]
    L.ungar_b200_tape_special_info.argtypes = [c_vp, c_i64_p]
    L.ungar_b200_tape_special_wait.argtypes = [c_vp]
    L.ungar_b200_tape_kernel_source.argtypes = [c_vp, c_i32, ctypes.c_char_p, c_i64, c_i64_p, ctypes.POINTER(c_i32)]
    for name in ("ungar_b200_tape_jacobian_pattern", "ungar_b200_tape_hessian_pattern"):
        getattr(L, name).argtypes = [c_vp, ctypes.POINTER(c_i64_p), ctypes.POINTER(c_i64_p), c_i64_p]
    for name in ("ungar_b200_tape_set_jacobian_elements", "ungar_b200_tape_set_hessian_elements"):
        getattr(L, name).argtypes = [c_vp, c_vp, c_vp, c_i64]
    for name in ("ungar_b200_tape_forward_zero", "ungar_b200_tape_sparse_jacobian"):
        getattr(L, name).argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i32, c_vp]
    L.ungar_b200_tape_sparse_hessian.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i32, c_vp]
    L.ungar_b200_kkt_solve_csc.argtypes = [c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f64, c_f64, c_vp, c_vp, c_i32]
    L.ungar_b200_set_profiling.argtypes = [c_i32]
    L.ungar_b200_sweep_times.argtypes = [ctypes.POINTER(ctypes.c_float), c_i32, ctypes.POINTER(c_i32)]
    L.ungar_b200_launch_count.restype = c_i64
    L.ungar_b200_last_error.restype = ctypes.c_char_p
    L.ungar_b200_abi_version.restype = c_i32
    _lib = L
    return L


def check(code: int) -> None:
    if code != OK:
        raise UngarB200Error(code, load().ungar_b200_last_error().decode())
