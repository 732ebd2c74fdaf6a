"""Python mirror of the reference's evaluator interface, over the C ABI.

``Function`` keeps the public surface of ``Ungar::Autodiff::Function``
(include/ungar/autodiff/function.hpp:180-361: ``Evaluate``, ``operator()``, ``Jacobian``, ``Hessian``,
``Implements*``, ``*Size``) so parity tests read like test/autodiff/function.test.cpp; every method also accepts
a batch ``xp[B, nx + np]``.  ``Model`` stands where ``MakeFunction`` / ``MakeNLPProblem`` stand in the examples
(quadruped.example.cpp:343-363) and adds the batched KKT sweep replacing
``SoftSQPOptimizer::AssembleOSQPInstance`` (optimization/soft_sqp.hpp:141-158).

numpy arrays are host buffers (the library stages them: H2D, kernels, D2H); torch CUDA tensors are device
buffers and the call is asynchronous on torch's current stream.  There is no CPU evaluation path here.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import (EQUALITIES, F32, F64, INEQUALITIES, MEM_DEVICE, MEM_HOST, OBJECTIVE, SOFT_INEQUALITIES, KktLayout,
                   ModelDesc, check)
from .workloads import MODEL_IDS

_NP_DTYPE = {F32: np.float32, F64: np.float64}


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _torch_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


class Model:
    """One of the three reference NMPC problems at horizon N, backed by the sm_100a kernels."""

    def __init__(self, kind, horizon: int, dtype: str = "f64", device: int = 0, barrier=(100.0, 2e-5), record_format: str = "dense"):
        self.kind = MODEL_IDS[kind] if isinstance(kind, str) else int(kind)
        self.horizon = int(horizon)
        self.dtype = {"f32": F32, "f64": F64}[dtype]
        self.np_dtype = _NP_DTYPE[self.dtype]
        self.device = int(device)
        self._lib = _lib.load()
        self.barrier = (float(barrier[0]), float(barrier[1]))
        self.record_format = {"dense": _lib.RECORD_DENSE, "compact": _lib.RECORD_COMPACT}[record_format]
        desc = ModelDesc(self.kind, self.horizon, self.dtype, self.device, float(barrier[0]), float(barrier[1]), self.record_format, 0)
        handle = ctypes.c_void_p()
        check(self._lib.ungar_b200_model_create(ctypes.byref(desc), ctypes.byref(handle)))
        self._handle = handle
        lay = KktLayout()
        check(self._lib.ungar_b200_kkt_layout_get(self._handle, ctypes.byref(lay)))
        self.layout = lay.as_dict()
        self.n_xp = self.layout["n_dec"] + self.layout["n_par"]
        self.compact = bool(self.layout["compact"])
        self._c2d = None
        self.objective = Function(self, OBJECTIVE)
        self.equalityConstraints = Function(self, EQUALITIES)
        self.inequalityConstraints = Function(self, INEQUALITIES)
        self.softInequalityConstraints = Function(self, SOFT_INEQUALITIES)

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self._lib.ungar_b200_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -------------------------------------------------------------------------------------------
    def _buffers(self, xp, out, out_cols: int):
        """Normalise (xp, out) to pointers / strides / memory kind.  Returns (xp, out, batch, mem, stream, squeeze)."""
        if _is_torch(xp):
            import torch

            if not xp.is_cuda:
                raise ValueError("torch inputs must be CUDA tensors (use numpy for host buffers)")
            want = torch.float32 if self.dtype == F32 else torch.float64
            if xp.dtype != want:
                raise ValueError(f"xp dtype {xp.dtype} does not match the model dtype")
            squeeze = xp.dim() == 1
            xp2 = xp.unsqueeze(0) if squeeze else xp
            if xp2.stride(-1) != 1:
                raise ValueError("xp rows must be contiguous")
            if out is None:
                out = torch.empty((xp2.shape[0], out_cols), dtype=want, device=xp.device)
            out2 = out.unsqueeze(0) if out.dim() == 1 else out
            return xp2, out2, MEM_DEVICE, _torch_stream(), squeeze
        xp2 = np.ascontiguousarray(xp, dtype=self.np_dtype)
        squeeze = xp2.ndim == 1
        if squeeze:
            xp2 = xp2[None]
        if out is None:
            out = np.empty((xp2.shape[0], out_cols), dtype=self.np_dtype)
        out2 = out[None] if out.ndim == 1 else out
        if out2.dtype != self.np_dtype or not out2.flags.c_contiguous:
            raise ValueError("output buffer must be a C-contiguous array of the model dtype")
        return xp2, out2, MEM_HOST, None, squeeze

    def _dev(self, t, name: str, cols: int, int32: bool = False):
        """Validates one torch argument of a device-pointer entry point (the C side can only compare strides with widths): CUDA tensor on
        the model's device, the model's dtype (or int32), 2-D, contiguous rows, at least `cols` columns.  Returns (pointer, row stride)."""
        import torch

        if t is None:
            return None, 0
        if not _is_torch(t) or not t.is_cuda:
            raise ValueError(f"{name} must be a CUDA tensor")
        if t.device.index != self.device:
            raise ValueError(f"{name} lives on cuda:{t.device.index}, the model on cuda:{self.device}")
        want = torch.int32 if int32 else (torch.float32 if self.dtype == F32 else torch.float64)
        if t.dtype != want:
            raise ValueError(f"{name} has dtype {t.dtype}, expected {want}")
        if t.dim() != 2 or t.shape[1] < cols:
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected [B, >= {cols}]")
        if t.stride(-1) != 1:
            raise ValueError(f"{name} rows must be contiguous")
        return t.data_ptr(), self._ld(t)

    @staticmethod
    def _ptr(a):
        return a.data_ptr() if _is_torch(a) else a.ctypes.data

    @staticmethod
    def _ld(a):
        ld = a.stride(0) if _is_torch(a) else a.strides[0] // a.itemsize
        return max(int(ld), int(a.shape[-1]))  # a size-1 leading dimension may report any stride

    def kkt_blocks(self, xp, out=None):
        """All values, Jacobian blocks and Gauss-Newton Hessian blocks of every node: ``records[B, layout.size]``."""
        xp2, out2, mem, stream, squeeze = self._buffers(xp, out, self.layout["size"])
        check(self._lib.ungar_b200_kkt_blocks(self._handle, self._ptr(xp2), xp2.shape[0], self._ld(xp2), self._ptr(out2),
                                              self._ld(out2), mem, stream))
        return out2[0] if squeeze else out2

    def jacobian_blocks(self, xp, out=None):
        """The Jacobian sweep alone: ``g`` and ``A`` (quadruped: also ``C``) of ``records[B, layout.size]``; other blocks unspecified."""
        xp2, out2, mem, stream, squeeze = self._buffers(xp, out, self.layout["size"])
        check(self._lib.ungar_b200_jacobian_blocks(self._handle, self._ptr(xp2), xp2.shape[0], self._ld(xp2), self._ptr(out2),
                                                   self._ld(out2), mem, stream))
        return out2[0] if squeeze else out2

    def summaries(self, xp, records, out=None):
        """``[B, 32]`` per-trajectory summaries (device tensors only): the all-gather payload."""
        import torch

        if out is None:
            out = torch.empty((xp.shape[0], _lib.SUMMARY_SIZE), dtype=xp.dtype, device=xp.device)
        (px, lx), (pr, lr), (po, _) = self._dev(xp, "xp", self.n_xp), self._dev(records, "records", self.layout["size"]), self._dev(out, "out", _lib.SUMMARY_SIZE)
        if records.shape[0] != xp.shape[0] or out.shape[0] != xp.shape[0] or not out.is_contiguous():
            raise ValueError("xp, records and out must have the same number of rows (out contiguous)")
        check(self._lib.ungar_b200_summaries(self._handle, px, xp.shape[0], lx, pr, lr, po, _torch_stream()))
        return out

    def step(self, xp, records=None, summaries=None):
        """One outer-iteration step: KKT records stay in HBM (``records``: CUDA tensor or None for the handle's
        workspace), per-trajectory summaries ``[B, 32]`` come back where ``xp`` lives (numpy -> host)."""
        xp2, out2, mem, stream, _ = self._buffers(xp, summaries, _lib.SUMMARY_SIZE)
        rec_ptr, rec_ld = self._dev(records, "records", self.layout["size"])
        check(self._lib.ungar_b200_kkt_step(self._handle, self._ptr(xp2), xp2.shape[0], self._ld(xp2), rec_ptr, rec_ld,
                                            self._ptr(out2), mem, stream))
        return out2

    def set_parameters(self, parameters) -> None:
        """Cache the parameter block ``parameters[B, n_par]`` (numpy -> host, torch CUDA tensor -> device) in the handle; ``step_x`` then
        takes only the decision variables (the MPC data flow: parameters change per control cycle, decision variables per iteration)."""
        n_par = self.layout["n_par"]
        if parameters.shape[-1] != n_par:
            raise ValueError(f"parameters have {parameters.shape[-1]} columns, expected {n_par}")
        p2, _, mem, stream, _ = self._buffers(parameters, parameters, n_par)
        check(self._lib.ungar_b200_set_parameters(self._handle, self._ptr(p2), p2.shape[0], self._ld(p2), mem, stream))

    def step_x(self, x, records=None, summaries=None):
        """``step`` with the decision variables ``x[B, n_dec]`` only; the parameters come from ``set_parameters``."""
        if x.shape[-1] < self.layout["n_dec"]:
            raise ValueError(f"x has {x.shape[-1]} columns, expected at least {self.layout['n_dec']}")
        x2, out2, mem, stream, _ = self._buffers(x, summaries, _lib.SUMMARY_SIZE)
        rec_ptr, rec_ld = self._dev(records, "records", self.layout["size"])
        check(self._lib.ungar_b200_kkt_step_x(self._handle, self._ptr(x2), x2.shape[0], self._ld(x2), rec_ptr, rec_ld,
                                              self._ptr(out2), mem, stream))
        return out2

    def qp_solve(self, records, steps=None, multipliers=None, want_multipliers: bool = True):
        """Batched solve of the equality-constrained QP of the records (CUDA tensors): ``(steps[B, n_dec], multipliers[B, m_eq])``.
        Replaces the OSQP call of SoftSQPOptimizer::SolveLocalQPProblem (soft_sqp.hpp:193-233).  The factorisation is fp64; an f32
        model's fp32 records are widened on the device."""
        import torch

        B = records.shape[0]
        if steps is None:
            steps = torch.empty((B, self.layout["n_dec"]), dtype=records.dtype, device=records.device)
        if multipliers is None and want_multipliers:
            multipliers = torch.empty((B, self.layout["m_eq"]), dtype=records.dtype, device=records.device)
        pr, lr = self._dev(records, "records", self.layout["size"])
        ps, ls = self._dev(steps, "steps", self.layout["n_dec"])
        mp, ml = self._dev(multipliers, "multipliers", self.layout["m_eq"])
        if steps.shape[0] != B or (multipliers is not None and multipliers.shape[0] != B):
            raise ValueError("records, steps and multipliers must have the same number of rows")
        check(self._lib.ungar_b200_qp_solve(self._handle, pr, B, lr, ps, ls, mp, ml, _torch_stream()))
        return steps, multipliers

    def sqp_options(self, **overrides) -> "_lib.SqpOptions":
        """Reference defaults (soft_sqp.hpp:44-50, backtracking_line_search.hpp:70-76) with keyword overrides."""
        opts = _lib.SqpOptions()
        check(self._lib.ungar_b200_sqp_options_default(ctypes.byref(opts)))
        for k, v in overrides.items():
            if not hasattr(opts, k):
                raise TypeError(f"unknown SQP option {k!r}")
            setattr(opts, k, v)
        return opts

    def line_search(self, xp, steps, options=None, status=None, info=None):
        """Batched BacktrackingLineSearch::Do (backtracking_line_search.hpp:81-165) on CUDA tensors: updates ``xp[:, :n_dec]``
        in place where a step is accepted and returns ``info[B, 8]`` = (alpha, theta, phi, f | theta0, phi0, f0, grad f . dw)."""
        import torch

        options = options or self.sqp_options()
        if info is None:
            info = torch.empty((xp.shape[0], _lib.LINE_SEARCH_INFO_SIZE), dtype=xp.dtype, device=xp.device)
        (px, lx), (ps, ls) = self._dev(xp, "xp", self.n_xp), self._dev(steps, "steps", self.layout["n_dec"])
        pst, _ = self._dev(status, "status", 2, int32=True)
        pi, _ = self._dev(info, "info", _lib.LINE_SEARCH_INFO_SIZE)
        if steps.shape[0] != xp.shape[0] or info.shape[0] != xp.shape[0] or not info.is_contiguous() or (
                status is not None and (status.shape != (xp.shape[0], 2) or not status.is_contiguous())):
            raise ValueError("steps / info / status must have one contiguous row per trajectory of xp")
        check(self._lib.ungar_b200_line_search(self._handle, px, xp.shape[0], lx, ps, ls, ctypes.byref(options), pst, pi, _torch_stream()))
        return info

    def sqp_solve(self, xp, options=None, want_info: bool = True):
        """SoftSQPOptimizer::Optimize (soft_sqp.hpp:63-109) for a batch, on the device.  ``xp`` (CUDA tensor or numpy array
        ``[B, n_xp]``) is updated in place; returns ``(status[B, 2] int32 = (status, iterations), info[B, 8] | None)``."""
        options = options or self.sqp_options()
        if _is_torch(xp):
            import torch

            B = xp.shape[0]
            status = torch.empty((B, 2), dtype=torch.int32, device=xp.device)
            info = torch.empty((B, _lib.LINE_SEARCH_INFO_SIZE), dtype=xp.dtype, device=xp.device) if want_info else None
            px, lx = self._dev(xp, "xp", self.n_xp)
            check(self._lib.ungar_b200_sqp_solve(self._handle, px, B, lx, ctypes.byref(options),
                                                 status.data_ptr(), info.data_ptr() if want_info else None, MEM_DEVICE,
                                                 _torch_stream()))
            return status, info
        if not (isinstance(xp, np.ndarray) and xp.dtype == self.np_dtype and xp.ndim == 2 and xp.flags.c_contiguous):
            raise ValueError("host xp must be a C-contiguous [B, n_xp] array of the model dtype (it is updated in place)")
        B = xp.shape[0]
        status = np.empty((B, 2), dtype=np.int32)
        info = np.empty((B, _lib.LINE_SEARCH_INFO_SIZE), dtype=self.np_dtype) if want_info else None
        check(self._lib.ungar_b200_sqp_solve(self._handle, xp.ctypes.data, B, xp.shape[1], ctypes.byref(options),
                                             status.ctypes.data, info.ctypes.data if want_info else None, MEM_HOST, None))
        return status, info

    def set_profiling(self, enabled: bool) -> None:
        check(self._lib.ungar_b200_set_profiling(int(enabled)))

    def sweep_times_ms(self) -> list:
        """Device durations (ms) of the sweep-kernel launches since profiling was enabled (synchronises)."""
        buf = (ctypes.c_float * 512)()
        n = ctypes.c_int32()
        check(self._lib.ungar_b200_sweep_times(buf, 512, ctypes.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def launch_count(self) -> int:
        return int(self._lib.ungar_b200_launch_count())

    def compact_map(self) -> np.ndarray:
        """COMPACT handles: int32 ``map[e]`` = offset of compact slot ``e`` in the dense arrangement, -2 for pad slots."""
        if self._c2d is None:
            ptr, n = ctypes.POINTER(ctypes.c_int32)(), ctypes.c_int64()
            check(self._lib.ungar_b200_kkt_compact_map(self._handle, ctypes.byref(ptr), ctypes.byref(n)))
            self._c2d = np.ctypeslib.as_array(ptr, (int(n.value),)).copy()
        return self._c2d

    def to_dense(self, rec):
        """Dense arrangement of COMPACT record(s) (numpy): one scatter through ``compact_map``; slots the compact record does not
        hold are structural zeros.  DENSE handles return ``rec`` unchanged."""
        if not self.compact:
            return rec
        rec = np.asarray(rec)
        m = self.compact_map()
        keep = m >= 0
        out = np.zeros(rec.shape[:-1] + (self.layout["dense_size"],), dtype=rec.dtype)
        out[..., m[keep]] = rec[..., :m.size][..., keep]
        return out

    def split_record(self, rec) -> dict:
        """Views of one record (or a batch of records) by block name, reshaped per ungar_b200_kkt_layout (COMPACT records are
        expanded to the dense arrangement first)."""
        if self.compact and rec.shape[-1] != self.layout["dense_size"]:
            rec = self.to_dense(rec)
        L = self.layout
        N, nx, nu, nz = L["horizon"], L["nx"], L["nu"], L["nz"]
        lead = rec.shape[:-1]

        def cut(off, count, *shape):
            return rec[..., off:off + count].reshape(*lead, *shape)

        out = {"g": cut(L["g"], L["m_eq"], L["m_eq"]), "A": cut(L["A"], N * nx * nz, N, nx, nz),
               "h": cut(L["h"], L["m_ineq"], L["m_ineq"]), "cost": cut(L["cost"], 2, 2),
               "grad": cut(L["grad"], L["n_dec"], L["n_dec"]), "H": cut(L["H"], N * L["tri"], N, L["tri"]),
               "HN": cut(L["HN"], L["tri_terminal"], L["tri_terminal"])}
        if L["legs"]:
            out["C"] = cut(L["C"], N * L["legs"] * 80, N, L["legs"], 4, 20)
        if L["hc_per_node"]:
            out["Hc"] = cut(L["Hc"], (N - 1) * nu, N - 1, nu)
        return out


class SoftSQPOptimizer:
    """Mirror of ``Ungar::SoftSQPOptimizer`` (optimization/soft_sqp.hpp:42-61): same constructor arguments, ``Optimize`` takes
    the NLP problem (a ``Model``) and the flat vector(s) ``xp``.  The barrier (stiffness, epsilon) lives in the model handle,
    where the reference JIT-compiles it into its own Function (soft_sqp.hpp:114-138); ``Optimize`` checks they agree."""

    def __init__(self, verbose: bool = False, constraintViolationMultiplier: float = 1.0, maxIterations: int = 10,
                 stiffness: float = 100.0, epsilon: float = 2e-5):
        self.verbose = verbose
        self.constraintViolationMultiplier = float(constraintViolationMultiplier)
        self.maxIterations = int(maxIterations)
        self.stiffness, self.epsilon = float(stiffness), float(epsilon)
        self.status = None
        self.info = None

    def Optimize(self, nlpProblem: "Model", xp):
        if (nlpProblem.barrier[0], nlpProblem.barrier[1]) != (self.stiffness, self.epsilon):
            raise ValueError("the model was created with a different barrier (stiffness, epsilon) than this optimizer")
        opts = nlpProblem.sqp_options(max_iterations=self.maxIterations,
                                      constraint_violation_multiplier=self.constraintViolationMultiplier)
        single = not _is_torch(xp) and np.ndim(xp) == 1
        buf = np.array(xp, dtype=nlpProblem.np_dtype)[None] if single else xp
        self.status, self.info = nlpProblem.sqp_solve(buf, opts)
        n = nlpProblem.layout["n_dec"]
        return buf[0, :n] if single else buf[:, :n]


class Function:
    """Mirror of ``Ungar::Autodiff::Function`` for one function of a ``Model``."""

    def __init__(self, model: Model, which: int):
        self._m = model
        self._which = which
        v = [ctypes.c_int64() for _ in range(5)]
        check(model._lib.ungar_b200_function_info(model._handle, which, *[ctypes.byref(x) for x in v]))
        self._nx, self._np, self._ny, self._nnz_jac, self._nnz_hes = (int(x.value) for x in v)

    # sizes (function.hpp:351-361)
    def IndependentVariableSize(self) -> int:
        return self._nx

    def ParameterSize(self) -> int:
        return self._np

    def DependentVariableSize(self) -> int:
        return self._ny

    def ImplementsFunction(self) -> bool:
        return True

    def ImplementsJacobian(self) -> bool:
        return True

    def ImplementsHessian(self) -> bool:
        return self._which in (OBJECTIVE, SOFT_INEQUALITIES)

    def _sparsity(self, fn_name):
        rows, cols, nnz = _lib.c_i64_p(), _lib.c_i64_p(), ctypes.c_int64()
        check(getattr(self._m._lib, fn_name)(self._m._handle, self._which, ctypes.byref(rows), ctypes.byref(cols),
                                             ctypes.byref(nnz)))
        n = int(nnz.value)
        return (np.ctypeslib.as_array(rows, (n,)).copy() if n else np.zeros(0, np.int64),
                np.ctypeslib.as_array(cols, (n,)).copy() if n else np.zeros(0, np.int64))

    def JacobianSparsity(self):
        """(rows, cols): row-major, columns ascending within a row (GenericModel::JacobianSparsity)."""
        return self._sparsity("ungar_b200_jacobian_sparsity")

    def HessianSparsity(self):
        """(rows, cols) of the upper triangle (GenericModel::HessianSparsity(0, ...))."""
        return self._sparsity("ungar_b200_hessian_sparsity")

    def _call(self, entry, xp, out, cols):
        m = self._m
        xp2, out2, mem, stream, squeeze = m._buffers(xp, out, max(cols, 1))
        if xp2.shape[-1] != self._nx + self._np:
            raise ValueError(f"xp has {xp2.shape[-1]} entries, expected {self._nx + self._np}")
        check(getattr(m._lib, entry)(m._handle, self._which, m._ptr(xp2), xp2.shape[0], m._ld(xp2), m._ptr(out2),
                                     m._ld(out2), mem, stream))
        out2 = out2[..., :cols]
        return out2[0] if squeeze else out2

    # Evaluate / operator() (function.hpp:180-214)
    def Evaluate(self, xp, y=None):
        return self._call("ungar_b200_forward_zero", xp, y, self._ny)

    __call__ = Evaluate

    def JacobianValues(self, xp, out=None):
        """Nonzeros of dy/dx in JacobianSparsity() order; ``[B, nnz]`` for a batch."""
        return self._call("ungar_b200_sparse_jacobian", xp, out, self._nnz_jac)

    def HessianValues(self, xp, out=None):
        return self._call("ungar_b200_sparse_hessian", xp, out, self._nnz_hes)

    # Jacobian / Hessian as sparse matrices for ONE xp (function.hpp:216-274)
    def Jacobian(self, xp):
        import scipy.sparse as sp

        vals = np.asarray(self.JacobianValues(np.asarray(xp)), dtype=np.float64)
        rows, cols = self.JacobianSparsity()
        return sp.csr_matrix((vals, (rows, cols)), shape=(self._ny, self._nx))

    def Hessian(self, xp):
        """Upper-triangular view, like the reference (function.hpp:232-235)."""
        import scipy.sparse as sp

        vals = np.asarray(self.HessianValues(np.asarray(xp)), dtype=np.float64)
        rows, cols = self.HessianSparsity()
        return sp.csr_matrix((vals, (rows, cols)), shape=(self._nx, self._nx))
