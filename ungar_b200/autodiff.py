"""Generic path behind ``MakeFunction``: arbitrary user lambdas, taped on the host and evaluated on the GPU.

Python mirror of the reference's autodiff front end (include/ungar/autodiff/function.hpp, data_types.hpp):

* ``AD``            — the tracing scalar standing where ``ad_scalar_t = CppAD::AD<CppAD::cg::CG<double>>`` stands
                      (autodiff/data_types.hpp:39-41): records an operation tape with CppAD's folding rules (operations on
                      parameters are folded; ``x * 0``, ``x + 0``, ``x * 1``, ``x / 1``, ``0 / x`` are identities; integer powers
                      are repeated products; ``CondExp*`` differentiates the selected branch; ``abs'(0) = 0``);
* ``Blueprint``     — ``Function::Blueprint`` (function.hpp:44-75): the lambda ``f(xp, y)`` plus sizes, name, enabled derivatives;
* ``MakeFunction``  — ``Autodiff::MakeFunction`` (function.hpp:607-613): where the reference tapes, generates C, runs gcc and
                      dlopens, this tapes and hands the tape to ``ungar_b200_tape_create``;
* ``TapeFunction``  — ``Autodiff::Function`` (function.hpp:77-361) over a tape handle: same sizes, the parameter columns trimmed
                      from the Jacobian pattern (function.hpp:529-550), upper-triangular x-x Hessian of scalar functions
                      (function.hpp:552-574), ``Evaluate / Jacobian / Hessian``; every call also takes a batch ``xp[B, nx + np]``.

All values and derivatives are computed by the register-machine kernels of ``csrc/tape_machine.cuh``; there is no CPU
evaluation path here (tracing records operations, it does not evaluate the function for the caller).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, check

(OP_INDEP, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SQRT, OP_SIN, OP_COS, OP_TAN, OP_ATAN, OP_ACOS, OP_ASIN, OP_EXP,
 OP_LOG, OP_ABS, OP_POW, OP_ATAN2, OP_CLT, OP_CLE, OP_CGT, OP_CGE, OP_CEQ) = range(24)

NODE_DTYPE = np.dtype([("op", "u1"), ("a", "i4"), ("b", "i4"), ("c", "i4"), ("d", "i4"), ("k", "f8")], align=True)
assert NODE_DTYPE.itemsize == 32  # sizeof(ungar_b200_tape_node)

# EnabledDerivatives bitmask (autodiff/data_types.hpp:95-111)
NONE, JACOBIAN, HESSIAN, ALL = 0, 1, 2, 3


class Tape:
    def __init__(self, n_indep: int):
        self.nodes = []  # (op, a, b, c, d, k)
        self.n_indep = n_indep

    def push(self, op, a=-1, b=-1, c=-1, d=-1, k=0.0) -> int:
        self.nodes.append((op, a, b, c, d, k))
        return len(self.nodes) - 1

    def array(self) -> np.ndarray:
        arr = np.zeros(len(self.nodes), dtype=NODE_DTYPE)
        if self.nodes:
            cols = list(zip(*self.nodes))
            for name, col in zip(("op", "a", "b", "c", "d", "k"), cols):
                arr[name] = col
        return arr


_recording: Tape | None = None


class AD:
    """Tracing scalar: ``v`` is the value at the taping point, ``id`` the tape node (-1: a parameter/constant)."""

    __slots__ = ("v", "id")

    def __init__(self, v=0.0, id=-1):
        self.v = float(v)
        self.id = id

    @property
    def variable(self) -> bool:
        return self.id >= 0

    # -- arithmetic with CppAD's "identical" folding rules -----------------------------------------------------------------
    def __add__(self, o):
        o = _ad(o)
        if self.variable and not o.variable and o.v == 0.0:
            return self
        if o.variable and not self.variable and self.v == 0.0:
            return o
        return _binary(OP_ADD, self, o, self.v + o.v)

    __radd__ = __add__

    def __sub__(self, o):
        o = _ad(o)
        if self.variable and not o.variable and o.v == 0.0:
            return self
        return _binary(OP_SUB, self, o, self.v - o.v)

    def __rsub__(self, o):
        return _ad(o).__sub__(self)

    def __mul__(self, o):
        o = _ad(o)
        for x, c in ((self, o), (o, self)):
            if x.variable and not c.variable:
                if c.v == 0.0:
                    return AD(0.0)
                if c.v == 1.0:
                    return x
        return _binary(OP_MUL, self, o, self.v * o.v)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = _ad(o)
        if self.variable and not o.variable and o.v == 1.0:
            return self
        if o.variable and not self.variable and self.v == 0.0:
            return AD(0.0)
        return _binary(OP_DIV, self, o, self.v / o.v if o.v != 0.0 else math.copysign(math.inf, self.v) if self.v else math.nan)

    def __rtruediv__(self, o):
        return _ad(o).__truediv__(self)

    def __neg__(self):
        return _unary(OP_NEG, self, -self.v)

    def __pos__(self):
        return self

    def __pow__(self, e):
        return pow(self, e)

    def __abs__(self):
        return abs_(self)

    # comparisons act on the taping-point values (the reference records with "no_compare_op", function.hpp:466)
    def __lt__(self, o): return self.v < _ad(o).v  # noqa: E704
    def __le__(self, o): return self.v <= _ad(o).v  # noqa: E704
    def __gt__(self, o): return self.v > _ad(o).v  # noqa: E704
    def __ge__(self, o): return self.v >= _ad(o).v  # noqa: E704
    def __float__(self): return self.v  # noqa: E704

    def __repr__(self):
        return f"AD({self.v}, node={self.id})"


def _ad(x) -> AD:
    return x if isinstance(x, AD) else AD(x)


def _node_of(x: AD) -> int:
    return x.id if x.variable else _recording.push(OP_CONST, k=x.v)


def _unary(op, x: AD, value: float) -> AD:
    if not x.variable or _recording is None:
        return AD(value)
    return AD(value, _recording.push(op, x.id))


def _binary(op, a: AD, b: AD, value: float) -> AD:
    if (not a.variable and not b.variable) or _recording is None:
        return AD(value)
    ia, ib = _node_of(a), _node_of(b)
    return AD(value, _recording.push(op, ia, ib))


def _safe(fn, x, default=math.nan):
    try:
        return fn(x)
    except (ValueError, OverflowError):
        return default


def sqrt(x): x = _ad(x); return _unary(OP_SQRT, x, _safe(math.sqrt, x.v))  # noqa: E702
def sin(x): x = _ad(x); return _unary(OP_SIN, x, math.sin(x.v))  # noqa: E702
def cos(x): x = _ad(x); return _unary(OP_COS, x, math.cos(x.v))  # noqa: E702
def tan(x): x = _ad(x); return _unary(OP_TAN, x, math.tan(x.v))  # noqa: E702
def atan(x): x = _ad(x); return _unary(OP_ATAN, x, math.atan(x.v))  # noqa: E702
def acos(x): x = _ad(x); return _unary(OP_ACOS, x, _safe(math.acos, x.v))  # noqa: E702
def asin(x): x = _ad(x); return _unary(OP_ASIN, x, _safe(math.asin, x.v))  # noqa: E702
def exp(x): x = _ad(x); return _unary(OP_EXP, x, _safe(math.exp, x.v, math.inf))  # noqa: E702
def log(x): x = _ad(x); return _unary(OP_LOG, x, _safe(math.log, x.v))  # noqa: E702
def abs_(x): x = _ad(x); return _unary(OP_ABS, x, math.fabs(x.v))  # noqa: E702


def atan2(y, x):
    y, x = _ad(y), _ad(x)
    return _binary(OP_ATAN2, y, x, math.atan2(y.v, x.v))


def pow(x, e):  # noqa: A001  (mirrors CppAD::pow / Utils::Pow, utils.hpp:820-837)
    x = _ad(x)
    if isinstance(e, int) and not isinstance(e, bool):  # CppAD: integer powers by repeated multiplication
        if e < 0:
            return AD(1.0) / pow(x, -e)
        p = AD(1.0)
        for _ in range(e):
            p = p * x
        return p
    e = _ad(e)
    try:
        value = math.pow(x.v, e.v)
    except (ValueError, OverflowError):
        value = math.nan
    return _binary(OP_POW, x, e, value)


def _cond(op, take_true: bool, a, b, t, f) -> AD:
    a, b, t, f = _ad(a), _ad(b), _ad(t), _ad(f)
    sel = t if take_true else f
    if _recording is None or (not a.variable and not b.variable):
        return sel  # decided by parameters: no operation
    if not t.variable and not f.variable and t.v == f.v:
        return sel
    return AD(sel.v, _recording.push(op, _node_of(a), _node_of(b), _node_of(t), _node_of(f)))


def CondExpLt(a, b, t, f): return _cond(OP_CLT, _ad(a).v < _ad(b).v, a, b, t, f)  # noqa: E704
def CondExpLe(a, b, t, f): return _cond(OP_CLE, _ad(a).v <= _ad(b).v, a, b, t, f)  # noqa: E704
def CondExpGt(a, b, t, f): return _cond(OP_CGT, _ad(a).v > _ad(b).v, a, b, t, f)  # noqa: E704
def CondExpGe(a, b, t, f): return _cond(OP_CGE, _ad(a).v >= _ad(b).v, a, b, t, f)  # noqa: E704
def CondExpEq(a, b, t, f): return _cond(OP_CEQ, _ad(a).v == _ad(b).v, a, b, t, f)  # noqa: E704


def record(fn, x0) -> tuple:
    """CppAD::Independent(x) ... CppAD::ADFun(x, y) (function.hpp:456-465): returns (node array, dependents, dependent constants)."""
    global _recording
    x0 = np.asarray(x0, dtype=np.float64)
    if _recording is not None:
        raise RuntimeError("a tape is already being recorded")
    _recording = Tape(x0.size)
    try:
        xs = [AD(float(v), _recording.push(OP_INDEP, i)) for i, v in enumerate(x0)]
        ys = fn(xs)
        ys = [_ad(y) for y in (ys if isinstance(ys, (list, tuple)) else [ys])]
        nodes = _recording.array()
    finally:
        _recording = None
    return nodes, np.array([y.id for y in ys], dtype=np.int32), np.array([y.v for y in ys], dtype=np.float64)


class TapeHandle:
    """Owner of one ``ungar_b200_tape`` (raw tape in, analysed program out).  Host-side analysis only until the first evaluation."""

    def __init__(self, nodes: np.ndarray, n_independent: int, dependents, dependent_constants=None, device: int = 0):
        self._lib = _lib.load()
        nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        deps = np.ascontiguousarray(dependents, dtype=np.int32)
        consts = np.ascontiguousarray(dependent_constants if dependent_constants is not None else np.zeros(deps.size), dtype=np.float64)
        handle = ctypes.c_void_p()
        check(self._lib.ungar_b200_tape_create(nodes.ctypes.data, nodes.size, int(n_independent), deps.ctypes.data, consts.ctypes.data,
                                               deps.size, int(device), ctypes.byref(handle)))
        self._h = handle
        self.n_independent, self.n_dependent = int(n_independent), int(deps.size)
        self._nnz_jac = self._nnz_hes = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ungar_b200_tape_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def special_info(self) -> dict:
        """State of the NVRTC-specialised kernels per order (0 values, 1 Jacobian, 2 Hessian, 3 reverse sweep of a scalar function): ``state`` 0 not tried / 1 specialised /
        -1 interpreter / 2 compiling in the background, ``from_cache``, ``key`` = content hash the compiled kernel is cached under."""
        buf = (ctypes.c_int64 * 16)()
        check(self._lib.ungar_b200_tape_special_info(self._h, buf))
        return {o: {"state": int(buf[4 * o]), "from_cache": bool(buf[4 * o + 1]), "key": (int(buf[4 * o + 3]) << 32) | int(buf[4 * o + 2])}
                for o in range(4)}  # 3: the reverse sweep of a scalar function

    def wait_specialised(self) -> dict:
        """Blocks until no background compile of this tape's kernels is in flight (long tapes compile on a worker thread while the
        interpreter serves); returns special_info()."""
        check(self._lib.ungar_b200_tape_special_wait(self._h))
        return self.special_info()

    def kernel_source(self, order: int = 0):
        """(CUDA source the NVRTC path generates for ``order``, number of kernels) — host-only, nothing is compiled."""
        need, parts = ctypes.c_int64(), ctypes.c_int32()
        check(self._lib.ungar_b200_tape_kernel_source(self._h, int(order), None, 0, ctypes.byref(need), ctypes.byref(parts)))
        buf = ctypes.create_string_buffer(int(need.value))
        check(self._lib.ungar_b200_tape_kernel_source(self._h, int(order), buf, int(need.value), None, None))
        return buf.value.decode(), int(parts.value)

    def info(self) -> dict:
        buf = (ctypes.c_int64 * 6)()
        check(self._lib.ungar_b200_tape_info(self._h, buf))
        return dict(zip(("independents", "dependents", "live_nodes", "slots", "jacobian_colors", "hessian_directions"), map(int, buf)))

    def _pattern(self, entry):
        rows, cols, nnz = _lib.c_i64_p(), _lib.c_i64_p(), ctypes.c_int64()
        check(getattr(self._lib, entry)(self._h, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(nnz)))
        n = int(nnz.value)
        if n == 0:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        return np.ctypeslib.as_array(rows, (n,)).copy(), np.ctypeslib.as_array(cols, (n,)).copy()

    def jacobian_pattern(self):
        """GenericModel::JacobianSparsitySet over ALL independents (rows, cols), row-major."""
        return self._pattern("ungar_b200_tape_jacobian_pattern")

    def hessian_pattern(self):
        """GenericModel::HessianSparsitySet: full symmetric pattern, union over the dependents."""
        return self._pattern("ungar_b200_tape_hessian_pattern")

    def set_jacobian_elements(self, rows, cols):
        rows, cols = np.ascontiguousarray(rows, dtype=np.int64), np.ascontiguousarray(cols, dtype=np.int64)
        check(self._lib.ungar_b200_tape_set_jacobian_elements(self._h, rows.ctypes.data, cols.ctypes.data, rows.size))
        self._nnz_jac = rows.size

    def set_hessian_elements(self, rows, cols):
        rows, cols = np.ascontiguousarray(rows, dtype=np.int64), np.ascontiguousarray(cols, dtype=np.int64)
        check(self._lib.ungar_b200_tape_set_hessian_elements(self._h, rows.ctypes.data, cols.ctypes.data, rows.size))
        self._nnz_hes = rows.size

    # ---- evaluation (numpy = host buffers, torch CUDA tensors = device buffers on the current stream) ---------------------------
    def _run(self, entry, x, n_out, extra=()):
        if type(x).__module__.startswith("torch"):
            import torch

            if not x.is_cuda or x.dtype != torch.float64:
                raise ValueError("torch inputs must be float64 CUDA tensors")
            x2 = x.unsqueeze(0) if x.dim() == 1 else x
            if x2.dim() != 2 or x2.shape[1] != self.n_independent:
                raise ValueError(f"x has shape {tuple(x.shape)}, expected [B, {self.n_independent}]")
            if x2.stride(-1) != 1:
                raise ValueError("x rows must be contiguous")
            out = torch.empty((x2.shape[0], max(n_out, 1)), dtype=torch.float64, device=x.device)
            ld_x = max(int(x2.stride(0)), self.n_independent)  # a size-1 leading dimension may report any stride
            check(getattr(self._lib, entry)(self._h, x2.data_ptr(), *extra, x2.shape[0], ld_x, out.data_ptr(), out.stride(0),
                                            MEM_DEVICE, torch.cuda.current_stream().cuda_stream))
            out = out[:, :n_out]
            return out[0] if x.dim() == 1 else out
        x2 = np.ascontiguousarray(x, dtype=np.float64)
        squeeze = x2.ndim == 1
        if squeeze:
            x2 = x2[None]
        if x2.shape[1] != self.n_independent:
            raise ValueError(f"x has {x2.shape[1]} entries, expected {self.n_independent}")
        out = np.empty((x2.shape[0], max(n_out, 1)))
        check(getattr(self._lib, entry)(self._h, x2.ctypes.data, *extra, x2.shape[0], x2.shape[1], out.ctypes.data, out.shape[1], MEM_HOST,
                                        None))
        out = out[:, :n_out]
        return out[0] if squeeze else out

    def forward_zero(self, x):
        return self._run("ungar_b200_tape_forward_zero", x, self.n_dependent)

    def sparse_jacobian(self, x):
        if self._nnz_jac is None:
            r, c = self.jacobian_pattern()
            self.set_jacobian_elements(r, c)
        return self._run("ungar_b200_tape_sparse_jacobian", x, self._nnz_jac)

    def sparse_hessian(self, x, weights=None):
        if self._nnz_hes is None:
            r, c = self.hessian_pattern()
            self.set_hessian_elements(r, c)
        w = np.ones(self.n_dependent) if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._w_keepalive = w
        return self._run("ungar_b200_tape_sparse_hessian", x, self._nnz_hes, extra=(w.ctypes.data,))


class Blueprint:
    """Function::Blueprint (function.hpp:44-75): ``functionImpl(xp) -> y`` (a list of AD / floats), sizes, name, enabled derivatives.
    Like the reference (:53-58) the lambda is evaluated once, on a fixed pseudo-random point, to learn the number of dependents."""

    def __init__(self, functionImpl, independentVariableSize: int, parameterSize: int, name: str = "function",
                 enabledDerivatives: int = ALL):
        self.functionImpl = functionImpl
        self.independentVariableSize = int(independentVariableSize)
        self.parameterSize = int(parameterSize)
        self.name = name
        self.enabledDerivatives = enabledDerivatives
        n = self.independentVariableSize + self.parameterSize
        self.tapingPoint = 0.25 + 0.5 * np.random.default_rng(20240807).random(n)


class TapeFunction:
    """Mirror of Ungar::Autodiff::Function (function.hpp:77-361) for an arbitrary taped lambda."""

    def __init__(self, blueprint: Blueprint, device: int = 0, _recorded=None):
        bp = blueprint
        self._nx, self._np = bp.independentVariableSize, bp.parameterSize
        # `_recorded`: a tape saved by save() — the counterpart of the reference loading an existing model library instead of taping
        # the lambda again (function.hpp:420-451), except that what is stored is the tape itself, so nothing can go stale
        nodes, deps, consts = _recorded if _recorded is not None else record(bp.functionImpl, bp.tapingPoint)
        self._recorded = (nodes, deps, consts)
        self._tape = TapeHandle(nodes, self._nx + self._np, deps, consts, device)
        self._ny = deps.size
        self.name = bp.name
        self._enabled = int(bp.enabledDerivatives)
        self._jac = bool(bp.enabledDerivatives & JACOBIAN)
        self._hes = bool(bp.enabledDerivatives & HESSIAN) and self._ny == 1  # scalar functions only (function.hpp:136-137)
        self._jr = self._jc = self._hr = self._hc = np.zeros(0, np.int64)
        if self._jac:  # parameter columns trimmed (function.hpp:529-550)
            r, c = self._tape.jacobian_pattern()
            keep = c < self._nx
            self._jr, self._jc = r[keep], c[keep]
            self._tape.set_jacobian_elements(self._jr, self._jc)
        if self._hes:  # x-x block, upper triangle (function.hpp:552-574)
            r, c = self._tape.hessian_pattern()
            keep = (r < self._nx) & (c < self._nx) & (c >= r)
            self._hr, self._hc = r[keep], c[keep]
            self._tape.set_hessian_elements(self._hr, self._hc)

    def save(self, path: str) -> None:
        """Writes the recorded tape (nodes, dependents, sizes) to a compressed .npz; load() rebuilds the function without the lambda."""
        nodes, deps, consts = self._recorded
        np.savez_compressed(path, nodes=nodes, dependents=deps, dependent_constants=consts, sizes=np.array([self._nx, self._np, self._enabled]),
                            name=np.array(self.name))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "TapeFunction":
        z = np.load(path)
        nx, npar, enabled = (int(v) for v in z["sizes"])
        bp = Blueprint.__new__(Blueprint)
        bp.functionImpl, bp.independentVariableSize, bp.parameterSize = None, nx, npar
        bp.name, bp.enabledDerivatives, bp.tapingPoint = str(z["name"]), enabled, None
        return cls(bp, device, _recorded=(z["nodes"].astype(NODE_DTYPE), z["dependents"], z["dependent_constants"]))

    def IndependentVariableSize(self): return self._nx  # noqa: E704
    def ParameterSize(self): return self._np  # noqa: E704
    def DependentVariableSize(self): return self._ny  # noqa: E704
    def ImplementsFunction(self): return True  # noqa: E704
    def ImplementsJacobian(self): return self._jac  # noqa: E704
    def ImplementsHessian(self): return self._hes  # noqa: E704
    def JacobianSparsity(self): return self._jr.copy(), self._jc.copy()  # noqa: E704
    def HessianSparsity(self): return self._hr.copy(), self._hc.copy()  # noqa: E704
    def tape_info(self): return self._tape.info()  # noqa: E704

    def Evaluate(self, xp):
        return self._tape.forward_zero(xp)

    __call__ = Evaluate

    def JacobianValues(self, xp):
        if not self._jac:
            raise _lib.UngarB200Error(_lib.EUNSUPPORTED, "the Jacobian was not enabled in the blueprint")
        return self._tape.sparse_jacobian(xp)

    def HessianValues(self, xp, dependentVariableIndex: int = 0):
        if not self._hes:
            raise _lib.UngarB200Error(_lib.EUNSUPPORTED, "the Hessian is implemented only for scalar functions (function.hpp:136-137)")
        w = np.zeros(self._ny)
        w[dependentVariableIndex] = 1.0
        return self._tape.sparse_hessian(xp, w)

    def Jacobian(self, xp):
        import scipy.sparse as sp

        return sp.csr_matrix((np.asarray(self.JacobianValues(np.asarray(xp))), (self._jr, self._jc)), shape=(self._ny, self._nx))

    def Hessian(self, dependentVariableIndex: int, xp):
        """Upper-triangular view, like the reference (function.hpp:232-258)."""
        import scipy.sparse as sp

        return sp.csr_matrix((np.asarray(self.HessianValues(np.asarray(xp), dependentVariableIndex)), (self._hr, self._hc)),
                             shape=(self._nx, self._nx))


def MakeFunction(blueprint: Blueprint, recompileLibraries: bool = False, device: int = 0) -> TapeFunction:
    """Autodiff::MakeFunction (function.hpp:607-613).  ``recompileLibraries`` is accepted for signature parity: nothing is compiled,
    so there is no stale-library hazard (the reference caches by NAME only, function.hpp:420-451)."""
    return TapeFunction(blueprint, device)


# ---------------------------------------------------------------------------------------------------------------------------
# The optimiser over taped functions: mirror of Ungar::SoftSQPOptimizer for ARBITRARY NLP problems
# ---------------------------------------------------------------------------------------------------------------------------
class RelaxedPolyBarrierFunction:
    """optimization/soft_inequality_constraint.hpp:131-205: coefficients (:136-145) and the taped evaluation (:181-190)."""

    def __init__(self, rhs: float, stiffness: float = 1.0, epsilon: float = 2e-5):
        self.rhs, self.eps = rhs, epsilon
        self.a1 = stiffness
        self.b1 = -0.5 * self.a1 * epsilon
        self.c1 = -1.0 / 3.0 * (-self.b1 - self.a1 * epsilon) * epsilon - 0.5 * self.a1 * epsilon ** 2 - self.b1 * epsilon
        self.a2 = (-self.b1 - self.a1 * epsilon) / epsilon ** 2
        self.b2, self.c2, self.d2 = self.a1, self.b1, self.c1

    def Evaluate(self, lhs):
        """Sum over the entries of ``lhs`` (a list of AD) of the piecewise barrier of ``lhs_i - rhs``."""
        total = AD(0.0)
        for v in lhs:
            x = _ad(v) - self.rhs
            quad = 0.5 * self.a1 * pow(x, 2) + self.b1 * x + self.c1
            cubic = 1.0 / 3.0 * self.a2 * pow(x, 3) + 0.5 * self.b2 * pow(x, 2) + self.c2 * x + self.d2
            total = total + CondExpLt(x, AD(0.0), quad, CondExpLt(x, AD(self.eps), cubic, AD(0.0)))
        return total


class NLPProblem:
    """optimization/concepts.hpp:153-161; ``None`` stands for ``hana::nothing`` (MakeNLPProblem overloads, :190-262)."""

    def __init__(self, objective, equalityConstraints=None, inequalityConstraints=None):
        self.objective, self.equalityConstraints, self.inequalityConstraints = objective, equalityConstraints, inequalityConstraints


def MakeNLPProblem(objective, equalityConstraints=None, inequalityConstraints=None) -> NLPProblem:
    return NLPProblem(objective, equalityConstraints, inequalityConstraints)


class SoftSQPOptimizer:
    """Ungar::SoftSQPOptimizer (optimization/soft_sqp.hpp:42-109) for NLP problems made of TapeFunctions.  The control flow of
    Optimize and of BacktrackingLineSearch::Do (backtracking_line_search.hpp:81-165) runs on the host exactly as in the reference;
    every function value, Jacobian, Hessian and the local QP (ungar_b200_kkt_solve_csc) are evaluated on the device.  For the three
    reference MPC problems use ungar_b200.SoftSQPOptimizer / Model.sqp_solve instead: there the loop itself runs on the device."""

    def __init__(self, verbose: bool = False, constraintViolationMultiplier: float = 1.0, maxIterations: int = 10,
                 stiffness: float = 100.0, epsilon: float = 2e-5, device: int = 0):
        self.verbose, self.mult, self.maxIterations = verbose, float(constraintViolationMultiplier), int(maxIterations)
        self.stiffness, self.epsilon, self.device = float(stiffness), float(epsilon), device
        self._soft = None
        self.iterations = 0
        self.lineSearchParameters = dict(alphaMin=1e-4, thetaMin=1e-6, thetaMax=1e-2, eta=1e-4, gammaPhi=1e-6, gammaTheta=1e-6,
                                         gammaAlpha=0.5)

    def MakeSoftInequalityConstraintFunction(self, nlp: NLPProblem) -> TapeFunction:
        """soft_sqp.hpp:114-138: Zsoft(z) = RelaxedPolyBarrierFunction{0, stiffness, epsilon}.Evaluate(-z), its own taped Function."""
        m = nlp.inequalityConstraints.DependentVariableSize()
        barrier = RelaxedPolyBarrierFunction(0.0, self.stiffness, self.epsilon)
        return MakeFunction(Blueprint(lambda z: [barrier.Evaluate([-v for v in z])], m, 0,
                                      f"soft_sqp_relaxed_poly__sz_{m}_k_{self.stiffness}_eps_{self.epsilon}", ALL), device=self.device)

    # -- pieces of AssembleOSQPInstance (soft_sqp.hpp:141-158, :236-264) -------------------------------------------------------
    def _soft_value(self, nlp, xp) -> float:
        if nlp.inequalityConstraints is None:
            return 0.0
        return float(self._soft(nlp.inequalityConstraints(xp))[0])

    def _qp(self, nlp, xp):
        import scipy.sparse as sp

        n = nlp.objective.IndependentVariableSize()
        Hu = nlp.objective.Hessian(0, xp)
        P = Hu + sp.identity(n) * 1e-6                       # upper triangle, as handed to OSQP
        grad_f = np.asarray(nlp.objective.Jacobian(xp).todense()).ravel()
        q = grad_f.copy()
        if nlp.inequalityConstraints is not None:
            z = nlp.inequalityConstraints(xp)
            Jh = nlp.inequalityConstraints.Jacobian(xp)
            q += (self._soft.Jacobian(z) @ Jh).toarray().ravel()
            GN = Jh.T @ self._soft.Hessian(0, z) @ Jh        # the soft Hessian is diagonal: its upper triangle is all of it
            P = P + sp.triu(GN)
        if nlp.equalityConstraints is not None:
            A = nlp.equalityConstraints.Jacobian(xp)
            b = -nlp.equalityConstraints(xp)
        else:
            A, b = sp.csr_matrix((0, n)), np.zeros(0)
        return sp.csc_matrix(P), q, sp.csc_matrix(A), b, grad_f

    def _solve_qp(self, P, q, A, b):
        lib = _lib.load()
        n, m = P.shape[0], A.shape[0]
        P.sort_indices()
        A.sort_indices()
        x, y = np.zeros(n), np.zeros(max(m, 1))
        pc, pr, pv = P.indptr.astype(np.int32), P.indices.astype(np.int32), np.ascontiguousarray(P.data, dtype=np.float64)
        ac, ar, av = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        check(lib.ungar_b200_kkt_solve_csc(n, m, pc.ctypes.data, pr.ctypes.data, pv.ctypes.data, q.ctypes.data,
                                           ac.ctypes.data if m else None, ar.ctypes.data if m else None, av.ctypes.data if m else None,
                                           b.ctypes.data if m else None, 1e-9, 1e-9, x.ctypes.data, y.ctypes.data if m else None,
                                           self.device))
        return x

    def _line_search(self, grad, dw, phi, theta, w):
        """BacktrackingLineSearch::Do (backtracking_line_search.hpp:81-165)."""
        p = self.lineSearchParameters
        proj = float(np.sum(grad * dw))
        alpha, th0, ph0 = 1.0, theta(w), phi(w)
        while alpha >= p["alphaMin"]:
            wn = w + alpha * dw
            thn, phn = theta(wn), phi(wn)
            if thn > p["thetaMax"]:
                ok = thn < (1.0 - p["gammaTheta"]) * th0
            elif max(th0, thn) < p["thetaMin"] and proj < 0.0:
                ok = phn < ph0 + p["eta"] * alpha * proj
            else:
                ok = phn < (1.0 - p["gammaPhi"]) * ph0 or thn < (1.0 - p["gammaTheta"]) * th0
            if ok:
                return True, wn
            alpha *= p["gammaAlpha"]
        return False, w

    def Optimize(self, nlp: NLPProblem, xp):
        """soft_sqp.hpp:63-109.  Returns the optimised decision variables (the reference returns _cache.xp.head(n))."""
        xp = np.array(xp, dtype=np.float64)
        n = nlp.objective.IndependentVariableSize()
        assert xp.size == n + nlp.objective.ParameterSize()
        if nlp.inequalityConstraints is not None and self._soft is None:
            self._soft = self.MakeSoftInequalityConstraintFunction(nlp)

        def full(x):
            v = xp.copy()
            v[:n] = x
            return v

        def phi(x):
            return float(nlp.objective(full(x))[0]) + self._soft_value(nlp, full(x))

        def theta(x):
            if nlp.equalityConstraints is None:
                return 0.0
            g = nlp.equalityConstraints(full(x))
            return self.mult * float(np.sqrt(np.dot(g, g)))

        self.iterations = 0
        for _ in range(self.maxIterations):
            self.iterations += 1
            objective = float(nlp.objective(xp)[0])
            P, q, A, b, grad_f = self._qp(nlp, xp)
            d = self._solve_qp(P, q, A, b)
            accepted, w = self._line_search(grad_f, d, phi, theta, xp[:n].copy())
            if not accepted:
                break
            xp[:n] = w
            diff = float(nlp.objective(xp)[0]) - objective
            if diff < 0.0 and abs(diff) < 1e-6:
                break
        return xp[:n].copy()
