"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU restatement in this directory (``oracle.cpp`` / ``models.hpp`` /
``stage_port.cpp``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py``
(``cpu_baseline`` and ``--impl reference``) may import this package; the product package
``ungar_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)

OBJECTIVE, EQUALITIES, INEQUALITIES = 0, 1, 2


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_fast.so with the Makefile in this directory."""
    args = ["make", "-C", _HERE, "-s"] + (["-B"] if force else [])
    subprocess.run(args, check=True)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_c_dbl_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_c_int_p)


class Oracle:
    """fp64 CPU evaluator of the reference path (values, sparse Jacobian/Hessian, KKT block record)."""

    def __init__(self, fast: bool = False):
        name = "liboracle_fast.so" if fast else "liboracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.oracle_sizes.argtypes = [ctypes.c_int, ctypes.c_int, _c_int_p]
        L.oracle_record_layout.argtypes = [ctypes.c_int, ctypes.c_int, _c_int_p]
        L.oracle_eval.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_dbl_p, _c_dbl_p]
        L.oracle_jacobian.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_dbl_p, ctypes.c_int64,
                                      _c_int_p, _c_int_p, _c_dbl_p]
        L.oracle_jacobian.restype = ctypes.c_int64
        L.oracle_hessian.argtypes = [ctypes.c_int, ctypes.c_int, _c_dbl_p, ctypes.c_int64,
                                     _c_int_p, _c_int_p, _c_dbl_p]
        L.oracle_hessian.restype = ctypes.c_int64
        L.oracle_barrier.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, _c_dbl_p, _c_dbl_p,
                                     _c_dbl_p, _c_dbl_p]
        L.oracle_kkt_record.argtypes = [ctypes.c_int, ctypes.c_int, _c_dbl_p, ctypes.c_double,
                                        ctypes.c_double, _c_dbl_p]
        L.oracle_dynamics.argtypes = [ctypes.c_int, ctypes.c_int, _c_dbl_p, ctypes.c_int, _c_dbl_p]
        L.oracle_approx_exp.argtypes = [_c_dbl_p, _c_dbl_p, _c_dbl_p]
        L.oracle_stage_sweep.argtypes = [ctypes.c_int, ctypes.c_int, _c_dbl_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_double, ctypes.c_double, _c_dbl_p, ctypes.c_int64,
                                         ctypes.c_int]
        L.oracle_stage_sweep.restype = ctypes.c_int
        L.oracle_qp_solve.argtypes = [ctypes.c_int, _c_dbl_p, ctypes.c_int64, ctypes.c_int64, _c_dbl_p, ctypes.c_int64, ctypes.c_int]
        L.oracle_sqp_solve.argtypes = [ctypes.c_int, _c_dbl_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        L.oracle_qp_solve_model.argtypes = [ctypes.c_int] + L.oracle_qp_solve.argtypes
        L.oracle_sqp_solve_model.argtypes = [ctypes.c_int] + L.oracle_sqp_solve.argtypes

    # ------------------------------------------------------------------ sizes / layout
    def sizes(self, model: int, N: int) -> dict:
        v = np.zeros(7, dtype=np.int32)
        if self.lib.oracle_sizes(model, N, _ip(v)) != 0:
            raise ValueError("oracle_sizes failed")
        keys = ["nx", "nu", "N", "n_dec", "n_par", "m_eq", "m_ineq"]
        return {k: int(x) for k, x in zip(keys, v)}

    def record_layout(self, model: int, N: int) -> dict:
        v = np.zeros(15, dtype=np.int32)
        if self.lib.oracle_record_layout(model, N, _ip(v)) != 0:
            raise ValueError("oracle_record_layout failed")
        keys = ["g", "A", "C", "h", "cost", "grad", "H", "HN", "Hc", "size", "nz", "tri", "ntri_N",
                "n_legs", "hc_per_node"]
        return {k: int(x) for k, x in zip(keys, v)}

    # ------------------------------------------------------------------ Function API restated
    def evaluate(self, model: int, fn: int, N: int, xp: np.ndarray) -> np.ndarray:
        s = self.sizes(model, N)
        xp = np.ascontiguousarray(xp, dtype=np.float64)
        assert xp.shape == (s["n_dec"] + s["n_par"],)
        ny = {OBJECTIVE: 1, EQUALITIES: s["m_eq"], INEQUALITIES: s["m_ineq"]}[fn]
        y = np.zeros(ny)
        n = self.lib.oracle_eval(model, fn, N, _dp(xp), _dp(y))
        assert n == ny, (n, ny)
        return y

    def _triplets(self, call):
        nnz = call(0, None, None, None)
        if nnz < 0:
            raise ValueError("oracle derivative call failed")
        rows = np.zeros(nnz, dtype=np.int32)
        cols = np.zeros(nnz, dtype=np.int32)
        vals = np.zeros(nnz)
        call(nnz, _ip(rows), _ip(cols), _dp(vals))
        return rows, cols, vals

    def jacobian(self, model: int, fn: int, N: int, xp: np.ndarray):
        """(rows, cols, vals) of dy/dx, row-major, columns ascending (parameter columns trimmed)."""
        xp = np.ascontiguousarray(xp, dtype=np.float64)
        return self._triplets(lambda cap, r, c, v: self.lib.oracle_jacobian(model, fn, N, _dp(xp), cap, r, c, v))

    def hessian(self, model: int, N: int, xp: np.ndarray):
        """Upper-triangular (rows, cols, vals) of the objective Hessian."""
        xp = np.ascontiguousarray(xp, dtype=np.float64)
        return self._triplets(lambda cap, r, c, v: self.lib.oracle_hessian(model, N, _dp(xp), cap, r, c, v))

    def barrier(self, stiffness: float, epsilon: float, z: np.ndarray):
        z = np.ascontiguousarray(z, dtype=np.float64)
        val = ctypes.c_double(0.0)
        dz, d2z = np.zeros_like(z), np.zeros_like(z)
        self.lib.oracle_barrier(stiffness, epsilon, z.size, _dp(z), ctypes.byref(val), _dp(dz), _dp(d2z))
        return val.value, dz, d2z

    def dynamics(self, model: int, N: int, xp: np.ndarray, k: int) -> np.ndarray:
        xp = np.ascontiguousarray(xp, dtype=np.float64)
        out = np.zeros(13)
        n = self.lib.oracle_dynamics(model, N, _dp(xp), k, _dp(out))
        return out[:n]

    def approx_exp(self, v: np.ndarray):
        """Utils::ApproximateExponentialMap(v).coeffs() (x, y, z, w) and its 4x3 Jacobian."""
        v = np.ascontiguousarray(v, dtype=np.float64)
        q, jac = np.zeros(4), np.zeros(12)
        self.lib.oracle_approx_exp(_dp(v), _dp(q), _dp(jac))
        return q, jac.reshape(4, 3)

    # ------------------------------------------------------------------ KKT block record
    def kkt_record(self, model: int, N: int, xp: np.ndarray, stiffness: float, epsilon: float) -> np.ndarray:
        """Per-trajectory block record assembled from the monolithic sparse matrices (slow, exact)."""
        xp = np.ascontiguousarray(xp, dtype=np.float64)
        rec = np.zeros(self.record_layout(model, N)["size"])
        rc = self.lib.oracle_kkt_record(model, N, _dp(xp), stiffness, epsilon, _dp(rec))
        if rc != 0:
            raise RuntimeError(f"oracle_kkt_record: block decomposition is not lossless (code {rc})")
        return rec

    def stage_sweep(self, model: int, N: int, xp: np.ndarray, stiffness: float, epsilon: float,
                    threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
        """Stage-wise CPU port of the same record for a batch ``xp[B, n_xp]`` (the timed CPU baseline)."""
        xp = np.ascontiguousarray(xp, dtype=np.float64)
        B = xp.shape[0]
        size = self.record_layout(model, N)["size"]
        if out is None:
            out = np.zeros((B, size))
        rc = self.lib.oracle_stage_sweep(model, N, _dp(xp), B, xp.shape[1], stiffness, epsilon, _dp(out),
                                         out.shape[1], threads)
        if rc != 0:
            raise RuntimeError(f"oracle_stage_sweep failed ({rc})")
        return out

    # ------------------------------------------------------------------ CPU port of the consumers of the record (quadruped)
    def qp_solve_port(self, N: int, records: np.ndarray, threads: int = 1, model: int = 2) -> np.ndarray:
        """Stage-wise exact QP solve of block records ``[B, size]`` -> steps ``[B, n_dec]`` (sqp_port.cpp; quadruped by default)."""
        records = np.ascontiguousarray(records, dtype=np.float64)
        B = records.shape[0]
        steps = np.zeros((B, self.sizes(model, N)["n_dec"]))
        if self.lib.oracle_qp_solve_model(model, N, _dp(records), B, records.shape[1], _dp(steps), steps.shape[1], threads) != 0:
            raise RuntimeError("oracle_qp_solve failed")
        return steps

    def sqp_solve_port(self, N: int, xp: np.ndarray, stiffness: float, epsilon: float, multiplier: float, iterations: int,
                       threads: int = 1, model: int = 2):
        """SoftSQPOptimizer::Optimize for a batch of problems on the CPU (sqp_port.cpp): returns (xp_final, status[B, 2])."""
        out = np.array(xp, dtype=np.float64, order="C")
        B = out.shape[0]
        status = np.zeros((B, 2), dtype=np.int32)
        rc = self.lib.oracle_sqp_solve_model(model, N, _dp(out), B, out.shape[1], stiffness, epsilon, multiplier, iterations,
                                             status.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), threads)
        if rc != 0:
            raise RuntimeError("oracle_sqp_solve failed")
        return out, status
