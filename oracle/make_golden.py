#!/usr/bin/env python
"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz from oracle/_ref.

The golden vectors are produced by the REFERENCE'S OWN CODE: the unchanged lambdas of example/mpc/*.example.cpp, taped by
the reference's own MakeFunction / Function (include/ungar/autodiff/function.hpp) over the tracing shim of oracle/refshim
(see oracle/build_ref.py).  /root/reference does not exist on the GPU box, so the vectors are committed as small fixtures:
values in full, derivatives as matrix-vector probes (J v, J^T w, H v with seeded v, w) which pin every entry.

While generating, the script also compares the restated oracle (oracle/liboracle.so) with the reference tapes entry by
entry and fails on any disagreement above 1e-11 — this is the pin of the oracle.

Run here (needs /root/reference):  python oracle/build_ref.py && python oracle/make_golden.py
"""
from __future__ import annotations

import ctypes
import glob
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from ungar_b200 import workloads as W  # noqa: E402

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int64)
CONFIGS = [("quadrotor", 30, (100.0, 2e-5)), ("rc_car", 30, (100.0, 1e-2)), ("rc_car", 60, (100.0, 1e-2)),
           ("quadruped", 30, (1.0, 1.0)), ("quadruped", 100, (1.0, 1.0))]


class RefTape:
    """One tape 'library' written by the reference's MakeFunction (evaluated by the shim's GenericModel)."""
    _lib = None

    def __init__(self, path):
        if RefTape._lib is None:
            L = ctypes.CDLL(os.path.join(HERE, "_ref", "libreftape.so"))
            L.reftape_open.restype = ctypes.c_void_p
            L.reftape_open.argtypes = [ctypes.c_char_p]
            L.reftape_info.argtypes = [ctypes.c_void_p, c_ip]
            L.reftape_eval.argtypes = [ctypes.c_void_p, c_dp, c_dp]
            L.reftape_jacobian.argtypes = [ctypes.c_void_p, c_dp, c_ip, c_ip, c_dp]
            L.reftape_hessian.argtypes = [ctypes.c_void_p, c_dp, c_dp, c_ip, c_ip, c_dp]
            RefTape._lib = L
        self.h = RefTape._lib.reftape_open(path.encode())
        assert self.h, path
        info = np.zeros(5, dtype=np.int64)
        RefTape._lib.reftape_info(self.h, info.ctypes.data_as(c_ip))
        self.n_in, self.n_out, self.nnz_j, self.nnz_h, self.n_nodes = (int(v) for v in info)

    def __call__(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n_out)
        RefTape._lib.reftape_eval(self.h, x.ctypes.data_as(c_dp), y.ctypes.data_as(c_dp))
        return y

    def jacobian(self, x, n_cols):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r, c, v = np.zeros(self.nnz_j, np.int64), np.zeros(self.nnz_j, np.int64), np.zeros(self.nnz_j)
        RefTape._lib.reftape_jacobian(self.h, x.ctypes.data_as(c_dp), r.ctypes.data_as(c_ip), c.ctypes.data_as(c_ip),
                                      v.ctypes.data_as(c_dp))
        return sp.csr_matrix((v, (r, c)), shape=(self.n_out, n_cols)), (r, c)

    def hessian(self, x, n):
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = np.ones(self.n_out)
        r, c, v = np.zeros(self.nnz_h, np.int64), np.zeros(self.nnz_h, np.int64), np.zeros(self.nnz_h)
        RefTape._lib.reftape_hessian(self.h, x.ctypes.data_as(c_dp), w.ctypes.data_as(c_dp), r.ctypes.data_as(c_ip),
                                     c.ctypes.data_as(c_ip), v.ctypes.data_as(c_dp))
        return sp.csr_matrix((v, (r, c)), shape=(n, n)), (r, c)


def find(tdir, stem):
    hits = glob.glob(os.path.join(tdir, stem, "cppad_cg", "*_lib.so"))
    assert len(hits) == 1, (tdir, stem, hits)
    return RefTape(hits[0])


def close(a, b, what, tol=1e-11):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    err = float(np.max(np.abs(a - b))) / scale if b.size else 0.0
    assert err < tol, f"oracle != reference tapes for {what}: {err:.3e}"
    return err


def main():
    orc = oracle.Oracle()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, N, (k, eps) in CONFIGS:
        mid = W.MODEL_IDS[name]
        s = orc.sizes(mid, N)
        n, m_eq, m_in = s["n_dec"], s["m_eq"], s["m_ineq"]
        tdir = os.path.join(HERE, "_ref", "tapes", f"{name}_N{N}")
        obj, eqs, ins = find(tdir, f"{name}_mpc_obj"), find(tdir, f"{name}_mpc_eqs"), find(tdir, f"{name}_mpc_ineqs")
        soft = find(tdir, "soft_sqp_relaxed_poly_*")
        assert (obj.n_in, eqs.n_out, ins.n_out, soft.n_in) == (n + s["n_par"], m_eq, m_in, m_in)
        points = np.vstack([W.nominal(mid, N, 0.0)[None], W.synthetic_batch(mid, N, 2, seed=4242)])
        rng = np.random.default_rng(99)
        fix = {"xp": points, "barrier": np.array([k, eps])}
        worst = 0.0
        for i, xp in enumerate(points):
            v, w_eq, w_in = rng.standard_normal(n), rng.standard_normal(m_eq), rng.standard_normal(m_in)
            f, g, h = obj(xp), eqs(xp), ins(xp)
            Jf, _ = obj.jacobian(xp, n)
            Jg, (gr, gc) = eqs.jacobian(xp, n)
            Jh, (hr, hc) = ins.jacobian(xp, n)
            Hf, _ = obj.hessian(xp, n)
            Hfull = Hf + sp.triu(Hf, 1).T
            Z = soft(h)
            dZ, _ = soft.jacobian(h, m_in)
            d2Z, _ = soft.hessian(h, m_in)
            # ---- pin the restated oracle against the reference's own lambdas ------------------------------------
            worst = max(worst, close(orc.evaluate(mid, 0, N, xp), f, "objective"), close(orc.evaluate(mid, 1, N, xp), g, "equalities"),
                        close(orc.evaluate(mid, 2, N, xp), h, "inequalities"))
            Jf_pat = obj.jacobian(xp, n)[1]
            for fn, Jref, rows, pat in ((0, Jf, 1, Jf_pat), (1, Jg, m_eq, (gr, gc)), (2, Jh, m_in, (hr, hc))):
                r, c, vals = orc.jacobian(mid, fn, N, xp)
                Jo = sp.csr_matrix((vals, (r, c)), shape=(rows, n))
                worst = max(worst, close((Jo - Jref).toarray() if rows * n < 4e6 else abs(Jo - Jref).max(), 0 * np.zeros(1), f"Jacobian {fn}"))
                ref_pat = set(zip(pat[0].tolist(), pat[1].tolist()))  # structural, as reported by the reference's model
                assert ref_pat == set(zip(r.tolist(), c.tolist())), "structural pattern differs from the reference tape's"
            r, c, vals = orc.hessian(mid, N, xp)
            Hpat = obj.hessian(xp, n)[1]
            assert set(zip(Hpat[0].tolist(), Hpat[1].tolist())) == set(zip(r.tolist(), c.tolist())), "Hessian pattern"
            Ho = sp.csr_matrix((vals, (r, c)), shape=(n, n))
            worst = max(worst, close(abs(Ho - Hf).max(), np.zeros(1), "objective Hessian"))
            bz, bdz, bd2z = orc.barrier(k, eps, h)
            worst = max(worst, close(bz, Z[0], "barrier"), close(bdz, dZ.toarray()[0], "barrier Jacobian"),
                        close(bd2z, d2Z.diagonal(), "barrier Hessian"))
            assert abs(d2Z - sp.diags(d2Z.diagonal())).max() == 0.0
            # ---- fixture ------------------------------------------------------------------------------------------------
            fix.update({f"f{i}": f, f"g{i}": g, f"h{i}": h, f"gradf{i}": Jf.toarray()[0], f"v{i}": v, f"w_eq{i}": w_eq,
                        f"w_in{i}": w_in, f"Jg_v{i}": Jg @ v, f"JgT_w{i}": Jg.T @ w_eq, f"Jh_v{i}": Jh @ v, f"JhT_w{i}": Jh.T @ w_in,
                        f"Hf_v{i}": Hfull @ v, f"Z{i}": Z, f"dZ{i}": dZ.toarray()[0], f"d2Z{i}": d2Z.diagonal(),
                        f"nnz{i}": np.array([Jg.nnz, Jh.nnz, Hf.nnz, eqs.n_nodes, ins.n_nodes, obj.n_nodes])})
        path = os.path.join(out_dir, f"{name}_N{N}.npz")
        np.savez_compressed(path, **fix)
        nnz_o = [len(orc.jacobian(mid, fn, N, points[1])[2]) for fn in (1, 2)]
        print(f"{name} N={N}: oracle == reference tapes (worst {worst:.2e}); nnz J_g ref {eqs.nnz_j} / oracle {nnz_o[0]}, "
              f"J_h ref {ins.nnz_j} / oracle {nnz_o[1]}; tape nodes eq {eqs.n_nodes}; {os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    main()
