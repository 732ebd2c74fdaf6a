"""The timed CPU baseline of kind "codegen" (oracle/codegen_baseline.py: straight-line C generated from the reference's own tapes, the
stand-in for CppADCodeGen's output, function.hpp:453-522) must compute what the oracle computes, or timing it means nothing.
Runs on the CPU; skipped when oracle/_ref (tapes / prebuilt libraries) is absent."""
import os

import numpy as np
import pytest

from ungar_b200 import EXAMPLE_BARRIER
from ungar_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def baseline(name, N):
    from oracle import codegen_baseline as CG

    if not os.path.exists(os.path.join(CG.OUT, f"{name}_N{N}.npz")):  # never built inside a test: the quadruped takes ~30 minutes of gcc
        pytest.skip("oracle/_ref/codegen libraries not built (python oracle/build_ref.py; python oracle/codegen_baseline.py)")
    return CG.Baseline(name, N)


@pytest.mark.parametrize("name,N", [("quadrotor", 30), ("rc_car", 60), ("quadruped", 100)])
def test_generated_code_equals_the_oracle(oracle, name, N):
    mid = W.MODEL_IDS[name]
    bl = baseline(name, N)
    m = bl.meta
    k, eps = EXAMPLE_BARRIER[mid]
    xp = W.synthetic_batch(mid, N, 3, seed=11, perturb_params=True)
    xp[1, m["n_dec"] // 2:int(m["n_dec"])] *= 1.3  # push some inequalities into the barrier's active region
    _, out = bl.run(xp, 2)
    rec = oracle.stage_sweep(mid, N, xp, k, eps)
    L = oracle.record_layout(mid, N)
    for b in range(3):
        for fn, fid in (("obj", 0), ("eqs", 1), ("ineqs", 2)):
            y = oracle.evaluate(mid, fid, N, xp[b])
            got = out[b, int(m[fn + "_y"]):int(m[fn + "_y"]) + int(m[fn + "_ny"])]
            assert np.allclose(got, y, rtol=1e-12, atol=1e-13), fn
            rows, cols, vals = oracle.jacobian(mid, fid, N, xp[b])
            assert np.array_equal(rows, m[fn + "_rows"]) and np.array_equal(cols, m[fn + "_cols"]), f"structural pattern of {fn}"
            gv = out[b, int(m[fn + "_jac"]):int(m[fn + "_jac"]) + int(m[fn + "_nnz"])]
            assert np.allclose(gv, vals, rtol=1e-12, atol=1e-13), fn
        rows, cols, vals = oracle.hessian(mid, N, xp[b])
        hp = m["obj_hes_pattern"]
        assert np.array_equal(rows, hp[:, 0]) and np.array_equal(cols, hp[:, 1])
        assert np.allclose(out[b, int(m["obj_hes"]):int(m["obj_hes"]) + int(m["obj_nnz_hes"])], vals, rtol=1e-12, atol=1e-13)
        # the assembly on top: q = grad f + J_h^T dZ and the barrier value, against the oracle's block record
        n_dec = int(m["n_dec"])
        assert np.allclose(out[b, int(m["q"]):int(m["q"]) + n_dec], rec[b, L["grad"]:L["grad"] + n_dec], rtol=1e-11, atol=1e-12)
        assert np.isclose(out[b, int(m["z"])], rec[b, L["cost"] + 1], rtol=1e-11, atol=1e-12)
