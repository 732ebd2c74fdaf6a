"""Generic tape path on the GPU (ungar_b200_tape_*, csrc/tape_machine.cuh) — parity anchors:

  * the reference's own known answers for Function: test/autodiff/function.test.cpp:33-59 (ApproximateExponentialMap), :70-96
    (y = [p |x|^2, 2 x0^2], J = [[2 p x], [4 x0, 0, 0, 0]]), :120-136 (H = 2 p I), each at 1024 random points like the reference;
  * symbolic differentiation (sympy) of a function that uses every tape operation, first and second order;
  * the reference's OWN MPC lambdas: the tapes that oracle/_ref recorded from the unchanged example sources, evaluated by the
    register machine and compared with the restated oracle and with the hand-written kernels (three independent computations).
"""
import glob
import os

import numpy as np
import pytest

from ungar_b200 import EXAMPLE_BARRIER
from ungar_b200 import autodiff as A
from ungar_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_function_test_known_answers():
    rng = np.random.default_rng(0)
    f = A.MakeFunction(A.Blueprint(lambda xp: [xp[4] * sum(x * x for x in xp[:4]), 2.0 * A.pow(xp[0], 2)], 4, 1, "jacobian_test", A.JACOBIAN))
    xp = rng.uniform(-1, 1, (1024, 5))
    y = f(xp)
    x, p = xp[:, :4], xp[:, 4]
    assert np.allclose(y[:, 0], p * (x * x).sum(1), rtol=1e-14) and np.allclose(y[:, 1], 2 * x[:, 0] ** 2, rtol=1e-14)
    J = f.JacobianValues(xp)  # order (0,0) (0,1) (0,2) (0,3) (1,0)
    assert np.allclose(J[:, :4], 2 * p[:, None] * x, rtol=1e-14, atol=1e-15) and np.allclose(J[:, 4], 4 * x[:, 0], rtol=1e-14)
    Jm = f.Jacobian(xp[3]).toarray()
    assert np.allclose(Jm, np.vstack([2 * p[3] * x[3], [4 * x[3, 0], 0, 0, 0]]))
    h = A.MakeFunction(A.Blueprint(lambda xp: [xp[4] * sum(x * x for x in xp[:4])], 4, 1, "hessian_test", A.ALL))
    H = h.HessianValues(xp)
    assert np.allclose(H, 2 * p[:, None] * np.ones((1, 4)), rtol=1e-13, atol=1e-14)
    assert np.allclose(h.Hessian(0, xp[5]).toarray(), 2 * p[5] * np.eye(4))


def approximate_exponential_map(v):  # Utils::ApproximateExponentialMap (utils/utils.hpp:731-749)
    n = A.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + 2.220446049250313e-16)
    k = A.sin(0.5 * n) / n
    return [v[0] * k, v[1] * k, v[2] * k, A.cos(0.5 * n)]


def test_exponential_map_known_answers():
    f = A.MakeFunction(A.Blueprint(approximate_exponential_map, 3, 0, "exponential_map_test", A.JACOBIAN))
    assert np.allclose(f(np.zeros(3)), [0, 0, 0, 1], atol=1e-8)
    assert np.allclose(f.Jacobian(np.zeros(3)).toarray(), np.vstack([0.5 * np.eye(3), np.zeros((1, 3))]), atol=1e-7)
    v = np.random.default_rng(1).uniform(-1, 1, (1024, 3))
    th = np.linalg.norm(v, axis=1, keepdims=True)
    exact = np.hstack([v / th * np.sin(th / 2), np.cos(th / 2)])
    assert np.allclose(f(v), exact, atol=1e-8)


def zoo(ops, xp):
    """One function through every tape operation; `ops` is the namespace (ungar_b200.autodiff or a sympy adaptor)."""
    a, b, c, d, p = xp
    u = ops.sin(a * b) + ops.cos(c) * ops.exp(-d * d) + ops.sqrt(a * a + 1.5) / (b * b + 2.0)
    w = ops.atan(c * d) + ops.tan(0.3 * a) + ops.log(b * b + c * c + 1.0) + ops.acos(0.5 * ops.sin(d)) + ops.asin(0.4 * ops.cos(a))
    z = ops.atan2(a + 2.0, b * c + 3.0) + ops.pow(a * a + 1.0, 1.7) + ops.pow(b, 3) - ops.abs_(c - 0.1) * d
    return [p * u + w * z, u - z, ops.CondExpGt(a, b, a * a * c, ops.sin(b) * d) + ops.CondExpLe(c, d, c * d * d, a)]


class SympyOps:
    def __init__(self):
        import sympy as sp

        self.sp = sp
        for name in ("sin", "cos", "tan", "exp", "log", "sqrt", "atan", "acos", "asin", "atan2"):
            setattr(self, name, getattr(sp, name))

    def pow(self, x, e): return x ** e  # noqa: E704
    def abs_(self, x): return self.sp.sqrt(x * x)  # noqa: E704  smooth away from 0, derivative sign(x) like CppAD
    def CondExpGt(self, a, b, t, f): return self.sp.Piecewise((t, a > b), (f, True))  # noqa: E704
    def CondExpLe(self, a, b, t, f): return self.sp.Piecewise((t, a <= b), (f, True))  # noqa: E704


def test_every_operation_against_symbolic_derivatives():
    import sympy as sp

    syms = sp.symbols("a b c d p")
    exprs = zoo(SympyOps(), syms)
    J_sym = [[sp.diff(e, s) for s in syms[:4]] for e in exprs]
    f = A.MakeFunction(A.Blueprint(lambda xp: zoo(A, xp), 4, 1, "zoo", A.JACOBIAN))
    rows, cols = f.JacobianSparsity()
    assert rows.size == 12  # dense 3 x 4
    rng = np.random.default_rng(2)
    xp = rng.uniform(-0.9, 0.9, (64, 5))
    y, J = f(xp), f.JacobianValues(xp)
    fy = sp.lambdify(syms, exprs, "numpy")
    fJ = sp.lambdify(syms, J_sym, "numpy")
    for b in range(64):
        assert np.allclose(y[b], np.array(fy(*xp[b]), dtype=float), rtol=1e-12, atol=1e-13)
        ref = np.array(fJ(*xp[b]), dtype=float)
        assert np.allclose(J[b], ref[rows, cols], rtol=1e-11, atol=1e-12), (b, J[b], ref[rows, cols])
    # second order, one scalar function at a time (weights select the dependent)
    for i in range(3):
        g = A.MakeFunction(A.Blueprint(lambda xp, i=i: [zoo(A, xp)[i]], 4, 1, f"zoo{i}", A.ALL))
        hr, hc = g.HessianSparsity()
        H_sym = sp.hessian(exprs[i], syms[:4])
        fH = sp.lambdify(syms, H_sym, "numpy")
        H = g.HessianValues(xp)
        for b in range(64):
            ref = np.array(fH(*xp[b]), dtype=float)
            dense = np.zeros((4, 4))
            dense[hr, hc] = H[b]
            assert np.allclose(dense, np.triu(ref), rtol=1e-10, atol=1e-11), (i, b)
    # the whole weighted sum through the raw tape interface: H(sum_r w_r y_r)
    t = f._tape
    r, c = t.hessian_pattern()
    t.set_hessian_elements(r, c)
    w = np.array([0.3, -1.2, 2.0])
    Hw = t.sparse_hessian(xp[:4], w)
    Hs = sp.lambdify(syms, sp.hessian(sum(wi * e for wi, e in zip(w, exprs)), syms), "numpy")
    for b in range(4):
        ref = np.array(Hs(*xp[b]), dtype=float)
        assert np.allclose(Hw[b], ref[r, c], rtol=1e-10, atol=1e-11)


def test_cppad_semantics_at_kinks_and_branches():
    f = A.MakeFunction(A.Blueprint(lambda xp: [A.abs_(xp[0]), A.CondExpLt(xp[0], xp[1], xp[0] * xp[0], 3.0 * xp[1])], 2, 0, "kinks", A.JACOBIAN))
    J = f.JacobianValues(np.array([[0.0, 1.0], [2.0, 1.0], [-2.0, 1.0]]))  # pattern: (0,0) (1,0) (1,1)
    assert J[0].tolist() == [0.0, 0.0, 0.0]   # abs'(0) = 0; branch x0^2 at x0 = 0
    assert J[1].tolist() == [1.0, 0.0, 3.0]   # the branch is chosen at the EVALUATION point, not at the taping point
    assert J[2].tolist() == [-1.0, -4.0, 0.0]


def test_device_buffers_and_batch_consistency():
    import torch

    f = A.MakeFunction(A.Blueprint(lambda xp: zoo(A, xp), 4, 1, "zoo", A.JACOBIAN))
    xp = np.random.default_rng(3).uniform(-0.9, 0.9, (300, 5))
    d = torch.from_numpy(xp).cuda()
    Jd = f.JacobianValues(d)
    torch.cuda.synchronize()
    Jh = f.JacobianValues(xp)  # (second call of this order: served by the NVRTC-specialised kernel, same op functions, other instruction order)
    assert np.allclose(Jd.cpu().numpy(), Jh, rtol=1e-13, atol=1e-15)
    one = np.stack([f.JacobianValues(xp[b]) for b in (0, 17, 299)])
    assert np.array_equal(one, Jh[[0, 17, 299]])


# ---------------------------------------------------------------------------------------------------------------------------
# The reference's own lambdas, as taped by oracle/_ref from the unchanged example sources
# ---------------------------------------------------------------------------------------------------------------------------
def load_reference_tape(path):
    """Tape file written by oracle/refshim (CppAD-compatible tracing of the reference's lambdas): header, nodes, dependents."""
    raw = open(path, "rb").read()
    magic, nn, nd, ni, flags = np.frombuffer(raw, dtype=np.int64, count=5)
    assert int(magic) == 0x32455041545F4255
    off = 40
    nodes = np.frombuffer(raw, dtype=A.NODE_DTYPE, count=int(nn), offset=off)
    off += int(nn) * 32
    dep_id = np.frombuffer(raw, dtype=np.int32, count=int(nd), offset=off)
    off += int(nd) * 4
    dep_const = np.frombuffer(raw, dtype=np.float64, count=int(nd), offset=off)
    return nodes, int(ni), dep_id, dep_const


def tape_path(config, function):
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "tapes", config, function, "cppad_cg", "*_lib.so"))
    return hits[0] if hits else None


@pytest.mark.parametrize("name,N,batch", [("quadrotor", 30, 16), ("rc_car", 60, 16), ("quadruped", 30, 8), ("quadruped", 100, 2)])
def test_reference_lambda_tapes_on_the_register_machine(oracle, name, N, batch):
    import ungar_b200

    mid = W.MODEL_IDS[name]
    config = f"{name}_N{N}"
    if tape_path(config, f"{name}_mpc_eqs") is None:
        pytest.skip("oracle/_ref tapes were not built (needs /root/reference at build time)")
    s = W.sizes(mid, N)
    nx = s["n_dec"]
    xp = W.synthetic_batch(mid, N, batch, seed=11)
    model = ungar_b200.Model(name, N, dtype="f64", barrier=EXAMPLE_BARRIER[mid])
    hand = {0: model.objective, 1: model.equalityConstraints, 2: model.inequalityConstraints}
    for fn, suffix in ((0, "obj"), (1, "eqs"), (2, "ineqs")):
        nodes, ni, dep_id, dep_const = load_reference_tape(tape_path(config, f"{name}_mpc_{suffix}"))
        assert ni == s["n_xp"]
        t = A.TapeHandle(nodes, ni, dep_id, dep_const)
        y = t.forward_zero(xp)
        ref_y = np.stack([oracle.evaluate(mid, fn, N, xp[b]) for b in range(batch)])
        scale = max(1.0, np.max(np.abs(xp[:, :nx])))
        assert np.max(np.abs(y - ref_y)) <= 1e-12 * max(scale, np.max(np.abs(ref_y)))
        # Jacobian: structural pattern trimmed like function.hpp:529-550 must equal the oracle's, values to rounding
        r, c = t.jacobian_pattern()
        keep = c < nx
        r, c = r[keep], c[keep]
        t.set_jacobian_elements(r, c)
        o_r, o_c, _ = oracle.jacobian(mid, fn, N, xp[0])
        assert np.array_equal(r, o_r) and np.array_equal(c, o_c)
        J = t.sparse_jacobian(xp)
        for b in (0, batch - 1):
            ref_J = oracle.jacobian(mid, fn, N, xp[b])[2]
            assert np.max(np.abs(J[b] - ref_J)) <= 1e-11 * max(1.0, np.max(np.abs(ref_J))), (suffix, b)
        if name == "quadruped" and N == 30 and fn == 1:
            # second calls: this 20 k-instruction tape (713 slots) now runs as four segmented NVRTC kernels whose cross-kernel values
            # travel through the scratch array — same values, same Jacobian
            t.forward_zero(xp), t.sparse_jacobian(xp)   # start the background compiles (these calls are still served by the interpreter)
            info2 = t.wait_specialised()
            assert info2[0]["state"] == 1 and info2[1]["state"] == 1, info2
            y2, J2 = t.forward_zero(xp), t.sparse_jacobian(xp)
            assert np.max(np.abs(y2 - y)) <= 1e-13 * max(scale, np.max(np.abs(y))) and np.max(np.abs(J2 - J)) <= 1e-12 * max(1.0, np.max(np.abs(J)))
        # ... and the hand-written kernels agree with the register machine (same reference-format arrays)
        Jk = hand[fn].JacobianValues(xp[:2])
        hr, hc = hand[fn].JacobianSparsity()
        assert np.array_equal(hr, r) and np.array_equal(hc, c)
        assert np.max(np.abs(Jk - J[:2])) <= 1e-9 * max(1.0, np.max(np.abs(J[:2])))
        if fn == 0:  # objective Hessian, upper triangle of the x-x block (function.hpp:552-574)
            r2, c2 = t.hessian_pattern()
            keep = (r2 < nx) & (c2 < nx) & (c2 >= r2)
            r2, c2 = r2[keep], c2[keep]
            t.set_hessian_elements(r2, c2)
            H = t.sparse_hessian(xp[:2])
            o_r, o_c, o_v = oracle.hessian(mid, N, xp[0])
            got = {(int(i), int(j)): v for i, j, v in zip(r2, c2, H[0])}
            ref = {(int(i), int(j)): v for i, j, v in zip(o_r, o_c, o_v)}
            assert set(ref) <= set(got)  # the structural pattern may only be a superset of the oracle's
            for key, v in got.items():
                assert abs(v - ref.get(key, 0.0)) <= 1e-10 * max(1.0, abs(ref.get(key, 0.0))), key
        info = t.info()
        assert info["slots"] < 0.2 * info["live_nodes"] + 64  # liveness-based slots: the scratch is far smaller than the tape


# ---------------------------------------------------------------------------------------------------------------------------
# The reference's UNCHANGED Function class and example over the product's CppAD-compatible header (tests/build_ref_gpu.py)
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("binary", ["function_tests_gpu", "function_example_gpu"])
def test_reference_sources_run_on_the_gpu_through_the_product_header(binary):
    """function_tests_gpu: known answers of test/autodiff/function.test.cpp:33-142 through include/ungar/autodiff/function.hpp;
    function_example_gpu: example/autodiff/function.example.cpp as it lies (VariableMap + MakeFunction + TestJacobian/TestHessian,
    UNGAR_ASSERT active).  Both evaluate every value and derivative with the register-machine kernels."""
    import subprocess

    exe = os.path.join(ROOT, "tests", "_ref_gpu", binary)
    if not os.path.exists(exe):
        pytest.skip("tests/_ref_gpu was not built (needs /root/reference at build time)")
    proc = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    if binary == "function_tests_gpu":
        assert "all reference known answers reproduced" in proc.stdout


def test_reference_soft_sqp_test_runs_on_the_gpu_through_the_product_headers():
    """test/optimization/soft_sqp.test.cpp:34-111 through the reference's UNCHANGED SoftSQPOptimizer: functions taped and evaluated
    on the device (cppad/cg.hpp), local QPs solved on the device (osqp++.h -> ungar_b200_kkt_solve_csc)."""
    import subprocess

    exe = os.path.join(ROOT, "tests", "_ref_gpu", "soft_sqp_tests_gpu")
    if not os.path.exists(exe):
        pytest.skip("tests/_ref_gpu was not built (needs /root/reference at build time)")
    proc = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "all reference optima reproduced" in proc.stdout


def _mpc_log(exe, env_name, solves):
    """Numbers of every 't = ..., obj = ..., eqs = ...' line an example logs (quadrotor.example.cpp:412-426)."""
    import re
    import subprocess

    proc = subprocess.run([exe], capture_output=True, text=True, timeout=1200, env=dict(os.environ, **{env_name: str(solves)}),
                          cwd=os.path.dirname(exe))
    assert proc.returncode == 0, proc.stdout[-1500:] + proc.stderr[-1500:]
    rows = []
    for line in (proc.stdout + proc.stderr).splitlines():
        if " t = " in line:
            rows.append([float(v) for v in re.findall(r"-?\d+\.\d+", line.split(" t = ", 1)[1])])
    return rows


@pytest.mark.parametrize("example,solves", [("rc_car", 60), ("quadrotor", 60), ("quadruped", 24)])
def test_unchanged_mpc_examples_on_the_gpu_match_the_cpu_build(example, solves):
    """example/mpc/{rc_car,quadrotor,quadruped}.example.cpp compiled AS THEY LIE twice: against oracle/refshim (CPU tape evaluator + CPU sparse
    LU: the oracle) and against the product headers (register machine + device KKT solve).  The closed-loop MPC logs — time, objective,
    constraint violations, tracked outputs, applied inputs, printed with three decimals — must agree step by step."""
    gpu = os.path.join(ROOT, "tests", "_ref_gpu", f"example_{example}_gpu")
    cpu = os.path.join(ROOT, "oracle", "_ref", f"example_{example}_N30")
    if not (os.path.exists(gpu) and os.path.exists(cpu)):
        pytest.skip("reference example binaries were not built (needs /root/reference at build time)")
    got = _mpc_log(gpu, "UNGAR_B200_MAX_QP_SOLVES", solves)
    ref = _mpc_log(cpu, "UNGAR_REF_MAX_SOLVES", solves)
    assert len(ref) >= 5 and len(got) == len(ref), (len(got), len(ref))
    for step, (a, b) in enumerate(zip(got, ref)):
        assert len(a) == len(b) and np.allclose(a, b, rtol=0, atol=2.1e-3), (step, a, b)  # three printed decimals


def test_python_soft_sqp_reaches_the_reference_optima():
    """test/optimization/soft_sqp.test.cpp:34-111 through the Python mirror: taped objective / constraints / barrier Functions on the
    register machine, local QPs through ungar_b200_kkt_solve_csc; optima (3, 1), (2, 1), (1, 1) within isApprox(1e-1)."""
    obj = lambda: A.MakeFunction(A.Blueprint(lambda v: [A.pow(v[0] - 3.0, 2) + A.pow(v[1] - 2.0, 2)], 2, 0, "obj_soft_sqp_test"))  # noqa: E731
    eqs = lambda: A.MakeFunction(A.Blueprint(lambda v: [v[0] - v[1]], 2, 0, "eqs_soft_sqp_test", A.JACOBIAN))  # noqa: E731
    ineqs1 = lambda: A.MakeFunction(A.Blueprint(lambda v: [v[1] - 1.0, -v[0]], 2, 0, "ineqs_1_soft_sqp_test", A.JACOBIAN))  # noqa: E731
    ineqs2 = lambda: A.MakeFunction(A.Blueprint(lambda v: [A.pow(v[0], 2) - v[1] - 3.0, v[1] - 1.0, -v[0]], 2, 0,  # noqa: E731
                                                "ineqs_2_soft_sqp_test", A.JACOBIAN))
    cases = [(A.MakeNLPProblem(obj(), None, ineqs1()), (3.0, 1.0)), (A.MakeNLPProblem(obj(), None, ineqs2()), (2.0, 1.0)),
             (A.MakeNLPProblem(obj(), eqs(), ineqs2()), (1.0, 1.0))]
    for nlp, optimum in cases:
        x = A.SoftSQPOptimizer(False, 1.0, 100, 100.0, 2e-8).Optimize(nlp, np.zeros(2))
        ref = np.array(optimum)
        assert np.linalg.norm(x - ref) <= 1e-1 * min(np.linalg.norm(x), np.linalg.norm(ref)), (x, optimum)


# ---------------------------------------------------------------------------------------------------------------------------
# NVRTC specialisation of hot tapes + content-hashed kernel cache (the reference caches by NAME, function.hpp:420-451)
# ---------------------------------------------------------------------------------------------------------------------------
def test_specialised_kernels_equal_the_interpreter_and_are_cached_by_content(tmp_path, monkeypatch):
    import time

    import torch

    cache = tmp_path / "kernels"
    monkeypatch.setenv("UNGAR_B200_KERNEL_CACHE", str(cache))

    def make(scale):
        # "the same name, another lambda": only `scale` differs between two recordings
        return A.MakeFunction(A.Blueprint(lambda v: [scale * v[4] * sum(x * A.sin(x) for x in v[:4]) + A.pow(v[1], 3), A.sqrt(1.0 + v[0] * v[0]) / (2.0 + A.cos(v[2]))],
                                          4, 1, "same_name", A.JACOBIAN))

    rng = np.random.default_rng(3)
    X = rng.standard_normal((64, 5))
    f = make(1.0)
    y0, J0 = f._tape.forward_zero(X), f._tape.sparse_jacobian(X)          # first calls: the register machine
    assert all(v["state"] == 0 for v in f._tape.special_info().values())
    y1, J1 = f._tape.forward_zero(X), f._tape.sparse_jacobian(X)          # second calls: specialised
    info = f._tape.special_info()
    assert info[0]["state"] == 1 and info[1]["state"] == 1 and not info[1]["from_cache"], info
    assert np.allclose(y1, y0, rtol=1e-14, atol=1e-15) and np.allclose(J1, J0, rtol=1e-13, atol=1e-15)
    assert len(list(cache.glob("*.cubin"))) == 2
    # the same lambda recorded again (a new handle, as after a restart): served from the cache under the same content hash
    g = make(1.0)
    g._tape.sparse_jacobian(X)
    Jg = g._tape.sparse_jacobian(X)
    assert g._tape.special_info()[1] == {"state": 1, "from_cache": True, "key": info[1]["key"]}
    assert np.array_equal(Jg, J1)
    # a CHANGED lambda under the same name: different content hash, compiled afresh, different values (no stale kernel)
    h = make(2.0)
    h._tape.sparse_jacobian(X)
    Jh = h._tape.sparse_jacobian(X)
    hi = h._tape.special_info()[1]
    assert hi["state"] == 1 and not hi["from_cache"] and hi["key"] != info[1]["key"]
    assert not np.allclose(Jh, J1) and len(list(cache.glob("*.cubin"))) == 3
    # UNGAR_B200_NO_NVRTC keeps the interpreter
    monkeypatch.setenv("UNGAR_B200_NO_NVRTC", "1")
    k = make(1.0)
    k._tape.forward_zero(X)
    k._tape.forward_zero(X)
    assert k._tape.special_info()[0]["state"] == -1


def test_specialised_quadrotor_jacobian_meets_the_latency_target(tmp_path, monkeypatch):
    """the quadrotor N = 30 equality Jacobian of the reference's own lambda (tape recorded by oracle/_ref) at batch
    1024 — the interpreter needs ~1.8 ms per call, the specialised kernel must stay under 0.2 ms (target 0.1 ms) and agree with it."""
    import torch

    monkeypatch.setenv("UNGAR_B200_KERNEL_CACHE", str(tmp_path / "kernels"))
    path = tape_path("quadrotor_N30", "quadrotor_mpc_eqs")
    if path is None:
        pytest.skip("oracle/_ref tapes were not built (needs /root/reference at build time)")
    nodes, ni, dep_id, dep_const = load_reference_tape(path)
    t = A.TapeHandle(nodes, ni, dep_id, dep_const)
    n_dec = 13 * 31 + 4 * 30
    rows, cols = t.jacobian_pattern()
    keep = cols < n_dec
    t.set_jacobian_elements(rows[keep], cols[keep])
    xp = torch.from_numpy(W.synthetic_batch(W.QUADROTOR, 30, 1024, seed=3)).cuda()

    def timed(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = t.sparse_jacobian(xp)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    def busy(seconds):  # the tests before this one may have left the GPU idle for tens of seconds (CPU oracle, NVRTC): wake its clocks up
        a = torch.randn(4096, 4096, device="cuda")
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            (a @ a).sum().item()

    import time

    busy(0.5)
    ms_interp, J_interp = timed(1)
    _, J_special = timed(1)  # compiles
    assert t.special_info()[1]["state"] == 1
    busy(0.5)
    timed(50)
    ms_special, J_special = timed(20)
    print(f"\n[tape] quadrotor N=30 equality Jacobian, batch 1024: interpreter {ms_interp:.3f} ms, NVRTC-specialised {ms_special:.3f} ms")
    assert torch.allclose(J_special, J_interp, rtol=1e-12, atol=1e-14)
    assert ms_special < 0.2


def test_long_tapes_run_as_segmented_specialised_kernels(tmp_path, monkeypatch):
    """A tape beyond the single-kernel limit (12 k instructions) is cut into kernels of 6 k instructions; values that cross a cut travel
    through the scratch array (liveness over the cuts).  A chained recurrence with long-lived values, CondExp and a late use of the first
    independents — so that every kind of value crosses several cuts — against the interpreter (first call) for values and Jacobian."""
    monkeypatch.setenv("UNGAR_B200_KERNEL_CACHE", str(tmp_path / "kernels"))
    n = 6

    def chain(v):
        a, b, c = v[0], v[1], v[2]
        keep = [v[3] * v[4], A.sin(v[5])]                 # live from the first to the last segment
        acc = 0.0
        for k in range(850):                             # ~7 instructions per round
            a, b, c = b * 0.999 + 0.001 * A.sin(c), c - 0.002 * a * b, A.CondExpGt(a, b, a, b) * 0.5 + 0.5 * c
            if k % 200 == 0:
                acc = acc + a * keep[0] + b * keep[1]
        return [a + acc, b * keep[0], c + v[0] * keep[1]]

    f = A.MakeFunction(A.Blueprint(chain, n, 0, "long_chain", A.JACOBIAN))
    assert f.tape_info()["live_nodes"] > 12500
    rng = np.random.default_rng(8)
    X = 0.5 + 0.2 * rng.standard_normal((96, n))
    y0, J0 = f._tape.forward_zero(X), f._tape.sparse_jacobian(X)          # interpreter
    yb, Jb = f._tape.forward_zero(X), f._tape.sparse_jacobian(X)          # second calls: a worker thread starts compiling, the interpreter serves
    assert np.array_equal(yb, y0) and np.array_equal(Jb, J0)
    assert all(v["state"] in (1, 2) for o, v in f._tape.special_info().items() if o < 2)
    info = f._tape.wait_specialised()
    assert info[0]["state"] == 1 and info[1]["state"] == 1, info
    y1, J1 = f._tape.forward_zero(X), f._tape.sparse_jacobian(X)          # segmented specialised kernels
    assert np.isfinite(y1).all() and np.isfinite(J1).all()
    assert np.allclose(y1, y0, rtol=1e-12, atol=1e-14) and np.allclose(J1, J0, rtol=1e-11, atol=1e-13)
    # second-order jets (Hessian of w . y): the accumulator crosses the cuts too
    w = np.array([0.7, -1.3, 0.4])
    H0 = f._tape.sparse_hessian(X[:8], w)
    f._tape.sparse_hessian(X[:8], w)
    assert f._tape.wait_specialised()[2]["state"] == 1
    H1 = f._tape.sparse_hessian(X[:8], w)
    assert H1.shape[1] > 0
    assert np.allclose(H1, H0, rtol=1e-10, atol=1e-12)
    monkeypatch.setenv("UNGAR_B200_NO_SEGMENTS", "1")                     # the switch keeps long tapes on the interpreter
    g = A.MakeFunction(A.Blueprint(chain, n, 0, "long_chain", A.JACOBIAN))
    g._tape.forward_zero(X)
    g._tape.forward_zero(X)
    assert g._tape.wait_specialised()[0]["state"] == -1


@pytest.mark.skipif(os.environ.get("UNGAR_B200_RUN_UNVALIDATED") != "1",
                    reason="the reverse sweep is opt-in and has not run on a GPU yet (GPU budget of the round was spent): its generator is validated "
                           "on the CPU by tests/test_tape_host.py::test_reverse_sweep_source_is_consistent_and_compiles")
def test_reverse_sweep_serves_the_gradient_of_scalar_functions(monkeypatch, tmp_path):
    """A scalar function has one dense Jacobian row: forward mode needs one direction per independent.  With UNGAR_B200_REVERSE=1 the
    gradient comes, from the second call on, from ONE generated reverse sweep per vector (csrc/tape.cu::generate_reverse_source,
    special_info()[3]); it must equal the forward-mode Jacobian of the first call."""
    from test_tape_host import _scalar_objective

    monkeypatch.setenv("UNGAR_B200_REVERSE", "1")
    monkeypatch.setenv("UNGAR_B200_KERNEL_CACHE", str(tmp_path / "kernels"))
    n = 24
    f = A.MakeFunction(A.Blueprint(_scalar_objective, n, 0, "scalar_objective", A.JACOBIAN))
    assert f.tape_info()["jacobian_colors"] == n
    rng = np.random.default_rng(4)
    X = 0.3 + 0.4 * rng.random((257, n))
    J0 = f._tape.sparse_jacobian(X)                       # forward mode, n directions per vector (interpreter)
    J1 = f._tape.sparse_jacobian(X)                       # reverse sweep, one thread per vector
    info = f._tape.special_info()
    assert info[3]["state"] == 1 and info[1]["state"] == 0, info
    assert np.isfinite(J1).all() and np.allclose(J1, J0, rtol=1e-12, atol=1e-14)
    import torch

    J2 = f._tape.sparse_jacobian(torch.from_numpy(X).cuda()).cpu().numpy()
    assert np.array_equal(J2, J1)
    monkeypatch.delenv("UNGAR_B200_REVERSE")              # without the switch: the forward path (which then specialises as before)
    g = A.MakeFunction(A.Blueprint(_scalar_objective, n, 0, "scalar_objective", A.JACOBIAN))
    g._tape.sparse_jacobian(X)
    Jg = g._tape.sparse_jacobian(X)
    gi = g._tape.special_info()
    assert gi[3]["state"] == 0 and gi[1]["state"] == 1 and np.allclose(Jg, J0, rtol=1e-12, atol=1e-14)
