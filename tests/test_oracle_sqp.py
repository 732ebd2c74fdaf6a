"""Pins the CPU restatement of the outer loop (oracle/sqp_reference.py) against the reference's own golden vectors for it:
test/optimization/soft_sqp.test.cpp:34-111 — three small NLPs whose optima (3, 1), (2, 1), (1, 1) SoftSQPOptimizer{false, 1.0,
100, 100.0, 2e-8} must reach within isApprox(1e-1) from x = 0 — and the acceptance rules of
backtracking_line_search.hpp:118-148 on hand-built cases."""
import numpy as np
import pytest

from oracle import sqp_reference as S


def problems(oracle):
    f = lambda x: (x[0] - 3.0) ** 2 + (x[1] - 2.0) ** 2  # noqa: E731  soft_sqp.test.cpp:43-46
    grad = lambda x: np.array([2 * (x[0] - 3.0), 2 * (x[1] - 2.0)])  # noqa: E731
    hess = lambda x: 2.0 * np.eye(2)  # noqa: E731
    g = lambda x: np.array([x[0] - x[1]])  # noqa: E731  :51
    Jg = lambda x: np.array([[1.0, -1.0]])  # noqa: E731
    h1 = lambda x: np.array([x[1] - 1.0, -x[0]])  # noqa: E731  :58-60
    Jh1 = lambda x: np.array([[0.0, 1.0], [-1.0, 0.0]])  # noqa: E731
    h2 = lambda x: np.array([x[0] ** 2 - x[1] - 3.0, x[1] - 1.0, -x[0]])  # noqa: E731  :81-85
    Jh2 = lambda x: np.array([[2 * x[0], -1.0], [0.0, 1.0], [-1.0, 0.0]])  # noqa: E731
    kw = dict(stiffness=100.0, epsilon=2e-8)
    return [(S.DenseProblem(oracle, 2, f, grad, hess, None, None, h1, Jh1, **kw), (3.0, 1.0)),
            (S.DenseProblem(oracle, 2, f, grad, hess, None, None, h2, Jh2, **kw), (2.0, 1.0)),
            (S.DenseProblem(oracle, 2, f, grad, hess, g, Jg, h2, Jh2, **kw), (1.0, 1.0))]


def test_soft_sqp_reaches_the_reference_optima(oracle):
    for i, (prob, optimum) in enumerate(problems(oracle)):
        x, status, iterations, log = S.soft_sqp_loop(prob, np.zeros(2), multiplier=1.0, max_iterations=100)
        ref = np.array(optimum)
        # Eigen isApprox(1e-1): |x - ref| <= 1e-1 * min(|x|, |ref|)
        assert np.linalg.norm(x - ref) <= 1e-1 * min(np.linalg.norm(x), np.linalg.norm(ref)), (i, x, status, iterations)
        assert iterations <= 100 and status in (S.RUNNING, S.CONVERGED, S.LINE_SEARCH_FAILED)


def test_line_search_acceptance_rules():
    p = S.LineSearchParameters()
    w = np.array([1.0])
    # (a) large violation: accepted only when theta decreases by the relative margin (:126-132), whatever phi does
    res, w1 = S.line_search(np.array([1.0]), np.array([-0.5]), lambda x: -x[0], lambda x: abs(x[0]), w, p)
    assert res.accepted and res.alpha == 1.0 and w1[0] == 0.5
    # (b) no decrease along the direction: backtracks alpha = 1, 1/2, ... down to alphaMin (14 trials), rejected, w unchanged
    res, w1 = S.line_search(np.array([1.0]), np.array([1.0]), lambda x: x[0], lambda x: abs(x[0]), w, p)
    assert not res.accepted and res.trials == 14 and w1[0] == 1.0 and res.alpha == 0.0
    # (c) feasible both sides and descent direction: Armijo on phi (:133-139)
    phi = lambda x: (x[0] - 0.25) ** 2  # noqa: E731
    res, w1 = S.line_search(np.array([2 * 0.75]), np.array([-3.0]), phi, lambda x: 0.0, w, p)
    # alpha = 1 overshoots to -2 (phi grows), 1/2 lands on the mirror point -1/2 (phi equal: no sufficient decrease), 1/4 passes
    assert res.accepted and res.alpha == 0.25 and res.trials == 3 and w1[0] == 0.25
    # (d) small violation on one side only: either relative decrease is enough (:140-147)
    res, _ = S.line_search(np.array([-1.0]), np.array([1e-3]), lambda x: -x[0], lambda x: 5e-3, w, p)
    assert res.accepted and res.alpha == 1.0


@pytest.mark.parametrize("model,N", [(0, 8), (1, 8), (2, 6)])
def test_monolithic_qp_step_is_a_kkt_point(oracle, model, N):
    """AssembleOSQPInstance restated (soft_sqp.hpp:141-158): the sparse-LU step satisfies the QP's optimality conditions, and for
    the quadruped equals the solve of the block record (oracle/qp_reference.py) that the device kernel is tested against."""
    from oracle import qp_reference as Q
    from ungar_b200 import EXAMPLE_BARRIER
    from ungar_b200 import workloads as W

    k, eps = EXAMPLE_BARRIER[model]
    xp = W.synthetic_batch(model, N, 1, seed=3)[0]
    P, q, A, g, grad_f = S.monolithic_qp(oracle, model, N, xp, k, eps)
    # the quadruped's constraint matrix is rank deficient (swing legs: all-zero contact rows): it needs the quasi-definite delta,
    # which perturbs A d = -g by delta * |lambda|
    delta = 1e-9 if model == W.QUADRUPED else 0.0
    d, lam = S.solve_qp(P, q, A, g, delta=delta)
    assert np.max(np.abs(A @ d + g)) <= 1e-9 * max(1.0, np.max(np.abs(g))) + 2 * delta * np.max(np.abs(lam))
    assert np.max(np.abs(P @ d + q + A.T @ lam)) < 1e-8 * max(1.0, np.max(np.abs(q)))
    assert np.allclose(P.toarray(), P.toarray().T)
    if model == W.QUADRUPED:
        rec = oracle.stage_sweep(model, N, xp[None], k, eps)[0]
        L = oracle.record_layout(model, N)
        s = oracle.sizes(model, N)
        L.update(horizon=N, n_dec=s["n_dec"], m_eq=s["m_eq"])
        d2, _ = Q.kkt_solve(rec, L)
        assert np.max(np.abs(d - d2)) < 1e-6 * np.max(np.abs(d))


@pytest.mark.parametrize("N,perturb", [(6, False), (30, True)])
def test_cpp_qp_port_equals_the_sparse_kkt_solve(oracle, N, perturb):
    """oracle/sqp_port.cpp (the timed CPU baseline of the SQP loop) against the sparse-LU oracle of the same record."""
    from oracle import qp_reference as Q
    from ungar_b200 import workloads as W

    xp = W.synthetic_batch(W.QUADRUPED, N, 3, seed=41, perturb_params=perturb)
    rec = oracle.stage_sweep(W.QUADRUPED, N, xp, 1.0, 1.0)
    steps = oracle.qp_solve_port(N, rec, threads=2)
    L = oracle.record_layout(W.QUADRUPED, N)
    s = oracle.sizes(W.QUADRUPED, N)
    L.update(horizon=N, n_dec=s["n_dec"], m_eq=s["m_eq"])
    for b in range(3):
        d_ref, _ = Q.kkt_solve(rec[b], L)
        assert np.max(np.abs(steps[b] - d_ref)) <= 1e-8 * np.max(np.abs(d_ref))


def test_cpp_sqp_port_equals_the_python_restatement(oracle):
    """Same statuses, iteration counts and iterates as oracle/sqp_reference.py::soft_sqp (two exact QP solvers: 1e-5 like the device
    loop's test), with trajectories split over threads."""
    from ungar_b200 import workloads as W

    N, iters = 10, 5
    xp = W.synthetic_batch(W.QUADRUPED, N, 4, seed=23)
    out, status = oracle.sqp_solve_port(N, xp, 1.0, 1.0, 1.0 / N, iters, threads=3)
    n = oracle.sizes(W.QUADRUPED, N)["n_dec"]
    for b in range(4):
        ref, ref_status, ref_iters, _ = S.soft_sqp(oracle, W.QUADRUPED, N, xp[b], 1.0, 1.0, 1.0 / N, iters)
        assert (status[b, 0], status[b, 1]) == (ref_status, ref_iters)
        assert np.max(np.abs(out[b, :n] - ref[:n])) <= 1e-5 * np.max(np.abs(ref[:n]))
        assert np.array_equal(out[b, n:], xp[b, n:])
    with pytest.raises(RuntimeError):
        oracle.sqp_solve_port(N, xp[:, :50], 1.0, 1.0, 1.0 / N, 1)


@pytest.mark.parametrize("name,N", [("quadruped", 10), ("quadrotor", 30), ("rc_car", 30)])
def test_sqp_golden_vectors_are_current(oracle, name, N):
    """tests/golden/sqp_*.npz (oracle/make_golden_sqp.py) still equal what the oracle computes; the fixtures also equal
    the C++ port.  The device loop is compared with the same files in tests/test_gpu_sqp.py."""
    import os

    from ungar_b200 import workloads as W

    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"sqp_{name}_N{N}.npz"))
    mid = W.MODEL_IDS[name]
    xp, iters = fx["xp"], int(fx["iterations"])
    assert np.array_equal(xp, W.synthetic_batch(mid, N, xp.shape[0], seed=int(fx["seed"])))
    for b in (0, xp.shape[0] - 1):
        x, st, it, log = S.soft_sqp(oracle, mid, N, xp[b], float(fx["stiffness"]), float(fx["epsilon"]), float(fx["multiplier"]), iters)
        assert (st, it) == tuple(fx["status"][b]) and np.allclose(x, fx["final"][b], rtol=0, atol=1e-12 * np.max(np.abs(x)))
        assert [l["ls"].alpha for l in log] == [a for a in fx["alphas"][b] if a >= 0]
    # the C++ port (Schur complement for the quadruped, Riccati for the other two) reaches the same statuses and iterates
    out, status = oracle.sqp_solve_port(N, xp, float(fx["stiffness"]), float(fx["epsilon"]), float(fx["multiplier"]), iters, threads=2,
                                        model=mid)
    n = oracle.sizes(mid, N)["n_dec"]
    assert np.array_equal(status, fx["status"])
    assert np.max(np.abs(out[:, :n] - fx["final"][:, :n])) <= 1e-5 * np.max(np.abs(fx["final"][:, :n]))


@pytest.mark.parametrize("model,N", [(0, 30), (0, 7), (1, 60), (1, 2)])
def test_cpp_riccati_port_equals_the_monolithic_kkt_solve(oracle, model, N):
    """Quadrotor / RC car: the Riccati port (input-rate coupling in the augmented state) against the QP assembled as
    AssembleOSQPInstance does and solved by one sparse LU."""
    from ungar_b200 import EXAMPLE_BARRIER
    from ungar_b200 import workloads as W

    k, eps = EXAMPLE_BARRIER[model]
    xp = W.synthetic_batch(model, N, 3, seed=43)
    steps = oracle.qp_solve_port(N, oracle.stage_sweep(model, N, xp, k, eps), threads=2, model=model)
    for b in range(3):
        P, q, A, g, _ = S.monolithic_qp(oracle, model, N, xp[b], k, eps)
        d_ref, _ = S.solve_qp(P, q, A, g, delta=0.0)
        assert np.max(np.abs(steps[b] - d_ref)) <= 1e-8 * np.max(np.abs(d_ref))
