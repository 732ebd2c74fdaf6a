"""Batched line search and soft-SQP loop against the CPU restatement of the reference's loop
(oracle/sqp_reference.py: backtracking_line_search.hpp:81-165, soft_sqp.hpp:63-109), through the C ABI."""
import numpy as np
import pytest

from ungar_b200 import EXAMPLE_BARRIER
from ungar_b200 import workloads as W

pytestmark = pytest.mark.gpu

# constraintViolationMultiplier of each example's optimizer (quadrotor.example.cpp:370 default 1.0; rc_car.example.cpp:363 and
# quadruped.example.cpp:444 pass dt = 1 / N)
def multiplier(model, N):
    return 1.0 if model == W.QUADROTOR else 1.0 / N


def rel(a, b):
    return abs(a - b) / max(1e-12, abs(b))


@pytest.mark.parametrize("model,N", [(W.QUADROTOR, 12), (W.RC_CAR, 20), (W.QUADRUPED, 10), (W.QUADRUPED, 30)])
def test_line_search_matches_reference_loop(oracle, model, N):
    """Directions: the exact QP step, an overshooting multiple of it (forces backtracking) and its negative (mostly rejected)."""
    import torch

    import ungar_b200
    from oracle import sqp_reference as S

    k, eps = EXAMPLE_BARRIER[model]
    mult = multiplier(model, N)
    m = ungar_b200.Model(W.MODEL_NAMES[model], N, dtype="f64", barrier=(k, eps))
    n = m.layout["n_dec"]
    B = 3
    xp = W.synthetic_batch(model, N, B, seed=17)
    base, grads = [], []
    for b in range(B):
        P, q, A, g, grad_f = S.monolithic_qp(oracle, model, N, xp[b], k, eps)
        base.append(S.solve_qp(P, q, A, g)[0])
        grads.append(grad_f)
    opts = m.sqp_options(constraint_violation_multiplier=mult)
    seen = set()
    for scale in (1.0, 6.0, -1.0, 1e-3):
        dw = np.stack(base) * scale
        d_xp = torch.from_numpy(xp.copy()).cuda()
        status = torch.zeros((B, 2), dtype=torch.int32, device="cuda")
        info = m.line_search(d_xp, torch.from_numpy(dw).cuda(), opts, status).cpu().numpy()
        got, st = d_xp.cpu().numpy(), status.cpu().numpy()
        for b in range(B):
            merit = S.Merit(oracle, model, N, xp[b], k, eps, mult)
            res, w_next = S.line_search(grads[b], dw[b], merit.phi, merit.theta, xp[b, :n].copy())
            seen.add((res.accepted, res.trials > 1))
            assert info[b, 0] == res.alpha, (scale, b, info[b], res)
            assert rel(info[b, 4], res.theta0) < 1e-9 and rel(info[b, 5], res.phi0) < 1e-9 and rel(info[b, 7], res.projection) < 1e-9
            if res.accepted:
                assert rel(info[b, 1], res.theta) < 1e-9 and rel(info[b, 2], res.phi) < 1e-9
                assert np.max(np.abs(got[b, :n] - w_next)) <= 1e-12 * np.max(np.abs(w_next))
                assert st[b, 0] in (0, 1) and st[b, 1] == 1
                f_next = merit.objective(w_next)
                f_prev = merit.objective(xp[b, :n].copy())
                assert st[b, 0] == (1 if (f_next - f_prev < 0 and abs(f_next - f_prev) < 1e-6) else 0)
            else:
                assert np.array_equal(got[b, :n], xp[b, :n]) and st[b, 0] == 2
            assert np.array_equal(got[b, n:], xp[b, n:])  # parameters are never touched
    assert (True, False) in seen and (True, True) in seen, seen  # both a first-trial accept and a backtracked one were exercised


def test_line_search_skips_stopped_trajectories_and_rejects_f32():
    import torch

    import ungar_b200
    from ungar_b200 import _lib

    N = 8
    m = ungar_b200.Model("rc_car", N, dtype="f64", barrier=EXAMPLE_BARRIER[W.RC_CAR])
    xp = torch.from_numpy(W.synthetic_batch(W.RC_CAR, N, 4, seed=2)).cuda()
    before = xp.clone()
    dw = torch.full((4, m.layout["n_dec"]), 1e-3, dtype=torch.float64, device="cuda")
    status = torch.tensor([[0, 0], [1, 3], [2, 5], [0, 1]], dtype=torch.int32, device="cuda")
    m.line_search(xp, dw, status=status)
    st = status.cpu().numpy()
    assert np.array_equal(st[1], [1, 3]) and np.array_equal(st[2], [2, 5]) and st[0, 1] == 1 and st[3, 1] == 2
    assert torch.equal(xp[1], before[1]) and torch.equal(xp[2], before[2])
    # an F32 handle searches in fp64 (twin handle, arguments widened on the device): same decisions as the F64 handle on the widened
    # arguments, iterate and report narrowed back
    m32 = ungar_b200.Model("rc_car", N, dtype="f32", barrier=EXAMPLE_BARRIER[W.RC_CAR])
    x32, d32 = before.float().clone(), dw.float()
    x64 = x32.double()
    info64 = m.line_search(x64, d32.double())
    info32 = m32.line_search(x32, d32)
    assert info32.dtype == torch.float32 and torch.equal(info32[:, 0].double(), info64[:, 0])            # same step sizes
    assert torch.allclose(info32.double(), info64, rtol=1e-6, atol=1e-30) and torch.equal(x32, x64.float())


@pytest.mark.parametrize("name,N,iters,perturb", [("quadruped", 10, 5, False), ("quadruped", 30, 4, True), ("quadrotor", 30, 6, False),
                                                  ("rc_car", 30, 6, False)])
def test_sqp_solve_matches_reference_loop(oracle, name, N, iters, perturb):
    """Whole loop on the device vs SoftSQPOptimizer::Optimize restated on the CPU (exact sparse-LU QP), same iterates."""
    import torch

    import ungar_b200
    from oracle import sqp_reference as S

    mid = W.MODEL_IDS[name]
    k, eps = EXAMPLE_BARRIER[mid]
    mult = multiplier(mid, N)
    m = ungar_b200.Model(name, N, dtype="f64", barrier=(k, eps))
    n = m.layout["n_dec"]
    B = 4
    xp = W.synthetic_batch(mid, N, B, seed=23, perturb_params=perturb)
    d_xp = torch.from_numpy(xp.copy()).cuda()
    status, info = m.sqp_solve(d_xp, m.sqp_options(max_iterations=iters, constraint_violation_multiplier=mult))
    got, st, info = d_xp.cpu().numpy(), status.cpu().numpy(), info.cpu().numpy()
    for b in range(B):
        ref, ref_status, ref_iters, log = S.soft_sqp(oracle, mid, N, xp[b], k, eps, mult, iters)
        assert (st[b, 0], st[b, 1]) == (ref_status, ref_iters), (b, st[b], ref_status, ref_iters)
        assert info[b, 0] == log[-1]["ls"].alpha
        # Two exact QP solvers (stage-wise Schur complement on the device, sparse LU of the KKT system here) agree to ~1e-7 per
        # step (tests/test_gpu_qp.py); the nonlinear loop compounds that: measured 1.0e-6 after four iterations at N = 30.
        scale = np.max(np.abs(ref[:n]))
        assert np.max(np.abs(got[b, :n] - ref[:n])) <= 1e-5 * scale, (b, np.max(np.abs(got[b, :n] - ref[:n])) / scale)
        assert np.array_equal(got[b, n:], xp[b, n:])
    # lock-step: feeding the DEVICE QP step of every iteration into the CPU loop removes the solver difference; what remains is
    # the line search and the bookkeeping, which must agree to rounding
    def device_step(b):
        def qp(x):
            rec = m.kkt_blocks(torch.from_numpy(x[None].copy()).cuda(), torch.zeros((1, m.layout["size"]), dtype=torch.float64, device="cuda"))
            return m.qp_solve(rec, want_multipliers=False)[0][0].cpu().numpy()
        return qp

    for b in range(2):
        ref, ref_status, ref_iters, _ = S.soft_sqp(oracle, mid, N, xp[b], k, eps, mult, iters, qp=device_step(b))
        assert (st[b, 0], st[b, 1]) == (ref_status, ref_iters)
        assert np.max(np.abs(got[b, :n] - ref[:n])) <= 1e-10 * np.max(np.abs(ref[:n]))
    # host-buffer entry point (MEM_HOST) and the SoftSQPOptimizer mirror give the same iterates
    host = xp.copy()
    opt = ungar_b200.SoftSQPOptimizer(False, mult, iters, k, eps)
    sol = opt.Optimize(m, host)
    assert np.array_equal(sol, got[:, :n]) and np.array_equal(opt.status, st)
    one = opt.Optimize(m, xp[0])
    assert np.array_equal(one, got[0, :n])


def test_sqp_solve_full_batch_properties():
    """BASELINE config 4 size (N = 100, 1024 trajectories): the bookkeeping is consistent with the acceptance rules and with
    the Function path evaluated at the final iterates."""
    import torch

    import ungar_b200

    N, B, iters = 100, 1024, 4  # quadruped.example.cpp:444: maxIterations 4, multiplier dt
    k, eps = EXAMPLE_BARRIER[W.QUADRUPED]
    m = ungar_b200.Model("quadruped", N, dtype="f64", barrier=(k, eps))
    n = m.layout["n_dec"]
    xp0 = W.synthetic_batch(W.QUADRUPED, N, B, seed=5)
    d_xp = torch.from_numpy(xp0.copy()).cuda()
    opts = m.sqp_options(max_iterations=iters, constraint_violation_multiplier=1.0 / N)
    status, info = m.sqp_solve(d_xp, opts)
    torch.cuda.synchronize()
    st, info = status.cpu().numpy(), info.cpu().numpy()
    assert torch.isfinite(d_xp).all() and np.isfinite(info).all()
    assert set(np.unique(st[:, 0])) <= {0, 1, 2} and st[:, 1].min() >= 1 and st[:, 1].max() <= iters
    assert np.all(st[st[:, 0] == 0, 1] == iters)  # still running <=> stopped by max_iterations
    alpha, th, ph, th0, ph0 = info[:, 0], info[:, 1], info[:, 2], info[:, 4], info[:, 5]
    acc = alpha > 0
    assert np.array_equal(acc, st[:, 0] != 2)
    k2 = np.round(np.log2(alpha[acc]))
    assert np.all(alpha[acc] == 2.0 ** k2) and alpha[acc].min() >= opts.alpha_min
    ok = np.where(th > opts.theta_max, th < (1 - opts.gamma_theta) * th0,
                  (ph < (1 - opts.gamma_phi) * ph0) | (th < (1 - opts.gamma_theta) * th0) | (np.maximum(th, th0) < opts.theta_min))
    assert np.all(ok[acc])
    # the merit values the line search reports are those of the Function path at the final iterate
    sample = [0, 1, B // 2, B - 1]
    xs = d_xp[sample].cpu().numpy()
    g = m.equalityConstraints(xs)
    f = m.objective(xs)[:, 0]
    assert np.allclose(np.sqrt((g * g).sum(1)) / N, th[sample], rtol=1e-9)
    assert np.allclose(f, info[sample, 3], rtol=1e-9)
    # determinism
    d2 = torch.from_numpy(xp0.copy()).cuda()
    s2, _ = m.sqp_solve(d2, opts)
    assert torch.equal(d2, d_xp) and torch.equal(s2, status)


@pytest.mark.parametrize("name,N", [("quadrotor", 30), ("rc_car", 60)])
def test_sqp_solve_on_an_f32_handle_runs_the_fp64_loop(name, N):
    """BASELINE configs 2 and 3 (fp32): ungar_b200_sqp_solve on an F32 handle widens the iterate on the device, runs the F64 loop on
    a twin handle and narrows the result — statuses, iteration counts and step sizes equal the F64 handle's on the widened input,
    the iterate equals its fp32 rounding; device and host buffers."""
    import torch

    import ungar_b200
    from ungar_b200 import EXAMPLE_BARRIER
    from ungar_b200 import workloads as W

    mid = W.MODEL_IDS[name]
    m32 = ungar_b200.Model(name, N, dtype="f32", barrier=EXAMPLE_BARRIER[mid])
    m64 = ungar_b200.Model(name, N, dtype="f64", barrier=EXAMPLE_BARRIER[mid])
    opts = m64.sqp_options(max_iterations=4, constraint_violation_multiplier=1.0 if name == "quadrotor" else 1.0 / N)
    xp32 = torch.from_numpy(W.synthetic_batch(mid, N, 16, seed=9)).float().cuda()
    x64 = xp32.double()
    st64, info64 = m64.sqp_solve(x64, opts)
    x32 = xp32.clone()
    st32, info32 = m32.sqp_solve(x32, opts)
    assert torch.equal(st32, st64) and info32.dtype == torch.float32
    assert torch.equal(info32[:, 0].double(), info64[:, 0]) and torch.equal(x32, x64.float())
    assert not torch.equal(x32, xp32)                                                   # the iterate moved
    host = xp32.cpu().numpy().copy()
    sth, _ = m32.sqp_solve(host, opts)                                                  # host buffers
    assert np.array_equal(sth, st64.cpu().numpy()) and np.array_equal(host[:, :m32.layout["n_dec"]], x32.cpu().numpy()[:, :m32.layout["n_dec"]])


@pytest.mark.parametrize("name,N", [("quadruped", 10), ("quadrotor", 30), ("rc_car", 30)])
def test_sqp_solve_matches_the_committed_golden_vectors(name, N):
    """The same comparison as test_sqp_solve_matches_reference_loop, against tests/golden/sqp_*.npz (written by
    oracle/make_golden_sqp.py from the restated SoftSQPOptimizer::Optimize; kept current by tests/test_oracle_sqp.py)."""
    import os

    import torch

    import ungar_b200

    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"sqp_{name}_N{N}.npz"))
    m = ungar_b200.Model(name, N, dtype="f64", barrier=(float(fx["stiffness"]), float(fx["epsilon"])))
    n = m.layout["n_dec"]
    d_xp = torch.from_numpy(fx["xp"].copy()).cuda()
    status, info = m.sqp_solve(d_xp, m.sqp_options(max_iterations=int(fx["iterations"]), constraint_violation_multiplier=float(fx["multiplier"])))
    got, st, info = d_xp.cpu().numpy(), status.cpu().numpy(), info.cpu().numpy()
    assert np.array_equal(st, fx["status"])
    for b in range(got.shape[0]):
        last = [a for a in fx["alphas"][b] if a >= 0][-1]
        assert info[b, 0] == last
        scale = np.max(np.abs(fx["final"][b, :n]))
        assert np.max(np.abs(got[b, :n] - fx["final"][b, :n])) <= 1e-5 * scale
        assert np.array_equal(got[b, n:], fx["xp"][b, n:])
