"""Pins the CPU oracle (oracle/) before anything is compared against it.

Anchors:
  * known answers derived from the reference examples' own constants
    (quadrotor hover, quadruped stance, RC-car first step — rc_car.example.cpp:164-179,320-343);
  * test/autodiff/function.test.cpp:33-59 — ApproximateExponentialMap ~ ExponentialMap on [-1,1]^3,
    value (0,0,0,1) and Jacobian [I/2; 0] at 0;
  * the reference's own self-check method: AD vs 2nd-order central finite differences
    (include/ungar/autodiff/function.hpp:285-325) — here at a far stricter tolerance than
    Utils::CompareMatrices (utils/utils.hpp:1070-1090: rel 1e-2 AND abs 1e-3);
  * include/ungar/optimization/soft_inequality_constraint.hpp:133-190 — barrier pieces;
  * monolithic assembly (soft_sqp.hpp:141-158) == stage-wise port, and the block cut is lossless.
"""
import numpy as np
import pytest

from ungar_b200 import workloads as W

OBJ, EQ, INEQ = 0, 1, 2
MODELS = [W.QUADROTOR, W.RC_CAR, W.QUADRUPED]


def dense(rows, cols, vals, shape):
    m = np.zeros(shape)
    m[rows, cols] = vals
    return m


@pytest.mark.parametrize("model,N", [(W.QUADROTOR, 30), (W.RC_CAR, 30), (W.QUADRUPED, 30), (W.RC_CAR, 60),
                                     (W.QUADRUPED, 100)])
def test_sizes_match_reference_examples(oracle, model, N):
    s = oracle.sizes(model, N)
    w = W.sizes(model, N)
    for k in ("nx", "nu", "n_dec", "n_par", "m_eq", "m_ineq"):
        assert s[k] == w[k]
    # SURVEY.md §8 "Model dimensions" (N = 30 rows verified against the compiled reference headers).
    expect = {(W.QUADROTOR, 30): (523, 437, 403, 240), (W.RC_CAR, 30): (246, 83, 186, 90),
              (W.QUADRUPED, 30): (1123, 948, 883, 360), (W.RC_CAR, 60): (486, 143, 366, 180),
              (W.QUADRUPED, 100): (3713, 2978, 2913, 1200)}[(model, N)]
    assert (s["n_dec"], s["n_par"], s["m_eq"], s["m_ineq"]) == expect


def test_quadrotor_hover_is_an_equilibrium(oracle):
    xp = W.quadrotor_nominal(30, 0.0)
    g = oracle.evaluate(W.QUADROTOR, EQ, 30, xp)
    assert g.shape == (403,)
    assert np.abs(g).max() == 0.0  # u = sqrt(m g / (4 b)) (quadrotor.example.cpp:356-358)
    h = oracle.evaluate(W.QUADROTOR, INEQ, 30, xp)
    u = np.sqrt(1.5 * 9.80665 / 0.015 / 4.0)
    assert np.allclose(h[0::2], u - 100.0) and np.allclose(h[1::2], -u)


def test_rc_car_first_step_known_answer(oracle):
    xp = W.rc_car_nominal(30, 0.0)
    xn = oracle.dynamics(W.RC_CAR, 30, xp, 0)
    vx = 1.0 - (1.0 / 30.0) * (0.0518 + 0.00035) / 0.041
    assert np.allclose(xn, [vx / 30.0, 0.0, 0.0, vx, 0.0, 0.0], rtol=0, atol=1e-15)
    assert abs(xn[3] - 0.957601626) < 1e-9 and abs(xn[0] - 0.03192005) < 1e-8


def test_quadruped_stance_known_answers(oracle):
    for N in (30, 100):
        xp = W.quadruped_nominal(N, 0.0)
        g = oracle.evaluate(W.QUADRUPED, EQ, N, xp)
        assert np.abs(g[:13 + 13 * N]).max() < 1e-15  # f_i = m g / 4 e_z (quadruped.example.cpp:416-419)
        foot = g[13 + 13 * N:].reshape(N, 4, 4)
        assert np.allclose(foot[0], [[0.0, 0.0, 0.0, 0.38]] * 4)  # foot rows for k = 0
        assert np.abs(foot[1:]).max() < 1e-15


def test_approximate_exponential_map(oracle):
    q, J = oracle.approx_exp(np.zeros(3))
    assert np.allclose(q, [0, 0, 0, 1], atol=1e-8)
    assert np.allclose(J, np.vstack([0.5 * np.eye(3), np.zeros((1, 3))]), atol=1e-7)
    rng = np.random.default_rng(0)
    for _ in range(1024):  # function.test.cpp:52-57
        v = rng.uniform(-1, 1, 3)
        th = np.linalg.norm(v)
        exact = np.concatenate([v / th * np.sin(th / 2), [np.cos(th / 2)]])
        q, J = oracle.approx_exp(v)
        assert np.allclose(q, exact, rtol=0, atol=1e-12)
        fd = np.zeros((4, 3))
        for j in range(3):
            e = np.zeros(3)
            e[j] = 1e-6
            fd[:, j] = (oracle.approx_exp(v + e)[0] - oracle.approx_exp(v - e)[0]) / 2e-6
        assert np.allclose(J, fd, rtol=0, atol=1e-8)


def test_poly_barrier_pieces(oracle):
    for k, eps in [(100.0, 2e-5), (100.0, 1e-2), (1.0, 1.0)]:  # per-example settings, SURVEY.md A.4
        a1, b1 = k, -0.5 * k * eps
        c1 = -1.0 / 3.0 * (-b1 - a1 * eps) * eps - 0.5 * a1 * eps ** 2 - b1 * eps
        a2 = (-b1 - a1 * eps) / eps ** 2
        z = np.array([3.0 * eps, 0.5 * eps, -0.5 * eps, -2.0 * eps])  # Zsoft(z) = sum b(-z)
        x = -z
        val, dz, d2z = oracle.barrier(k, eps, z)
        b = [0.5 * a1 * x[0] ** 2 + b1 * x[0] + c1, 0.5 * a1 * x[1] ** 2 + b1 * x[1] + c1,
             a2 * x[2] ** 3 / 3 + 0.5 * a1 * x[2] ** 2 + b1 * x[2] + c1, 0.0]
        assert np.isclose(val, sum(b), rtol=1e-13)
        assert np.allclose(dz, [-(a1 * x[0] + b1), -(a1 * x[1] + b1), -(a2 * x[2] ** 2 + a1 * x[2] + b1), 0.0], rtol=1e-12)
        assert np.allclose(d2z, [a1, a1, 2 * a2 * x[2] + a1, 0.0], rtol=1e-12)
        # C2 at both knots
        for knot in (0.0, -eps):
            lo = oracle.barrier(k, eps, np.array([knot - 1e-9 * max(eps, 1e-3)]))
            hi = oracle.barrier(k, eps, np.array([knot + 1e-9 * max(eps, 1e-3)]))
            assert abs(lo[0] - hi[0]) < 1e-6 * k * eps ** 2 + 1e-15
            assert abs(lo[1][0] - hi[1][0]) < 1e-6 * k * max(eps, 1e-3)


@pytest.mark.parametrize("model", MODELS)
def test_ad_jacobians_match_central_differences(oracle, model):
    N = 4
    s = W.sizes(model, N)
    xp = W.synthetic_batch(model, N, 3, seed=11, perturb_params=True)[2]
    for fn, ny in ((OBJ, 1), (EQ, s["m_eq"]), (INEQ, s["m_ineq"])):
        r, c, v = oracle.jacobian(model, fn, N, xp)
        J = dense(r, c, v, (ny, s["n_dec"]))
        fd = np.zeros_like(J)
        for j in range(s["n_dec"]):
            e = np.zeros_like(xp)
            e[j] = 1e-6 * max(1.0, abs(xp[j]))
            fd[:, j] = (oracle.evaluate(model, fn, N, xp + e) - oracle.evaluate(model, fn, N, xp - e)) / (2 * e[j])
        scale = np.maximum(np.abs(fd), 1.0)
        assert np.max(np.abs(J - fd) / scale) < 2e-7, (model, fn)
        # structural pattern must cover every numerically nonzero FD entry
        assert np.all((np.abs(fd) > 1e-6) <= (J != 0) | (np.abs(fd) < 1e-6))


@pytest.mark.parametrize("model", MODELS)
def test_ad_hessian_matches_differences_of_gradients(oracle, model):
    N = 3
    s = W.sizes(model, N)
    xp = W.synthetic_batch(model, N, 2, seed=5)[1]
    r, c, v = oracle.hessian(model, N, xp)
    assert np.all(c >= r)  # upper triangle only (function.hpp:563-571)
    H = dense(r, c, v, (s["n_dec"], s["n_dec"]))
    H = H + np.triu(H, 1).T
    fd = np.zeros_like(H)

    def grad(x):
        rr, cc, vv = oracle.jacobian(model, OBJ, N, x)
        return dense(rr, cc, vv, (1, s["n_dec"]))[0]

    for j in range(s["n_dec"]):
        e = np.zeros_like(xp)
        e[j] = 1e-5
        fd[:, j] = (grad(xp + e) - grad(xp - e)) / 2e-5
    assert np.max(np.abs(H - fd)) < 1e-6


@pytest.mark.parametrize("model,N,bar", [(W.QUADROTOR, 30, (100.0, 2e-5)), (W.RC_CAR, 60, (100.0, 1e-2)),
                                         (W.QUADRUPED, 30, (1.0, 1.0)), (W.QUADRUPED, 100, (1.0, 1.0))])
def test_stage_port_equals_monolithic_assembly(oracle, model, N, bar):
    """soft_sqp.hpp:141-158 on the monolithic matrices == node-by-node port; the cut is lossless."""
    xps = W.synthetic_batch(model, N, 3, seed=3, perturb_params=True)
    recs = oracle.stage_sweep(model, N, xps, bar[0], bar[1], threads=2)
    for b in (0, 2):
        mono = oracle.kkt_record(model, N, xps[b], bar[0], bar[1])  # raises if a nonzero is left out
        assert np.max(np.abs(mono - recs[b]) / np.maximum(np.abs(mono), 1.0)) < 1e-13


def test_kkt_record_reassembles_to_the_monolithic_matrices(oracle):
    """Scatter the record back into P, q, A and compare with the sparse monolithic pieces."""
    model, N, (k, eps) = W.QUADRUPED, 6, (1.0, 1.0)
    s, L = oracle.sizes(model, N), oracle.record_layout(model, N)
    xp = W.synthetic_batch(model, N, 1, seed=9)[0]
    rec = oracle.kkt_record(model, N, xp, k, eps)
    nx, nu, nz, n = s["nx"], s["nu"], L["nz"], s["n_dec"]
    # reference-side pieces
    rh, ch, vh = oracle.hessian(model, N, xp)
    Hf = dense(rh, ch, vh, (n, n))
    r, c, v = oracle.jacobian(model, INEQ, N, xp)
    Jh = dense(r, c, v, (s["m_ineq"], n))
    hval = oracle.evaluate(model, INEQ, N, xp)
    _, dB, d2B = oracle.barrier(k, eps, hval)
    P = Hf + np.triu(Jh.T @ np.diag(d2B) @ Jh) + 1e-6 * np.eye(n)
    r, c, v = oracle.jacobian(model, OBJ, N, xp)
    q = dense(r, c, v, (1, n))[0] + Jh.T @ dB
    assert np.allclose(rec[L["grad"]:L["grad"] + n], q, rtol=1e-13, atol=1e-15)
    P2 = np.zeros((n, n))
    iu = np.triu_indices(nz)
    for kk in range(N):
        idx = np.concatenate([np.arange(nx * kk, nx * kk + nx), nx * (N + 1) + nu * kk + np.arange(nu)])
        blk = np.zeros((nz, nz))
        blk[iu] = rec[L["H"] + kk * L["tri"]:L["H"] + (kk + 1) * L["tri"]]
        P2[np.ix_(idx, idx)] += blk
    blk = np.zeros((nx, nx))
    blk[np.triu_indices(nx)] = rec[L["HN"]:L["HN"] + L["ntri_N"]]
    idx = np.arange(nx * N, nx * N + nx)
    P2[np.ix_(idx, idx)] += blk
    assert np.allclose(P2, P, rtol=1e-13, atol=1e-18)
