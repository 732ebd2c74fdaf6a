"""Synthetic NMPC inputs in Ungar's flat ``variables = [X | U | parameters]`` layout.

Host-side (numpy) data only: the nominal points restate the initialisation code of the three reference
examples and the per-trajectory perturbation follows SURVEY.md §8(d) "Synthetic inputs".

Reference anchors
  quadrotor : example/mpc/quadrotor.example.cpp:323-358 (constants, hover guess), :379-395 (references)
  rc_car    : example/mpc/rc_car.example.cpp:317-352, :371-376
  quadruped : example/mpc/quadruped.example.cpp:378-430, :463-493
The reference fixes N = 30 and dt = 1/N; N is a free argument here (BASELINE.json uses 30/60/100).
"""
from __future__ import annotations

import math

import numpy as np

QUADROTOR, RC_CAR, QUADRUPED = 0, 1, 2
MODEL_NAMES = {QUADROTOR: "quadrotor", RC_CAR: "rc_car", QUADRUPED: "quadruped"}
MODEL_IDS = {v: k for k, v in MODEL_NAMES.items()}


def sizes(model: int, N: int) -> dict:
    """nx, nu, n_dec, n_par, m_eq, m_ineq of a model (SURVEY.md §8 'Model dimensions')."""
    if model == QUADROTOR:
        nx, nu = 13, 4
        n_par = 21 + 13 * (N + 1) + 13
        m_eq, m_ineq = nx * (N + 1), 2 * nu * N
    elif model == RC_CAR:
        nx, nu = 6, 2
        n_par = 15 + 2 * (N + 1) + 6
        m_eq, m_ineq = nx * (N + 1), 3 * N
    elif model == QUADRUPED:
        nx, nu = 13, 24
        n_par = 29 * (N + 1) + 49
        m_eq, m_ineq = nx * (N + 1) + 16 * N, 12 * N
    else:
        raise ValueError(f"unknown model {model}")
    n_dec = nx * (N + 1) + nu * N
    return dict(nx=nx, nu=nu, N=N, n_dec=n_dec, n_par=n_par, n_xp=n_dec + n_par, m_eq=m_eq, m_ineq=m_ineq)


def _zquat(angle: float) -> np.ndarray:
    """Utils::ElementaryZQuaternion in (x, y, z, w) storage (utils/utils.hpp:920-925)."""
    return np.array([0.0, 0.0, math.sin(angle / 2.0), math.cos(angle / 2.0)])


def quadrotor_nominal(N: int = 30, time: float = 0.0) -> np.ndarray:
    s = sizes(QUADROTOR, N)
    xp = np.zeros(s["n_xp"])
    P = s["n_dec"]
    dt, m, g, b = 1.0 / N, 1.5, 9.80665, 0.015
    xp[P + 0], xp[P + 1] = dt, m
    xp[P + 2:P + 5] = 3e-2
    xp[P + 5:P + 17] = [0.2, 0.2, 0.0, -0.2, 0.2, 0.0, -0.2, -0.2, 0.0, 0.2, -0.2, 0.0]
    xp[P + 17], xp[P + 18], xp[P + 19], xp[P + 20] = g, b, 0.1, 1e2
    xm = np.zeros(13)
    xm[2], xm[6] = 4.0, 1.0
    xp[P + 21 + 13 * (N + 1):P + 21 + 13 * (N + 1) + 13] = xm
    for k in range(N + 1):
        xp[13 * k:13 * k + 13] = xm
    xp[13 * (N + 1):P] = math.sqrt(m * g / b / 4.0)
    z_period, z_amp, yaw_rate, t_start = 4.0, 1.0, math.pi, 4.0
    for k in range(N + 1):
        t = time + k * dt
        on = float(t > t_start)
        xp[P + 21 + 3 * k + 2] = 4.0 + on * z_amp * math.sin(2.0 * math.pi / z_period * t)
        xp[P + 21 + 3 * (N + 1) + 4 * k:P + 21 + 3 * (N + 1) + 4 * k + 4] = (
            _zquat(yaw_rate * t) if on else np.array([0.0, 0.0, 0.0, 1.0]))
        xp[P + 21 + 7 * (N + 1) + 3 * k + 2] = (on * 2.0 * math.pi / z_period * z_amp *
                                                 math.cos(2.0 * math.pi / z_period * t))
        xp[P + 21 + 10 * (N + 1) + 3 * k + 2] = on * yaw_rate
    return xp


def rc_car_nominal(N: int = 30, time: float = 0.0) -> np.ndarray:
    s = sizes(RC_CAR, N)
    xp = np.zeros(s["n_xp"])
    P = s["n_dec"]
    dt = 1.0 / N
    xp[P:P + 15] = [dt, 0.041, 27.8e-6, 0.029, 0.033, 2.579, 1.2, 0.192, 3.3852, 1.2691, 0.1737,
                    0.287, 0.0545, 0.0518, 0.00035]
    xm = np.array([0.0, 0.0, 0.0, 1.0, 0.0, 0.0])
    xp[P + 15 + 2 * (N + 1):P + 15 + 2 * (N + 1) + 6] = xm
    for k in range(N + 1):
        xp[6 * k:6 * k + 6] = xm
        xp[6 * k] = xm[3] * (k * dt)
        t = time + k * dt
        xp[P + 15 + 2 * k] = 1.0 * t
        xp[P + 15 + 2 * k + 1] = 0.2 * math.sin(2.0 * math.pi / 8.0 * t)
    return xp


def quadruped_nominal(N: int = 30, time: float = 0.0) -> np.ndarray:
    s = sizes(QUADRUPED, N)
    xp = np.zeros(s["n_xp"])
    U0, P0 = 13 * (N + 1), s["n_dec"]
    Rho = P0 + 29 * (N + 1)
    dt, m, g, L = 1.0 / N, 25.0, 9.80665, 0.42
    hips = np.array([[0.2, 0.15, -0.1], [0.2, -0.15, -0.1], [-0.2, 0.15, -0.1], [-0.2, -0.15, -0.1]])
    feet = np.array([[0.2, 0.1, 0.0], [0.2, -0.1, 0.0], [-0.2, 0.1, 0.0], [-0.2, -0.1, 0.0]])
    xp[Rho + 0], xp[Rho + 1] = dt, m
    xp[Rho + 2:Rho + 5] = [0.048125, 0.093125, 0.055625]
    xp[Rho + 5:Rho + 17] = hips.ravel()
    xp[Rho + 17], xp[Rho + 18], xp[Rho + 19] = L, g, 0.7
    xm = np.zeros(13)
    xm[2], xm[6] = 0.38, 1.0
    xp[Rho + 20:Rho + 33] = xm
    for i in range(4):
        xp[Rho + 33 + 4 * i] = 1.0
        xp[Rho + 34 + 4 * i:Rho + 37 + 4 * i] = feet[i]
    for k in range(N + 1):
        xp[13 * k:13 * k + 13] = xm
    for k in range(N):
        for i in range(4):
            xp[U0 + 24 * k + 6 * i + 2] = m * g / 4.0
            xp[U0 + 24 * k + 6 * i + 3:U0 + 24 * k + 6 * i + 6] = feet[i]  # identity^-1 * foot
    z_period, z_amp, yaw_rate, t_start, gait = 8.0, 0.03, math.pi / 6.0, 2.0, 0.4
    for k in range(N + 1):
        t = time + k * dt
        on = float(t > t_start)
        pk = P0 + 29 * k
        xp[pk + 2] = 0.38 + on * z_amp * math.sin(2.0 * math.pi / z_period * (t - t_start))
        xp[pk + 3:pk + 7] = _zquat(yaw_rate * (t - t_start)) if on else np.array([0.0, 0.0, 0.0, 1.0])
        xp[pk + 9] = on * 2.0 * math.pi / z_period * z_amp * math.cos(2.0 * math.pi / z_period * (t - t_start))
        xp[pk + 12] = on * yaw_rate
        phase = math.sin(2.0 * math.pi / gait * t) > 0.0
        for i in range(4):
            xp[pk + 13 + 4 * i] = (float(phase if (i & 1) else (not phase)) if on else 1.0)
            xp[pk + 14 + 4 * i:pk + 17 + 4 * i] = hips[i] - 0.8 * L * np.array([0.0, 0.0, 1.0])
    return xp


_NOMINAL = {QUADROTOR: quadrotor_nominal, RC_CAR: rc_car_nominal, QUADRUPED: quadruped_nominal}
# quadruped references at time = 2.5 so that the pace gait pattern is active.
_DEFAULT_TIME = {QUADROTOR: 4.5, RC_CAR: 0.0, QUADRUPED: 2.5}


def nominal(model: int, N: int, time: float | None = None) -> np.ndarray:
    return _NOMINAL[model](N, _DEFAULT_TIME[model] if time is None else time)


def perturb_parameters(model: int, N: int, xp: np.ndarray, rng) -> None:
    """In-place, per-trajectory perturbation of the physical parameters and references (test coverage: anisotropic
    inertia, asymmetric geometry, arbitrary contact schedules — nothing the nominal examples exercise)."""
    s = sizes(model, N)
    B, P = xp.shape[0], s["n_dec"]
    scale = lambda n: rng.uniform(0.8, 1.25, (B, n))  # noqa: E731
    if model == QUADROTOR:
        xp[:, P + 1:P + 5] *= scale(4)              # mass, inertia (each axis separately)
        xp[:, P + 5:P + 17] *= scale(12)            # propeller positions
        xp[:, P + 18:P + 20] *= scale(2)            # thrust / drag constants
        xp[:, P + 21:P + 21 + 3 * (N + 1)] += rng.uniform(-0.2, 0.2, (B, 3 * (N + 1)))
        xp[:, P + 21 + 7 * (N + 1):P + 21 + 13 * (N + 1)] += 0.1 * rng.standard_normal((B, 6 * (N + 1)))
    elif model == RC_CAR:
        xp[:, P + 1:P + 15] *= scale(14)
        xp[:, P + 15:P + 15 + 2 * (N + 1)] += rng.uniform(-0.2, 0.2, (B, 2 * (N + 1)))
    else:
        Rho = P + 29 * (N + 1)
        xp[:, Rho + 1:Rho + 5] *= scale(4)          # mass, inertia
        xp[:, Rho + 5:Rho + 18] *= scale(13)        # hips, leg length
        xp[:, Rho + 19] *= scale(1)[:, 0]           # friction coefficient
        par = xp[:, P:Rho].reshape(B, N + 1, 29)
        par[:, :, 0:3] += rng.uniform(-0.05, 0.05, (B, N + 1, 3))
        par[:, :, 7:13] += 0.1 * rng.standard_normal((B, N + 1, 6))
        par[:, :, 13::4] = (rng.uniform(0, 1, (B, N + 1, 4)) > 0.4).astype(np.float64)   # contact schedule
        par[:, :, 14:17] += rng.uniform(-0.03, 0.03, (B, N + 1, 3))
        xp[:, Rho + 33:Rho + 49:4] = (rng.uniform(0, 1, (B, 4)) > 0.3).astype(np.float64)  # measured contacts
        xp[:, Rho + 20:Rho + 23] += rng.uniform(-0.05, 0.05, (B, 3))


def synthetic_batch(model: int, N: int, batch: int, seed: int = 20240807, time: float | None = None,
                    perturb_params: bool = False) -> np.ndarray:
    """``[batch, n_xp]`` float64 trajectories: nominal point + seeded perturbation of X and U only.

    positions += U(-0.1, 0.1); quaternions = normalise(q + 0.1 N(0,1)^4) with the sign making q.q_ref > 0;
    velocities += N(0, 0.1); inputs *= 1 + 0.05 U(-1, 1) (rc_car inputs += 0.5 U(-1, 1) since the nominal
    inputs are zero); rc_car v_x >= 0.5.  Parameters stay at their nominal values.
    """
    s = sizes(model, N)
    rng = np.random.Generator(np.random.Philox(seed + 7919 * model))
    base = nominal(model, N, time)
    xp = np.tile(base, (batch, 1))
    nx, nu = s["nx"], s["nu"]
    X = xp[:, :nx * (N + 1)].reshape(batch, N + 1, nx)
    U = xp[:, nx * (N + 1):s["n_dec"]].reshape(batch, N, nu)
    if model in (QUADROTOR, QUADRUPED):
        X[:, :, 0:3] += rng.uniform(-0.1, 0.1, (batch, N + 1, 3))
        q = X[:, :, 3:7] + 0.1 * rng.standard_normal((batch, N + 1, 4))
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
        if model == QUADROTOR:
            P = s["n_dec"]
            qref = base[P + 21 + 3 * (N + 1):P + 21 + 7 * (N + 1)].reshape(N + 1, 4)
        else:
            qref = base[s["n_dec"]:s["n_dec"] + 29 * (N + 1)].reshape(N + 1, 29)[:, 3:7]
        sign = np.where(np.sum(q * qref[None], axis=-1, keepdims=True) < 0.0, -1.0, 1.0)
        X[:, :, 3:7] = q * sign
        X[:, :, 7:13] += 0.1 * rng.standard_normal((batch, N + 1, 6))
        U *= 1.0 + 0.05 * rng.uniform(-1.0, 1.0, U.shape)
    else:
        X[:, :, 0:2] += rng.uniform(-0.1, 0.1, (batch, N + 1, 2))
        X[:, :, 2] += 0.1 * rng.standard_normal((batch, N + 1))
        X[:, :, 3:6] += 0.1 * rng.standard_normal((batch, N + 1, 3))
        X[:, :, 3] = np.maximum(X[:, :, 3], 0.5)
        U += 0.5 * rng.uniform(-1.0, 1.0, U.shape)
    if perturb_params:
        perturb_parameters(model, N, xp, rng)
    return np.ascontiguousarray(xp)
