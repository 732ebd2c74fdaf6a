"""Rigid-body dynamics evaluator behind the generic path.

In the reference, ``Ungar::Robot<Scalar>`` wraps Pinocchio v2.7.0 (include/ungar/rbd/robot.hpp:40-104, rbd/evaluator.hpp:45-58):
a URDF is loaded behind a free-flyer root joint and the quantities of rbd/quantities/*.hpp are evaluated by Pinocchio's algorithms;
instantiated with ``ad_scalar_t`` the articulated-body algorithm is TAPED and becomes one more ``Autodiff::Function`` of
``[q; v; tau]`` whose values and Jacobian are evaluated by generated code (test/rbd/robot.test.cpp:109-162).  Pinocchio is absent
from the reference tree and from this image, so this module is a from-scratch statement of the same pipeline on top of the
generic path of this repository:

* ``load_urdf``   URDF -> kinematic tree behind a free-flyer root (links attached by fixed joints are merged into their parent body,
                  as Pinocchio's URDF parser does);
* ``aba`` / ``rnea`` / ``crba``  Featherstone's articulated-body, recursive Newton-Euler and composite-rigid-body algorithms in
                  body coordinates, written over a GENERIC scalar: with ``ungar_b200.autodiff.AD`` scalars they record a tape, and
                  ``Robot.MakeFunction`` hands that tape to the register-machine kernels — batched forward dynamics and their
                  Jacobian on the GPU (``tests/test_rbd.py`` shows the shape of test/rbd/robot.test.cpp);
* conventions     Pinocchio's: ``q = [p(3), quaternion (x, y, z, w), joint angles]``, ``v = [linear(3), angular(3)]`` of the base
                  in the BASE frame followed by joint rates, generalized forces ordered like ``v``, gravity 9.81 along -z by default.

Status: the algorithms are validated on the CPU against an independent oracle (``oracle/rbd_reference.py``: CRBA + RNEA with 6x6
spatial matrices) and against identities (RNEA o ABA = id, M symmetric positive definite, energy balance).  Parity with Pinocchio
itself is UNPINNED (no copy of it exists here).  Evaluating the functions needs a GPU (there is no CPU evaluation path for tapes);
plain-float calls of ``aba`` / ``rnea`` exist so that tests can check the algorithm that gets taped.
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET

import numpy as np

from . import autodiff as A

GRAVITY = 9.81  # pinocchio::Model::gravity981


# ---------------------------------------------------------------------------------------------------------------------------
# Small fixed-size algebra over a generic scalar (float or autodiff.AD): 3-vectors and 3x3 matrices as Python lists
# ---------------------------------------------------------------------------------------------------------------------------
def _sin(x):
    return A.sin(x) if isinstance(x, A.AD) else math.sin(x)


def _cos(x):
    return A.cos(x) if isinstance(x, A.AD) else math.cos(x)


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _mv(M, v):
    return [M[i][0] * v[0] + M[i][1] * v[1] + M[i][2] * v[2] for i in range(3)]


def _mtv(M, v):
    return [M[0][i] * v[0] + M[1][i] * v[1] + M[2][i] * v[2] for i in range(3)]


def _mm(Am, Bm):
    return [[Am[i][0] * Bm[0][j] + Am[i][1] * Bm[1][j] + Am[i][2] * Bm[2][j] for j in range(3)] for i in range(3)]


def _tr(M):
    return [[M[j][i] for j in range(3)] for i in range(3)]


def _add(a, b):
    return [x + y for x, y in zip(a, b)]


def _sub(a, b):
    return [x - y for x, y in zip(a, b)]


def _scale(s, a):
    return [s * x for x in a]


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return [[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr]]


def _rodrigues(axis, q):
    """Rotation about the unit vector ``axis`` by the (generic-scalar) angle q."""
    s, c = _sin(q), _cos(q)
    x, y, z = axis
    t = 1.0 - c
    return [[t * x * x + c, t * x * y - s * z, t * x * z + s * y], [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
            [t * x * z - s * y, t * y * z + s * x, t * z * z + c]]


def _quat_matrix(x, y, z, w):
    return [[1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - w * z), 2.0 * (x * z + w * y)],
            [2.0 * (x * y + w * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - w * x)],
            [2.0 * (x * z - w * y), 2.0 * (y * z + w * x), 1.0 - 2.0 * (x * x + y * y)]]


# Spatial vectors are pairs (angular, linear) of 3-lists in BODY coordinates.  A parent-to-child transform is (E, r): E rotates
# parent coordinates into child coordinates, r is the child origin in parent coordinates.
def _xm(E, r, m):  # motion vector parent -> child
    w, v = m
    return _mv(E, w), _mv(E, _sub(v, _cross(r, w)))


def _xtf(E, r, f):  # force vector child -> parent  (X^T f)
    n, l = f
    lp = _mtv(E, l)
    return _add(_mtv(E, n), _cross(r, lp)), lp


def _crm(a, b):  # a x b for motion vectors
    return _cross(a[0], b[0]), _add(_cross(a[0], b[1]), _cross(a[1], b[0]))


def _crf(a, f):  # a x* f for force vectors
    return _add(_cross(a[0], f[0]), _cross(a[1], f[1])), _cross(a[0], f[1])


class Body:
    """One movable body: its joint (placement in the parent body frame, type, axis) and its spatial inertia about the body origin."""

    def __init__(self, name, parent, R, p, jtype, axis):
        self.name, self.parent, self.R, self.p, self.jtype, self.axis = name, parent, R, p, jtype, axis
        self.mass, self.h, self.I = 0.0, [0.0, 0.0, 0.0], [[0.0] * 3 for _ in range(3)]  # m, m c, inertia about the origin

    def add_inertia(self, mass, com, Ic):
        """Adds a rigid part: mass, centre of mass and inertia about that centre, all in this body's frame."""
        self.mass += mass
        self.h = _add(self.h, _scale(mass, com))
        c = com
        shift = [[mass * ((c[0] ** 2 + c[1] ** 2 + c[2] ** 2) * (1.0 if i == j else 0.0) - c[i] * c[j]) for j in range(3)] for i in range(3)]
        self.I = [[self.I[i][j] + Ic[i][j] + shift[i][j] for j in range(3)] for i in range(3)]

    def inertia_apply(self, m):  # f = I m  (spatial inertia about the origin: [I, h x; -h x, m])
        w, v = m
        return _add(_mv(self.I, w), _cross(self.h, v)), _sub(_scale(self.mass, v), _cross(self.h, w))


class RobotModel:
    def __init__(self, name, bodies):
        self.name, self.bodies = name, bodies
        self.nq = 7 + sum(1 for b in bodies[1:])
        self.nv = 6 + sum(1 for b in bodies[1:])
        self.njoints = len(bodies) + 1  # Pinocchio counts the "universe" joint

    @property
    def total_mass(self):
        return sum(b.mass for b in self.bodies)


def _floats(text, default):
    return [float(t) for t in text.split()] if text else list(default)


def load_urdf(source: str) -> RobotModel:
    """URDF (file name or XML string) -> tree behind a free-flyer root joint (pinocchio::urdf::buildModel with JointModelFreeFlyer,
    rbd/robot.hpp:43-50).  Revolute / continuous / prismatic joints become 1-DoF joints; links behind fixed joints are merged."""
    root = ET.fromstring(source) if source.lstrip().startswith("<") else ET.parse(source).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    child_of = {j.find("child").get("link"): j for j in joints}
    base = [n for n in links if n not in child_of]
    if len(base) != 1:
        raise ValueError(f"the URDF must have exactly one root link, found {base}")
    children = {}
    for j in joints:
        children.setdefault(j.find("parent").get("link"), []).append(j)

    def origin(elem):
        o = elem.find("origin") if elem is not None else None
        xyz = _floats(o.get("xyz") if o is not None else None, (0, 0, 0))
        rpy = _floats(o.get("rpy") if o is not None else None, (0, 0, 0))
        return _rpy(*rpy), xyz

    bodies = []

    def attach_inertia(body, link, R, p):
        """The link's frame sits at (R, p) in the body's frame."""
        inert = link.find("inertial")
        if inert is None:
            return
        mass = float(inert.find("mass").get("value"))
        Ri, pi = origin(inert)
        it = inert.find("inertia")
        Il = [[float(it.get("ixx")), float(it.get("ixy")), float(it.get("ixz"))],
              [float(it.get("ixy")), float(it.get("iyy")), float(it.get("iyz"))],
              [float(it.get("ixz")), float(it.get("iyz")), float(it.get("izz"))]]
        Rc = _mm(R, Ri)                        # inertial frame in body coordinates
        com = _add(p, _mv(R, pi))
        body.add_inertia(mass, com, _mm(_mm(Rc, Il), _tr(Rc)))

    def visit(link_name, body_index, R, p):
        attach_inertia(bodies[body_index], links[link_name], R, p)
        for j in children.get(link_name, []):
            Rj, pj = origin(j)
            Rc, pc = _mm(R, Rj), _add(p, _mv(R, pj))   # joint frame in the current body's coordinates
            jtype = j.get("type")
            child = j.find("child").get("link")
            if jtype == "fixed":
                visit(child, body_index, Rc, pc)
            elif jtype in ("revolute", "continuous", "prismatic"):
                ax = _floats(j.find("axis").get("xyz") if j.find("axis") is not None else None, (1, 0, 0))
                n = math.sqrt(sum(a * a for a in ax))
                bodies.append(Body(j.get("name"), body_index, Rc, pc, "prismatic" if jtype == "prismatic" else "revolute", [a / n for a in ax]))
                visit(child, len(bodies) - 1, [[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]], [0.0, 0.0, 0.0])
            else:
                raise ValueError(f"joint type {jtype!r} is not supported")

    bodies.append(Body("root_joint", -1, None, None, "free_flyer", None))
    visit(base[0], 0, [[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]], [0.0, 0.0, 0.0])
    return RobotModel(root.get("name"), bodies)


# ---------------------------------------------------------------------------------------------------------------------------
# Kinematics shared by the algorithms
# ---------------------------------------------------------------------------------------------------------------------------
def _joint_transforms(model, q):
    """(E, r) parent->child of every body and the motion subspace S_i of the 1-DoF joints (in child coordinates)."""
    X, S = [], []
    E0 = _tr(_quat_matrix(q[3], q[4], q[5], q[6]))
    X.append((E0, [q[0], q[1], q[2]]))
    S.append(None)
    for i, b in enumerate(model.bodies[1:], start=1):
        qi = q[6 + i]
        if b.jtype == "revolute":
            Rj = _mm(b.R, _rodrigues(b.axis, qi))     # child frame in parent coordinates
            X.append((_tr(Rj), b.p))
            S.append((b.axis, [0.0, 0.0, 0.0]))
        else:
            X.append((_tr(b.R), _add(b.p, _mv(b.R, _scale(qi, b.axis)))))
            S.append(([0.0, 0.0, 0.0], b.axis))
    return X, S


def _split_v(v):
    """Pinocchio's base ordering [linear; angular] -> the (angular, linear) spatial pair."""
    return [v[3], v[4], v[5]], [v[0], v[1], v[2]]


def rnea(model: RobotModel, q, v, a, gravity: float = GRAVITY):
    """Inverse dynamics tau = M(q) a + h(q, v) (pinocchio::rnea; rbd/quantities/joint_torques.hpp)."""
    nb = len(model.bodies)
    X, S = _joint_transforms(model, q)
    vel, acc, f = [None] * nb, [None] * nb, [None] * nb
    a0 = ([0.0, 0.0, 0.0], [0.0, 0.0, gravity])  # fictitious upward acceleration of the world = gravity
    vel[0] = _split_v(v)
    ab = _split_v(a)
    g0 = _xm(X[0][0], X[0][1], a0)
    acc[0] = (_add(g0[0], ab[0]), _add(g0[1], ab[1]))
    for i in range(1, nb):
        b = model.bodies[i]
        vj = (_scale(v[5 + i], S[i][0]), _scale(v[5 + i], S[i][1]))
        vp = _xm(X[i][0], X[i][1], vel[b.parent])
        vel[i] = (_add(vp[0], vj[0]), _add(vp[1], vj[1]))
        ap = _xm(X[i][0], X[i][1], acc[b.parent])
        c = _crm(vel[i], vj)
        acc[i] = (_add(_add(ap[0], _scale(a[5 + i], S[i][0])), c[0]), _add(_add(ap[1], _scale(a[5 + i], S[i][1])), c[1]))
    for i in range(nb):
        b = model.bodies[i]
        Ia, Iv = b.inertia_apply(acc[i]), b.inertia_apply(vel[i])
        c = _crf(vel[i], Iv)
        f[i] = (_add(Ia[0], c[0]), _add(Ia[1], c[1]))
    tau = [0.0] * model.nv
    for i in range(nb - 1, 0, -1):
        tau[5 + i] = sum(S[i][0][k] * f[i][0][k] + S[i][1][k] * f[i][1][k] for k in range(3))
        fp = _xtf(X[i][0], X[i][1], f[i])
        p = model.bodies[i].parent
        f[p] = (_add(f[p][0], fp[0]), _add(f[p][1], fp[1]))
    tau[0:3] = f[0][1]
    tau[3:6] = f[0][0]
    return tau


def _sym6_solve(M, rhs):
    """Solves the symmetric positive definite 6x6 system M x = rhs over a generic scalar (LDL^T, no pivoting)."""
    n = 6
    L = [[0.0] * n for _ in range(n)]
    D = [0.0] * n
    for j in range(n):
        d = M[j][j]
        for k in range(j):
            d = d - L[j][k] * L[j][k] * D[k]
        D[j] = d
        for i in range(j + 1, n):
            s = M[i][j]
            for k in range(j):
                s = s - L[i][k] * L[j][k] * D[k]
            L[i][j] = s / d
    y = list(rhs)
    for i in range(n):
        for k in range(i):
            y[i] = y[i] - L[i][k] * y[k]
    y = [y[i] / D[i] for i in range(n)]
    for i in range(n - 1, -1, -1):
        for k in range(i + 1, n):
            y[i] = y[i] - L[k][i] * y[k]
    return y


def aba(model: RobotModel, q, v, tau, gravity: float = GRAVITY):
    """Forward dynamics a = M(q)^-1 (tau - h(q, v)) by the articulated-body algorithm (pinocchio::aba;
    rbd/quantities/generalized_accelerations.hpp), O(number of bodies), no matrix assembled except the 6x6 of the floating base."""
    nb = len(model.bodies)
    X, S = _joint_transforms(model, q)
    vel, c = [None] * nb, [None] * nb
    vel[0] = _split_v(v)
    for i in range(1, nb):
        b = model.bodies[i]
        vj = (_scale(v[5 + i], S[i][0]), _scale(v[5 + i], S[i][1]))
        vp = _xm(X[i][0], X[i][1], vel[b.parent])
        vel[i] = (_add(vp[0], vj[0]), _add(vp[1], vj[1]))
        c[i] = _crm(vel[i], vj)
    # articulated inertias as 6x6 nested lists in (angular, linear) block order, bias forces as spatial pairs
    IA, pA = [], []
    for i in range(nb):
        b = model.bodies[i]
        hx = [[0.0, -b.h[2], b.h[1]], [b.h[2], 0.0, -b.h[0]], [-b.h[1], b.h[0], 0.0]]
        M6 = [[0.0] * 6 for _ in range(6)]
        for r in range(3):
            for s in range(3):
                M6[r][s] = b.I[r][s]
                M6[r][3 + s] = hx[r][s]
                M6[3 + r][s] = -hx[r][s]
                M6[3 + r][3 + s] = b.mass if r == s else 0.0
        IA.append(M6)
        pA.append(_crf(vel[i], b.inertia_apply(vel[i])))
    U, d, u = [None] * nb, [None] * nb, [None] * nb

    def mat6_vec(M6, m):
        x = list(m[0]) + list(m[1])
        y = [sum(M6[r][k] * x[k] for k in range(6)) for r in range(6)]
        return y[:3], y[3:]

    for i in range(nb - 1, 0, -1):
        Si = list(S[i][0]) + list(S[i][1])
        Ui = [sum(IA[i][r][k] * Si[k] for k in range(6)) for r in range(6)]
        di = sum(Si[k] * Ui[k] for k in range(6))
        ui = tau[5 + i] - sum(Si[k] * (pA[i][0] + pA[i][1])[k] for k in range(6))
        U[i], d[i], u[i] = Ui, di, ui
        Ia = [[IA[i][r][s] - Ui[r] * Ui[s] / di for s in range(6)] for r in range(6)]
        Iac = mat6_vec(Ia, c[i])
        pa = (_add(_add(pA[i][0], Iac[0]), _scale(ui / di, Ui[:3])), _add(_add(pA[i][1], Iac[1]), _scale(ui / di, Ui[3:])))
        # transform to the parent: IA_p += X^T Ia X, pA_p += X^T pa, with X = [E 0; -E rx E]
        E, r = X[i]
        rx = [[0.0, -r[2], r[1]], [r[2], 0.0, -r[0]], [-r[1], r[0], 0.0]]
        Erx = _mm(E, rx)
        X6 = [[0.0] * 6 for _ in range(6)]
        for a_ in range(3):
            for b_ in range(3):
                X6[a_][b_] = E[a_][b_]
                X6[3 + a_][b_] = -Erx[a_][b_]
                X6[3 + a_][3 + b_] = E[a_][b_]
        T = [[sum(Ia[r][k] * X6[k][s] for k in range(6)) for s in range(6)] for r in range(6)]
        p = model.bodies[i].parent
        for r_ in range(6):
            for s_ in range(6):
                IA[p][r_][s_] = IA[p][r_][s_] + sum(X6[k][r_] * T[k][s_] for k in range(6))
        fp = _xtf(E, r, pa)
        pA[p] = (_add(pA[p][0], fp[0]), _add(pA[p][1], fp[1]))
    # floating base: IA_0 a_0 = tau_0 - pA_0 (S = identity), the world accelerates upwards with gravity
    tb = _split_v(tau)
    rhs = [tb[0][k] - pA[0][0][k] for k in range(3)] + [tb[1][k] - pA[0][1][k] for k in range(3)]
    a0 = _sym6_solve(IA[0], rhs)
    acc = [None] * nb
    acc[0] = (a0[:3], a0[3:])            # acceleration of the base relative to the fictitious world, body coordinates
    g0 = _xm(X[0][0], X[0][1], ([0.0, 0.0, 0.0], [0.0, 0.0, gravity]))
    out = [0.0] * model.nv
    base_dd = (_sub(acc[0][0], g0[0]), _sub(acc[0][1], g0[1]))
    out[0:3] = base_dd[1]
    out[3:6] = base_dd[0]
    for i in range(1, nb):
        b = model.bodies[i]
        ap = _xm(X[i][0], X[i][1], acc[b.parent])
        ap = (_add(ap[0], c[i][0]), _add(ap[1], c[i][1]))
        x = list(ap[0]) + list(ap[1])
        qdd = (u[i] - sum(U[i][k] * x[k] for k in range(6))) / d[i]
        out[5 + i] = qdd
        acc[i] = (_add(ap[0], _scale(qdd, S[i][0])), _add(ap[1], _scale(qdd, S[i][1])))
    return out


def crba(model: RobotModel, q):
    """Joint-space inertia matrix (pinocchio::crba; rbd/quantities/joint_space_inertia_matrix.hpp) as columns of unit accelerations
    through rnea without velocity and gravity — O(nv) inverse-dynamics calls, used for tests and small models."""
    zero = [0.0] * model.nv
    cols = []
    for j in range(model.nv):
        e = list(zero)
        e[j] = 1.0
        cols.append(rnea(model, q, zero, e, gravity=0.0))
    return [[cols[j][i] for j in range(model.nv)] for i in range(model.nv)]


# ---------------------------------------------------------------------------------------------------------------------------
# The remaining quantities of include/ungar/rbd/quantities/*.hpp (generic scalar, world frame where Pinocchio uses it)
# ---------------------------------------------------------------------------------------------------------------------------
def _world_kinematics(model, q, v=None, a=None):
    """World pose (R, p) of every body; with v (and a): body-coordinate spatial velocities (and gravity-free accelerations)."""
    X, S = _joint_transforms(model, q)
    nb = len(model.bodies)
    Rw, pw = [None] * nb, [None] * nb
    Rw[0], pw[0] = _tr(X[0][0]), list(X[0][1])
    for i in range(1, nb):
        par = model.bodies[i].parent
        Rw[i] = _mm(Rw[par], _tr(X[i][0]))
        pw[i] = _add(pw[par], _mv(Rw[par], X[i][1]))
    vel = acc = None
    if v is not None:
        vel = [None] * nb
        vel[0] = _split_v(v)
        if a is not None:
            acc = [None] * nb
            acc[0] = _split_v(a)
        for i in range(1, nb):
            par = model.bodies[i].parent
            vj = (_scale(v[5 + i], S[i][0]), _scale(v[5 + i], S[i][1]))
            vp = _xm(X[i][0], X[i][1], vel[par])
            vel[i] = (_add(vp[0], vj[0]), _add(vp[1], vj[1]))
            if a is not None:
                ap = _xm(X[i][0], X[i][1], acc[par])
                c = _crm(vel[i], vj)
                acc[i] = (_add(_add(ap[0], _scale(a[5 + i], S[i][0])), c[0]), _add(_add(ap[1], _scale(a[5 + i], S[i][1])), c[1]))
    return Rw, pw, vel, acc


def com_position(model, q):
    """pinocchio::centerOfMass(q) -> data.com[0]: centre of mass of the whole tree in the world frame."""
    Rw, pw, _, _ = _world_kinematics(model, q)
    tot = [0.0, 0.0, 0.0]
    for i, b in enumerate(model.bodies):
        tot = _add(tot, _add(_scale(b.mass, pw[i]), _mv(Rw[i], b.h)))
    return _scale(1.0 / model.total_mass, tot)


def com_velocity(model, q, v):
    """data.vcom[0]: velocity of the centre of mass in the world frame."""
    Rw, _, vel, _ = _world_kinematics(model, q, v)
    tot = [0.0, 0.0, 0.0]
    for i, b in enumerate(model.bodies):  # m (v_o + w x c) = m v_o + w x h
        tot = _add(tot, _mv(Rw[i], _add(_scale(b.mass, vel[i][1]), _cross(vel[i][0], b.h))))
    return _scale(1.0 / model.total_mass, tot)


def com_acceleration(model, q, v, a):
    """data.acom[0]: classical acceleration of the centre of mass in the world frame (gravity is not part of it)."""
    Rw, _, vel, acc = _world_kinematics(model, q, v, a)
    tot = [0.0, 0.0, 0.0]
    for i, b in enumerate(model.bodies):
        w, vo = vel[i]
        al, ao = acc[i]
        lin = _add(_scale(b.mass, _add(ao, _cross(w, vo))), _add(_cross(al, b.h), _cross(w, _cross(w, b.h))))
        tot = _add(tot, _mv(Rw[i], lin))
    return _scale(1.0 / model.total_mass, tot)


def centroidal_momentum(model, q, v):
    """pinocchio::ccrba -> data.hg: [linear momentum; angular momentum about the centre of mass], world-aligned axes."""
    Rw, pw, vel, _ = _world_kinematics(model, q, v)
    com = com_position(model, q)
    lin, ang = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    for i, b in enumerate(model.bodies):
        n, l = b.inertia_apply(vel[i])                 # momentum about the body origin, body coordinates
        lw = _mv(Rw[i], l)
        lin = _add(lin, lw)
        ang = _add(ang, _add(_mv(Rw[i], n), _cross(_sub(pw[i], com), lw)))
    return lin + ang


def centroidal_momentum_matrix(model, q):
    """data.Ag (6 x nv): hg = Ag v, one unit velocity per column."""
    cols = []
    for j in range(model.nv):
        e = [0.0] * model.nv
        e[j] = 1.0
        cols.append(centroidal_momentum(model, q, e))
    return [[cols[j][i] for j in range(model.nv)] for i in range(6)]


def composite_rigid_body_inertia(model, q):
    """data.Ig: (total mass, rotational inertia of the whole tree about its centre of mass in world-aligned axes, 3 x 3)."""
    Rw, pw, _, _ = _world_kinematics(model, q)
    com = com_position(model, q)
    I = [[0.0] * 3 for _ in range(3)]
    for i, b in enumerate(model.bodies):
        # inertia about the body origin rotated to world axes, moved to the centre of mass with the parallel-axis theorem twice:
        # I_com_total += R I_o R^T - m [(c.c) 1 - c c^T] + m [(d.d) 1 - d d^T],  c = body com - body origin, d = body com - total com
        Iw = _mm(_mm(Rw[i], b.I), _tr(Rw[i]))
        if b.mass == 0.0:
            continue
        c = _mv(Rw[i], _scale(1.0 / b.mass, b.h))
        d = _sub(_add(pw[i], c), com)
        cc, dd = c[0] * c[0] + c[1] * c[1] + c[2] * c[2], d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
        for r in range(3):
            for s_ in range(3):
                I[r][s_] = I[r][s_] + Iw[r][s_] - b.mass * ((cc if r == s_ else 0.0) - c[r] * c[s_]) + b.mass * ((dd if r == s_ else 0.0) - d[r] * d[s_])
    return model.total_mass, I


def kinetic_energy(model, q, v):
    """pinocchio::computeKineticEnergy."""
    _, _, vel, _ = _world_kinematics(model, q, v)
    e = 0.0
    for i, b in enumerate(model.bodies):
        n, l = b.inertia_apply(vel[i])
        e = e + 0.5 * sum(vel[i][0][k] * n[k] + vel[i][1][k] * l[k] for k in range(3))
    return e


def potential_energy(model, q, gravity: float = GRAVITY):
    """pinocchio::computePotentialEnergy: -m g . com with g = (0, 0, -gravity)."""
    return model.total_mass * gravity * com_position(model, q)[2]


def generalized_gravity(model, q, gravity: float = GRAVITY):
    """pinocchio::computeGeneralizedGravity: rnea(q, 0, 0)."""
    z = [0.0] * model.nv
    return rnea(model, q, z, z, gravity)


def nonlinear_effects(model, q, v, gravity: float = GRAVITY):
    """pinocchio::nonLinearEffects: rnea(q, v, 0) = Coriolis + centrifugal + gravity."""
    return rnea(model, q, v, [0.0] * model.nv, gravity)


def joint_space_inertia_matrix_inverse(model, q):
    """pinocchio::computeMinverse: columns of unit generalized forces through the articulated-body algorithm (no velocity, no gravity)."""
    zero = [0.0] * model.nv
    cols = []
    for j in range(model.nv):
        e = list(zero)
        e[j] = 1.0
        cols.append(aba(model, q, zero, e, gravity=0.0))
    return [[cols[j][i] for j in range(model.nv)] for i in range(model.nv)]


def frames(model, q):
    """pinocchio::framesForwardKinematics -> data.oMf for the movable bodies: {joint name: (R world<-body, p world)}."""
    Rw, pw, _, _ = _world_kinematics(model, q)
    return {b.name: (Rw[i], pw[i]) for i, b in enumerate(model.bodies)}


def integrate(model, q, v, dt: float = 1.0):
    """pinocchio::integrate(q, v dt): the base moves by its body-frame twist (SE(3) exponential), the joints by v dt.  Floats only."""
    w = np.array(v[3:6], dtype=float) * dt
    u = np.array(v[0:3], dtype=float) * dt
    th = float(np.linalg.norm(w))
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])
    if th < 1e-9:
        Rinc, V = np.eye(3) + K, np.eye(3) + 0.5 * K
    else:
        Rinc = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * K @ K
    R = np.array(_quat_matrix(*[float(x) for x in q[3:7]]))
    p = np.array(q[0:3], dtype=float) + R @ (V @ u)
    Rn = R @ Rinc
    qw = 0.5 * math.sqrt(max(1e-300, 1.0 + Rn[0, 0] + Rn[1, 1] + Rn[2, 2]))
    quat = [(Rn[2, 1] - Rn[1, 2]) / (4 * qw), (Rn[0, 2] - Rn[2, 0]) / (4 * qw), (Rn[1, 0] - Rn[0, 1]) / (4 * qw), qw]
    return list(p) + quat + [float(q[7 + i]) + float(v[6 + i]) * dt for i in range(model.nq - 7)]


QUANTITIES = {  # include/ungar/rbd/quantities/*.hpp -> (function, arguments)
    "generalized_accelerations": (aba, "qvt"), "joint_torques": (rnea, "qva"), "joint_space_inertia_matrix": (crba, "q"),
    "joint_space_inertia_matrix_inverse": (joint_space_inertia_matrix_inverse, "q"), "generalized_gravity": (generalized_gravity, "q"),
    "nonlinear_effects": (nonlinear_effects, "qv"), "com_position": (com_position, "q"), "com_velocity": (com_velocity, "qv"),
    "com_acceleration": (com_acceleration, "qva"), "centroidal_momentum": (centroidal_momentum, "qv"),
    "centroidal_momentum_matrix": (centroidal_momentum_matrix, "q"), "composite_rigid_body_inertia": (composite_rigid_body_inertia, "q"),
    "kinetic_energy": (kinetic_energy, "qv"), "potential_energy": (potential_energy, "q"), "frames": (frames, "q"),
}


def _flatten(y):
    """Quantity value -> flat list of scalars (matrices row-major; the composite inertia as [mass, I row-major]; frames as
    [R row-major, p] per body in tree order)."""
    if isinstance(y, dict):
        return [x for R, p in y.values() for x in ([e for row in R for e in row] + list(p))]
    if isinstance(y, tuple):
        return [y[0]] + [e for row in y[1] for e in row]
    if isinstance(y, list) and y and isinstance(y[0], list):
        return [e for row in y for e in row]
    return list(y) if isinstance(y, list) else [y]


class _Evaluator:
    """``robot.Compute(quantity).At(q, v, ...)`` (rbd/evaluator.hpp:45-58): evaluates the taped quantity on the GPU."""

    def __init__(self, robot, quantity):
        self.robot, self.quantity = robot, quantity

    def At(self, *args):
        f = self.robot.MakeFunction(self.quantity)
        x = np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in args])
        self.robot._results[self.quantity] = np.asarray(f(x))
        return self


class Robot:
    """Mirror of ``Ungar::Robot`` (rbd/robot.hpp:40-104): ``Compute(quantity).At(q, v, ...)`` / ``Get(quantity)`` for the fifteen
    quantities of rbd/quantities/*.hpp, each a taped Function of the stacked arguments evaluated by the register machine."""

    def __init__(self, urdf: str, gravity: float = GRAVITY, device: int = 0):
        self.model = load_urdf(urdf)
        self.gravity, self.device = gravity, device
        self._functions, self._results = {}, {}

    def Model(self):
        return self.model

    def Compute(self, quantity: str) -> _Evaluator:
        return _Evaluator(self, quantity)

    def Get(self, quantity: str):
        return self._results[quantity]

    def MakeFunction(self, quantity: str = "generalized_accelerations", name: str | None = None, scale: float = 1.0):
        """The Autodiff::Function of test/rbd/robot.test.cpp:121-133: input = the quantity's arguments stacked ([q; v; tau] for the
        generalized accelerations), output = the flattened quantity times ``scale``; Jacobian enabled."""
        key = (quantity, scale)
        if key in self._functions:
            return self._functions[key]
        m = self.model
        algo, args = QUANTITIES[quantity]
        sizes = [m.nq if c == "q" else m.nv for c in args]
        takes_gravity = quantity in ("generalized_accelerations", "joint_torques", "generalized_gravity", "nonlinear_effects", "potential_energy")

        def impl(x):
            parts, off = [], 0
            for n in sizes:
                parts.append(x[off:off + n])
                off += n
            y = algo(m, *parts, self.gravity) if takes_gravity else algo(m, *parts)
            return [scale * e for e in _flatten(y)]

        bp = A.Blueprint(impl, sum(sizes), 0, name or f"{m.name}_{quantity}", A.JACOBIAN)
        x0 = []  # tape at a regular point: unit quaternion, distinct small joint angles, small rates
        for c in args:
            x0 += ([0.0] * 6 + [1.0] + [0.1 * k for k in range(1, m.nq - 6)]) if c == "q" else [0.05] * m.nv
        bp.tapingPoint = np.array(x0)
        f = A.MakeFunction(bp, device=self.device)
        self._functions[key] = f
        return f
