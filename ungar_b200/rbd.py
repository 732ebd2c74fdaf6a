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


class Robot:
    """Mirror of ``Ungar::Robot`` (rbd/robot.hpp:40-104) for the quantities that are functions of (q, v, tau) / (q, v, a)."""

    def __init__(self, urdf: str, gravity: float = GRAVITY):
        self.model = load_urdf(urdf)
        self.gravity = gravity

    def Model(self):
        return self.model

    def MakeFunction(self, quantity: str = "generalized_accelerations", name: str | None = None, scale: float = 1.0, device: int = 0):
        """The Autodiff::Function of test/rbd/robot.test.cpp:121-133: input [q; v; tau] (or [q; v; a] for "joint_torques"), output
        the quantity times ``scale``; values and Jacobian are evaluated on the GPU by the register machine."""
        m = self.model
        algo = {"generalized_accelerations": aba, "joint_torques": rnea}[quantity]

        def impl(x):
            q, v, w = x[:m.nq], x[m.nq:m.nq + m.nv], x[m.nq + m.nv:]
            return [scale * y for y in algo(m, q, v, w, self.gravity)]

        bp = A.Blueprint(impl, m.nq + 2 * m.nv, 0, name or f"{m.name}_{quantity}", A.JACOBIAN)
        # tape at a regular configuration: unit quaternion, small joint angles
        x0 = np.zeros(m.nq + 2 * m.nv)
        x0[6] = 1.0
        x0[7:m.nq] = 0.1 * np.arange(1, m.nq - 6)
        x0[m.nq:] = 0.05
        bp.tapingPoint = x0
        return A.MakeFunction(bp, device=device)
