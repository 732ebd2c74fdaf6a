"""ORACLE — TEST INFRASTRUCTURE ONLY.  Independent numpy statement of rigid-body dynamics for a URDF tree behind a free-flyer root,
used to check ``ungar_b200/rbd.py`` (the algorithms that get taped for the GPU).

Different algorithm and different code from the product on purpose: forward dynamics here is  a = M(q)^-1 (tau - h(q, v))  with
M from the composite-rigid-body algorithm and h from recursive Newton-Euler, everything in dense 6x6 spatial matrices
(Featherstone, "Rigid Body Dynamics Algorithms", tables 5.1 and 6.2), whereas the product tapes the articulated-body algorithm
written on 3-vectors.  Conventions are Pinocchio's (the library behind include/ungar/rbd/robot.hpp:40-104, absent here):
q = [p, quaternion (x, y, z, w), joint angles]; v = [linear, angular] of the base in the base frame, then joint rates.

PARITY UNPINNED: the reference pins its rbd layer only against Pinocchio itself (test/rbd/robot.test.cpp:109-162), which is not in
/root/reference nor in this image; what this oracle pins is algorithm-vs-algorithm agreement plus physical identities.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET

import numpy as np


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def rpy_matrix(r, p, y):
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def axis_angle(axis, q):
    K = skew(axis)
    return np.eye(3) + np.sin(q) * K + (1.0 - np.cos(q)) * K @ K


def quat_matrix(x, y, z, w):
    q = np.array([x, y, z])
    return (w * w - q @ q) * np.eye(3) + 2.0 * np.outer(q, q) + 2.0 * w * skew(q)


def plucker(E, r):
    """Motion transform parent -> child: rotation E (parent coords -> child coords), child origin r in parent coords."""
    X = np.zeros((6, 6))
    X[:3, :3] = E
    X[3:, 3:] = E
    X[3:, :3] = -E @ skew(r)
    return X


def crm(v):
    M = np.zeros((6, 6))
    M[:3, :3] = skew(v[:3])
    M[3:, 3:] = skew(v[:3])
    M[3:, :3] = skew(v[3:])
    return M


def spatial_inertia(mass, com, Ic):
    C = skew(com)
    M = np.zeros((6, 6))
    M[:3, :3] = Ic + mass * C @ C.T
    M[:3, 3:] = mass * C
    M[3:, :3] = mass * C.T
    M[3:, 3:] = mass * np.eye(3)
    return M


class Tree:
    """parent[i], fixed transform to the joint frame (Rt, pt), joint type/axis and 6x6 inertia of every movable body (0 = base)."""

    def __init__(self, urdf: str):
        root = ET.fromstring(urdf) if urdf.lstrip().startswith("<") else ET.parse(urdf).getroot()
        links = {l.get("name"): l for l in root.findall("link")}
        joints = root.findall("joint")
        is_child = {j.find("child").get("link") for j in joints}
        base = [n for n in links if n not in is_child][0]
        self.parent, self.Rt, self.pt, self.kind, self.axis, self.I = [-1], [None], [None], ["free"], [None], [np.zeros((6, 6))]

        def pose(elem):
            o = elem.find("origin") if elem is not None else None
            xyz = np.array([float(t) for t in (o.get("xyz") if o is not None and o.get("xyz") else "0 0 0").split()])
            rpy = [float(t) for t in (o.get("rpy") if o is not None and o.get("rpy") else "0 0 0").split()]
            return rpy_matrix(*rpy), xyz

        def walk(link, body, R, p):
            inert = links[link].find("inertial")
            if inert is not None:
                Ri, pi = pose(inert)
                it = inert.find("inertia")
                Il = np.array([[float(it.get("ixx")), float(it.get("ixy")), float(it.get("ixz"))],
                               [float(it.get("ixy")), float(it.get("iyy")), float(it.get("iyz"))],
                               [float(it.get("ixz")), float(it.get("iyz")), float(it.get("izz"))]])
                Rc = R @ Ri
                self.I[body] = self.I[body] + spatial_inertia(float(inert.find("mass").get("value")), p + R @ pi, Rc @ Il @ Rc.T)
            for j in joints:
                if j.find("parent").get("link") != link:
                    continue
                Rj, pj = pose(j)
                Rc, pc = R @ Rj, p + R @ pj
                child = j.find("child").get("link")
                if j.get("type") == "fixed":
                    walk(child, body, Rc, pc)
                else:
                    ax = j.find("axis")
                    a = np.array([float(t) for t in (ax.get("xyz") if ax is not None else "1 0 0").split()])
                    self.parent.append(body)
                    self.Rt.append(Rc)
                    self.pt.append(pc)
                    self.kind.append("prismatic" if j.get("type") == "prismatic" else "revolute")
                    self.axis.append(a / np.linalg.norm(a))
                    self.I.append(np.zeros((6, 6)))
                    walk(child, len(self.parent) - 1, np.eye(3), np.zeros(3))

        walk(base, 0, np.eye(3), np.zeros(3))
        self.nb = len(self.parent)
        self.nq, self.nv = 7 + self.nb - 1, 6 + self.nb - 1

    def transforms(self, q):
        X = [plucker(quat_matrix(*q[3:7]).T, q[:3])]
        S = [np.eye(6)]
        for i in range(1, self.nb):
            if self.kind[i] == "revolute":
                R = self.Rt[i] @ axis_angle(self.axis[i], q[6 + i])
                X.append(plucker(R.T, self.pt[i]))
                S.append(np.concatenate([self.axis[i], np.zeros(3)])[:, None])
            else:
                X.append(plucker(self.Rt[i].T, self.pt[i] + self.Rt[i] @ (q[6 + i] * self.axis[i])))
                S.append(np.concatenate([np.zeros(3), self.axis[i]])[:, None])
        return X, S


P = np.block([[np.zeros((3, 3)), np.eye(3)], [np.eye(3), np.zeros((3, 3))]])  # [linear; angular] <-> [angular; linear]


def bias_forces(tree: Tree, q, v, gravity=9.81):
    """h(q, v) = rnea(q, v, 0) in Pinocchio's ordering."""
    X, S = tree.transforms(q)
    vj = [P @ v[:6]] + [S[i][:, 0] * v[5 + i] for i in range(1, tree.nb)]
    vel, acc, f = [None] * tree.nb, [None] * tree.nb, [None] * tree.nb
    a_world = np.array([0, 0, 0, 0, 0, gravity])
    for i in range(tree.nb):
        if i == 0:
            vel[0], acc[0] = vj[0], X[0] @ a_world
        else:
            p = tree.parent[i]
            vel[i] = X[i] @ vel[p] + vj[i]
            acc[i] = X[i] @ acc[p] + crm(vel[i]) @ vj[i]
        f[i] = tree.I[i] @ acc[i] - crm(vel[i]).T @ (tree.I[i] @ vel[i])
    h = np.zeros(tree.nv)
    for i in range(tree.nb - 1, 0, -1):
        h[5 + i] = S[i][:, 0] @ f[i]
        f[tree.parent[i]] = f[tree.parent[i]] + X[i].T @ f[i]
    h[:6] = P @ f[0]
    return h


def mass_matrix(tree: Tree, q):
    """Composite-rigid-body algorithm (Featherstone table 6.2), Pinocchio's ordering."""
    X, S = tree.transforms(q)
    Ic = [m.copy() for m in tree.I]
    for i in range(tree.nb - 1, 0, -1):
        Ic[tree.parent[i]] = Ic[tree.parent[i]] + X[i].T @ Ic[i] @ X[i]
    M = np.zeros((tree.nv, tree.nv))
    cols = [list(range(6))] + [[5 + i] for i in range(1, tree.nb)]
    for i in range(tree.nb):
        F = Ic[i] @ S[i]
        M[np.ix_(cols[i], cols[i])] = S[i].T @ F
        j = i
        while tree.parent[j] >= 0:
            F = X[j].T @ F
            j = tree.parent[j]
            M[np.ix_(cols[i], cols[j])] = F.T @ S[j]
            M[np.ix_(cols[j], cols[i])] = (F.T @ S[j]).T
    Pn = np.eye(tree.nv)
    Pn[:6, :6] = P
    return Pn @ M @ Pn.T


def forward_dynamics(tree: Tree, q, v, tau, gravity=9.81):
    return np.linalg.solve(mass_matrix(tree, q), tau - bias_forces(tree, q, v, gravity))


def kinetic_energy(tree: Tree, q, v):
    X, S = tree.transforms(q)
    vel = [P @ v[:6]]
    for i in range(1, tree.nb):
        vel.append(X[i] @ vel[tree.parent[i]] + S[i][:, 0] * v[5 + i])
    return 0.5 * sum(vel[i] @ tree.I[i] @ vel[i] for i in range(tree.nb))


def world_poses(tree: Tree, q):
    """(R, p) of every movable body in the world, from the products of the Pluecker transforms."""
    X, _ = tree.transforms(q)
    Xw = [X[0]]
    for i in range(1, tree.nb):
        Xw.append(X[i] @ Xw[tree.parent[i]])   # world -> body i
    poses = []
    for Xi in Xw:
        E = Xi[:3, :3]
        rx = -E.T @ Xi[3:, :3]                  # skew(r)
        poses.append((E.T, np.array([rx[2, 1], rx[0, 2], rx[1, 0]])))
    return poses, Xw


def body_mass_com(I6):
    m = I6[3, 3]
    h = np.array([I6[2, 4], I6[0, 5], I6[1, 3]])   # m * skew(c) block
    return m, (h / m if m > 0 else np.zeros(3))


def com(tree: Tree, q):
    poses, _ = world_poses(tree, q)
    tot, M = np.zeros(3), 0.0
    for (R, p), I6 in zip(poses, tree.I):
        m, c = body_mass_com(I6)
        tot += m * (p + R @ c)
        M += m
    return tot / M, M


def centroidal_momentum(tree: Tree, q, v):
    """[linear; angular about the centre of mass] in world-aligned axes: sum of the body momenta moved with force transforms."""
    X, S = tree.transforms(q)
    poses, Xw = world_poses(tree, q)
    c, _ = com(tree, q)
    vel = [P @ v[:6]]
    for i in range(1, tree.nb):
        vel.append(X[i] @ vel[tree.parent[i]] + S[i][:, 0] * v[5 + i])
    Xc = plucker(np.eye(3), c)                   # world -> frame at the centre of mass, world axes
    h = np.zeros(6)
    for i in range(tree.nb):
        # momentum of body i in its own coordinates -> world coordinates (X^T) -> centre-of-mass frame (X^-T)
        h += np.linalg.inv(Xc).T @ (Xw[i].T @ (tree.I[i] @ vel[i]))
    return P @ h
