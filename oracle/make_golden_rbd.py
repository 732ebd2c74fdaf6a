#!/usr/bin/env python
"""ORACLE — TEST INFRASTRUCTURE ONLY.  Writes the rigid-body-dynamics fixtures of SURVEY.md §8f-3 for ANYmal B (nq 19 / nv 18), the
robot of test/rbd/robot.test.cpp:89-162, whose URDF lives in /root/reference and therefore cannot travel to the GPU box:

  tests/golden/rbd_anymal_b_tape.npz   the recorded tape of the generalized accelerations a = ABA(q, v, tau) (Robot.MakeFunction,
                                       i.e. the Autodiff::Function of robot.test.cpp:121-133), so that the GPU test can rebuild the
                                       function — and NVRTC the batched straight-line kernel — without the URDF
  tests/golden/rbd_anymal_b.npz        1024 seeded states X = [q; v; tau], the accelerations of the INDEPENDENT numpy oracle
                                       (oracle/rbd_reference.py: M^-1 (tau - h) with CRBA + RNEA on 6 x 6 spatial matrices) for all
                                       of them, and for the first 16 states the derivative of the oracle along the configuration
                                       manifold: d a / d (xi, v, tau) by central differences, xi = the body-frame tangent of
                                       pinocchio::integrate, together with the tangent map T = d q / d xi (19 x 18) so that the test can
                                       compare  [J_q T, J_v, J_tau]  of the device Jacobian with it.

Parity with Pinocchio itself is UNPINNED (absent here); the oracle is another algorithm in other code.
Re-run after a change of rbd.py or of the oracle:  python oracle/make_golden_rbd.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import rbd_reference as R  # noqa: E402

ANYMAL = "/root/reference/data/robots/anymal_b_description/robots/anymal.urdf"
OUT = os.path.join(ROOT, "tests", "golden")
B, B_JAC, SEED = 1024, 16, 11


def retract(q, xi):
    """q (+) xi: base position moves by R(q) xi_lin, the quaternion (x, y, z, w) by the body-frame rotation vector xi_ang (exact
    exponential), the joints by xi_joints — the first-order behaviour of pinocchio::integrate, which is all a derivative needs."""
    q = np.array(q, dtype=float)
    Rm = R.quat_matrix(*q[3:7])
    out = q.copy()
    out[0:3] = q[0:3] + Rm @ xi[0:3]
    w = xi[3:6]
    th = np.linalg.norm(w)
    dq = np.concatenate([0.5 * w, [1.0]]) if th < 1e-12 else np.concatenate([np.sin(th / 2) * w / th, [np.cos(th / 2)]])
    x1, y1, z1, w1 = q[3:7]
    x2, y2, z2, w2 = dq
    out[3:7] = [w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2]
    out[7:] = q[7:] + xi[6:]
    return out


def main():
    if not os.path.exists(ANYMAL):
        raise SystemExit("the ANYmal B URDF of the reference is needed (/root/reference)")
    from ungar_b200 import rbd

    robot, tree = rbd.Robot(ANYMAL), R.Tree(ANYMAL)
    m = robot.Model()
    assert (m.nq, m.nv) == (19, 18) == (tree.nq, tree.nv)
    f = robot.MakeFunction("generalized_accelerations")
    f.save(os.path.join(OUT, "rbd_anymal_b_tape.npz"))
    rng = np.random.default_rng(SEED)
    X = np.zeros((B, m.nq + 2 * m.nv))
    for b in range(B):
        q = rng.standard_normal(m.nq)
        q[3:7] /= np.linalg.norm(q[3:7])
        X[b] = np.concatenate([q, rng.standard_normal(m.nv), 5.0 * rng.standard_normal(m.nv)])
    nq, nv = m.nq, m.nv
    fd = lambda x: R.forward_dynamics(tree, x[:nq], x[nq:nq + nv], x[nq + nv:], rbd.GRAVITY)  # noqa: E731
    A = np.stack([fd(X[b]) for b in range(B)])
    D = np.zeros((B_JAC, nv, 3 * nv))
    T = np.zeros((B_JAC, nq, nv))
    eps = 1e-6
    for b in range(B_JAC):
        x = X[b]
        for k in range(nv):
            e = np.zeros(nv)
            e[k] = eps
            qp, qm = retract(x[:nq], e), retract(x[:nq], -e)
            T[b, :, k] = (qp - qm) / (2 * eps)
            D[b, :, k] = (fd(np.concatenate([qp, x[nq:]])) - fd(np.concatenate([qm, x[nq:]]))) / (2 * eps)
        for k in range(2 * nv):
            e = np.zeros(nq + 2 * nv)
            e[nq + k] = eps * max(1.0, abs(x[nq + k]))
            D[b, :, nv + k] = (fd(x + e) - fd(x - e)) / (2 * e[nq + k])
    np.savez_compressed(os.path.join(OUT, "rbd_anymal_b.npz"), X=X, A=A, D=D, T=T, gravity=rbd.GRAVITY, seed=SEED)
    print("tape:", f.tape_info(), " states:", X.shape, " max |a|:", np.abs(A).max())


if __name__ == "__main__":
    main()
