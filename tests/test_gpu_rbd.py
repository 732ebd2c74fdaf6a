"""Batched rigid-body dynamics of ANYmal B on the GPU (reference: rbd/robot.hpp:40-104, rbd/evaluator.hpp:45-58,
rbd/quantities/generalized_accelerations.hpp, test/rbd/robot.test.cpp:109-162).

The batched ABA kernel is GENERATED: ungar_b200/rbd.py states Featherstone's articulated-body algorithm over the tracing scalar, the
tree of the robot is unrolled when the function is made (Robot.MakeFunction: 8.4 k live tape nodes for ANYmal B, nq 19 / nv 18), and
csrc/tape.cu hands the straight-line program to NVRTC: one sm_100a kernel per order (values; Jacobian by column colours), one thread per
(state, colour), every intermediate in a register.  These tests run that kernel on 1024 random (q, v, tau) and compare the accelerations
and their derivatives with the INDEPENDENT numpy oracle (oracle/rbd_reference.py, M^-1 (tau - h) on 6 x 6 spatial matrices) through the
committed fixtures of oracle/make_golden_rbd.py (the URDF lives in /root/reference and does not travel to the GPU box).

Parity with Pinocchio — what the reference's own test compares against — is UNPINNED: Pinocchio is not available offline."""
import os

import numpy as np
import pytest

from ungar_b200 import autodiff as A

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NQ, NV = 19, 18


def fixtures():
    return np.load(os.path.join(GOLDEN, "rbd_anymal_b.npz")), os.path.join(GOLDEN, "rbd_anymal_b_tape.npz")


def test_the_fixture_tape_is_the_tape_of_the_product_algorithm():
    """Host-side analysis only (no GPU): the committed tape has the sizes of robot.test.cpp:89-107 and — where the reference's URDF is
    present — is exactly what Robot.MakeFunction records today."""
    z, tape = fixtures()
    f = A.TapeFunction.load(tape)
    assert (f.IndependentVariableSize(), f.ParameterSize(), f.DependentVariableSize()) == (NQ + 2 * NV, 0, NV)
    info = f.tape_info()
    assert info["live_nodes"] > 5000 and info["slots"] < 400 and info["jacobian_colors"] <= NQ + 2 * NV
    rows, cols = f.JacobianSparsity()
    assert not np.any(cols < 3) and set(range(3, NQ + 2 * NV)) == set(cols.tolist())  # gravity is uniform: no dependence on the position
    urdf = "/root/reference/data/robots/anymal_b_description/robots/anymal.urdf"
    if os.path.exists(urdf):
        from ungar_b200 import rbd

        g = rbd.Robot(urdf).MakeFunction("generalized_accelerations")
        stored = np.load(tape)
        assert np.array_equal(g._recorded[0], stored["nodes"].astype(A.NODE_DTYPE)) and np.array_equal(g._recorded[1], stored["dependents"])


@pytest.mark.gpu
def test_batched_aba_of_anymal_b_matches_the_oracle_at_1024_states():
    z, tape = fixtures()
    X, A_ref = z["X"], z["A"]
    f = A.TapeFunction.load(tape)
    a1 = f(X)                       # first call of the order: the register machine (interpreter)
    a2 = f(X)                       # second call: the NVRTC-specialised straight-line kernel
    assert f._tape.special_info()[0]["state"] == 1, "the batched ABA kernel was not generated (NVRTC unavailable?)"
    scale = np.maximum(1.0, np.abs(A_ref).max(axis=1, keepdims=True))
    assert np.max(np.abs(a2 - A_ref) / scale) <= 1e-9
    assert np.max(np.abs(a1 - a2) / scale) <= 1e-12    # same op functions, another instruction order
    # device buffers: the same kernel on a torch tensor, results stay in HBM
    import torch

    a3 = f(torch.from_numpy(X).cuda()).cpu().numpy()
    assert np.array_equal(a3, a2)


@pytest.mark.gpu
def test_aba_derivatives_of_anymal_b_match_the_oracle_along_the_configuration_manifold():
    """d a / d (q, v, tau) from the Jacobian-order kernel (52 colours) against central differences of the oracle.  q lives on
    R^3 x S^3 x R^12: the derivative with respect to the raw quaternion entries depends on how an implementation extends the rotation
    off the unit sphere, so the comparison is made along the tangent of pinocchio::integrate, J_q T with T = d q / d xi."""
    z, tape = fixtures()
    X, D_ref, T = z["X"], z["D"], z["T"]
    f = A.TapeFunction.load(tape)
    rows, cols = f.JacobianSparsity()
    vals = f.JacobianValues(X)      # interpreter
    vals = f.JacobianValues(X)      # specialised kernel, 1024 states x 52 colours
    assert f._tape.special_info()[1]["state"] == 1
    assert vals.shape == (X.shape[0], rows.size) and np.isfinite(vals).all()
    for b in range(D_ref.shape[0]):
        J = np.zeros((NV, NQ + 2 * NV))
        J[rows, cols] = vals[b]
        Jt = np.concatenate([J[:, :NQ] @ T[b], J[:, NQ:]], axis=1)
        assert np.max(np.abs(Jt - D_ref[b])) <= 2e-6 * max(1.0, np.abs(D_ref[b]).max())   # bounded by the finite differences
        Minv = J[:, NQ + NV:]                                                                # d a / d tau = M^-1
        assert np.allclose(Minv, Minv.T, rtol=1e-9, atol=1e-11) and np.linalg.eigvalsh(0.5 * (Minv + Minv.T)).min() > 0.0
    # the single-vector call of the reference's Function::Jacobian (function.hpp:196-230) agrees with the batched one
    J0 = f.Jacobian(X[5]).toarray()
    J5 = np.zeros_like(J0)
    J5[rows, cols] = vals[5]
    assert np.allclose(J0, J5, rtol=1e-12, atol=1e-13)
