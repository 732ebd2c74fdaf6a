"""Batched QP solve against the oracle: one sparse LU of the KKT system assembled from the same record
(oracle/qp_reference.py::kkt_solve), plus size-independent optimality properties at the full BASELINE batch."""
import numpy as np
import pytest

from ungar_b200 import EXAMPLE_BARRIER
from ungar_b200 import workloads as W

pytestmark = pytest.mark.gpu


def layout_for_reference(model):
    L = dict(model.layout)
    return L


@pytest.mark.parametrize("N,perturb", [(6, False), (30, True), (100, True), (100, False)])
def test_qp_step_matches_sparse_kkt_solve(oracle, N, perturb):
    import torch

    import ungar_b200
    from oracle import qp_reference as Q

    k, eps = EXAMPLE_BARRIER[W.QUADRUPED]
    model = ungar_b200.Model("quadruped", N, dtype="f64", barrier=(k, eps))
    xp = W.synthetic_batch(W.QUADRUPED, N, 4, seed=41, perturb_params=perturb)
    rec = model.kkt_blocks(torch.from_numpy(xp).cuda(), torch.zeros((4, model.layout["size"]), dtype=torch.float64, device="cuda"))
    steps, mult = model.qp_solve(rec)
    torch.cuda.synchronize()
    steps, mult, rec_h = steps.cpu().numpy(), mult.cpu().numpy(), rec.cpu().numpy()
    L = layout_for_reference(model)
    for b in range(4):
        d_ref, lam_ref = Q.kkt_solve(rec_h[b], L)
        assert np.max(np.abs(steps[b] - d_ref)) <= 1e-7 * np.max(np.abs(d_ref)), (b, np.max(np.abs(steps[b] - d_ref)) / np.max(np.abs(d_ref)))
        assert np.max(np.abs(mult[b] - lam_ref)) <= 1e-6 * max(1e-12, np.max(np.abs(lam_ref)))


def test_qp_optimality_at_full_batch():
    """KKT residuals of every trajectory of BASELINE config 4: A d = -g and P d + q + A^T lambda = 0."""
    import torch

    import ungar_b200
    from oracle import qp_reference as Q

    N, B = 100, 1024
    k, eps = EXAMPLE_BARRIER[W.QUADRUPED]
    model = ungar_b200.Model("quadruped", N, dtype="f64", barrier=(k, eps))
    xp = W.synthetic_batch(W.QUADRUPED, N, B, seed=101)
    d_xp = torch.from_numpy(xp).cuda()
    rec = model.kkt_blocks(d_xp, torch.zeros((B, model.layout["size"]), dtype=torch.float64, device="cuda"))
    steps, mult = model.qp_solve(rec)
    torch.cuda.synchronize()
    assert torch.isfinite(steps).all() and torch.isfinite(mult).all()
    # residuals of a sample through the reference-format sparse matrices
    L = model.layout
    for b in (0, B // 3, B - 1):
        x = xp[b]
        Jg = model.equalityConstraints.Jacobian(x)
        g = model.equalityConstraints(x)
        d, lam = steps[b].cpu().numpy(), mult[b].cpu().numpy()
        assert np.max(np.abs(Jg @ d + g)) < 1e-6 * max(1.0, np.max(np.abs(g)))  # A d = -g (up to the 1e-9 regulariser)
        H, q, U, V, bb = Q.stage_blocks(rec[b].cpu().numpy(), dict(L))
        r = Q.kkt_solve(rec[b].cpu().numpy(), dict(L))[0]
        assert np.max(np.abs(d - r)) <= 1e-7 * np.max(np.abs(r))
    # determinism
    steps2, _ = model.qp_solve(rec)
    assert torch.equal(steps2, steps)


@pytest.mark.parametrize("name,N", [("quadrotor", 30), ("quadrotor", 7), ("rc_car", 60), ("rc_car", 2)])
def test_riccati_qp_matches_monolithic_kkt_solve(oracle, name, N):
    """Quadrotor / RC car: the Riccati step and multipliers against one sparse LU of the KKT system assembled the way
    AssembleOSQPInstance does (oracle/sqp_reference.py::monolithic_qp, soft_sqp.hpp:141-158) — including the u_k - u_{k+1}
    coupling of the input-rate cost, which the recursion carries in the augmented state."""
    import torch

    import ungar_b200
    from oracle import sqp_reference as S

    mid = W.MODEL_IDS[name]
    k, eps = EXAMPLE_BARRIER[mid]
    model = ungar_b200.Model(name, N, dtype="f64", barrier=(k, eps))
    B = 5
    xp = W.synthetic_batch(mid, N, B, seed=43)
    xp[1, model.layout["nx"] * (N + 1):model.layout["n_dec"]] *= 1.5  # push some inputs into the barrier's active region
    rec = model.kkt_blocks(torch.from_numpy(xp).cuda(), torch.zeros((B, model.layout["size"]), dtype=torch.float64, device="cuda"))
    steps, mult = model.qp_solve(rec)
    torch.cuda.synchronize()
    steps, mult = steps.cpu().numpy(), mult.cpu().numpy()
    for b in range(B):
        P, q, A, g, _ = S.monolithic_qp(oracle, mid, N, xp[b], k, eps)
        d_ref, lam_ref = S.solve_qp(P, q, A, g, delta=0.0)
        assert np.max(np.abs(steps[b] - d_ref)) <= 1e-7 * np.max(np.abs(d_ref)), (b, np.max(np.abs(steps[b] - d_ref)) / np.max(np.abs(d_ref)))
        assert np.max(np.abs(mult[b] - lam_ref)) <= 1e-6 * max(1e-12, np.max(np.abs(lam_ref)))
        # optimality of the device solution on its own: A d = -g,  P d + q + A^T lambda = 0
        assert np.max(np.abs(A @ steps[b] + g)) <= 1e-9 * max(1.0, np.max(np.abs(g)))
        assert np.max(np.abs(P @ steps[b] + q + A.T @ mult[b])) <= 1e-7 * max(1.0, np.max(np.abs(q)))


@pytest.mark.parametrize("name,N", [("quadrotor", 30), ("rc_car", 60), ("quadruped", 10)])
def test_qp_on_an_f32_handle_factorises_in_fp64(name, N):
    """BASELINE configs 2 and 3 are fp32.  An F32 handle sweeps in fp32 and hands its fp32 record to the same fp64 factorisation
    (widened on the device, on a twin F64 handle): the step equals the F64 handle's solve of the widened record up to the fp32
    rounding of the output, and it solves the QP the fp32 record states."""
    import torch

    import ungar_b200
    from ungar_b200 import EXAMPLE_BARRIER
    from ungar_b200 import workloads as W

    mid = W.MODEL_IDS[name]
    m32 = ungar_b200.Model(name, N, dtype="f32", barrier=EXAMPLE_BARRIER[mid])
    m64 = ungar_b200.Model(name, N, dtype="f64", barrier=EXAMPLE_BARRIER[mid])
    xp = torch.from_numpy(W.synthetic_batch(mid, N, 6, seed=3)).cuda()
    rec32 = m32.kkt_blocks(xp.float())
    assert rec32.dtype == torch.float32
    steps32, mult32 = m32.qp_solve(rec32)
    steps64, mult64 = m64.qp_solve(rec32.double())
    assert steps32.dtype == torch.float32 and torch.isfinite(steps32).all()
    scale = steps64.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
    assert float(((steps32.double() - steps64).abs() / scale).max()) <= 2e-7        # one fp32 rounding of the output
    mscale = mult64.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
    assert float(((mult32.double() - mult64).abs() / mscale).max()) <= 2e-7
