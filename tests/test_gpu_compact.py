"""COMPACT KKT records (include/ungar_b200.h ungar_b200_record_format, csrc/compact.cuh): only the structurally non-zero slots of the
quadruped's blocks (quadruped.example.cpp:162-200, :216-244, :269-303), one chunk per shooting node.  Parity: dense-from-compact
against the oracle with the strict metric, bit-identity with the dense sweep, and every consumer (QP solve, SQP loop, summaries,
host-buffer paths) on the compact format."""
import numpy as np
import pytest

from ungar_b200 import EXAMPLE_BARRIER, parity
from ungar_b200 import workloads as W

pytestmark = pytest.mark.gpu
K, EPS = EXAMPLE_BARRIER[W.QUADRUPED]


def models(N):
    import ungar_b200

    return (ungar_b200.Model("quadruped", N, dtype="f64", barrier=(K, EPS)),
            ungar_b200.Model("quadruped", N, dtype="f64", barrier=(K, EPS), record_format="compact"))


@pytest.mark.parametrize("N,B", [(100, 33), (30, 9), (7, 5), (2, 3), (101, 4)])
def test_compact_records_are_the_dense_records_without_their_zeros(oracle, N, B):
    import torch

    dense, comp = models(N)
    assert comp.layout["compact"] == 1 and comp.layout["size"] == N * comp.layout["node_stride"] + 44 < dense.layout["size"]
    assert comp.layout["dense_size"] == dense.layout["size"]
    xp = W.synthetic_batch(W.QUADRUPED, N, B, seed=5, perturb_params=True)
    d_xp = torch.from_numpy(xp).cuda()
    rd = dense.kkt_blocks(d_xp, torch.zeros((B, dense.layout["size"]), dtype=torch.float64, device="cuda"))
    rc = comp.kkt_blocks(d_xp, torch.full((B, comp.layout["size"]), float("nan"), dtype=torch.float64, device="cuda"))
    torch.cuda.synchronize()
    rd, rc = rd.cpu().numpy(), rc.cpu().numpy()
    assert np.isfinite(rc).all()  # every slot of a compact record is written (pads as zeros)
    m = comp.compact_map()
    assert m.size == comp.layout["size"] and np.all(rc[:, m < 0] == 0.0)
    assert np.unique(m[m >= 0]).size == (m >= 0).sum()  # no dense slot is held twice
    # bit-identical to the dense sweep on every slot it holds (same arithmetic; odd horizons: the dense handle falls back to the generic
    # dual-number sweep, which agrees to rounding) ...
    if N % 2 == 0:
        assert np.array_equal(rc[:, m >= 0], rd[:, m[m >= 0]])
    else:
        assert np.allclose(rc[:, m >= 0], rd[:, m[m >= 0]], rtol=1e-11, atol=1e-12)
    # ... and the slots it does not hold are structural zeros of the ORACLE's record (the cut loses nothing)
    ref = oracle.stage_sweep(W.QUADRUPED, N, xp, K, EPS)
    held = np.zeros(dense.layout["size"], dtype=bool)
    held[m[m >= 0]] = True
    nonzero_outside = np.abs(ref[:, ~held]).max()
    assert nonzero_outside == 0.0, nonzero_outside
    # dense-from-compact against the oracle, strict metric
    rep = parity.compare_records(dense.split_record, comp.to_dense(rc), ref, xp, 13 * (N + 1), "f64")
    assert rep["ok"], rep
    # host buffers through the same entry point
    assert np.array_equal(comp.kkt_blocks(xp), rc)
    # summaries from the compact record == from the dense one
    sc = comp.summaries(d_xp, torch.from_numpy(rc).cuda()).cpu().numpy()
    sd = dense.summaries(d_xp, torch.from_numpy(rd).cuda()).cpu().numpy()
    assert np.array_equal(sc, sd) if N % 2 == 0 else np.allclose(sc, sd, rtol=1e-11, atol=1e-12)
    # reference-format calls of a compact handle are served from a dense internal record: same values as the dense handle
    assert np.array_equal(comp.equalityConstraints.JacobianValues(xp[0]), dense.equalityConstraints.JacobianValues(xp[0]))


@pytest.mark.parametrize("N", [100, 7, 2])
def test_qp_solve_and_sqp_loop_on_compact_records(N):
    import torch

    from oracle import qp_reference as Q

    dense, comp = models(N)
    B = 6
    xp = W.synthetic_batch(W.QUADRUPED, N, B, seed=9, perturb_params=True)
    d_xp = torch.from_numpy(xp).cuda()
    rd = dense.kkt_blocks(d_xp, torch.zeros((B, dense.layout["size"]), dtype=torch.float64, device="cuda"))
    rc = comp.kkt_blocks(d_xp)
    sd, md = dense.qp_solve(rd)
    sc, mc = comp.qp_solve(rc)
    torch.cuda.synchronize()
    if N % 2 == 0:
        assert torch.equal(sd, sc) and torch.equal(md, mc)  # the dense handle gathers into the same compact form first
    else:
        assert float((sd - sc).abs().max()) <= 1e-6 * float(sd.abs().max())  # the solve amplifies the rounding differences of its inputs
    d_ref, lam_ref = Q.kkt_solve(rd[0].cpu().numpy(), dict(dense.layout))
    assert np.max(np.abs(sc[0].cpu().numpy() - d_ref)) <= 1e-7 * np.max(np.abs(d_ref))
    # the SQP loop keeps its records compact internally for both handles: identical iterates
    opts = dense.sqp_options(max_iterations=3, constraint_violation_multiplier=1.0 / N)
    xa, xb = d_xp.clone(), d_xp.clone()
    sta, _ = dense.sqp_solve(xa, opts)
    stb, _ = comp.sqp_solve(xb, opts)
    torch.cuda.synchronize()
    assert torch.equal(xa, xb) and torch.equal(sta, stb)
    # kkt_step with host buffers on the compact handle
    rec_dev = torch.zeros((B, comp.layout["size"]), dtype=torch.float64, device="cuda")
    summ = comp.step(xp, records=rec_dev)
    assert torch.equal(rec_dev, rc) and np.array_equal(summ, comp.summaries(d_xp, rc).cpu().numpy())


def test_compact_is_rejected_where_it_does_not_exist():
    import ungar_b200
    from ungar_b200 import _lib

    with pytest.raises(_lib.UngarB200Error):
        ungar_b200.Model("quadrotor", 30, dtype="f64", record_format="compact")
    with pytest.raises(_lib.UngarB200Error):
        ungar_b200.Model("quadruped", 30, dtype="f32", barrier=(K, EPS), record_format="compact")
    dense, comp = models(4)
    with pytest.raises(_lib.UngarB200Error):
        dense.compact_map()
