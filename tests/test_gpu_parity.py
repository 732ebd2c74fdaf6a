"""Parity of the CUDA path (through the C ABI) against the CPU oracle — run on the B200 with `-m gpu`.

Tolerances (BASELINE.json north_star): 1e-6 relative in fp64, 1e-3 relative in fp32, measured per entry with the
STRICT metric of `ungar_b200/parity.py`:
    |gpu - oracle| / max(|oracle|, 1e-9) <= rtol
An entry that misses it must be a characterised cancellation: its absolute error within 64 units of roundoff of the magnitude of the
operands it is built from (for the defects g: the states, because g = x_{k+1} - f(x_k, u_k) cancels O(|x|) quantities; for every
other block: the block's own largest entry).  Structural sparsity (CSR index arrays) must match the oracle's exactly.
"""
import numpy as np
import pytest

from ungar_b200 import EXAMPLE_BARRIER, parity
from ungar_b200 import workloads as W

pytestmark = pytest.mark.gpu

OBJ, EQ, INEQ = 0, 1, 2
RTOL = {"f64": 1e-6, "f32": 1e-3}
CONFIGS = [  # BASELINE.json configs at oracle-friendly batch sizes, plus ragged horizons (tile tails)
    ("quadrotor", 30, "f32"), ("quadrotor", 30, "f64"), ("rc_car", 60, "f32"), ("rc_car", 60, "f64"),
    ("quadruped", 100, "f64"), ("quadruped", 30, "f32"), ("quadruped", 7, "f64"), ("quadrotor", 17, "f64"),
    ("rc_car", 2, "f64"), ("rc_car", 31, "f32"),
]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("the gpu-marked tests need a CUDA device (no CPU fallback exists)")
    return torch


def make_model(name, N, dtype):
    import ungar_b200

    return ungar_b200.Model(name, N, dtype=dtype, barrier=EXAMPLE_BARRIER[W.MODEL_IDS[name]])


def assert_close(name, got, ref, dtype, scale=None):
    """`dtype`: "f64" / "f32" (or, for legacy call sites, the matching rtol)."""
    if not isinstance(dtype, str):
        dtype = "f64" if dtype <= 1e-5 else "f32"
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    rep = parity.check_block(got, ref, dtype, scale)
    assert rep["ok"], (name, rep)
    return rep


def compare_records(model, got, ref, xp, dtype):
    if not isinstance(dtype, str):
        dtype = "f64" if dtype <= 1e-5 else "f32"
    nX = model.layout["nx"] * (model.layout["horizon"] + 1)
    rep = parity.compare_records(model.split_record, got, ref, xp, nX, dtype)
    assert rep["ok"], rep
    return rep


@pytest.mark.parametrize("name,N,dtype", CONFIGS)
def test_kkt_blocks_match_oracle(torch_cuda, oracle, name, N, dtype):
    torch = torch_cuda
    mid = W.MODEL_IDS[name]
    B = 5
    model = make_model(name, N, dtype)
    assert model.layout["size"] == oracle.record_layout(mid, N)["size"]
    for key in ("g", "A", "C", "h", "cost", "grad", "H", "HN", "Hc"):
        assert model.layout[key] == oracle.record_layout(mid, N)[key]
    xp64 = W.synthetic_batch(mid, N, B, seed=17, perturb_params=N % 2 == 1 or dtype == "f64")
    xp = xp64.astype(model.np_dtype)
    k, eps = EXAMPLE_BARRIER[mid]
    ref = oracle.stage_sweep(mid, N, xp.astype(np.float64), k, eps)  # oracle sees the rounded inputs
    mono = oracle.kkt_record(mid, N, xp[0].astype(np.float64), k, eps)
    assert np.max(np.abs(mono - ref[0]) / np.maximum(np.abs(mono), 1.0)) < 1e-12
    # device buffers, padded strides
    d_xp = torch.zeros((B, model.n_xp + 3), dtype=torch.float64 if dtype == "f64" else torch.float32, device="cuda")
    d_xp[:, :model.n_xp] = torch.from_numpy(xp).cuda()
    d_rec = torch.full((B, model.layout["size"] + 8), float("nan"), dtype=d_xp.dtype, device="cuda")
    out = model.kkt_blocks(d_xp[:, :model.n_xp], d_rec[:, :model.layout["size"]])
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.isnan(d_rec[:, model.layout["size"]:].cpu().numpy()).all()  # nothing written past the record
    compare_records(model, got, ref, xp, RTOL[dtype])
    # host buffers through the same entry point (H2D + kernels + D2H inside the call)
    got_host = model.split_record(model.kkt_blocks(xp))
    for key, blk in model.split_record(got).items():  # (padding between blocks is never written)
        assert np.array_equal(got_host[key], blk), key
    # summaries
    summ = model.summaries(d_xp[:, :model.n_xp], out).cpu().numpy()
    nu, nX = model.layout["nu"], model.layout["nx"] * (N + 1)
    assert np.array_equal(summ[:, :nu], xp[:, nX:nX + nu])
    blocks = model.split_record(got)
    assert np.allclose(summ[:, 24], blocks["cost"][:, 0]) and np.allclose(summ[:, 26], np.abs(blocks["g"]).max(axis=1))
    assert np.allclose(summ[:, 27], blocks["h"].max(axis=1))


@pytest.mark.parametrize("name,N,dtype", [("quadrotor", 30, "f64"), ("rc_car", 60, "f64"), ("quadruped", 30, "f64"),
                                          ("quadruped", 100, "f64"), ("quadrotor", 30, "f32"), ("rc_car", 5, "f32")])
def test_function_api_matches_oracle(torch_cuda, oracle, name, N, dtype):
    """Evaluate / Jacobian / Hessian of the three functions in the reference's own format (CSR, upper triangle)."""
    mid = W.MODEL_IDS[name]
    model = make_model(name, N, dtype)
    rtol = RTOL[dtype]
    xps = W.synthetic_batch(mid, N, 3, seed=23).astype(model.np_dtype)
    xp = xps[1]
    x64 = xp.astype(np.float64)
    state_scale = float(np.max(np.abs(xp[:model.layout["nx"] * (N + 1)])))
    for fn, F in ((OBJ, model.objective), (EQ, model.equalityConstraints), (INEQ, model.inequalityConstraints)):
        s = oracle.sizes(mid, N)
        assert F.IndependentVariableSize() == s["n_dec"] and F.ParameterSize() == s["n_par"]
        y_ref = oracle.evaluate(mid, fn, N, x64)
        assert F.DependentVariableSize() == y_ref.size
        assert_close(f"y{fn}", F(xp), y_ref, rtol, scale=state_scale if fn == EQ else None)
        rows, cols, vals = oracle.jacobian(mid, fn, N, x64)
        grows, gcols = F.JacobianSparsity()
        assert np.array_equal(grows, rows) and np.array_equal(gcols, cols), f"Jacobian pattern of function {fn}"
        assert_close(f"J{fn}", F.JacobianValues(xp), vals, rtol)
        J = F.Jacobian(xp)
        assert J.shape == (y_ref.size, s["n_dec"]) and J.nnz == vals.size
    rows, cols, vals = oracle.hessian(mid, N, x64)
    hrows, hcols = model.objective.HessianSparsity()
    assert np.array_equal(hrows, rows) and np.array_equal(hcols, cols)
    assert_close("Hf", model.objective.HessianValues(xp), vals, rtol)
    assert model.objective.ImplementsHessian() and not model.equalityConstraints.ImplementsHessian()
    # batched call == per-trajectory calls
    yb = model.equalityConstraints(xps)
    for b in range(3):
        assert np.array_equal(yb[b], model.equalityConstraints(xps[b]))
    # barrier function of soft_sqp.hpp:114-138
    k, eps = EXAMPLE_BARRIER[mid]
    z = oracle.evaluate(mid, INEQ, N, x64)
    z[::3] = -0.5 * eps  # exercise the cubic piece too
    val, dz, d2z = oracle.barrier(k, eps, z)
    S = model.softInequalityConstraints
    zt = z.astype(model.np_dtype)
    val, dz, d2z = oracle.barrier(k, eps, zt.astype(np.float64))
    assert_close("Z", S(zt), [val], rtol)
    assert_close("dZ", S.JacobianValues(zt), dz, rtol)
    assert_close("d2Z", S.HessianValues(zt), d2z, rtol)


def test_known_answers_on_device(torch_cuda):
    """The reference-derived known answers of SURVEY.md §8c, evaluated by the kernels."""
    m = make_model("quadrotor", 30, "f64")
    g = m.equalityConstraints(W.quadrotor_nominal(30, 0.0))
    assert np.abs(g).max() < 1e-14  # hover equilibrium
    m = make_model("quadruped", 30, "f64")
    g = m.equalityConstraints(W.quadruped_nominal(30, 0.0))
    assert np.abs(g[:13 + 13 * 30]).max() < 1e-14
    assert np.allclose(g[13 + 13 * 30:].reshape(30, 4, 4)[0], [[0, 0, 0, 0.38]] * 4)
    m = make_model("rc_car", 30, "f64")
    xp = W.rc_car_nominal(30, 0.0)
    g = m.equalityConstraints(xp)
    vx = 1.0 - (1.0 / 30.0) * (0.0518 + 0.00035) / 0.041
    assert abs((xp[6 + 3] - g[6 + 3]) - vx) < 1e-14  # x_1 - g = f(x_0, u_0)


def test_edge_cases(torch_cuda):
    import ungar_b200
    from ungar_b200 import _lib

    m = make_model("quadruped", 30, "f64")
    assert m.kkt_blocks(np.zeros((0, m.n_xp))).shape == (0, m.layout["size"])  # empty batch
    with pytest.raises(_lib.UngarB200Error):  # ld_rec too small
        m.kkt_blocks(np.zeros((1, m.n_xp)), np.zeros((1, 16)))
    with pytest.raises(_lib.UngarB200Error):  # Hessian of a vector function (function.hpp:136-137)
        m.equalityConstraints.HessianValues(np.zeros(m.n_xp))
    with pytest.raises(ValueError):
        m.objective(np.zeros(m.n_xp - 1))
    with pytest.raises(_lib.UngarB200Error):
        ungar_b200.Model("quadruped", 30, device=99)


@pytest.mark.parametrize("name,N,dtype,B", [("quadrotor", 30, "f32", 4096), ("rc_car", 60, "f32", 8192),
                                            ("quadruped", 100, "f64", 1024)])
def test_full_size_properties(torch_cuda, oracle, name, N, dtype, B):
    """BASELINE.json batch sizes: ALL trajectories against the oracle (strict metric) + size-independent properties."""
    torch = torch_cuda
    mid = W.MODEL_IDS[name]
    model = make_model(name, N, dtype)
    k, eps = EXAMPLE_BARRIER[mid]
    xp = W.synthetic_batch(mid, N, B, seed=101).astype(model.np_dtype)
    d_xp = torch.from_numpy(xp).cuda()
    zeros = lambda: torch.zeros((B, model.layout["size"]), dtype=d_xp.dtype, device="cuda")  # noqa: E731
    rec = model.kkt_blocks(d_xp, zeros())  # (the padding between blocks is never written: pre-zero it)
    torch.cuda.synchronize()
    got = rec.cpu().numpy()
    assert np.isfinite(got).all()
    # every trajectory of the batch against the oracle (the stage-wise port on all host threads; seconds at these sizes)
    import os

    ref = oracle.stage_sweep(mid, N, xp.astype(np.float64), k, eps, threads=os.cpu_count() or 1)
    rep = compare_records(model, got, ref, xp, dtype)
    print(f"\n[parity {name} N={N} {dtype} B={B}] strict max {rep['max_rel_err']:.3e} (tol {RTOL[dtype]:g}), cancellation entries "
          f"{rep['cancellation_entries']} of {rep['entries']} (worst {rep['cancellation_worst_ulps']:.1f} ulps of {rep['cancellation_bound_ulps']:.0f}), "
          f"legacy metric {rep['legacy_max_rel_err']:.3e}")
    # permutation equivariance: trajectories are independent work items
    perm = np.random.default_rng(0).permutation(B)
    rec2 = model.kkt_blocks(d_xp[torch.from_numpy(perm).cuda()], zeros())
    torch.cuda.synchronize()
    assert torch.equal(rec2, rec[torch.from_numpy(perm).cuda()])
    # idempotence / determinism
    assert torch.equal(model.kkt_blocks(d_xp, zeros()), rec)
    # checksum of checksums: cost = sum of what the per-function call reports
    f = model.objective(d_xp[:64])
    torch.cuda.synchronize()
    blocks = model.split_record(got[:64])
    assert_close("cost", f.cpu().numpy()[:, 0], blocks["cost"][:, 0], RTOL[dtype])
    # first-order consistency of A with g (linearity): g(x + dz) - g(x) ~ J dz for a tiny step in x_k, u_k
    blocks = model.split_record(got[:2])
    L = model.layout
    step = 1e-6 if dtype == "f64" else 1e-3
    dz = np.random.default_rng(1).standard_normal(L["n_dec"]) * step
    xq = xp[:2].astype(np.float64).copy()
    xq[:, :L["n_dec"]] += dz
    g1 = model.equalityConstraints(xq.astype(model.np_dtype)).astype(np.float64)
    J = model.equalityConstraints.Jacobian(xp[0])
    lin = J @ dz
    # second-order remainder (and fp32 rounding of g) relative to the size of the linear term
    assert np.max(np.abs(g1[0] - blocks["g"][0] - lin)) < (1e-4 if dtype == "f64" else 5e-2) * np.max(np.abs(lin))


@pytest.mark.parametrize("name,N,dtype,B", [("quadruped", 100, "f64", 1024), ("quadruped", 30, "f64", 301), ("rc_car", 60, "f32", 700),
                                            ("quadrotor", 30, "f64", 100)])
def test_host_step_pipeline_equals_device_step(torch_cuda, name, N, dtype, B):
    """ungar_b200_kkt_step with HOST buffers (chunked H2D overlapped with the sweep) returns the same summaries and leaves the
    same records in HBM as the device-pointer call; ragged batch sizes exercise the last, shorter chunk."""
    torch = torch_cuda
    mid = W.MODEL_IDS[name]
    model = make_model(name, N, dtype)
    xp = W.synthetic_batch(mid, N, B, seed=7).astype(model.np_dtype)
    padded = np.zeros((B, model.n_xp + 3), dtype=model.np_dtype)  # non-compact host stride
    padded[:, :model.n_xp] = xp
    tdt = torch.float64 if dtype == "f64" else torch.float32
    rec_h = torch.zeros((B, model.layout["size"]), dtype=tdt, device="cuda")
    rec_d = torch.zeros_like(rec_h)
    sum_d = model.step(torch.from_numpy(xp).cuda(), records=rec_d)
    torch.cuda.synchronize()
    for host in (xp, padded[:, :model.n_xp]):
        rec_h.zero_()
        sum_h = np.zeros((B, 32), dtype=model.np_dtype)
        if host is xp:
            model.step(host, records=rec_h, summaries=sum_h)
        else:  # strided rows straight through the ABI
            from ungar_b200 import _lib
            _lib.check(model._lib.ungar_b200_kkt_step(model._handle, padded.ctypes.data, B, padded.shape[1], rec_h.data_ptr(), rec_h.stride(0),
                                                      sum_h.ctypes.data, _lib.MEM_HOST, None))
        assert np.array_equal(sum_h, sum_d.cpu().numpy())
        assert torch.equal(rec_h, rec_d)


@pytest.mark.parametrize("name,N,dtype,B", [("quadrotor", 30, "f32", 257), ("quadrotor", 17, "f64", 9), ("rc_car", 60, "f32", 130),
                                            ("rc_car", 31, "f64", 5), ("quadruped", 100, "f64", 12), ("quadruped", 7, "f32", 3)])
def test_jacobian_blocks_equal_the_full_sweep(torch_cuda, name, N, dtype, B):
    """ungar_b200_jacobian_blocks (the Jacobian sweep of BASELINE configs[1]): g and A (quadruped: also C) are bit-identical to what
    ungar_b200_kkt_blocks writes; ragged horizons and batch sizes exercise the tails; the host-buffer path agrees with the device one."""
    torch = torch_cuda
    mid = W.MODEL_IDS[name]
    model = make_model(name, N, dtype)
    xp = W.synthetic_batch(mid, N, B, seed=19).astype(model.np_dtype)
    d_xp = torch.from_numpy(xp).cuda()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    full = model.kkt_blocks(d_xp, torch.zeros((B, model.layout["size"]), dtype=tdt, device="cuda"))
    jac = model.jacobian_blocks(d_xp, torch.full((B, model.layout["size"]), float("nan"), dtype=tdt, device="cuda"))
    torch.cuda.synchronize()
    fb, jb = model.split_record(full.cpu().numpy()), model.split_record(jac.cpu().numpy())
    for key in ("g", "A") + (("C",) if model.layout["legs"] else ()):
        assert np.array_equal(fb[key], jb[key]), key
    host = model.jacobian_blocks(xp)
    hb = model.split_record(host)
    assert np.array_equal(hb["g"], fb["g"]) and np.array_equal(hb["A"], fb["A"])
