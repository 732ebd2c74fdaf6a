"""Throughput of the generic tape path (register-machine interpreter and NVRTC-specialised kernels) on the reference's own MPC lambdas (oracle/_ref tapes), next to the
hand-written kernels for the same functions: equality-constraint Jacobian in the reference's CSR value format."""
import glob
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ungar_b200  # noqa: E402
from test_gpu_tape import load_reference_tape  # noqa: E402
from ungar_b200 import autodiff as A  # noqa: E402
from ungar_b200 import workloads as W  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for name, N in (("quadrotor", 30), ("rc_car", 60), ("quadruped", 30), ("quadruped", 100)):
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "tapes", f"{name}_N{N}", f"{name}_mpc_eqs", "cppad_cg", "*_lib.so"))
    if not hits:
        continue
    mid = W.MODEL_IDS[name]
    nodes, ni, dep_id, dep_const = load_reference_tape(hits[0])
    t = A.TapeHandle(nodes, ni, dep_id, dep_const)
    nx = W.sizes(mid, N)["n_dec"]
    r, c = t.jacobian_pattern()
    keep = c < nx
    t.set_jacobian_elements(r[keep], c[keep])
    info = t.info()
    model = ungar_b200.Model(name, N, dtype="f64", barrier=ungar_b200.EXAMPLE_BARRIER[mid])
    print(f"{name} N={N}: {nodes.size} tape nodes -> {info['live_nodes']} live, {info['slots']} slots, {info['jacobian_colors']} colours, "
          f"nnz(J_g) = {int(keep.sum())}")
    os.environ["UNGAR_B200_NO_NVRTC"] = "1"   # a second handle that stays on the interpreter (the switch is read when a kernel would be compiled)
    ti = A.TapeHandle(nodes, ni, dep_id, dep_const)
    ti.set_jacobian_elements(r[keep], c[keep])
    x1 = torch.from_numpy(W.synthetic_batch(mid, N, 1, seed=3)).cuda()
    ti.sparse_jacobian(x1); ti.sparse_jacobian(x1)
    ti.wait_specialised()  # (a long tape's worker thread reads the switch when it starts compiling)
    del os.environ["UNGAR_B200_NO_NVRTC"]
    t0 = time.perf_counter()
    t.sparse_jacobian(x1); t.sparse_jacobian(x1)  # the second call of an order compiles (or loads from the cache) the specialised kernel(s)
    t.wait_specialised()                          # (long tapes compile on a worker thread while the interpreter serves)
    torch.cuda.synchronize()
    print(f"   NVRTC specialisation: state {t.special_info()[1]['state']}, {time.perf_counter() - t0:.1f} s including the compile"
          f" ({'from the cache' if t.special_info()[1]['from_cache'] else 'compiled'})")
    for B in (1, 64, 1024):
        xp = torch.from_numpy(W.synthetic_batch(mid, N, B, seed=3)).cuda()
        ms_i = timed(lambda: ti.sparse_jacobian(xp), reps=3)
        ms_t = timed(lambda: t.sparse_jacobian(xp), reps=3)
        ms_k = timed(lambda: model.equalityConstraints.JacobianValues(xp), reps=3)
        thread_instr = info["live_nodes"] * info["jacobian_colors"] * B
        print(f"   batch {B:5d}: interpreter {ms_i:8.3f} ms | specialised {ms_t:8.3f} ms ({thread_instr / ms_t / 1e6:8.1f} G thread-instr/s, {B * N / ms_t / 1e3:8.3f} M nodes/s)"
              f" | hand-written kernels {ms_k:8.3f} ms ({B * N / ms_k / 1e3:9.3f} M nodes/s)")
