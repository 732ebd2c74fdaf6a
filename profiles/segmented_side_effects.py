import os, sys, time, ctypes, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import ungar_b200
from ungar_b200 import autodiff as A, workloads as W
from test_gpu_tape import load_reference_tape, tape_path
rt = ctypes.CDLL("libcudart.so.12")
def stack():
    v = ctypes.c_size_t(); rt.cudaDeviceGetLimit(ctypes.byref(v), 0); return v.value
def handle(cfg, fn, N, mid):
    nodes, ni, dep_id, dep_const = load_reference_tape(tape_path(cfg, fn))
    t = A.TapeHandle(nodes, ni, dep_id, dep_const)
    nx = W.sizes(mid, N)["n_dec"]; r, c = t.jacobian_pattern(); keep = c < nx; t.set_jacobian_elements(r[keep], c[keep]); return t
def bench(label, fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{label}: gpu {e0.elapsed_time(e1)/reps:.3f} ms  wall {(t1-t0)*1e3/reps:.3f} ms  stack limit {stack()}", flush=True)
xq = torch.from_numpy(W.synthetic_batch(W.QUADROTOR, 30, 1024, seed=3)).cuda()
tq = handle("quadrotor_N30", "quadrotor_mpc_eqs", 30, W.QUADROTOR)
mq = ungar_b200.Model("quadrotor", 30, dtype="f64")
bench("before: quadrotor interpreter", lambda: tq.sparse_jacobian(xq))
bench("before: hand-written        ", lambda: mq.equalityConstraints.JacobianValues(xq))
tp = handle("quadruped_N30", "quadruped_mpc_eqs", 30, W.QUADRUPED)
xp = torch.from_numpy(W.synthetic_batch(W.QUADRUPED, 30, 8, seed=3)).cuda()
tp.sparse_jacobian(xp); tp.sparse_jacobian(xp); print(tp.wait_specialised()[1], flush=True)
bench("segmented quadruped N=30    ", lambda: tp.sparse_jacobian(xp))
tq2 = handle("quadrotor_N30", "quadrotor_mpc_eqs", 30, W.QUADROTOR)
os.environ["UNGAR_B200_NO_NVRTC"] = "1"
bench("after: quadrotor interpreter (new handle)", lambda: tq2.sparse_jacobian(xq))
bench("after: quadrotor interpreter (old handle)", lambda: tq.sparse_jacobian(xq))
bench("after: hand-written        ", lambda: mq.equalityConstraints.JacobianValues(xq))
del tp; import gc; gc.collect()
bench("after del: quadrotor interpreter", lambda: tq2.sparse_jacobian(xq))
