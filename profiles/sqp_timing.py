"""Times one soft-SQP iteration on the device (quadruped N = 100, 1024 trajectories, fp64): KKT sweep, QP solve, line search,
and the whole ungar_b200_sqp_solve loop (quadruped.example.cpp:444: 4 iterations, multiplier dt)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ungar_b200
from ungar_b200 import workloads as W

name = sys.argv[1] if len(sys.argv) > 1 else "quadruped"
N, B = {"quadruped": (100, 1024), "quadrotor": (30, 4096), "rc_car": (60, 8192)}[name]
mid = W.MODEL_IDS[name]
# the quadruped's consumers work on the compact record (csrc/compact.cuh), which is what ungar_b200_sqp_solve uses internally
m = ungar_b200.Model(name, N, dtype="f64", barrier=ungar_b200.EXAMPLE_BARRIER[mid], record_format="compact" if name == "quadruped" else "dense")
xp0 = torch.from_numpy(W.synthetic_batch(mid, N, B)).cuda()
xp = xp0.clone()
rec = m.kkt_blocks(xp)
steps, _ = m.qp_solve(rec, want_multipliers=False)
opts = m.sqp_options(max_iterations=4, constraint_violation_multiplier=1.0 if name == "quadrotor" else 1.0 / N)
torch.cuda.synchronize()


def timed(fn, reps=10, setup=None):
    tot = 0.0
    for _ in range(reps):
        if setup:
            setup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


print(name)
print("kkt_blocks  %d x N=%d: %.3f ms" % (B, N, timed(lambda: m.kkt_blocks(xp0, rec))))
print("qp_solve    %d x N=%d: %.3f ms" % (B, N, timed(lambda: m.qp_solve(rec, steps, want_multipliers=False))))
info = m.line_search(xp, steps, opts)
print("line_search %d x N=%d: %.3f ms (mean trials %.2f)" % (B, N, timed(lambda: m.line_search(xp, steps, opts), setup=lambda: xp.copy_(xp0)),
                                                          float((-torch.log2(info[:, 0].clamp_min(2.0 ** -14))).mean()) + 1))
m.sqp_solve(xp.copy_(xp0), opts)  # warm-up: the first call allocates the loop's workspaces (0.5 GB of compact records) and loads its kernels
t = timed(lambda: m.sqp_solve(xp, opts), setup=lambda: xp.copy_(xp0))
st, _ = m.sqp_solve(xp.copy_(xp0), opts)
print("sqp_solve   %d x N=%d, 4 iterations: %.3f ms  (status counts %s)" % (B, N, t, torch.bincount(st[:, 0], minlength=3).tolist()))
