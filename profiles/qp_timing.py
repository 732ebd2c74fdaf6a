"""Times the batched QP solve next to the KKT sweep that feeds it (quadruped N = 100, 1024 trajectories, fp64)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ungar_b200
from ungar_b200 import workloads as W

m = ungar_b200.Model("quadruped", 100, dtype="f64", barrier=(1.0, 1.0))
xp = torch.from_numpy(W.synthetic_batch(2, 100, 1024)).cuda()
rec = m.kkt_blocks(xp)
steps, mult = m.qp_solve(rec)
torch.cuda.synchronize()


def timed(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("kkt_blocks 1024 x N=100: %.3f ms" % timed(lambda: m.kkt_blocks(xp, rec)))
print("qp_solve   1024 x N=100: %.3f ms" % timed(lambda: m.qp_solve(rec, steps, mult)))

mc = ungar_b200.Model("quadruped", 100, dtype="f64", barrier=(1.0, 1.0), record_format="compact")
recc = mc.kkt_blocks(xp)
sc, mu = mc.qp_solve(recc)
torch.cuda.synchronize()
print("compact kkt_blocks 1024 x N=100: %.3f ms" % timed(lambda: mc.kkt_blocks(xp, recc)))
print("compact qp_solve   1024 x N=100: %.3f ms" % timed(lambda: mc.qp_solve(recc, sc, mu)))
