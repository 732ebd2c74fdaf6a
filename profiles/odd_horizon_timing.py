import os, sys, torch
sys.path.insert(0, os.getcwd())
import ungar_b200
from ungar_b200 import workloads as W
for N in (100, 101):
    m = ungar_b200.Model("quadruped", N, dtype="f64", barrier=(1.0, 1.0))
    xp = torch.from_numpy(W.synthetic_batch(W.QUADRUPED, N, 1024)).cuda()
    rec = m.kkt_blocks(xp); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m.kkt_blocks(xp, rec)
    e1.record(); torch.cuda.synchronize()
    print("dense record, N =", N, "%.3f ms per 1024 trajectories" % (e0.elapsed_time(e1)/20))
