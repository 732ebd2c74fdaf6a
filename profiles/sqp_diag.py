import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import ungar_b200
from ungar_b200 import workloads as W
N,B=100,1024
m = ungar_b200.Model("quadruped", N, dtype="f64", barrier=ungar_b200.EXAMPLE_BARRIER[W.QUADRUPED], record_format="compact")
xp0 = torch.from_numpy(W.synthetic_batch(W.QUADRUPED, N, B)).cuda(); xp = xp0.clone()
opts = m.sqp_options(max_iterations=4, constraint_violation_multiplier=1.0/N)
m.sqp_solve(xp, opts); torch.cuda.synchronize()
for rep in range(6):
    xp.copy_(xp0); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record(); m.sqp_solve(xp, opts); e1.record(); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print("rep",rep,"gpu %.3f ms"%e0.elapsed_time(e1),"enqueue %.3f ms"%((t1-t0)*1e3),"total wall %.3f ms"%((t2-t0)*1e3))
