"""Throughput of the generated batched ABA kernel (ANYmal B, nq 19 / nv 18): accelerations and their Jacobian for
batches of states resident in HBM, CUDA events on the launching stream; the numpy oracle (CRBA + RNEA + dense solve) on one host core
beside it.  Run on a GPU box:  python profiles/rbd_timing.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ungar_b200 import autodiff as A  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "rbd_anymal_b.npz"))
f = A.TapeFunction.load(os.path.join(ROOT, "tests", "golden", "rbd_anymal_b_tape.npz"))
print("tape:", f.tape_info())


def timed(fn, reps):
    fn()
    fn()  # the second call of an order compiles / loads the specialised kernel
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(3)
for B in (1024, 16384, 131072):
    X = torch.from_numpy(z["X"][rng.integers(0, z["X"].shape[0], B)]).cuda()
    ms = timed(lambda: f._tape.forward_zero(X), 20)
    print(f"ABA values    batch {B:7d}: {ms:8.3f} ms  {B / ms * 1e3 / 1e6:8.2f} M states/s   kernel state {f._tape.special_info()[0]['state']}")
    if B <= 16384:
        ms = timed(lambda: f._tape.sparse_jacobian(X), 10)
        print(f"ABA Jacobian  batch {B:7d}: {ms:8.3f} ms  {B / ms * 1e3 / 1e6:8.2f} M states/s   kernel state {f._tape.special_info()[1]['state']}")

try:
    from oracle import rbd_reference as R  # noqa: E402

    urdf = "/root/reference/data/robots/anymal_b_description/robots/anymal.urdf"
    if os.path.exists(urdf):
        tree = R.Tree(urdf)
        x = z["X"][0]
        t0 = time.perf_counter()
        for _ in range(50):
            R.forward_dynamics(tree, x[:19], x[19:37], x[37:])
        print(f"numpy oracle, one core: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per state")
except Exception as e:  # the oracle needs the reference's URDF, which does not travel
    print("oracle timing skipped:", e)
