"""Run in a subprocess by tests/test_gpu_variants.py with kernel-selection environment variables set: checks every model's
KKT record against the oracle on a small batch.  Prints 'OK <max err>' per configuration; exits non-zero on failure."""
import sys

import numpy as np

import oracle
import ungar_b200
from ungar_b200 import workloads as W

orc = oracle.Oracle()
for name, N, dtype, rtol in (("quadrotor", 30, "f32", 1e-3), ("quadrotor", 17, "f64", 1e-6), ("rc_car", 60, "f32", 1e-3),
                             ("rc_car", 31, "f64", 1e-6), ("quadruped", 30, "f64", 1e-6), ("quadruped", 100, "f64", 1e-6)):
    mid = W.MODEL_IDS[name]
    k, eps = ungar_b200.EXAMPLE_BARRIER[mid]
    model = ungar_b200.Model(name, N, dtype=dtype, barrier=(k, eps))
    xp = W.synthetic_batch(mid, N, 9, seed=31, perturb_params=True).astype(model.np_dtype)
    ref = model.split_record(orc.stage_sweep(mid, N, xp.astype(np.float64), k, eps))
    got = model.split_record(model.kkt_blocks(xp).astype(np.float64))
    worst = 0.0
    nX = model.layout["nx"] * (N + 1)
    for key, b in ref.items():
        scale = float(np.max(np.abs(xp[:, :nX]))) if key == "g" else float(np.max(np.abs(b)))
        err = float(np.max(np.abs(got[key] - b) / (np.abs(b) + 1e-3 * scale + 1e-300)))
        worst = max(worst, err)
        if not err <= rtol:
            print(f"FAIL {name} N={N} {dtype} block {key}: {err:.3e}")
            sys.exit(1)
    print(f"OK {name} N={N} {dtype} {worst:.2e}")
