#!/usr/bin/env python
"""ORACLE — TEST INFRASTRUCTURE ONLY.  Writes tests/golden/sqp_<model>_N<N>.npz: seeded inputs and the iterates / statuses / line-search
records of SoftSQPOptimizer::Optimize restated in oracle/sqp_reference.py (soft_sqp.hpp:63-109, backtracking_line_search.hpp:81-165),
so that the device loop is also compared with committed vectors.  Re-run after any change of the oracle:  python oracle/make_golden_sqp.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import sqp_reference as S  # noqa: E402
from ungar_b200 import EXAMPLE_BARRIER  # noqa: E402
from ungar_b200 import workloads as W  # noqa: E402

CASES = [("quadruped", 10, 5, 4, 23), ("quadrotor", 30, 6, 4, 23), ("rc_car", 30, 6, 4, 23)]  # model, N, iterations, batch, seed


def main():
    orc = oracle.Oracle()
    for name, N, iters, B, seed in CASES:
        mid = W.MODEL_IDS[name]
        k, eps = EXAMPLE_BARRIER[mid]
        mult = 1.0 if mid == W.QUADROTOR else 1.0 / N
        xp = W.synthetic_batch(mid, N, B, seed=seed)
        finals, status, alphas = [], [], []
        for b in range(B):
            x, st, it, log = S.soft_sqp(orc, mid, N, xp[b], k, eps, mult, iters)
            finals.append(x)
            status.append((st, it))
            alphas.append([l["ls"].alpha for l in log] + [-1.0] * (iters - len(log)))
        path = os.path.join(ROOT, "tests", "golden", f"sqp_{name}_N{N}.npz")
        np.savez_compressed(path, xp=xp, final=np.stack(finals), status=np.array(status, dtype=np.int32), alphas=np.array(alphas),
                            iterations=iters, multiplier=mult, stiffness=k, epsilon=eps, seed=seed)
        print(path, np.array(status).tolist())


if __name__ == "__main__":
    main()
