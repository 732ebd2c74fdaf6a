#!/usr/bin/env python
"""Tracked summaries of profiles/run_sqp_profile.sh: <tag>_sqp_launches.md (per-kernel shares of one SQP timing run under ncu),
<tag>_sqp_timing.txt (CUDA-event timings, not under a profiler) and <tag>_ncu_sqp.json (full-set metrics of the QP and
line-search kernels)."""
import csv
import io
import json
import os
import shutil
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ncu_summary  # noqa: E402


def main(tag):
    out = os.path.join(ROOT, "gpurun_out")
    shutil.copy(os.path.join(out, f"sqp_timing_{tag}.log"), os.path.join(HERE, f"{tag}_sqp_timing.txt"))
    rows = [r for r in csv.reader(open(os.path.join(out, f"sqp_launches_{tag}.csv"))) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += float(r[vi].replace(",", ""))
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(HERE, f"{tag}_sqp_launches.md"), "w") as f:
        f.write(f"# ncu launch list of the SQP components, tag {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 "
                "python profiles/sqp_timing.py quadruped` (quadruped N = 100, 1024 trajectories, fp64; cold-cache, serialised launches: "
                "compare shares, not absolutes; torch's own copy / fill kernels are part of the script, not of the library)\n\n"
                "| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / 1e3 / v[0]:.1f} | {v[1] / total:.3f} |\n")
    summ = []
    for rep in (f"sqp_{tag}.ncu-rep", f"ls_{tag}.ncu-rep", f"riccati_{tag}.ncu-rep"):
        path = os.path.join(out, rep)
        if not os.path.exists(path):
            continue
        buf = io.StringIO()
        stdout, sys.stdout = sys.stdout, buf
        ncu_summary.main(path)
        sys.stdout = stdout
        summ += json.loads(buf.getvalue())
    json.dump(summ, open(os.path.join(HERE, f"{tag}_ncu_sqp.json"), "w"), indent=1)
    print("kernels summarised:", [k["kernel"][:40] for k in summ])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
