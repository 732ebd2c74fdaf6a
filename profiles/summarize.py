#!/usr/bin/env python
"""Turns the scratch artefacts of profiles/run_profile.sh (gpurun_out/*_<tag>.*) into the tracked summaries:
profiles/<tag>_bench.json, <tag>_launches.md, <tag>_ncu_sweep.json and the traffic entry bench.py reports."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ncu_summary  # noqa: E402


def main(tag):
    out = os.path.join(ROOT, "gpurun_out")
    bench = json.load(open(os.path.join(out, f"bench_{tag}.json")))
    json.dump(bench, open(os.path.join(HERE, f"{tag}_bench.json"), "w"), indent=1)
    # launch list: per-kernel totals and shares (cold-cache, serialised: compare SHARES)
    rows = [r for r in csv.reader(open(os.path.join(out, f"launches_{tag}.csv"))) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += float(r[vi].replace(",", ""))
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(HERE, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list, tag {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python bench.py "
                f"--steps 20 --warmup 3 --no-cpu-baseline` (cold-cache, serialised launches: compare shares, not absolutes; the capture "
                f"spans the device-resident leg, one launch per step, and the host-buffer leg, whose steps launch one sweep per H2D chunk)\n\n"
                "| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / 1e3 / v[0]:.1f} | {v[1] / total:.3f} |\n")
        share = bench["roofline"]["kernel_share_of_step"]
        f.write(f"\nLive share of the sweep kernel inside bench.py's timed region (CUDA events): {share:.3f}\n")
    # full capture of the sweep kernel
    rep = os.path.join(out, f"sweep_{tag}.ncu-rep")
    buf = io.StringIO()
    stdout, sys.stdout = sys.stdout, buf
    ncu_summary.main(rep)
    sys.stdout = stdout
    summ = json.loads(buf.getvalue())
    json.dump(summ, open(os.path.join(HERE, f"{tag}_ncu_sweep.json"), "w"), indent=1)

    def num(s):
        v, unit = s.split()
        return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]

    traffic = sum(num(k["dram__bytes_read.sum"]) + num(k["dram__bytes_write.sum"]) for k in summ) / len(summ)
    cfg = bench["config"]
    key = f"{cfg['model_problem']}_{bench['dtype']}_N{cfg['horizon']}_B{cfg['batch_per_gpu']}"
    path = os.path.join(HERE, "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[key] = traffic
    table[key + "_source"] = f"profiles/{tag}_ncu_sweep.json (dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(summ)} launches)"
    json.dump(table, open(path, "w"), indent=1)
    print(key, traffic)


if __name__ == "__main__":
    main(sys.argv[1])
