#!/usr/bin/env python
"""SASS evidence: opcode histogram of every kernel in ungar_b200/libungar_b200.so (cuobjdump -sass), so that the statements in
DESIGN.md about TMA bulk copies (UBLKCP), cp.async (LDGSTS), mbarriers (SYNCS), FP64 tensor cores (DMMA) and FP64 FMAs (DFMA) can be
checked without rebuilding.  Writes profiles/sass_r02.txt.   python profiles/sass_histogram.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ungar_b200", "libungar_b200.so")
WATCH = ["UBLKCP", "UTMACMDFLUSH", "LDGSTS", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR", "ATOMG", "RED", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    lines = ["SASS opcode histogram of ungar_b200/libungar_b200.so (sm_100a), one block per kernel.",
             "columns: total instructions | watched opcode families (prefix match) | ten most frequent opcodes", ""]
    for name, cnt in kernels.items():
        total = sum(cnt.values())
        fam = {w: sum(v for k, v in cnt.items() if k.startswith(w)) for w in WATCH}
        fam = {k: v for k, v in fam.items() if v}
        short = demangle(name)
        short = re.sub(r"\(.*", "", short)[:110]
        lines.append(f"{short}")
        lines.append(f"  total {total}   " + "  ".join(f"{k}:{v}" for k, v in fam.items()))
        lines.append("  top: " + ", ".join(f"{k} {v}" for k, v in cnt.most_common(10)))
        lines.append("")
    path = os.path.join(ROOT, "profiles", "sass_r02.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines))
    print(path, len(kernels), "kernels")


if __name__ == "__main__":
    sys.exit(main())
