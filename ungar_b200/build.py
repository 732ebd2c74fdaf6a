"""In-tree build of the CUDA library (sm_100a only).  `python -m ungar_b200.build` or build.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libungar_b200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("ungar_b200.cu", "tape.cu", "kkt_dense.cu")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in ("dual.cuh", "models.cuh", "sweep.cuh", "sweep_structured.cuh", "sweep_structured_compact.cuh", "sweep_tpn.cuh", "sweep_small.cuh", "qp_schur.cuh", "qp_twisted.cuh", "compact.cuh", "qp_riccati.cuh", "line_search.cuh", "tape_machine.cuh", "abi_internal.h")] + [
    os.path.join(ROOT, "include", "ungar_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--expt-relaxed-constexpr", "-diag-suppress", "20011,20013,20014,20015",
    "-Xcompiler", "-fPIC", "-shared",
    # cuSOLVER's dense LU backs the generic QP fallback (csrc/kkt_dense.cu); everything else is hand-written
    "-lcusolver", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
    # the generic tape path specialises hot tapes with NVRTC and loads the cubin through the driver API (csrc/tape.cu); both are
    # bound with dlopen at run time so that the library still loads on hosts without a GPU driver
    "-ldl",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the ungar_b200 CUDA library cannot be built")


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in SOURCES + HEADERS if os.path.exists(f))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ungar_b200/libungar_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
