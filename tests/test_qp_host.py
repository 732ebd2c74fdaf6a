"""The arithmetic of the QP kernel (ungar_b200/csrc/qp_twisted.cuh) without a GPU: the header's device functions are compiled for the
CPU (tests/host/*.cpp define away the CUDA qualifiers) and run (a) per lane against dense products built independently from the same
compact chunk, (b) as a whole warp — 32 host threads in lockstep, shuffles and __syncwarp through a std::barrier — against plain dense
Cholesky / triangular-solve loops.  This is host-logic coverage of the kernel's building blocks; the kernel itself is checked against
the sparse-LU oracle on the device in tests/test_gpu_qp.py."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["qp_rows_host", "qp_warp_host"])
def test_qp_kernel_building_blocks_on_the_host(tmp_path, name):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = tmp_path / name
    build = subprocess.run([gxx, "-std=c++20", "-O1", "-w", "-pthread", "-I", os.path.join(ROOT, "ungar_b200", "csrc"), "-o", str(exe),
                            os.path.join(ROOT, "tests", "host", name + ".cpp")], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout + build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
