"""The C++ host mirror (ungar_b200/include/ungar_b200/function.hpp) and the example built on it compile and link against
the C-ABI library; without a GPU the program fails loudly (no CPU fallback), with one it reproduces the stance known
answers of quadruped.example.cpp."""
import os
import subprocess

import pytest

from ungar_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compile_example(tmp_path, name="kkt_sweep"):
    lib = build.build()
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror", f"-I{ROOT}/include", f"-I{ROOT}/ungar_b200/include", "-o", exe,
           f"{ROOT}/examples/{name}.cpp", lib, f"-Wl,-rpath,{os.path.dirname(lib)}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="behaviour without a GPU")
def test_cpp_example_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    proc = subprocess.run([compile_example(tmp_path)], capture_output=True, text=True)
    assert proc.returncode == 2 and "no CPU fallback" in proc.stderr


@pytest.mark.gpu
def test_cpp_example_reproduces_the_stance_known_answers(tmp_path):
    proc = subprocess.run([compile_example(tmp_path)], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "nnz(J_g)=14167" in proc.stdout and "foot row = 0.38" in proc.stdout
    assert "soft SQP: status" in proc.stdout


@pytest.mark.skipif(_has_gpu(), reason="behaviour without a GPU")
def test_generic_function_example_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    """examples/generic_function.cpp: the CppAD-compatible tracing scalar + GenericModel of ungar_b200/include/cppad/cg.hpp used on
    their own (no Eigen, no reference headers)."""
    proc = subprocess.run([compile_example(tmp_path, "generic_function")], capture_output=True, text=True)
    assert proc.returncode == 2 and "no CPU fallback" in proc.stderr


@pytest.mark.gpu
def test_generic_function_example_matches_the_closed_form(tmp_path):
    proc = subprocess.run([compile_example(tmp_path, "generic_function")], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "max error vs closed form" in proc.stdout
