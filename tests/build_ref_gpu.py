#!/usr/bin/env python
"""Builds tests/_ref_gpu/: the reference's UNCHANGED sources compiled against the PRODUCT's CppAD-compatible header
(ungar_b200/include/cppad/cg.hpp -> C ABI -> register-machine kernels), so the reference's own Function class and its own example
evaluate their lambdas on the GPU.

  function_tests_gpu     the known answers of test/autodiff/function.test.cpp:33-142 (GoogleTest is absent: the bodies are restated
                         with plain checks in oracle/ref_drivers/function_tests.cpp) through the reference's include/ungar/autodiff/function.hpp
  function_example_gpu   example/autodiff/function.example.cpp compiled as it lies: VariableMap + MakeFunction + TestJacobian /
                         TestHessian (AD vs finite differences) with UNGAR_ASSERT active
  soft_sqp_tests_gpu     the reference's unchanged SoftSQPOptimizer on the three NLPs of test/optimization/soft_sqp.test.cpp:34-111
                         (tests/ref_drivers/soft_sqp_tests.cpp): functions AND local QPs on the device (osqp++.h stand-in)
  example_<name>_gpu     example/mpc/{quadrotor,rc_car,quadruped}.example.cpp compiled as they lie (tests/ref_drivers/example_driver.cpp bounds
                         their endless MPC loop through UNGAR_B200_MAX_QP_SOLVES)

Needs /root/reference (absent on the GPU box: the prebuilt binaries travel with the snapshot; tests/_ref_gpu is git-ignored).
Nothing of the reference is copied into the repository.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref_gpu")
DEPS = os.path.join(ROOT, "oracle", "_ref", "deps")  # the zips the reference bundles, unpacked by oracle/build_ref.py
CXX = os.environ.get("CXX", "g++")


def main(force: bool = False) -> bool:
    if not os.path.isdir(REF):
        print("build_ref_gpu: /root/reference is absent: using the prebuilt tests/_ref_gpu if any")
        return False
    sys.path.insert(0, ROOT)
    from ungar_b200 import build

    lib = build.build()
    os.makedirs(DEPS, exist_ok=True)
    for z in ("eigen/eigen-3.4.0.zip", "hana/hana-boost-1.84.0.zip", "preprocessor/preprocessor-1.84.0-ungar.zip"):
        marker = os.path.join(DEPS, os.path.basename(z) + ".done")
        if not os.path.exists(marker):
            zipfile.ZipFile(os.path.join(REF, "external/config", z)).extractall(DEPS)
            open(marker, "w").close()
    spdlog = glob.glob("/opt/prime-rl/.venv/lib/python3*/site-packages/flashinfer/data/spdlog/include")
    os.makedirs(os.path.join(OUT, "tapes"), exist_ok=True)
    flags = ["-std=c++20", "-O1", "-DUNGAR_CONFIG_ENABLE_AUTODIFF", "-DUNGAR_CONFIG_ENABLE_OPTIMIZATION", "-DFMT_HEADER_ONLY", f'-DUNGAR_CODEGEN_FOLDER="{OUT}/tapes"',
             "-ftemplate-backtrace-limit=1", "-fconstexpr-depth=2147483647", "-fconstexpr-loop-limit=2147483647",
             "-fconstexpr-cache-depth=2147483647", "-fconstexpr-ops-limit=2147483647",
             f"-I{ROOT}/ungar_b200/include", f"-I{ROOT}/include", f"-I{REF}/include", f"-I{DEPS}/eigen-3.4.0",
             f"-I{DEPS}/hana-boost-1.84.0/include", f"-I{DEPS}/preprocessor-1.84.0-ungar/include"]
    if spdlog:
        flags += ["-DUNGAR_CONFIG_ENABLE_LOGGING", f"-I{spdlog[0]}"]
    link = [lib, f"-Wl,-rpath,{os.path.dirname(lib)}", "-Wl,-rpath,$ORIGIN/../../ungar_b200"]
    headers = [os.path.join(ROOT, "ungar_b200/include", h) for h in ("cppad/cg.hpp", "osqp++.h")] + [os.path.join(ROOT, "include/ungar_b200.h")]
    driver = os.path.join(HERE, "ref_drivers/example_driver.cpp")
    targets = [("function_tests_gpu", os.path.join(ROOT, "oracle/ref_drivers/function_tests.cpp"), []),
               ("function_example_gpu", os.path.join(REF, "example/autodiff/function.example.cpp"), []),
               ("soft_sqp_tests_gpu", os.path.join(HERE, "ref_drivers/soft_sqp_tests.cpp"), [])]
    for tag, coeff in (("a", "2.0"), ("b", "3.0")):  # same function name, same folder, different lambda (tests/test_tape_host.py)
        targets.append((f"stale_tape_{tag}", os.path.join(HERE, "ref_drivers/stale_tape_driver.cpp"), [f"-DCOEFF={coeff}"]))
    for example in ("quadrotor", "rc_car", "quadruped"):
        targets.append((f"example_{example}_gpu", driver, [f'-DUNGAR_EXAMPLE_SOURCE="{REF}/example/mpc/{example}.example.cpp"']))
    for name, src, extra in targets:
        exe = os.path.join(OUT, name)
        if not force and os.path.exists(exe) and os.path.getmtime(exe) >= max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in headers]):
            continue
        cmd = [CXX] + flags + extra + ["-o", exe, src] + link
        print("+", " ".join(cmd[:4]), "...", src, flush=True)
        subprocess.run(cmd, check=True)
    print("build_ref_gpu: ok")
    return True


if __name__ == "__main__":
    main(force="--force" in sys.argv)
