#!/usr/bin/env python
"""ORACLE — TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/ from the reference sources WHERE THEY LIE under /root/reference.

Nothing of the reference is copied into the repository: outputs go only to oracle/_ref/ (git-ignored, but shipped to the
GPU box by gpurun).  The reference's own build system is not used (it needs CMake + network downloads).

  1. unpack the three zips the reference bundles (Eigen 3.4.0, Boost.Hana 1.84, Boost.Preprocessor 1.84;
     external/config/*/*.zip) into oracle/_ref/deps
  2. compile oracle/ref_drivers/function_tests.cpp: the reference's UNCHANGED include/ungar/autodiff/function.hpp
     driven over oracle/refshim (stand-in for the absent CppAD / CppADCodeGen / finite-diff) and run it — it must
     reproduce the known answers of test/autodiff/function.test.cpp
  3. for each MPC example: compile the UNCHANGED example source (or, for N != 30, a copy under oracle/_ref/src whose only
     edit is the `constexpr auto N = 30_c` line) with oracle/ref_drivers/example_driver.cpp and run it with
     UNGAR_REF_MAX_SOLVES=0: the reference's own MakeFunction tapes the example's own lambdas -> oracle/_ref/tapes
  4. compile oracle/ref_drivers/tape_tool.cpp -> oracle/_ref/libreftape.so (used by oracle/make_golden.py)
"""
from __future__ import annotations

import glob
import os
import re
import subprocess
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
DEPS = os.path.join(OUT, "deps")
CXX = os.environ.get("CXX", "g++")

EXAMPLES = [  # (example file, horizon) — BASELINE.json configs use N = 30 / 60 / 100
    ("quadrotor", 30), ("rc_car", 30), ("rc_car", 60), ("quadruped", 30), ("quadruped", 100),
]


def spdlog_include() -> str:
    hits = glob.glob("/opt/prime-rl/.venv/lib/python3*/site-packages/flashinfer/data/spdlog/include")
    if not hits:
        raise RuntimeError("header-only spdlog not found (the reference's logging.hpp needs it)")
    return hits[0]


def flags(tapes_dir: str) -> list[str]:
    return ["-std=c++20", "-O1", "-DUNGAR_CONFIG_ENABLE_AUTODIFF", "-DUNGAR_CONFIG_ENABLE_OPTIMIZATION",
            "-DUNGAR_CONFIG_ENABLE_LOGGING", "-DFMT_HEADER_ONLY", "-DUNGAR_CONFIG_ENABLE_RELEASE_MODE",
            f'-DUNGAR_CODEGEN_FOLDER="{tapes_dir}"', "-ftemplate-backtrace-limit=1", "-fconstexpr-depth=2147483647",
            "-fconstexpr-loop-limit=2147483647", "-fconstexpr-cache-depth=2147483647", "-fconstexpr-ops-limit=2147483647",
            f"-I{HERE}/refshim", f"-I{REF}/include", f"-I{DEPS}/eigen-3.4.0", f"-I{DEPS}/hana-boost-1.84.0/include",
            f"-I{DEPS}/preprocessor-1.84.0-ungar/include", f"-I{spdlog_include()}"]


def run(cmd, **kw):
    print("+", " ".join(cmd[:6]), "..." if len(cmd) > 6 else "", flush=True)
    subprocess.run(cmd, check=True, **kw)


def newer(target: str, *sources: str) -> bool:
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in sources if os.path.exists(s))


def main(force: bool = False) -> None:
    if not os.path.isdir(REF):
        print("build_ref: /root/reference is absent (GPU box): using the prebuilt oracle/_ref if any")
        return
    os.makedirs(DEPS, exist_ok=True)
    for z in ("eigen/eigen-3.4.0.zip", "hana/hana-boost-1.84.0.zip", "preprocessor/preprocessor-1.84.0-ungar.zip"):
        marker = os.path.join(DEPS, os.path.basename(z) + ".done")
        if not os.path.exists(marker):
            zipfile.ZipFile(os.path.join(REF, "external/config", z)).extractall(DEPS)
            open(marker, "w").close()
    tapes = os.path.join(OUT, "tapes")
    os.makedirs(tapes, exist_ok=True)
    shim = [os.path.join(HERE, "refshim/cppad/cg.hpp"), os.path.join(HERE, "refshim/osqp++.h"), os.path.join(HERE, "ad.hpp")]

    exe = os.path.join(OUT, "function_tests")
    src = os.path.join(HERE, "ref_drivers/function_tests.cpp")
    if force or not newer(exe, src, *shim):
        run([CXX] + flags(tapes) + ["-o", exe, src])
    run([exe], cwd=HERE, stdout=subprocess.DEVNULL)

    os.makedirs(os.path.join(OUT, "src"), exist_ok=True)
    driver = os.path.join(HERE, "ref_drivers/example_driver.cpp")
    for name, N in EXAMPLES:
        ref_src = os.path.join(REF, "example/mpc", f"{name}.example.cpp")
        tdir = os.path.join(tapes, f"{name}_N{N}")
        done = os.path.join(tdir, ".done")
        if not force and newer(done, ref_src, driver, *shim):
            continue
        example = ref_src
        if N != 30:  # the ONLY edit: the horizon constant (quadrotor.example.cpp:52, rc_car:50, quadruped:57)
            text = open(ref_src).read()
            text, count = re.subn(r"constexpr auto N(\s*)= 30_c;", rf"constexpr auto N\1= {N}_c;", text)
            assert count == 1, f"horizon constant not found in {ref_src}"
            example = os.path.join(OUT, "src", f"{name}_N{N}.example.cpp")
            open(example, "w").write(text)
        os.makedirs(tdir, exist_ok=True)
        exe = os.path.join(OUT, f"example_{name}_N{N}")
        run([CXX] + flags(tdir) + [f'-DUNGAR_EXAMPLE_SOURCE="{example}"', "-o", exe, driver])
        run([exe], cwd=HERE, env=dict(os.environ, UNGAR_REF_MAX_SOLVES="0"), stdout=subprocess.DEVNULL)
        open(done, "w").close()

    lib = os.path.join(OUT, "libreftape.so")
    src = os.path.join(HERE, "ref_drivers/tape_tool.cpp")
    if force or not newer(lib, src, *shim):
        run([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", f"-I{HERE}/refshim", "-o", lib, src])
    print("build_ref: ok")


if __name__ == "__main__":
    main(force="--force" in sys.argv)
