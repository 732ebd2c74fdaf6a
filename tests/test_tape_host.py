"""Host side of the generic tape path (no GPU needed): tracing scalar, tape analysis in ungar_b200_tape_create, structural sparsity,
parameter trimming and colouring.  Known patterns: test/autodiff/function.test.cpp:70-96 (J = [[2 p x], [4 x0, 0, 0, 0]]), :120-136
(H = 2 p I)."""
import ctypes

import numpy as np
import pytest

from ungar_b200 import _lib
from ungar_b200 import autodiff as A


def norm2_times_p(xp):
    x, p = xp[:4], xp[4]
    return p * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3])


def test_reference_test_functions_have_the_reference_patterns():
    f = A.MakeFunction(A.Blueprint(lambda xp: [norm2_times_p(xp), 2.0 * A.pow(xp[0], 2)], 4, 1, "jacobian_test", A.JACOBIAN))
    assert (f.IndependentVariableSize(), f.ParameterSize(), f.DependentVariableSize()) == (4, 1, 2)
    rows, cols = f.JacobianSparsity()
    assert rows.tolist() == [0, 0, 0, 0, 1] and cols.tolist() == [0, 1, 2, 3, 0]  # parameter column trimmed (function.hpp:529-550)
    assert f.ImplementsJacobian() and not f.ImplementsHessian()
    h = A.MakeFunction(A.Blueprint(lambda xp: [norm2_times_p(xp)], 4, 1, "hessian_test", A.ALL))
    rows, cols = h.HessianSparsity()
    assert rows.tolist() == [0, 1, 2, 3] and cols.tolist() == [0, 1, 2, 3]  # upper triangle of the x-x block (function.hpp:552-574)
    full_r, full_c = h._tape.hessian_pattern()  # over all independents: x_i interacts with p
    assert {(0, 4), (4, 0), (3, 4)} <= set(zip(full_r.tolist(), full_c.tolist())) and (4, 4) not in set(zip(full_r.tolist(), full_c.tolist()))
    # a vector-valued function has no Hessian (function.hpp:136-137)
    v = A.MakeFunction(A.Blueprint(lambda xp: [xp[0] * xp[1], xp[1]], 2, 0, "vec", A.ALL))
    assert not v.ImplementsHessian()
    with pytest.raises(_lib.UngarB200Error):
        v.HessianValues(np.zeros(2))


def test_folding_rules_and_dead_nodes():
    def f(xp):
        x, y = xp
        dead = A.sin(x) * A.cos(y)  # noqa: F841  never reaches a dependent
        return [x * 0.0 + y * 1.0, (x + 0.0) / 1.0, 0.0 / x, A.pow(x, 3), 2.0 * 3.0, A.CondExpGt(1.0, 0.0, x, y)]
    nodes, deps, consts = A.record(f, [0.5, 0.7])
    ops = nodes["op"].tolist()
    # y*1 -> y, x*0 -> 0, 0 + y -> y; (x+0)/1 -> x; 0/x -> 0; x^3 = two multiplications; 6 is a parameter; CondExp decided by parameters
    assert deps[0] == 1 and deps[1] == 0 and deps[2] == -1 and consts[2] == 0.0 and deps[4] == -1 and consts[4] == 6.0 and deps[5] == 0
    assert ops.count(A.OP_MUL) == 1 + 2 and ops.count(A.OP_SIN) == 1  # the dead product + x*x, (x*x)*x
    t = A.TapeHandle(nodes, 2, deps, consts)
    info = t.info()
    assert info["live_nodes"] == 2 + 2  # the two independents and the two products of x^3; sin / cos / their product are eliminated
    rows, cols = t.jacobian_pattern()
    assert list(zip(rows.tolist(), cols.tolist())) == [(0, 1), (1, 0), (3, 0), (5, 0)]


def test_conditional_pattern_is_the_union_of_both_branches():
    f = A.MakeFunction(A.Blueprint(lambda xp: [A.CondExpLt(xp[0], xp[1], xp[2] * xp[2], A.sin(xp[3]))], 4, 0, "cond", A.ALL))
    assert f.JacobianSparsity()[1].tolist() == [2, 3]  # the compared values carry no derivative
    assert list(zip(*[a.tolist() for a in f.HessianSparsity()])) == [(2, 2), (3, 3)]


def test_colouring_separates_columns_that_share_a_row():
    n = 40

    def chain(xp):  # banded: row k touches x_k, x_{k+1}, x_{k+2}
        return [xp[k] * xp[k + 1] + A.sin(xp[k + 2]) for k in range(n - 2)]
    f = A.MakeFunction(A.Blueprint(chain, n, 0, "banded", A.JACOBIAN))
    info = f.tape_info()
    assert info["jacobian_colors"] == 3 and info["slots"] <= 8  # bandwidth 3 -> 3 directions; liveness keeps the scratch tiny
    rows, cols = f.JacobianSparsity()
    assert rows.size == 3 * (n - 2)


def test_tape_validation_errors():
    lib = _lib.load()
    nodes = np.zeros(2, dtype=A.NODE_DTYPE)
    nodes[0] = (A.OP_INDEP, 0, -1, -1, -1, 0.0)
    nodes[1] = (A.OP_ADD, 0, 1, -1, -1, 0.0)  # operand 1 is the node itself: not yet defined
    deps = np.array([1], dtype=np.int32)
    h = ctypes.c_void_p()
    assert lib.ungar_b200_tape_create(nodes.ctypes.data, 2, 1, deps.ctypes.data, None, 1, 0, ctypes.byref(h)) == _lib.EINVAL
    assert b"operand out of range" in lib.ungar_b200_last_error()
    nodes[1] = (77, 0, 0, -1, -1, 0.0)
    assert lib.ungar_b200_tape_create(nodes.ctypes.data, 2, 1, deps.ctypes.data, None, 1, 0, ctypes.byref(h)) == _lib.EINVAL
    nodes[1] = (A.OP_MUL, 0, 0, -1, -1, 0.0)
    assert lib.ungar_b200_tape_create(nodes.ctypes.data, 2, 1, deps.ctypes.data, None, 1, 0, ctypes.byref(h)) == _lib.OK
    rows = np.array([0], dtype=np.int64)
    cols = np.array([5], dtype=np.int64)
    assert lib.ungar_b200_tape_set_jacobian_elements(h, rows.ctypes.data, cols.ctypes.data, 1) == _lib.EINVAL
    lib.ungar_b200_tape_destroy(h)


def test_reference_binaries_over_the_product_header_fail_loudly_without_a_gpu():
    """tests/_ref_gpu (the reference's unchanged function.hpp / function.example.cpp compiled against ungar_b200/include/cppad/cg.hpp):
    taping and analysis run on the host, the first evaluation needs the device — there is no CPU fallback behind the header."""
    import os
    import subprocess

    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("behaviour without a GPU")
    except ImportError:
        pass
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref_gpu", "function_example_gpu")
    if not os.path.exists(exe):
        pytest.skip("tests/_ref_gpu was not built (needs /root/reference at build time)")
    proc = subprocess.run([exe], capture_output=True, text=True)
    assert proc.returncode != 0 and "no CPU fallback" in proc.stderr


def _reference_tape(config, function):
    """Tape recorded by oracle/_ref from the reference's own lambda (see tests/test_gpu_tape.py::load_reference_tape)."""
    import glob
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hits = glob.glob(os.path.join(root, "oracle", "_ref", "tapes", config, function, "cppad_cg", "*_lib.so"))
    if not hits:
        pytest.skip("oracle/_ref tapes were not built (needs /root/reference at build time)")
    raw = open(hits[0], "rb").read()
    magic, nn, nd, ni, flags = np.frombuffer(raw, dtype=np.int64, count=5)
    assert int(magic) == 0x32455041545F4255
    nodes = np.frombuffer(raw, dtype=A.NODE_DTYPE, count=int(nn), offset=40)
    dep_id = np.frombuffer(raw, dtype=np.int32, count=int(nd), offset=40 + int(nn) * 32)
    dep_const = np.frombuffer(raw, dtype=np.float64, count=int(nd), offset=40 + int(nn) * 32 + int(nd) * 4)
    return nodes, int(ni), dep_id, dep_const


@pytest.mark.parametrize("name,N", [("quadrotor", 30), ("rc_car", 60), ("quadruped", 30), ("quadruped", 100)])
def test_structural_analysis_of_the_reference_lambdas_matches_the_oracle(oracle, name, N):
    """Host-side analysis of the reference's OWN taped lambdas (no device needed): the Jacobian patterns of the three functions, trimmed
    like function.hpp:529-550, equal the restated oracle's CSR patterns entry for entry; the objective's Hessian pattern (upper
    triangle of the x-x block, function.hpp:552-574) contains the oracle's; liveness keeps the scratch far below the tape length."""
    from ungar_b200 import workloads as W

    mid = W.MODEL_IDS[name]
    s = W.sizes(mid, N)
    nx = s["n_dec"]
    xp = W.synthetic_batch(mid, N, 1, seed=11)[0]
    for fn, suffix in ((0, "obj"), (1, "eqs"), (2, "ineqs")):
        nodes, ni, dep_id, dep_const = _reference_tape(f"{name}_N{N}", f"{name}_mpc_{suffix}")
        assert ni == s["n_xp"]
        t = A.TapeHandle(nodes, ni, dep_id, dep_const)
        r, c = t.jacobian_pattern()
        keep = c < nx
        o_r, o_c, _ = oracle.jacobian(mid, fn, N, xp)
        assert np.array_equal(r[keep], o_r) and np.array_equal(c[keep], o_c), suffix
        t.set_jacobian_elements(r[keep], c[keep])
        info = t.info()
        assert info["live_nodes"] <= nodes.size and info["slots"] < 0.2 * info["live_nodes"] + 64
        if fn == 0:  # a scalar function has one dense row: forward-mode column compression cannot merge its columns (no reverse sweep)
            assert info["jacobian_colors"] == int(keep.sum())
        else:        # block-banded: a few stage widths
            assert 1 <= info["jacobian_colors"] <= 3 * (s["nx"] + s["nu"])
        if fn == 0:
            r2, c2 = t.hessian_pattern()
            keep2 = (r2 < nx) & (c2 < nx) & (c2 >= r2)
            o_r, o_c, _ = oracle.hessian(mid, N, xp)
            assert set(zip(o_r.tolist(), o_c.tolist())) <= set(zip(r2[keep2].tolist(), c2[keep2].tolist()))


@pytest.mark.parametrize("example,functions", [("quadrotor", ("quadrotor_mpc_obj", "quadrotor_mpc_eqs", "quadrotor_mpc_ineqs")),
                                               ("rc_car", ("rc_car_mpc_obj", "rc_car_mpc_eqs", "rc_car_mpc_ineqs"))])
def test_product_header_records_the_same_tapes_as_the_oracle_shim(example, functions, tmp_path):
    """The UNCHANGED reference example compiled against the product's cppad/cg.hpp (tests/_ref_gpu) tapes its three lambdas through
    the reference's own MakeFunction before it evaluates anything, so the tapes exist even without a GPU.  Node for node they equal
    the tapes the oracle shim recorded from the same source (oracle/_ref/tapes) — the ones the golden fixtures were made from."""
    import glob
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_ref_gpu", f"example_{example}_gpu")
    if not os.path.exists(exe) or not os.path.isdir(os.path.join(root, "oracle", "_ref", "tapes", f"{example}_N30")):
        pytest.skip("reference example binaries / oracle tapes were not built (needs /root/reference at build time)")
    subprocess.run([exe], capture_output=True, env=dict(os.environ, UNGAR_B200_MAX_QP_SOLVES="0"), cwd=os.path.dirname(exe))

    def load(path, magic_expected):
        raw = open(path, "rb").read()
        magic, nn, nd, ni, _ = np.frombuffer(raw, dtype=np.int64, count=5)
        assert int(magic) == magic_expected
        nodes = np.frombuffer(raw, dtype=A.NODE_DTYPE, count=int(nn), offset=40)
        deps = np.frombuffer(raw, dtype=np.int32, count=int(nd), offset=40 + int(nn) * 32)
        return nodes, int(ni), deps

    for fn in functions:
        mine = glob.glob(os.path.join(root, "tests", "_ref_gpu", "tapes", fn, "cppad_cg", "*_lib.so"))
        theirs = glob.glob(os.path.join(root, "oracle", "_ref", "tapes", f"{example}_N30", fn, "cppad_cg", "*_lib.so"))
        assert mine and theirs, fn
        a, b = load(mine[0], 0x3130505430303242), load(theirs[0], 0x32455041545F4255)
        assert a[1] == b[1] and np.array_equal(a[2], b[2]) and a[0].size == b[0].size
        for field in ("op", "a", "b", "c", "d", "k"):
            assert np.array_equal(a[0][field], b[0][field]), (fn, field)


def test_a_changed_lambda_under_the_same_name_is_taped_again():
    """SURVEY.md §5 / function.hpp:420-451: the reference keys its generated library by NAME, so a changed lambda under an unchanged
    name silently runs the old code.  Two builds of one driver (tests/ref_drivers/stale_tape_driver.cpp: y = COEFF x0 x1, COEFF = 2 / 3)
    share the function name and the codegen folder and call the reference's unchanged MakeFunction with recompileLibraries = false.
    The product's tapes carry the identity of the executable that recorded them: the same build finds its tape again ("Loading"), the
    other build finds it stale, deletes it before the reference looks for it and tapes its own lambda ("Compiling")."""
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = {t: os.path.join(root, "tests", "_ref_gpu", f"stale_tape_{t}") for t in "ab"}
    if not all(os.path.exists(e) for e in exe.values()):
        pytest.skip("tests/_ref_gpu was not built (needs /root/reference at build time)")
    for d in ("stale_probe", "stale_probe_internal"):
        shutil.rmtree(os.path.join(root, "tests", "_ref_gpu", "tapes", d), ignore_errors=True)
    tape = os.path.join(root, "tests", "_ref_gpu", "tapes", "stale_probe", "cppad_cg", "stale_probe_lib.so")

    def run(tag):
        proc = subprocess.run([exe[tag]], capture_output=True, text=True, cwd=os.path.dirname(exe[tag]))
        assert proc.returncode == 0, proc.stderr
        raw = open(tape, "rb").read()
        nn = int(np.frombuffer(raw, dtype=np.int64, count=5)[1])
        consts = np.frombuffer(raw, dtype=A.NODE_DTYPE, count=nn, offset=40)["k"]
        return ("Loading shared library" in proc.stdout, "Compiling shared library" in proc.stdout, set(consts.tolist()))

    loaded, compiled, k = run("a")
    assert compiled and not loaded and 2.0 in k and 3.0 not in k
    loaded, compiled, k = run("a")          # same build: the tape is found again by name, exactly like the reference's library
    assert loaded and not compiled and 2.0 in k
    loaded, compiled, k = run("b")          # other lambda, same name: NOT served from a's tape
    assert compiled and not loaded and 3.0 in k and 2.0 not in k
    loaded, compiled, k = run("a")
    assert compiled and not loaded and 2.0 in k and 3.0 not in k


def _long_chain(v):
    a, b, c = v[0], v[1], v[2]
    keep = [v[3] * v[4], A.sin(v[5])]
    acc = 0.0
    for k in range(850):
        a, b, c = b * 0.999 + 0.001 * A.sin(c), c - 0.002 * a * b, A.CondExpGt(a, b, a, b) * 0.5 + 0.5 * c
        if k % 200 == 0:
            acc = acc + a * keep[0] + b * keep[1]
    return [a + acc, b * keep[0], c + v[0] * keep[1]]


def test_segmented_kernel_source_never_reads_an_undefined_register():
    """Tapes beyond the single-kernel limit are cut into kernels of 6 k instructions (csrc/tape.cu::generate_segmented_source).  Host-only
    check of the liveness logic on the generated text: inside every kernel a register is loaded from the scratch array or assigned before
    it is read, and whatever is loaded — across a cut, or after being parked because its next use was far away — is the CURRENT value of
    its slot (stored after the slot's last assignment)."""
    import re

    import glob
    import os

    handles = [A.MakeFunction(A.Blueprint(_long_chain, 6, 0, "long_chain_host", A.JACOBIAN))._tape]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hits = glob.glob(os.path.join(root, "oracle", "_ref", "tapes", "quadruped_N30", "quadruped_mpc_eqs", "cppad_cg", "*_lib.so"))
    if hits:  # the reference's own quadruped equality lambda: 20 k instructions, 713 slots alive at once
        raw = open(hits[0], "rb").read()
        magic, nn, nd, ni, _ = np.frombuffer(raw, dtype=np.int64, count=5)
        nodes = np.frombuffer(raw, dtype=A.NODE_DTYPE, count=int(nn), offset=40)
        deps = np.frombuffer(raw, dtype=np.int32, count=int(nd), offset=40 + int(nn) * 32)
        consts = np.frombuffer(raw, dtype=np.float64, count=int(nd), offset=40 + int(nn) * 32 + int(nd) * 4)
        handles.append(A.TapeHandle(nodes, int(ni), deps, consts))
    for tape, order in [(h, o) for h in handles for o in (0, 1, 2)]:
        src, parts = tape.kernel_source(order)
        assert parts >= 3 and src.count("extern \"C\" __global__") == parts
        kernels = re.split(r'extern "C" __global__', src)[1:]
        scratch_valid = set()                            # slots whose CURRENT value sits in the scratch array
        crossed = 0
        for k, body in enumerate(kernels):
            defined = set()                              # registers holding the current value of their slot (they do not survive a kernel)
            for line in body.splitlines():
                line = line.strip()
                m = re.match(r"r(\d+) = load_slot<ORDER>\(scratch, stride, t, (\d+)\);", line)
                if m:
                    assert m.group(1) == m.group(2) and int(m.group(1)) in scratch_valid, (order, k, line)
                    crossed += int(m.group(1)) not in defined
                    defined.add(int(m.group(1)))
                    continue
                m = re.match(r"store_slot<ORDER>\(scratch, stride, t, (\d+), r(\d+)\);", line)
                if m:
                    assert m.group(1) == m.group(2) and int(m.group(1)) in defined, (order, k, line)
                    scratch_valid.add(int(m.group(1)))
                    continue
                m = re.match(r"r(\d+) = ([^;]*);", line)  # (an independent's line goes on to seed its own .d: not a read)
                if m:
                    reads = {int(x) for x in re.findall(r"\br(\d+)\b", m.group(2))}
                    assert reads <= defined, (order, k, line, sorted(reads - defined))
                    defined.add(int(m.group(1)))
                    scratch_valid.discard(int(m.group(1)))  # a new value: whatever the scratch array holds for this slot is stale now
                    continue
                if line.startswith("out[") or line.startswith("{ const int e") or line.startswith("acc +="):
                    reads = {int(x) for x in re.findall(r"\br(\d+)\b", line)}
                    assert reads <= defined, (order, k, line)
        assert crossed > 0                                # values did cross the cuts


def test_generated_kernels_compile_for_sm_100a_without_a_gpu():
    """NVRTC needs no device: the generated single-kernel and segmented sources compile to sm_100a cubins on a CPU host (the driver's
    "does it build" check extended to the generated code)."""
    try:
        from cuda.bindings import nvrtc
    except Exception:
        try:
            from cuda import nvrtc
        except Exception:
            pytest.skip("cuda-python (NVRTC bindings) not importable")
    import os

    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ungar_b200", "csrc")
    small = A.MakeFunction(A.Blueprint(lambda v: [v[4] * sum(x * A.sin(x) for x in v[:4]) + A.pow(v[1], 3)], 4, 1, "small_host", A.JACOBIAN))
    long_ = A.MakeFunction(A.Blueprint(_long_chain, 6, 0, "long_chain_host", A.JACOBIAN))
    for f, order in ((small, 1), (small, 2), (long_, 0)):  # (order 1 of the long tape compiles for a minute: the GPU tests do that)
        src, parts = f._tape.kernel_source(order)
        err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"tape_special.cu", 0, [], [])
        assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS
        opts = [b"--gpu-architecture=sm_100a", f"-I{hdr}".encode(), b"-I/usr/local/cuda/include", b"--std=c++17", b"-default-device", b"--fmad=true"]
        (err,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
        if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
            _, n = nvrtc.nvrtcGetProgramLogSize(prog)
            log = b" " * n
            nvrtc.nvrtcGetProgramLog(prog, log)
            raise AssertionError(log.decode(errors="replace")[-2000:])
        err, n = nvrtc.nvrtcGetCUBINSize(prog)
        assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS and n > 1000
        nvrtc.nvrtcDestroyProgram(prog)


def _scalar_objective(v):
    """A scalar function touching every tape operation the reference's objectives use, with slot reuse and a CondExp."""
    n = len(v)
    acc = 0.0
    for i in range(n - 1):
        d = v[i] - 0.5 * v[i + 1]
        acc = acc + d * d + 0.1 * A.sin(v[i]) * A.cos(v[i + 1]) + A.sqrt(1.0 + v[i] * v[i]) / (2.0 + A.exp(-v[i + 1]))
        acc = acc + A.CondExpGt(v[i], v[i + 1], v[i] * v[i + 1], A.pow(v[i], 2)) + A.atan2(v[i], 1.5 + v[i + 1] * v[i + 1]) - A.log(2.0 + A.abs_(d))
    return [acc]


def test_reverse_sweep_source_is_consistent_and_compiles():
    """Reverse sweep of a scalar function (csrc/tape.cu::generate_reverse_source): host-only replay of the generated text with plain
    floats — forward statements through a tiny interpreter of the op functions, backward statements as written — against central
    differences, and an NVRTC compile for sm_100a."""
    import math
    import re

    n = 24
    f = A.MakeFunction(A.Blueprint(_scalar_objective, n, 0, "scalar_objective_host", A.JACOBIAN))
    src, parts = f._tape.kernel_source(3)
    assert parts == 1 and "tape_reverse" in src
    rng = np.random.default_rng(2)
    x = 0.3 + 0.4 * rng.random(n)
    rows, cols = f.JacobianSparsity()
    elem = -np.ones(n, dtype=int)
    elem[cols] = np.arange(cols.size)

    # --- replay: translate the generated statements into Python (values only; the op functions restated here)
    OPS = {2: lambda a, b: a + b, 3: lambda a, b: a - b, 4: lambda a, b: a * b, 5: lambda a, b: a / b, 18: lambda a, b: math.atan2(a, b),
           17: lambda a, b: a ** b}
    UN = {6: lambda a: -a, 7: math.sqrt, 8: math.sin, 9: math.cos, 10: math.tan, 11: math.atan, 12: math.acos, 13: math.asin, 14: math.exp,
          15: math.log, 16: abs}
    CMP = {19: lambda a, b: a < b, 20: lambda a, b: a <= b, 21: lambda a, b: a > b, 22: lambda a, b: a >= b, 23: lambda a, b: a == b}
    r, a, V, out = {}, {}, {}, np.zeros(cols.size)
    body = src[src.index("tape_reverse"):]
    for line in body.splitlines():
        line = line.strip()
        m = re.match(r"Jet<ORDER> r(\d+); double a(\d+) = 0.0;", line)
        if m:
            a[int(m.group(2))] = 0.0
            continue
        m = re.match(r"V\((\d+)\) = r(\d+)\.v;", line)
        if m:
            V[int(m.group(1))] = r[int(m.group(2))]
            continue
        m = re.match(r"r(\d+) = jet_const<ORDER>\(x\[(\d+)\]\);", line)
        if m:
            r[int(m.group(1))] = float(x[int(m.group(2))])
            continue
        m = re.match(r"r(\d+) = jet_const<ORDER>\((\S+)\);", line)
        if m:
            r[int(m.group(1))] = float.fromhex(m.group(2))
            continue
        m = re.match(r"r(\d+) = jet_binary<ORDER>\((\d+), r(\d+), r(\d+)\);", line)
        if m:
            r[int(m.group(1))] = OPS[int(m.group(2))](r[int(m.group(3))], r[int(m.group(4))])
            continue
        m = re.match(r"r(\d+) = jet_pow_const<ORDER>\(r(\d+), (\S+)\);", line)
        if m:
            r[int(m.group(1))] = r[int(m.group(2))] ** float.fromhex(m.group(3))
            continue
        m = re.match(r"r(\d+) = jet_unary<ORDER>\((\d+), r(\d+)\);", line)
        if m:
            r[int(m.group(1))] = UN[int(m.group(2))](r[int(m.group(3))])
            continue
        m = re.match(r"r(\d+) = jet_compare\((\d+), r(\d+)\.v, r(\d+)\.v\) \? r(\d+) : r(\d+);", line)
        if m:
            r[int(m.group(1))] = r[int(m.group(5))] if CMP[int(m.group(2))](r[int(m.group(3))], r[int(m.group(4))]) else r[int(m.group(6))]
            continue
        # ---- backward statements: C expressions over a<k>, V(i), g and a few locals — evaluated as written
        if line.startswith("{ const int e = elem["):
            m = re.match(r"\{ const int e = elem\[(\d+)\]; if \(e >= 0\) out\[e\] \+= a(\d+); a(\d+) = 0.0; \}", line)
            j, s_ = int(m.group(1)), int(m.group(2))
            if elem[j] >= 0:
                out[elem[j]] += a[s_]
            a[s_] = 0.0
            continue
        m = re.match(r"a(\d+) \+= 1\.0;", line)
        if m:
            a[int(m.group(1))] += 1.0
            continue
        m = re.match(r"a(\d+) = 0\.0;", line)
        if m:
            a[int(m.group(1))] = 0.0
            continue
        if line.startswith("{ const double g = a"):
            env = {"pow": math.pow, "log": math.log, "sin": math.sin, "cos": math.cos, "rsqrt": lambda z: 1.0 / math.sqrt(z),
                   "double": float, "jet_compare": lambda op, p, q: CMP[op](p, q)}
            stmts = [t.strip() for t in line.strip("{} ").split(";") if t.strip()]
            for st in stmts:
                st = re.sub(r"V\((-?\d+)\)", lambda mm: repr(V[int(mm.group(1))]), st)
                st = re.sub(r"\ba(\d+)\b", r"A[\1]", st)
                st = re.sub(r"0x[0-9a-fA-F.]+p[+-]?\d+", lambda mm: repr(float.fromhex(mm.group(0))), st)
                if st.startswith("const double "):
                    for piece in st[len("const double "):].split(", "):
                        name, expr = piece.split(" = ", 1)
                        env[name.strip()] = eval(expr, {"A": a}, env)
                elif st.startswith("if ("):
                    mm = re.match(r"if \((.*)\) (A\[\d+\]) \+= g$", st)
                    cond = eval(mm.group(1), {"A": a}, env)
                    pending_else = (mm.group(2), cond)
                    if cond:
                        exec(f"{mm.group(2)} += g", {"A": a}, env)
                elif st.startswith("else "):
                    if not pending_else[1]:
                        exec(st[5:], {"A": a}, env)
                else:
                    exec(st, {"A": a}, env)
            continue
    ref = np.zeros(cols.size)
    fx = lambda z: float(_scalar_objective(list(z))[0])  # noqa: E731
    for e, j in enumerate(cols):
        h = 1e-6
        xp_, xm_ = x.copy(), x.copy()
        xp_[j] += h
        xm_[j] -= h
        ref[e] = (fx(xp_) - fx(xm_)) / (2 * h)
    assert np.allclose(out, ref, rtol=2e-6, atol=1e-8), np.max(np.abs(out - ref))

    try:
        from cuda.bindings import nvrtc
    except Exception:
        try:
            from cuda import nvrtc
        except Exception:
            return
    import os

    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ungar_b200", "csrc")
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"tape_reverse.cu", 0, [], [])
    opts = [b"--gpu-architecture=sm_100a", f"-I{hdr}".encode(), b"-I/usr/local/cuda/include", b"--std=c++17", b"-default-device", b"--fmad=true"]
    (err,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        _, nlog = nvrtc.nvrtcGetProgramLogSize(prog)
        log = b" " * nlog
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise AssertionError(log.decode(errors="replace")[-2000:])
