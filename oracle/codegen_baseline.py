"""ORACLE — TEST INFRASTRUCTURE ONLY (timed CPU baseline, kind "codegen").

The CPU baseline SURVEY.md §8d / BASELINE.md §3 specify: what CppADCodeGen does for the reference — tape -> straight-line C ->
``gcc -O3 -g -march=native -mtune=native -ffast-math -shared`` (include/ungar/autodiff/function.hpp:516-522, :610-611) -> dlopen —
produced by our own generator because CppAD / CppADCodeGen are absent from the image.

Input: the tapes ``oracle/_ref`` recorded from the reference's UNCHANGED example lambdas (oracle/build_ref.py; they travel to the GPU
box with the prebuilt files).  For every function of a problem (objective, equalities, inequalities) the generator emits

  * ``forward_zero``:  one statement per live tape node over a scratch array ``v[]`` (CppADCodeGen's generated code has the same shape),
  * ``sparse_jacobian``: sparse forward-mode AD done AT GENERATION TIME — every node carries the set of decision variables it depends
    on and one scratch slot per structurally non-zero partial — so the emitted code computes exactly the structural non-zeros
    (parameter columns trimmed, function.hpp:529-550) and none of the zeros the dense stage-wise port (stage_port.cpp) multiplies,
  * objective only: ``sparse_hessian`` (upper triangle) by second-order sparse forward mode,
  * the assembly the SQP adds on top (soft_sqp.hpp:141-158, :245-264): barrier derivatives, q = grad f + J_h^T dZ and the
    upper-triangular J_h^T d2Z J_h accumulated through slot indices fixed at generation time.

The statements are split into functions of a few thousand lines (gcc's time and memory are superlinear in function size;
CppADCodeGen splits its output for the same reason), compiled in parallel, linked into ``oracle/_ref/codegen/<config>.so`` with an
OpenMP batch driver.  ``timed()`` reports nodes/s at 1 thread (the reference's real behaviour: ``Function`` has no threading) and on
all cores.  ``-march=native`` binaries only run on the CPU they were built for: the build keeps a portable ``-march=x86-64-v3`` twin
and ``load()`` picks the native one only when this host's CPU flags cover the build host's.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TAPES = os.path.join(HERE, "_ref", "tapes")
OUT = os.path.join(HERE, "_ref", "codegen")
(OP_INDEP, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SQRT, OP_SIN, OP_COS, OP_TAN, OP_ATAN, OP_ACOS, OP_ASIN, OP_EXP, OP_LOG,
 OP_ABS, OP_POW, OP_ATAN2, OP_CLT, OP_CLE, OP_CGT, OP_CGE, OP_CEQ) = range(24)
NODE = np.dtype([("op", np.uint8), ("a", np.int32), ("b", np.int32), ("c", np.int32), ("d", np.int32), ("k", np.float64)], align=True)
BARRIER = {"quadrotor": (100.0, 2e-5), "rc_car": (100.0, 1e-2), "quadruped": (1.0, 1.0)}
SIZES = {"quadrotor": (13, 4), "rc_car": (6, 2), "quadruped": (13, 24)}
CHUNK = 400  # statements per generated function (gcc -O3 is superlinear in function size: 3000 took 5x longer per line)


def read_tape(path: str):
    raw = open(path, "rb").read()
    magic, nn, nd, ni, flags = np.frombuffer(raw, dtype=np.int64, count=5)
    assert magic == 0x32455041545F4255, f"not a tape file: {path}"
    off = 40
    nodes = np.frombuffer(raw, dtype=NODE, count=nn, offset=off)
    off += nn * NODE.itemsize
    dep_id = np.frombuffer(raw, dtype=np.int32, count=nd, offset=off)
    off += nd * 4
    dep_const = np.frombuffer(raw, dtype=np.float64, count=nd, offset=off)
    return nodes, int(ni), dep_id, dep_const


def tape_path(model: str, N: int, fn: str) -> str:
    return os.path.join(TAPES, f"{model}_N{N}", f"{model}_mpc_{fn}", "cppad_cg", f"{model}_mpc_{fn}_lib.so")


class Emitter:
    """Straight-line statements over a scratch array, cut into functions."""

    def __init__(self, prefix: str):
        self.prefix = prefix
        self.funcs: list[list[str]] = [[]]
        self.n_slots = 0

    def slot(self) -> int:
        self.n_slots += 1
        return self.n_slots - 1

    def emit(self, stmt: str) -> None:
        if len(self.funcs[-1]) >= CHUNK:
            self.funcs.append([])
        self.funcs[-1].append(stmt)

    def assign(self, expr: str) -> str:
        s = self.slot()
        self.emit(f"v[{s}] = {expr};")
        return f"v[{s}]"

    LOCALS_MAX = 300000  # statements; beyond this gcc -O3 needs hours for the local-scalar form (quadruped N = 100: 1.7 M statements)

    def reuse_slots(self) -> None:
        """Temporaries that die inside the function that defines them become local `const double` scalars (registers); only values
        that cross a function boundary keep a slot in the scratch array, and those slots are recycled by liveness — CppADCodeGen
        reuses its temporaries the same way.  (One array slot per statement made the scratch ~14 MB per trajectory: memory-bound.)"""
        import re

        stmts = [st for body in self.funcs for st in body]
        if len(stmts) > self.LOCALS_MAX:
            return  # one array slot per statement: compiles in minutes, runs ~2x slower (measured on the quadrotor)
        tok = re.compile(r"v\[(\d+)\]")
        first, last = {}, {}
        for idx, st in enumerate(stmts):
            for mm in tok.finditer(st):
                k = int(mm.group(1))
                first.setdefault(k, idx)
                last[k] = idx
        local = {k for k in first if first[k] // CHUNK == last[k] // CHUNK}
        expire = {}
        for slot, idx in last.items():
            if slot not in local:
                expire.setdefault(idx, []).append(slot)
        free, mapping, n_new = [], {}, 0
        out = []
        for idx, st in enumerate(stmts):
            lhs = tok.match(st)
            if lhs:
                old = int(lhs.group(1))
                if old in local:
                    st = "const double t" + str(old) + st[lhs.end():]
                elif old not in mapping:
                    if free:
                        mapping[old] = free.pop()
                    else:
                        mapping[old] = n_new
                        n_new += 1
            out.append(tok.sub(lambda mm: f"t{mm.group(1)}" if int(mm.group(1)) in local else f"v[{mapping[int(mm.group(1))]}]", st))
            for old in expire.get(idx, ()):  # the slot's value is dead after this statement
                if old in mapping:
                    free.append(mapping[old])
        self.n_slots = max(n_new, 1)
        self.funcs = [out[i:i + CHUNK] for i in range(0, len(out), CHUNK)] or [[]]

    def source(self, signature_args: str) -> tuple[str, str]:
        """(C source of the chunk functions, body of the driver that calls them in order)."""
        self.reuse_slots()
        parts, calls = [], []
        for i, body in enumerate(self.funcs):
            name = f"{self.prefix}_{i}"
            parts.append(f"void {name}({signature_args}) {{\n" + "\n".join(body) + "\n}\n")
            calls.append(f"{name}(x, v, out);")
        return parts, "\n".join(calls)


def lit(x: float) -> str:
    return repr(float(x))


def gen_function(em: Emitter, nodes, n_indep, dep_id, dep_const, n_dec: int, out_off: dict, order: int):
    """Emits value (+ Jacobian, + Hessian for order 2) code of one tape.  Returns (ny, jac pattern rows, cols, hes pattern).
    Values land at out[out_off['y'] + r]; Jacobian non-zeros at out[out_off['jac'] + e] in row-major / ascending-column order."""
    n = len(nodes)
    live = np.zeros(n, dtype=bool)
    stack = [int(d) for d in dep_id if d >= 0]
    ops, A, Bv, C, D, Kc = nodes["op"], nodes["a"], nodes["b"], nodes["c"], nodes["d"], nodes["k"]
    while stack:
        i = stack.pop()
        if live[i]:
            continue
        live[i] = True
        op = ops[i]
        if op in (OP_INDEP, OP_CONST):
            continue
        stack.append(int(A[i]))
        if op in (OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_ATAN2) or op >= OP_CLT:
            stack.append(int(Bv[i]))
        if op >= OP_CLT:
            stack.append(int(C[i]))
            stack.append(int(D[i]))
    val: dict[int, str] = {}    # node -> C expression of its value (a slot, an input or a literal)
    grad: dict[int, dict] = {}  # node -> {decision variable j: expression of d node / d x_j}
    hess: dict[int, dict] = {}  # node -> {(j, k) j <= k: expression}
    const: dict[int, float] = {}

    def add_terms(terms):  # sum of expressions (strings); '' if none
        terms = [t for t in terms if t]
        return " + ".join(terms)

    def scaled(coef: str, expr: str) -> str:
        if coef == "1.0":
            return expr
        if coef == "-1.0":
            return f"-({expr})"
        return f"({coef}) * ({expr})"

    def lin(i, parts):
        """grad / hess of node i as a linear combination  sum_k coef_k * node_k  (first order chain rule)."""
        g: dict = {}
        for coef, src in parts:
            for j, e in grad.get(src, {}).items():
                g.setdefault(j, []).append(scaled(coef, e))
        grad[i] = {j: em.assign(add_terms(ts)) for j, ts in g.items()}
        if order == 2:
            h: dict = {}
            for coef, src in parts:
                for jk, e in hess.get(src, {}).items():
                    h.setdefault(jk, []).append(scaled(coef, e))
            hess[i] = {jk: ts for jk, ts in h.items()}  # finalised by the caller (second-order terms may be added)

    def finish_hess(i, extra: dict | None = None):
        if order != 2:
            return
        h = hess.get(i, {})
        for jk, ts in (extra or {}).items():
            h.setdefault(jk, []).extend(ts)
        hess[i] = {jk: em.assign(add_terms(ts)) for jk, ts in h.items() if ts}

    def outer(ga: dict, gb: dict, coef: str, sym: bool) -> dict:
        """coef * (ga gb^T + gb ga^T) (sym) or coef * ga ga^T restricted to the upper triangle."""
        out: dict = {}
        for j, ej in ga.items():
            for k, ek in gb.items():
                a, b = (j, k) if j <= k else (k, j)
                term = scaled(coef, f"({ej}) * ({ek})")
                if sym and j == k:
                    term = f"2.0 * ({term})"  # (ga gb^T + gb ga^T)_jj = 2 ga_j gb_j
                if not sym:  # ga ga^T: every unordered pair once, off-diagonal pairs appear twice in the loop
                    if j > k:
                        continue
                    out.setdefault((a, b), []).append(term if j == k else f"2.0 * ({term})")
                else:
                    out.setdefault((a, b), []).append(term)
        return out

    for i in range(n):
        if not live[i]:
            continue
        op, a, b = int(ops[i]), int(A[i]), int(Bv[i])
        if op == OP_INDEP:
            val[i] = f"x[{a}]"
            if a < n_dec:
                grad[i] = {a: "1.0"}
            continue
        if op == OP_CONST:
            val[i] = lit(Kc[i])
            const[i] = float(Kc[i])
            continue
        va = val[a]
        vb = val[b] if b >= 0 and b in val else None
        if op == OP_ADD:
            val[i] = em.assign(f"{va} + {vb}")
            lin(i, [("1.0", a), ("1.0", b)]); finish_hess(i)
        elif op == OP_SUB:
            val[i] = em.assign(f"{va} - {vb}")
            lin(i, [("1.0", a), ("-1.0", b)]); finish_hess(i)
        elif op == OP_NEG:
            val[i] = em.assign(f"-{va}")
            lin(i, [("-1.0", a)]); finish_hess(i)
        elif op == OP_MUL:
            val[i] = em.assign(f"{va} * {vb}")
            lin(i, [(vb, a), (va, b)])
            finish_hess(i, outer(grad.get(a, {}), grad.get(b, {}), "1.0", True) if order == 2 else None)
        elif op == OP_DIV:
            inv = em.assign(f"1.0 / {vb}")
            val[i] = em.assign(f"{va} * {inv}")
            if b in grad and grad[b]:
                mq = em.assign(f"-{val[i]} * {inv}")
                lin(i, [(inv, a), (mq, b)])
                if order == 2:  # d2(a/b) = (ga gb^T + gb ga^T)(-1/b^2) + 2a/b^3 gb gb^T
                    c1 = em.assign(f"-{inv} * {inv}")
                    c2 = em.assign(f"-2.0 * {mq} * {inv}")
                    ex = outer(grad.get(a, {}), grad.get(b, {}), c1, True)
                    for jk, ts in outer(grad.get(b, {}), grad.get(b, {}), c2, False).items():
                        ex.setdefault(jk, []).extend(ts)
                    finish_hess(i, ex)
            else:
                lin(i, [(inv, a)]); finish_hess(i)
        elif op in (OP_SQRT, OP_SIN, OP_COS, OP_ATAN, OP_ABS, OP_EXP, OP_LOG, OP_TAN):
            if op == OP_SQRT:
                val[i] = em.assign(f"sqrt({va})"); d1 = em.assign(f"0.5 / {val[i]}"); d2 = f"-0.5 * {d1} / {va}"
            elif op == OP_SIN:
                val[i] = em.assign(f"sin({va})"); d1 = em.assign(f"cos({va})"); d2 = f"-{val[i]}"
            elif op == OP_COS:
                val[i] = em.assign(f"cos({va})"); d1 = em.assign(f"-sin({va})"); d2 = f"-{val[i]}"
            elif op == OP_ATAN:
                val[i] = em.assign(f"atan({va})"); d1 = em.assign(f"1.0 / (1.0 + {va} * {va})"); d2 = f"-2.0 * {va} * {d1} * {d1}"
            elif op == OP_ABS:  # CppAD: abs'(0) = 0
                val[i] = em.assign(f"fabs({va})"); d1 = em.assign(f"({va} > 0.0) - ({va} < 0.0)"); d2 = "0.0"
            elif op == OP_EXP:
                val[i] = em.assign(f"exp({va})"); d1 = val[i]; d2 = val[i]
            elif op == OP_LOG:
                val[i] = em.assign(f"log({va})"); d1 = em.assign(f"1.0 / {va}"); d2 = f"-{d1} * {d1}"
            else:
                val[i] = em.assign(f"tan({va})"); d1 = em.assign(f"1.0 + {val[i]} * {val[i]}"); d2 = f"2.0 * {val[i]} * {d1}"
            lin(i, [(d1, a)])
            if order == 2 and grad.get(a):
                c2 = em.assign(d2)
                finish_hess(i, outer(grad[a], grad[a], c2, False))
            else:
                finish_hess(i)
        elif op == OP_POW:
            if b in const and float(const[b]).is_integer():
                p = int(const[b])
                val[i] = em.assign(f"pow({va}, {p})")
                d1 = em.assign(f"{p}.0 * pow({va}, {p - 1})")
                lin(i, [(d1, a)])
                if order == 2 and grad.get(a):
                    c2 = em.assign(f"{p * (p - 1)}.0 * pow({va}, {p - 2})")
                    finish_hess(i, outer(grad[a], grad[a], c2, False))
                else:
                    finish_hess(i)
            else:
                raise NotImplementedError("pow with a non-integer exponent does not occur in the three MPC models")
        elif op >= OP_CLT:
            cmp = {OP_CLT: "<", OP_CLE: "<=", OP_CGT: ">", OP_CGE: ">=", OP_CEQ: "=="}[op]
            c, d = int(C[i]), int(D[i])
            cond = em.assign(f"({va} {cmp} {vb}) ? 1.0 : 0.0")
            ncond = em.assign(f"1.0 - {cond}")
            val[i] = em.assign(f"{cond} != 0.0 ? {val[c]} : {val[d]}")
            lin(i, [(cond, c), (ncond, d)]); finish_hess(i)  # the selected branch is differentiated (CppAD CondExp semantics)
        else:
            raise NotImplementedError(f"tape op {op}")
    # outputs
    rows, cols = [], []
    e = 0
    for r, did in enumerate(dep_id):
        did = int(did)
        em.emit(f"out[{out_off['y'] + r}] = {val[did] if did >= 0 else lit(dep_const[r])};")
        if did >= 0:
            for j in sorted(grad.get(did, {})):
                em.emit(f"out[{out_off['jac'] + e}] = {grad[did][j]};")
                rows.append(r); cols.append(j); e += 1
    hes = []
    if order == 2:
        did = int(dep_id[0])
        for (j, k) in sorted(hess.get(did, {})):
            em.emit(f"out[{out_off['hes'] + len(hes)}] = {hess[did][(j, k)]};")
            hes.append((j, k))
    return len(dep_id), np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64), hes


def generate(model: str, N: int):
    """(chunk functions, driver source, output layout) of the whole per-trajectory evaluation."""
    nx, nu = SIZES[model]
    n_dec = nx * (N + 1) + nu * N
    tapes = {fn: read_tape(tape_path(model, N, fn)) for fn in ("obj", "eqs", "ineqs")}
    m_eq, m_in = len(tapes["eqs"][2]), len(tapes["ineqs"][2])
    # output layout: f | g | h | grad f (dense n_dec) ... placed after a first pass tells the nnz counts
    layout = {}
    src_parts, drivers = [], []
    off = 0
    pats = {}
    for fn, order in (("obj", 2), ("eqs", 1), ("ineqs", 1)):
        nodes, ni, dep_id, dep_const = tapes[fn]
        em = Emitter(f"{fn}")
        ny = len(dep_id)
        # generous provisional offsets, compacted below
        oo = {"y": off, "jac": off + ny, "hes": 0}
        # first pass needs the nnz to place the Hessian: generate with hes after a bound of the Jacobian size
        bound = ny * 64 if fn != "obj" else n_dec
        oo["hes"] = oo["jac"] + bound
        ny, rows, cols, hes = gen_function(em, nodes, ni, dep_id, dep_const, n_dec, oo, order)
        assert rows.size <= bound
        body, calls = em.source("const double* __restrict__ x, double* __restrict__ v, double* __restrict__ out")
        layout[fn] = {"y": oo["y"], "ny": ny, "jac": oo["jac"], "nnz": int(rows.size), "hes": oo["hes"], "nnz_hes": len(hes),
                      "rows": rows, "cols": cols, "hes_pattern": hes, "slots": em.n_slots}
        src_parts.extend(body)
        drivers.append(calls)
        off = oo["hes"] + len(hes) + 8
        pats[fn] = (rows, cols)
    # assembly on top (soft_sqp.hpp:141-158, :245-264): barrier, q, upper-triangular Gauss-Newton block in slot order
    k_bar, eps = BARRIER[model]
    a1 = k_bar; b1 = -0.5 * a1 * eps
    c1 = -1.0 / 3.0 * (-b1 - a1 * eps) * eps - 0.5 * a1 * eps * eps - b1 * eps
    a2 = (-b1 - a1 * eps) / (eps * eps)
    Li = layout["ineqs"]
    q_off = off; off += n_dec
    z_off = off; off += 1
    rows_h, cols_h = pats["ineqs"]
    gn = {}
    asm = [f"for (int j = 0; j < {n_dec}; ++j) out[{q_off} + j] = 0.0;"]
    for e, (r, c) in enumerate(zip(layout["obj"]["rows"], layout["obj"]["cols"])):
        asm.append(f"out[{q_off + int(c)}] = out[{layout['obj']['jac'] + e}];")
    asm.append("double zsum = 0.0;")
    asm.append(f"for (int i = 0; i < {m_in}; ++i) {{ const double xx = -out[{Li['y']} + i]; double b0, dz, d2;"
               f" if (xx < 0.0) {{ b0 = 0.5*{lit(a1)}*xx*xx + {lit(b1)}*xx + {lit(c1)}; dz = -({lit(a1)}*xx + {lit(b1)}); d2 = {lit(a1)}; }}"
               f" else if (xx < {lit(eps)}) {{ b0 = {lit(a2 / 3.0)}*xx*xx*xx + 0.5*{lit(a1)}*xx*xx + {lit(b1)}*xx + {lit(c1)}; dz = -({lit(a2)}*xx*xx + {lit(a1)}*xx + {lit(b1)}); d2 = 2.0*{lit(a2)}*xx + {lit(a1)}; }}"
               f" else {{ b0 = 0.0; dz = 0.0; d2 = 0.0; }} zsum += b0; v[i] = dz; v[{m_in} + i] = d2; }}")
    asm.append(f"out[{z_off}] = zsum;")
    start = 0
    gn_slots = {}
    for r in range(m_in):
        es = [e for e in range(start, len(rows_h)) if rows_h[e] == r]
        start += len(es)
        for e in es:
            asm.append(f"out[{q_off + int(cols_h[e])}] += v[{r}] * out[{Li['jac'] + e}];")
        for ia, ea in enumerate(es):
            for eb in es[ia:]:
                key = (int(cols_h[ea]), int(cols_h[eb]))
                if key not in gn_slots:
                    gn_slots[key] = len(gn_slots)
    gn_off = off; off += len(gn_slots)
    asm.append(f"for (int s = 0; s < {len(gn_slots)}; ++s) out[{gn_off} + s] = 0.0;")
    start = 0
    for r in range(m_in):
        es = [e for e in range(start, len(rows_h)) if rows_h[e] == r]
        start += len(es)
        for ia, ea in enumerate(es):
            for eb in es[ia:]:
                s_ = gn_slots[(int(cols_h[ea]), int(cols_h[eb]))]
                asm.append(f"out[{gn_off + s_}] += v[{m_in + r}] * out[{Li['jac'] + ea}] * out[{Li['jac'] + eb}];")
    total = off
    n_slots = max(max(layout[f]["slots"] for f in layout), 2 * m_in) + 8
    # assembly: first function = up to and including the barrier loop and the z store, the rest in chunks
    SIG = "const double* __restrict__ x, double* __restrict__ v, double* __restrict__ out"
    cut = next(i for i, st in enumerate(asm) if st.startswith(f"out[{z_off}]")) + 1
    src_parts.append(f"void asm_0({SIG}) {{\n" + "\n".join(asm[:cut]) + "\n}\n")
    calls = ["asm_0(x, v, out);"]
    for ci, i in enumerate(range(cut, len(asm), CHUNK)):
        src_parts.append(f"void asm_{ci + 1}({SIG}) {{\n" + "\n".join(asm[i:i + CHUNK]) + "\n}\n")
        calls.append(f"asm_{ci + 1}(x, v, out);")
    all_calls = "\n".join(drivers) + "\n" + "\n".join(calls)
    protos = "\n".join(f"void {c.split('(')[0]}({SIG});" for c in all_calls.split("\n") if c.strip())
    src = ["#include <math.h>", "#include <stdlib.h>", "#include <omp.h>", protos, ""]
    src.append("void codegen_eval(const double* x, double* v, double* out) {\n" + all_calls + "\n}\n")
    src.append(f"int codegen_out_size(void) {{ return {total}; }}\nint codegen_scratch_size(void) {{ return {n_slots}; }}\n")
    src.append("""void codegen_batch(const double* xp, long long batch, long long ld_xp, double* out, long long ld_out, int threads) {
#pragma omp parallel num_threads(threads)
    {
        double* v = (double*)malloc(sizeof(double) * (size_t)codegen_scratch_size());
#pragma omp for schedule(static)
        for (long long b = 0; b < batch; ++b) codegen_eval(xp + b * ld_xp, v, out + b * ld_out);
        free(v);
    }
}
""")
    meta = {"total": total, "q": q_off, "z": z_off, "gn": gn_off, "gn_pattern": sorted(gn_slots, key=gn_slots.get), "layout": layout,
            "n_dec": n_dec, "m_eq": m_eq, "m_in": m_in}
    return src_parts, "\n".join(src), meta


def cpu_flags() -> set:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def build(model: str, N: int, force: bool = False) -> dict:
    """Generates and compiles oracle/_ref/codegen/<model>_N<N>_{native,v3}.so (needs oracle/_ref/tapes).  Returns paths + metadata."""
    os.makedirs(OUT, exist_ok=True)
    base = os.path.join(OUT, f"{model}_N{N}")
    meta_path = base + ".npz"
    want = [base + "_native.so", base + "_v3.so", meta_path]
    if not force and all(os.path.exists(p) for p in want):
        return {"base": base}
    t0 = time.time()
    funcs, driver, meta = generate(model, N)
    gen_s = time.time() - t0
    workers = max(1, os.cpu_count() or 1)
    n_units = max(1, min(4 * workers, len(funcs) // 8))
    units = []
    for u in range(n_units):
        path = f"{base}_u{u}.c"
        with open(path, "w") as f:
            f.write("#include <math.h>\n" + "\n".join(funcs[u::n_units]))
        units.append(path)
    with open(base + "_driver.c", "w") as f:
        f.write(driver)
    units.append(base + "_driver.c")
    lines = sum(fn.count("\n") for fn in funcs) + driver.count("\n")
    # -g of function.hpp:610 is left out: it does not change the generated code and doubles gcc's time on 1e6-line inputs
    for tag, arch in (("native", ["-march=native", "-mtune=native"]), ("v3", ["-march=x86-64-v3"])):
        objs, running = [], []
        for path in units:
            obj = path[:-2] + f"_{tag}.o"
            objs.append(obj)
            running.append(subprocess.Popen(["gcc", "-O3", "-ffast-math", "-fopenmp", "-fPIC", *arch, "-c", "-o", obj, path],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            while len(running) >= workers:
                p = running.pop(0)
                out, _ = p.communicate()
                if p.returncode != 0:
                    raise RuntimeError(f"gcc failed for the {tag} codegen baseline:\n{out[-2000:]}")
        for p in running:
            out, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError(f"gcc failed for the {tag} codegen baseline:\n{out[-2000:]}")
        subprocess.run(["gcc", "-shared", "-fopenmp", "-o", f"{base}_{tag}.so", *objs, "-lm"], check=True)
        for o in objs:
            os.remove(o)
    for path in units:
        os.remove(path)
    np.savez(meta_path, total=meta["total"], q=meta["q"], z=meta["z"], gn=meta["gn"], n_dec=meta["n_dec"], m_eq=meta["m_eq"], m_in=meta["m_in"],
             gn_pattern=np.array(meta["gn_pattern"], dtype=np.int64).reshape(-1, 2),
             flags=np.array(sorted(cpu_flags())), source_lines=lines, build_seconds=time.time() - t0, generate_seconds=gen_s,
             **{f"{fn}_{k}": np.asarray(meta["layout"][fn][k]) for fn in meta["layout"] for k in ("y", "ny", "jac", "nnz", "hes", "nnz_hes", "rows", "cols")},
             obj_hes_pattern=np.array(meta["layout"]["obj"]["hes_pattern"], dtype=np.int64).reshape(-1, 2))
    return {"base": base}


class Baseline:
    def __init__(self, model: str, N: int):
        base = os.path.join(OUT, f"{model}_N{N}")
        if not os.path.exists(base + ".npz"):
            if os.path.isdir(os.path.join(TAPES, f"{model}_N{N}")):
                build(model, N)
            else:
                raise FileNotFoundError(f"no codegen baseline and no tapes for {model} N={N}")
        self.meta = np.load(base + ".npz", allow_pickle=False)
        built_flags = set(self.meta["flags"].tolist())
        self.variant = "native" if built_flags and built_flags <= cpu_flags() and os.path.exists(base + "_native.so") else "v3"
        self.lib = ctypes.CDLL(f"{base}_{self.variant}.so")
        self.lib.codegen_batch.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
        self.total = int(self.lib.codegen_out_size())
        self.N, self.model = N, model

    def run(self, xp: np.ndarray, threads: int, out: np.ndarray | None = None):
        xp = np.ascontiguousarray(xp, dtype=np.float64)
        if out is None:
            out = np.empty((xp.shape[0], self.total))
        t0 = time.perf_counter()
        self.lib.codegen_batch(xp.ctypes.data, xp.shape[0], xp.shape[1], out.ctypes.data, out.shape[1], threads)
        return time.perf_counter() - t0, out


def timed(model: str, N: int, pool: np.ndarray, seconds: float = 8.0) -> dict:
    """nodes/s of the generated straight-line code at 1 thread and on all host threads, on a bounded sample of `pool`."""
    bl = Baseline(model, N)
    threads = os.cpu_count() or 1
    res = {}
    for label, nt in (("one_thread", 1), ("all_cores", threads)):
        probe = pool[:max(nt, 2)]
        bl.run(probe, nt)
        t, _ = bl.run(probe, nt)
        per = t / len(probe)
        n = int(max(nt, min(len(pool), (seconds / 2) / max(per, 1e-9))))
        n = max(nt, (n // nt) * nt)
        t_total, passes = 0.0, 0
        while t_total < seconds / 2 and passes < 1000:
            t, _ = bl.run(pool[:n], nt)
            t_total += t
            passes += 1
        res[label] = {"value": n * N * passes / t_total, "unit": "nodes/s", "cores": nt,
                      "sample": f"first {n} trajectories x {passes} passes ({t_total:.1f} s)"}
    res["kind"] = "codegen"
    res["variant"] = bl.variant
    res["what"] = ("straight-line C generated from the reference's own tapes (oracle/_ref): f, grad f, Hessian of f, g, J_g, h, J_h as sparse "
                   "non-zeros only + barrier, q and the Gauss-Newton block; gcc -O3 -g -ffast-math -march=" + ("native -mtune=native" if bl.variant == "native" else "x86-64-v3 (this host's CPU differs from the build host's)") +
                   f" (function.hpp:610-611); {int(bl.meta['source_lines'])} lines of C")
    return res


if __name__ == "__main__":
    cfgs = [("quadrotor", 30), ("rc_car", 60), ("quadruped", 100)] if len(sys.argv) < 2 else [(sys.argv[1], int(sys.argv[2]))]
    for m, n in cfgs:
        t0 = time.time()
        build(m, n, force="--force" in sys.argv)
        print(f"codegen baseline {m} N={n}: built in {time.time() - t0:.1f} s")
