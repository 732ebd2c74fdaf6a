// Generic path behind MakeFunction: a recorded operation tape of an ARBITRARY user lambda is
// evaluated on the GPU by a small register machine — values, first and second directional derivatives.
//
// Replaces, for functions that have no hand-written kernel, the straight-line C that CppADCodeGen emits and gcc compiles
// (include/ungar/autodiff/function.hpp:453-503) and the three entry points of the generated library the reference calls
// (function.hpp:186-189 forward_zero, :224-228 sparse_jacobian, :252-257 sparse_hessian).  Semantics are CppAD's: CondExp
// differentiates the selected branch, abs'(0) = 0, structural sparsity comes from dependency propagation.
//
// Execution model.  One thread = one (trajectory, direction) pair; all threads run the same instruction stream, so a warp
// never diverges except for predicated selects.  Node values live in a scratch array in HBM/L2 laid out
// [slot][component][thread] (slots are assigned by liveness on the host, typically a few hundred for a 100 k-node tape), so
// every access is a coalesced 256-byte line per warp.  ORDER selects the jet that is propagated:
//     0  value                                  -> forward_zero
//     1  value + d/dt along a seed direction    -> sparse Jacobian by column compression: direction = colour class
//     2  value + d/dt + d2/dt2                  -> sparse Hessian: d^T H d for d = e_i and d = e_i + e_j (no reverse sweep,
//                                                  no step size:  H_ij = (q(e_i + e_j) - q(e_i) - q(e_j)) / 2 exactly)
// This is the generic fallback: throughput comes from the batch x directions parallelism, not from per-model structure
// (the three reference MPC problems have hand-written kernels, sweep_*.cuh).
#pragma once

#include <cuda_runtime.h>

namespace ub {
namespace tape {

enum Op : int {
    T_INDEP = 0, T_CONST, T_ADD, T_SUB, T_MUL, T_DIV, T_NEG, T_SQRT, T_SIN, T_COS, T_TAN, T_ATAN, T_ACOS, T_ASIN, T_EXP, T_LOG, T_ABS,
    T_POW, T_ATAN2, T_CLT, T_CLE, T_CGT, T_CGE, T_CEQ,
    T_OUTPUT,       // dependent b <- slot a
    T_OUTPUT_CONST  // dependent b <- consts[a]
};

struct Instr {
    int op, dst, a, b, c, d;
};

struct Program {
    const Instr* code;
    const double* consts;
    int n_instr, n_slots, n_indep, n_dep;
};

// Seed of direction `dir` on independent i:  kind 0: color[i] == dir  (Jacobian);  kind 1: i == pi[dir] || i == pj[dir] (Hessian).
struct Seeds {
    int kind;
    const int* color;
    const int* pi;
    const int* pj;
};

template <int ORDER>
struct Jet {
    double v, d, dd;
};

template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_const(double k) { return {k, 0.0, 0.0}; }

template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_add(const Jet<ORDER>& a, const Jet<ORDER>& b) { return {a.v + b.v, a.d + b.d, a.dd + b.dd}; }
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_sub(const Jet<ORDER>& a, const Jet<ORDER>& b) { return {a.v - b.v, a.d - b.d, a.dd - b.dd}; }
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_neg(const Jet<ORDER>& a) { return {-a.v, -a.d, -a.dd}; }
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_mul(const Jet<ORDER>& a, const Jet<ORDER>& b) {
    Jet<ORDER> r{a.v * b.v, 0.0, 0.0};
    if (ORDER >= 1) r.d = a.d * b.v + a.v * b.d;
    if (ORDER >= 2) r.dd = a.dd * b.v + 2.0 * a.d * b.d + a.v * b.dd;
    return r;
}
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_div(const Jet<ORDER>& a, const Jet<ORDER>& b) {
    const double inv = 1.0 / b.v;
    Jet<ORDER> r{a.v * inv, 0.0, 0.0};
    if (ORDER >= 1) r.d = (a.d - r.v * b.d) * inv;
    if (ORDER >= 2) r.dd = (a.dd - 2.0 * r.d * b.d - r.v * b.dd) * inv;
    return r;
}
// y = phi(a) with phi' = f1, phi'' = f2 at a.v
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_chain(const Jet<ORDER>& a, double y, double f1, double f2) {
    Jet<ORDER> r{y, 0.0, 0.0};
    if (ORDER >= 1) r.d = f1 * a.d;
    if (ORDER >= 2) r.dd = f1 * a.dd + f2 * a.d * a.d;
    return r;
}
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_sqrt(const Jet<ORDER>& a) {
    const double y = sqrt(a.v), f1 = 0.5 / y;
    return jet_chain(a, y, f1, ORDER >= 2 ? -0.5 * f1 / a.v : 0.0);
}
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_exp(const Jet<ORDER>& a) {
    const double y = exp(a.v);
    return jet_chain(a, y, y, y);
}
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_log(const Jet<ORDER>& a) {
    const double inv = 1.0 / a.v;
    return jet_chain(a, log(a.v), inv, -inv * inv);
}

// ---- one function per tape operation: the interpreter below and the NVRTC-specialised kernels (tape.cu) share them, so both
// evaluate a tape with the same arithmetic.
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_unary(int op, const Jet<ORDER>& a) {
    switch (op) {
        case T_NEG: return jet_neg(a);
        case T_SQRT: return jet_sqrt(a);
        case T_SIN: { double s, c; sincos(a.v, &s, &c); return jet_chain(a, s, c, -s); }
        case T_COS: { double s, c; sincos(a.v, &s, &c); return jet_chain(a, c, -s, -c); }
        case T_TAN: { const double y = tan(a.v), f1 = 1.0 + y * y; return jet_chain(a, y, f1, 2.0 * y * f1); }
        case T_ATAN: { const double f1 = 1.0 / (1.0 + a.v * a.v); return jet_chain(a, atan(a.v), f1, -2.0 * a.v * f1 * f1); }
        case T_ACOS: { const double s2 = 1.0 - a.v * a.v, f1 = -rsqrt(s2); return jet_chain(a, acos(a.v), f1, a.v * f1 / s2); }
        case T_ASIN: { const double s2 = 1.0 - a.v * a.v, f1 = rsqrt(s2); return jet_chain(a, asin(a.v), f1, a.v * f1 / s2); }
        case T_EXP: return jet_exp(a);
        case T_LOG: return jet_log(a);
        default: {  // T_ABS — CppAD: abs'(x) = sign(x), sign(0) = 0
            const double sg = double((a.v > 0.0) - (a.v < 0.0));
            return Jet<ORDER>{fabs(a.v), sg * a.d, sg * a.dd};
        }
    }
}
// pow(a, k) with a CONSTANT exponent k: closed-form derivatives k a^(k-1), k (k-1) a^(k-2) — finite for a <= 0 (the
// general exp(k log a) form gives NaN derivatives there, e.g. pow(x, 2.0) at x < 0).
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_pow_const(const Jet<ORDER>& a, double k) {
    const double y = pow(a.v, k);
    const double f1 = ORDER >= 1 ? k * pow(a.v, k - 1.0) : 0.0;
    const double f2 = ORDER >= 2 ? k * (k - 1.0) * pow(a.v, k - 2.0) : 0.0;
    return jet_chain(a, y, f1, f2);
}
template <int ORDER>
__device__ __forceinline__ Jet<ORDER> jet_binary(int op, const Jet<ORDER>& a, const Jet<ORDER>& b) {
    switch (op) {
        case T_ADD: return jet_add(a, b);
        case T_SUB: return jet_sub(a, b);
        case T_MUL: return jet_mul(a, b);
        case T_DIV: return jet_div(a, b);
        case T_POW: {  // general power: exp(b log a)
            Jet<ORDER> r = jet_exp(jet_mul(b, jet_log(a)));
            r.v = pow(a.v, b.v);
            return r;
        }
        default: {  // T_ATAN2: atan2(y = a, x = b)
            const double r2 = b.v * b.v + a.v * a.v, inv = 1.0 / r2;
            Jet<ORDER> r = jet_const<ORDER>(atan2(a.v, b.v));
            if (ORDER >= 1) {
                const double num = b.v * a.d - a.v * b.d;
                r.d = num * inv;
                if (ORDER >= 2) r.dd = ((b.v * a.dd - a.v * b.dd) - r.d * 2.0 * (b.v * b.d + a.v * a.d)) * inv;
            }
            return r;
        }
    }
}
__device__ __forceinline__ bool jet_compare(int op, double l, double r) {
    return op == T_CLT ? l < r : op == T_CLE ? l <= r : op == T_CGT ? l > r : op == T_CGE ? l >= r : l == r;
}

template <int ORDER>
__device__ __forceinline__ Jet<ORDER> load_slot(const double* __restrict__ scratch, long long stride, long long t, int slot) {
    constexpr int NC = ORDER + 1;
    const double* p = scratch + (long long)slot * NC * stride + t;
    Jet<ORDER> r{p[0], 0.0, 0.0};
    if (ORDER >= 1) r.d = p[stride];
    if (ORDER >= 2) r.dd = p[2 * stride];
    return r;
}
template <int ORDER>
__device__ __forceinline__ void store_slot(double* __restrict__ scratch, long long stride, long long t, int slot, const Jet<ORDER>& j) {
    constexpr int NC = ORDER + 1;
    double* p = scratch + (long long)slot * NC * stride + t;
    p[0] = j.v;
    if (ORDER >= 1) p[stride] = j.d;
    if (ORDER >= 2) p[2 * stride] = j.dd;
}

// ORDER 0: out = y[batch][ld_out].   ORDER 1: out = vals[batch][ld_out], out_slot[dep * ndir + dir] = nonzero index or -1.
// ORDER 2: out = q[batch][ld_out] with q[b][dir] = sum_dep weights[dep] * d2/dt2 dep.
template <int ORDER>
__global__ void __launch_bounds__(128)
tape_kernel(Program P, Seeds S, const double* __restrict__ x_all, long long ld_x, long long batch, int ndir,
            double* __restrict__ scratch, long long stride, double* __restrict__ out_all, long long ld_out,
            const int* __restrict__ out_slot, const double* __restrict__ weights) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= batch * ndir) return;
    const long long b = t / ndir;
    const int dir = int(t - b * ndir);
    const double* __restrict__ x = x_all + b * ld_x;
    double* __restrict__ out = out_all + b * ld_out;
    int s0 = -1, s1 = -1;
    if (ORDER >= 1 && S.kind == 1) { s0 = S.pi[dir]; s1 = S.pj[dir]; }
    double acc = 0.0;
    for (int pc = 0; pc < P.n_instr; ++pc) {
        const Instr in = P.code[pc];
        Jet<ORDER> r;
        switch (in.op) {
            case T_INDEP: {
                r = jet_const<ORDER>(x[in.a]);
                if (ORDER >= 1) r.d = (S.kind == 0 ? S.color[in.a] == dir : (in.a == s0 || in.a == s1)) ? 1.0 : 0.0;
                break;
            }
            case T_CONST: r = jet_const<ORDER>(P.consts[in.a]); break;
            case T_ADD: case T_SUB: case T_MUL: case T_DIV: case T_ATAN2:
                r = jet_binary(in.op, load_slot<ORDER>(scratch, stride, t, in.a), load_slot<ORDER>(scratch, stride, t, in.b));
                break;
            case T_POW:  // in.c >= 0: the exponent is the constant consts[in.c] (closed-form derivatives)
                r = in.c >= 0 ? jet_pow_const(load_slot<ORDER>(scratch, stride, t, in.a), P.consts[in.c])
                              : jet_binary(T_POW, load_slot<ORDER>(scratch, stride, t, in.a), load_slot<ORDER>(scratch, stride, t, in.b));
                break;
            case T_NEG: case T_SQRT: case T_SIN: case T_COS: case T_TAN: case T_ATAN: case T_ACOS: case T_ASIN: case T_EXP: case T_LOG: case T_ABS:
                r = jet_unary(in.op, load_slot<ORDER>(scratch, stride, t, in.a));
                break;
            case T_CLT: case T_CLE: case T_CGT: case T_CGE: case T_CEQ: {
                const double l = scratch[(long long)in.a * (ORDER + 1) * stride + t], rr = scratch[(long long)in.b * (ORDER + 1) * stride + t];
                const bool take = jet_compare(in.op, l, rr);
                r = load_slot<ORDER>(scratch, stride, t, take ? in.c : in.d);
                break;
            }
            case T_OUTPUT: case T_OUTPUT_CONST: {
                const Jet<ORDER> a = in.op == T_OUTPUT ? load_slot<ORDER>(scratch, stride, t, in.a) : jet_const<ORDER>(P.consts[in.a]);
                if (ORDER == 0) out[in.b] = a.v;
                if (ORDER == 1) {
                    const int e = out_slot[(long long)in.b * ndir + dir];
                    if (e >= 0) out[e] = a.d;
                }
                if (ORDER == 2) acc += weights[in.b] * a.dd;
                continue;
            }
            default: continue;
        }
        store_slot<ORDER>(scratch, stride, t, in.dst, r);
    }
    if (ORDER == 2) out[dir] = acc;
}

// Hessian values from the directional second derivatives: vals[b][e] = q[dia_i[e]] if i == j, else (q[pair[e]] - q[dia_i] - q[dia_j]) / 2.
__global__ void hessian_combine_kernel(const double* __restrict__ q_all, long long ld_q, const int* __restrict__ di, const int* __restrict__ dj,
                                       const int* __restrict__ pr, int nnz, double* __restrict__ vals_all, long long ld_vals, long long batch) {
    const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (id >= batch * nnz) return;
    const long long b = id / nnz;
    const int e = int(id - b * nnz);
    const double* q = q_all + b * ld_q;
    vals_all[b * ld_vals + e] = pr[e] < 0 ? q[di[e]] : 0.5 * (q[pr[e]] - q[di[e]] - q[dj[e]]);
}

}  // namespace tape
}  // namespace ub
