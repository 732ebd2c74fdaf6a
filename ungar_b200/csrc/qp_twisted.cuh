// Batched equality-constrained QP solve for the quadruped NMPC on the COMPACT record (compact.cuh) — round-2 replacement of
// qp_schur.cuh.  Same mathematics (numpy statement: oracle/qp_reference.py::schur_twisted):
//     min_d 1/2 d^T P d + q^T d   s.t.  A d = -g         (what SoftSQPOptimizer hands to OSQP, soft_sqp.hpp:141-158, :193-233)
// P is block diagonal over the stage variables w_j = [x_j; u_j] with a closed-form inverse, the constraint rows are grouped as
//     nu_j = [ defect of stage j-1 (x_0 - x_measured for j = 0) ; contact rows of stage j ]                    (29 rows)
// every group touches only w_{j-1} (through V_{j-1} = [A_{j-1}; Cp_j]) and w_j (through U_j = [I 0; Cs_j]), so the Schur complement
// S = A P^-1 A^T + delta I on the multipliers is block tridiagonal with 29 x 29 blocks.
//
// What changed against round 1 (2.27 ms, 0.24 of the HBM peak, 255 registers with spills, 7 warps / SM):
//   * TWISTED elimination: groups 0 .. m-1 are eliminated top-down by warp 0, groups N .. m+1 bottom-up by warp 1 of the same
//     64-thread CTA; the two dependent chains are half as long and run concurrently (14 warps / SM for 1024 trajectories); they
//     meet at group m = N / 2 and the substitutions run outward from there, again one direction per warp.
//   * the record is the compact one: a stage is two TMA bulk loads (2832 B + 2080 B, cp.async.bulk + mbarrier) instead of ~90
//     8-byte cp.async with index arithmetic, and 2.6x fewer bytes.
//   * the triangular solve of the off-diagonal block is FUSED into the Cholesky sweep of the diagonal block that produces its
//     operand: column c of L is broadcast once through shared memory and serves both the trailing update of S and the
//     right-looking update of the next group's coupling rows.  No L image in shared memory, no separate TRSM pass.
//   * L leaves as packed columns with coalesced 8-byte stores straight from registers (3.9 KB per group instead of 7.2 KB).
//   * rows are handled in two register forms, a dense 31-column "pattern" row (q, w and the 24 inputs) plus two scalars for the
//     p / v columns, and a 12-entry sparse row for U; the register peak is the fused sweep (two 29-entry rows), so the kernel fits
//     7 CTAs per SM without the 24 KB images of round 1.
// Lane i < 29 owns ROW i of every 29-row block.  FP64 tensor cores (mma.sync.m8n8k4.f64) do the rank-29 update S -= Lo Lo^T.
#pragma once

#include <cstdint>

#include "compact.cuh"
#include "qp_schur.cuh"  // qp_dmma, qp_ld2, qp_st2

namespace ub {

struct QpT {
    static constexpr int G = 29, LS = 30;
    using K = Compact;
    // per-warp shared memory (doubles)
    static constexpr int oLO = 0, oSM0 = G * LS, oSM1 = oSM0 + K::SMALL, oAB = oSM1 + K::SMALL, oY = oAB + K::APART, oCOL = oY + 32,
                         oBAR = oCOL + 64, PER_WARP = oBAR + 4;
    static constexpr int SMEM_BYTES = 2 * PER_WARP * 8;
    // per-group global workspace (doubles): packed columns of L (column c: rows c..28), 1 / L_ii, y
    static constexpr int wsInv = 435, wsY = 464, WS_GROUP = 494;
    // scratch vectors of the outward pass live behind the L image in the LO region
    static constexpr int oVA = 496, oVC = 536, oNU2 = 576;
    static_assert((PER_WARP * 8) % 16 == 0 && (oSM0 * 8) % 16 == 0 && (oSM1 * 8) % 16 == 0 && (oAB * 8) % 16 == 0 && (WS_GROUP * 8) % 16 == 0, "TMA alignment");
    __host__ __device__ static constexpr int colstart(int c) { return 29 * c - (c * (c - 1)) / 2; }
};

// ---------------------------------------------------------------------------------------------------------------------------------
// mbarrier + bulk-copy helpers (one barrier per buffer per warp; lane 0 arms and issues, every lane waits on the phase)
// ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned qt_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void qt_bar_init(uint64_t* bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(qt_smem(bar)) : "memory"); }
__device__ __forceinline__ void qt_bar_expect(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(qt_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void qt_bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(qt_smem(dst)), "l"(src), "r"(bytes),
                 "r"(qt_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void qt_bar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(qt_smem(bar)),
        "r"(parity)
        : "memory");
}

// One buffer with its barrier and phase bit.  load(): all lanes must have finished reading the buffer (the caller syncs the warp).
struct QtBuf {
    double* buf;
    uint64_t* bar;
    unsigned phase;
    __device__ __forceinline__ void load(const double* src, unsigned bytes, int lane) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was modified in place through the generic proxy
            qt_bar_expect(bar, bytes);
            qt_bulk_load(buf, src, bytes, bar);
        }
    }
    __device__ __forceinline__ void wait() {
        qt_bar_wait(bar, phase);
        phase ^= 1u;
    }
};

// ---------------------------------------------------------------------------------------------------------------------------------
// Stage data accessors.  `sm`: small part of a chunk (Cs, Cp, g, q -> t, Hd/Hb -> P^-1 in place); `ab`: AQ | AP of a chunk.
// ---------------------------------------------------------------------------------------------------------------------------------
using QK = Compact;

// P_j^-1 and t_j = P_j^-1 q_j, in place.  Hd -> reciprocals; Hb[b] (upper triangle of a symmetric 3x3) -> upper triangle of its inverse.
__device__ __forceinline__ void qt_pinv_t(double* __restrict__ sm, int lane, bool inputs) {
    if (lane < 13) {
        const double r = 1.0 / sm[QK::oHd + lane];
        sm[QK::oHd + lane] = r;
        sm[QK::oQ + lane] *= r;
    } else if (lane < 21) {
        const int b = lane - 13;
        double* h = sm + QK::oHb + 6 * b;
        double* qv = sm + QK::oQ + 13 + 3 * b;
        if (inputs) {
            const double a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5];
            const double c00 = d * f - e * e, c01 = c * e - bb * f, c02 = bb * e - c * d;
            const double id = 1.0 / (a * c00 + bb * c01 + c * c02);
            const double i00 = c00 * id, i01 = c01 * id, i02 = c02 * id, i11 = (a * f - c * c) * id, i12 = (bb * c - a * e) * id, i22 = (a * d - bb * bb) * id;
            h[0] = i00; h[1] = i01; h[2] = i02; h[3] = i11; h[4] = i12; h[5] = i22;
            const double q0 = qv[0], q1 = qv[1], q2 = qv[2];
            qv[0] = i00 * q0 + i01 * q1 + i02 * q2;
            qv[1] = i01 * q0 + i11 * q1 + i12 * q2;
            qv[2] = i02 * q0 + i12 * q1 + i22 * q2;
        } else {
            h[0] = h[1] = h[2] = h[3] = h[4] = h[5] = 0.0;
            qv[0] = qv[1] = qv[2] = 0.0;
        }
    }
    __syncwarp();
}

// y = P^-1 x for a 37-vector in shared memory (lanes 0..12 the diagonal part, lanes 13..20 one 3x3 block each); in place allowed.
__device__ __forceinline__ void qt_apply_pinv(const double* __restrict__ sm, const double* x, double* y, int lane) {
    if (lane < 13) y[lane] = sm[QK::oHd + lane] * x[lane];
    else if (lane < 21) {
        const int b = lane - 13;
        const double* h = sm + QK::oHb + 6 * b;
        const double x0 = x[13 + 3 * b], x1 = x[14 + 3 * b], x2 = x[15 + 3 * b];
        y[13 + 3 * b] = h[0] * x0 + h[1] * x1 + h[2] * x2;
        y[14 + 3 * b] = h[1] * x0 + h[3] * x1 + h[4] * x2;
        y[15 + 3 * b] = h[2] * x0 + h[4] * x1 + h[5] * x2;
    }
    __syncwarp();
}

// Dense "pattern" row: entries at the 31 columns (q0..q3, w0..w2, u0..u23) + one entry at column p_xc + one at column v_xc.
struct QtRow {
    double p[31];
    double xp, xv;
    int xc;
};
// Sparse row of U: p_xc, q0..q3, v_xc, w0..w2, r0..r2 of leg `leg`.
struct QtRowU {
    double xp, xv, q[4], w[3], r[3];
    int xc, leg;
};

// Row `lane` of V_j = [A_j (13 rows); Cp_{j+1} (16 rows)].  `smc`: the small part that holds Cp_{j+1} (null rows when `have_cp` is false).
__device__ __forceinline__ void qt_load_v_row(const double* __restrict__ ab, const double* __restrict__ smc, bool have_cp, int lane, QtRow& v) {
#pragma unroll
    for (int k = 0; k < 31; ++k) v.p[k] = 0.0;
    v.xp = 0.0; v.xv = 0.0; v.xc = 0;
    const int qr = lane >= 3 && lane < 7 ? lane - 3 : (lane >= 10 && lane < 13 ? lane - 6 : -1);
    if (qr >= 0) {
        const double* row = ab + qr * 32;
#pragma unroll
        for (int k = 0; k < 30; k += 2) { const double2 t2 = qp_ld2(row + k); v.p[k] = t2.x; v.p[k + 1] = t2.y; }
        v.p[30] = row[30];
    } else if (lane < 13) {
        const bool prow = lane < 3;
        const int c = prow ? lane : lane - 7;
        const double* ap = ab + 224 + (prow ? c : 3 + c) * 6;
        v.xc = c;
        v.xp = prow ? ap[0] : 0.0;
        v.xv = prow ? ap[1] : ap[0];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const double fl = prow ? ap[2 + l] : ap[1 + l];
#pragma unroll
            for (int m = 0; m < 3; ++m) v.p[7 + 6 * l + m] = m == c ? fl : 0.0;
        }
    } else if (lane < 29 && have_cp) {
        const int r = lane - 13, leg = r >> 2, rr = r & 3;
        if (rr > 0) {
            const double* cp = smc + QK::oCp + (leg * 3 + rr - 1) * 8;
            v.xc = rr - 1;
            v.xp = cp[0];
#pragma unroll
            for (int i = 0; i < 4; ++i) v.p[i] = cp[1 + i];
#pragma unroll
            for (int l = 0; l < 4; ++l)
#pragma unroll
                for (int m = 0; m < 3; ++m) v.p[7 + 6 * l + 3 + m] = l == leg ? cp[5 + m] : 0.0;
        }
    }
}

// v . t  (t = P^-1 q in place of q in `sm`)
__device__ __forceinline__ double qt_row_dot_t(const QtRow& v, const double* __restrict__ sm) {
    const double* t = sm + QK::oQ;
    double a0 = v.xp * t[v.xc], a1 = v.xv * t[7 + v.xc];
#pragma unroll
    for (int i = 0; i < 4; ++i) a0 += v.p[i] * t[3 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) a1 += v.p[4 + i] * t[10 + i];
#pragma unroll
    for (int n = 0; n < 24; n += 2) { a0 += v.p[7 + n] * t[13 + n]; a1 += v.p[8 + n] * t[14 + n]; }
    return a0 + a1;
}

// w = v P^-1 in place
__device__ __forceinline__ void qt_row_times_pinv(QtRow& v, const double* __restrict__ sm) {
    const double* pd = sm + QK::oHd;
    v.xp *= pd[v.xc];
    v.xv *= pd[7 + v.xc];
#pragma unroll
    for (int i = 0; i < 4; ++i) v.p[i] *= pd[3 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) v.p[4 + i] *= pd[10 + i];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const double* h = sm + QK::oHb + 6 * b;
        const double2 h01 = qp_ld2(h), h23 = qp_ld2(h + 2), h45 = qp_ld2(h + 4);
        const double x0 = v.p[7 + 3 * b], x1 = v.p[8 + 3 * b], x2 = v.p[9 + 3 * b];
        v.p[7 + 3 * b] = x0 * h01.x + x1 * h01.y + x2 * h23.x;
        v.p[8 + 3 * b] = x0 * h01.y + x1 * h23.y + x2 * h45.x;
        v.p[9 + 3 * b] = x0 * h23.x + x1 * h45.x + x2 * h45.y;
    }
}

// out[c] (+)= w . (row c of V),  c = 0 .. 28.  `smc` holds the Cp rows of V (have_cp false: they are null).
template <bool ADD>
__device__ __forceinline__ void qt_row_dot_v_rows(const QtRow& w, const double* __restrict__ ab, const double* __restrict__ smc, bool have_cp, double* out) {
#pragma unroll
    for (int c = 0; c < 29; ++c) {
        double acc;
        if ((c >= 3 && c < 7) || (c >= 10 && c < 13)) {
            const double* row = ab + (c < 7 ? c - 3 : c - 6) * 32;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int k = 0; k < 30; k += 2) { const double2 t2 = qp_ld2(row + k); a0 += w.p[k] * t2.x; a1 += w.p[k + 1] * t2.y; }
            acc = a0 + a1 + w.p[30] * row[30];
        } else if (c < 3) {
            const double* ap = ab + 224 + c * 6;
            const double2 d2 = qp_ld2(ap), f01 = qp_ld2(ap + 2), f23 = qp_ld2(ap + 4);
            acc = (w.xc == c ? w.xp * d2.x + w.xv * d2.y : 0.0) + (w.p[7 + c] * f01.x + w.p[13 + c] * f01.y) + (w.p[19 + c] * f23.x + w.p[25 + c] * f23.y);
        } else if (c < 10) {
            const int cc = c - 7;
            const double* ap = ab + 224 + (3 + cc) * 6;
            const double2 d2 = qp_ld2(ap), f12 = qp_ld2(ap + 2);
            acc = (w.xc == cc ? w.xv * d2.x : 0.0) + (w.p[7 + cc] * d2.y + w.p[13 + cc] * f12.x) + (w.p[19 + cc] * f12.y + w.p[25 + cc] * ap[4]);
        } else {
            const int rc = c - 13, legc = rc >> 2, rrc = rc & 3;
            if (rrc == 0) acc = 0.0;
            else {
                const double* cp = smc + QK::oCp + (legc * 3 + rrc - 1) * 8;
                const double2 c01 = qp_ld2(cp), c23 = qp_ld2(cp + 2), c45 = qp_ld2(cp + 4), c67 = qp_ld2(cp + 6);
                acc = (w.xc == rrc - 1 ? w.xp * c01.x : 0.0) + (w.p[0] * c01.y + w.p[1] * c23.x) + (w.p[2] * c23.y + w.p[3] * c45.x) +
                      (w.p[7 + 6 * legc + 3] * c45.y + w.p[7 + 6 * legc + 4] * c67.x + w.p[7 + 6 * legc + 5] * c67.y);
                acc = have_cp ? acc : 0.0;
            }
        }
        if (ADD) out[c] += acc; else out[c] = acc;
    }
}

// e[c] = w . (row c of U_j),  U_j = [I 0; Cs_j]
__device__ __forceinline__ void qt_row_dot_u_rows(const QtRow& w, const double* __restrict__ sm, double* e) {
#pragma unroll
    for (int c = 13; c < 29; ++c) {
        const int rc = c - 13, legc = rc >> 2, rrc = rc & 3, pcc = rrc == 0 ? 2 : rrc - 1;
        const double* cs = sm + QK::oCs + (legc * 4 + rrc) * 8;
        const double2 c01 = qp_ld2(cs), c23 = qp_ld2(cs + 2), c45 = qp_ld2(cs + 4), c67 = qp_ld2(cs + 6);
        e[c] = (w.xc == pcc ? w.xp * c01.x : 0.0) + (w.p[0] * c01.y + w.p[1] * c23.x) + (w.p[2] * c23.y + w.p[3] * c45.x) +
               (w.p[7 + 6 * legc + 3] * c45.y + w.p[7 + 6 * legc + 4] * c67.x + w.p[7 + 6 * legc + 5] * c67.y);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { e[c] = w.xc == c ? w.xp : 0.0; e[7 + c] = w.xc == c ? w.xv : 0.0; e[10 + c] = w.p[4 + c]; }
#pragma unroll
    for (int c = 0; c < 4; ++c) e[3 + c] = w.p[c];
}

// Row `lane` of U_j in sparse form (state lanes: unit vectors; contact lanes: the Cs row; lanes >= 29: null).
__device__ __forceinline__ void qt_load_u_row(const double* __restrict__ sm, int lane, QtRowU& u) {
    u.xp = 0.0; u.xv = 0.0; u.xc = 0; u.leg = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) u.q[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { u.w[i] = 0.0; u.r[i] = 0.0; }
    if (lane < 13) {
        if (lane < 3) { u.xp = 1.0; u.xc = lane; }
        else if (lane < 7) {
#pragma unroll
            for (int i = 0; i < 4; ++i) u.q[i] = lane - 3 == i ? 1.0 : 0.0;
        } else if (lane < 10) { u.xv = 1.0; u.xc = lane - 7; }
        else {
#pragma unroll
            for (int i = 0; i < 3; ++i) u.w[i] = lane - 10 == i ? 1.0 : 0.0;
        }
    } else if (lane < 29) {
        const int r = lane - 13, leg = r >> 2, rr = r & 3;
        const double* cs = sm + QK::oCs + (leg * 4 + rr) * 8;
        u.leg = leg;
        u.xc = rr == 0 ? 2 : rr - 1;
        u.xp = cs[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) u.q[i] = cs[1 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) u.r[i] = cs[5 + i];
    }
}

__device__ __forceinline__ double qt_urow_dot(const QtRowU& u, const double* __restrict__ x) {  // u . x for a 37-vector x in shared memory
    double a = u.xp * x[u.xc] + u.xv * x[7 + u.xc];
#pragma unroll
    for (int i = 0; i < 4; ++i) a += u.q[i] * x[3 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) a += u.w[i] * x[10 + i] + u.r[i] * x[13 + 6 * u.leg + 3 + i];
    return a;
}

__device__ __forceinline__ void qt_urow_times_pinv(QtRowU& u, const double* __restrict__ sm) {
    const double* pd = sm + QK::oHd;
    u.xp *= pd[u.xc];
    u.xv *= pd[7 + u.xc];
#pragma unroll
    for (int i = 0; i < 4; ++i) u.q[i] *= pd[3 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) u.w[i] *= pd[10 + i];
    const double* h = sm + QK::oHb + 6 * (2 * u.leg + 1);
    const double x0 = u.r[0], x1 = u.r[1], x2 = u.r[2];
    u.r[0] = x0 * h[0] + x1 * h[1] + x2 * h[2];
    u.r[1] = x0 * h[1] + x1 * h[3] + x2 * h[4];
    u.r[2] = x0 * h[2] + x1 * h[4] + x2 * h[5];
}

// out[c] += wu . (row c of U_j)
__device__ __forceinline__ void qt_urow_dot_u_rows(const QtRowU& wu, const double* __restrict__ sm, double* out) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { out[c] += wu.xc == c ? wu.xp : 0.0; out[7 + c] += wu.xc == c ? wu.xv : 0.0; out[10 + c] += wu.w[c]; }
#pragma unroll
    for (int c = 0; c < 4; ++c) out[3 + c] += wu.q[c];
#pragma unroll
    for (int c = 13; c < 29; ++c) {
        const int rc = c - 13, legc = rc >> 2, rrc = rc & 3, pcc = rrc == 0 ? 2 : rrc - 1;
        const double* cs = sm + QK::oCs + (legc * 4 + rrc) * 8;
        const double2 c01 = qp_ld2(cs), c23 = qp_ld2(cs + 2), c45 = qp_ld2(cs + 4), c67 = qp_ld2(cs + 6);
        const double own = wu.r[0] * c45.y + wu.r[1] * c67.x + wu.r[2] * c67.y;
        out[c] += (wu.xc == pcc ? wu.xp * c01.x : 0.0) + (wu.q[0] * c01.y + wu.q[1] * c23.x) + (wu.q[2] * c23.y + wu.q[3] * c45.x) + (wu.leg == legc ? own : 0.0);
    }
}

// e[c] = wu . (row c of V)
__device__ __forceinline__ void qt_urow_dot_v_rows(const QtRowU& wu, const double* __restrict__ ab, const double* __restrict__ smc, bool have_cp, double* e) {
    const int ro = 7 + 6 * wu.leg + 3;
#pragma unroll
    for (int c = 0; c < 29; ++c) {
        double acc;
        if ((c >= 3 && c < 7) || (c >= 10 && c < 13)) {
            const double* row = ab + (c < 7 ? c - 3 : c - 6) * 32;
            const double2 r01 = qp_ld2(row), r23 = qp_ld2(row + 2), r45 = qp_ld2(row + 4);
            acc = (wu.q[0] * r01.x + wu.q[1] * r01.y) + (wu.q[2] * r23.x + wu.q[3] * r23.y) + (wu.w[0] * r45.x + wu.w[1] * r45.y) + wu.w[2] * row[6] +
                  (wu.r[0] * row[ro] + wu.r[1] * row[ro + 1] + wu.r[2] * row[ro + 2]);
        } else if (c < 3) {
            const double2 d2 = qp_ld2(ab + 224 + c * 6);
            acc = wu.xc == c ? wu.xp * d2.x + wu.xv * d2.y : 0.0;
        } else if (c < 10) {
            acc = wu.xc == c - 7 ? wu.xv * ab[224 + (3 + c - 7) * 6] : 0.0;
        } else {
            const int rc = c - 13, legc = rc >> 2, rrc = rc & 3;
            if (rrc == 0) acc = 0.0;
            else {
                const double* cp = smc + QK::oCp + (legc * 3 + rrc - 1) * 8;
                const double2 c01 = qp_ld2(cp), c23 = qp_ld2(cp + 2), c45 = qp_ld2(cp + 4), c67 = qp_ld2(cp + 6);
                const double own = wu.r[0] * c45.y + wu.r[1] * c67.x + wu.r[2] * c67.y;
                acc = (wu.xc == rrc - 1 ? wu.xp * c01.x : 0.0) + (wu.q[0] * c01.y + wu.q[1] * c23.x) + (wu.q[2] * c23.y + wu.q[3] * c45.x) + (wu.leg == legc ? own : 0.0);
                acc = have_cp ? acc : 0.0;
            }
        }
        e[c] = acc;
    }
}

// S -= Lo Lo^T on the FP64 tensor cores.  The image (29 rows, stride LS) holds Lo; it is overwritten by the product, of which every
// lane then subtracts its own row.  Ten lower 8x8 tiles x eight k-steps of mma.m8n8k4 (rows / k beyond 28 are masked).
__device__ __forceinline__ void qt_syrk(double* __restrict__ img, double* s, int lane) {
    constexpr int G = QpT::G, LS = QpT::LS;
    double acc[10][2];
#pragma unroll
    for (int t = 0; t < 10; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    const int fr = lane >> 2, fq = lane & 3;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        double f[4];
#pragma unroll
        for (int I = 0; I < 4; ++I) {
            f[I] = img[min(8 * I + fr, G - 1) * LS + min(4 * kk + fq, G - 1)];
            if (kk == 7 && fq != 0) f[I] = 0.0;
        }
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J <= I; ++J) qp_dmma(acc[(I * (I + 1)) / 2 + J][0], acc[(I * (I + 1)) / 2 + J][1], f[I], f[J]);
    }
    __syncwarp();
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) {
            const int row = 8 * I + fr, col = 8 * J + 2 * fq;
            if (row < G && col < LS) qp_st2(img + row * LS + col, acc[(I * (I + 1)) / 2 + J][0], acc[(I * (I + 1)) / 2 + J][1]);
        }
    __syncwarp();
    const double* mine = img + min(lane, G - 1) * LS;
#pragma unroll
    for (int c = 0; c + 1 < G; c += 2) { const double2 t2 = qp_ld2(mine + c); s[c] -= t2.x; s[c + 1] -= t2.y; }
    s[G - 1] -= mine[G - 1];
    __syncwarp();
}

// Cholesky of the block held row-per-lane in s[] (lower triangle), fused with y = L^-1 rhs and — when TRSM — with the right-looking
// solve e <- e L^-T of the NEXT group's coupling rows.  Column c of L goes through a 2-slot shared buffer once and serves both
// updates; it leaves for the workspace with one coalesced store.  Returns y (entry `lane`); inv_own = 1 / L_ii of the own row.
template <bool TRSM>
__device__ __forceinline__ double qt_cholesky(double* s, double* e, double rhs, double* __restrict__ col2, double* __restrict__ wsg, int lane, double& inv_own) {
    constexpr int G = QpT::G;
    constexpr unsigned FULL = 0xffffffffu;
    double rk = rhs;
    inv_own = 0.0;
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double rinv = rsqrt(__shfl_sync(FULL, s[c], c));
        const double l = lane >= c ? s[c] * rinv : 0.0;
        s[c] = l;
        if (lane == c) inv_own = rinv;
        double* col = col2 + (c & 1) * 32;
        col[lane] = l;
        if (lane >= c && lane < G) wsg[QpT::colstart(c) + lane - c] = l;
        const double yc = __shfl_sync(FULL, rk, c) * rinv;
        if (lane == c) rk = yc;
        else if (lane > c) rk -= l * yc;
        double ec = 0.0;
        if (TRSM) { ec = e[c] * rinv; e[c] = ec; }
        __syncwarp();
        if (c + 1 < G) {
            if ((c + 1) & 1) {
                const double lc = col[c + 1];
                s[c + 1] -= l * lc;
                if (TRSM) e[c + 1] -= ec * lc;
            }
#pragma unroll
            for (int cc = (c + 2) & ~1; cc < G; cc += 2) {
                const double2 t2 = qp_ld2(col + cc);
                s[cc] -= l * t2.x;
                if (TRSM) e[cc] -= ec * t2.x;
                if (cc + 1 < G) {
                    s[cc + 1] -= l * t2.y;
                    if (TRSM) e[cc + 1] -= ec * t2.y;
                }
            }
        }
    }
    return rk;
}

// (V_j^T nu)[k] for the local variable k = 0..36: nu = multipliers of group j+1 (29 entries in shared memory); `smc` holds Cp_{j+1}.
__device__ __forceinline__ double qt_vT_nu(const double* __restrict__ ab, const double* __restrict__ smc, bool have_cp, const double* __restrict__ nu, int k) {
    double acc = 0.0;
    const int col = k < 3 ? -1 : k < 7 ? k - 3 : k < 10 ? -1 : k - 6;
    if (col >= 0) {
#pragma unroll
        for (int qr = 0; qr < 7; ++qr) acc += ab[qr * 32 + col] * nu[qr < 4 ? 3 + qr : 6 + qr];
    }
    const double* ap = ab + 224;
    if (k < 3) acc += ap[k * 6] * nu[k];
    else if (k >= 7 && k < 10) acc += ap[(k - 7) * 6 + 1] * nu[k - 7] + ap[(3 + k - 7) * 6] * nu[k];
    else if (k >= 13) {
        const int n = k - 13, l = n / 6, mm = n - 6 * l;
        if (mm < 3) acc += ap[mm * 6 + 2 + l] * nu[mm] + ap[(3 + mm) * 6 + 1 + l] * nu[7 + mm];
        else if (have_cp) {
#pragma unroll
            for (int r1 = 0; r1 < 3; ++r1) acc += smc[QK::oCp + (l * 3 + r1) * 8 + 5 + mm - 3] * nu[13 + 4 * l + r1 + 1];
        }
    }
    if (have_cp) {
        if (k < 3) {
#pragma unroll
            for (int l = 0; l < 4; ++l) acc += smc[QK::oCp + (l * 3 + k) * 8] * nu[13 + 4 * l + k + 1];
        } else if (k < 7) {
#pragma unroll
            for (int l = 0; l < 4; ++l)
#pragma unroll
                for (int r1 = 0; r1 < 3; ++r1) acc += smc[QK::oCp + (l * 3 + r1) * 8 + 1 + k - 3] * nu[13 + 4 * l + r1 + 1];
        }
    }
    return acc;
}

// (U_j^T nu)[k]: nu = multipliers of group j; U_j = [I 0; Cs_j].
__device__ __forceinline__ double qt_uT_nu(const double* __restrict__ sm, bool have_cs, const double* __restrict__ nu, int k) {
    double acc = k < 13 ? nu[k] : 0.0;
    if (!have_cs) return acc;
    const double* cs = sm + QK::oCs;
    if (k < 3) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            acc += cs[(l * 4 + k + 1) * 8] * nu[13 + 4 * l + k + 1];
            if (k == 2) acc += cs[(l * 4) * 8] * nu[13 + 4 * l];
        }
    } else if (k < 7) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) acc += cs[(l * 4 + rr) * 8 + 1 + k - 3] * nu[13 + 4 * l + rr];
    } else if (k >= 13) {
        const int n = k - 13, l = n / 6, mm = n - 6 * l;
        if (mm >= 3) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) acc += cs[(l * 4 + rr) * 8 + 5 + mm - 3] * nu[13 + 4 * l + rr];
        }
    }
    return acc;
}

// Solves L z = x then L^T nu = y - z with the packed-column image of L (`lp`: columns, then 1 / L_ii at wsInv, y at wsY): returns nu_lane.
__device__ __forceinline__ double qt_outward_solve(const double* __restrict__ lp, double x, int lane) {
    constexpr int G = QpT::G;
    constexpr unsigned FULL = 0xffffffffu;
    const int ln = min(lane, G - 1);
    double rk = x;
#pragma unroll
    for (int i = 0; i < G; ++i) {  // forward: lane > i needs L[lane][i]
        const double zi = __shfl_sync(FULL, rk, i) * lp[QpT::wsInv + i];
        const double lv = lp[QpT::colstart(i) + max(ln - i, 0)];
        if (lane == i) rk = zi;
        else if (lane > i) rk -= lv * zi;
    }
    rk = (lane < G ? lp[QpT::wsY + ln] : 0.0) - rk;
    const int cs = QpT::colstart(ln);
#pragma unroll
    for (int i = G - 1; i >= 0; --i) {  // backward with L^T: lane < i needs L[i][lane] = column `lane`, row i
        const double ni = __shfl_sync(FULL, rk, i) * lp[QpT::wsInv + i];
        const double lv = lp[cs + max(i - ln, 0)];
        if (lane == i) rk = ni;
        else if (lane < i) rk -= lv * ni;
    }
    return lane < G ? rk : 0.0;
}

// =================================================================================================================================
// One CTA of two warps per trajectory.  The two chains are separate (non-inlined) functions so that each gets its own register
// allocation: inlined into one kernel, the union of the two paths spilled ~1.5 KB per thread.
// =================================================================================================================================
__device__ __noinline__ void qt_top_chain(const double* __restrict__ rec_all, long long ld_rec, double* __restrict__ ws_all, double* __restrict__ step_all, long long ld_step,
        double* __restrict__ mult_all, long long ld_mult, int N, double delta, long long b, int lane, int wib, unsigned char* smem_raw) {
    using Q = QpT;
    constexpr int G = Q::G, LS = Q::LS;
    constexpr unsigned SMALL_B = QK::SMALL * 8, APART_B = QK::APART * 8, WS_B = Q::WS_GROUP * 8;
    double* const cta = reinterpret_cast<double*>(smem_raw);
    double* const sm  = cta + wib * Q::PER_WARP;
    double* const img = sm + Q::oLO;
    double* const sY  = sm + Q::oY;
    double* const sC  = sm + Q::oCOL;
    double* const other_img = cta + (wib ^ 1) * Q::PER_WARP + Q::oLO;
    double* const other_y   = cta + (wib ^ 1) * Q::PER_WARP + Q::oY;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + Q::oBAR);
    QtBuf sa{sm + Q::oSM0, bars + 0, 0u}, sb{sm + Q::oSM1, bars + 1, 0u}, abuf{sm + Q::oAB, bars + 2, 0u}, lbuf{img, bars + 3, 0u};

    const double* __restrict__ rec  = rec_all + b * ld_rec;
    const double* __restrict__ tail = rec + QK::tail(N);
    double* __restrict__ ws   = ws_all + b * (long long)(N + 1) * Q::WS_GROUP;
    double* __restrict__ step = step_all + b * ld_step;
    double* __restrict__ mult = mult_all ? mult_all + b * ld_mult : nullptr;
    const int nX = 13 * (N + 1);
    const int m  = N / 2;  // the groups meet here
    const bool act = lane < G, st = lane < 13;
    auto chunk = [&](int j) { return rec + (long long)j * QK::NODE; };
    auto swap_bufs = [&]() { const QtBuf t = sa; sa = sb; sb = t; };

    double s[G], e[G];
#pragma unroll
    for (int c = 0; c < G; ++c) { s[c] = 0.0; e[c] = 0.0; }
    double rpart = 0.0, inv_own;
    // ======================================================== top-down: groups 0 .. m-1
    sa.load(chunk(0), SMALL_B, lane);
    sb.load(chunk(1), SMALL_B, lane);
    abuf.load(chunk(0) + QK::oAQ, APART_B, lane);
    double gdef = st ? tail[QK::tG0 + lane] : 0.0;  // x_0 - x_measured
    sa.wait();
    for (int j = 0; j < m; ++j) {  // sa = small_j (landed), sb = small_{j+1} and abuf = A_j (in flight)
        qt_pinv_t(sa.buf, lane, true);
        {   // S_jj += U P^-1 U^T + delta I ;  rhs = g - U t + rpart
            QtRowU u;
            qt_load_u_row(sa.buf, lane, u);
            const double ut = qt_urow_dot(u, sa.buf + QK::oQ);
            rpart += (st ? gdef : (act ? sa.buf[QK::oG + lane] : 0.0)) - ut;
            qt_urow_times_pinv(u, sa.buf);
            qt_urow_dot_u_rows(u, sa.buf, s);
#pragma unroll
            for (int c = 0; c < G; ++c) s[c] += c == lane ? delta : 0.0;
        }
        if (j > 0) qt_syrk(img, s, lane);
        sb.wait();
        abuf.wait();
        double carry;
        {   // coupling rows of group j+1: e = (V P^-1 U^T) row
            QtRow v;
            qt_load_v_row(abuf.buf, sb.buf, true, lane, v);
            carry = qt_row_dot_t(v, sa.buf);
            qt_row_times_pinv(v, sa.buf);
            qt_row_dot_u_rows(v, sa.buf, e);
        }
        double* wsg = ws + (long long)j * Q::WS_GROUP;
        const double yj = qt_cholesky<true>(s, e, rpart, sC, wsg, lane, inv_own);
        if (act) {
            sY[lane] = yj;
            wsg[Q::wsY + lane] = yj;
            wsg[Q::wsInv + lane] = inv_own;
#pragma unroll
            for (int c = 0; c + 1 < G; c += 2) qp_st2(img + lane * LS + c, e[c], e[c + 1]);
            qp_st2(img + lane * LS + G - 1, e[G - 1], 0.0);
        }
        __syncwarp();
        {   // rpart of group j+1 = -Lo y_j - V t_j
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int c = 0; c + 1 < G; c += 2) { const double2 t2 = qp_ld2(sY + c); d0 += e[c] * t2.x; d1 += e[c + 1] * t2.y; }
            d0 += e[G - 1] * sY[G - 1];
            rpart = -(d0 + d1) - carry;
        }
        {   // S_{j+1,j+1} part: V P^-1 V^T row
            QtRow v;
            qt_load_v_row(abuf.buf, sb.buf, true, lane, v);
            qt_row_times_pinv(v, sa.buf);
            qt_row_dot_v_rows<false>(v, abuf.buf, sb.buf, true, s);
        }
        gdef = st ? sa.buf[QK::oG + lane] : 0.0;  // defect of stage j: the state rows of group j+1
        __syncwarp();
        swap_bufs();  // sa = small_{j+1}
        if (j + 1 < m) {
            sb.load(chunk(j + 2), SMALL_B, lane);
            abuf.load(chunk(j + 1) + QK::oAQ, APART_B, lane);
        }
    }
    // ---- middle group m: top part S = V P^-1 V^T - Lo Lo^T, rhs = defect - V t - Lo y
    qt_syrk(img, s, lane);
    rpart += st ? gdef : 0.0;
    __syncthreads();  // the bottom warp has parked its part of S_mm and of the right-hand side in its image
    if (act) {
        const double* mine = other_img + lane * LS;
#pragma unroll
        for (int c = 0; c + 1 < G; c += 2) { const double2 t2 = qp_ld2(mine + c); s[c] += t2.x; s[c + 1] += t2.y; }
        const double2 t2 = qp_ld2(mine + G - 1);
        s[G - 1] += t2.x;
        rpart += t2.y;
    }
    double* wsg = ws + (long long)m * Q::WS_GROUP;
    const double ym = qt_cholesky<false>(s, e, rpart, sC, wsg, lane, inv_own);
    // nu_m = L^-T y_m from the rows in registers (column access = across lanes: one warp reduction per entry)
    double num = 0.0;
#pragma unroll
    for (int i = G - 1; i >= 0; --i) {
        double part = lane > i && act ? s[i] * num : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        const double yi = __shfl_sync(0xffffffffu, ym, i), ii = __shfl_sync(0xffffffffu, inv_own, i);
        if (lane == i) num = (yi - part) * ii;
    }
    if (act) {
        sY[lane] = num;
        other_y[lane] = num;
        if (mult) {
            if (st) mult[13 * m + lane] = num;
            else mult[nX + 16 * m + (lane - 13)] = num;
        }
    } else { sY[lane] = 0.0; other_y[lane] = 0.0; }
    __syncthreads();  // nu_m is visible to the bottom warp

    // ======================================================== outward: groups m-1 .. 0
    double* const vA = img + Q::oVA;
    double* const vC = img + Q::oVC;
    sb.load(chunk(m), SMALL_B, lane);  // Cp_m
    sb.wait();
    for (int j = m - 1; j >= 0; --j) {
        sa.load(chunk(j), SMALL_B, lane);
        abuf.load(chunk(j) + QK::oAQ, APART_B, lane);
        lbuf.load(ws + (long long)j * Q::WS_GROUP, WS_B, lane);
        sa.wait();
        abuf.wait();
        qt_pinv_t(sa.buf, lane, true);
        for (int k = lane; k < 37; k += 32) vA[k] = qt_vT_nu(abuf.buf, sb.buf, true, sY, k);  // a = V_j^T nu_{j+1}
        __syncwarp();
        qt_apply_pinv(sa.buf, vA, vC, lane);
        QtRowU u;
        qt_load_u_row(sa.buf, lane, u);
        const double z = qt_urow_dot(u, vC);  // (U P^-1 a) row
        lbuf.wait();
        const double nu = qt_outward_solve(img, z, lane);
        __syncwarp();
        sY[lane] = nu;
        __syncwarp();
        for (int k = lane; k < 37; k += 32) vA[k] += qt_uT_nu(sa.buf, true, sY, k);
        __syncwarp();
        qt_apply_pinv(sa.buf, vA, vC, lane);
        for (int k = lane; k < 37; k += 32) {
            const long long dst = k < 13 ? 13 * j + k : nX + 24 * j + (k - 13);
            step[dst] = -(sa.buf[QK::oQ + k] + vC[k]);  // d_j = -(t_j + P^-1 (a + U^T nu_j))
        }
        if (mult && act) {
            if (st) mult[13 * j + lane] = nu;
            else mult[nX + 16 * j + (lane - 13)] = nu;
        }
        __syncwarp();
        swap_bufs();  // sb = small_j: group j-1 needs its Cp rows
    }
}

__device__ __noinline__ void qt_bottom_chain(const double* __restrict__ rec_all, long long ld_rec, double* __restrict__ ws_all, double* __restrict__ step_all, long long ld_step,
        double* __restrict__ mult_all, long long ld_mult, int N, double delta, long long b, int lane, int wib, unsigned char* smem_raw) {
    using Q = QpT;
    constexpr int G = Q::G, LS = Q::LS;
    constexpr unsigned SMALL_B = QK::SMALL * 8, APART_B = QK::APART * 8, WS_B = Q::WS_GROUP * 8;
    double* const cta = reinterpret_cast<double*>(smem_raw);
    double* const sm  = cta + wib * Q::PER_WARP;
    double* const img = sm + Q::oLO;
    double* const sY  = sm + Q::oY;
    double* const sC  = sm + Q::oCOL;
    double* const other_img = cta + (wib ^ 1) * Q::PER_WARP + Q::oLO;
    double* const other_y   = cta + (wib ^ 1) * Q::PER_WARP + Q::oY;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + Q::oBAR);
    QtBuf sa{sm + Q::oSM0, bars + 0, 0u}, sb{sm + Q::oSM1, bars + 1, 0u}, abuf{sm + Q::oAB, bars + 2, 0u}, lbuf{img, bars + 3, 0u};

    const double* __restrict__ rec  = rec_all + b * ld_rec;
    const double* __restrict__ tail = rec + QK::tail(N);
    double* __restrict__ ws   = ws_all + b * (long long)(N + 1) * Q::WS_GROUP;
    double* __restrict__ step = step_all + b * ld_step;
    double* __restrict__ mult = mult_all ? mult_all + b * ld_mult : nullptr;
    const int nX = 13 * (N + 1);
    const int m  = N / 2;  // the groups meet here
    const bool act = lane < G, st = lane < 13;
    auto chunk = [&](int j) { return rec + (long long)j * QK::NODE; };
    auto swap_bufs = [&]() { const QtBuf t = sa; sa = sb; sb = t; };

    double s[G], e[G];
#pragma unroll
    for (int c = 0; c < G; ++c) { s[c] = 0.0; e[c] = 0.0; }
    double rpart = 0.0, inv_own;
    // ======================================================== bottom-up: groups N .. m+1
    // sa = small_j (synthesised for j = N: no inputs, no contact rows), sb = small_{j-1}, abuf = A_{j-1}
    for (int k = lane; k < QK::SMALL; k += 32) sa.buf[k] = 0.0;
    __syncwarp();
    if (st) {
        sa.buf[QK::oQ + lane]  = tail[QK::tQN + lane];
        sa.buf[QK::oHd + lane] = tail[QK::tHN + lane];
    }
    __syncwarp();
    sb.load(chunk(N - 1), SMALL_B, lane);
    abuf.load(chunk(N - 1) + QK::oAQ, APART_B, lane);
    qt_pinv_t(sa.buf, lane, false);
    for (int j = N; j > m; --j) {
        const bool last = j == N;  // group N: 13 rows, no contact rows
        {   // S_jj = U P^-1 U^T + delta I ;  rhs = contact values - U t + rpart
            QtRowU u;
            qt_load_u_row(sa.buf, lane, u);
            if (last && !st) { u.xp = 0.0; u.q[0] = u.q[1] = u.q[2] = u.q[3] = 0.0; u.r[0] = u.r[1] = u.r[2] = 0.0; }
            const double ut = qt_urow_dot(u, sa.buf + QK::oQ);
            rpart += (!st && act ? sa.buf[QK::oG + lane] : 0.0) - ut;
            qt_urow_times_pinv(u, sa.buf);
#pragma unroll
            for (int c = 0; c < G; ++c) s[c] = c == lane ? delta : 0.0;
            qt_urow_dot_u_rows(u, sa.buf, s);
        }
        if (!last) qt_syrk(img, s, lane);
        sb.wait();
        abuf.wait();
        qt_pinv_t(sb.buf, lane, true);
        rpart += st ? sb.buf[QK::oG + lane] : 0.0;  // defect of stage j-1
        {   // S_jj += V P^-1 V^T ;  rhs -= V t_{j-1}
            QtRow v;
            qt_load_v_row(abuf.buf, sa.buf, !last, lane, v);
            rpart -= qt_row_dot_t(v, sb.buf);
            qt_row_times_pinv(v, sb.buf);
            qt_row_dot_v_rows<true>(v, abuf.buf, sa.buf, !last, s);
        }
        {   // coupling rows of group j-1: e = (U_{j-1} P^-1 V^T) row
            QtRowU u;
            qt_load_u_row(sb.buf, lane, u);
            qt_urow_times_pinv(u, sb.buf);
            qt_urow_dot_v_rows(u, abuf.buf, sa.buf, !last, e);
        }
        __syncwarp();
        if (j - 2 >= m) abuf.load(chunk(j - 2) + QK::oAQ, APART_B, lane);
        double* wsg = ws + (long long)j * Q::WS_GROUP;
        const double yj = qt_cholesky<true>(s, e, rpart, sC, wsg, lane, inv_own);
        if (act) {
            sY[lane] = yj;
            wsg[Q::wsY + lane] = yj;
            wsg[Q::wsInv + lane] = inv_own;
#pragma unroll
            for (int c = 0; c + 1 < G; c += 2) qp_st2(img + lane * LS + c, e[c], e[c + 1]);
            qp_st2(img + lane * LS + G - 1, e[G - 1], 0.0);
        }
        __syncwarp();
        {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int c = 0; c + 1 < G; c += 2) { const double2 t2 = qp_ld2(sY + c); d0 += e[c] * t2.x; d1 += e[c + 1] * t2.y; }
            d0 += e[G - 1] * sY[G - 1];
            rpart = -(d0 + d1);
        }
        __syncwarp();
        swap_bufs();  // sa = small_{j-1} (P^-1 and t already in place)
        if (j - 2 >= m) sb.load(chunk(j - 2), SMALL_B, lane);
    }
    // ---- middle group m, bottom part: U P^-1 U^T + delta I - Uo Uo^T ; rhs = contact values - U t - Uo y   (sa = small_m)
    {
        QtRowU u;
        qt_load_u_row(sa.buf, lane, u);
        const double ut = qt_urow_dot(u, sa.buf + QK::oQ);
        rpart += (!st && act ? sa.buf[QK::oG + lane] : 0.0) - ut;
        qt_urow_times_pinv(u, sa.buf);
#pragma unroll
        for (int c = 0; c < G; ++c) s[c] = c == lane ? delta : 0.0;
        qt_urow_dot_u_rows(u, sa.buf, s);
    }
    qt_syrk(img, s, lane);
    if (act) {
#pragma unroll
        for (int c = 0; c + 1 < G; c += 2) qp_st2(img + lane * LS + c, s[c], s[c + 1]);
        qp_st2(img + lane * LS + G - 1, s[G - 1], rpart);
    }
    __syncthreads();  // parked
    __syncthreads();  // nu_m has arrived in sY

    // ======================================================== outward: groups m+1 .. N (and the steps d_m .. d_N)
    // invariant at group j: sa = small_{j-1} (P^-1, t in place), abuf = A_{j-1}, sY = nu_{j-1}
    double* const vA = img + Q::oVA;
    double* const vC = img + Q::oVC;
    double* const nu2 = img + Q::oNU2;
    abuf.load(chunk(m) + QK::oAQ, APART_B, lane);
    abuf.wait();
    for (int j = m + 1; j <= N; ++j) {
        const bool last = j == N;
        lbuf.load(ws + (long long)j * Q::WS_GROUP, WS_B, lane);
        if (!last) {
            sb.load(chunk(j), SMALL_B, lane);
            sb.wait();
        }
        for (int k = lane; k < 37; k += 32) vA[k] = qt_uT_nu(sa.buf, true, sY, k);  // b = U_{j-1}^T nu_{j-1}
        __syncwarp();
        qt_apply_pinv(sa.buf, vA, vC, lane);
        double z;
        {   // (V_{j-1} P^-1 b) row
            QtRow v;
            qt_load_v_row(abuf.buf, sb.buf, !last, lane, v);
            const double* x = vC;
            double a0 = v.xp * x[v.xc], a1 = v.xv * x[7 + v.xc];
#pragma unroll
            for (int i = 0; i < 4; ++i) a0 += v.p[i] * x[3 + i];
#pragma unroll
            for (int i = 0; i < 3; ++i) a1 += v.p[4 + i] * x[10 + i];
#pragma unroll
            for (int n = 0; n < 24; n += 2) { a0 += v.p[7 + n] * x[13 + n]; a1 += v.p[8 + n] * x[14 + n]; }
            z = a0 + a1;
        }
        lbuf.wait();
        const double nu = qt_outward_solve(img, z, lane);
        __syncwarp();
        nu2[lane] = nu;
        __syncwarp();
        for (int k = lane; k < 37; k += 32) vA[k] += qt_vT_nu(abuf.buf, sb.buf, !last, nu2, k);
        __syncwarp();
        qt_apply_pinv(sa.buf, vA, vC, lane);
        for (int k = lane; k < 37; k += 32) {
            const long long dst = k < 13 ? 13 * (j - 1) + k : nX + 24 * (j - 1) + (k - 13);
            step[dst] = -(sa.buf[QK::oQ + k] + vC[k]);  // d_{j-1}
        }
        if (mult && act) {
            if (st) mult[13 * j + lane] = nu;
            else if (!last) mult[nX + 16 * j + (lane - 13)] = nu;
        }
        sY[lane] = nu;
        __syncwarp();
        if (!last) {
            swap_bufs();  // sa = small_j
            qt_pinv_t(sa.buf, lane, true);
            abuf.load(chunk(j) + QK::oAQ, APART_B, lane);
            abuf.wait();
        }
    }
    // d_N = -P_N^-1 (q_N + nu_N)
    if (st) step[13 * N + lane] = -(tail[QK::tQN + lane] + sY[lane]) / tail[QK::tHN + lane];
}

__global__ void __maxnreg__(144)
qp_twisted_kernel(const double* __restrict__ rec_all, long long ld_rec, double* __restrict__ ws_all, double* __restrict__ step_all, long long ld_step,
                  double* __restrict__ mult_all, long long ld_mult, int N, long long batch, double delta, const int* __restrict__ skip_status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long b = blockIdx.x;
    if (b >= batch) return;
    if (skip_status && skip_status[2 * b] != 0) return;  // SQP loop: this trajectory has stopped (CTA-uniform)
    {
        uint64_t* const bars = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem_raw) + wib * QpT::PER_WARP + QpT::oBAR);
        if (lane < 4) qt_bar_init(bars + lane);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    if (wib == 0) qt_top_chain(rec_all, ld_rec, ws_all, step_all, ld_step, mult_all, ld_mult, N, delta, b, lane, wib, smem_raw);
    else qt_bottom_chain(rec_all, ld_rec, ws_all, step_all, ld_step, mult_all, ld_mult, N, delta, b, lane, wib, smem_raw);
}

}  // namespace ub
