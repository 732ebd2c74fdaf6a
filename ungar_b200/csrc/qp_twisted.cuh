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
//     64-thread CTA; the two dependent chains are half as long and run concurrently (7 CTAs = 14 warps per SM: the 1024
//     trajectories of the headline configuration are one wave); they meet at group m = N / 2 and the substitutions run outward
//     from there, again one direction per warp.
//   * the record is the compact one: a stage is two TMA bulk loads (2832 B + 2080 B, cp.async.bulk + mbarrier) instead of ~90
//     8-byte cp.async with index arithmetic, and 2.6x fewer bytes.
//   * registers: a lane owns ROW i of every 29-row block, but only ONE 29-entry row is live at a time (S during the Cholesky
//     sweep, then the coupling row during the triangular solve); the operands V P^-1 are formed in pieces (state part + one leg
//     at a time).  128 registers, no spills (round 1: 255 with 319 k local loads).
//   * the Cholesky sweep keeps the columns UNSCALED in shared memory (column c = S'[:, c], pivot on top): no pivot shuffle, no
//     per-column selects; 1 / sqrt(pivot) is folded into the scalar every update is multiplied with.  The same image serves the
//     triangular solve of the next group's coupling rows and leaves for the workspace with one coalesced store per column.
// FP64 tensor cores (mma.sync.m8n8k4.f64) do the rank-29 update S -= Lo Lo^T.
#pragma once

#include <cstdint>

#include "compact.cuh"
#ifndef QT_HOST_TEST  // tests/host/*.cpp compile the arithmetic below for the CPU and supply qp_ld2 / qp_st2 / the warp intrinsics themselves
#include "qp_schur.cuh"  // qp_dmma, qp_ld2, qp_st2
#endif

namespace ub {

struct QpT {
    static constexpr int G = 29, LS = 30;
    using K = Compact;
    // per-warp shared memory (doubles)
    static constexpr int oLO = 0, oSM0 = G * LS, oSM1 = oSM0 + K::SMALL, oAB = oSM1 + K::SMALL, oY = oAB + K::APART, oMISC = oY + 32,
                         oBAR = oMISC + 4, PER_WARP = oBAR + 4;
    static constexpr int SLOT_BYTES = 2 * PER_WARP * 8 + 64;  // one trajectory: two warps' regions + its pointers (QtArgs) behind them
    // SLOTS trajectories share a CTA (one CTA per SM) so that their chains can run in LOCKSTEP: the unrolled column loops are ~0.5 MB of
    // straight-line code, the SM's instruction cache holds 32 KB, and with every warp at its own program counter the kernel was bound by
    // instruction fetch (ncu r02d: sm__icc hit rate 55 %, the GPC-level instruction cache at 74 % of its request peak, stall_no_instruction
    // the largest stall).  The warps of one direction meet at a named barrier once per group, stay within a few hundred instructions of
    // each other, and the line the first one fetches serves the other six.
    static constexpr int SLOTS = 7;
    static constexpr int SMEM_BYTES = SLOTS * SLOT_BYTES;
    static constexpr int THREADS = 64 * SLOTS;
    // factor image (shared memory during a group, global workspace afterwards): unscaled columns, 1 / sqrt(pivot), y
    static constexpr int fRI = 450, fY = 480, WS_GROUP = 510;
    // scratch vectors of the outward pass live behind the factor image in the LO region
    static constexpr int oVA = 512, oVC = 552, oNU2 = 592;
    static_assert(SLOT_BYTES % 16 == 0 && (PER_WARP * 8) % 16 == 0 && (oSM0 * 8) % 16 == 0 && (oSM1 * 8) % 16 == 0 && (oAB * 8) % 16 == 0 && (WS_GROUP * 8) % 16 == 0, "TMA alignment");
    // column c of the factor (rows c .. 28) sits at bc(c) + row: even bases, so that rows (2k, 2k+1) are one 16-byte load
    __host__ __device__ static constexpr int bc(int c) { return 28 * c - (c * (c - 1)) / 2 + c / 2; }
};
static_assert(QpT::bc(28) + 28 < QpT::fRI, "factor image");

#ifndef QT_HOST_TEST
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

// Per-warp context.  Deliberately tiny: everything else (buffer addresses, barrier addresses, the trajectory's pointers, which live in
// shared memory behind the two warps' regions) is recomputed at the point of use, because the chains keep a 29-entry row in registers
// and ptxas needs slack to batch the broadcast loads (with ~75 registers of loop invariants it serialised every load -> FMA pair).
struct QtArgs {
    const double* rec;   // this trajectory's compact record
    double* ws;          // this trajectory's workspace, (N + 1) groups
    double* step;
    double* mult;        // may be null
    int N;
    int slot;          // which trajectory of the CTA: named barrier 1 + slot pairs its two warps
    double delta;
    int lock_threads;  // 32 x (active trajectories of the CTA): arrival count of the two lockstep barriers
};
static_assert(sizeof(QtArgs) <= 64, "QtArgs lives in the 64 bytes behind the two warps' regions");
// named barriers: 1 + slot = the two warps of a trajectory; 8 / 9 = all top / all bottom warps of the CTA (lockstep, see QpT::SLOTS)
__device__ __forceinline__ void qt_named_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
struct QtCtx {
    double* sm;       // this warp's shared memory
    int lane;
    int cur;          // which of the two small buffers is "a"
    unsigned phases;  // bit b: phase of barrier b (0, 1: small buffers; 2: A part; 3: factor image)
    __device__ __forceinline__ double* small(int which) const { return sm + QpT::oSM0 + ((cur ^ which) & 1) * Compact::SMALL; }  // 0: a, 1: b
    __device__ __forceinline__ double* ab() const { return sm + QpT::oAB; }
    __device__ __forceinline__ uint64_t* bar(int b) const { return reinterpret_cast<uint64_t*>(sm + QpT::oBAR) + b; }
    __device__ __forceinline__ void issue(int b, double* dst, const double* src, unsigned bytes) const {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was modified in place through the generic proxy
            qt_bar_expect(bar(b), bytes);
            qt_bulk_load(dst, src, bytes, bar(b));
        }
    }
    __device__ __forceinline__ void wait(int b) {
        qt_bar_wait(bar(b), (phases >> b) & 1u);
        phases ^= 1u << b;
    }
    // all lanes must have finished reading the destination (the caller syncs the warp)
    __device__ __forceinline__ void load_small(int which, const double* chunk) const { issue((cur ^ which) & 1, small(which), chunk, Compact::SMALL * 8); }
    __device__ __forceinline__ void wait_small(int which) { wait((cur ^ which) & 1); }
    __device__ __forceinline__ void load_ab(const double* chunk) const { issue(2, ab(), chunk + Compact::oAQ, Compact::APART * 8); }
    __device__ __forceinline__ void load_factor(const double* wsg) const { issue(3, sm + QpT::oLO, wsg, QpT::WS_GROUP * 8); }
    __device__ __forceinline__ void swap() { cur ^= 1; }
};

#endif  // QT_HOST_TEST

// ---------------------------------------------------------------------------------------------------------------------------------
// Stage data.  `sm`: small part of a chunk (Cs, Cp, g, q -> t, Hd/Hb -> P^-1 in place); `ab`: AQ | AP of a chunk.
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

// y = P^-1 x for a 37-vector in shared memory (lanes 0..12 the diagonal part, lanes 13..20 one 3x3 block each).
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

// Which row of V_j = [A_j (13 rows); Cp_{j+1} (16 rows)] / U_j = [I 0; Cs_j] a lane owns: decoded once per kernel.
struct QtLane {
    int qr;    // AQ row (0..6) for the q+ / w+ rows, else -1
    int pc;    // p+ row c (0..2), else -1
    int vc;    // v+ row c (0..2), else -1
    int leg;   // contact lanes: leg 0..3 (else 0)
    int rr;    // contact lanes: row 0..3 within the leg (else -1)
    __device__ __forceinline__ explicit QtLane(int lane) {
        qr = lane >= 3 && lane < 7 ? lane - 3 : (lane >= 10 && lane < 13 ? lane - 6 : -1);
        pc = lane < 3 ? lane : -1;
        vc = lane >= 7 && lane < 10 ? lane - 7 : -1;
        const int r = lane - 13;
        leg = lane >= 13 && lane < 29 ? r >> 2 : 0;
        rr  = lane >= 13 && lane < 29 ? r & 3 : -1;
    }
};

// A row of V in pieces: the state part ...
struct QtV0 {
    double q[4], w[3], xp[3], xv[3];  // entries at the columns q, w, p_c, v_c
};
// ... and the input part of one leg.
struct QtVL {
    double f[3], r[3];
};

__device__ __forceinline__ void qt_load_v0(const double* __restrict__ ab, const double* __restrict__ smc, const QtLane& L, QtV0& v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v.q[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { v.w[i] = 0.0; v.xp[i] = 0.0; v.xv[i] = 0.0; }
    if (L.qr >= 0) {
        const double* row = ab + L.qr * 32;
        const double2 a = qp_ld2(row), b = qp_ld2(row + 2), c = qp_ld2(row + 4);
        v.q[0] = a.x; v.q[1] = a.y; v.q[2] = b.x; v.q[3] = b.y; v.w[0] = c.x; v.w[1] = c.y; v.w[2] = row[6];
    } else if (L.pc >= 0) {
        const double2 d = qp_ld2(ab + 224 + L.pc * 6);
#pragma unroll
        for (int i = 0; i < 3; ++i) { v.xp[i] = i == L.pc ? d.x : 0.0; v.xv[i] = i == L.pc ? d.y : 0.0; }
    } else if (L.vc >= 0) {
        const double d = ab[224 + (3 + L.vc) * 6];
#pragma unroll
        for (int i = 0; i < 3; ++i) v.xv[i] = i == L.vc ? d : 0.0;
    } else if (L.rr > 0) {
        const double* cp = smc + QK::oCp + (L.leg * 3 + L.rr - 1) * 8;
        const double2 a = qp_ld2(cp), b = qp_ld2(cp + 2);
        const double c4 = cp[4];
#pragma unroll
        for (int i = 0; i < 3; ++i) v.xp[i] = i == L.rr - 1 ? a.x : 0.0;
        v.q[0] = a.y; v.q[1] = b.x; v.q[2] = b.y; v.q[3] = c4;
    }
}

template <int LEG>
__device__ __forceinline__ void qt_load_vl(const double* __restrict__ ab, const double* __restrict__ smc, const QtLane& L, QtVL& v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { v.f[i] = 0.0; v.r[i] = 0.0; }
    if (L.qr >= 0) {
        const double* row = ab + L.qr * 32 + 8 + 6 * LEG;
        const double2 a = qp_ld2(row), b = qp_ld2(row + 2), c = qp_ld2(row + 4);
        v.f[0] = a.x; v.f[1] = a.y; v.f[2] = b.x; v.r[0] = b.y; v.r[1] = c.x; v.r[2] = c.y;
    } else if (L.pc >= 0) {
        const double fl = ab[224 + L.pc * 6 + 2 + LEG];
#pragma unroll
        for (int i = 0; i < 3; ++i) v.f[i] = i == L.pc ? fl : 0.0;
    } else if (L.vc >= 0) {
        const double fl = ab[224 + (3 + L.vc) * 6 + 1 + LEG];
#pragma unroll
        for (int i = 0; i < 3; ++i) v.f[i] = i == L.vc ? fl : 0.0;
    } else if (L.rr > 0 && L.leg == LEG) {
        const double* cp = smc + QK::oCp + (LEG * 3 + L.rr - 1) * 8;
        v.r[0] = cp[5]; v.r[1] = cp[6]; v.r[2] = cp[7];
    }
}

// piece . x  for a 37-vector x in shared memory
__device__ __forceinline__ double qt_v0_dot(const QtV0& v, const double* __restrict__ x) {
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { a0 += v.xp[i] * x[i]; a1 += v.xv[i] * x[7 + i]; a0 += v.w[i] * x[10 + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) a1 += v.q[i] * x[3 + i];
    return a0 + a1;
}
template <int LEG>
__device__ __forceinline__ double qt_vl_dot(const QtVL& v, const double* __restrict__ x) {
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { a0 += v.f[i] * x[13 + 6 * LEG + i]; a1 += v.r[i] * x[16 + 6 * LEG + i]; }
    return a0 + a1;
}

// piece <- piece P^-1
__device__ __forceinline__ void qt_v0_pinv(QtV0& v, const double* __restrict__ sm) {
    const double* pd = sm + QK::oHd;
#pragma unroll
    for (int i = 0; i < 3; ++i) { v.xp[i] *= pd[i]; v.xv[i] *= pd[7 + i]; v.w[i] *= pd[10 + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) v.q[i] *= pd[3 + i];
}
__device__ __forceinline__ void qt_block3(const double* __restrict__ h, double* x) {  // x <- x B, B symmetric 3x3 given by its upper triangle
    const double2 h01 = qp_ld2(h), h23 = qp_ld2(h + 2), h45 = qp_ld2(h + 4);
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    x[0] = x0 * h01.x + x1 * h01.y + x2 * h23.x;
    x[1] = x0 * h01.y + x1 * h23.y + x2 * h45.x;
    x[2] = x0 * h23.x + x1 * h45.x + x2 * h45.y;
}
template <int LEG>
__device__ __forceinline__ void qt_vl_pinv(QtVL& v, const double* __restrict__ sm) {
    qt_block3(sm + QK::oHb + 6 * (2 * LEG), v.f);
    qt_block3(sm + QK::oHb + 6 * (2 * LEG + 1), v.r);
}

// out[c] (+)= piece . (row c of V),  c = 0 .. 28   (state piece: sets or adds; leg pieces: add)
template <bool ADD>
__device__ __forceinline__ void qt_v0_dot_v_rows(const QtV0& w, const double* __restrict__ ab, const double* __restrict__ smc, double* out) {
#pragma unroll
    for (int c = 0; c < 29; ++c) {
        double acc;
        if ((c >= 3 && c < 7) || (c >= 10 && c < 13)) {
            const double* row = ab + (c < 7 ? c - 3 : c - 6) * 32;
            const double2 a = qp_ld2(row), b = qp_ld2(row + 2), d = qp_ld2(row + 4);
            acc = (w.q[0] * a.x + w.q[1] * a.y) + (w.q[2] * b.x + w.q[3] * b.y) + (w.w[0] * d.x + w.w[1] * d.y) + w.w[2] * row[6];
        } else if (c < 3) {
            const double2 d = qp_ld2(ab + 224 + c * 6);
            acc = w.xp[c] * d.x + w.xv[c] * d.y;
        } else if (c < 10) {
            acc = w.xv[c - 7] * ab[224 + (3 + c - 7) * 6];
        } else {
            const int rc = c - 13, legc = rc >> 2, rrc = rc & 3;
            if (rrc == 0) acc = 0.0;
            else {
                const double* cp = smc + QK::oCp + (legc * 3 + rrc - 1) * 8;
                const double2 a = qp_ld2(cp), b = qp_ld2(cp + 2);
                acc = w.xp[rrc - 1] * a.x + (w.q[0] * a.y + w.q[1] * b.x) + (w.q[2] * b.y + w.q[3] * cp[4]);
            }
        }
        if (ADD) out[c] += acc; else out[c] = acc;
    }
}
template <int LEG>
__device__ __forceinline__ void qt_vl_dot_v_rows(const QtVL& w, const double* __restrict__ ab, const double* __restrict__ smc, double* out) {
#pragma unroll
    for (int c = 0; c < 29; ++c) {
        if ((c >= 3 && c < 7) || (c >= 10 && c < 13)) {
            const double* row = ab + (c < 7 ? c - 3 : c - 6) * 32 + 8 + 6 * LEG;
            const double2 a = qp_ld2(row), b = qp_ld2(row + 2), d = qp_ld2(row + 4);
            out[c] += (w.f[0] * a.x + w.f[1] * a.y) + (w.f[2] * b.x + w.r[0] * b.y) + (w.r[1] * d.x + w.r[2] * d.y);
        } else if (c < 3) {
            out[c] += w.f[c] * ab[224 + c * 6 + 2 + LEG];
        } else if (c < 10) {
            out[c] += w.f[c - 7] * ab[224 + (3 + c - 7) * 6 + 1 + LEG];
        } else {
            const int rc = c - 13, legc = rc >> 2, rrc = rc & 3;
            if (legc == LEG && rrc > 0) {
                const double* cp = smc + QK::oCp + (legc * 3 + rrc - 1) * 8;
                out[c] += w.r[0] * cp[5] + w.r[1] * cp[6] + w.r[2] * cp[7];
            }
        }
    }
}

// e[c] = piece . (row c of U_j),  U_j = [I 0; Cs_j]: the state piece sets all 29 entries, a leg piece adds to its four contact rows
__device__ __forceinline__ void qt_v0_dot_u_rows(const QtV0& w, const double* __restrict__ sm, double* e) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { e[c] = w.xp[c]; e[7 + c] = w.xv[c]; e[10 + c] = w.w[c]; }
#pragma unroll
    for (int c = 0; c < 4; ++c) e[3 + c] = w.q[c];
#pragma unroll
    for (int c = 13; c < 29; ++c) {
        const int rc = c - 13, legc = rc >> 2, rrc = rc & 3, pcc = rrc == 0 ? 2 : rrc - 1;
        const double* cs = sm + QK::oCs + (legc * 4 + rrc) * 8;
        const double2 a = qp_ld2(cs), b = qp_ld2(cs + 2);
        e[c] = w.xp[pcc] * a.x + (w.q[0] * a.y + w.q[1] * b.x) + (w.q[2] * b.y + w.q[3] * cs[4]);
    }
}
template <int LEG>
__device__ __forceinline__ void qt_vl_dot_u_rows(const QtVL& w, const double* __restrict__ sm, double* e) {
#pragma unroll
    for (int rrc = 0; rrc < 4; ++rrc) {
        const double* cs = sm + QK::oCs + (LEG * 4 + rrc) * 8;
        e[13 + 4 * LEG + rrc] += w.r[0] * cs[5] + w.r[1] * cs[6] + w.r[2] * cs[7];
    }
}

// Row `lane` of U_j in sparse form (state lanes: unit vectors; contact lanes: the Cs row; lanes >= 29: null).  lm = indicator of the leg.
struct QtU {
    double xp[3], xv[3], q[4], w[3], r[3], lm[4];
    int leg;
};
__device__ __forceinline__ void qt_load_u(const double* __restrict__ sm, int lane, const QtLane& L, QtU& u) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { u.xp[i] = lane == i ? 1.0 : 0.0; u.xv[i] = lane == 7 + i ? 1.0 : 0.0; u.w[i] = lane == 10 + i ? 1.0 : 0.0; u.r[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { u.q[i] = lane == 3 + i ? 1.0 : 0.0; u.lm[i] = 0.0; }
    u.leg = L.leg;
    if (L.rr >= 0) {
        const double* cs = sm + QK::oCs + (L.leg * 4 + L.rr) * 8;
        const double2 a = qp_ld2(cs), b = qp_ld2(cs + 2), c = qp_ld2(cs + 4), d = qp_ld2(cs + 6);
        const int pc = L.rr == 0 ? 2 : L.rr - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) u.xp[i] = i == pc ? a.x : 0.0;
        u.q[0] = a.y; u.q[1] = b.x; u.q[2] = b.y; u.q[3] = c.x;
        u.r[0] = c.y; u.r[1] = d.x; u.r[2] = d.y;
#pragma unroll
        for (int i = 0; i < 4; ++i) u.lm[i] = i == L.leg ? 1.0 : 0.0;
    }
}
__device__ __forceinline__ double qt_u_dot(const QtU& u, const double* __restrict__ x) {  // u . x for a 37-vector x in shared memory
    double a0 = 0.0, a1 = 0.0;
    const double* xr = x + 13 + 6 * u.leg + 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) { a0 += u.xp[i] * x[i] + u.w[i] * x[10 + i]; a1 += u.xv[i] * x[7 + i] + u.r[i] * xr[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) a0 += u.q[i] * x[3 + i];
    return a0 + a1;
}
__device__ __forceinline__ void qt_u_pinv(QtU& u, const double* __restrict__ sm) {
    const double* pd = sm + QK::oHd;
#pragma unroll
    for (int i = 0; i < 3; ++i) { u.xp[i] *= pd[i]; u.xv[i] *= pd[7 + i]; u.w[i] *= pd[10 + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) u.q[i] *= pd[3 + i];
    qt_block3(sm + QK::oHb + 6 * (2 * u.leg + 1), u.r);
}
// out[c] += wu . (row c of U_j)
__device__ __forceinline__ void qt_u_dot_u_rows(const QtU& wu, const double* __restrict__ sm, double* out) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { out[c] += wu.xp[c]; out[7 + c] += wu.xv[c]; out[10 + c] += wu.w[c]; }
#pragma unroll
    for (int c = 0; c < 4; ++c) out[3 + c] += wu.q[c];
#pragma unroll
    for (int c = 13; c < 29; ++c) {
        const int rc = c - 13, legc = rc >> 2, rrc = rc & 3, pcc = rrc == 0 ? 2 : rrc - 1;
        const double* cs = sm + QK::oCs + (legc * 4 + rrc) * 8;
        const double2 a = qp_ld2(cs), b = qp_ld2(cs + 2), d = qp_ld2(cs + 4), f = qp_ld2(cs + 6);
        const double own = wu.r[0] * d.y + wu.r[1] * f.x + wu.r[2] * f.y;
        out[c] += wu.xp[pcc] * a.x + (wu.q[0] * a.y + wu.q[1] * b.x) + (wu.q[2] * b.y + wu.q[3] * d.x) + wu.lm[legc] * own;
    }
}
// e[c] = wu . (row c of V)
__device__ __forceinline__ void qt_u_dot_v_rows(const QtU& wu, const double* __restrict__ ab, const double* __restrict__ smc, double* e) {
    const int ro = 8 + 6 * wu.leg + 3;
#pragma unroll
    for (int c = 0; c < 29; ++c) {
        if ((c >= 3 && c < 7) || (c >= 10 && c < 13)) {
            const double* row = ab + (c < 7 ? c - 3 : c - 6) * 32;
            const double2 a = qp_ld2(row), b = qp_ld2(row + 2), d = qp_ld2(row + 4);
            e[c] = (wu.q[0] * a.x + wu.q[1] * a.y) + (wu.q[2] * b.x + wu.q[3] * b.y) + (wu.w[0] * d.x + wu.w[1] * d.y) + wu.w[2] * row[6] +
                   (wu.r[0] * row[ro] + wu.r[1] * row[ro + 1] + wu.r[2] * row[ro + 2]);
        } else if (c < 3) {
            const double2 d = qp_ld2(ab + 224 + c * 6);
            e[c] = wu.xp[c] * d.x + wu.xv[c] * d.y;
        } else if (c < 10) {
            e[c] = wu.xv[c - 7] * ab[224 + (3 + c - 7) * 6];
        } else {
            const int rc = c - 13, legc = rc >> 2, rrc = rc & 3;
            if (rrc == 0) e[c] = 0.0;
            else {
                const double* cp = smc + QK::oCp + (legc * 3 + rrc - 1) * 8;
                const double2 a = qp_ld2(cp), b = qp_ld2(cp + 2), d = qp_ld2(cp + 4), f = qp_ld2(cp + 6);
                const double own = wu.r[0] * d.y + wu.r[1] * f.x + wu.r[2] * f.y;
                e[c] = wu.xp[rrc - 1] * a.x + (wu.q[0] * a.y + wu.q[1] * b.x) + (wu.q[2] * b.y + wu.q[3] * d.x) + wu.lm[legc] * own;
            }
        }
    }
}

// S (+)= (V_j P_j^-1 V_j^T) row, formed piece by piece; returns (V_j t_j) entry of the lane.  SET: the sum starts from zero.
template <bool SET>
__device__ __forceinline__ double qt_vpv(const double* __restrict__ ab, const double* __restrict__ smc, const double* __restrict__ smp, const QtLane& L, double* s) {
    double vt;
    {
        QtV0 v;
        qt_load_v0(ab, smc, L, v);
        vt = qt_v0_dot(v, smp + QK::oQ);
        qt_v0_pinv(v, smp);
        qt_v0_dot_v_rows<!SET>(v, ab, smc, s);
    }
#define QT_LEG_PIECE(LEG)                                  \
    {                                                      \
        QtVL v;                                            \
        qt_load_vl<LEG>(ab, smc, L, v);                    \
        vt += qt_vl_dot<LEG>(v, smp + QK::oQ);             \
        qt_vl_pinv<LEG>(v, smp);                           \
        qt_vl_dot_v_rows<LEG>(v, ab, smc, s);              \
    }
    QT_LEG_PIECE(0) QT_LEG_PIECE(1) QT_LEG_PIECE(2) QT_LEG_PIECE(3)
#undef QT_LEG_PIECE
    return vt;
}

// e = (V_j P_j^-1 U_j^T) row; returns (V_j t_j) entry of the lane.
__device__ __forceinline__ double qt_vpu(const double* __restrict__ ab, const double* __restrict__ smc, const double* __restrict__ smp, const QtLane& L, double* e) {
    double vt;
    {
        QtV0 v;
        qt_load_v0(ab, smc, L, v);
        vt = qt_v0_dot(v, smp + QK::oQ);
        qt_v0_pinv(v, smp);
        qt_v0_dot_u_rows(v, smp, e);
    }
#define QT_LEG_PIECE(LEG)                                  \
    {                                                      \
        QtVL v;                                            \
        qt_load_vl<LEG>(ab, smc, L, v);                    \
        vt += qt_vl_dot<LEG>(v, smp + QK::oQ);             \
        qt_vl_pinv<LEG>(v, smp);                           \
        qt_vl_dot_u_rows<LEG>(v, smp, e);                  \
    }
    QT_LEG_PIECE(0) QT_LEG_PIECE(1) QT_LEG_PIECE(2) QT_LEG_PIECE(3)
#undef QT_LEG_PIECE
    return vt;
}

// (V_j x) entry of the lane for a 37-vector x in shared memory
__device__ __forceinline__ double qt_v_dot(const double* __restrict__ ab, const double* __restrict__ smc, const QtLane& L, const double* __restrict__ x) {
    double acc;
    {
        QtV0 v;
        qt_load_v0(ab, smc, L, v);
        acc = qt_v0_dot(v, x);
    }
#define QT_LEG_PIECE(LEG)                    \
    {                                        \
        QtVL v;                              \
        qt_load_vl<LEG>(ab, smc, L, v);      \
        acc += qt_vl_dot<LEG>(v, x);         \
    }
    QT_LEG_PIECE(0) QT_LEG_PIECE(1) QT_LEG_PIECE(2) QT_LEG_PIECE(3)
#undef QT_LEG_PIECE
    return acc;
}

#ifndef QT_HOST_TEST
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

#endif  // QT_HOST_TEST

// x[cc] += a * col[cc] for cc = FIRST .. 28, the column read from shared memory in batches of up to eight 16-byte loads that are
// all issued before the first dependent FMA (left to itself the compiler reuses one load register and serialises load -> FMA pairs).
// (`first` is a compile-time constant at every call site once the column loops are unrolled.)
__device__ __forceinline__ void qt_axpy_column(double* x, double a, const double* col, const int first) {
    constexpr int G = QpT::G;
    if (first & 1) x[first] += a * col[first];
    const int e0 = (first + 1) & ~1;  // first even index
#pragma unroll
    for (int base = e0; base < G; base += 16) {
        double2 t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (base + 2 * k < G) t[k] = qp_ld2(col + base + 2 * k);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (base + 2 * k < G) x[base + 2 * k] += a * t[k].x;
            if (base + 2 * k + 1 < G) x[base + 2 * k + 1] += a * t[k].y;
        }
    }
}

// Cholesky of the block held row-per-lane in s[] (lower triangle), fused with y = L^-1 rhs.  Column c is parked UNSCALED in the
// factor image (fimg[bc(c) + row] = S'[row][c], the pivot on top) and in the workspace; r_c = 1 / sqrt(pivot) goes to fimg[fRI + c].
// L[row][c] = S'[row][c] r_c; every update multiplies with r_c^2 instead.  Returns y (entry `lane`).
// (`stash` is re-read every other column after another lane wrote it: volatile, and no __restrict__ on the shared-memory operands.)
__device__ __forceinline__ double qt_cholesky(double* s, double rhs, double* fimg, double* __restrict__ wsg, volatile double* stash, int lane) {
    constexpr int G = QpT::G;
    double rk = rhs;
    double* const il = fimg + lane;
    double* const wl = wsg + lane;
#pragma unroll
    for (int c = 0; c < G; ++c) {
        if (unsigned(lane - c) < unsigned(G - c)) {  // c <= lane < 29
            il[QpT::bc(c)] = s[c];
            wl[QpT::bc(c)] = s[c];
        }
        if (lane == c) stash[c & 1] = rk;
        __syncwarp();
        const double rinv = rsqrt(fimg[QpT::bc(c) + c]);
        const double r2 = rinv * rinv;
        fimg[QpT::fRI + c] = rinv;  // every lane writes the same value
        const double tc = stash[c & 1] * r2;
        if (lane > c) rk -= s[c] * tc;
        if (c + 1 < G) {
            const double l2 = s[c] * r2;
            qt_axpy_column(s, -l2, fimg + QpT::bc(c), c + 1);
        }
    }
    __syncwarp();
    const double ri = fimg[QpT::fRI + min(lane, G - 1)];
    if (lane < G) wsg[QpT::fRI + lane] = ri;
    return rk * ri;
}

// e <- e L^-T (right-looking) with the unscaled factor image: the coupling row of the next group.
__device__ __forceinline__ void qt_trsm(double* e, const double* __restrict__ fimg) {
    constexpr int G = QpT::G;
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double rinv = fimg[QpT::fRI + c];
        const double e2 = e[c] * (rinv * rinv);
        e[c] *= rinv;
        if (c + 1 < G) qt_axpy_column(e, -e2, fimg + QpT::bc(c), c + 1);
    }
}

// (V_j^T nu)[k] for the local variable k = 0..36: nu = multipliers of group j+1 (29 entries in shared memory); `smc` holds Cp_{j+1}.
__device__ __forceinline__ double qt_vT_nu(const double* __restrict__ ab, const double* __restrict__ smc, const double* __restrict__ nu, int k) {
    double acc = 0.0;
    const int col = k < 3 ? -1 : k < 7 ? k - 3 : k < 10 ? -1 : k < 13 ? k - 6 : k - 5;
    if (col >= 0) {
#pragma unroll
        for (int qr = 0; qr < 7; ++qr) acc += ab[qr * 32 + col] * nu[qr < 4 ? 3 + qr : 6 + qr];
    }
    const double* ap = ab + 224;
    if (k < 3) {
        acc += ap[k * 6] * nu[k];
#pragma unroll
        for (int l = 0; l < 4; ++l) acc += smc[QK::oCp + (l * 3 + k) * 8] * nu[13 + 4 * l + k + 1];
    } else if (k < 7) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
#pragma unroll
            for (int r1 = 0; r1 < 3; ++r1) acc += smc[QK::oCp + (l * 3 + r1) * 8 + 1 + k - 3] * nu[13 + 4 * l + r1 + 1];
    } else if (k < 10) acc += ap[(k - 7) * 6 + 1] * nu[k - 7] + ap[(3 + k - 7) * 6] * nu[k];
    else if (k >= 13) {
        const int n = k - 13, l = n / 6, mm = n - 6 * l;
        if (mm < 3) acc += ap[mm * 6 + 2 + l] * nu[mm] + ap[(3 + mm) * 6 + 1 + l] * nu[7 + mm];
        else {
#pragma unroll
            for (int r1 = 0; r1 < 3; ++r1) acc += smc[QK::oCp + (l * 3 + r1) * 8 + 5 + mm - 3] * nu[13 + 4 * l + r1 + 1];
        }
    }
    return acc;
}

// (U_j^T nu)[k]: nu = multipliers of group j; U_j = [I 0; Cs_j].
__device__ __forceinline__ double qt_uT_nu(const double* __restrict__ sm, const double* __restrict__ nu, int k) {
    double acc = k < 13 ? nu[k] : 0.0;
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

// Solves L z = x, then L^T nu = y - z, with the factor image `f` (unscaled columns, r = 1 / sqrt(pivot) at fRI, y at fY): returns nu_lane.
// Both sweeps keep the running entries unnormalised (z_i = rk_i r_i, nu_i = u_i r_i^2), so no lane needs a per-step select.
__device__ __forceinline__ double qt_outward_solve(const double* __restrict__ f, double x, int lane) {
    constexpr int G = QpT::G;
    constexpr unsigned FULL = 0xffffffffu;
    const int ln = min(lane, G - 1);
    const double* own_row = f + ln;           // raw[ln][i] = f[bc(i) + ln]
    const double* own_col = f + QpT::bc(ln);  // raw[i][ln] = f[bc(ln) + i]
    const double r_own = f[QpT::fRI + ln];
    double rk = x;
#pragma unroll
    for (int i = 0; i < G - 1; ++i) {
        const double ri = f[QpT::fRI + i];
        const double ti = __shfl_sync(FULL, rk, i) * (ri * ri);
        if (lane > i) rk -= own_row[QpT::bc(i)] * ti;
    }
    double u = ((lane < G ? f[QpT::fY + ln] : 0.0) - rk * r_own) * (own_col[ln] * r_own);  // (y - z) d,  d = pivot r = sqrt(pivot)
#pragma unroll
    for (int i = G - 1; i > 0; --i) {
        const double ri = f[QpT::fRI + i];
        const double ni = __shfl_sync(FULL, u, i) * (ri * ri);
        if (lane < i) u -= own_col[i] * ni;
    }
    return lane < G ? u * (r_own * r_own) : 0.0;
}

#ifndef QT_HOST_TEST
// =================================================================================================================================
// One CTA of two warps per trajectory; the chains are separate (non-inlined) functions so that each gets its own register allocation.
// =================================================================================================================================

// Coupling row -> image, and the part of the next right-hand side it carries: returns -(row . y).
__device__ __forceinline__ double qt_store_coupling(const double* e, double* img, const double* sY, int lane) {
    constexpr int G = QpT::G, LS = QpT::LS;
    if (lane < G) {
#pragma unroll
        for (int c = 0; c + 1 < G; c += 2) qp_st2(img + lane * LS + c, e[c], e[c + 1]);
        qp_st2(img + lane * LS + G - 1, e[G - 1], 0.0);
    }
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int c = 0; c + 1 < G; c += 2) { const double2 t2 = qp_ld2(sY + c); d0 += e[c] * t2.x; d1 += e[c + 1] * t2.y; }
    d0 += e[G - 1] * sY[G - 1];
    return -(d0 + d1);
}

// S (+)= U P^-1 U^T + delta I of the chunk in `sma`;  returns the lane's part of the right-hand side: g - U t (contact lanes; state
// lanes: -t, their g comes from the neighbouring chunk).
template <bool SET>
__device__ __forceinline__ double qt_upu(const double* sma, int lane, double delta, double* s) {
    constexpr int G = QpT::G;
    const QtLane LN(lane);
    QtU u;
    qt_load_u(sma, lane, LN, u);
    const double r = (lane >= 13 && lane < G ? sma[QK::oG + lane] : 0.0) - qt_u_dot(u, sma + QK::oQ);
    qt_u_pinv(u, sma);
    if (SET) {
#pragma unroll
        for (int c = 0; c < G; ++c) s[c] = c == lane ? delta : 0.0;
    } else {
#pragma unroll
        for (int c = 0; c < G; ++c) s[c] += c == lane ? delta : 0.0;
    }
    qt_u_dot_u_rows(u, sma, s);
    return r;
}

#define QT_ARGS (*reinterpret_cast<const QtArgs*>(reinterpret_cast<const double*>(smem_raw) + 2 * QpT::PER_WARP))
#define QT_CHUNK(j) (QT_ARGS.rec + (long long)(j) * QK::NODE)
#define QT_WSG(j) (QT_ARGS.ws + (long long)(j) * QpT::WS_GROUP)
#define QT_PAIR_SYNC() qt_named_sync(1 + QT_ARGS.slot, 64)
#define QT_LOCKSTEP(dir) qt_named_sync(8 + (dir), QT_ARGS.lock_threads)

__device__ __forceinline__ void qt_store_step(const unsigned char* smem_raw, const double* sma, const double* vC, int j, int lane) {
    const int N = QT_ARGS.N;
    double* step = QT_ARGS.step;
    for (int k = lane; k < 37; k += 32) {
        const long long dst = k < 13 ? 13 * j + k : 13 * (N + 1) + 24 * j + (k - 13);
        step[dst] = -(sma[QK::oQ + k] + vC[k]);  // d_j = -(t_j + P^-1 (...))
    }
}
__device__ __forceinline__ void qt_store_mult(const unsigned char* smem_raw, double nu, int j, int lane) {
    double* mult = QT_ARGS.mult;
    const int N = QT_ARGS.N;
    if (mult && lane < QpT::G) {
        if (lane < 13) mult[13 * j + lane] = nu;
        else if (j < N) mult[13 * (N + 1) + 16 * j + (lane - 13)] = nu;
    }
}

__device__ __noinline__ void qt_top_chain(int lane, unsigned char* smem_raw) {
    using Q = QpT;
    constexpr int G = Q::G, LS = Q::LS;
    QtCtx cx{reinterpret_cast<double*>(smem_raw), lane, 0, 0u};
    double* const img = cx.sm + Q::oLO;
    double* const sY  = cx.sm + Q::oY;
    const int m = QT_ARGS.N / 2;
    const bool act = lane < G, st = lane < 13;
    double s[G];
    double rpart = 0.0;
    // ======================================================== top-down: groups 0 .. m-1
#pragma unroll
    for (int c = 0; c < G; ++c) s[c] = 0.0;
    cx.load_small(0, QT_CHUNK(0));
    cx.load_small(1, QT_CHUNK(1));
    cx.load_ab(QT_CHUNK(0));
    double gdef = st ? QT_ARGS.rec[QK::tail(QT_ARGS.N) + QK::tG0 + lane] : 0.0;  // x_0 - x_measured
    cx.wait_small(0);
    for (int j = 0; j < m; ++j) {  // a = small_j (landed), b = small_{j+1} and A_j (in flight)
        QT_LOCKSTEP(0);
        qt_pinv_t(cx.small(0), lane, true);
        rpart += gdef + qt_upu<false>(cx.small(0), lane, QT_ARGS.delta, s);  // S_jj += U P^-1 U^T + delta I ;  rhs = g - U t + rpart
        if (j > 0) qt_syrk(img, s, lane);
        QT_LOCKSTEP(0);
        const double yj = qt_cholesky(s, rpart, img, QT_WSG(j), cx.sm + Q::oMISC, lane);
        if (act) {
            sY[lane] = yj;
            QT_WSG(j)[Q::fY + lane] = yj;
        }
        cx.wait_small(1);
        cx.wait(2);
        QT_LOCKSTEP(0);
        {   // coupling rows of group j+1: (V P^-1 U^T) row, solved against L_j, parked in the image
            double e[G];
            const QtLane LN(lane);
            const double carry = qt_vpu(cx.ab(), cx.small(1), cx.small(0), LN, e);
            qt_trsm(e, img);
            __syncwarp();
            rpart = qt_store_coupling(e, img, sY, lane) - carry;  // -Lo y_j - V t_j
        }
        QT_LOCKSTEP(0);
        {
            const QtLane LN(lane);
            qt_vpv<true>(cx.ab(), cx.small(1), cx.small(0), LN, s);  // S_{j+1,j+1} part: V P^-1 V^T row
        }
        gdef = st ? cx.small(0)[QK::oG + lane] : 0.0;  // defect of stage j: the state rows of group j+1
        __syncwarp();
        cx.swap();  // a = small_{j+1}
        if (j + 1 < m) {
            cx.load_small(1, QT_CHUNK(j + 2));
            cx.load_ab(QT_CHUNK(j + 1));
        }
    }
    // ---- middle group m: top part S = V P^-1 V^T - Lo Lo^T, rhs = defect - V t - Lo y
    qt_syrk(img, s, lane);
    rpart += gdef;
    QT_PAIR_SYNC();  // the bottom warp has parked its part of S_mm and of the right-hand side in its image
    if (act) {
        const double* mine = cx.sm + Q::PER_WARP + Q::oLO + lane * LS;
#pragma unroll
        for (int c = 0; c + 1 < G; c += 2) { const double2 t2 = qp_ld2(mine + c); s[c] += t2.x; s[c + 1] += t2.y; }
        const double2 t2 = qp_ld2(mine + G - 1);
        s[G - 1] += t2.x;
        rpart += t2.y;
    }
    {
        const double ym = qt_cholesky(s, rpart, img, QT_WSG(m), cx.sm + Q::oMISC, lane);
        if (act) img[Q::fY + lane] = ym;
        __syncwarp();
        const double num = qt_outward_solve(img, 0.0, lane);  // nu_m = L^-T y_m
        __syncwarp();
        sY[lane] = num;
        cx.sm[Q::PER_WARP + Q::oY + lane] = num;
        qt_store_mult(smem_raw, num, m, lane);
    }
    QT_PAIR_SYNC();  // nu_m is visible to the bottom warp

    // ======================================================== outward: groups m-1 .. 0
    // the factor images in the workspace were written through the generic proxy and come back through the async proxy (TMA)
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncwarp();
    double* const vA = img + Q::oVA;
    double* const vC = img + Q::oVC;
    // every load is issued one group ahead, as soon as the last reader of its buffer is done (the first version issued and awaited them
    // at the top of each group: a full HBM latency per group on a 50-group dependent chain)
    cx.load_small(1, QT_CHUNK(m));  // Cp_m
    if (m > 0) {
        cx.load_small(0, QT_CHUNK(m - 1));
        cx.load_ab(QT_CHUNK(m - 1));
        cx.load_factor(QT_WSG(m - 1));
    }
    cx.wait_small(1);
    for (int j = m - 1; j >= 0; --j) {  // a = small_j, A_j, factor_j (in flight or landed), b = small_{j+1}
        QT_LOCKSTEP(0);
        cx.wait_small(0);
        cx.wait(2);
        qt_pinv_t(cx.small(0), lane, true);
        for (int k = lane; k < 37; k += 32) vA[k] = qt_vT_nu(cx.ab(), cx.small(1), sY, k);  // a = V_j^T nu_{j+1}
        __syncwarp();
        if (j > 0) {  // A_j and Cp_{j+1} have been consumed: their buffers take the next group's data
            cx.load_small(1, QT_CHUNK(j - 1));
            cx.load_ab(QT_CHUNK(j - 1));
        }
        qt_apply_pinv(cx.small(0), vA, vC, lane);
        double z;
        {
            const QtLane LN(lane);
            QtU u;
            qt_load_u(cx.small(0), lane, LN, u);
            z = qt_u_dot(u, vC);  // (U P^-1 a) row
        }
        cx.wait(3);
        const double nu = qt_outward_solve(img, z, lane);
        __syncwarp();
        if (j > 0) cx.load_factor(QT_WSG(j - 1));  // vA / vC live behind the factor image
        sY[lane] = nu;
        __syncwarp();
        for (int k = lane; k < 37; k += 32) vA[k] += qt_uT_nu(cx.small(0), sY, k);
        __syncwarp();
        qt_apply_pinv(cx.small(0), vA, vC, lane);
        qt_store_step(smem_raw, cx.small(0), vC, j, lane);
        qt_store_mult(smem_raw, nu, j, lane);
        __syncwarp();
        cx.swap();  // a = small_{j-1} (in flight), b = small_j: group j-1 needs its Cp rows
    }
}

__device__ __noinline__ void qt_bottom_chain(int lane, unsigned char* smem_raw) {
    using Q = QpT;
    constexpr int G = Q::G, LS = Q::LS;
    QtCtx cx{reinterpret_cast<double*>(smem_raw) + Q::PER_WARP, lane, 0, 0u};
    double* const img = cx.sm + Q::oLO;
    double* const sY  = cx.sm + Q::oY;
    const int N = QT_ARGS.N, m = N / 2;
    const bool act = lane < G, st = lane < 13;
    double s[G];
    double rpart = 0.0;
    // ======================================================== bottom-up: groups N .. m+1
    // a = small_j (synthesised for j = N: no inputs, no contact rows), b = small_{j-1}, A_{j-1}
    for (int k = lane; k < QK::SMALL; k += 32) cx.small(0)[k] = 0.0;
    __syncwarp();
    if (st) {
        const double* tail = QT_ARGS.rec + QK::tail(N);
        cx.small(0)[QK::oQ + lane]  = tail[QK::tQN + lane];
        cx.small(0)[QK::oHd + lane] = tail[QK::tHN + lane];
    }
    __syncwarp();
    cx.load_small(1, QT_CHUNK(N - 1));
    cx.load_ab(QT_CHUNK(N - 1));
    qt_pinv_t(cx.small(0), lane, false);
    for (int j = N; j > m; --j) {
        QT_LOCKSTEP(1);
        rpart += qt_upu<true>(cx.small(0), lane, QT_ARGS.delta, s);  // S_jj = U P^-1 U^T + delta I ;  rhs = contact values - U t + rpart
        if (j < N) qt_syrk(img, s, lane);
        cx.wait_small(1);
        cx.wait(2);
        QT_LOCKSTEP(1);
        qt_pinv_t(cx.small(1), lane, true);
        rpart += st ? cx.small(1)[QK::oG + lane] : 0.0;  // defect of stage j-1
        {
            const QtLane LN(lane);
            rpart -= qt_vpv<false>(cx.ab(), cx.small(0), cx.small(1), LN, s);  // S_jj += V P^-1 V^T ;  rhs -= V t_{j-1}
        }
        QT_LOCKSTEP(1);
        const double yj = qt_cholesky(s, rpart, img, QT_WSG(j), cx.sm + Q::oMISC, lane);
        if (act) {
            sY[lane] = yj;
            QT_WSG(j)[Q::fY + lane] = yj;
        }
        QT_LOCKSTEP(1);
        {   // coupling rows of group j-1: (U_{j-1} P^-1 V^T) row, solved against M_j, parked in the image
            double e[G];
            const QtLane LN(lane);
            QtU u;
            qt_load_u(cx.small(1), lane, LN, u);
            qt_u_pinv(u, cx.small(1));
            qt_u_dot_v_rows(u, cx.ab(), cx.small(0), e);
            __syncwarp();
            if (j - 2 >= m) cx.load_ab(QT_CHUNK(j - 2));
            qt_trsm(e, img);
            __syncwarp();
            rpart = qt_store_coupling(e, img, sY, lane);  // -Uo y_j
        }
        __syncwarp();
        cx.swap();  // a = small_{j-1} (P^-1 and t already in place)
        if (j - 2 >= m) cx.load_small(1, QT_CHUNK(j - 2));
    }
    // ---- middle group m, bottom part: U P^-1 U^T + delta I - Uo Uo^T ; rhs = contact values - U t - Uo y   (a = small_m)
    rpart += qt_upu<true>(cx.small(0), lane, QT_ARGS.delta, s);
    qt_syrk(img, s, lane);
    if (act) {
#pragma unroll
        for (int c = 0; c + 1 < G; c += 2) qp_st2(img + lane * LS + c, s[c], s[c + 1]);
        qp_st2(img + lane * LS + G - 1, s[G - 1], rpart);
    }
    QT_PAIR_SYNC();  // parked
    QT_PAIR_SYNC();  // nu_m has arrived in sY

    // ======================================================== outward: groups m+1 .. N (and the steps d_m .. d_N)
    // invariant at group j: a = small_{j-1} (P^-1, t in place), A_{j-1} landed, sY = nu_{j-1}
    asm volatile("fence.proxy.async.global;" ::: "memory");  // factor images: generic-proxy stores, async-proxy (TMA) loads
    __syncwarp();
    double* const vA = img + Q::oVA;
    double* const vC = img + Q::oVC;
    double* const nu2 = img + Q::oNU2;
    cx.load_ab(QT_CHUNK(m));
    if (m + 1 <= N) cx.load_factor(QT_WSG(m + 1));
    if (m + 1 < N) cx.load_small(1, QT_CHUNK(m + 1));
    for (int j = m + 1; j <= N; ++j) {  // loads are issued one group ahead (see the top chain)
        QT_LOCKSTEP(1);
        const bool last = j == N;
        if (!last) cx.wait_small(1);
        else {
            for (int k = lane; k < 96; k += 32) cx.small(1)[QK::oCp + k] = 0.0;  // group N has no contact rows
            __syncwarp();
        }
        for (int k = lane; k < 37; k += 32) vA[k] = qt_uT_nu(cx.small(0), sY, k);  // b = U_{j-1}^T nu_{j-1}
        __syncwarp();
        qt_apply_pinv(cx.small(0), vA, vC, lane);
        cx.wait(2);  // A_{j-1}
        double z;
        {
            const QtLane LN(lane);
            z = qt_v_dot(cx.ab(), cx.small(1), LN, vC);  // (V_{j-1} P^-1 b) row
        }
        cx.wait(3);
        const double nu = qt_outward_solve(img, z, lane);
        __syncwarp();
        if (!last) cx.load_factor(QT_WSG(j + 1));  // vA / vC / nu2 live behind the factor image
        nu2[lane] = nu;
        __syncwarp();
        for (int k = lane; k < 37; k += 32) vA[k] += qt_vT_nu(cx.ab(), cx.small(1), nu2, k);
        __syncwarp();
        if (!last) cx.load_ab(QT_CHUNK(j));  // A_{j-1} has been consumed
        qt_apply_pinv(cx.small(0), vA, vC, lane);
        qt_store_step(smem_raw, cx.small(0), vC, j - 1, lane);  // d_{j-1}
        qt_store_mult(smem_raw, nu, j, lane);
        sY[lane] = nu;
        __syncwarp();
        if (!last) {
            cx.swap();  // a = small_j, b = the buffer small_{j-1} leaves
            if (j + 1 < N) cx.load_small(1, QT_CHUNK(j + 1));
            qt_pinv_t(cx.small(0), lane, true);
        }
    }
    // d_N = -P_N^-1 (q_N + nu_N)
    if (st) {
        const double* tail = QT_ARGS.rec + QK::tail(N);
        QT_ARGS.step[13 * N + lane] = -(tail[QK::tQN + lane] + sY[lane]) / tail[QK::tHN + lane];
    }
}
#undef QT_PAIR_SYNC
#undef QT_LOCKSTEP
#undef QT_ARGS
#undef QT_CHUNK
#undef QT_WSG

__global__ void __launch_bounds__(QpT::THREADS, 1)
qp_twisted_kernel(const double* __restrict__ rec_all, long long ld_rec, double* __restrict__ ws_all, double* __restrict__ step_all, long long ld_step,
                  double* __restrict__ mult_all, long long ld_mult, int N, long long batch, double delta, const int* __restrict__ skip_status) {
    extern __shared__ __align__(16) unsigned char smem_cta[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, slot = warp >> 1, wib = warp & 1;
    unsigned char* const smem_raw = smem_cta + slot * QpT::SLOT_BYTES;  // this trajectory's two regions + QtArgs
    const long long b = (long long)blockIdx.x * QpT::SLOTS + slot;
    // SQP loop: a trajectory that has stopped is skipped (uniform over its two warps)
    const bool active = b < batch && !(skip_status && skip_status[2 * b] != 0);
    if (active) {
        uint64_t* const bars = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem_raw) + wib * QpT::PER_WARP + QpT::oBAR);
        if (lane < 4) qt_bar_init(bars + lane);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int n_active = __syncthreads_count(active && lane == 0 && wib == 0);
    if (!active) return;  // from here on only named barriers with explicit arrival counts
    if (wib == 0 && lane == 0)
        *reinterpret_cast<QtArgs*>(reinterpret_cast<double*>(smem_raw) + 2 * QpT::PER_WARP) =
            QtArgs{rec_all + b * ld_rec, ws_all + b * (long long)(N + 1) * QpT::WS_GROUP, step_all + b * ld_step,
                   mult_all ? mult_all + b * ld_mult : nullptr, N, slot, delta, 32 * n_active};
    qt_named_sync(1 + slot, 64);
    if (wib == 0) qt_top_chain(lane, smem_raw);
    else qt_bottom_chain(lane, smem_raw);
}

#endif  // QT_HOST_TEST

}  // namespace ub
