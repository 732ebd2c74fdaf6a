// Batched equality-constrained QP solve for the quadruped NMPC, consuming the KKT block records in place.
//
// Replaces what SoftSQPOptimizer::SolveLocalQPProblem hands to OSQP (include/ungar/optimization/soft_sqp.hpp:193-233):
//     min_d 1/2 d^T P d + q^T d   s.t.  A d = -g            (l = u = -g: the soft SQP only poses equality-constrained QPs)
// OSQP v0.6.3 (ADMM) is absent from the reference tree; this is an exact stage-wise factorisation instead.
//
// Algorithm (one warp per trajectory; numpy statement and derivation: oracle/qp_reference.py::schur_stagewise).
// P is block diagonal over the stage variables w_j = [x_j; u_j] (quadruped: no input-rate term), and each H_j is a diagonal
// (13 state entries) plus eight 3x3 blocks (f_i, r_i of each leg), so P^-1 is closed form.  Group the constraint rows as
//     nu_j = [ dynamics defect of stage j-1 (or x_0 - x_measured for j = 0) ; contact rows of stage j ]      (29 rows)
// Every group touches only w_{j-1} and w_j:   nu_j rows = V_{j-1} w_{j-1} + U_j w_j,   U_j = [I_x 0; Cs_j],  V_j = [A_j; Cp_{j+1}]
// so the Schur complement S = A P^-1 A^T + delta I is block TRIDIAGONAL with 29x29 blocks
//     S_jj = U_j P_j^-1 U_j^T + V_{j-1} P_{j-1}^-1 V_{j-1}^T + delta I,      S_{j+1,j} = V_j P_j^-1 U_j^T
// and is factorised by a block Cholesky sweep; then nu from two block substitutions and d_j = -P_j^-1 (q_j + U_j^T nu_j +
// V_j^T nu_{j+1}).  delta = 1e-9 keeps rank-deficient rows (swing legs: all-zero contact rows) harmless, like OSQP's rho/sigma.
//
// Mapping.  Lane i < 29 owns ROW i of every 29-row block and keeps it in REGISTERS (S_jj row, S_{j,j-1} row, V P^-1 row); only
// the operands that every lane needs (rows of L_{j-1,j-1}, of Lo, of A_j, Cs_j, Cp_{j+1}) sit in shared memory and are read as
// 16-byte broadcasts.  The sparsity of U and of the contact rows is unrolled at compile time: a contact row has ten non-zeros
// (seven base-state columns + the foot position of its own leg), U's state rows are the identity.
// Forward sweep per stage:  S_jj row update -> TRSM of the S_{j,j-1} row against L_{j-1,j-1} -> SYRK -> Cholesky fused with the
// forward substitution (one column per step, pivot by shuffle, column through a 2-slot shared buffer) -> L_jj leaves through
// the TMA engine (one cp.async.bulk per stage) -> V P^-1 row and its three products for stage j+1.
// Backward sweep per stage: only L_jj and y_j are read back; Lo_{j+1}^T nu_{j+1} = L_jj^-1 U_j P_j^-1 V_j^T nu_{j+1} is rebuilt
// from the stage data with sparse mat-vecs and one extra substitution, so Lo never goes to global memory.
// Global latency: the stage data of stage j+1 (A, Cs, Cp, q, g, the 61 entries of H that P^-1 needs) is fetched with cp.async
// while stage j computes; the landing zone is whichever of two 870-double regions does not hold stage j's data, and the same
// region serves as the Lo image in between (ping-pong, no extra shared memory).  The backward sweep re-carves the three big
// regions into two (packed L, stage data) pairs and prefetches stage j-1 the same way.
// Shared memory 24.1 KB per warp -> 8 warps per SM, 1184 trajectories per wave (the 1024-trajectory headline is one wave).
#pragma once

#include "sweep.cuh"
#include "sweep_structured.cuh"  // bulk_store

namespace ub {

struct QpShape {
    static constexpr int NX = 13, NU = 24, NZ = 37, TRI = 703, G = 29;
    static constexpr int LS = 30;  // row stride of the L / Lo images: even (16-byte rows); column 29 of L holds 1 / L_ii
    static constexpr int LA = 38;  // row stride of the A image (column 37 is a zero pad)
    // per-warp shared memory (doubles); every offset is even so that double2 accesses are aligned
    // forward:  L image (870) | X (870) | Y (870), X and Y alternating between "stage data" and "Lo image"
    // backward: packed L (464) x 2 | stage data (814) x 2 in the same 2610 doubles
    static constexpr int oL = 0, oX = G * LS, oY = 2 * G * LS, big = 3 * G * LS;
    static constexpr int dA = 0, dCs = NX * LA, dCp = dCs + 160, dTotal = dCp + 160;  // inside a stage-data region
    static constexpr int LP = 464;                                                   // packed lower triangle (435) + 1 / L_ii (29)
    static constexpr int bD0 = 2 * LP;
    static constexpr int oP = big, oT = oP + 14 + 8 * 10, oQ = oT + 38, oNu = oQ + 38, oC = oNu + 32, oSt = oC + 64, total = oSt + 136;
    static constexpr int stQ = 0, stG = 38, stH = 70;  // staging of the small per-stage vectors: q (38) | g or y (32) | H entries (64)
    static constexpr int WARPS = 4;
    static constexpr int SMEM_BYTES = WARPS * total * 8;
    // per-trajectory global workspace (doubles): per group  L_jj image (G * LS) | y_j (32)
    static constexpr int WS_GROUP = G * LS + 32;
    static_assert(dTotal <= G * LS && bD0 + 2 * dTotal <= big, "region carving");
    static_assert(total % 2 == 0 && dCs % 2 == 0 && oP % 2 == 0 && oT % 2 == 0 && oNu % 2 == 0 && oC % 2 == 0 && oSt % 2 == 0, "alignment");
    static_assert((WS_GROUP * 8) % 16 == 0, "bulk copies need 16-byte multiples");
};

__device__ __forceinline__ double2 qp_ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void qp_st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
__device__ __forceinline__ void qp_cp8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// D (8x8) += A (8x4, row) * B (4x8, col) in fp64: lane holds A[lane / 4][lane % 4], B[lane % 4][lane / 4], D[lane / 4][2 (lane % 4) + {0, 1}].
__device__ __forceinline__ void qp_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Closed-form inverse of a symmetric 3x3 block [a b c; b d e; c e f] (row-major 9 entries out).
__device__ __forceinline__ void inv_sym3(double a, double b, double c, double d, double e, double f, double* out) {
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02, id = 1.0 / det;
    out[0] = c00 * id; out[1] = c01 * id; out[2] = c02 * id;
    out[3] = out[1];   out[4] = (a * f - c * c) * id; out[5] = (b * c - a * e) * id;
    out[6] = out[2];   out[7] = out[5]; out[8] = (a * d - b * b) * id;
}

// P^-1 image: 13 reciprocals at [0, 13), block `blk` (row-major 3x3) at 14 + 10 * blk.
// y = P^-1 x cooperatively: lanes 0..12 the diagonal state part, lanes 13..20 one 3x3 input block each.
__device__ __forceinline__ void apply_pinv_warp(const double* __restrict__ pinv, const double* __restrict__ x, double* __restrict__ y, int lane) {
    if (lane < 13) y[lane] = pinv[lane] * x[lane];
    else if (lane < 21) {
        const int blk = lane - 13;
        const double* m = pinv + 14 + 10 * blk;
        const double* xv = x + 13 + 3 * blk;
        y[13 + 3 * blk + 0] = m[0] * xv[0] + m[1] * xv[1] + m[2] * xv[2];
        y[13 + 3 * blk + 1] = m[3] * xv[0] + m[4] * xv[1] + m[5] * xv[2];
        y[13 + 3 * blk + 2] = m[6] * xv[0] + m[7] * xv[1] + m[8] * xv[2];
    }
    __syncwarp();
}

// Starts the asynchronous fetch of stage j into a stage-data region sD (A_j 13 x 37 at stride 38, Cs_j / Cp_{j+1} 16 x 10 each: the
// non-zero columns of the contact rows of stage j and of stage j+1 on w_j) and into the staging area sSt (q_j, the entries of H_j
// that P_j^-1 needs, and g_j when `with_g`).  Zero-filled where the horizon ends (j = N: only the state part).  No commit.
__device__ __forceinline__ void qp_fetch_stage(const double* __restrict__ rec, const RecLayout& L, int N, int j, double* __restrict__ sD,
                                               double* __restrict__ sSt, int lane, bool with_g) {
    using Q = QpShape;
    const bool live = j < N, next = j + 1 < N;
    const int nX = Q::NX * (N + 1);
    const double* cj = rec + L.C + (long long)j * 320;
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        const int e = lane + 32 * it, row = e / 10, col = e - 10 * row;
        if (live) qp_cp8(sD + Q::dCs + e, cj + row * 20 + col); else sD[Q::dCs + e] = 0.0;
        if (next) qp_cp8(sD + Q::dCp + e, cj + 320 + row * 20 + 10 + col); else sD[Q::dCp + e] = 0.0;
    }
    const double* aj = rec + L.A + (long long)j * 481;
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int e = lane + 32 * it;
        if (e < 481) {
            const int row = e / 37, col = e - 37 * row;
            if (live) qp_cp8(sD + Q::dA + row * Q::LA + col, aj + e); else sD[Q::dA + row * Q::LA + col] = 0.0;
        }
    }
    if (lane < 13) sD[Q::dA + lane * Q::LA + 37] = 0.0;  // pad column, read by the 16-byte row loads
    for (int k = lane; k < 37; k += 32) {
        if (k < 13) qp_cp8(sSt + Q::stQ + k, rec + L.grad + Q::NX * j + k);
        else if (live) qp_cp8(sSt + Q::stQ + k, rec + L.grad + nX + Q::NU * j + (k - 13));
        else sSt[Q::stQ + k] = 0.0;
    }
    if (with_g) {
        if (lane < 13) qp_cp8(sSt + Q::stG + lane, rec + L.g + Q::NX * j + lane);
        else if (live && lane < Q::G) qp_cp8(sSt + Q::stG + lane, rec + L.g + nX + 16 * j + (lane - 13));
        else sSt[Q::stG + lane] = 0.0;
    }
    if (live) {  // diagonal of the state block and the six distinct entries of each 3x3 input block, from the packed upper triangle of H_j
        const double* H = rec + L.H + (long long)j * Q::TRI;
        if (lane < 13) qp_cp8(sSt + Q::stH + lane, H + (lane * 37 - (lane * (lane - 1)) / 2));
        else if (lane < 21) {
            const int a = 13 + 3 * (lane - 13);
            const int d0 = a * 37 - (a * (a - 1)) / 2, d1 = (a + 1) * 37 - ((a + 1) * a) / 2, d2 = (a + 2) * 37 - ((a + 2) * (a + 1)) / 2;
            double* dst = sSt + Q::stH + 16 + 6 * (lane - 13);
            qp_cp8(dst + 0, H + d0); qp_cp8(dst + 1, H + d0 + 1); qp_cp8(dst + 2, H + d0 + 2);
            qp_cp8(dst + 3, H + d1); qp_cp8(dst + 4, H + d1 + 1); qp_cp8(dst + 5, H + d2);
        }
    } else if (lane < 13) qp_cp8(sSt + Q::stH + lane, rec + L.HN + (lane * 13 - (lane * (lane - 1)) / 2));
}

// P_j^-1 from the staged entries of H_j.
__device__ __forceinline__ void qp_build_pinv(const double* __restrict__ sSt, double* __restrict__ sP, bool live, int lane) {
    using Q = QpShape;
    if (lane < 13) sP[lane] = 1.0 / sSt[Q::stH + lane];
    else if (lane < 21) {
        double* out = sP + 14 + 10 * (lane - 13);
        if (live) {
            const double* h = sSt + Q::stH + 16 + 6 * (lane - 13);
            inv_sym3(h[0], h[1], h[2], h[3], h[4], h[5], out);
        } else {
#pragma unroll
            for (int e = 0; e < 9; ++e) out[e] = 0.0;
        }
    }
    __syncwarp();
}

// sum_r Cx[r][column k of w] * nu[13 + r] over the sixteen contact rows: entry k of Cx^T nu (Cx = Cs or Cp image).
__device__ __forceinline__ double qp_contact_T(const double* __restrict__ sCx, const double* __restrict__ nu, int k) {
    double acc = 0.0;
    if (k < 7) {
#pragma unroll
        for (int r = 0; r < 16; ++r) acc += sCx[r * 10 + k] * nu[13 + r];
    } else if (k >= 13) {
        const int q = k - 13, leg = q / 6, m = q - 6 * leg - 3;
        if (m >= 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc += sCx[(4 * leg + i) * 10 + 7 + m] * nu[13 + 4 * leg + i];
        }
    }
    return acc;
}

// Forward sweep (factorisation + forward substitution), then backward sweep (multipliers and step).
__global__ void __launch_bounds__(QpShape::WARPS * 32, 2)
qp_schur_kernel(const double* __restrict__ rec_all, long long ld_rec, double* __restrict__ ws_all, double* __restrict__ step_all,
                long long ld_step, double* __restrict__ mult_all, long long ld_mult, int N, long long batch, RecLayout L, double delta,
                const int* __restrict__ skip_status) {
    using Q = QpShape;
    constexpr int G = Q::G, LS = Q::LS, LA = Q::LA;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* const sm = reinterpret_cast<double*>(smem_raw) + wib * Q::total;
    double *sL = sm + Q::oL, *sP = sm + Q::oP, *sT = sm + Q::oT, *sQ = sm + Q::oQ, *sY = sm + Q::oNu, *sC = sm + Q::oC, *sSt = sm + Q::oSt;
    const long long b = (long long)blockIdx.x * Q::WARPS + wib;
    if (b >= batch) return;
    if (skip_status && skip_status[2 * b] != 0) return;  // SQP loop: this trajectory has stopped (warp-uniform)
    const double* __restrict__ rec = rec_all + b * ld_rec;
    double* __restrict__ ws = ws_all + b * (long long)(N + 1) * Q::WS_GROUP;
    const int nX = Q::NX * (N + 1);

    // lane roles: state rows (identity in U, A rows in V) or contact rows (Cs rows in U, Cp rows in V); lanes 29..31 shadow row 28
    const bool act = lane < G, st = lane < 13;
    const int r = st ? 0 : min(lane - 13, 15), leg = r >> 2;

    if (lane < 8) sP[14 + 10 * lane + 9] = 0.0;
    if (lane == 0) { sP[13] = 0.0; sT[37] = 0.0; }
    sY[lane] = 0.0;
    double* sD = sm + Q::oX;  // stage data of the current stage
    double* sO = sm + Q::oY;  // the other region: Lo image, then landing zone of the next stage's data
    qp_fetch_stage(rec, L, N, 0, sD, sSt, lane, true);
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ================================================================ forward sweep over the groups nu_0 .. nu_N
    double s[G], e[G];  // row `lane` of S_jj (lower triangle meaningful) and of S_{j,j-1} / Lo_j
#pragma unroll
    for (int c = 0; c < G; ++c) { s[c] = 0.0; e[c] = 0.0; }
    double carry = 0.0;  // (V_{j-1} t_{j-1})[lane]
    for (int j = 0; j <= N; ++j) {
        const bool live = j < N;
        const double *sA = sD + Q::dA, *sCs = sD + Q::dCs, *sCp = sD + Q::dCp;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const double gval = sSt[Q::stG + lane];
        qp_build_pinv(sSt, sP, live, lane);
        apply_pinv_warp(sP, sSt + Q::stQ, sT, lane);  // t_j = P_j^-1 q_j

        // ---- own U row (contact lanes: the ten non-zeros) and its product with P^-1 -------------------------------------------
        double cs[10], wu[10];
#pragma unroll
        for (int k = 0; k < 10; k += 2) { const double2 t2 = qp_ld2(sCs + r * 10 + k); cs[k] = t2.x; cs[k + 1] = t2.y; }
#pragma unroll
        for (int k = 0; k < 7; ++k) wu[k] = cs[k] * sP[k];
        {
            const double* Bm = sP + 14 + 10 * (2 * leg + 1);
#pragma unroll
            for (int m = 0; m < 3; ++m) wu[7 + m] = cs[7] * Bm[m] + cs[8] * Bm[3 + m] + cs[9] * Bm[6 + m];
        }
        // ---- S_jj += U P^-1 U^T + delta I (lower triangle) ;  rhs_j = g_j - U t_j - carry --------------------------------------
        const double dself = sP[min(lane, 12)];
#pragma unroll
        for (int c = 0; c < 13; ++c) s[c] += st ? (c == lane ? dself + delta : 0.0) : (c < 7 ? wu[c] : 0.0);
#pragma unroll
        for (int c = 13; c < G; ++c) {
            const int rc = c - 13, legc = rc >> 2;
            const double* row = sCs + rc * 10;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int k = 0; k < 6; k += 2) { const double2 t2 = qp_ld2(row + k); a0 += wu[k] * t2.x; a1 += wu[k + 1] * t2.y; }
            const double2 t6 = qp_ld2(row + 6), t8 = qp_ld2(row + 8);
            a0 += wu[6] * t6.x;
            const double own = wu[7] * t6.y + wu[8] * t8.x + wu[9] * t8.y;
            s[c] += st ? 0.0 : a0 + a1 + (leg == legc ? own : 0.0) + (c == lane ? delta : 0.0);
        }
        double rhs;
        {
            double ut;
            if (st) ut = sT[lane];
            else {
                ut = 0.0;
#pragma unroll
                for (int k = 0; k < 7; ++k) ut += cs[k] * sT[k];
#pragma unroll
                for (int m = 0; m < 3; ++m) ut += cs[7 + m] * sT[13 + 6 * leg + 3 + m];
            }
            rhs = gval - ut - carry;
        }
        if (j > 0) {
            // ---- Lo_j row = S_{j,j-1} row * L_{j-1,j-1}^-T  (left-looking, L rows broadcast) -----------------------------------
#pragma unroll
            for (int c = 0; c < G; ++c) {
                const double* Lr = sL + c * LS;
                double a0 = e[c], a1 = 0.0;
#pragma unroll
                for (int k = 0; k + 1 < c; k += 2) { const double2 t2 = qp_ld2(Lr + k); a0 -= e[k] * t2.x; a1 -= e[k + 1] * t2.y; }
                if (c & 1) a0 -= e[c - 1] * Lr[c - 1];
                e[c] = (a0 + a1) * Lr[G];
            }
            {
                double d0 = 0.0, d1 = 0.0;  // rhs_j -= Lo_j y_{j-1}
#pragma unroll
                for (int c = 0; c + 1 < G; c += 2) { const double2 t2 = qp_ld2(sY + c); d0 += e[c] * t2.x; d1 += e[c + 1] * t2.y; }
                d0 += e[G - 1] * sY[G - 1];
                rhs -= d0 + d1;
            }
            if (act) {
#pragma unroll
                for (int c = 0; c + 1 < G; c += 2) qp_st2(sO + lane * LS + c, e[c], e[c + 1]);
                qp_st2(sO + lane * LS + G - 1, e[G - 1], 0.0);
            }
            __syncwarp();
            // ---- S_jj -= Lo_j Lo_j^T on the FP64 tensor cores: ten lower 8x8 tiles x eight k-steps of mma.m8n8k4 -------------------------
            // fragment (I, kk) = Lo[8 I + lane / 4][4 kk + lane % 4] serves as the A operand of tile row I and the B operand of tile column I
            {
                double acc[10][2];
#pragma unroll
                for (int t = 0; t < 10; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
                const int fr = lane >> 2, fq = lane & 3;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    double f[4];
#pragma unroll
                    for (int I = 0; I < 4; ++I) {
                        f[I] = sO[min(8 * I + fr, G - 1) * LS + 4 * kk + fq];  // rows 29..31 shadow row 28 (their tiles' rows are dropped)
                        if (kk == 7 && fq != 0) f[I] = 0.0;                     // k = 29..31 do not exist
                    }
#pragma unroll
                    for (int I = 0; I < 4; ++I)
#pragma unroll
                        for (int J = 0; J <= I; ++J) qp_dmma(acc[(I * (I + 1)) / 2 + J][0], acc[(I * (I + 1)) / 2 + J][1], f[I], f[J]);
                }
                __syncwarp();  // all fragments are read: the product overwrites the Lo image, then every lane subtracts its row
#pragma unroll
                for (int I = 0; I < 4; ++I)
#pragma unroll
                    for (int J = 0; J <= I; ++J) {
                        const int row = 8 * I + fr, col = 8 * J + 2 * fq;
                        if (row < G && col < LS) qp_st2(sO + row * LS + col, acc[(I * (I + 1)) / 2 + J][0], acc[(I * (I + 1)) / 2 + J][1]);
                    }
                __syncwarp();
                const double* mine = sO + min(lane, G - 1) * LS;
#pragma unroll
                for (int c = 0; c + 1 < G; c += 2) { const double2 t2 = qp_ld2(mine + c); s[c] -= t2.x; s[c + 1] -= t2.y; }
                s[G - 1] -= mine[G - 1];
            }
            __syncwarp();  // every lane is done with the Lo image
        }
        // ---- the other region is free until the next stage's TRSM: fetch stage j+1 into it while this stage factorises ----------------
        if (live) {
            qp_fetch_stage(rec, L, N, j + 1, sO, sSt, lane, true);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // ---- Cholesky of S_jj fused with y_j = L_jj^-1 rhs_j: one column per step ---------------------------------------------------
        double rk = rhs, inv_own = 0.0;
#pragma unroll
        for (int c = 0; c < G; ++c) {
            const double rinv = rsqrt(__shfl_sync(FULL, s[c], c));
            const double l = lane >= c ? s[c] * rinv : 0.0;
            s[c] = l;
            if (lane == c) inv_own = rinv;
            double* col = sC + (c & 1) * 32;
            col[lane] = l;
            const double yc = __shfl_sync(FULL, rk, c) * rinv;
            if (lane == c) rk = yc;
            else if (lane > c) rk -= l * yc;
            __syncwarp();
            if (c + 1 < G) {
                if ((c + 1) & 1) s[c + 1] -= l * col[c + 1];
#pragma unroll
                for (int cc = (c + 2) & ~1; cc + 1 < G + 1; cc += 2) {
                    const double2 t2 = qp_ld2(col + cc);
                    s[cc] -= l * t2.x;
                    if (cc + 1 < G) s[cc + 1] -= l * t2.y;
                }
            }
        }
        // ---- L_jj (rows from registers) and y_j leave: the L image through the TMA engine -------------------------------------------
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (act) {
#pragma unroll
            for (int c = 0; c + 1 < G; c += 2) qp_st2(sL + lane * LS + c, s[c], s[c + 1]);
            qp_st2(sL + lane * LS + G - 1, s[G - 1], inv_own);
            sY[lane] = rk;
            ws[(long long)j * Q::WS_GROUP + G * LS + lane] = rk;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            bulk_store(ws + (long long)j * Q::WS_GROUP, sL, G * LS * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        // ---- stage j+1: S_{j+1,j} row = (V P^-1 U^T) row ;  carry rows (V P^-1 V^T) and V t_j -------------------------------------
        if (live) {
            double w[37];
            {
                double v[37];
                if (st) {
#pragma unroll
                    for (int k = 0; k < 36; k += 2) { const double2 t2 = qp_ld2(sA + lane * LA + k); v[k] = t2.x; v[k + 1] = t2.y; }
                    v[36] = sA[lane * LA + 36];
                } else {
                    double cp[10];
#pragma unroll
                    for (int k = 0; k < 10; k += 2) { const double2 t2 = qp_ld2(sCp + r * 10 + k); cp[k] = t2.x; cp[k + 1] = t2.y; }
#pragma unroll
                    for (int k = 0; k < 13; ++k) v[k] = k < 7 ? cp[k] : 0.0;
#pragma unroll
                    for (int lg = 0; lg < 4; ++lg)
#pragma unroll
                        for (int m = 0; m < 3; ++m) { v[13 + 6 * lg + m] = 0.0; v[13 + 6 * lg + 3 + m] = leg == lg ? cp[7 + m] : 0.0; }
                }
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int k = 0; k < 36; k += 2) { const double2 t2 = qp_ld2(sT + k); c0 += v[k] * t2.x; c1 += v[k + 1] * t2.y; }
                carry = c0 + c1 + v[36] * sT[36];
#pragma unroll
                for (int k = 0; k < 13; ++k) w[k] = v[k] * sP[k];
#pragma unroll
                for (int blk = 0; blk < 8; ++blk) {
                    const double* Bm = sP + 14 + 10 * blk;
                    const double2 b0 = qp_ld2(Bm), b2 = qp_ld2(Bm + 2), b4 = qp_ld2(Bm + 4), b6 = qp_ld2(Bm + 6);
                    const double b8 = Bm[8];
                    const double x0 = v[13 + 3 * blk], x1 = v[14 + 3 * blk], x2 = v[15 + 3 * blk];
                    w[13 + 3 * blk + 0] = x0 * b0.x + x1 * b2.y + x2 * b6.x;
                    w[13 + 3 * blk + 1] = x0 * b0.y + x1 * b4.x + x2 * b6.y;
                    w[13 + 3 * blk + 2] = x0 * b2.x + x1 * b4.y + x2 * b8;
                }
            }
#pragma unroll
            for (int c = 0; c < 13; ++c) {
                e[c] = w[c];
                const double* row = sA + c * LA;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int k = 0; k < 36; k += 4) {
                    const double2 t0 = qp_ld2(row + k), t2 = qp_ld2(row + k + 2);
                    a0 += w[k] * t0.x; a1 += w[k + 1] * t0.y; a2 += w[k + 2] * t2.x; a3 += w[k + 3] * t2.y;
                }
                s[c] = (a0 + a1) + (a2 + a3) + w[36] * row[36];
            }
#pragma unroll
            for (int c = 13; c < G; ++c) {
                const int rc = c - 13, legc = rc >> 2, o = 13 + 6 * legc + 3;
                {
                    const double* row = sCs + rc * 10;
                    const double2 t0 = qp_ld2(row), t2 = qp_ld2(row + 2), t4 = qp_ld2(row + 4), t6 = qp_ld2(row + 6), t8 = qp_ld2(row + 8);
                    e[c] = (w[0] * t0.x + w[1] * t0.y + w[2] * t2.x + w[3] * t2.y) + (w[4] * t4.x + w[5] * t4.y + w[6] * t6.x) +
                           (w[o] * t6.y + w[o + 1] * t8.x + w[o + 2] * t8.y);
                }
                {
                    const double* row = sCp + rc * 10;
                    const double2 t0 = qp_ld2(row), t2 = qp_ld2(row + 2), t4 = qp_ld2(row + 4), t6 = qp_ld2(row + 6), t8 = qp_ld2(row + 8);
                    s[c] = (w[0] * t0.x + w[1] * t0.y + w[2] * t2.x + w[3] * t2.y) + (w[4] * t4.x + w[5] * t4.y + w[6] * t6.x) +
                           (w[o] * t6.y + w[o + 1] * t8.x + w[o + 2] * t8.y);
                }
            }
        }
        double* const tmp = sD; sD = sO; sO = tmp;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();

    // ================================================================ backward sweep: nu_j, then d_j
    // sY holds nu_{j+1} (zero beyond the horizon).  The big regions become two (packed L_jj, stage data) pairs.
    sY[lane] = 0.0;
    double* __restrict__ step = step_all + b * ld_step;
    auto fetch_back = [&](int j, int buf) {
        qp_fetch_stage(rec, L, N, j, sm + Q::bD0 + buf * Q::dTotal, sSt, lane, false);
        const double* wj = ws + (long long)j * Q::WS_GROUP;
        double* sLp = sm + buf * Q::LP;
        if (act) {
            qp_cp8(sSt + Q::stG + lane, wj + G * LS + lane);  // y_j
            qp_cp8(sLp + 435 + lane, wj + lane * LS + G);      // 1 / L_ii
        }
#pragma unroll
        for (int i = 0; i < G; ++i)
            if (lane <= i) qp_cp8(sLp + (i * (i + 1)) / 2 + lane, wj + i * LS + lane);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __syncwarp();
    fetch_back(N, 0);
    for (int j = N; j >= 0; --j) {
        const int nz = j < N ? 37 : 13, buf = (N - j) & 1;
        const double* sLp = sm + buf * Q::LP;
        const double* sDb = sm + Q::bD0 + buf * Q::dTotal;
        const double *sA = sDb + Q::dA, *sCs = sDb + Q::dCs, *sCp = sDb + Q::dCp;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const double yj = act ? sSt[Q::stG + lane] : 0.0;
        for (int k = lane; k < 37; k += 32) sQ[k] = sSt[Q::stQ + k];
        qp_build_pinv(sSt, sP, j < N, lane);
        if (j > 0) fetch_back(j - 1, buf ^ 1);
        // a = V_j^T nu_{j+1}  (entries k = lane, lane + 32) -> sC ;  P^-1 a -> sT
        for (int k = lane; k < 37; k += 32) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 13; ++i) acc += sA[i * LA + k] * sY[i];
            sC[k] = acc + qp_contact_T(sCp, sY, k);
        }
        __syncwarp();
        apply_pinv_warp(sP, sC, sT, lane);
        // z = U_j P^-1 a (row per lane), then z' = L_jj^-1 z
        double cs[10];
#pragma unroll
        for (int k = 0; k < 10; k += 2) { const double2 t2 = qp_ld2(sCs + r * 10 + k); cs[k] = t2.x; cs[k + 1] = t2.y; }
        double rk;
        if (st) rk = sT[lane];
        else {
            rk = 0.0;
#pragma unroll
            for (int k = 0; k < 7; ++k) rk += cs[k] * sT[k];
#pragma unroll
            for (int m = 0; m < 3; ++m) rk += cs[7 + m] * sT[13 + 6 * leg + 3 + m];
        }
        {
            const double* own = sLp + (min(lane, G - 1) * (min(lane, G - 1) + 1)) / 2;  // own packed row: entries 0..lane
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const double zi = __shfl_sync(FULL, rk, i) * sLp[435 + i];
                if (lane == i) rk = zi;
                else if (lane > i && act) rk -= own[i] * zi;
            }
        }
        rk = yj - rk;  // y_j - Lo_{j+1}^T nu_{j+1}
#pragma unroll
        for (int i = G - 1; i >= 0; --i) {  // nu_j = L_jj^-T rhs, column-oriented from the last row up; lane k owns entry k
            const double ni = __shfl_sync(FULL, rk, i) * sLp[435 + i];
            if (lane == i) rk = ni;
            else if (lane < i) rk -= sLp[(i * (i + 1)) / 2 + lane] * ni;
        }
        __syncwarp();
        if (act) sY[lane] = rk;
        __syncwarp();
        // d_j = -P^-1 (q_j + a + U_j^T nu_j)
        for (int k = lane; k < 37; k += 32) sQ[k] += sC[k] + (k < 13 ? sY[k] : 0.0) + qp_contact_T(sCs, sY, k);
        __syncwarp();
        apply_pinv_warp(sP, sQ, sT, lane);
        for (int k = lane; k < nz; k += 32) {
            const long long dst = k < 13 ? Q::NX * j + k : nX + Q::NU * j + (k - 13);
            step[dst] = -sT[k];
        }
        if (mult_all && act) {  // multipliers in the reference's row order [x0 | defects | contact rows]
            double* mult = mult_all + b * ld_mult;
            if (lane < 13) mult[Q::NX * j + lane] = rk;
            else if (j < N) mult[nX + 16 * j + (lane - 13)] = rk;
        }
        __syncwarp();
    }
}

}  // namespace ub
