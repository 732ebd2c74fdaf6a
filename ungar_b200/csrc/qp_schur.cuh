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
// Lanes own rows: lane i < 29 holds row i of every 29x.. block; operands of the small products are broadcast from shared memory.
#pragma once

#include "sweep.cuh"

namespace ub {

struct QpShape {
    static constexpr int NX = 13, NU = 24, NZ = 37, TRI = 703, G = 29, LD = 37;  // G rows per group; LD = odd row stride
    // per-warp shared memory (doubles)
    static constexpr int oU = 0, oV = oU + G * LD, oW = oV + G * LD, oS = oW + G * LD, oE = oS + G * G, oL = oE + G * G,
                         oP = oL + G * G, oT = oP + 13 + 8 * 9, total = ((oT + 2 * NZ + 2 * G + 3) & ~3);
    static constexpr int WARPS = 4;
    static constexpr int SMEM_BYTES = WARPS * total * 8;
    // per-trajectory global workspace (doubles): per group  Ld (G*G) | Lo (G*G) | y (G)
    static constexpr int WS_GROUP = 2 * G * G + G;
};

// Closed-form inverse of a symmetric 3x3 block (row-major 9 entries out).
__device__ __forceinline__ void inv_sym3(double a, double b, double c, double d, double e, double f, double* out) {
    // [a b c; b d e; c e f]
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02, id = 1.0 / det;
    out[0] = c00 * id; out[1] = c01 * id; out[2] = c02 * id;
    out[3] = out[1];   out[4] = (a * f - c * c) * id; out[5] = (b * c - a * e) * id;
    out[6] = out[2];   out[7] = out[5]; out[8] = (a * d - b * b) * id;
}

// y = P^-1 x cooperatively: lanes 0..12 the diagonal state part, lanes 13..20 one 3x3 input block each.
__device__ __forceinline__ void apply_pinv_warp(const double* __restrict__ pinv, const double* __restrict__ x, double* __restrict__ y,
                                                int nz, int lane) {
    if (lane < 13) y[lane] = pinv[lane] * x[lane];
    else if (lane < 21) {
        const int blk = lane - 13;
        const double* m = pinv + 13 + 9 * blk;
        const double* xv = x + 13 + 3 * blk;
        const bool on = nz > 13;
        y[13 + 3 * blk + 0] = on ? m[0] * xv[0] + m[1] * xv[1] + m[2] * xv[2] : 0.0;
        y[13 + 3 * blk + 1] = on ? m[3] * xv[0] + m[4] * xv[1] + m[5] * xv[2] : 0.0;
        y[13 + 3 * blk + 2] = on ? m[6] * xv[0] + m[7] * xv[1] + m[8] * xv[2] : 0.0;
    }
    __syncwarp();
}

// y = P^-1 x for one stage (x, y: 37 entries; pinv: 13 reciprocals + 8 blocks of 9).  Executed by one lane.
__device__ __forceinline__ void apply_pinv(const double* __restrict__ pinv, const double* __restrict__ x, double* __restrict__ y, int nz) {
#pragma unroll
    for (int i = 0; i < 13; ++i) y[i] = pinv[i] * x[i];
    if (nz > 13) {
#pragma unroll
        for (int blk = 0; blk < 8; ++blk) {
            const double* m = pinv + 13 + 9 * blk;
            const double* xv = x + 13 + 3 * blk;
            y[13 + 3 * blk + 0] = m[0] * xv[0] + m[1] * xv[1] + m[2] * xv[2];
            y[13 + 3 * blk + 1] = m[3] * xv[0] + m[4] * xv[1] + m[5] * xv[2];
            y[13 + 3 * blk + 2] = m[6] * xv[0] + m[7] * xv[1] + m[8] * xv[2];
        }
    } else {
#pragma unroll
        for (int i = 13; i < 37; ++i) y[i] = 0.0;
    }
}

// Loads stage j of a trajectory into shared memory: U_j (29 x 37), V_j (29 x 37, rows of nu_{j+1} on w_j), P_j^-1.
// All 32 lanes participate.  For j = N only the state part exists (U = [I_x], V = 0).
__device__ __forceinline__ void load_stage(const double* __restrict__ rec, const RecLayout& L, int N, int j, double* __restrict__ sU,
                                           double* __restrict__ sV, double* __restrict__ sP, int lane) {
    using Q = QpShape;
    for (int e = lane; e < Q::G * Q::LD; e += 32) { sU[e] = 0.0; sV[e] = 0.0; }
    __syncwarp();
    if (lane < 13) sU[lane * Q::LD + lane] = 1.0;  // I_x rows (dynamics defect of stage j-1, or x_0 - x_measured)
    if (j < N) {
        // contact rows of stage j on w_j (Cs_j) and contact rows of stage j+1 on w_j (Cp_{j+1})
        for (int e = lane; e < 16 * 10; e += 32) {
            const int row = e / 10, col = e % 10, leg = row >> 2;
            const int dst = col < 7 ? col : 13 + 6 * leg + 3 + (col - 7);
            sU[(13 + row) * Q::LD + dst] = rec[L.C + (long long)j * 320 + row * 20 + col];
            if (j + 1 < N) sV[(13 + row) * Q::LD + dst] = rec[L.C + (long long)(j + 1) * 320 + row * 20 + 10 + col];
        }
        for (int e = lane; e < 13 * 37; e += 32) sV[(e / 37) * Q::LD + e % 37] = rec[L.A + (long long)j * 481 + e];
        // P_j^-1 from the packed upper triangle of H_j
        const double* H = rec + L.H + (long long)j * Q::TRI;
        if (lane < 13) sP[lane] = 1.0 / H[lane * 37 - (lane * (lane - 1)) / 2];
        else if (lane < 21) {
            const int a = 13 + 3 * (lane - 13);
            const int d0 = a * 37 - (a * (a - 1)) / 2, d1 = (a + 1) * 37 - ((a + 1) * a) / 2, d2 = (a + 2) * 37 - ((a + 2) * (a + 1)) / 2;
            inv_sym3(H[d0], H[d0 + 1], H[d0 + 2], H[d1], H[d1 + 1], H[d2], sP + 13 + 9 * (lane - 13));
        }
    } else {
        const double* H = rec + L.HN;
        if (lane < 13) sP[lane] = 1.0 / H[lane * 13 - (lane * (lane - 1)) / 2];
    }
    __syncwarp();
}

// W = M P^-1 row by row (lane i < 29 owns row i).
__device__ __forceinline__ void times_pinv(const double* __restrict__ sM, const double* __restrict__ sP, double* __restrict__ sW,
                                           int lane, int nz) {
    using Q = QpShape;
    if (lane < Q::G) {
        double x[37], y[37];
#pragma unroll
        for (int k = 0; k < 37; ++k) x[k] = sM[lane * Q::LD + k];
        apply_pinv(sP, x, y, nz);
#pragma unroll
        for (int k = 0; k < 37; ++k) sW[lane * Q::LD + k] = y[k];
    }
    __syncwarp();
}

// out[i][c] (+)= sum_k W[i][k] M[c][k]   (29 x 29; lane i owns row i; M rows broadcast from shared memory)
__device__ __forceinline__ void gemm_wmT(const double* __restrict__ sW, const double* __restrict__ sM, double* __restrict__ out,
                                         int lane, bool accumulate) {
    using Q = QpShape;
    if (lane < Q::G) {
        double w[37];
#pragma unroll
        for (int k = 0; k < 37; ++k) w[k] = sW[lane * Q::LD + k];
        for (int c = 0; c < Q::G; ++c) {
            double acc = accumulate ? out[lane * Q::G + c] : 0.0;
#pragma unroll
            for (int k = 0; k < 37; ++k) acc += w[k] * sM[c * Q::LD + k];
            out[lane * Q::G + c] = acc;
        }
    }
    __syncwarp();
}

// Forward sweep (factorisation + forward substitution), then backward sweep (multipliers and step).
__global__ void __launch_bounds__(QpShape::WARPS * 32)
qp_schur_kernel(const double* __restrict__ rec_all, long long ld_rec, double* __restrict__ ws_all, double* __restrict__ step_all,
                long long ld_step, double* __restrict__ mult_all, long long ld_mult, int N, long long batch, RecLayout L, double delta) {
    using Q = QpShape;
    constexpr int G = Q::G, LD = Q::LD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* const sm = reinterpret_cast<double*>(smem_raw) + wib * Q::total;
    double *sU = sm + Q::oU, *sV = sm + Q::oV, *sW = sm + Q::oW, *sS = sm + Q::oS, *sE = sm + Q::oE, *sL = sm + Q::oL,
           *sP = sm + Q::oP, *sT = sm + Q::oT;  // sT: t_j = P^-1 q_j (37) | q_j (37) | y_prev (29) | scratch (29)
    double* const sQ = sT + 37;
    double* const sY = sQ + 37;
    double* const sR = sY + G;
    const long long b = (long long)blockIdx.x * Q::WARPS + wib;
    if (b >= batch) return;
    const double* __restrict__ rec = rec_all + b * ld_rec;
    double* __restrict__ ws = ws_all + b * (long long)(N + 1) * Q::WS_GROUP;
    const int nX = Q::NX * (N + 1);

    // ================================================================ forward sweep over the groups nu_0 .. nu_N
    // carry = V_{j-1} P_{j-1}^-1 V_{j-1}^T (into S_jj) and V_{j-1} t_{j-1} (into the right-hand side): computed at the end of stage j-1
    for (int e = lane; e < G * G; e += 32) sS[e] = 0.0;
    if (lane < G) sR[lane] = 0.0;
    __syncwarp();
    for (int j = 0; j <= N; ++j) {
        const int nz = j < N ? 37 : 13;
        load_stage(rec, L, N, j, sU, sV, sP, lane);
        // q_j and t_j = P_j^-1 q_j
        for (int k = lane; k < 37; k += 32)
            sQ[k] = k < 13 ? rec[L.grad + Q::NX * j + k] : (j < N ? rec[L.grad + nX + Q::NU * j + (k - 13)] : 0.0);
        __syncwarp();
        apply_pinv_warp(sP, sQ, sT, nz, lane);
        // S_jj += U P^-1 U^T + delta I ;   rhs_j = -(b_j + U t_j + carry),  b_j = -g(rows of nu_j)
        times_pinv(sU, sP, sW, lane, nz);
        gemm_wmT(sW, sU, sS, lane, true);
        if (lane < G) {
            sS[lane * G + lane] += delta;
            double ut = 0.0;
#pragma unroll
            for (int k = 0; k < 37; ++k) ut += sU[lane * LD + k] * sT[k];
            double gval = 0.0;
            if (lane < 13) gval = rec[L.g + Q::NX * j + lane];
            else if (j < N) gval = rec[L.g + nX + 16 * j + (lane - 13)];
            sR[lane] = gval - ut - sR[lane];  // -(b + U t + carry) with b = -g
        }
        __syncwarp();
        // off-diagonal block S_{j,j-1} is in sE (from the previous stage): L_{j,j-1} = S_{j,j-1} L_{j-1,j-1}^-T ;  S_jj -= Lo Lo^T
        if (j > 0) {
            if (lane < G) {  // row `lane` of Lo, kept in registers (fully unrolled: no local memory)
                double row[G];
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    double acc = sE[lane * G + c];
#pragma unroll
                    for (int k = 0; k < c; ++k) acc -= row[k] * sL[c * G + k];
                    row[c] = acc / sL[c * G + c];
                }
                double dot = 0.0;
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    sE[lane * G + c] = row[c];
                    dot += row[c] * sY[c];
                }
                sR[lane] -= dot;  // rhs_j - Lo y_{j-1}
            }
            __syncwarp();
            if (lane < G) {
                double lo[G];
#pragma unroll
                for (int k = 0; k < G; ++k) lo[k] = sE[lane * G + k];
                for (int c = 0; c <= lane; ++c) {  // only the lower triangle of S_jj is used by the Cholesky
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < G; ++k) acc += lo[k] * sE[c * G + k];
                    sS[lane * G + c] -= acc;
                }
            }
            __syncwarp();
            for (int e = lane; e < G * G; e += 32) ws[(long long)j * Q::WS_GROUP + G * G + e] = sE[e];  // Lo_j
        }
        // Cholesky of S_jj (right-looking; lane i owns row i), result in sL (lower)
        for (int c = 0; c < G; ++c) {
            const double piv = sqrt(sS[c * G + c]);
            __syncwarp();
            if (lane < G && lane >= c) sL[lane * G + c] = lane == c ? piv : sS[lane * G + c] / piv;
            __syncwarp();
            if (lane < G && lane > c) {
                const double lic = sL[lane * G + c];
                for (int cc = c + 1; cc <= lane; ++cc) sS[lane * G + cc] -= lic * sL[cc * G + c];
            }
            __syncwarp();
        }
        // forward substitution y_j = L_jj^-1 rhs_j (serial over rows, lane 0) — 29 x 29 / 2 operations
        {  // column-oriented: y_i is final once columns < i have been eliminated; lane k owns entry k
            double rk = lane < G ? sR[lane] : 0.0;
            for (int i = 0; i < G; ++i) {
                const double yi = __shfl_sync(0xffffffffu, rk, i) / sL[i * G + i];
                if (lane == i) rk = yi;
                else if (lane > i && lane < G) rk -= sL[lane * G + i] * yi;
            }
            if (lane < G) sY[lane] = rk;
        }
        __syncwarp();
        for (int e = lane; e < G * G; e += 32) {
            const int i = e / G, c = e % G;
            ws[(long long)j * Q::WS_GROUP + e] = c <= i ? sL[e] : 0.0;  // Ld_j
        }
        if (lane < G) ws[(long long)j * Q::WS_GROUP + 2 * G * G + lane] = sY[lane];
        // prepare stage j+1: S_{j+1,j} = V P^-1 U^T -> sE ;  carry V P^-1 V^T -> sS ;  V t_j -> sR
        if (j < N) {
            times_pinv(sV, sP, sW, lane, nz);
            gemm_wmT(sW, sU, sE, lane, false);
            gemm_wmT(sW, sV, sS, lane, false);
            if (lane < G) {
                double vt = 0.0;
#pragma unroll
                for (int k = 0; k < 37; ++k) vt += sV[lane * LD + k] * sT[k];
                sR[lane] = vt;
            }
            __syncwarp();
        }
    }

    // ================================================================ backward sweep: nu_j, then d_j
    // sY holds nu_{j+1} (zero beyond the horizon); sW row 0 is reused for the stage vector v = q + U^T nu_j + V^T nu_{j+1}
    if (lane < G) sY[lane] = 0.0;
    __syncwarp();
    double* __restrict__ step = step_all + b * ld_step;
    for (int j = N; j >= 0; --j) {
        const int nz = j < N ? 37 : 13;
        load_stage(rec, L, N, j, sU, sV, sP, lane);
        for (int e = lane; e < G * G; e += 32) sL[e] = ws[(long long)j * Q::WS_GROUP + e];
        if (j < N)
            for (int e = lane; e < G * G; e += 32) sE[e] = ws[(long long)(j + 1) * Q::WS_GROUP + G * G + e];  // Lo_{j+1}
        if (lane < G) sR[lane] = ws[(long long)j * Q::WS_GROUP + 2 * G * G + lane];                          // y_j
        __syncwarp();
        // V^T nu_{j+1} needs nu_{j+1} (still in sY) BEFORE it is overwritten: accumulate the stage vector first
        for (int k = lane; k < 37; k += 32) {
            double acc = k < 13 ? rec[L.grad + Q::NX * j + k] : (j < N ? rec[L.grad + nX + Q::NU * j + (k - 13)] : 0.0);
            if (j < N)
                for (int i = 0; i < G; ++i) acc += sV[i * LD + k] * sY[i];
            sQ[k] = acc;
        }
        // rhs = y_j - Lo_{j+1}^T nu_{j+1}
        if (lane < G && j < N) {
            double acc = 0.0;
            for (int i = 0; i < G; ++i) acc += sE[i * G + lane] * sY[i];
            sR[lane] -= acc;
        }
        __syncwarp();
        {  // nu_j = L_jj^-T rhs, column-oriented from the last row up; lane k owns entry k
            double rk = lane < G ? sR[lane] : 0.0;
            for (int i = G - 1; i >= 0; --i) {
                const double ni = __shfl_sync(0xffffffffu, rk, i) / sL[i * G + i];
                if (lane == i) rk = ni;
                else if (lane < i) rk -= sL[i * G + lane] * ni;
            }
            __syncwarp();
            if (lane < G) sY[lane] = rk;
        }
        __syncwarp();
        for (int k = lane; k < 37; k += 32) {
            double acc = sQ[k];
            for (int i = 0; i < G; ++i) acc += sU[i * LD + k] * sY[i];
            sQ[k] = acc;
        }
        __syncwarp();
        apply_pinv_warp(sP, sQ, sT, nz, lane);
        for (int k = lane; k < nz; k += 32) {
            const long long dst = k < 13 ? Q::NX * j + k : nX + Q::NU * j + (k - 13);
            step[dst] = -sT[k];
        }
        if (mult_all && lane < G) {  // multipliers in the reference's row order [x0 | defects | contact rows]
            double* mult = mult_all + b * ld_mult;
            if (lane < 13) mult[Q::NX * j + lane] = sY[lane];
            else if (j < N) mult[nX + 16 * j + (lane - 13)] = sY[lane];
        }
        __syncwarp();
    }
}

}  // namespace ub
