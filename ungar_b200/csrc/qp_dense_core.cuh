// Dense core of the twisted QP solve (qp_twisted.cuh) on the FP64 tensor cores: the 29 x 29 blocks of the block-tridiagonal Schur
// complement live in the C-FRAGMENT layout of mma.sync.m8n8k4.f64 — a 4 x 4 grid of 8 x 8 tiles, lane (g = lane / 4, t = lane % 4)
// holds entries (g, 2t) and (g, 2t + 1) of every tile — and every O(n^3) update is a rank-4 DMMA:
//
//   qt3_cholesky    right-looking Cholesky in PANELS of 4 columns.  Per panel: the panel's columns go from the fragments to a
//                   row-per-lane buffer in shared memory (lane = row), every lane factors the 4 x 4 diagonal block redundantly and
//                   solves its own row against it (a dozen FMAs), the forward substitution of the right-hand side rides along, the
//                   panel of L goes to the factor image (shared memory + workspace), and the trailing matrix takes ONE DMMA per
//                   8 x 8 tile:  S(I, J) -= F_I F_J^T  with  F_I[lane] = L[8 I + g][4 p + t]  (30 DMMAs instead of 406 DFMAs + 217
//                   broadcast loads per lane).
//   qt3_trsm_syrk   coupling rows  Lo = E L^-T  in the same panels — own row against the 4 x 4 block, then E(I, J) -= Lo_I F_J^T on the
//                   tensor cores — and, with the same Lo fragments, the rank update the NEXT group needs:  P += Lo_I Lo_J^T.  Lo itself
//                   is never stored: the next group receives P in registers (20 doubles) and -Lo y in one scalar.
//
// Rows / columns 29 .. 31 are phantom (identity on the diagonal).  Factor image (doubles): panel p (columns 4p .. 4p + 3) holds rows
// 4p .. 28, four entries per row, at fb(p); 1 / L_ii at fRI; y at fY.  Everything below is plain C++ over shared-memory pointers plus
// qp_dmma / __syncwarp / __shfl_sync, so tests/host/qp_core_host.cpp runs it on 32 host threads in lockstep.
#pragma once

namespace ub {

struct Qt3 {
    static constexpr int G = 29;
    __host__ __device__ static constexpr int fb(int p) { return 116 * p - 8 * p * (p - 1); }  // sum_{q < p} 4 (29 - 4 q)
    static constexpr int fRI = 480, fY = 512, WS_GROUP = 544;
    // staging behind the factor image (all inside the 870-double LO region of qp_twisted.cuh)
    static constexpr int oPB = 544, oRK = 672, oLP = 672, oST = 544, REGION = 800;  // PB [32][4]; RK [32] / LP [32][4]; ST [32][8] aliases PB + LP
    __host__ __device__ static constexpr int tile(int I, int J) { return (I * (I + 1)) / 2 + J; }  // lower tiles, J <= I
};
static_assert(Qt3::fb(8) == 480, "factor image");

// Row-per-lane block (entries 0 .. 28 of row `lane`) -> lower C-fragment tiles, through the [32][8] staging buffer, one tile column at a
// time.  sign = +1: S = rows;  ACC: S = rows - S_in (the incoming fragments hold the rank update P to subtract).
template <bool ACC>
__device__ __forceinline__ void qt3_rows_to_lower_frags(const double* row, double* __restrict__ st, double (*S)[2], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int J = 0; J < 4; ++J) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const int c = 8 * J + k;
            const double a = (c < Qt3::G && lane < Qt3::G) ? row[c < Qt3::G ? c : 0] : (lane == c ? 1.0 : 0.0);
            const double b = (c + 1 < Qt3::G && lane < Qt3::G) ? row[c + 1 < Qt3::G ? c + 1 : 0] : (lane == c + 1 ? 1.0 : 0.0);
            qp_st2(st + lane * 8 + k, a, b);
        }
        __syncwarp();
#pragma unroll
        for (int I = J; I < 4; ++I) {
            const double2 v = qp_ld2(st + (8 * I + g) * 8 + 2 * t);
            double* s = S[Qt3::tile(I, J)];
            if (ACC) { s[0] = v.x - s[0]; s[1] = v.y - s[1]; }
            else { s[0] = v.x; s[1] = v.y; }
        }
    }
    __syncwarp();
}

// Same for a full (non-symmetric) 29 x 29 block: 16 tiles E[4 I + J].
__device__ __forceinline__ void qt3_rows_to_frags(const double* row, double* __restrict__ st, double (*E)[2], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int J = 0; J < 4; ++J) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const int c = 8 * J + k;
            const double a = (c < Qt3::G && lane < Qt3::G) ? row[c < Qt3::G ? c : 0] : 0.0;
            const double b = (c + 1 < Qt3::G && lane < Qt3::G) ? row[c + 1 < Qt3::G ? c + 1 : 0] : 0.0;
            qp_st2(st + lane * 8 + k, a, b);
        }
        __syncwarp();
#pragma unroll
        for (int I = 0; I < 4; ++I) {
            const double2 v = qp_ld2(st + (8 * I + g) * 8 + 2 * t);
            E[4 * I + J][0] = v.x;
            E[4 * I + J][1] = v.y;
        }
    }
    __syncwarp();
}

// 4 x 4 lower-triangular factor of a symmetric block given by its lower entries d[i][j] (j <= i); r[i] = 1 / l[i][i].
struct Qt3Block {
    double l10, l20, l21, l30, l31, l32, r0, r1, r2, r3;
};
__device__ __forceinline__ void qt3_factor4(const double* __restrict__ pb4, Qt3Block& B) {  // pb4: 4 rows x 4 (row-major), lower part used
    const double2 a0 = qp_ld2(pb4), a1 = qp_ld2(pb4 + 4), a2 = qp_ld2(pb4 + 8), a2b = qp_ld2(pb4 + 10), a3 = qp_ld2(pb4 + 12), a3b = qp_ld2(pb4 + 14);
    const double d00 = a0.x, d10 = a1.x, d11 = a1.y, d20 = a2.x, d21 = a2.y, d22 = a2b.x, d30 = a3.x, d31 = a3.y, d32 = a3b.x, d33 = a3b.y;
    B.r0 = rsqrt(d00);
    B.l10 = d10 * B.r0; B.l20 = d20 * B.r0; B.l30 = d30 * B.r0;
    B.r1 = rsqrt(d11 - B.l10 * B.l10);
    B.l21 = (d21 - B.l20 * B.l10) * B.r1; B.l31 = (d31 - B.l30 * B.l10) * B.r1;
    B.r2 = rsqrt(d22 - B.l20 * B.l20 - B.l21 * B.l21);
    B.l32 = (d32 - B.l30 * B.l20 - B.l31 * B.l21) * B.r2;
    B.r3 = rsqrt(d33 - B.l30 * B.l30 - B.l31 * B.l31 - B.l32 * B.l32);
}
// x <- x L_dd^-T for a 4-entry row x
__device__ __forceinline__ void qt3_solve4(const Qt3Block& B, double& x0, double& x1, double& x2, double& x3) {
    x0 *= B.r0;
    x1 = (x1 - x0 * B.l10) * B.r1;
    x2 = (x2 - x0 * B.l20 - x1 * B.l21) * B.r2;
    x3 = (x3 - x0 * B.l30 - x1 * B.l31 - x2 * B.l32) * B.r3;
}

// Cholesky of the block in the lower C-fragment tiles S[10][2], fused with y = L^-1 rhs (rhs = entry `lane`, 0 for lanes >= 29).
// `f`: factor image + staging (shared memory, Qt3::REGION doubles); `wsg`: the group's slot in the workspace.  Returns y_lane.
__device__ __forceinline__ double qt3_cholesky(double (*S)[2], double rhs, double* __restrict__ f, double* __restrict__ wsg, int lane) {
    constexpr int G = Qt3::G;
    const int g = lane >> 2, t = lane & 3;
    double* const pb = f + Qt3::oPB;
    double* const rkb = f + Qt3::oRK;
    double rk = rhs, y_own = 0.0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        constexpr int dummy = 0; (void)dummy;
        const int Jp = p >> 1, half = p & 1, c0 = 4 * p;
        // ---- panel columns: fragments -> row-per-lane buffer; right-hand side entries next to them
        if ((t >> 1) == half) {
#pragma unroll
            for (int I = Jp; I < 4; ++I) qp_st2(pb + (8 * I + g) * 4 + 2 * (t & 1), S[Qt3::tile(I, Jp)][0], S[Qt3::tile(I, Jp)][1]);
        }
        rkb[lane] = rk;
        __syncwarp();
        Qt3Block B;
        qt3_factor4(pb + c0 * 4, B);
        const double2 x01 = qp_ld2(pb + lane * 4), x23 = qp_ld2(pb + lane * 4 + 2);
        double x0 = x01.x, x1 = x01.y, x2 = x23.x, x3 = x23.y;
        qt3_solve4(B, x0, x1, x2, x3);
        // forward substitution inside the panel (every lane, redundantly) and on the own entry
        const double2 k01 = qp_ld2(rkb + c0), k23 = qp_ld2(rkb + c0 + 2);
        double y0 = k01.x, y1 = k01.y, y2 = k23.x, y3 = k23.y;
        qt3_solve4(B, y0, y1, y2, y3);  // L_dd y = rk_block  is the same recurrence as a row solve against L_dd^T
        const int i = lane - c0;        // row index inside / below the panel
        // rows of the block itself: zeros above the diagonal; rows above the panel: nothing
        if (i < 0) { x0 = x1 = x2 = x3 = 0.0; }
        else if (i < 4) {
            if (i < 1) x1 = 0.0;
            if (i < 2) x2 = 0.0;
            if (i < 3) x3 = 0.0;
            y_own = i == 0 ? y0 : (i == 1 ? y1 : (i == 2 ? y2 : y3));
        } else rk -= x0 * y0 + x1 * y1 + x2 * y2 + x3 * y3;
        // L panel -> factor image (shared memory and workspace); 1 / L_ii
        if (i >= 0 && lane < G) {
            double* dst = f + Qt3::fb(p) + i * 4;
            qp_st2(dst, x0, x1); qp_st2(dst + 2, x2, x3);
            double* wd = wsg + Qt3::fb(p) + i * 4;
            qp_st2(wd, x0, x1); qp_st2(wd + 2, x2, x3);
        }
        if (lane == 0) {
            qp_st2(f + Qt3::fRI + c0, B.r0, B.r1); qp_st2(f + Qt3::fRI + c0 + 2, B.r2, B.r3);
            qp_st2(wsg + Qt3::fRI + c0, B.r0, B.r1); qp_st2(wsg + Qt3::fRI + c0 + 2, B.r2, B.r3);
        }
        __syncwarp();
        // ---- trailing update on the tensor cores: S(I, J) -= F_I F_J^T,  F_I[lane] = L[8 I + g][4 p + t]
        if (p < 7) {
            const int Jlo = half ? Jp + 1 : Jp;
            double F[4];
#pragma unroll
            for (int I = 0; I < 4; ++I) {
                const int r = 8 * I + g - c0;
                F[I] = (I >= Jlo && r >= 0 && 8 * I + g < G) ? f[Qt3::fb(p) + (r < 0 ? 0 : r) * 4 + t] : 0.0;
            }
#pragma unroll
            for (int I = 0; I < 4; ++I)
#pragma unroll
                for (int J = 0; J <= I; ++J)
                    if (J >= Jlo) qp_dmma(S[Qt3::tile(I, J)][0], S[Qt3::tile(I, J)][1], -F[I], F[J]);
        }
    }
    return y_own;
}

// Coupling rows of the next group: Lo = E L^-T with the factor image `f` of qt3_cholesky (still in shared memory), P = Lo Lo^T in lower
// C-fragment tiles (returned in P[10][2]), and -(Lo y) for the lane's row (returned).  `e_row`: entries 0 .. 28 of row `lane` of E.
__device__ __forceinline__ double qt3_trsm_syrk(const double* e_row, double* __restrict__ f, const double* __restrict__ y, double (*P)[2], int lane) {
    constexpr int G = Qt3::G;
    const int g = lane >> 2, t = lane & 3;
    double* const pb = f + Qt3::oPB;
    double* const lp = f + Qt3::oLP;
    double E[16][2];
    qt3_rows_to_frags(e_row, f + Qt3::oST, E, lane);
#pragma unroll
    for (int q = 0; q < 10; ++q) { P[q][0] = 0.0; P[q][1] = 0.0; }
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int Jp = p >> 1, half = p & 1, c0 = 4 * p;
        if ((t >> 1) == half) {
#pragma unroll
            for (int I = 0; I < 4; ++I) qp_st2(pb + (8 * I + g) * 4 + 2 * (t & 1), E[4 * I + Jp][0], E[4 * I + Jp][1]);
        }
        __syncwarp();
        // the 4 x 4 diagonal block of L and its reciprocal diagonal, from the factor image
        Qt3Block B;
        {
            const double* d = f + Qt3::fb(p);
            const double2 r01 = qp_ld2(f + Qt3::fRI + c0), r23 = qp_ld2(f + Qt3::fRI + c0 + 2);
            B.r0 = r01.x; B.r1 = r01.y; B.r2 = r23.x; B.r3 = r23.y;
            if (p < 7) {
                const double2 a1 = qp_ld2(d + 4), a2 = qp_ld2(d + 8), a3 = qp_ld2(d + 12), a3b = qp_ld2(d + 14);
                B.l10 = a1.x; B.l20 = a2.x; B.l21 = a2.y; B.l30 = a3.x; B.l31 = a3.y; B.l32 = a3b.x;
            } else {  // columns 29 .. 31 are phantom
                B.l10 = B.l20 = B.l21 = B.l30 = B.l31 = B.l32 = 0.0;
                B.r1 = B.r2 = B.r3 = 1.0;
            }
        }
        const double2 x01 = qp_ld2(pb + lane * 4), x23 = qp_ld2(pb + lane * 4 + 2);
        double x0 = x01.x, x1 = x01.y, x2 = x23.x, x3 = x23.y;
        qt3_solve4(B, x0, x1, x2, x3);
        if (p == 7) { x1 = x2 = x3 = 0.0; }
        if (lane >= G) { x0 = x1 = x2 = x3 = 0.0; }
        {
            const double2 y01 = qp_ld2(y + c0), y23 = qp_ld2(y + c0 + 2);
            acc += x0 * y01.x + (p < 7 ? x1 * y01.y + x2 * y23.x + x3 * y23.y : 0.0);
        }
        qp_st2(lp + lane * 4, x0, x1); qp_st2(lp + lane * 4 + 2, x2, x3);
        __syncwarp();
        double LF[4], FF[4];
        const int Jlo = half ? Jp + 1 : Jp;
#pragma unroll
        for (int I = 0; I < 4; ++I) {
            LF[I] = lp[(8 * I + g) * 4 + t];
            const int r = 8 * I + g - c0;
            FF[I] = (p < 7 && I >= Jlo && r >= 0 && 8 * I + g < G) ? f[Qt3::fb(p) + (r < 0 ? 0 : r) * 4 + t] : 0.0;
        }
#pragma unroll
        for (int I = 0; I < 4; ++I) {
#pragma unroll
            for (int J = 0; J < 4; ++J)
                if (p < 7 && J >= Jlo) qp_dmma(E[4 * I + J][0], E[4 * I + J][1], -LF[I], FF[J]);
#pragma unroll
            for (int J = 0; J <= I; ++J) qp_dmma(P[Qt3::tile(I, J)][0], P[Qt3::tile(I, J)][1], LF[I], LF[J]);
        }
        __syncwarp();
    }
    return -acc;
}

// Solves L z = x, then L^T nu = y - z, with the factor image `f` (panels of L, 1 / L_ii at fRI, y at fY): returns nu_lane.
__device__ __forceinline__ double qt3_outward_solve(const double* __restrict__ f, double x, int lane) {
    constexpr int G = Qt3::G;
    constexpr unsigned FULL = 0xffffffffu;
    const int ln = lane < G ? lane : G - 1;
    const double r_own = f[Qt3::fRI + ln];
    double rk = x;
#pragma unroll
    for (int i = 0; i < G - 1; ++i) {  // forward: lane > i needs L[lane][i]
        const double zi = __shfl_sync(FULL, rk, i) * f[Qt3::fRI + i];
        const int p = i >> 2;
        const double lv = f[Qt3::fb(p) + (ln - 4 * p < 0 ? 0 : ln - 4 * p) * 4 + (i & 3)];
        if (lane > i) rk -= lv * zi;
    }
    double R = (lane < G ? f[Qt3::fY + ln] : 0.0) - rk * r_own;
    const int pl = ln >> 2;
    const double* own_col = f + Qt3::fb(pl) + (ln & 3) - 4 * pl * 4;  // L[i][ln] = own_col[i * 4] for i >= 4 pl
#pragma unroll
    for (int i = G - 1; i > 0; --i) {  // backward with L^T: lane < i needs L[i][lane]
        const double ni = __shfl_sync(FULL, R, i) * f[Qt3::fRI + i];
        const double lv = own_col[(i < 4 * pl ? 4 * pl : i) * 4];
        if (lane < i) R -= lv * ni;
    }
    return lane < G ? R * r_own : 0.0;
}

}  // namespace ub
