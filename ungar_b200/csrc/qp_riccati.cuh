// Batched equality-constrained QP solve for the problems whose only equalities are the initial condition and the dynamics
// defects (quadrotor, RC car): a Riccati recursion over the KKT block records.
//
// Replaces what SoftSQPOptimizer::SolveLocalQPProblem hands to OSQP (include/ungar/optimization/soft_sqp.hpp:193-233):
//     min_d 1/2 d^T P d + q^T d   s.t.  A d = -g
// In stage form (record blocks of ungar_b200_kkt_layout; z_k = [dx_k; du_k]):
//     cost   sum_k 1/2 z_k^T H_k z_k + q_k^T z_k  +  sum_k du_k^T diag(Hc_k) du_{k+1}  +  1/2 dx_N^T HN dx_N + q_N^T dx_N
//     s.t.   dx_0 = -g_0,     dx_{k+1} = -A_k z_k - g_{k+1}
// The input-rate term of the objectives (quadrotor.example.cpp:219-225, rc_car.example.cpp:211-217) couples du_k and du_{k+1};
// it is carried by augmenting the state with the previous input, s_k = [dx_k; v_k], v_k = du_{k-1}.  The value function
// V_k(s) = 1/2 s^T [Pxx Pxv; Pxv^T Pvv] s + [px; pv]^T s is propagated backwards with the augmented blocks kept separate (the
// transition of the v part is the identity on du_k and has no state dependence, so only A_k^T Pxx A_k is a real product):
//     G    = A^T Pxx A,   T = A^T Pxv                                   (nz x nz, nz x nu)
//     M    = H_k + G  - [0 T] - [0 T]^T + [0 0; 0 Pvv]                  the z-z block after substituting the dynamics
//     m    = q_k - A^T (px - Pxx g_{k+1})  + [0; pv - Pxv^T g_{k+1}]
//     Y    = M_uu^-1 [M_ux | D | m_u],   D = diag(Hc_{k-1})             (Cholesky of the nu x nu block)
//     du_k = -Y [dx_k; v_k; 1]
//     Pxx' = M_xx - M_xu Y_x,  Pxv' = -M_xu Y_v,  Pvv' = -D Y_v,  px' = m_x - M_xu y_m,  pv' = -D y_m
// then a forward rollout gives the step, and one more backward sweep the multipliers from the stationarity rows
//     lambda_N = -(HN dx_N + q_N),    lambda_k = -(H_xx dx_k + H_xu du_k + q_x,k + A_x,k^T lambda_{k+1}).
// Exact (no ADMM iterations); the reference's OSQP v0.6.3 is absent from the tree.
//
// Mapping: one warp per trajectory; every matrix of a stage lives in shared memory (quadrotor: 16 KB per warp) and the lanes
// stride over output entries.  The stage data (A_k, H_k, q_k, g_{k+1}, Hc_{k-1}; in the rollout also the gains Y_k) is fetched with
// cp.async ONE STAGE AHEAD into the other of two stage buffers, so the three sweeps never wait on HBM inside a stage.  The blocks
// are tiny (13 x 17, 6 x 8): the kernel runs next to sweeps that are 10-100x larger.
#pragma once

#include "sweep.cuh"

namespace ub {

template <class Mdl>
struct RiccatiShape {
    static constexpr int NX = Mdl::NX, NU = Mdl::NU, NZ = Mdl::NZ, NS = NX + NU, TRI = NZ * (NZ + 1) / 2;
    // stage buffer (one of two, filled with cp.async one stage ahead):  A_k | packed H_k | q_k | g_{k+1} | diag(Hc_{k-1}) | Y_k
    static constexpr int bA = 0, bH = bA + NX * NZ, bQ = bH + TRI, bG = bQ + NZ, bD = bG + NX, bY = bD + NU,
                         STAGE = (bY + NU * (NS + 1) + 1) & ~1;
    // per-warp shared memory (doubles)
    static constexpr int oS = 0, oM = oS + 2 * STAGE, oW = oM + NZ * NZ, oPxx = oW + NX * NZ, oPxv = oPxx + NX * NX, oPvv = oPxv + NX * NU,
                         oT = oPvv + NU * NU, oY = oT + NZ * NU, oVec = oY + NU * (NS + 1),
                         // small vectors: px (NX) pv (NU) wx (NX) wv (NU) m (NZ) g (NX) q (NZ) D (NU) s (NS) du (NU) lam (NX)
                         vPx = 0, vPv = vPx + NX, vWx = vPv + NU, vWv = vWx + NX, vM = vWv + NU, vG = vM + NZ, vQ = vG + NX, vD = vQ + NZ,
                         vS = vD + NU, vDu = vS + NS, vLam = vDu + NU, vEnd = vLam + NX,
                         total = (oVec + vEnd + 1) & ~1;
    static constexpr int WARPS = 4;
    static constexpr int SMEM_BYTES = WARPS * total * 8;
    static constexpr int WS_STAGE = NU * (NS + 1);  // Y_k per stage in the global workspace
};

template <class Mdl>
__global__ void __launch_bounds__(RiccatiShape<Mdl>::WARPS * 32)
qp_riccati_kernel(const double* __restrict__ rec_all, long long ld_rec, double* __restrict__ ws_all, double* __restrict__ step_all,
                  long long ld_step, double* __restrict__ mult_all, long long ld_mult, int N, long long batch, RecLayout L,
                  const int* __restrict__ skip_status) {
    using R = RiccatiShape<Mdl>;
    constexpr int NX = R::NX, NU = R::NU, NZ = R::NZ, NS = R::NS, TRI = R::TRI;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* const sm = reinterpret_cast<double*>(smem_raw) + wib * R::total;
    double *sS = sm + R::oS, *sM = sm + R::oM, *sW = sm + R::oW, *sPxx = sm + R::oPxx, *sPxv = sm + R::oPxv, *sPvv = sm + R::oPvv,
           *sT = sm + R::oT, *sY = sm + R::oY, *sv = sm + R::oVec;
    double *px = sv + R::vPx, *pv = sv + R::vPv, *wx = sv + R::vWx, *wv = sv + R::vWv, *mv = sv + R::vM, *ss = sv + R::vS,
           *sdu = sv + R::vDu, *slam = sv + R::vLam;
    const long long b = (long long)blockIdx.x * R::WARPS + wib;
    if (b >= batch) return;
    if (skip_status && skip_status[2 * b] != 0) return;  // SQP loop: this trajectory has stopped (warp-uniform)
    const double* __restrict__ rec = rec_all + b * ld_rec;
    double* __restrict__ ws = ws_all + b * (long long)N * R::WS_STAGE;
    const int uoff = NX * (N + 1);  // first input entry of the decision vector / gradient

    auto cp8 = [](double* dst, const double* src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    // Asynchronous fetch of stage k into buffer `buf`; `gains`: also Y_k from the workspace (forward rollout).
    auto fetch = [&](int k, int buf, bool gains) {
        double* S = sS + buf * R::STAGE;
        const double* Ak = rec + L.A + (long long)k * NX * NZ;
        const double* Hk = rec + L.H + (long long)k * TRI;
        for (int e = lane; e < NX * NZ; e += 32) cp8(S + R::bA + e, Ak + e);
        for (int e = lane; e < TRI; e += 32) cp8(S + R::bH + e, Hk + e);
        for (int e = lane; e < NZ; e += 32) cp8(S + R::bQ + e, e < NX ? rec + L.grad + NX * k + e : rec + L.grad + uoff + NU * k + (e - NX));
        for (int e = lane; e < NX; e += 32) cp8(S + R::bG + e, rec + L.g + NX * (k + 1) + e);
        for (int e = lane; e < NU; e += 32) {
            if (k > 0) cp8(S + R::bD + e, rec + L.Hc + NU * (k - 1) + e);
            else S[R::bD + e] = 0.0;
        }
        if (gains)
            for (int e = lane; e < R::WS_STAGE; e += 32) cp8(S + R::bY + e, ws + (long long)k * R::WS_STAGE + e);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto arrive = [] {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    };

    fetch(N - 1, 0, false);
    // ---- terminal value function: Pxx = HN, px = q_N, everything else zero --------------------------------------------------------
    for (int e = lane; e < NX * NX; e += 32) {
        const int i = e / NX, j = e - i * NX;
        sPxx[e] = rec[L.HN + (i <= j ? tri_index(NX, i, j) : tri_index(NX, j, i))];
    }
    for (int e = lane; e < NX * NU; e += 32) sPxv[e] = 0.0;
    for (int e = lane; e < NU * NU; e += 32) sPvv[e] = 0.0;
    for (int e = lane; e < NX; e += 32) px[e] = rec[L.grad + NX * N + e];
    for (int e = lane; e < NU; e += 32) pv[e] = 0.0;
    __syncwarp();

    // ================================================================ backward Riccati sweep
    for (int k = N - 1; k >= 0; --k) {
        const int buf = (N - 1 - k) & 1;
        arrive();
        if (k > 0) fetch(k - 1, buf ^ 1, false);
        else fetch(0, buf ^ 1, false);  // first stage of the forward rollout (its gains Y_0 stay in sY)
        const double* S = sS + buf * R::STAGE;
        const double *sA = S + R::bA, *sHp = S + R::bH, *sq = S + R::bQ, *sg = S + R::bG, *sD = S + R::bD;
        // wx = px - Pxx g,  wv = pv - Pxv^T g
        for (int e = lane; e < NS; e += 32) {
            double acc = e < NX ? px[e] : pv[e - NX];
            if (e < NX) { for (int r = 0; r < NX; ++r) acc -= sPxx[e * NX + r] * sg[r]; wx[e] = acc; }
            else        { for (int r = 0; r < NX; ++r) acc -= sPxv[r * NU + (e - NX)] * sg[r]; wv[e - NX] = acc; }
        }
        // W = Pxx A,  T = A^T Pxv
        for (int e = lane; e < NX * NZ; e += 32) {
            const int i = e / NZ, j = e - i * NZ;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < NX; ++r) acc += sPxx[i * NX + r] * sA[r * NZ + j];
            sW[e] = acc;
        }
        for (int e = lane; e < NZ * NU; e += 32) {
            const int i = e / NU, j = e - i * NU;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < NX; ++r) acc += sA[r * NZ + i] * sPxv[r * NU + j];
            sT[e] = acc;
        }
        __syncwarp();
        // M = H + A^T W - [0 T] - [0 T]^T + [0 0; 0 Pvv] ;  m = q - A^T wx + [0; wv]
        for (int e = lane; e < NZ * NZ; e += 32) {
            const int i = e / NZ, j = e - i * NZ;
            double acc = sHp[i <= j ? tri_index(NZ, i, j) : tri_index(NZ, j, i)];
#pragma unroll
            for (int r = 0; r < NX; ++r) acc += sA[r * NZ + i] * sW[r * NZ + j];
            if (j >= NX) acc -= sT[i * NU + (j - NX)];
            if (i >= NX) acc -= sT[j * NU + (i - NX)];
            if (i >= NX && j >= NX) acc += sPvv[(i - NX) * NU + (j - NX)];
            sM[e] = acc;
        }
        for (int e = lane; e < NZ; e += 32) {
            double acc = sq[e];
#pragma unroll
            for (int r = 0; r < NX; ++r) acc -= sA[r * NZ + e] * wx[r];
            if (e >= NX) acc += wv[e - NX];
            mv[e] = acc;
        }
        __syncwarp();
        // Cholesky of M_uu, redundantly in every lane (nu <= 4); then one right-hand side per lane: Y = M_uu^-1 [M_ux | D | m_u]
        double Lc[NU][NU];
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double acc = sM[(NX + i) * NZ + NX + j];
#pragma unroll
                for (int r = 0; r < j; ++r) acc -= Lc[i][r] * Lc[j][r];
                Lc[i][j] = i == j ? sqrt(acc) : acc / Lc[j][j];
            }
        for (int c = lane; c <= NS; c += 32) {
            double y[NU];
#pragma unroll
            for (int a = 0; a < NU; ++a) y[a] = c < NX ? sM[(NX + a) * NZ + c] : c < NS ? (a == c - NX ? sD[a] : 0.0) : mv[NX + a];
#pragma unroll
            for (int a = 0; a < NU; ++a) {
#pragma unroll
                for (int r = 0; r < a; ++r) y[a] -= Lc[a][r] * y[r];
                y[a] /= Lc[a][a];
            }
#pragma unroll
            for (int a = NU - 1; a >= 0; --a) {
#pragma unroll
                for (int r = a + 1; r < NU; ++r) y[a] -= Lc[r][a] * y[r];
                y[a] /= Lc[a][a];
            }
#pragma unroll
            for (int a = 0; a < NU; ++a) sY[a * (NS + 1) + c] = y[a];
        }
        __syncwarp();
        if (k > 0)
            for (int e = lane; e < R::WS_STAGE; e += 32) ws[(long long)k * R::WS_STAGE + e] = sY[e];
        // value function of stage k
        for (int e = lane; e < NX * NX; e += 32) {
            const int i = e / NX, j = e - i * NX;
            double acc = sM[i * NZ + j];
#pragma unroll
            for (int a = 0; a < NU; ++a) acc -= sM[i * NZ + NX + a] * sY[a * (NS + 1) + j];
            sPxx[e] = acc;
        }
        for (int e = lane; e < NX * NU; e += 32) {
            const int i = e / NU, j = e - i * NU;
            double acc = 0.0;
#pragma unroll
            for (int a = 0; a < NU; ++a) acc -= sM[i * NZ + NX + a] * sY[a * (NS + 1) + NX + j];
            sPxv[e] = acc;
        }
        for (int e = lane; e < NU * NU; e += 32) {
            const int i = e / NU, j = e - i * NU;
            sPvv[e] = -sD[i] * sY[i * (NS + 1) + NX + j];
        }
        for (int e = lane; e < NS; e += 32) {
            if (e < NX) {
                double acc = mv[e];
#pragma unroll
                for (int a = 0; a < NU; ++a) acc -= sM[e * NZ + NX + a] * sY[a * (NS + 1) + NS];
                px[e] = acc;
            } else pv[e - NX] = -sD[e - NX] * sY[(e - NX) * (NS + 1) + NS];
        }
        __syncwarp();
    }

    // ================================================================ forward rollout: s_0 = [-g_0; 0]
    // Stage 0 was fetched (buffer N & 1) while the backward sweep finished; its gains Y_0 are still in sY (never written to the
    // workspace), every later stage's gains come back from the workspace one stage ahead.
    double* __restrict__ step = step_all + b * ld_step;
    for (int e = lane; e < NS; e += 32) ss[e] = e < NX ? -rec[L.g + e] : 0.0;
    __syncwarp();
    for (int k = 0; k < N; ++k) {
        const int buf = (N + k) & 1;
        arrive();
        if (k + 1 < N) fetch(k + 1, buf ^ 1, true);
        const double* S = sS + buf * R::STAGE;
        const double *sA = S + R::bA, *sg = S + R::bG, *Yk = k == 0 ? sY : S + R::bY;
        for (int a = lane; a < NU; a += 32) {
            double acc = -Yk[a * (NS + 1) + NS];
#pragma unroll
            for (int c = 0; c < NS; ++c) acc -= Yk[a * (NS + 1) + c] * ss[c];
            sdu[a] = acc;
        }
        __syncwarp();
        for (int e = lane; e < NZ; e += 32) {
            if (e < NX) step[NX * k + e] = ss[e];
            else step[uoff + NU * k + (e - NX)] = sdu[e - NX];
        }
        double nxt = 0.0;
        if (lane < NX) {
            nxt = -sg[lane];
#pragma unroll
            for (int c = 0; c < NX; ++c) nxt -= sA[lane * NZ + c] * ss[c];
#pragma unroll
            for (int c = 0; c < NU; ++c) nxt -= sA[lane * NZ + NX + c] * sdu[c];
        }
        __syncwarp();
        if (lane < NX) ss[lane] = nxt;
        else if (lane < NS) ss[lane] = sdu[lane - NX];
        __syncwarp();
    }
    if (lane < NX) step[NX * N + lane] = ss[lane];
    if (!mult_all) return;

    // ================================================================ multipliers, reference row order [x_0 - x_m | defects]
    double* __restrict__ mult = mult_all + b * ld_mult;
    if (lane < NX) {
        double acc = rec[L.grad + NX * N + lane];
        for (int c = 0; c < NX; ++c) {
            const int i = lane <= c ? lane : c, j = lane <= c ? c : lane;
            acc += rec[L.HN + tri_index(NX, i, j)] * ss[c];
        }
        slam[lane] = -acc;
        mult[NX * N + lane] = -acc;
    }
    __syncwarp();
    for (int k = N - 1; k >= 0; --k) {
        const double* Ak = rec + L.A + (long long)k * NX * NZ;
        const double* Hk = rec + L.H + (long long)k * TRI;
        double lam = 0.0;
        if (lane < NX) {
            double acc = rec[L.grad + NX * k + lane];
            for (int c = 0; c < NZ; ++c) {
                const int i = lane <= c ? lane : c, j = lane <= c ? c : lane;
                const double d = c < NX ? step[NX * k + c] : step[uoff + NU * k + (c - NX)];
                acc += Hk[tri_index(NZ, i, j)] * d;
            }
            for (int r = 0; r < NX; ++r) acc += Ak[r * NZ + lane] * slam[r];
            lam = -acc;
        }
        __syncwarp();
        if (lane < NX) { slam[lane] = lam; mult[NX * k + lane] = lam; }
        __syncwarp();
    }
}

}  // namespace ub
