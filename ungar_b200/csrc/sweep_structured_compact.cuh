// KKT stage sweep of the quadruped NMPC writing the COMPACT record (compact.cuh): the arithmetic of sweep_structured.cuh (phase 0 /
// phase A / phase B are verbatim the same code, so every value is bit-identical to the dense sweep's), but a node's
// structurally non-zero slots — A 250, H 61, C 224 — and its g / h / q entries all land in ONE 626-double chunk image per warp,
// which leaves with ONE cp.async.bulk per node.  Per node 5008 B instead of 12 656 B (dense A + H + C + g + h + q), and no
// scattered small stores.  Replaces, like the dense sweep, one pass of SoftSQPOptimizer::AssembleOSQPInstance (soft_sqp.hpp:141-158).
#pragma once

#include "compact.cuh"
#include "sweep_structured.cuh"

namespace ub {

struct QuadrupedCompactSweep {
    static constexpr int NX = 13, NU = 24, NP = 29, RUN = 10, CORE = 77, NRHO = 49;
    static constexpr int STAGE = 2 * Compact::NODE;  // one chunk image per warp
    static constexpr int oCORE = STAGE, oXS = oCORE + 848, oUS = oXS + (RUN + 2) * NX, oPS = oUS + (RUN + 1) * NU,
                         oRHO = oPS + (RUN + 1) * NP + 1, PER_WARP = oRHO + NRHO + 1;
    static_assert((PER_WARP * 8) % 16 == 0 && (RUN + 1) * CORE <= 848, "shared memory layout");
    static constexpr int WARPS = 2;
    static constexpr int SMEM_BYTES = PER_WARP * 8;
    // resident teams per SM: the compact image is 5 KB per warp (23 KB per team), so registers set the limit — 128 per thread
    // (no spills) leave room for 8 teams = 16 warps (9 teams at 96 registers measured slower: 0.196 vs 0.187 ms); the dense sweep's 24 KB pair image caps it at 6
    static constexpr int TEAMS_PER_SM = 8;
    static constexpr int cR = 0, cQ = 9, cE = 21, cXN = 25, cSGN = 38, cM = 39;
};

// One CTA = one TEAM of two warps sharing the run's inputs; warp w owns node 2 p + w of every node pair and its OWN staging image: one
// compact chunk (compact.cuh), handed to the TMA engine with ONE bulk store per node (5008 B).
template <bool BARRIER>
__global__ void __launch_bounds__(64, QuadrupedCompactSweep::TEAMS_PER_SM)
quadruped_compact_kernel(const double* __restrict__ xp_all, long long ld_xp, double* __restrict__ rec_all, long long ld_rec,
                            double* __restrict__ partials, int N, int run_len, int runs_per_traj, long long total_runs,
                            BarrierCoef<double> bar, unsigned int* __restrict__ sched, const int* __restrict__ active = nullptr,
                            const unsigned int* __restrict__ n_active = nullptr) {
    using Q = QuadrupedCompactSweep;
    using K = Compact;
    using Mdl = Quadruped;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double* const wsm   = reinterpret_cast<double*>(smem_raw);
    double* const cw    = wsm + w * K::NODE;  // this warp's chunk image
    double* const cores = wsm + Q::oCORE;
    double* const xs    = wsm + Q::oXS;
    double* const us    = wsm + Q::oUS;
    double* const ps    = wsm + Q::oPS;
    double* const rs    = wsm + Q::oRHO;
    for (int e = tid; e < Q::STAGE; e += 64) wsm[e] = 0.0;  // pads and the slots no node writes stay zero
    __syncthreads();

    // ---- per-lane roles and staging addresses (the slot is fixed per warp, so all of these are loop constants) ----------
    const int leg = lane >> 3, c = lane & 7;  // column lanes: c < 6 -> (f0 f1 f2 r0 r1 r2) of `leg`
    const bool col_lane = c < 6, fcol = c < 3;
    const int c3 = fcol ? c : (c < 6 ? c - 3 : 0);
    const int cq = c & 3;
    const bool qcol = lane < 4, wcol = lane >= 8 && lane < 11, pvlane = lane >= 16 && lane < 19;
    double* const aqIn = cw + K::oAQ + 8 + 6 * leg + c;        // input column of this lane in the q+ / w+ rows (valid if col_lane)
    double* const aqSt = cw + K::oAQ + (qcol ? c : 4 + c);     // state column of this lane (valid if qcol || wcol)
    double* const ap   = cw + K::oAP;                          // p+ / v+ rows
    double* const hIn  = cw + K::oHb + 6 * (2 * leg + (c >= 3 ? 1 : 0)) + (c3 == 0 ? 0 : (c3 == 1 ? 3 : 5));  // row c3 of the 3x3 block
    double* const hSt  = cw + K::oHd + lane;                   // valid if lane < 13
    double* const csl  = cw + K::oCs + leg * 32;               // Cs[leg][4][8]
    double* const cpl  = cw + K::oCp + leg * 24;               // Cp[leg][3][8]
    const double ec0 = c3 == 0 ? 1.0 : 0.0, ec1 = c3 == 1 ? 1.0 : 0.0, ec2 = c3 == 2 ? 1.0 : 0.0;
    bool pending = false;  // a bulk store may still be reading the staging image (CTA-uniform)

    // Runs are CLAIMED, not statically strided: the first one is the CTA's index, every further one comes from an atomic counter
    // (sched[0]).  A CTA that becomes resident late — another kernel (the NCCL all-gather of the previous step, a neighbour's H2D
    // chunk sweep) holds part of an SM — then simply claims fewer runs instead of stretching the launch by a second wave.
    __shared__ long long s_next;
    // `active` / `n_active` (SQP loop): only the trajectories still RUNNING are swept, listed by build_active_kernel
    if (n_active) total_runs = (long long)*n_active * runs_per_traj;
    long long run = blockIdx.x;
    while (run < total_runs) {
        const long long item = run / runs_per_traj;
        const long long b = active ? (long long)active[item] : item;
        const int run_in_traj = int(run - item * runs_per_traj);
        const int k0    = run_in_traj * run_len;
        const int nodes = min(N, k0 + run_len) - k0;
        const double* __restrict__ x = xp_all + b * ld_xp;
        double* __restrict__ r       = rec_all + b * ld_rec;
        double* __restrict__ rT      = r + K::tail(N);
        const double* __restrict__ Rho = x + Mdl::rho_off(N);

        // ---- phase 0: the run's inputs -> shared memory (slot j of xs/us/ps = node k0 - 1 + j) ------------------------
        {
            const int halo = k0 > 0 ? 0 : 1;  // no node -1
            const double* gx = x + Mdl::x_off(N, k0 - 1 + halo);
            for (int e = tid + halo * Q::NX; e < (nodes + 2) * Q::NX; e += 64) async_copy8(xs + e, gx + (e - halo * Q::NX));
            const double* gu = x + Mdl::u_off(N, k0 - 1 + halo);
            for (int e = tid + halo * Q::NU; e < (nodes + 1) * Q::NU; e += 64) async_copy8(us + e, gu + (e - halo * Q::NU));
            const double* gp = x + Mdl::p_off(N, k0 - 1 + halo);
            for (int e = tid + halo * Q::NP; e < (nodes + 1) * Q::NP; e += 64) async_copy8(ps + e, gp + (e - halo * Q::NP));
            if (tid < Q::NRHO) async_copy8(rs + tid, Rho + tid);  // the shared parameters travel with the same group: one exposed latency
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double dt = rs[0], mass = rs[1], I0 = rs[2], I1 = rs[3], I2 = rs[4], Llen = rs[17], g0 = rs[18], mu = rs[19];
        const double iI0 = __drcp_rn(I0), iI1 = __drcp_rn(I1), iI2 = __drcp_rn(I2), inv_m = __drcp_rn(mass);  // = 1.0 / x, correctly rounded
        const double hip0 = rs[5 + 3 * leg], hip1 = rs[6 + 3 * leg], hip2 = rs[7 + 3 * leg];
        const double sdt0 = dt * iI0, sdt1 = dt * iI1, sdt2 = dt * iI2;

        // ---- phase A: thread-per-node primal cores (thread 0 = halo node k0 - 1: only its rotation matrix) -------------
        if (tid >= 1 && tid <= nodes) {
            node_core(xs + tid * Q::NX, us + tid * Q::NU, ps + tid * Q::NP, dt, inv_m, g0, I0, I1, I2, iI0, iI1, iI2,
                      cores + tid * Q::CORE);
        } else if (tid == 0 && k0 > 0) {
            rot_matrix(xs[3], xs[4], xs[5], xs[6], cores + Q::cR);
            drot_matrices(xs[3], xs[4], xs[5], xs[6], cores + Q::cM);
        }
        __syncthreads();

        double cost_acc = 0.0, bar_acc = 0.0, gmax = 0.0, hmax = -INFINITY;  // per-lane partials over this warp's nodes
        if (k0 == 0 && w == 0 && lane < 13) {  // x_0 - x_measured (:266-268)
            const double gv = x[lane] - x[Mdl::xm_off(N) + lane];
            rT[K::tG0 + lane] = gv;
            if (lane == 0) rT[K::tG0 + 13] = 0.0;  // pad
            gmax = fabs(gv);
        }

        // ---- phase B: warp-per-node block fill; warp w owns node 2 p + w of pair p -----------------------------------------
        const int pairs = (nodes + 1) >> 1;
        for (int p = 0; p < pairs; ++p) {
            const int n = 2 * p + w, k = k0 + n;
            const bool active = n < nodes;
            const double* __restrict__ xk = xs + (n + 1) * Q::NX;
            const double* __restrict__ uk = us + (n + 1) * Q::NU + 6 * leg;
            const double* __restrict__ pk = ps + (n + 1) * Q::NP;
            const double* __restrict__ co = cores + (n + 1) * Q::CORE;
            const double* __restrict__ R  = co + Q::cR;
            const double* __restrict__ Qw = co + Q::cQ;

            // -------- values that do not touch the staging image: computed while the previous bulk store drains ------------
            double W0 = 0, W1 = 0, W2 = 0, z0 = 0, z1 = 0, z2 = 0, s = 0, f0 = 0, f1 = 0, f2 = 0, r0 = 0, r1 = 0, r2 = 0;
            double Rc0 = 0, Rc1 = 0, Rc2 = 0, M0 = 0, M1 = 0, M2 = 0, M3 = 0, M4 = 0, M5 = 0, M6 = 0, M7 = 0, M8 = 0;
            if (active) {
                {  // this lane's matrix d(R v)/dq_cq of the node (used for f here and for r in the contact rows)
                    const double* __restrict__ Mc = co + Q::cM + 9 * cq;
                    M0 = Mc[0]; M1 = Mc[1]; M2 = Mc[2]; M3 = Mc[3]; M4 = Mc[4]; M5 = Mc[5]; M6 = Mc[6]; M7 = Mc[7]; M8 = Mc[8];
                }
                f0 = uk[0]; f1 = uk[1]; f2 = uk[2]; r0 = uk[3]; r1 = uk[4]; r2 = uk[5];
                s = pk[13 + 4 * leg];
                Rc0 = R[c3]; Rc1 = R[3 + c3]; Rc2 = R[6 + c3];  // column c3 of R
                // this lane's column of W = d w+ / d z:  f column: dt I^-1 s (r x R[:, c]);  r column: dt I^-1 s (e_c x R f)
                const double Rf0 = R[0] * f0 + R[1] * f1 + R[2] * f2, Rf1 = R[3] * f0 + R[4] * f1 + R[5] * f2,
                             Rf2 = R[6] * f0 + R[7] * f1 + R[8] * f2;
                const double u0 = fcol ? r0 : ec0, u1 = fcol ? r1 : ec1, u2 = fcol ? r2 : ec2;
                const double v0 = fcol ? Rc0 : Rf0, v1 = fcol ? Rc1 : Rf1, v2 = fcol ? Rc2 : Rf2;
                W0 = s * sdt0 * (u1 * v2 - u2 * v1); W1 = s * sdt1 * (u2 * v0 - u0 * v2); W2 = s * sdt2 * (u0 * v1 - u1 * v0);
                // q columns: dt I^-1 sum_i s_i r_i x d(R f_i)/dq_c — every leg adds its part, xor-reduced over the 4 legs
                const double d0 = M0 * f0 + M1 * f1 + M2 * f2, d1 = M3 * f0 + M4 * f1 + M5 * f2, d2 = M6 * f0 + M7 * f1 + M8 * f2;
                z0 = s * (r1 * d2 - r2 * d1); z1 = s * (r2 * d0 - r0 * d2); z2 = s * (r0 * d1 - r1 * d0);
            }
            z0 += __shfl_xor_sync(0xffffffffu, z0, 8);  z1 += __shfl_xor_sync(0xffffffffu, z1, 8);  z2 += __shfl_xor_sync(0xffffffffu, z2, 8);
            z0 += __shfl_xor_sync(0xffffffffu, z0, 16); z1 += __shfl_xor_sync(0xffffffffu, z1, 16); z2 += __shfl_xor_sync(0xffffffffu, z2, 16);

            if (pending) {  // this warp's previous bulk store must have finished reading its image
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                pending = false;
            }
            __syncwarp();

            if (active) {
                // ---- A = d(x_{k+1} - f)/dz: rows w+ (10-12) = -W, rows q+ (3-6) = -(Qw W [+ Rmat(e)]), rows p+/v+ constants --
                if (col_lane) {
                    aqIn[4 * 32] = -W0; aqIn[5 * 32] = -W1; aqIn[6 * 32] = -W2;
                    aqIn[0 * 32] = -(Qw[0] * W0 + Qw[1] * W1 + Qw[2] * W2);
                    aqIn[1 * 32] = -(Qw[3] * W0 + Qw[4] * W1 + Qw[5] * W2);
                    aqIn[2 * 32] = -(Qw[6] * W0 + Qw[7] * W1 + Qw[8] * W2);
                    aqIn[3 * 32] = -(Qw[9] * W0 + Qw[10] * W1 + Qw[11] * W2);
                    if (fcol) {
                        const double sm = s * inv_m * dt;
                        ap[c * 6 + 2 + leg]       = -dt * sm;
                        ap[(3 + c) * 6 + 1 + leg] = -sm;
                    }
                }
                if (qcol || wcol) {  // state columns: lanes 0..3 -> q_c, lanes 8..10 -> w_c
                    double G0, G1, G2, add0 = 0, add1 = 0, add2 = 0, add3 = 0;
                    if (qcol) {
                        const double e0 = co[Q::cE], e1 = co[Q::cE + 1], e2 = co[Q::cE + 2], e3 = co[Q::cE + 3];
                        G0 = sdt0 * z0; G1 = sdt1 * z1; G2 = sdt2 * z2;
                        add0 = c == 0 ? e3 : c == 1 ? e2 : c == 2 ? -e1 : e0;   // Rmat(e) column c: d(q (x) e)/dq_c
                        add1 = c == 0 ? -e2 : c == 1 ? e3 : c == 2 ? e0 : e1;
                        add2 = c == 0 ? e1 : c == 1 ? -e0 : c == 2 ? e3 : e2;
                        add3 = c == 0 ? -e0 : c == 1 ? -e1 : c == 2 ? -e2 : e3;
                    } else {  // e_c + dt I^-1 (Iw x e_c - I_c (w x e_c))
                        const double w0 = xk[10], w1 = xk[11], w2 = xk[12];
                        const double Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I2 * w2;
                        const double Ic = pick3(c, I0, I1, I2);
                        G0 = ec0 + sdt0 * ((Iw1 * ec2 - Iw2 * ec1) - Ic * (w1 * ec2 - w2 * ec1));
                        G1 = ec1 + sdt1 * ((Iw2 * ec0 - Iw0 * ec2) - Ic * (w2 * ec0 - w0 * ec2));
                        G2 = ec2 + sdt2 * ((Iw0 * ec1 - Iw1 * ec0) - Ic * (w0 * ec1 - w1 * ec0));
                    }
                    aqSt[4 * 32] = -G0; aqSt[5 * 32] = -G1; aqSt[6 * 32] = -G2;
                    aqSt[0 * 32] = -(add0 + Qw[0] * G0 + Qw[1] * G1 + Qw[2] * G2);
                    aqSt[1 * 32] = -(add1 + Qw[3] * G0 + Qw[4] * G1 + Qw[5] * G2);
                    aqSt[2 * 32] = -(add2 + Qw[6] * G0 + Qw[7] * G1 + Qw[8] * G2);
                    aqSt[3 * 32] = -(add3 + Qw[9] * G0 + Qw[10] * G1 + Qw[11] * G2);
                }
                if (pvlane) {  // lanes 16..18: the constant p / v entries
                    ap[c * 6] = -1.0; ap[c * 6 + 1] = -dt; ap[(3 + c) * 6] = -1.0;
                }

                // ---- state part of the objective and the defects: lanes 0..12 own state entry `lane` -----------------------
                if (lane < 13) {
                    const double sgn = co[Q::cSGN];
                    const double wgt = lane < 2 ? 0.1 : (lane == 2 ? 10.0 : 1.0);  // Vector3r{0.1, 0.1, 10} (:228)
                    const bool isq   = lane >= 3 && lane < 7;
                    const double res = wgt * (isq ? xk[lane] + sgn * pk[lane] : xk[lane] - pk[lane]);
                    cost_acc += res * res;
                    cw[K::oQ + lane] = 2.0 * wgt * res;
                    *hSt = 2.0 * wgt * wgt + (BARRIER ? 1e-6 : 0.0);
                    const double gv = xk[13 + lane] - co[Q::cXN + lane];  // x_{k+1} - f(x_k, u_k)   (:276)
                    cw[K::oG + lane] = gv;
                    gmax = fmax(gmax, fabs(gv));
                }

                // ---- this lane's input entry: objective, inequalities of its leg, barrier, Gauss-Newton rows -----------------
                if (col_lane) {
                    if (fcol) {  // f_c:  h_A = -s f_z,  h_B = s |f_xy|_eps - mu f_z   (:330-331)
                        const double f2xy = f0 * f0 + f1 * f1 + UB_EPS;
                        const double inv_fxy = rsqrt(f2xy), fxy = f2xy * inv_fxy;
                        const double hA = -s * f2, hB = s * fxy - mu * f2;
                        double bA = 0, dA = 0, ddA = 0, bB = 0, dB = 0, ddB = 0;
                        if (BARRIER) {
                            barrier_eval_sel(bar, hA, bA, dA, ddA);
                            barrier_eval_sel(bar, hB, bB, dB, ddB);
                        }
                        const double gB0 = s * f0 * inv_fxy, gB1 = s * f1 * inv_fxy, gB2 = -mu;  // grad h_B wrt f
                        const double gA_c = c == 2 ? -s : 0.0, gB_c = pick3(c, gB0, gB1, gB2), fc = uk[c];
                        cost_acc += 1e-8 * fc * fc;
                        cw[K::oQ + 13 + 6 * leg + c] = 2e-8 * fc + dA * gA_c + dB * gB_c;
                        if (c < 2) {
                            const double hv = c == 0 ? hA : hB;
                            cw[K::oHi + 3 * leg + c] = hv;
                            bar_acc += c == 0 ? bA : bB;
                            hmax = fmax(hmax, hv);
                        }
                        hIn[0] = ddA * gA_c * gA_c + ddB * gB_c * gB_c + 2e-8 + (BARRIER ? 1e-6 : 0.0);
                        if (c < 2) hIn[1] = ddA * gA_c * (c == 1 ? -s : 0.0) + ddB * gB_c * (c == 0 ? gB1 : gB2);
                        if (c < 1) hIn[2] = ddA * gA_c * (-s) + ddB * gB_c * gB2;
                    } else {  // r_c':  h_C = s |r - hip|_eps - L   (:332-333)
                        const double dr0 = r0 - hip0, dr1 = r1 - hip1, dr2 = r2 - hip2;
                        const double n2 = dr0 * dr0 + dr1 * dr1 + dr2 * dr2 + UB_EPS;
                        const double inv_nr = rsqrt(n2), nr = n2 * inv_nr;
                        const double hC = s * nr - Llen;
                        double bC = 0, dC = 0, ddC = 0;
                        if (BARRIER) barrier_eval_sel(bar, hC, bC, dC, ddC);
                        const double gC0 = s * dr0 * inv_nr, gC1 = s * dr1 * inv_nr, gC2 = s * dr2 * inv_nr;
                        const double gC_c = pick3(c3, gC0, gC1, gC2);
                        const double rc = uk[c] - pk[14 + 4 * leg + c3];
                        cost_acc += rc * rc;
                        cw[K::oQ + 13 + 6 * leg + c] = 2.0 * rc + dC * gC_c;
                        if (c == 3) {
                            cw[K::oHi + 3 * leg + 2] = hC;
                            bar_acc += bC;
                            hmax = fmax(hmax, hC);
                        }
                        hIn[0] = ddC * gC_c * gC_c + 2.0 + (BARRIER ? 1e-6 : 0.0);
                        if (c3 < 2) hIn[1] = ddC * gC_c * (c3 == 0 ? gC1 : gC2);
                        if (c3 < 1) hIn[2] = ddC * gC_c * gC2;
                    }
                }

                // ---- contact rows of this lane's leg (:279-303); previous-node kinematics recomputed from shared memory ---------
                {
                    const double ft0 = xk[0] + R[0] * r0 + R[1] * r1 + R[2] * r2;
                    const double ft1 = xk[1] + R[3] * r0 + R[4] * r1 + R[5] * r2;
                    const double ft2 = xk[2] + R[6] * r0 + R[7] * r1 + R[8] * r2;
                    double Rp0 = 0, Rp1 = 0, Rp2 = 0, Dp0 = 0, Dp1 = 0, Dp2 = 0, fp0, fp1, fp2, s_prev;
                    if (k > 0) {
                        const double* xq = xk - Q::NX;
                        const double* Rh = co - Q::CORE + Q::cR;
                        const double q0 = uk[3 - Q::NU], q1 = uk[4 - Q::NU], q2 = uk[5 - Q::NU];  // r_{k-1, leg}
                        Rp0 = Rh[c3]; Rp1 = Rh[3 + c3]; Rp2 = Rh[6 + c3];
                        const double* __restrict__ Mp = co - Q::CORE + Q::cM + 9 * cq;  // previous node's d(R v)/dq_cq
                        Dp0 = Mp[0] * q0 + Mp[1] * q1 + Mp[2] * q2;
                        Dp1 = Mp[3] * q0 + Mp[4] * q1 + Mp[5] * q2;
                        Dp2 = Mp[6] * q0 + Mp[7] * q1 + Mp[8] * q2;
                        fp0 = xq[0] + Rh[0] * q0 + Rh[1] * q1 + Rh[2] * q2;
                        fp1 = xq[1] + Rh[3] * q0 + Rh[4] * q1 + Rh[5] * q2;
                        fp2 = xq[2] + Rh[6] * q0 + Rh[7] * q1 + Rh[8] * q2;
                        s_prev = pk[13 + 4 * leg - Q::NP];
                    } else {
                        fp0 = rs[34 + 4 * leg]; fp1 = rs[35 + 4 * leg]; fp2 = rs[36 + 4 * leg];  // measured foot (:296)
                        s_prev = rs[33 + 4 * leg];
                    }
                    const double c0 = (1.0 - s_prev) * s, ss = s_prev * s;
                    const double D0 = M0 * r0 + M1 * r1 + M2 * r2, D1 = M3 * r0 + M4 * r1 + M5 * r2, D2 = M6 * r0 + M7 * r1 + M8 * r2;
                    if (c < 4) {  // d foot / d q_c, and the contact values (row c)
                        csl[1 + c] = c0 * D2;
                        csl[8 + 1 + c] = ss * D0; csl[16 + 1 + c] = ss * D1; csl[24 + 1 + c] = ss * D2;
                        cpl[1 + c] = -ss * Dp0; cpl[8 + 1 + c] = -ss * Dp1; cpl[16 + 1 + c] = -ss * Dp2;
                        const double val = c == 0 ? c0 * ft2 : ss * (c == 1 ? ft0 - fp0 : c == 2 ? ft1 - fp1 : ft2 - fp2);
                        cw[K::oG + 13 + 4 * leg + c] = val;
                        gmax = fmax(gmax, fabs(val));
                    }
                    if (c < 3) {  // d foot / d r_c = R[:, c];  d foot / d p = I
                        csl[5 + c] = c0 * Rc2;
                        csl[8 + 5 + c] = ss * Rc0; csl[16 + 5 + c] = ss * Rc1; csl[24 + 5 + c] = ss * Rc2;
                        cpl[5 + c] = -ss * Rp0; cpl[8 + 5 + c] = -ss * Rp1; cpl[16 + 5 + c] = -ss * Rp2;
                        csl[(c + 1) * 8] = ss;
                        cpl[c * 8]       = k > 0 ? -ss : 0.0;
                        if (c == 2) csl[0] = c0;
                    }
                }
            }

            // ---- node complete: this warp hands its chunk image to the TMA engine (one bulk store) -------------------------------------------
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (active) {
                if (lane == 0) {
                    bulk_store(r + (long long)k * K::NODE, cw, K::NODE * 8);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                pending = true;
            }
        }

        // ---- terminal state x_N: objective gradient and diagonal block (warp 0 of the trajectory's last run) ------------------------
        if (k0 + nodes == N && w == 0) {
            const double* __restrict__ xN = xs + (nodes + 1) * Q::NX;
            const double* __restrict__ pN = x + Mdl::p_off(N, N);
            double dm = 0.0, dp = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a = xN[3 + i] - pN[3 + i], bq = xN[3 + i] + pN[3 + i];
                dm += a * a; dp += bq * bq;
            }
            const double sgn = dm > dp ? 1.0 : -1.0;
            double hdiag = 0.0;
            if (lane < 13) {
                const double wgt = lane < 2 ? 0.1 : (lane == 2 ? 10.0 : 1.0);
                const bool isq   = lane >= 3 && lane < 7;
                const double res = wgt * (isq ? xN[lane] + sgn * pN[lane] : xN[lane] - pN[lane]);
                cost_acc += res * res;
                hdiag = 2.0 * wgt * wgt + (BARRIER ? 1e-6 : 0.0);
                rT[K::tQN + lane] = 2.0 * wgt * res;
                rT[K::tHN + lane] = hdiag;
                if (lane == 0) { rT[K::tQN + 13] = 0.0; rT[K::tHN + 13] = 0.0; }  // pads
            }
        }
        // ---- per-(run, warp) partials: objective, barrier, |g|_inf, max h — one xor-tree per run instead of per node ----------------
        cost_acc = warp_sum(cost_acc);
        bar_acc  = warp_sum(bar_acc);
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
            hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
        }
        if (lane == 0) {
            double* pt = partials + ((long long)b * (2 * runs_per_traj) + 2 * run_in_traj + w) * 4;
            pt[0] = cost_acc; pt[1] = bar_acc; pt[2] = gmax; pt[3] = hmax;
        }
        if (tid == 0) s_next = (long long)gridDim.x + atomicAdd(&sched[0], 1u);
        __syncthreads();  // everyone is done with xs/us/ps/cores before the next run overwrites them
        run = s_next;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (tid == 0) {
        // the last CTA to leave re-arms the scheduler for the next launch (stream order makes the zeros visible to it)
        __threadfence();
        if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
            __threadfence();
        }
    }
}

// 32-scalar per-trajectory summary from a COMPACT record: u_0 (24), f, Zsoft, |g|_inf, max h, zero padding (one warp per trajectory).
__global__ void summary_compact_kernel(const double* __restrict__ xp_all, long long ld_xp, const double* __restrict__ rec_all, long long ld_rec,
                                       double* __restrict__ out_all, int N, int u0_off) {
    using K = Compact;
    const long long b = blockIdx.x;
    const double* rec = rec_all + b * ld_rec;
    const double* tail = rec + K::tail(N);
    const int l = threadIdx.x;
    double gmax = l < 13 ? fabs(tail[K::tG0 + l]) : 0.0, hmax = -INFINITY;
    for (int k = 0; k < N; ++k) {
        const double* ch = rec + (long long)k * K::NODE;
        if (l < 29) gmax = fmax(gmax, fabs(ch[K::oG + l]));
        if (l < 12) hmax = fmax(hmax, ch[K::oHi + l]);
    }
    for (int o = 16; o; o >>= 1) {
        gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
    }
    double v = 0.0;
    if (l < 24) v = xp_all[b * ld_xp + u0_off + l];
    else if (l == 24) v = tail[K::tCost];
    else if (l == 25) v = tail[K::tCost + 1];
    else if (l == 26) v = gmax;
    else if (l == 27) v = hmax;
    out_all[b * 32 + l] = v;
}

// List of the trajectories whose SQP status is RUNNING (order irrelevant: trajectories are independent) and their count.
__global__ void build_active_kernel(const int* __restrict__ status, long long batch, int* __restrict__ active, unsigned int* __restrict__ n_active) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < batch && status[2 * b] == 0) active[atomicAdd(n_active, 1u)] = int(b);
}

}  // namespace ub
