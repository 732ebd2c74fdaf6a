// KKT stage sweep of the quadruped single-rigid-body NMPC, structured version: ONE WARP OWNS ONE SHOOTING
// NODE at a time and walks a run of consecutive nodes of one trajectory.
//
// Why a second kernel: the generic sweep (sweep.cuh) pushes one dense tangent per thread through the whole
// stage and needs ~250 registers in fp64.  Here the chain rule is applied by hand through the model's
// bottlenecks — everything downstream of (q, w, f_i, r_i) goes through the 3-vector w+ = w + dt I^-1 tau and
// the 4x3 matrix Qw = d q+ / d w+ — so a node costs ~700 fp64 instructions per lane instead of ~50 000 per
// node, and the kernel becomes what the roofline says it should be: a stream of stores.
//
// Data flow per warp (persistent, grid = #SMs x resident CTAs):
//   inputs   flat Ungar vector [X | U | P | Rho] of the trajectory, read through L1 (4 % of the traffic)
//   staging  per-warp shared-memory image of the A, H and C blocks of TWO consecutive nodes (24 064 B).  It is
//            zeroed once; every node rewrites exactly the structurally non-zero slots, so the ~65 % zeros of the
//            dense blocks never cost an instruction again
//   stores   one elected lane issues three TMA bulk copies (cp.async.bulk.global.shared::cta, SASS UBLKCP) per
//            node pair: 7 696 + 11 248 + 5 120 B, 16-byte aligned because nodes are paired; the small vectors
//            (g, h, grad: 5 % of the bytes) go out as plain coalesced stores from registers
//
// Reference lines restated: quadruped.example.cpp:148-203 (dynamics), :209-251 (objective), :279-303 (contact
// rows), :312-338 (inequalities); soft_sqp.hpp:141-158, :245-264 (assembly); soft_inequality_constraint.hpp:133-190.
#pragma once

#include "sweep.cuh"

namespace ub {

struct QuadrupedStructured {
    static constexpr int NX = 13, NU = 24, NZ = 37, TRI = 703, NA = NX * NZ, NC = 320;
    static constexpr int PAIR_A = 2 * NA, PAIR_H = 2 * TRI, PAIR_C = 2 * NC;
    static constexpr int STAGE = PAIR_A + PAIR_H + PAIR_C;  // doubles per warp = 3008 (24 064 B)
    static constexpr int WARPS = 3;                          // per CTA; 3 CTAs / SM -> 9 warps, 216.6 KB smem
    static constexpr int SMEM_BYTES = WARPS * STAGE * 8;
};

__device__ __forceinline__ double pick3(int c, double a0, double a1, double a2) { return c == 0 ? a0 : (c == 1 ? a1 : a2); }

// d (R(q) v) / d q_c for the Eigen rotation formula v + 2 w (qv x v) + 2 qv x (qv x v); c = 0..2 -> qv_c, 3 -> w.
__device__ __forceinline__ void drot_dq(int c, double qx, double qy, double qz, double qw, double v0, double v1, double v2,
                                        double& o0, double& o1, double& o2) {
    const double e0 = c == 0 ? 1.0 : 0.0, e1 = c == 1 ? 1.0 : 0.0, e2 = c == 2 ? 1.0 : 0.0;
    const double qv_v = qx * v0 + qy * v1 + qz * v2;
    const double vc = pick3(c, v0, v1, v2), qc = pick3(c, qx, qy, qz);
    // c < 3: 2 [ e_c (qv.v) + qv v_c - 2 v qv_c ] + 2 w (e_c x v)
    const double a0 = 2.0 * (e0 * qv_v + qx * vc - 2.0 * v0 * qc) + 2.0 * qw * (e1 * v2 - e2 * v1);
    const double a1 = 2.0 * (e1 * qv_v + qy * vc - 2.0 * v1 * qc) + 2.0 * qw * (e2 * v0 - e0 * v2);
    const double a2 = 2.0 * (e2 * qv_v + qz * vc - 2.0 * v2 * qc) + 2.0 * qw * (e0 * v1 - e1 * v0);
    // c == 3: 2 (qv x v)
    const bool isw = c == 3;
    o0 = isw ? 2.0 * (qy * v2 - qz * v1) : a0;
    o1 = isw ? 2.0 * (qz * v0 - qx * v2) : a1;
    o2 = isw ? 2.0 * (qx * v1 - qy * v0) : a2;
}

struct Rot3 {
    double m00, m01, m02, m10, m11, m12, m20, m21, m22;
};
__device__ __forceinline__ Rot3 rot_matrix(double x, double y, double z, double w) {
    Rot3 R;
    R.m00 = 1.0 - 2.0 * (y * y + z * z); R.m01 = 2.0 * (x * y - w * z);       R.m02 = 2.0 * (x * z + w * y);
    R.m10 = 2.0 * (x * y + w * z);       R.m11 = 1.0 - 2.0 * (x * x + z * z); R.m12 = 2.0 * (y * z - w * x);
    R.m20 = 2.0 * (x * z - w * y);       R.m21 = 2.0 * (y * z + w * x);       R.m22 = 1.0 - 2.0 * (x * x + y * y);
    return R;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void bulk_store(void* gmem, const void* smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem),
                 "r"((unsigned)__cvta_generic_to_shared(smem)), "r"(bytes)
                 : "memory");
}

// One warp = one run of `run_len` consecutive nodes (even start, even length) of one trajectory.
template <bool BARRIER>
__global__ void __launch_bounds__(QuadrupedStructured::WARPS * 32, 3)
quadruped_structured_kernel(const double* __restrict__ xp_all, long long ld_xp, double* __restrict__ rec_all, long long ld_rec,
                            double* __restrict__ stage_cost, int N, int run_len, int runs_per_traj, long long total_runs,
                            RecLayout L, BarrierCoef<double> bar) {
    using Q = QuadrupedStructured;
    using Mdl = Quadruped;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* const stage = reinterpret_cast<double*>(smem_raw) + wib * Q::STAGE;
    double* const stA = stage;
    double* const stH = stage + Q::PAIR_A;
    double* const stC = stage + Q::PAIR_A + Q::PAIR_H;
    for (int e = lane; e < Q::STAGE; e += 32) stage[e] = 0.0;
    __syncwarp();

    const int leg = lane >> 3, c = lane & 7;  // column lanes: c < 6 -> (f0 f1 f2 r0 r1 r2) of `leg`
    const bool col_lane = c < 6;
    const int c3 = c < 3 ? c : c - 3;
    const long long warp_id = (long long)blockIdx.x * Q::WARPS + wib;
    const long long n_warps = (long long)gridDim.x * Q::WARPS;

    for (long long run = warp_id; run < total_runs; run += n_warps) {
        const long long b = run / runs_per_traj;
        const int k_begin = int(run - b * runs_per_traj) * run_len;
        const int k_end   = min(N, k_begin + run_len);
        const double* __restrict__ x = xp_all + b * ld_xp;
        double* __restrict__ r       = rec_all + b * ld_rec;
        const double* __restrict__ Rho = x + Mdl::rho_off(N);
        const double dt = Rho[0], mass = Rho[1], I0 = Rho[2], I1 = Rho[3], I2 = Rho[4], Llen = Rho[17], g0 = Rho[18],
                     mu = Rho[19];
        const double iI0 = 1.0 / I0, iI1 = 1.0 / I1, iI2 = 1.0 / I2, inv_m = 1.0 / mass;
        const double hip0 = Rho[5 + 3 * leg], hip1 = Rho[6 + 3 * leg], hip2 = Rho[7 + 3 * leg];

        if (k_begin == 0 && lane < 13) r[L.g + lane] = x[lane] - x[Mdl::xm_off(N) + lane];  // x_0 - x_measured

        // ---- carried "previous node" foot kinematics of this lane's leg (contact rows) ---------------------
        double Rp0 = 0, Rp1 = 0, Rp2 = 0;      // column c3 of R(q_{k-1})
        double Dp0 = 0, Dp1 = 0, Dp2 = 0;      // d(R r_{k-1,leg}) / d q_c   (c < 4)
        double fp0, fp1, fp2;                  // previous foot position
        double s_prev;
        if (k_begin > 0) {
            const double* xq = x + Mdl::x_off(N, k_begin - 1);
            const double* ur = x + Mdl::u_off(N, k_begin - 1) + 6 * leg + 3;
            const double qx = xq[3], qy = xq[4], qz = xq[5], qw = xq[6], r0 = ur[0], r1 = ur[1], r2 = ur[2];
            const Rot3 R = rot_matrix(qx, qy, qz, qw);
            Rp0 = pick3(c3, R.m00, R.m01, R.m02); Rp1 = pick3(c3, R.m10, R.m11, R.m12); Rp2 = pick3(c3, R.m20, R.m21, R.m22);
            drot_dq(c & 3, qx, qy, qz, qw, r0, r1, r2, Dp0, Dp1, Dp2);
            fp0 = xq[0] + R.m00 * r0 + R.m01 * r1 + R.m02 * r2;
            fp1 = xq[1] + R.m10 * r0 + R.m11 * r1 + R.m12 * r2;
            fp2 = xq[2] + R.m20 * r0 + R.m21 * r1 + R.m22 * r2;
            s_prev = x[Mdl::p_off(N, k_begin - 1) + 13 + 4 * leg];
        } else {
            fp0 = Rho[34 + 4 * leg]; fp1 = Rho[35 + 4 * leg]; fp2 = Rho[36 + 4 * leg];  // measured foot
            s_prev = Rho[33 + 4 * leg];
        }

        for (int k = k_begin; k < k_end; ++k) {
            const int slot = (k - k_begin) & 1;
            double* const sA = stA + slot * Q::NA;
            double* const sH = stH + slot * Q::TRI;
            double* const sC = stC + slot * Q::NC;
            const double* __restrict__ xk = x + Mdl::x_off(N, k);
            const double* __restrict__ uk = x + Mdl::u_off(N, k);
            const double* __restrict__ pk = x + Mdl::p_off(N, k);

            // ---- node-global primal pass (identical in every lane) ------------------------------------------
            const double qx = xk[3], qy = xk[4], qz = xk[5], qw = xk[6];
            const double w0 = xk[10], w1 = xk[11], w2 = xk[12];
            const Rot3 R = rot_matrix(qx, qy, qz, qw);
            double a0 = 0.0, a1 = 0.0, a2 = -g0;                                  // linear acceleration
            const double Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I2 * w2;
            double t0 = -(w1 * Iw2 - w2 * Iw1), t1 = -(w2 * Iw0 - w0 * Iw2), t2 = -(w0 * Iw1 - w1 * Iw0);  // torque
            double my_f0 = 0, my_f1 = 0, my_f2 = 0, my_r0 = 0, my_r1 = 0, my_r2 = 0, my_s = 0;
            double my_Rf0 = 0, my_Rf1 = 0, my_Rf2 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double f0 = uk[6 * i], f1 = uk[6 * i + 1], f2 = uk[6 * i + 2];
                const double r0 = uk[6 * i + 3], r1 = uk[6 * i + 4], r2 = uk[6 * i + 5];
                const double s = pk[13 + 4 * i];
                const double Rf0 = R.m00 * f0 + R.m01 * f1 + R.m02 * f2;
                const double Rf1 = R.m10 * f0 + R.m11 * f1 + R.m12 * f2;
                const double Rf2 = R.m20 * f0 + R.m21 * f1 + R.m22 * f2;
                a0 += s * f0 * inv_m; a1 += s * f1 * inv_m; a2 += s * f2 * inv_m;
                t0 += s * (r1 * Rf2 - r2 * Rf1); t1 += s * (r2 * Rf0 - r0 * Rf2); t2 += s * (r0 * Rf1 - r1 * Rf0);
                if (i == leg) {
                    my_f0 = f0; my_f1 = f1; my_f2 = f2; my_r0 = r0; my_r1 = r1; my_r2 = r2; my_s = s;
                    my_Rf0 = Rf0; my_Rf1 = Rf1; my_Rf2 = Rf2;
                }
            }
            const double vn0 = xk[7] + dt * a0, vn1 = xk[8] + dt * a1, vn2 = xk[9] + dt * a2;
            const double wn0 = w0 + dt * (t0 * iI0), wn1 = w1 + dt * (t1 * iI1), wn2 = w2 + dt * (t2 * iI2);
            const double pn0 = xk[0] + dt * vn0, pn1 = xk[1] + dt * vn1, pn2 = xk[2] + dt * vn2;
            const double y0 = dt * wn0, y1 = dt * wn1, y2 = dt * wn2;           // argument of the exponential map
            const double nn = sqrt(y0 * y0 + y1 * y1 + y2 * y2 + UB_EPS);
            double sh, ch;
            sincos(0.5 * nn, &sh, &ch);
            const double inv_n = 1.0 / nn;
            const double kap = sh * inv_n;
            const double e0 = y0 * kap, e1 = y1 * kap, e2 = y2 * kap, e3 = ch;   // e = aexp(dt w+)
            const double qn0 = qw * e0 + qx * e3 + qy * e2 - qz * e1;
            const double qn1 = qw * e1 + qy * e3 + qz * e0 - qx * e2;
            const double qn2 = qw * e2 + qz * e3 + qx * e1 - qy * e0;
            const double qn3 = qw * e3 - qx * e0 - qy * e1 - qz * e2;
            // Qw = d q+ / d w+ = dt * Lmat(q) * E,  E = d e / d y:  E[a][b] = kap d_ab + beta y_a y_b, E[3][b] = -kap/2 y_b
            const double beta = (0.5 * ch - kap) * inv_n * inv_n;
            // Lmat(q) columns (x y z w):  [qw qz -qy -qx], [-qz qw qx -qy], [qy -qx qw -qz], [qx qy qz qw]
            const double Ly0 = qw * y0 - qz * y1 + qy * y2, Ly1 = qz * y0 + qw * y1 - qx * y2,
                         Ly2 = -qy * y0 + qx * y1 + qw * y2, Ly3 = -qx * y0 - qy * y1 - qz * y2;
            const double m0 = beta * Ly0 - 0.5 * kap * qx, m1 = beta * Ly1 - 0.5 * kap * qy,
                         m2 = beta * Ly2 - 0.5 * kap * qz, m3 = beta * Ly3 - 0.5 * kap * qw;
            // Qw[a][b] = dt (kap Lmat[a][b] + y_b m_a)
            const double Q00 = dt * (kap * qw + y0 * m0), Q01 = dt * (-kap * qz + y1 * m0), Q02 = dt * (kap * qy + y2 * m0);
            const double Q10 = dt * (kap * qz + y0 * m1), Q11 = dt * (kap * qw + y1 * m1), Q12 = dt * (-kap * qx + y2 * m1);
            const double Q20 = dt * (-kap * qy + y0 * m2), Q21 = dt * (kap * qx + y1 * m2), Q22 = dt * (kap * qw + y2 * m2);
            const double Q30 = dt * (-kap * qx + y0 * m3), Q31 = dt * (-kap * qy + y1 * m3), Q32 = dt * (-kap * qz + y2 * m3);

            // ---- this lane's column of W = d w+ / d z ------------------------------------------------------------
            // leg lanes: f column c:  dt I^-1 s (r x R[:, c]);  r column c':  dt I^-1 s (e_c' x R f)
            double W0 = 0, W1 = 0, W2 = 0;
            {
                const double Rc0 = pick3(c3, R.m00, R.m01, R.m02), Rc1 = pick3(c3, R.m10, R.m11, R.m12),
                             Rc2 = pick3(c3, R.m20, R.m21, R.m22);
                const double ec0 = c3 == 0 ? 1.0 : 0.0, ec1 = c3 == 1 ? 1.0 : 0.0, ec2 = c3 == 2 ? 1.0 : 0.0;
                const bool fcol = c < 3;
                const double u0 = fcol ? my_r0 : ec0, u1 = fcol ? my_r1 : ec1, u2 = fcol ? my_r2 : ec2;
                const double v0 = fcol ? Rc0 : my_Rf0, v1 = fcol ? Rc1 : my_Rf1, v2 = fcol ? Rc2 : my_Rf2;
                const double sc = dt * my_s;
                W0 = sc * iI0 * (u1 * v2 - u2 * v1); W1 = sc * iI1 * (u2 * v0 - u0 * v2); W2 = sc * iI2 * (u0 * v1 - u1 * v0);
            }
            if (col_lane) {
                const int col = 13 + 6 * leg + c;
                sA[10 * 37 + col] = -W0; sA[11 * 37 + col] = -W1; sA[12 * 37 + col] = -W2;
                sA[3 * 37 + col] = -(Q00 * W0 + Q01 * W1 + Q02 * W2);
                sA[4 * 37 + col] = -(Q10 * W0 + Q11 * W1 + Q12 * W2);
                sA[5 * 37 + col] = -(Q20 * W0 + Q21 * W1 + Q22 * W2);
                sA[6 * 37 + col] = -(Q30 * W0 + Q31 * W1 + Q32 * W2);
                if (c < 3) {
                    sA[c * 37 + col]       = -(dt * dt) * (my_s * inv_m);
                    sA[(7 + c) * 37 + col] = -dt * (my_s * inv_m);
                }
            }
            // q columns: dt I^-1 sum_i s_i r_i x d(R f_i)/dq_c  — every leg's lanes c < 4 add their leg, then xor-reduce
            {
                double d0, d1, d2;
                drot_dq(c & 3, qx, qy, qz, qw, my_f0, my_f1, my_f2, d0, d1, d2);
                double z0 = my_s * (my_r1 * d2 - my_r2 * d1), z1 = my_s * (my_r2 * d0 - my_r0 * d2),
                       z2 = my_s * (my_r0 * d1 - my_r1 * d0);
                z0 += __shfl_xor_sync(0xffffffffu, z0, 8); z1 += __shfl_xor_sync(0xffffffffu, z1, 8); z2 += __shfl_xor_sync(0xffffffffu, z2, 8);
                z0 += __shfl_xor_sync(0xffffffffu, z0, 16); z1 += __shfl_xor_sync(0xffffffffu, z1, 16); z2 += __shfl_xor_sync(0xffffffffu, z2, 16);
                if (leg == 0 && c < 4) {  // lanes 0..3 own the q columns
                    const double G0 = dt * iI0 * z0, G1 = dt * iI1 * z1, G2 = dt * iI2 * z2;
                    const int col = 3 + c;
                    // Rmat(e) column c: d(q (x) e)/dq_c
                    const double r0c = c == 0 ? e3 : c == 1 ? e2 : c == 2 ? -e1 : e0;
                    const double r1c = c == 0 ? -e2 : c == 1 ? e3 : c == 2 ? e0 : e1;
                    const double r2c = c == 0 ? e1 : c == 1 ? -e0 : c == 2 ? e3 : e2;
                    const double r3c = c == 0 ? -e0 : c == 1 ? -e1 : c == 2 ? -e2 : e3;
                    sA[10 * 37 + col] = -G0; sA[11 * 37 + col] = -G1; sA[12 * 37 + col] = -G2;
                    sA[3 * 37 + col] = -(r0c + Q00 * G0 + Q01 * G1 + Q02 * G2);
                    sA[4 * 37 + col] = -(r1c + Q10 * G0 + Q11 * G1 + Q12 * G2);
                    sA[5 * 37 + col] = -(r2c + Q20 * G0 + Q21 * G1 + Q22 * G2);
                    sA[6 * 37 + col] = -(r3c + Q30 * G0 + Q31 * G1 + Q32 * G2);
                }
                if (leg == 1 && c < 3) {  // lanes 8..10 own the w columns: e_c + dt I^-1 (Iw x e_c - I_c (w x e_c))
                    const double ec0 = c == 0 ? 1.0 : 0.0, ec1 = c == 1 ? 1.0 : 0.0, ec2 = c == 2 ? 1.0 : 0.0;
                    const double Ic = pick3(c, I0, I1, I2);
                    const double G0 = ec0 + dt * iI0 * ((Iw1 * ec2 - Iw2 * ec1) - Ic * (w1 * ec2 - w2 * ec1));
                    const double G1 = ec1 + dt * iI1 * ((Iw2 * ec0 - Iw0 * ec2) - Ic * (w2 * ec0 - w0 * ec2));
                    const double G2 = ec2 + dt * iI2 * ((Iw0 * ec1 - Iw1 * ec0) - Ic * (w0 * ec1 - w1 * ec0));
                    const int col = 10 + c;
                    sA[10 * 37 + col] = -G0; sA[11 * 37 + col] = -G1; sA[12 * 37 + col] = -G2;
                    sA[3 * 37 + col] = -(Q00 * G0 + Q01 * G1 + Q02 * G2);
                    sA[4 * 37 + col] = -(Q10 * G0 + Q11 * G1 + Q12 * G2);
                    sA[5 * 37 + col] = -(Q20 * G0 + Q21 * G1 + Q22 * G2);
                    sA[6 * 37 + col] = -(Q30 * G0 + Q31 * G1 + Q32 * G2);
                }
                if (leg == 2 && c < 3) {  // lanes 16..18: the constant p / v entries
                    sA[c * 37 + c] = -1.0; sA[c * 37 + 7 + c] = -dt; sA[(7 + c) * 37 + 7 + c] = -1.0;
                }
            }

            // ---- state part of the objective, defects: lanes 0..12 own state entry `lane` -----------------------
            double cost_part = 0.0, bar_part = 0.0;
            {
                double dm = 0.0, dp = 0.0;  // Min(|q - qRef|^2, |q + qRef|^2) = CondExpGt(dm, dp, dp, dm)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double a = xk[3 + i] - pk[3 + i], bq = xk[3 + i] + pk[3 + i];
                    dm += a * a; dp += bq * bq;
                }
                const double sgn = dm > dp ? 1.0 : -1.0;
                if (lane < 13) {
                    const double wgt = lane < 2 ? 0.1 : (lane == 2 ? 10.0 : 1.0);
                    const bool isq = lane >= 3 && lane < 7;
                    const double res = wgt * (isq ? xk[lane] + sgn * pk[lane] : xk[lane] - pk[lane]);
                    cost_part = res * res;
                    r[L.grad + Mdl::x_off(N, k) + lane] = 2.0 * wgt * res;
                    sH[lane * 37 - (lane * (lane - 1)) / 2] = 2.0 * wgt * wgt + (BARRIER ? 1e-6 : 0.0);
                    const double xn = lane == 0 ? pn0 : lane == 1 ? pn1 : lane == 2 ? pn2 : lane == 3 ? qn0 : lane == 4 ? qn1
                                    : lane == 5 ? qn2 : lane == 6 ? qn3 : lane == 7 ? vn0 : lane == 8 ? vn1 : lane == 9 ? vn2
                                    : lane == 10 ? wn0 : lane == 11 ? wn1 : wn2;
                    r[L.g + 13 + 13 * k + lane] = x[Mdl::x_off(N, k + 1) + lane] - xn;
                }
            }

            // ---- this lane's input entry: objective, inequalities of its leg, barrier, Gauss-Newton rows -----------
            const double fxy = sqrt(my_f0 * my_f0 + my_f1 * my_f1 + UB_EPS);
            const double dr0 = my_r0 - hip0, dr1 = my_r1 - hip1, dr2 = my_r2 - hip2;
            const double nr = sqrt(dr0 * dr0 + dr1 * dr1 + dr2 * dr2 + UB_EPS);
            const double hA = -my_s * my_f2, hB = my_s * fxy - mu * my_f2, hC = my_s * nr - Llen;
            double bA = 0, dA = 0, ddA = 0, bB = 0, dB = 0, ddB = 0, bC = 0, dC = 0, ddC = 0;
            if (BARRIER) {
                barrier_eval(bar, hA, &bA, &dA, &ddA);
                barrier_eval(bar, hB, &bB, &dB, &ddB);
                barrier_eval(bar, hC, &bC, &dC, &ddC);
            }
            if (col_lane) {
                const int zi = 13 + 6 * leg + c;
                const int diag = zi * 37 - (zi * (zi - 1)) / 2;
                if (c < 3) {
                    const double inv_fxy = 1.0 / fxy;
                    const double gB0 = my_s * my_f0 * inv_fxy, gB1 = my_s * my_f1 * inv_fxy, gB2 = -mu;  // grad h_B wrt f
                    const double gA_c = c == 2 ? -my_s : 0.0;
                    const double gB_c = pick3(c, gB0, gB1, gB2);
                    const double fc = pick3(c, my_f0, my_f1, my_f2);
                    cost_part += 1e-8 * fc * fc;
                    r[L.grad + Mdl::u_off(N, k) + 6 * leg + c] = 2e-8 * fc + dA * gA_c + dB * gB_c;
                    r[L.h + 12 * k + 3 * leg + c] = pick3(c, hA, hB, hC);
                    bar_part = pick3(c, bA, bB, bC);
                    // row c of the f-f block: entries (c, c .. 2)
                    sH[diag] = ddA * gA_c * gA_c + ddB * gB_c * gB_c + 2e-8 + (BARRIER ? 1e-6 : 0.0);
                    if (c < 2) {
                        const double gA_1 = (c + 1 == 2) ? -my_s : 0.0;
                        const double gB_1 = c == 0 ? gB1 : gB2;
                        sH[diag + 1] = ddA * gA_c * gA_1 + ddB * gB_c * gB_1;
                    }
                    if (c < 1) sH[diag + 2] = ddA * gA_c * (-my_s) + ddB * gB_c * gB2;
                } else {
                    const double inv_nr = 1.0 / nr;
                    const double gC0 = my_s * dr0 * inv_nr, gC1 = my_s * dr1 * inv_nr, gC2 = my_s * dr2 * inv_nr;
                    const double gC_c = pick3(c3, gC0, gC1, gC2);
                    const double rc = pick3(c3, my_r0, my_r1, my_r2) - pk[14 + 4 * leg + c3];
                    cost_part += rc * rc;
                    r[L.grad + Mdl::u_off(N, k) + 6 * leg + c] = 2.0 * rc + dC * gC_c;
                    sH[diag] = ddC * gC_c * gC_c + 2.0 + (BARRIER ? 1e-6 : 0.0);
                    if (c3 < 2) sH[diag + 1] = ddC * gC_c * (c3 == 0 ? gC1 : gC2);
                    if (c3 < 1) sH[diag + 2] = ddC * gC_c * gC2;
                }
            }
            // stage cost / barrier value: deterministic xor-tree over the warp
            cost_part = warp_sum(cost_part);
            bar_part  = warp_sum(bar_part);
            if (lane == 0) {
                stage_cost[((long long)b * (N + 1) + k) * 2]     = cost_part;
                stage_cost[((long long)b * (N + 1) + k) * 2 + 1] = bar_part;
            }

            // ---- contact rows of this lane's leg ----------------------------------------------------------------------
            {
                const double ft0 = xk[0] + R.m00 * my_r0 + R.m01 * my_r1 + R.m02 * my_r2;
                const double ft1 = xk[1] + R.m10 * my_r0 + R.m11 * my_r1 + R.m12 * my_r2;
                const double ft2 = xk[2] + R.m20 * my_r0 + R.m21 * my_r1 + R.m22 * my_r2;
                const double c0 = (1.0 - s_prev) * my_s, ss = s_prev * my_s;
                const double Rc0 = pick3(c3, R.m00, R.m01, R.m02), Rc1 = pick3(c3, R.m10, R.m11, R.m12),
                             Rc2 = pick3(c3, R.m20, R.m21, R.m22);
                double D0, D1, D2;
                drot_dq(c & 3, qx, qy, qz, qw, my_r0, my_r1, my_r2, D0, D1, D2);
                double* const cl = sC + leg * 80;
                if (c < 4) {  // d foot / d q_c, and the contact values (row c)
                    cl[3 + c] = c0 * D2;
                    cl[20 + 3 + c] = ss * D0; cl[40 + 3 + c] = ss * D1; cl[60 + 3 + c] = ss * D2;
                    cl[20 + 13 + c] = -ss * Dp0; cl[40 + 13 + c] = -ss * Dp1; cl[60 + 13 + c] = -ss * Dp2;
                    const double val = c == 0 ? c0 * ft2 : ss * (c == 1 ? ft0 - fp0 : c == 2 ? ft1 - fp1 : ft2 - fp2);
                    r[L.g + 13 + 13 * N + 16 * k + 4 * leg + c] = val;
                }
                if (c < 3) {  // d foot / d r_c = R[:, c];  d foot / d p = I
                    const double pfac = k > 0 ? -ss : 0.0;
                    cl[7 + c] = c0 * Rc2;
                    cl[20 + 7 + c] = ss * Rc0; cl[40 + 7 + c] = ss * Rc1; cl[60 + 7 + c] = ss * Rc2;
                    cl[20 + 17 + c] = -ss * Rp0; cl[40 + 17 + c] = -ss * Rp1; cl[60 + 17 + c] = -ss * Rp2;
                    cl[(c + 1) * 20 + c]      = ss;
                    cl[(c + 1) * 20 + 10 + c] = pfac;
                    if (c == 2) cl[2] = c0;
                }
                // carry to the next node
                Rp0 = Rc0; Rp1 = Rc1; Rp2 = Rc2; Dp0 = D0; Dp1 = D1; Dp2 = D2;
                fp0 = ft0; fp1 = ft1; fp2 = ft2; s_prev = my_s;
            }

            // ---- node pair complete: hand the staged blocks to the TMA engine ------------------------------------------
            if (slot == 1 || k + 1 == k_end) {
                const int k_pair = k - slot;
                const int cnt    = slot + 1;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (cnt == 2) {
                    if (lane == 0) {
                        bulk_store(r + L.A + (long long)k_pair * Q::NA, stA, Q::PAIR_A * 8);
                        bulk_store(r + L.H + (long long)k_pair * Q::TRI, stH, Q::PAIR_H * 8);
                        bulk_store(r + L.C + (long long)k_pair * Q::NC, stC, Q::PAIR_C * 8);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                } else {  // odd tail (never taken when the run length is even): plain coalesced stores
                    for (int e = lane; e < Q::NA; e += 32) r[L.A + (long long)k_pair * Q::NA + e] = stA[e];
                    for (int e = lane; e < Q::TRI; e += 32) r[L.H + (long long)k_pair * Q::TRI + e] = stH[e];
                    for (int e = lane; e < Q::NC; e += 32) r[L.C + (long long)k_pair * Q::NC + e] = stC[e];
                }
                __syncwarp();
            }
        }

        // ---- terminal state x_N: objective gradient and diagonal block (last run of the trajectory) -----------------
        if (k_end == N) {
            const double* __restrict__ xN = x + Mdl::x_off(N, N);
            const double* __restrict__ pN = x + Mdl::p_off(N, N);
            double dm = 0.0, dp = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a = xN[3 + i] - pN[3 + i], bq = xN[3 + i] + pN[3 + i];
                dm += a * a; dp += bq * bq;
            }
            const double sgn = dm > dp ? 1.0 : -1.0;
            double cpart = 0.0, hdiag = 0.0;
            if (lane < 13) {
                const double wgt = lane < 2 ? 0.1 : (lane == 2 ? 10.0 : 1.0);
                const bool isq = lane >= 3 && lane < 7;
                const double res = wgt * (isq ? xN[lane] + sgn * pN[lane] : xN[lane] - pN[lane]);
                cpart = res * res;
                hdiag = 2.0 * wgt * wgt + (BARRIER ? 1e-6 : 0.0);
                r[L.grad + Mdl::x_off(N, N) + lane] = 2.0 * wgt * res;
            }
            cpart = warp_sum(cpart);
            if (lane == 0) {
                stage_cost[((long long)b * (N + 1) + N) * 2]     = cpart;
                stage_cost[((long long)b * (N + 1) + N) * 2 + 1] = 0.0;
            }
            for (int row = 0; row < 13; ++row) {  // packed upper triangle, row by row
                const double dv = __shfl_sync(0xffffffffu, hdiag, row);
                const int base  = row * 13 - (row * (row - 1)) / 2;
                if (lane < 13 - row) r[L.HN + base + lane] = lane == 0 ? dv : 0.0;
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace ub
