// KKT stage sweep of the quadruped single-rigid-body NMPC, structured version: a warp walks a run of up to 10
// consecutive shooting nodes of one trajectory and OWNS ONE NODE AT A TIME when it fills the blocks.
//
// Why a second kernel: the generic sweep (sweep.cuh) pushes one dense tangent per thread through the whole stage
// (~50 000 fp64 instructions per node, ~250 registers).  Here the chain rule is applied by hand through the model's
// bottlenecks — everything downstream of (q, w, f_i, r_i) goes through the 3-vector w+ = w + dt I^-1 tau and the 4x3
// matrix Qw = d q+ / d w+ — and the work is split in two phases so nothing is computed 32 times:
//
//   phase 0  the run's slice of the flat Ungar vector (x_{k0-1..k0+10}, u_{k0-1..k0+9}, p_{k0-1..k0+9}; 5.9 KB) is
//            copied into shared memory with cp.async (LDGSTS, 8-byte granules: the Ungar layout is not 16-byte aligned),
//            so the only exposed HBM latency is once per run and every later read is an LDS
//   phase A  thread-per-node: lane j computes the node-global primal "core" of node k0 + j - 1 (rotation matrix, Lie-Euler
//            step, exponential map and Qw; one sincos / sqrt / divide per node) and parks 39 doubles in shared memory
//   phase B  warp-per-node: for each node the 32 lanes take one column of d w+/d z each (24 input columns, 4 + 3 state
//            columns), the rows of the Gauss-Newton block and of the contact Jacobians of "their" leg, and write ONLY the
//            structurally non-zero slots into a per-warp shared-memory image of the A, H and C blocks of a node PAIR.
//            The image is zeroed once per kernel; the ~65 % zeros of the dense blocks never cost an instruction again
//   stores   one elected lane hands the image to the TMA engine: three cp.async.bulk.global.shared::cta copies per pair
//            (7 696 + 11 248 + 5 120 B; 16-byte aligned because nodes are paired; SASS UBLKCP).  The small vectors
//            (g, h, grad: 5 % of the bytes) leave as plain coalesced stores from registers.
//
// Reference lines restated: quadruped.example.cpp:148-203 (dynamics), :209-251 (objective), :279-303 (contact rows),
// :312-338 (inequalities); soft_sqp.hpp:141-158, :245-264 (assembly); soft_inequality_constraint.hpp:133-190.
#pragma once

#include "sweep.cuh"

namespace ub {

struct QuadrupedStructured {
    static constexpr int NX = 13, NU = 24, NZ = 37, NP = 29, TRI = 703, NA = NX * NZ, NC = 320;
    static constexpr int PAIR_A = 2 * NA, PAIR_H = 2 * TRI, PAIR_C = 2 * NC;
    static constexpr int STAGE = PAIR_A + PAIR_H + PAIR_C;  // 3008 doubles
    static constexpr int RUN = 10;                          // nodes per run (even)
    static constexpr int CORE = 77;                         // doubles per node core (odd stride: conflict-free)
    static constexpr int NRHO = 49;                         // the shared parameter block Rho (quadruped.example.cpp:129-134)
    static constexpr int oCORE = STAGE, oXS = oCORE + 848, oUS = oXS + (RUN + 2) * NX, oPS = oUS + (RUN + 1) * NU,
                         oRHO = oPS + (RUN + 1) * NP + 1,   // 4596
                         PER_WARP = 4646;                   // doubles per team (37 168 B, multiple of 16; 6 teams per SM)
    static_assert(oRHO + NRHO <= PER_WARP && (PER_WARP * 8) % 16 == 0, "per-warp shared memory layout");
    static_assert((RUN + 1) * CORE <= 848, "core slots: halo node + RUN nodes");
    static constexpr int WARPS = 2;                         // one team of two warps per CTA; 6 CTAs / SM -> 12 warps
    static constexpr int SMEM_BYTES = PER_WARP * 8;         // the team shares one staging image + inputs (33 600 B)
    // core slot layout
    static constexpr int cR = 0, cQ = 9, cE = 21, cXN = 25, cSGN = 38, cM = 39;  // cM: the four 3x3 matrices d(R(q) v)/dq_c
};

__device__ __forceinline__ double pick3(int c, double a0, double a1, double a2) { return c == 0 ? a0 : (c == 1 ? a1 : a2); }

// R with R v = v + 2 w (qv x v) + 2 qv x (qv x v)  (Eigen/src/Geometry/Quaternion.h:531-541), row-major into out[9].
__device__ __forceinline__ void rot_matrix(double x, double y, double z, double w, double* out) {
    out[0] = 1.0 - 2.0 * (y * y + z * z); out[1] = 2.0 * (x * y - w * z);       out[2] = 2.0 * (x * z + w * y);
    out[3] = 2.0 * (x * y + w * z);       out[4] = 1.0 - 2.0 * (x * x + z * z); out[5] = 2.0 * (y * z - w * x);
    out[6] = 2.0 * (x * z - w * y);       out[7] = 2.0 * (y * z + w * x);       out[8] = 1.0 - 2.0 * (x * x + y * y);
}

// The four matrices M_c with d(R(q) v)/dq_c = M_c v for the Eigen rotation formula (R v = v + 2 w (qv x v) + 2 qv x (qv x v)):
//   c < 3:  M_c = 2 w [e_c]x + 2 (e_c qv^T + qv e_c^T - 2 q_c I),      c = 3 (w):  M_3 = 2 [qv]x.        Row-major, 9 entries each.
// Computed once per node in phase A (thread-per-node); phase B applies them to f, r and the previous node's r as 3x3 mat-vecs.
__device__ __forceinline__ void drot_matrices(double x, double y, double z, double w, double* __restrict__ M) {
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z, tw = 2.0 * w;
    // c = 0
    M[0] = 0.0;        M[1] = ty;         M[2] = tz;
    M[3] = ty;         M[4] = -2.0 * tx;  M[5] = -tw;
    M[6] = tz;         M[7] = tw;         M[8] = -2.0 * tx;
    // c = 1
    M[9]  = -2.0 * ty; M[10] = tx;        M[11] = tw;
    M[12] = tx;        M[13] = 0.0;       M[14] = tz;
    M[15] = -tw;       M[16] = tz;        M[17] = -2.0 * ty;
    // c = 2
    M[18] = -2.0 * tz; M[19] = -tw;       M[20] = tx;
    M[21] = tw;        M[22] = -2.0 * tz; M[23] = ty;
    M[24] = tx;        M[25] = ty;        M[26] = 0.0;
    // c = 3
    M[27] = 0.0;       M[28] = -tz;       M[29] = ty;
    M[30] = tz;        M[31] = 0.0;       M[32] = -tx;
    M[33] = -ty;       M[34] = tx;        M[35] = 0.0;
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
__device__ __forceinline__ void async_copy8(double* smem, const double* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// Phase A: node-global primal core of one node (thread-per-node), written to `core`.
__device__ __forceinline__ void node_core(const double* __restrict__ xk, const double* __restrict__ uk,
                                          const double* __restrict__ pk, double dt, double inv_m, double g0, double I0,
                                          double I1, double I2, double iI0, double iI1, double iI2, double* __restrict__ core) {
    using Q = QuadrupedStructured;
    const double qx = xk[3], qy = xk[4], qz = xk[5], qw = xk[6];
    const double w0 = xk[10], w1 = xk[11], w2 = xk[12];
    double R[9];
    rot_matrix(qx, qy, qz, qw, R);
    double a0 = 0.0, a1 = 0.0, a2 = -g0;  // p'' = -g e_z + sum s_i f_i / m                   (quadruped.example.cpp:168,176)
    const double Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I2 * w2;
    double t0 = -(w1 * Iw2 - w2 * Iw1), t1 = -(w2 * Iw0 - w0 * Iw2), t2 = -(w0 * Iw1 - w1 * Iw0);  // -w x I w   (:169)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double f0 = uk[6 * i], f1 = uk[6 * i + 1], f2 = uk[6 * i + 2];
        const double r0 = uk[6 * i + 3], r1 = uk[6 * i + 4], r2 = uk[6 * i + 5];
        const double s = pk[13 + 4 * i];
        const double Rf0 = R[0] * f0 + R[1] * f1 + R[2] * f2, Rf1 = R[3] * f0 + R[4] * f1 + R[5] * f2,
                     Rf2 = R[6] * f0 + R[7] * f1 + R[8] * f2;
        a0 += s * f0 * inv_m; a1 += s * f1 * inv_m; a2 += s * f2 * inv_m;
        t0 += s * (r1 * Rf2 - r2 * Rf1); t1 += s * (r2 * Rf0 - r0 * Rf2); t2 += s * (r0 * Rf1 - r1 * Rf0);  // (:177)
    }
    // Lie-group semi-implicit Euler (:197-200)
    const double vn0 = xk[7] + dt * a0, vn1 = xk[8] + dt * a1, vn2 = xk[9] + dt * a2;
    const double wn0 = w0 + dt * (t0 * iI0), wn1 = w1 + dt * (t1 * iI1), wn2 = w2 + dt * (t2 * iI2);
    const double y0 = dt * wn0, y1 = dt * wn1, y2 = dt * wn2;
    const double nn = sqrt(y0 * y0 + y1 * y1 + y2 * y2 + UB_EPS);  // Utils::ApproximateNorm
    double sh, ch;
    sincos(0.5 * nn, &sh, &ch);
    const double inv_n = 1.0 / nn, kap = sh * inv_n;
    const double e0 = y0 * kap, e1 = y1 * kap, e2 = y2 * kap, e3 = ch;  // Utils::ApproximateExponentialMap
#pragma unroll
    for (int i = 0; i < 9; ++i) core[Q::cR + i] = R[i];
    core[Q::cE] = e0; core[Q::cE + 1] = e1; core[Q::cE + 2] = e2; core[Q::cE + 3] = e3;
    core[Q::cXN + 0] = xk[0] + dt * vn0; core[Q::cXN + 1] = xk[1] + dt * vn1; core[Q::cXN + 2] = xk[2] + dt * vn2;
    core[Q::cXN + 3] = qw * e0 + qx * e3 + qy * e2 - qz * e1;  // q (x) e, Eigen product (Quaternion.h:487-498)
    core[Q::cXN + 4] = qw * e1 + qy * e3 + qz * e0 - qx * e2;
    core[Q::cXN + 5] = qw * e2 + qz * e3 + qx * e1 - qy * e0;
    core[Q::cXN + 6] = qw * e3 - qx * e0 - qy * e1 - qz * e2;
    core[Q::cXN + 7] = vn0; core[Q::cXN + 8] = vn1; core[Q::cXN + 9] = vn2;
    core[Q::cXN + 10] = wn0; core[Q::cXN + 11] = wn1; core[Q::cXN + 12] = wn2;
    // Qw = d q+ / d w+ = dt Lmat(q) E,  E = d e / d y:  E[a][b] = kap d_ab + beta y_a y_b (a < 3),  E[3][b] = -kap/2 y_b
    const double beta = (0.5 * ch - kap) * inv_n * inv_n;
    const double Ly0 = qw * y0 - qz * y1 + qy * y2, Ly1 = qz * y0 + qw * y1 - qx * y2, Ly2 = -qy * y0 + qx * y1 + qw * y2,
                 Ly3 = -qx * y0 - qy * y1 - qz * y2;
    const double m0 = beta * Ly0 - 0.5 * kap * qx, m1 = beta * Ly1 - 0.5 * kap * qy, m2 = beta * Ly2 - 0.5 * kap * qz,
                 m3 = beta * Ly3 - 0.5 * kap * qw;
    core[Q::cQ + 0] = dt * (kap * qw + y0 * m0);  core[Q::cQ + 1] = dt * (-kap * qz + y1 * m0); core[Q::cQ + 2] = dt * (kap * qy + y2 * m0);
    core[Q::cQ + 3] = dt * (kap * qz + y0 * m1);  core[Q::cQ + 4] = dt * (kap * qw + y1 * m1);  core[Q::cQ + 5] = dt * (-kap * qx + y2 * m1);
    core[Q::cQ + 6] = dt * (-kap * qy + y0 * m2); core[Q::cQ + 7] = dt * (kap * qx + y1 * m2);  core[Q::cQ + 8] = dt * (kap * qw + y2 * m2);
    core[Q::cQ + 9] = dt * (-kap * qx + y0 * m3); core[Q::cQ + 10] = dt * (-kap * qy + y1 * m3); core[Q::cQ + 11] = dt * (-kap * qz + y2 * m3);
    // Min(|q - qRef|^2, |q + qRef|^2) = CondExpGt(dm, dp, dp, dm): '+' branch only when dm > dp (utils.hpp:976)
    double dm = 0.0, dp = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double a = xk[3 + i] - pk[3 + i], bq = xk[3 + i] + pk[3 + i];
        dm += a * a; dp += bq * bq;
    }
    core[Q::cSGN] = dm > dp ? 1.0 : -1.0;
    drot_matrices(qx, qy, qz, qw, core + Q::cM);
}

// Branch-free RelaxedPolyBarrierFunction pieces (same polynomials as barrier_eval, selected instead of branched).
__device__ __forceinline__ void barrier_eval_sel(const BarrierCoef<double>& B, double h, double& b0, double& dz, double& d2z) {
    const double x = -h;
    const double q0 = (0.5 * B.a1 * x + B.b1) * x + B.c1, q1 = B.a1 * x + B.b1;
    const double c0 = ((1.0 / 3.0 * B.a2 * x + 0.5 * B.b2) * x + B.c2) * x + B.d2, c1 = (B.a2 * x + B.b2) * x + B.c2,
                 c2 = 2.0 * B.a2 * x + B.b2;
    const bool neg = x < 0.0, mid = x < B.eps;
    b0  = neg ? q0 : (mid ? c0 : 0.0);
    dz  = -(neg ? q1 : (mid ? c1 : 0.0));
    d2z = neg ? B.a1 : (mid ? c2 : 0.0);
}

// One CTA = one TEAM of two warps sharing one staging image: warp w fills slot w of every node pair of the run.
template <bool BARRIER>
__global__ void __launch_bounds__(64, 6)
quadruped_structured_kernel(const double* __restrict__ xp_all, long long ld_xp, double* __restrict__ rec_all, long long ld_rec,
                            double* __restrict__ partials, int N, int run_len, int runs_per_traj, long long total_runs,
                            RecLayout L, BarrierCoef<double> bar, unsigned int* __restrict__ sched) {
    using Q = QuadrupedStructured;
    using Mdl = Quadruped;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double* const wsm   = reinterpret_cast<double*>(smem_raw);
    double* const stA   = wsm;
    double* const stH   = wsm + Q::PAIR_A;
    double* const stC   = wsm + Q::PAIR_A + Q::PAIR_H;
    double* const cores = wsm + Q::oCORE;
    double* const xs    = wsm + Q::oXS;
    double* const us    = wsm + Q::oUS;
    double* const ps    = wsm + Q::oPS;
    double* const rs    = wsm + Q::oRHO;
    for (int e = tid; e < Q::STAGE; e += 64) wsm[e] = 0.0;
    __syncthreads();

    // ---- per-lane roles and staging addresses (the slot is fixed per warp, so all of these are loop constants) ----------
    const int leg = lane >> 3, c = lane & 7;  // column lanes: c < 6 -> (f0 f1 f2 r0 r1 r2) of `leg`
    const bool col_lane = c < 6, fcol = c < 3;
    const int c3 = fcol ? c : (c < 6 ? c - 3 : 0);
    const int cq = c & 3;
    const bool qcol = lane < 4, wcol = lane >= 8 && lane < 11, pvlane = lane >= 16 && lane < 19;
    double* const sA  = stA + w * Q::NA;
    double* const sH  = stH + w * Q::TRI;
    double* const cl  = stC + w * Q::NC + leg * 80;
    double* const aCol = sA + 13 + 6 * leg + c;           // input column of this lane (valid if col_lane)
    double* const aSt  = sA + (qcol ? 3 + c : 10 + c);    // state column of this lane (valid if qcol || wcol)
    const int zi       = 13 + 6 * leg + c;
    double* const hIn  = sH + zi * 37 - (zi * (zi - 1)) / 2;
    double* const hSt  = sH + lane * 37 - (lane * (lane - 1)) / 2;  // valid if lane < 13
    const double ec0 = c3 == 0 ? 1.0 : 0.0, ec1 = c3 == 1 ? 1.0 : 0.0, ec2 = c3 == 2 ? 1.0 : 0.0;
    bool pending = false;  // a bulk store may still be reading the staging image (CTA-uniform)

    // Runs are CLAIMED, not statically strided: the first one is the CTA's index, every further one comes from an atomic counter
    // (sched[0]).  A CTA that becomes resident late — another kernel (the NCCL all-gather of the previous step, a neighbour's H2D
    // chunk sweep) holds part of an SM — then simply claims fewer runs instead of stretching the launch by a second wave.
    __shared__ long long s_next;
    long long run = blockIdx.x;
    while (run < total_runs) {
        const long long b = run / runs_per_traj;
        const int run_in_traj = int(run - b * runs_per_traj);
        const int k0    = run_in_traj * run_len;
        const int nodes = min(N, k0 + run_len) - k0;
        const double* __restrict__ x = xp_all + b * ld_xp;
        double* __restrict__ r       = rec_all + b * ld_rec;
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
            r[L.g + lane] = gv;
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

            if (pending) {  // the previous pair's bulk stores must have finished reading the image
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                pending = false;
            }
            __syncthreads();

            if (active) {
                // ---- A = d(x_{k+1} - f)/dz: rows w+ (10-12) = -W, rows q+ (3-6) = -(Qw W [+ Rmat(e)]), rows p+/v+ constants --
                if (col_lane) {
                    aCol[10 * 37] = -W0; aCol[11 * 37] = -W1; aCol[12 * 37] = -W2;
                    aCol[3 * 37] = -(Qw[0] * W0 + Qw[1] * W1 + Qw[2] * W2);
                    aCol[4 * 37] = -(Qw[3] * W0 + Qw[4] * W1 + Qw[5] * W2);
                    aCol[5 * 37] = -(Qw[6] * W0 + Qw[7] * W1 + Qw[8] * W2);
                    aCol[6 * 37] = -(Qw[9] * W0 + Qw[10] * W1 + Qw[11] * W2);
                    if (fcol) {
                        const double sm = s * inv_m * dt;
                        aCol[c * 37]       = -dt * sm;
                        aCol[(7 + c) * 37] = -sm;
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
                    aSt[10 * 37] = -G0; aSt[11 * 37] = -G1; aSt[12 * 37] = -G2;
                    aSt[3 * 37] = -(add0 + Qw[0] * G0 + Qw[1] * G1 + Qw[2] * G2);
                    aSt[4 * 37] = -(add1 + Qw[3] * G0 + Qw[4] * G1 + Qw[5] * G2);
                    aSt[5 * 37] = -(add2 + Qw[6] * G0 + Qw[7] * G1 + Qw[8] * G2);
                    aSt[6 * 37] = -(add3 + Qw[9] * G0 + Qw[10] * G1 + Qw[11] * G2);
                }
                if (pvlane) {  // lanes 16..18: the constant p / v entries
                    sA[c * 37 + c] = -1.0; sA[c * 37 + 7 + c] = -dt; sA[(7 + c) * 37 + 7 + c] = -1.0;
                }

                // ---- state part of the objective and the defects: lanes 0..12 own state entry `lane` -----------------------
                if (lane < 13) {
                    const double sgn = co[Q::cSGN];
                    const double wgt = lane < 2 ? 0.1 : (lane == 2 ? 10.0 : 1.0);  // Vector3r{0.1, 0.1, 10} (:228)
                    const bool isq   = lane >= 3 && lane < 7;
                    const double res = wgt * (isq ? xk[lane] + sgn * pk[lane] : xk[lane] - pk[lane]);
                    cost_acc += res * res;
                    r[L.grad + Mdl::x_off(N, k) + lane] = 2.0 * wgt * res;
                    *hSt = 2.0 * wgt * wgt + (BARRIER ? 1e-6 : 0.0);
                    const double gv = xk[13 + lane] - co[Q::cXN + lane];  // x_{k+1} - f(x_k, u_k)   (:276)
                    r[L.g + 13 + 13 * k + lane] = gv;
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
                        r[L.grad + Mdl::u_off(N, k) + 6 * leg + c] = 2e-8 * fc + dA * gA_c + dB * gB_c;
                        if (c < 2) {
                            const double hv = c == 0 ? hA : hB;
                            r[L.h + 12 * k + 3 * leg + c] = hv;
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
                        r[L.grad + Mdl::u_off(N, k) + 6 * leg + c] = 2.0 * rc + dC * gC_c;
                        if (c == 3) {
                            r[L.h + 12 * k + 3 * leg + 2] = hC;
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
                        cl[3 + c] = c0 * D2;
                        cl[20 + 3 + c] = ss * D0; cl[40 + 3 + c] = ss * D1; cl[60 + 3 + c] = ss * D2;
                        cl[20 + 13 + c] = -ss * Dp0; cl[40 + 13 + c] = -ss * Dp1; cl[60 + 13 + c] = -ss * Dp2;
                        const double val = c == 0 ? c0 * ft2 : ss * (c == 1 ? ft0 - fp0 : c == 2 ? ft1 - fp1 : ft2 - fp2);
                        r[L.g + 13 + 13 * N + 16 * k + 4 * leg + c] = val;
                        gmax = fmax(gmax, fabs(val));
                    }
                    if (c < 3) {  // d foot / d r_c = R[:, c];  d foot / d p = I
                        cl[7 + c] = c0 * Rc2;
                        cl[20 + 7 + c] = ss * Rc0; cl[40 + 7 + c] = ss * Rc1; cl[60 + 7 + c] = ss * Rc2;
                        cl[20 + 17 + c] = -ss * Rp0; cl[40 + 17 + c] = -ss * Rp1; cl[60 + 17 + c] = -ss * Rp2;
                        cl[(c + 1) * 20 + c]      = ss;
                        cl[(c + 1) * 20 + 10 + c] = k > 0 ? -ss : 0.0;
                        if (c == 2) cl[2] = c0;
                    }
                }
            }

            // ---- node pair complete: hand the staged blocks to the TMA engine ------------------------------------------------------
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            const int k_pair = k0 + 2 * p;
            if (2 * p + 1 < nodes) {
                if (tid == 0) {
                    bulk_store(r + L.A + (long long)k_pair * Q::NA, stA, Q::PAIR_A * 8);
                    bulk_store(r + L.H + (long long)k_pair * Q::TRI, stH, Q::PAIR_H * 8);
                    bulk_store(r + L.C + (long long)k_pair * Q::NC, stC, Q::PAIR_C * 8);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                pending = true;
            } else {  // odd tail (never taken when the run length is even): plain coalesced stores of slot 0
                for (int e = tid; e < Q::NA; e += 64) r[L.A + (long long)k_pair * Q::NA + e] = stA[e];
                for (int e = tid; e < Q::TRI; e += 64) r[L.H + (long long)k_pair * Q::TRI + e] = stH[e];
                for (int e = tid; e < Q::NC; e += 64) r[L.C + (long long)k_pair * Q::NC + e] = stC[e];
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
                r[L.grad + Mdl::x_off(N, N) + lane] = 2.0 * wgt * res;
            }
            for (int row = 0; row < 13; ++row) {  // packed upper triangle, row by row
                const double dv = __shfl_sync(0xffffffffu, hdiag, row);
                const int base  = row * 13 - (row * (row - 1)) / 2;
                if (lane < 13 - row) r[L.HN + base + lane] = lane == 0 ? dv : 0.0;
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
    if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        // the last CTA to leave re-arms the scheduler for the next launch (stream order makes the zeros visible to it)
        __threadfence();
        if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
            __threadfence();
        }
    }
}

}  // namespace ub
