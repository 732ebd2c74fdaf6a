// KKT stage sweep for the two small models (quadrotor nz = 17, RC car nz = 8): one warp per trajectory, EIGHT LANES PER
// SHOOTING NODE, four nodes per step.
//
// Why a third kernel: a node of these models produces only 1.07 KB (quadrotor fp32) / 456 B (RC car fp32), so the HBM
// roofline leaves ~190 / ~80 issue slots per node per SM.  The generic sweep (one thread per (node, tangent): 1063 / 355
// instructions per node) is issue-bound at ~0.2 of the roofline; the thread-per-node sweep (sweep_tpn.cuh) removes the
// redundancy but has to stage a whole trajectory per warp (30-54 KB), which caps residency at 4-7 warps per SM and makes it
// latency-bound.  This kernel keeps the cheap arithmetic of the latter and the occupancy of the former:
//   phase 0  the trajectory's flat Ungar vector -> shared memory with cp.async (4 / 8-byte granules)
//   phase A  thread-per-node "core": everything that is common to the node's columns (quadrotor: third column of R(q),
//            Lie-Euler step, exponential map, Qw = d q+/d w+;  RC car: the Pacejka force model differentiated with 5
//            tangents and the Euler / yaw-rotation step) -> <= 35 scalars per node in shared memory
//   phase B  8 lanes per node, 4 nodes per warp step: each lane writes the non-zero entries of "its" columns of A, its
//            diagonal entries of the Gauss-Newton block, its entries of g / h / grad, into a 4-node staging image that is
//            zeroed once per kernel (1.5 KB per node in fp32)
//   stores   lane 0 hands the image to the TMA engine (cp.async.bulk.global.shared::cta, two copies per 4 nodes, 16-byte
//            aligned because 4 node blocks always are); small vectors leave as plain stores
// Per-warp shared memory: 9-15 KB -> 15-24 warps per SM.
//
// Reference lines restated: quadrotor.example.cpp:126-190, :196-238, :271-291; rc_car.example.cpp:131-185, :197-231,
// :264-285; soft_sqp.hpp:141-158, :245-264.
#pragma once

#include "sweep_tpn.cuh"

namespace ub {

template <class Mdl>
struct SmallPolicy;

// ---------------------------------------------------------------------------------------------------------------------
// Quadrotor
// ---------------------------------------------------------------------------------------------------------------------
template <>
struct SmallPolicy<Quadrotor> {
    static constexpr int CORE = 35;  // odd stride
    // core layout: cc[3] | Q[12] | e[4] | xn[13] | sgn | s = T/m
    static constexpr int oC = 0, oQ = 3, oE = 15, oXN = 19, oSGN = 32, oS = 33;

    template <class T>
    __device__ __forceinline__ static void core(const T* __restrict__ sx, int N, int k, T* __restrict__ co) {
        const T* P = sx + Quadrotor::n_dec(N);
        const T* x = sx + Quadrotor::x_off(N, k);
        const T* u = sx + Quadrotor::u_off(N, k);
        const T dt = P[0], m = P[1], I0 = P[2], I1 = P[3], I2 = P[4], g0 = P[17], b = P[18], dd = P[19];
        const T inv_m = T(1) / m;
        const T qx = x[3], qy = x[4], qz = x[5], qw = x[6], w0 = x[10], w1 = x[11], w2 = x[12];
        T Tz = T(0), mo0 = T(0), mo1 = T(0), mo2 = T(0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const T r2 = u[i] * u[i];
            Tz += b * r2; mo0 += P[6 + 3 * i] * (b * r2); mo1 -= P[5 + 3 * i] * (b * r2); mo2 += ((i & 1) ? -dd : dd) * r2;
        }
        const T c0 = T(2) * (qw * qy + qz * qx), c1 = T(2) * (qz * qy - qw * qx), c2 = T(1) - T(2) * (qx * qx + qy * qy);
        const T a0 = Tz * c0 * inv_m, a1 = Tz * c1 * inv_m, a2 = (Tz * c2 - m * g0) * inv_m;
        const T Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I2 * w2;
        const T t0 = mo0 - (w1 * Iw2 - w2 * Iw1), t1 = mo1 - (w2 * Iw0 - w0 * Iw2), t2 = mo2 - (w0 * Iw1 - w1 * Iw0);
        const T vn0 = x[7] + dt * a0, vn1 = x[8] + dt * a1, vn2 = x[9] + dt * a2;
        const T wn0 = w0 + dt * (t0 / I0), wn1 = w1 + dt * (t1 / I1), wn2 = w2 + dt * (t2 / I2);
        const T y0 = dt * wn0, y1 = dt * wn1, y2 = dt * wn2;
        const T nn = m_sqrt(y0 * y0 + y1 * y1 + y2 * y2 + T(UB_EPS));
        T sh, ch;
        m_sincos(T(0.5) * nn, &sh, &ch);
        const T inv_n = T(1) / nn, kap = sh * inv_n;
        const T e0 = y0 * kap, e1 = y1 * kap, e2 = y2 * kap, e3 = ch;
        co[oC] = c0; co[oC + 1] = c1; co[oC + 2] = c2;
        co[oE] = e0; co[oE + 1] = e1; co[oE + 2] = e2; co[oE + 3] = e3;
        co[oXN + 0] = x[0] + dt * vn0; co[oXN + 1] = x[1] + dt * vn1; co[oXN + 2] = x[2] + dt * vn2;
        co[oXN + 3] = qw * e0 + qx * e3 + qy * e2 - qz * e1;
        co[oXN + 4] = qw * e1 + qy * e3 + qz * e0 - qx * e2;
        co[oXN + 5] = qw * e2 + qz * e3 + qx * e1 - qy * e0;
        co[oXN + 6] = qw * e3 - qx * e0 - qy * e1 - qz * e2;
        co[oXN + 7] = vn0; co[oXN + 8] = vn1; co[oXN + 9] = vn2; co[oXN + 10] = wn0; co[oXN + 11] = wn1; co[oXN + 12] = wn2;
        const T beta = (T(0.5) * ch - kap) * inv_n * inv_n;
        const T Ly0 = qw * y0 - qz * y1 + qy * y2, Ly1 = qz * y0 + qw * y1 - qx * y2, Ly2 = -qy * y0 + qx * y1 + qw * y2,
                Ly3 = -qx * y0 - qy * y1 - qz * y2;
        const T m0 = beta * Ly0 - T(0.5) * kap * qx, m1 = beta * Ly1 - T(0.5) * kap * qy, m2 = beta * Ly2 - T(0.5) * kap * qz,
                m3 = beta * Ly3 - T(0.5) * kap * qw;
        co[oQ + 0] = dt * (kap * qw + y0 * m0);  co[oQ + 1] = dt * (-kap * qz + y1 * m0); co[oQ + 2] = dt * (kap * qy + y2 * m0);
        co[oQ + 3] = dt * (kap * qz + y0 * m1);  co[oQ + 4] = dt * (kap * qw + y1 * m1);  co[oQ + 5] = dt * (-kap * qx + y2 * m1);
        co[oQ + 6] = dt * (-kap * qy + y0 * m2); co[oQ + 7] = dt * (kap * qx + y1 * m2);  co[oQ + 8] = dt * (kap * qw + y2 * m2);
        co[oQ + 9] = dt * (-kap * qx + y0 * m3); co[oQ + 10] = dt * (-kap * qy + y1 * m3); co[oQ + 11] = dt * (-kap * qz + y2 * m3);
        const T* qr = P + 21 + 3 * (N + 1) + 4 * k;
        T dm = T(0), dp = T(0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const T a = x[3 + i] - qr[i], bq = x[3 + i] + qr[i];
            dm += a * a; dp += bq * bq;
        }
        co[oSGN] = dm > dp ? T(1) : T(-1);
        co[oS]   = Tz * inv_m;
    }

    // Reference state of node k, entry i (field-major parameter block, quadrotor.example.cpp:104-116).
    template <class T>
    __device__ __forceinline__ static T ref_state(const T* __restrict__ sx, int N, int k, int i) {
        const T* P = sx + Quadrotor::n_dec(N) + 21;
        return i < 3 ? P[3 * k + i] : i < 7 ? P[3 * (N + 1) + 4 * k + (i - 3)] : i < 10 ? P[7 * (N + 1) + 3 * k + (i - 7)]
                                                                                           : P[10 * (N + 1) + 3 * k + (i - 10)];
    }
    template <class T> __device__ __forceinline__ static T state_weight(int) { return T(1); }

    // 8 lanes per node write the non-zero entries of A (the image is pre-zeroed).
    template <class T>
    __device__ __forceinline__ static void fill_A(int c, const T* __restrict__ sx, int N, int k, const T* __restrict__ co,
                                                  T* __restrict__ A) {
        constexpr int NZ = 17;
        const T* P = sx + Quadrotor::n_dec(N);
        const T* x = sx + Quadrotor::x_off(N, k);
        const T dt = P[0];
        const T* Q = co + oQ;
        if (c < 4) {
            // ---- input column u_c ----
            const T u = sx[Quadrotor::u_off(N, k) + c], b = P[18], dd = P[19], inv_m = T(1) / P[1];
            const T tu = T(2) * u, dT = b * tu * inv_m;
            const T G0 = dt * (P[6 + 3 * c] * b * tu) / P[2], G1 = dt * (-P[5 + 3 * c] * b * tu) / P[3],
                    G2 = dt * (((c & 1) ? -dd : dd) * tu) / P[4];
            T* a = A + 13 + c;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                a[r * NZ]       = -dt * dt * dT * co[oC + r];
                a[(7 + r) * NZ] = -dt * dT * co[oC + r];
            }
            a[10 * NZ] = -G0; a[11 * NZ] = -G1; a[12 * NZ] = -G2;
#pragma unroll
            for (int r = 0; r < 4; ++r) a[(3 + r) * NZ] = -(Q[3 * r] * G0 + Q[3 * r + 1] * G1 + Q[3 * r + 2] * G2);
            // ---- orientation column q_c:  d a / d q = s dc/dq;  d q+ / d q = Rmat(e) ----
            const T qx = x[3], qy = x[4], qz = x[5], qw = x[6], s = co[oS];
            // dc/d(qx,qy,qz,qw): row0 (2qz, 2qw, 2qx, 2qy), row1 (-2qw, 2qz, 2qy, -2qx), row2 (-4qx, -4qy, 0, 0)
            const T d0 = T(2) * (c == 0 ? qz : c == 1 ? qw : c == 2 ? qx : qy);
            const T d1 = T(2) * (c == 0 ? -qw : c == 1 ? qz : c == 2 ? qy : -qx);
            const T d2 = c == 0 ? -T(4) * qx : -T(4) * qy;
            T* q = A + 3 + c;
            q[0 * NZ] = -dt * dt * s * d0; q[7 * NZ] = -dt * s * d0;
            q[1 * NZ] = -dt * dt * s * d1; q[8 * NZ] = -dt * s * d1;
            if (c < 2) { q[2 * NZ] = -dt * dt * s * d2; q[9 * NZ] = -dt * s * d2; }  // a_z independent of q.z, q.w
            const T e0 = co[oE], e1 = co[oE + 1], e2 = co[oE + 2], e3 = co[oE + 3];
            q[3 * NZ] = -(c == 0 ? e3 : c == 1 ? e2 : c == 2 ? -e1 : e0);
            q[4 * NZ] = -(c == 0 ? -e2 : c == 1 ? e3 : c == 2 ? e0 : e1);
            q[5 * NZ] = -(c == 0 ? e1 : c == 1 ? -e0 : c == 2 ? e3 : e2);
            q[6 * NZ] = -(c == 0 ? -e0 : c == 1 ? -e1 : c == 2 ? -e2 : e3);
        } else if (c < 7) {
            // ---- angular-velocity column w_j: e_j + dt I^-1 (Iw x e_j - I_j (w x e_j)); constant p / v entries of row j ----
            const int j = c - 4, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            const T Ij = P[2 + j], I1v = P[2 + j1], I2v = P[2 + j2];
            const T w1v = x[10 + j1], w2v = x[10 + j2];
            T G[3];
            G[j]  = T(1);
            G[j1] = dt / I1v * (I2v * w2v - Ij * w2v);   // (Iw x e_j)[j+1] = +Iw[j+2],  (w x e_j)[j+1] = +w[j+2]
            G[j2] = dt / I2v * (-I1v * w1v + Ij * w1v);  // (Iw x e_j)[j+2] = -Iw[j+1],  (w x e_j)[j+2] = -w[j+1]
            T* a = A + 10 + j;
            a[10 * NZ] = -G[0]; a[11 * NZ] = -G[1]; a[12 * NZ] = -G[2];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[(3 + r) * NZ] = -(Q[3 * r] * G[0] + Q[3 * r + 1] * G[1] + Q[3 * r + 2] * G[2]);
            A[j * NZ + j] = -T(1); A[j * NZ + 7 + j] = -dt; A[(7 + j) * NZ + 7 + j] = -T(1);
        }
    }

    // Input entries: objective 1e-6 u^2 + 1e-6 (u - u_prev)^2 [+ the (k+1) term], inequalities r - rmax, -r (lanes c < 4).
    template <class T, bool BARRIER>
    __device__ __forceinline__ static void fill_inputs(int c, const T* __restrict__ sx, int N, int k, const BarrierCoef<T>& bar,
                                                       T* __restrict__ rec, const RecLayout& L, T* __restrict__ Hk, T& cost,
                                                       T& bsum, T& hmax) {
        if (c >= 4) return;
        constexpr int NZ = 17;
        const T u = sx[Quadrotor::u_off(N, k) + c], rmax = sx[Quadrotor::n_dec(N) + 20];
        T g = T(2e-6) * u, hd = T(2e-6) + (BARRIER ? T(1e-6) : T(0));
        cost += T(1e-6) * u * u;
        if (k) {
            const T e = u - sx[Quadrotor::u_off(N, k - 1) + c];
            cost += T(1e-6) * e * e; g += T(2e-6) * e; hd += T(2e-6);
        }
        if (k + 1 < N) {
            const T e = sx[Quadrotor::u_off(N, k + 1) + c] - u;
            g -= T(2e-6) * e; hd += T(2e-6);
            rec[L.Hc + k * 4 + c] = -T(2e-6);
        }
        const T h0 = u - rmax, h1 = -u;
        T b0 = T(0), dz0 = T(0), d20 = T(0), b1 = T(0), dz1 = T(0), d21 = T(0);
        if (BARRIER) {
            barrier_eval(bar, h0, &b0, &dz0, &d20);
            barrier_eval(bar, h1, &b1, &dz1, &d21);
        }
        bsum += b0 + b1;
        hmax = fmax(hmax, fmax(h0, h1));
        g += dz0 - dz1;
        hd += d20 + d21;
        rec[L.h + 8 * k + 2 * c]     = h0;
        rec[L.h + 8 * k + 2 * c + 1] = h1;
        rec[L.grad + Quadrotor::u_off(N, k) + c] = g;
        Hk[tri_index(NZ, 13 + c, 13 + c)] = hd;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// RC car
// ---------------------------------------------------------------------------------------------------------------------
template <>
struct SmallPolicy<RcCar> {
    static constexpr int CORE = 25;
    // core layout: dvx[5] | dvy[5] | dom[5] | xn[6] | sp | cp
    static constexpr int oDX = 0, oDY = 5, oDW = 10, oXN = 15, oSP = 21, oCP = 22;

    template <class T>
    __device__ __forceinline__ static void core(const T* __restrict__ sx, int N, int k, T* __restrict__ co) {
        using D5 = VDual<T, 5>;
        const T* x = sx + RcCar::x_off(N, k);
        const T* u = sx + RcCar::u_off(N, k);
        const T dt = sx[RcCar::n_dec(N)];
        D5 in[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) { in[i] = D5(i < 3 ? x[3 + i] : u[i - 3]); in[i].d[i] = T(1); }
        D5 acc[3];
        RcCar::accelerations(sx + RcCar::n_dec(N), in[0], in[1], in[2], in[3], in[4], acc);
        const T vxn = x[3] + dt * acc[0].v, vyn = x[4] + dt * acc[1].v, omn = x[5] + dt * acc[2].v;
        T sp, cp;
        m_sincos(x[2], &sp, &cp);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            co[oDX + j] = (j == 0 ? T(1) : T(0)) + dt * acc[0].d[j];
            co[oDY + j] = (j == 1 ? T(1) : T(0)) + dt * acc[1].d[j];
            co[oDW + j] = (j == 2 ? T(1) : T(0)) + dt * acc[2].d[j];
        }
        co[oXN + 0] = x[0] + dt * (vxn * cp - vyn * sp);
        co[oXN + 1] = x[1] + dt * (vxn * sp + vyn * cp);
        co[oXN + 2] = x[2] + dt * omn;
        co[oXN + 3] = vxn; co[oXN + 4] = vyn; co[oXN + 5] = omn;
        co[oSP] = sp; co[oCP] = cp;
    }
    template <class T>
    __device__ __forceinline__ static T ref_state(const T* __restrict__ sx, int N, int k, int i) {
        return sx[RcCar::n_dec(N) + 15 + 2 * k + i];  // only the position (i < 2) is tracked
    }
    template <class T> __device__ __forceinline__ static T state_weight(int i) { return i < 2 ? T(1) : T(0); }

    template <class T>
    __device__ __forceinline__ static void fill_A(int c, const T* __restrict__ sx, int N, int, const T* __restrict__ co,
                                                  T* __restrict__ A) {
        constexpr int NZ = 8;
        const T dt = sx[RcCar::n_dec(N)], sp = co[oSP], cp = co[oCP];
        if (c < 2) {
            A[c * NZ + c] = -T(1);
        } else if (c == 2) {
            const T vxn = co[oXN + 3], vyn = co[oXN + 4];
            A[0 * NZ + 2] = dt * (vxn * sp + vyn * cp);
            A[1 * NZ + 2] = -dt * (vxn * cp - vyn * sp);
            A[2 * NZ + 2] = -T(1);
        } else {
            const int j = c - 3;
            const T dvx = co[oDX + j], dvy = co[oDY + j], dom = co[oDW + j];
            A[0 * NZ + c] = -dt * (cp * dvx - sp * dvy);
            A[1 * NZ + c] = -dt * (sp * dvx + cp * dvy);
            A[2 * NZ + c] = -dt * dom;
            A[3 * NZ + c] = -dvx;
            if (j != 3) A[4 * NZ + c] = -dvy;  // vy+ does not depend on the duty cycle d
            A[5 * NZ + c] = -dom;
        }
    }

    // Input entries (lanes 6, 7 = d, delta): objective, |u| - 15;  lane 3 (v_x): 0.3 - v_x.
    template <class T, bool BARRIER>
    __device__ __forceinline__ static void fill_inputs(int c, const T* __restrict__ sx, int N, int k, const BarrierCoef<T>& bar,
                                                       T* __restrict__ rec, const RecLayout& L, T* __restrict__ Hk, T& cost,
                                                       T& bsum, T& hmax) {
        constexpr int NZ = 8;
        if (c >= 6) {
            const int i = c - 6;
            const T u = sx[RcCar::u_off(N, k) + i];
            T g = T(2e-6) * u, hd = T(2e-6) + (BARRIER ? T(1e-6) : T(0));
            cost += T(1e-6) * u * u;
            if (k) {
                const T e = u - sx[RcCar::u_off(N, k - 1) + i];
                cost += T(1e-6) * e * e; g += T(2e-6) * e; hd += T(2e-6);
            }
            if (k + 1 < N) {
                const T e = sx[RcCar::u_off(N, k + 1) + i] - u;
                g -= T(2e-6) * e; hd += T(2e-6);
                rec[L.Hc + k * 2 + i] = -T(2e-6);
            }
            const T h = m_abs(u) - T(15), sg = T((u > T(0)) - (u < T(0)));
            T b0 = T(0), dz = T(0), d2 = T(0);
            if (BARRIER) barrier_eval(bar, h, &b0, &dz, &d2);
            bsum += b0;
            hmax = fmax(hmax, h);
            rec[L.h + 3 * k + i] = h;
            rec[L.grad + RcCar::u_off(N, k) + i] = g + dz * sg;
            Hk[tri_index(NZ, c, c)] = hd + d2 * sg * sg;
        } else if (c == 3) {  // 0.3 - v_x: adds to the v_x entries written by the state pass (done there via `extra`)
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------------
template <class Mdl, class T>
struct SmallShape {
    static constexpr int NX = Mdl::NX, NU = Mdl::NU, NZ = Mdl::NZ, NH = Mdl::NH, TRI = NZ * (NZ + 1) / 2, G = 4;
    static constexpr int CORE = SmallPolicy<Mdl>::CORE;
    static constexpr int r4(int x) { return (x + 3) & ~3; }
    struct Offsets {
        int A, H, core, X, total;  // elements, per warp
    };
    static Offsets offsets(int N, int n_xp) {
        Offsets o;
        o.A = 0;
        o.H = r4(G * NX * NZ);
        o.core = r4(o.H + G * TRI);
        o.X = r4(o.core + N * CORE);
        o.total = r4(o.X + n_xp) + 4;
        return o;
    }
};

// JAC: the "Jacobian sweep" alone (BASELINE config 2): only the equality residuals g and the dynamics Jacobian blocks A are computed
// and stored; the objective / inequality / Gauss-Newton work and the H image are skipped.
template <class Mdl, class T, bool BARRIER, int WARPS, bool JAC = false>
__global__ void __launch_bounds__(WARPS * 32)
small_team_kernel(const T* __restrict__ xp_all, long long ld_xp, T* __restrict__ rec_all, long long ld_rec,
                  T* __restrict__ partials, int N, int n_xp, long long batch, RecLayout L, BarrierCoef<T> bar,
                  typename SmallShape<Mdl, T>::Offsets O, unsigned int* __restrict__ sched, const int* __restrict__ active = nullptr,
                  const unsigned int* __restrict__ n_active = nullptr) {
    using Sh = SmallShape<Mdl, T>;
    using Pol = SmallPolicy<Mdl>;
    constexpr int NX = Sh::NX, NU = Sh::NU, NZ = Sh::NZ, TRI = Sh::TRI, G = Sh::G, NA = NX * NZ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    T* const wsm  = reinterpret_cast<T*>(smem_raw) + (long long)wib * O.total;
    T* const stA  = wsm + O.A;
    T* const stH  = wsm + O.H;
    T* const core = wsm + O.core;
    T* const sx   = wsm + O.X;
    for (int e = lane; e < O.core; e += 32) wsm[e] = T(0);  // staging image: zeros persist
    __syncwarp();
    const int slot = lane >> 3, c = lane & 7;  // node slot within the group of 4, column role
    T* const sA = stA + slot * NA;
    T* const sH = stH + slot * TRI;
    const long long warp_id = (long long)blockIdx.x * WARPS + wib, n_warps = (long long)gridDim.x * WARPS;
    bool pending = false;

    // Trajectories are claimed from an atomic counter (sched[0]); the first one is the warp's index.  A warp that becomes resident
    // late then claims fewer trajectories instead of stretching the launch (same scheme as the quadruped sweep).
    // `active` / `n_active` (SQP loop): the work items are the trajectories still RUNNING, listed by build_active_kernel.
    const long long limit = n_active ? (long long)*n_active : batch;
    long long item = warp_id;
    while (item < limit) {
        const long long b = active ? (long long)active[item] : item;
        const T* __restrict__ x = xp_all + b * ld_xp;
        T* __restrict__ r       = rec_all + b * ld_rec;
        // ---- phase 0 ------------------------------------------------------------------------------------------------------
        for (int e = lane; e < n_xp; e += 32) {
            if (sizeof(T) == 4)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sx + e)), "l"(x + e) : "memory");
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sx + e)), "l"(x + e) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // ---- phase A: thread-per-node cores ---------------------------------------------------------------------------------
        for (int k = lane; k < N; k += 32) Pol::core(sx, N, k, core + k * Sh::CORE);
        __syncwarp();

        T cost = T(0), bsum = T(0), gmax = T(0), hmax = -INFINITY;
        if (lane < NX) {  // x_0 - x_measured
            const T gv = sx[lane] - sx[Mdl::xm_off(N) + lane];
            r[L.g + lane] = gv;
            gmax = m_abs(gv);
        }
        // ---- phase B: 8 lanes per node, 4 nodes per step ----------------------------------------------------------------------
        const int groups = (N + G - 1) / G;
        for (int g = 0; g < groups; ++g) {
            const int k = g * G + slot;
            const bool active = k < N;
            if (pending) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                pending = false;
            }
            if (active) {
                const T* co = core + k * Sh::CORE;
                Pol::template fill_A<T>(c, sx, N, k, co, sA);
                // state entries i = c, c + 8: defect, objective gradient / diagonal
                for (int i = c; i < NX; i += 8) {
                    const T gv = sx[Mdl::x_off(N, k + 1) + i] - co[Pol::oXN + i];
                    r[L.g + NX + NX * k + i] = gv;
                    gmax = fmax(gmax, m_abs(gv));
                    if constexpr (JAC) continue;
                    const T wgt = Pol::template state_weight<T>(i);
                    T gq = T(0), hd = BARRIER ? T(1e-6) : T(0);
                    if (wgt != T(0)) {
                        const T ref = Pol::template ref_state<T>(sx, N, k, i);
                        const bool isq = Mdl::KIND == 0 && i >= 3 && i < 7;
                        const T res = isq ? sx[Mdl::x_off(N, k) + i] + co[Mdl::KIND == 0 ? 32 : 0] * ref : sx[Mdl::x_off(N, k) + i] - ref;
                        cost += res * res;
                        gq = T(2) * res;
                        hd += T(2);
                    }
                    if (Mdl::KIND == 1 && i == 3) {  // RC car: 0.3 - v_x  (rc_car.example.cpp:281)
                        const T h = T(0.3) - sx[Mdl::x_off(N, k) + 3];
                        T b0 = T(0), dz = T(0), d2 = T(0);
                        if (BARRIER) barrier_eval(bar, h, &b0, &dz, &d2);
                        bsum += b0;
                        hmax = fmax(hmax, h);
                        r[L.h + 3 * k + 2] = h;
                        gq -= dz;
                        hd += d2;
                    }
                    r[L.grad + Mdl::x_off(N, k) + i] = gq;
                    sH[tri_index(NZ, i, i)] = hd;
                }
                if constexpr (!JAC) Pol::template fill_inputs<T, BARRIER>(c, sx, N, k, bar, r, L, sH, cost, bsum, hmax);
            }
            // ---- hand the 4-node image to the TMA engine ----------------------------------------------------------------------
            const int k0 = g * G, cnt = min(G, N - k0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (cnt == G) {
                if (lane == 0) {
                    tpn_bulk_store(r + L.A + (long long)k0 * NA, stA, G * NA * sizeof(T));
                    if constexpr (!JAC) tpn_bulk_store(r + L.H + (long long)k0 * TRI, stH, G * TRI * sizeof(T));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                pending = true;
            } else {  // ragged tail of the horizon: plain coalesced stores
                for (int e = lane; e < cnt * NA; e += 32) r[L.A + (long long)k0 * NA + e] = stA[e];
                if constexpr (!JAC)
                    for (int e = lane; e < cnt * TRI; e += 32) r[L.H + (long long)k0 * TRI + e] = stH[e];
                __syncwarp();
            }
        }
        // ---- terminal state x_N ---------------------------------------------------------------------------------------------------
        if (!JAC && lane < NX) {
            const int i = lane;
            const T wgt = Pol::template state_weight<T>(i);
            T gq = T(0), hd = BARRIER ? T(1e-6) : T(0);
            if (wgt != T(0)) {
                const T ref = Pol::template ref_state<T>(sx, N, N, i);
                T res = sx[Mdl::x_off(N, N) + i] - ref;
                if (Mdl::KIND == 0 && i >= 3 && i < 7) {
                    T dm = T(0), dp = T(0);
                    for (int j = 3; j < 7; ++j) {
                        const T rj = Pol::template ref_state<T>(sx, N, N, j), xj = sx[Mdl::x_off(N, N) + j];
                        dm += (xj - rj) * (xj - rj); dp += (xj + rj) * (xj + rj);
                    }
                    if (dm > dp) res = sx[Mdl::x_off(N, N) + i] + ref;
                }
                cost += res * res;
                gq = T(2) * res;
                hd += T(2);
            }
            r[L.grad + Mdl::x_off(N, N) + i] = gq;
            for (int j = i; j < NX; ++j) r[L.HN + tri_index(NX, i, j)] = j == i ? hd : T(0);
        }
        // ---- per-trajectory partials (one entry) ---------------------------------------------------------------------------------
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            cost += __shfl_xor_sync(0xffffffffu, cost, o);
            bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
            gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
            hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
        }
        unsigned int claimed = 0;
        if (lane == 0) {
            T* pt = partials + b * 4;
            pt[0] = cost; pt[1] = bsum; pt[2] = gmax; pt[3] = hmax;
            claimed = atomicAdd(&sched[0], 1u);
        }
        item = n_warps + (long long)__shfl_sync(0xffffffffu, claimed, 0);
    }
    if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __threadfence();
        if (atomicAdd(&sched[1], 1u) == (unsigned int)(n_warps - 1)) {  // the last warp to leave re-arms the counters
            sched[0] = 0u;
            sched[1] = 0u;
            __threadfence();
        }
    }
}

}  // namespace ub
