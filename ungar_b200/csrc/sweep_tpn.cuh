// KKT stage sweep, thread-per-node version for the small models (quadrotor: nz = 17, RC car: nz = 8).
//
// The generic sweep (sweep.cuh) spends one thread per (node, tangent) and therefore recomputes the primal pass nz times;
// on the fp32 configurations that made it issue-bound at ~0.2 of the HBM roofline (1063 / 355 instructions per node).
// Here ONE THREAD OWNS ONE SHOOTING NODE and carries all nz tangents in registers (VDual<T, NZ>): the primal value of every
// operation is computed once, the nz tangent updates are independent FMAs (ILP instead of redundancy), and a warp covers
// 32 nodes per instruction.  A CTA owns one trajectory at a time (persistent over trajectories):
//   inputs   the whole flat Ungar vector of the trajectory (2.5 - 3.8 KB) is copied to shared memory with coalesced loads;
//            the model functors then index that copy exactly as they would index global memory
//   staging  every block of the trajectory's record is assembled in shared memory with odd (bank-conflict-free) per-node
//            strides; the packed Hessian image is zeroed once and only its non-zero slots are rewritten
//   stores   the CTA streams the staged record out in record order with fully coalesced stores
#pragma once

#include "sweep.cuh"

namespace ub {

// Forward dual with N tangents held in registers.
template <class T, int N>
struct VDual {
    T v;
    T d[N];
    UB_HD VDual() : v(T(0)) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = T(0);
    }
    UB_HD VDual(T value) : v(value) {  // NOLINT
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = T(0);
    }
};
template <class T, int N> struct real_of<VDual<T, N>> { using type = T; };

#define UB_VD_LOOP _Pragma("unroll") for (int i_ = 0; i_ < N; ++i_)
template <class T, int N> UB_HD VDual<T, N> operator+(const VDual<T, N>& a, const VDual<T, N>& b) { VDual<T, N> r; r.v = a.v + b.v; UB_VD_LOOP r.d[i_] = a.d[i_] + b.d[i_]; return r; }
template <class T, int N> UB_HD VDual<T, N> operator-(const VDual<T, N>& a, const VDual<T, N>& b) { VDual<T, N> r; r.v = a.v - b.v; UB_VD_LOOP r.d[i_] = a.d[i_] - b.d[i_]; return r; }
template <class T, int N> UB_HD VDual<T, N> operator*(const VDual<T, N>& a, const VDual<T, N>& b) { VDual<T, N> r; r.v = a.v * b.v; UB_VD_LOOP r.d[i_] = a.d[i_] * b.v + a.v * b.d[i_]; return r; }
template <class T, int N> UB_HD VDual<T, N> operator/(const VDual<T, N>& a, const VDual<T, N>& b) { VDual<T, N> r; const T inv = T(1) / b.v; r.v = a.v * inv; UB_VD_LOOP r.d[i_] = (a.d[i_] - r.v * b.d[i_]) * inv; return r; }
template <class T, int N> UB_HD VDual<T, N> operator-(const VDual<T, N>& a) { VDual<T, N> r; r.v = -a.v; UB_VD_LOOP r.d[i_] = -a.d[i_]; return r; }
template <class T, int N> UB_HD VDual<T, N> operator+(const VDual<T, N>& a, T c) { VDual<T, N> r = a; r.v += c; return r; }
template <class T, int N> UB_HD VDual<T, N> operator+(T c, const VDual<T, N>& a) { VDual<T, N> r = a; r.v += c; return r; }
template <class T, int N> UB_HD VDual<T, N> operator-(const VDual<T, N>& a, T c) { VDual<T, N> r = a; r.v -= c; return r; }
template <class T, int N> UB_HD VDual<T, N> operator-(T c, const VDual<T, N>& a) { VDual<T, N> r; r.v = c - a.v; UB_VD_LOOP r.d[i_] = -a.d[i_]; return r; }
template <class T, int N> UB_HD VDual<T, N> operator*(const VDual<T, N>& a, T c) { VDual<T, N> r; r.v = a.v * c; UB_VD_LOOP r.d[i_] = a.d[i_] * c; return r; }
template <class T, int N> UB_HD VDual<T, N> operator*(T c, const VDual<T, N>& a) { return a * c; }
template <class T, int N> UB_HD VDual<T, N> operator/(const VDual<T, N>& a, T c) { return a * (T(1) / c); }
template <class T, int N> UB_HD VDual<T, N> operator/(T c, const VDual<T, N>& a) { VDual<T, N> r; const T inv = T(1) / a.v; r.v = c * inv; const T k = -r.v * inv; UB_VD_LOOP r.d[i_] = k * a.d[i_]; return r; }
template <class T, int N> UB_HD VDual<T, N>& operator+=(VDual<T, N>& a, const VDual<T, N>& b) { a.v += b.v; UB_VD_LOOP a.d[i_] += b.d[i_]; return a; }
template <class T, int N> UB_HD VDual<T, N>& operator-=(VDual<T, N>& a, const VDual<T, N>& b) { a.v -= b.v; UB_VD_LOOP a.d[i_] -= b.d[i_]; return a; }
template <class T, int N> UB_HD VDual<T, N> vd_chain(T f, T df, const VDual<T, N>& x) { VDual<T, N> r; r.v = f; UB_VD_LOOP r.d[i_] = df * x.d[i_]; return r; }
template <class T, int N> UB_HD T val(const VDual<T, N>& x) { return x.v; }
template <class T, int N> UB_HD VDual<T, N> m_sqrt(const VDual<T, N>& x) { const T r = m_sqrt(x.v); return vd_chain(r, T(0.5) / r, x); }
template <class T, int N> UB_HD VDual<T, N> m_atan(const VDual<T, N>& x) { return vd_chain(m_atan(x.v), T(1) / (T(1) + x.v * x.v), x); }
template <class T, int N> UB_HD VDual<T, N> m_abs(const VDual<T, N>& x) { return vd_chain(m_abs(x.v), T((x.v > T(0)) - (x.v < T(0))), x); }
template <class T, int N> UB_HD void m_sincos(const VDual<T, N>& x, VDual<T, N>* s, VDual<T, N>* c) {
    T sv, cv;
    m_sincos(x.v, &sv, &cv);
    *s = vd_chain(sv, cv, x);
    *c = vd_chain(cv, -sv, x);
}
template <class T, int N> UB_HD VDual<T, N> m_sin(const VDual<T, N>& x) { T s, c; m_sincos(x.v, &s, &c); return vd_chain(s, c, x); }
template <class T, int N> UB_HD VDual<T, N> m_cos(const VDual<T, N>& x) { T s, c; m_sincos(x.v, &s, &c); return vd_chain(c, -s, x); }
template <class T, int N> UB_HD VDual<T, N> select(bool c, const VDual<T, N>& t, const VDual<T, N>& f) { return c ? t : f; }
#undef UB_VD_LOOP


// ---------------------------------------------------------------------------------------------------------------------
// Affine function of ONE local variable, a * z_idx + b, as (value, slope, idx).  All objective residuals and, for the
// quadrotor and the RC car, all inequality rows are of this form (quadrotor.example.cpp:203-231, :280-288;
// rc_car.example.cpp:204-224, :271-282), so their derivatives need one slope, not nz tangents.
// ---------------------------------------------------------------------------------------------------------------------
template <class T>
struct OneVar {
    T v, d;
    int idx;
    UB_HD OneVar() : v(T(0)), d(T(0)), idx(0) {}
    UB_HD OneVar(T value) : v(value), d(T(0)), idx(0) {}  // NOLINT
    UB_HD OneVar(T value, T slope, int index) : v(value), d(slope), idx(index) {}
};
template <class T> struct real_of<OneVar<T>> { using type = T; };
template <class T> UB_HD OneVar<T> operator+(const OneVar<T>& a, T c) { return {a.v + c, a.d, a.idx}; }
template <class T> UB_HD OneVar<T> operator+(T c, const OneVar<T>& a) { return {c + a.v, a.d, a.idx}; }
template <class T> UB_HD OneVar<T> operator-(const OneVar<T>& a, T c) { return {a.v - c, a.d, a.idx}; }
template <class T> UB_HD OneVar<T> operator-(T c, const OneVar<T>& a) { return {c - a.v, -a.d, a.idx}; }
template <class T> UB_HD OneVar<T> operator*(const OneVar<T>& a, T c) { return {a.v * c, a.d * c, a.idx}; }
template <class T> UB_HD OneVar<T> operator*(T c, const OneVar<T>& a) { return {c * a.v, c * a.d, a.idx}; }
template <class T> UB_HD OneVar<T> operator-(const OneVar<T>& a) { return {-a.v, -a.d, a.idx}; }
template <class T> UB_HD T val(const OneVar<T>& x) { return x.v; }
template <class T> UB_HD OneVar<T> m_abs(const OneVar<T>& x) {
    const T s = T((x.v > T(0)) - (x.v < T(0)));
    return {m_abs(x.v), s * x.d, x.idx};
}

// ---------------------------------------------------------------------------------------------------------------------
// Hand-structured node Jacobians: A = d(x_{k+1} - f)/dz written as its structurally non-zero entries only (the shared-
// memory image of A is zeroed once), plus f(x_k, u_k).  `structured` = false -> the kernel uses the generic VDual pass.
// ---------------------------------------------------------------------------------------------------------------------
template <class Mdl>
struct NodeJacobian {
    static constexpr bool structured = false;
};

// Quadrotor (quadrotor.example.cpp:126-190): chain rule by hand through the thrust T e_z, the body moments, w+ = w + dt
// I^-1 tau and Qw = d q+ / d w+ (same derivation as sweep_structured.cuh).
template <>
struct NodeJacobian<Quadrotor> {
    static constexpr bool structured = true;
    template <class T>
    __device__ __forceinline__ static void run(const T* __restrict__ sx, int N, int k, T* __restrict__ A, T* __restrict__ xn) {
        constexpr int NZ = 17;
        const T* P = sx + Quadrotor::n_dec(N);
        const T* x = sx + Quadrotor::x_off(N, k);
        const T* u = sx + Quadrotor::u_off(N, k);
        const T dt = P[0], m = P[1], I0 = P[2], I1 = P[3], I2 = P[4], g0 = P[17], b = P[18], dd = P[19];
        const T iI0 = T(1) / I0, iI1 = T(1) / I1, iI2 = T(1) / I2, inv_m = T(1) / m;
        const T qx = x[3], qy = x[4], qz = x[5], qw = x[6], w0 = x[10], w1 = x[11], w2 = x[12];
        T Tz = T(0), mo0 = T(0), mo1 = T(0), mo2 = T(0);
        T dTz[4], dm0[4], dm1[4], dm2[4];  // d/du_i of thrust and body moments
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const T r2 = u[i] * u[i], sgn = (i & 1) ? -dd : dd, px = P[5 + 3 * i], py = P[6 + 3 * i];
            Tz += b * r2; mo0 += py * (b * r2); mo1 -= px * (b * r2); mo2 += sgn * r2;
            const T tu = T(2) * u[i];
            dTz[i] = b * tu; dm0[i] = py * b * tu; dm1[i] = -px * b * tu; dm2[i] = sgn * tu;
        }
        // third column of R(q): q * (0, 0, T) = T c
        const T c0 = T(2) * (qw * qy + qz * qx), c1 = T(2) * (qz * qy - qw * qx), c2 = T(1) - T(2) * (qx * qx + qy * qy);
        const T a0 = Tz * c0 * inv_m, a1 = Tz * c1 * inv_m, a2 = (Tz * c2 - m * g0) * inv_m;
        const T Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I2 * w2;
        const T t0 = mo0 - (w1 * Iw2 - w2 * Iw1), t1 = mo1 - (w2 * Iw0 - w0 * Iw2), t2 = mo2 - (w0 * Iw1 - w1 * Iw0);
        const T vn0 = x[7] + dt * a0, vn1 = x[8] + dt * a1, vn2 = x[9] + dt * a2;
        const T wn0 = w0 + dt * (t0 * iI0), wn1 = w1 + dt * (t1 * iI1), wn2 = w2 + dt * (t2 * iI2);
        const T y0 = dt * wn0, y1 = dt * wn1, y2 = dt * wn2;
        const T nn = m_sqrt(y0 * y0 + y1 * y1 + y2 * y2 + T(UB_EPS));
        T sh, ch;
        m_sincos(T(0.5) * nn, &sh, &ch);
        const T inv_n = T(1) / nn, kap = sh * inv_n;
        const T e0 = y0 * kap, e1 = y1 * kap, e2 = y2 * kap, e3 = ch;
        xn[0] = x[0] + dt * vn0; xn[1] = x[1] + dt * vn1; xn[2] = x[2] + dt * vn2;
        xn[3] = qw * e0 + qx * e3 + qy * e2 - qz * e1;
        xn[4] = qw * e1 + qy * e3 + qz * e0 - qx * e2;
        xn[5] = qw * e2 + qz * e3 + qx * e1 - qy * e0;
        xn[6] = qw * e3 - qx * e0 - qy * e1 - qz * e2;
        xn[7] = vn0; xn[8] = vn1; xn[9] = vn2; xn[10] = wn0; xn[11] = wn1; xn[12] = wn2;
        // Qw = dt Lmat(q) E
        const T beta = (T(0.5) * ch - kap) * inv_n * inv_n;
        const T Ly0 = qw * y0 - qz * y1 + qy * y2, Ly1 = qz * y0 + qw * y1 - qx * y2, Ly2 = -qy * y0 + qx * y1 + qw * y2,
                Ly3 = -qx * y0 - qy * y1 - qz * y2;
        const T m0 = beta * Ly0 - T(0.5) * kap * qx, m1 = beta * Ly1 - T(0.5) * kap * qy, m2 = beta * Ly2 - T(0.5) * kap * qz,
                m3 = beta * Ly3 - T(0.5) * kap * qw;
        const T Q[12] = {dt * (kap * qw + y0 * m0),  dt * (-kap * qz + y1 * m0), dt * (kap * qy + y2 * m0),
                         dt * (kap * qz + y0 * m1),  dt * (kap * qw + y1 * m1),  dt * (-kap * qx + y2 * m1),
                         dt * (-kap * qy + y0 * m2), dt * (kap * qx + y1 * m2),  dt * (kap * qw + y2 * m2),
                         dt * (-kap * qx + y0 * m3), dt * (-kap * qy + y1 * m3), dt * (-kap * qz + y2 * m3)};
        // ---- rows p+ (0-2) and v+ (7-9): d a / d q = (T/m) dc/dq, d a / d u_i = (dT_i / m) c ------------------------------
        const T s = Tz * inv_m, dts = dt * s, dt2s = dt * dts;
        // dc/d(qx, qy, qz, qw):  row 0: (2qz, 2qw, 2qx, 2qy);  row 1: (-2qw, 2qz, 2qy, -2qx);  row 2: (-4qx, -4qy, 0, 0)
        const T dc[3][4] = {{T(2) * qz, T(2) * qw, T(2) * qx, T(2) * qy}, {-T(2) * qw, T(2) * qz, T(2) * qy, -T(2) * qx},
                            {-T(4) * qx, -T(4) * qy, T(0), T(0)}};
        const T cc[3] = {c0, c1, c2};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            A[r * NZ + r]           = -T(1);   // d p+ / d p
            A[r * NZ + 7 + r]       = -dt;     // d p+ / d v
            A[(7 + r) * NZ + 7 + r] = -T(1);   // d v+ / d v
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (r < 2 || c < 2) {  // a_z does not depend on q.z, q.w
                    A[r * NZ + 3 + c]       = -dt2s * dc[r][c];
                    A[(7 + r) * NZ + 3 + c] = -dts * dc[r][c];
                }
                const T au = dTz[c] * inv_m * cc[r];
                A[r * NZ + 13 + c]       = -dt * dt * au;
                A[(7 + r) * NZ + 13 + c] = -dt * au;
            }
        }
        // ---- rows w+ (10-12) and q+ (3-6) ----------------------------------------------------------------------------------------
        const T Iv[3] = {I0, I1, I2}, iI[3] = {iI0, iI1, iI2}, wv[3] = {w0, w1, w2}, Iw[3] = {Iw0, Iw1, Iw2};
        const T ev[4] = {e0, e1, e2, e3};
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // w columns: e_c + dt I^-1 (Iw x e_c - I_c (w x e_c))
            const int c1i = (c + 1) % 3, c2i = (c + 2) % 3;
            T G[3];
            G[c]   = T(1);
            G[c1i] = dt * iI[c1i] * (Iw[c2i] - Iv[c] * wv[c2i]);    // (Iw x e_c)[c+1] = +Iw[c+2];  (w x e_c)[c+1] = +w[c+2]
            G[c2i] = dt * iI[c2i] * (-Iw[c1i] + Iv[c] * wv[c1i]);   // (Iw x e_c)[c+2] = -Iw[c+1];  (w x e_c)[c+2] = -w[c+1]
            const int col = 10 + c;
            A[10 * NZ + col] = -G[0]; A[11 * NZ + col] = -G[1]; A[12 * NZ + col] = -G[2];
#pragma unroll
            for (int r = 0; r < 4; ++r) A[(3 + r) * NZ + col] = -(Q[3 * r] * G[0] + Q[3 * r + 1] * G[1] + Q[3 * r + 2] * G[2]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // u columns: dt I^-1 d mom / d u_c;  q columns: Rmat(e)
            const T G0 = dt * iI0 * dm0[c], G1 = dt * iI1 * dm1[c], G2 = dt * iI2 * dm2[c];
            const int col = 13 + c;
            A[10 * NZ + col] = -G0; A[11 * NZ + col] = -G1; A[12 * NZ + col] = -G2;
#pragma unroll
            for (int r = 0; r < 4; ++r) A[(3 + r) * NZ + col] = -(Q[3 * r] * G0 + Q[3 * r + 1] * G1 + Q[3 * r + 2] * G2);
        }
        // Rmat(e): rows x [e3 e2 -e1 e0], y [-e2 e3 e0 e1], z [e1 -e0 e3 e2], w [-e0 -e1 -e2 e3]
        A[3 * NZ + 3] = -ev[3]; A[3 * NZ + 4] = -ev[2]; A[3 * NZ + 5] = ev[1];  A[3 * NZ + 6] = -ev[0];
        A[4 * NZ + 3] = ev[2];  A[4 * NZ + 4] = -ev[3]; A[4 * NZ + 5] = -ev[0]; A[4 * NZ + 6] = -ev[1];
        A[5 * NZ + 3] = -ev[1]; A[5 * NZ + 4] = ev[0];  A[5 * NZ + 5] = -ev[3]; A[5 * NZ + 6] = -ev[2];
        A[6 * NZ + 3] = ev[0];  A[6 * NZ + 4] = ev[1];  A[6 * NZ + 5] = ev[2];  A[6 * NZ + 6] = -ev[3];
    }
};

// RC car (rc_car.example.cpp:131-185): the Pacejka force model depends on (vx, vy, w, d, delta) only, so it is
// differentiated with 5 tangents and the Euler / yaw-rotation step is composed by hand.
template <>
struct NodeJacobian<RcCar> {
    static constexpr bool structured = true;
    template <class T>
    __device__ __forceinline__ static void run(const T* __restrict__ sx, int N, int k, T* __restrict__ A, T* __restrict__ xn) {
        constexpr int NZ = 8;
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
        xn[0] = x[0] + dt * (vxn * cp - vyn * sp);
        xn[1] = x[1] + dt * (vxn * sp + vyn * cp);
        xn[2] = x[2] + dt * omn;
        xn[3] = vxn; xn[4] = vyn; xn[5] = omn;
        A[0 * NZ + 0] = -T(1);
        A[1 * NZ + 1] = -T(1);
        A[0 * NZ + 2] = dt * (vxn * sp + vyn * cp);    // -(d px+/d phi) = -dt (-s vx+ - c vy+)
        A[1 * NZ + 2] = -dt * (vxn * cp - vyn * sp);
        A[2 * NZ + 2] = -T(1);
#pragma unroll
        for (int j = 0; j < 5; ++j) {  // columns vx vy w | d delta  (3..7)
            const T dvx = (j == 0 ? T(1) : T(0)) + dt * acc[0].d[j], dvy = (j == 1 ? T(1) : T(0)) + dt * acc[1].d[j],
                    dom = (j == 2 ? T(1) : T(0)) + dt * acc[2].d[j];
            A[0 * NZ + 3 + j] = -dt * (cp * dvx - sp * dvy);
            A[1 * NZ + 3 + j] = -dt * (sp * dvx + cp * dvy);
            A[2 * NZ + 3 + j] = -dt * dom;
            A[3 * NZ + 3 + j] = -dvx;
            if (j != 3) A[4 * NZ + 3 + j] = -dvy;  // vy+ does not depend on the duty cycle d
            A[5 * NZ + 3 + j] = -dom;
        }
    }
};

template <class Mdl>
struct TpnShape {
    static constexpr int NX = Mdl::NX, NU = Mdl::NU, NZ = Mdl::NZ, NH = Mdl::NH, TRI = NZ * (NZ + 1) / 2;
    static constexpr int odd(int x) { return x | 1; }
    // per-node strides in shared memory (odd -> conflict-free for thread-per-node accesses)
    static constexpr int sA = odd(NX * NZ), sH = odd(TRI), sG = odd(NX), sHv = odd(NH), sQ = odd(NZ), sHc = odd(NU);
    static constexpr int r4(int x) { return (x + 3) & ~3; }
    // array starts (elements) for a horizon of N nodes: every array begins on a multiple of 4 elements (16 B in fp32)
    struct Offsets {
        int A, H, G, Hv, Q, Hc, X, total;
    };
    static Offsets offsets(int N, int n_xp) {
        Offsets o;
        o.A = 0;
        o.H = r4(o.A + N * sA);
        o.G = r4(o.H + N * sH);
        o.Hv = r4(o.G + N * sG);
        o.Q = r4(o.Hv + N * sHv);
        o.Hc = r4(o.Q + N * sQ);
        o.X = r4(o.Hc + N * sHc);
        o.total = r4(o.X + n_xp) + 4;
        return o;
    }
};

__device__ __forceinline__ void tpn_bulk_store(void* gmem, const void* smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem),
                 "r"((unsigned)__cvta_generic_to_shared(smem)), "r"(bytes)
                 : "memory");
}

// Write `nodes` rows of `natural` elements (shared-memory row stride `stride`) to a contiguous global range.
// Unpadded arrays leave through the TMA engine (16-byte aligned prefix; SASS UBLKCP) when `bulk` allows it; padded arrays
// are copied element by element in global order (fully coalesced stores).
template <class T>
__device__ __forceinline__ void tpn_emit(T* __restrict__ g, const T* __restrict__ sm, int nodes, int natural, int stride, bool bulk) {
    const int tid = threadIdx.x;
    if (stride == natural) {
        const int total = nodes * natural;
        int done = 0;
        if (bulk) {
            const unsigned bytes = (unsigned(total) * unsigned(sizeof(T))) & ~15u;
            if (tid == 0 && bytes) tpn_bulk_store(g, sm, bytes);
            done = int(bytes / sizeof(T));
        }
        for (int e = done + tid; e < total; e += blockDim.x) g[e] = sm[e];
    } else {  // padded rows: linear over the global range (the divisions are by compile-time constants after inlining)
        const int total = nodes * natural;
        for (int e = tid; e < total; e += blockDim.x) g[e] = sm[(e / natural) * stride + e % natural];
    }
}

template <class Mdl, class T, bool BARRIER, bool STRUCT>
__global__ void tpn_sweep_kernel(const T* __restrict__ xp_all, long long ld_xp, T* __restrict__ rec_all, long long ld_rec,
                                 T* __restrict__ partials, int N, int n_xp, long long batch, RecLayout L, BarrierCoef<T> bar,
                                 typename TpnShape<Mdl>::Offsets O, int bulk) {
    using Sh = TpnShape<Mdl>;
    constexpr int NX = Sh::NX, NU = Sh::NU, NZ = Sh::NZ, NH = Sh::NH, TRI = Sh::TRI;
    using D = VDual<T, NZ>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* const sm  = reinterpret_cast<T*>(smem_raw);
    T* const pA  = sm + O.A;
    T* const pH  = sm + O.H;
    T* const pG  = sm + O.G;
    T* const pHv = sm + O.Hv;
    T* const pQ  = sm + O.Q;
    T* const pHc = sm + O.Hc;
    T* const sx  = sm + O.X;  // the trajectory's flat vector
    const int tid = threadIdx.x;
    for (int e = tid; e < N * Sh::sH; e += blockDim.x) pH[e] = T(0);  // packed Hessian image: zeros persist
    if (STRUCT)
        for (int e = tid; e < N * Sh::sA; e += blockDim.x) pA[e] = T(0);  // structured A: only non-zeros are rewritten

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const T* __restrict__ x = xp_all + b * ld_xp;
        T* __restrict__ r       = rec_all + b * ld_rec;
        if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();  // previous trajectory fully handed over before sx / staging are overwritten
        for (int e = tid; e < n_xp; e += blockDim.x) sx[e] = x[e];
        __syncthreads();

        for (int k = tid; k <= N; k += blockDim.x) {
            const bool terminal = k == N;
            T cval = T(0), bsum = T(0), gmax = T(0), hmax = -INFINITY;
            T gq[NZ], hd[NZ], hc[NZ];
#pragma unroll
            for (int i = 0; i < NZ; ++i) { gq[i] = T(0); hd[i] = BARRIER ? T(1e-6) : T(0); hc[i] = T(0); }
            T* const hA = pH + (terminal ? 0 : k) * Sh::sH;

            if constexpr (STRUCT && NodeJacobian<Mdl>::structured && Mdl::INEQ_SEPARABLE) {
                // ---- structured path: one-variable affine residuals, hand-derived A ------------------------------------------
                using O = OneVar<T>;
                O z[NZ];
#pragma unroll
                for (int i = 0; i < NX; ++i) z[i] = O(sx[Mdl::x_off(N, k) + i], T(1), i);
                if (!terminal) {
#pragma unroll
                    for (int i = 0; i < NU; ++i) z[NX + i] = O(sx[Mdl::u_off(N, k) + i], T(1), NX + i);
                }
                Mdl::cost_terms(sx, N, k, z, [&](T c, const O& res, bool counts) {
                    if (counts) cval += c * res.v * res.v;
#pragma unroll
                    for (int i = 0; i < NZ; ++i)
                        if (i == res.idx) {
                            gq[i] += T(2) * c * res.v * res.d;
                            hd[i] += T(2) * c * res.d * res.d;
                            if (!counts) hc[i] -= T(2) * c * res.d * res.d;
                        }
                });
                if (!terminal) {
                    T xn[NX];
                    NodeJacobian<Mdl>::run(sx, N, k, pA + k * Sh::sA, xn);
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const T gv = sx[Mdl::x_off(N, k + 1) + i] - xn[i];
                        pG[k * Sh::sG + i] = gv;
                        gmax = fmax(gmax, m_abs(gv));
                    }
                    O h[NH];
                    Mdl::inequalities(sx, N, k, z, h);
#pragma unroll
                    for (int row = 0; row < NH; ++row) {
                        T b0 = T(0), dz = T(0), d2z = T(0);
                        if (BARRIER) barrier_eval(bar, h[row].v, &b0, &dz, &d2z);
                        bsum += b0;
                        hmax = fmax(hmax, h[row].v);
                        pHv[k * Sh::sHv + row] = h[row].v;
#pragma unroll
                        for (int i = 0; i < NZ; ++i)
                            if (i == h[row].idx) {
                                gq[i] += dz * h[row].d;
                                hd[i] += d2z * h[row].d * h[row].d;
                            }
                    }
                }
            } else {
                // ---- generic path: nz tangents in registers ----------------------------------------------------------------------
                D z[NZ];
#pragma unroll
                for (int i = 0; i < NX; ++i) { z[i] = D(sx[Mdl::x_off(N, k) + i]); z[i].d[i] = T(1); }
                if (!terminal) {
#pragma unroll
                    for (int i = 0; i < NU; ++i) { z[NX + i] = D(sx[Mdl::u_off(N, k) + i]); z[NX + i].d[NX + i] = T(1); }
                }
                Mdl::cost_terms(sx, N, k, z, [&](T c, const D& res, bool counts) {
                    if (counts) cval += c * res.v * res.v;
#pragma unroll
                    for (int i = 0; i < NZ; ++i) {
                        gq[i] += T(2) * c * res.v * res.d[i];
                        hd[i] += T(2) * c * res.d[i] * res.d[i];
                        if (!counts) hc[i] -= T(2) * c * res.d[i] * res.d[i];
                    }
                });
                if (!terminal) {
                    {
                        D xn[NX];
                        Mdl::dynamics(sx, N, k, z, xn);
#pragma unroll
                        for (int i = 0; i < NX; ++i) {
                            const T gv = sx[Mdl::x_off(N, k + 1) + i] - xn[i].v;
                            pG[k * Sh::sG + i] = gv;
                            gmax = fmax(gmax, m_abs(gv));
#pragma unroll
                            for (int c = 0; c < NZ; ++c) pA[k * Sh::sA + i * NZ + c] = -xn[i].d[c];
                        }
                    }
                    D h[NH];
                    Mdl::inequalities(sx, N, k, z, h);
                    T d2[NH > 0 ? NH : 1];
#pragma unroll
                    for (int row = 0; row < NH; ++row) {
                        T b0 = T(0), dz = T(0), d2z = T(0);
                        if (BARRIER) barrier_eval(bar, h[row].v, &b0, &dz, &d2z);
                        d2[row] = d2z;
                        bsum += b0;
                        hmax = fmax(hmax, h[row].v);
                        pHv[k * Sh::sHv + row] = h[row].v;
#pragma unroll
                        for (int i = 0; i < NZ; ++i) {
                            gq[i] += dz * h[row].d[i];
                            hd[i] += d2z * h[row].d[i] * h[row].d[i];
                        }
                    }
                    if constexpr (!Mdl::INEQ_SEPARABLE) {  // off-diagonal Gauss-Newton entries
#pragma unroll
                        for (int i = 0; i < NZ; ++i)
#pragma unroll
                            for (int j = i + 1; j < NZ; ++j) {
                                T acc = T(0);
#pragma unroll
                                for (int row = 0; row < NH; ++row) acc += d2[row] * h[row].d[i] * h[row].d[j];
                                hA[tri_index(NZ, i, j)] = acc;
                            }
                    }
                }
            }

            if (terminal) {  // x_N: gradient and diagonal block straight to the record
#pragma unroll
                for (int i = 0; i < NX; ++i) r[L.grad + Mdl::x_off(N, N) + i] = gq[i];
                for (int i = 0; i < NX; ++i)
                    for (int j = i; j < NX; ++j) r[L.HN + tri_index(NX, i, j)] = i == j ? hd[i] : T(0);
            } else {
#pragma unroll
                for (int i = 0; i < NZ; ++i) {
                    hA[tri_index(NZ, i, i)] = hd[i];
                    pQ[k * Sh::sQ + i]      = gq[i];
                }
                if (Mdl::HC) {
#pragma unroll
                    for (int i = 0; i < NU; ++i) pHc[k * Sh::sHc + i] = hc[NX + i];
                }
            }
            T* pt = partials + ((long long)b * (N + 1) + k) * 4;
            pt[0] = cval; pt[1] = bsum; pt[2] = gmax; pt[3] = hmax;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // ---- write-out in record order: the two big blocks through the TMA engine, the small vectors row by row ---------
        tpn_emit(r + L.A, pA, N, NX * NZ, Sh::sA, bulk != 0);
        tpn_emit(r + L.H, pH, N, TRI, Sh::sH, bulk != 0);
        if (bulk && tid == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        tpn_emit(r + L.g + NX, pG, N, NX, Sh::sG, false);
        tpn_emit(r + L.h, pHv, N, NH, Sh::sHv, false);
        if (Mdl::HC) tpn_emit(r + L.Hc, pHc, N - 1, NU, Sh::sHc, false);
        {
            const int lane = tid & 31, warps = blockDim.x >> 5;
            for (int node = tid >> 5; node < N; node += warps)
                for (int i = lane; i < NZ; i += 32) {
                    const T v = pQ[node * Sh::sQ + i];
                    if (i < NX) r[L.grad + node * NX + i] = v;
                    else r[L.grad + NX * (N + 1) + node * NU + (i - NX)] = v;
                }
        }
        for (int e = tid; e < NX; e += blockDim.x) r[L.g + e] = sx[e] - sx[Mdl::xm_off(N) + e];
    }
    if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace ub
