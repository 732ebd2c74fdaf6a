// Per-shooting-node functors of the three Ungar NMPC problems.
//
// The reference writes each problem as three horizon-wide lambdas over one flat vector
// [X | U | parameters] (example/mpc/quadruped.example.cpp:343-363); the body of their `for k` loops is the
// per-node work restated here (SURVEY.md finding 0.3).  Every functor reads the trajectory's flat vector
// `xp` in the reference layout (so the VariableMap indices of the examples stay valid) and the node's
// local variables z = [x_k; u_k] as scalars of type S (plain, Dual or Dep; see dual.cuh).
//
// Contract shared by the models (used by sweep.cuh and by the host-side pattern builder):
//   NX, NU, NZ = NX + NU          sizes of x_k, u_k, z_k
//   NH                            inequality rows per node
//   LEGS                          contact-row groups per node (quadruped: 4, else 0)
//   HC                            1 if the objective couples u_k and u_{k+1} (diagonal block), else 0
//   dynamics(xp, N, k, z, xn)     x_{k+1} = f(x_k, u_k)
//   inequalities(xp, N, k, z, h)  NH rows
//   cost_terms(xp, N, k, z, nz, sink)   the objective as a sum of  c * r^2  with r affine in ONE local
//                                 variable; sink(c, r, counts) — counts = false for the u_{k+1} - u_k term
//                                 whose value belongs to node k + 1 but whose derivatives touch z_k.
#pragma once

#include "dual.cuh"

namespace ub {

// ---------------------------------------------------------------------------------------------
// Lie-group semi-implicit Euler (quadrotor.example.cpp:184-187, quadruped.example.cpp:197-200).
// ---------------------------------------------------------------------------------------------
template <class S>
UB_HD void lie_euler(const S* z, const real_t<S> dt, const Vec3<S>& pdd, const Vec3<S>& wd, S* xn) {
    const Vec3<S> vn{z[7] + dt * pdd.x, z[8] + dt * pdd.y, z[9] + dt * pdd.z};
    const Vec3<S> wn{z[10] + dt * wd.x, z[11] + dt * wd.y, z[12] + dt * wd.z};
    const Quat<S> qn = qmul(Quat<S>{z[3], z[4], z[5], z[6]}, approx_exp(Vec3<S>{dt * wn.x, dt * wn.y, dt * wn.z}));
    xn[0] = z[0] + dt * vn.x; xn[1] = z[1] + dt * vn.y; xn[2] = z[2] + dt * vn.z;
    xn[3] = qn.x; xn[4] = qn.y; xn[5] = qn.z; xn[6] = qn.w;
    xn[7] = vn.x; xn[8] = vn.y; xn[9] = vn.z;
    xn[10] = wn.x; xn[11] = wn.y; xn[12] = wn.z;
}

// State-tracking terms shared by quadrotor and quadruped: |W (p - pRef)|^2 + Min(|q - qRef|^2, |q + qRef|^2)
// + |v - vRef|^2 + |w - wRef|^2 (quadrotor.example.cpp:213-217, quadruped.example.cpp:228-232).
// Utils::Min = CondExpGt(a, b, b, a) (utils/utils.hpp:976): '+' branch only when dm > dp.
template <class S, class T, class Sink>
UB_HD void tracking_terms(const S* z, const T* pref, const T* qref, const T* vref, const T* wref, T w0, T w1, T w2,
                          Sink&& sink) {
    sink(T(1), w0 * (z[0] - pref[0]), true);
    sink(T(1), w1 * (z[1] - pref[1]), true);
    sink(T(1), w2 * (z[2] - pref[2]), true);
    T dm = T(0), dp = T(0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T a = T(val(z[3 + i])) - qref[i], b = T(val(z[3 + i])) + qref[i];
        dm += a * a;
        dp += b * b;
    }
    const T sgn = dm > dp ? T(1) : T(-1);
#pragma unroll
    for (int i = 0; i < 4; ++i) sink(T(1), z[3 + i] + sgn * qref[i], true);
#pragma unroll
    for (int i = 0; i < 3; ++i) sink(T(1), z[7 + i] - vref[i], true);
#pragma unroll
    for (int i = 0; i < 3; ++i) sink(T(1), z[10 + i] - wref[i], true);
}

// =============================================================================================
struct Quadrotor {  // example/mpc/quadrotor.example.cpp
    static constexpr int KIND = 0, NX = 13, NU = 4, NZ = 17, NH = 8, LEGS = 0, HC = 1;
    static constexpr bool INEQ_SEPARABLE = true;  // every inequality row touches one local variable (:280-288)
    static constexpr bool TPN_DEFAULT = false, TPN_STRUCTURED_DEFAULT = true;  // measured defaults, DESIGN.md §5.3
    UB_HD static int n_dec(int N) { return NX * (N + 1) + NU * N; }
    UB_HD static int n_par(int N) { return 21 + 13 * (N + 1) + 13; }
    UB_HD static int m_eq(int N) { return NX * (N + 1); }
    UB_HD static int x_off(int N, int k) { return NX * k; }
    UB_HD static int u_off(int N, int k) { return NX * (N + 1) + NU * k; }
    UB_HD static int xm_off(int N) { return n_dec(N) + 21 + 13 * (N + 1); }

    // quadrotorDynamics, quadrotor.example.cpp:126-190
    template <class S, class T>
    UB_HD static void dynamics(const T* xp, int N, int k, const S* z, S* xn) {
        const T* P = xp + n_dec(N);
        const T dt = P[0], m = P[1], g0 = P[17], b = P[18], d = P[19];
        S Tz(T(0));
        Vec3<S> mom{S(T(0)), S(T(0)), S(T(0))};
        for (int i = 0; i < 4; ++i) {
            const S r2 = z[13 + i] * z[13 + i];          // Utils::Pow(r, 2)
            const S t  = b * r2;                          // bT_i = b r^2 e_z
            Tz += t;
            mom.x += P[5 + 3 * i + 1] * t;                // pP x (0, 0, t) = (py t, -px t, 0)
            mom.y -= P[5 + 3 * i + 0] * t;
            mom.z += ((i & 1) ? -d : d) * r2;             // drag moment d r^2 e_z (-1)^i
        }
        // q * (0, 0, Tz): third column of R(q) times Tz.  CppAD folds the literal zeros of Vector3ad::UnitZ()
        // (quadrotor.example.cpp:162), so the reference tape's z row does not depend on q.z / q.w; same form here.
        const S &qx = z[3], &qy = z[4], &qz = z[5], &qw = z[6];
        const Vec3<S> qT{T(2) * (qw * qy + qz * qx) * Tz, T(2) * (qz * qy - qw * qx) * Tz,
                         Tz - T(2) * (qx * qx + qy * qy) * Tz};
        const Vec3<S> pdd{qT.x / m, qT.y / m, (qT.z - m * g0) / m};
        const Vec3<S> w{z[10], z[11], z[12]};
        const Vec3<S> Iw{P[2] * w.x, P[3] * w.y, P[4] * w.z};
        const Vec3<S> rhs = mom - cross(w, Iw);
        const Vec3<S> wd{(T(1) / P[2]) * rhs.x, (T(1) / P[3]) * rhs.y, (T(1) / P[4]) * rhs.z};
        lie_euler(z, dt, pdd, wd, xn);
    }
    // inequalityConstraints, quadrotor.example.cpp:271-291
    template <class S, class T>
    UB_HD static void inequalities(const T* xp, int N, int k, const S* z, S* h) {
        const T rmax = xp[n_dec(N) + 20];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            h[2 * i]     = z[13 + i] - rmax;
            h[2 * i + 1] = -z[13 + i];
        }
    }
    // objectiveFunction, quadrotor.example.cpp:196-238
    template <class S, class T, class Sink>
    UB_HD static void cost_terms(const T* xp, int N, int k, const S* z, Sink&& sink) {
        const T* P = xp + n_dec(N);
        tracking_terms(z, P + 21 + 3 * k, P + 21 + 3 * (N + 1) + 4 * k, P + 21 + 7 * (N + 1) + 3 * k,
                       P + 21 + 10 * (N + 1) + 3 * k, T(1), T(1), T(1), sink);
        if (k == N) return;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (k) sink(T(1e-6), z[13 + i] - xp[u_off(N, k - 1) + i], true);          // :219-225
            sink(T(1e-6), z[13 + i], true);                                            // :226-230
            if (k + 1 < N) sink(T(1e-6), xp[u_off(N, k + 1) + i] - z[13 + i], false);  // term of node k+1
        }
    }
};

// =============================================================================================
struct RcCar {  // example/mpc/rc_car.example.cpp
    static constexpr int KIND = 1, NX = 6, NU = 2, NZ = 8, NH = 3, LEGS = 0, HC = 1;
    static constexpr bool INEQ_SEPARABLE = true;  // rc_car.example.cpp:271-282
    static constexpr bool TPN_DEFAULT = true, TPN_STRUCTURED_DEFAULT = false;
    UB_HD static int n_dec(int N) { return NX * (N + 1) + NU * N; }
    UB_HD static int n_par(int N) { return 15 + 2 * (N + 1) + 6; }
    UB_HD static int m_eq(int N) { return NX * (N + 1); }
    UB_HD static int x_off(int N, int k) { return NX * k; }
    UB_HD static int u_off(int N, int k) { return NX * (N + 1) + NU * k; }
    UB_HD static int xm_off(int N) { return n_dec(N) + 15 + 2 * (N + 1); }

    // Force model of rcCarDynamics (rc_car.example.cpp:158-171): (vx, vy, om, d, delta) -> (vx', vy', om').  P = parameters.
    template <class S, class T>
    UB_HD static void accelerations(const T* P, const S& vx, const S& vy, const S& om, const S& d, const S& delta, S* acc) {
        const T m = P[1], Iz = P[2], lf = P[3], lr = P[4], Bf = P[5], Cf = P[6], Df = P[7], Br = P[8], Cr = P[9], Dr = P[10],
                Cm1 = P[11], Cm2 = P[12], Cr0 = P[13], Cr2 = P[14];
        const S den    = vx + T(UB_EPS);
        const S alphaf = delta - m_atan((om * lf + vy) / den);
        const S alphar = m_atan((om * lr - vy) / den);
        const S Ffy    = Df * m_sin(Cf * m_atan(Bf * alphaf));
        const S Fry    = Dr * m_sin(Cr * m_atan(Br * alphar));
        const S Frx    = (Cm1 - Cm2 * vx) * d - Cr0 - Cr2 * (vx * vx);
        S sd, cd;
        m_sincos(delta, &sd, &cd);
        acc[0] = (Frx - Ffy * sd + m * vy * om) / m;
        acc[1] = (Fry + Ffy * cd - m * vx * om) / m;
        acc[2] = (Ffy * lf * cd - Fry * lr) / Iz;
    }
    // rcCarDynamics, rc_car.example.cpp:131-185.  z = [px py phi vx vy om | d delta]
    template <class S, class T>
    UB_HD static void dynamics(const T* xp, int N, int k, const S* z, S* xn) {
        const T* P = xp + n_dec(N);
        const T dt = P[0];
        const S &phi = z[2], &vx = z[3], &vy = z[4], &om = z[5];
        S acc[3];
        accelerations(P, vx, vy, om, z[6], z[7], acc);
        S sp, cp;
        m_sincos(phi, &sp, &cp);
        const S vxn = vx + dt * acc[0], vyn = vy + dt * acc[1], omn = om + dt * acc[2];
        xn[0] = z[0] + dt * (vxn * cp - vyn * sp);
        xn[1] = z[1] + dt * (vxn * sp + vyn * cp);
        xn[2] = phi + dt * omn;
        xn[3] = vxn;
        xn[4] = vyn;
        xn[5] = omn;
    }
    // inequalityConstraints, rc_car.example.cpp:264-285
    template <class S, class T>
    UB_HD static void inequalities(const T* xp, int N, int k, const S* z, S* h) {
        h[0] = m_abs(z[6]) - T(15);
        h[1] = m_abs(z[7]) - T(15);
        h[2] = T(0.3) - z[3];
    }
    // objectiveFunction, rc_car.example.cpp:197-231
    template <class S, class T, class Sink>
    UB_HD static void cost_terms(const T* xp, int N, int k, const S* z, Sink&& sink) {
        const T* pref = xp + n_dec(N) + 15 + 2 * k;
        sink(T(1), z[0] - pref[0], true);
        sink(T(1), z[1] - pref[1], true);
        if (k == N) return;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            sink(T(1e-6), z[6 + i], true);
            if (k) sink(T(1e-6), z[6 + i] - xp[u_off(N, k - 1) + i], true);
            if (k + 1 < N) sink(T(1e-6), xp[u_off(N, k + 1) + i] - z[6 + i], false);
        }
    }
};

// =============================================================================================
struct Quadruped {  // example/mpc/quadruped.example.cpp (single-rigid-body model)
    static constexpr int KIND = 2, NX = 13, NU = 24, NZ = 37, NH = 12, LEGS = 4, HC = 0;
    static constexpr bool INEQ_SEPARABLE = false;  // friction-cone and leg-length rows couple three locals
    static constexpr bool TPN_DEFAULT = false, TPN_STRUCTURED_DEFAULT = false;
    static constexpr int NP = 29;  // per-node parameter p_k (quadruped.example.cpp:94-95)
    UB_HD static int n_dec(int N) { return NX * (N + 1) + NU * N; }
    UB_HD static int n_par(int N) { return NP * (N + 1) + 49; }
    UB_HD static int m_eq(int N) { return NX * (N + 1) + 16 * N; }
    UB_HD static int x_off(int N, int k) { return NX * k; }
    UB_HD static int u_off(int N, int k) { return NX * (N + 1) + NU * k; }
    UB_HD static int p_off(int N, int k) { return n_dec(N) + NP * k; }
    UB_HD static int rho_off(int N) { return n_dec(N) + NP * (N + 1); }
    UB_HD static int xm_off(int N) { return rho_off(N) + 20; }

    // quadrupedDynamics, quadruped.example.cpp:148-203.  z = [p q v w | (f r) x 4]
    template <class S, class T>
    UB_HD static void dynamics(const T* xp, int N, int k, const S* z, S* xn) {
        const T* R  = xp + rho_off(N);
        const T* pk = xp + p_off(N, k);
        const T dt = R[0], m = R[1], g0 = R[18];
        const Quat<S> q{z[3], z[4], z[5], z[6]};
        const Vec3<S> w{z[10], z[11], z[12]};
        Vec3<S> pdd{S(T(0)), S(T(0)), S(-g0)};
        const Vec3<S> wIw = cross(w, Vec3<S>{R[2] * w.x, R[3] * w.y, R[4] * w.z});
        Vec3<S> wd{-wIw.x, -wIw.y, -wIw.z};
        for (int i = 0; i < 4; ++i) {
            const Vec3<S> f{z[13 + 6 * i], z[14 + 6 * i], z[15 + 6 * i]};
            const Vec3<S> r{z[16 + 6 * i], z[17 + 6 * i], z[18 + 6 * i]};
            const T s = pk[13 + 4 * i];
            pdd.x += s * f.x / m; pdd.y += s * f.y / m; pdd.z += s * f.z / m;
            const Vec3<S> t = cross(r, rotate(q, f));
            wd.x += s * t.x; wd.y += s * t.y; wd.z += s * t.z;
        }
        wd = Vec3<S>{wd.x / R[2], wd.y / R[3], wd.z / R[4]};
        lie_euler(z, dt, pdd, wd, xn);
    }
    // inequalityConstraints, quadruped.example.cpp:312-338
    template <class S, class T>
    UB_HD static void inequalities(const T* xp, int N, int k, const S* z, S* h) {
        const T* R  = xp + rho_off(N);
        const T* pk = xp + p_off(N, k);
        const T mu = R[19], L = R[17];
        for (int i = 0; i < 4; ++i) {
            const T s = pk[13 + 4 * i];
            const S &fx = z[13 + 6 * i], &fy = z[14 + 6 * i], &fz = z[15 + 6 * i];
            const Vec3<S> dr{z[16 + 6 * i] - R[5 + 3 * i], z[17 + 6 * i] - R[6 + 3 * i], z[18 + 6 * i] - R[7 + 3 * i]};
            h[3 * i]     = -s * fz;
            h[3 * i + 1] = s * approx_norm2(fx, fy) - mu * fz;
            h[3 * i + 2] = s * approx_norm(dr) - L;
        }
    }
    // objectiveFunction, quadruped.example.cpp:209-251
    template <class S, class T, class Sink>
    UB_HD static void cost_terms(const T* xp, int N, int k, const S* z, Sink&& sink) {
        const T* pk = xp + p_off(N, k);
        tracking_terms(z, pk, pk + 3, pk + 7, pk + 10, T(0.1), T(0.1), T(10), sink);
        if (k == N) return;
        for (int i = 0; i < 4; ++i)
            for (int c = 0; c < 3; ++c) {
                sink(T(1), z[16 + 6 * i + c] - pk[14 + 4 * i + c], true);  // |r - rRef|^2
                sink(T(1e-8), z[13 + 6 * i + c], true);                    // 1e-8 |f|^2
            }
    }
    // Contact rows of equalityConstraints, quadruped.example.cpp:279-303.  Local variables
    // zl = [p_k q_k r_{k,i} | p_{k-1} q_{k-1} r_{k-1,i}] (20); for k = 0 the previous foot is measured.
    template <class S, class T>
    UB_HD static void contact_rows(const T* xp, int N, int k, int leg, const S* zl, S* rows) {
        const T* R = xp + rho_off(N);
        const T s  = xp[p_off(N, k) + 13 + 4 * leg];
        const T sp = k ? xp[p_off(N, k - 1) + 13 + 4 * leg] : R[33 + 4 * leg];
        const Vec3<S> foot = Vec3<S>{zl[0], zl[1], zl[2]} +
                             rotate(Quat<S>{zl[3], zl[4], zl[5], zl[6]}, Vec3<S>{zl[7], zl[8], zl[9]});
        Vec3<S> prev;
        if (k) {
            prev = Vec3<S>{zl[10], zl[11], zl[12]} +
                   rotate(Quat<S>{zl[13], zl[14], zl[15], zl[16]}, Vec3<S>{zl[17], zl[18], zl[19]});
        } else {
            prev = Vec3<S>{S(R[34 + 4 * leg]), S(R[35 + 4 * leg]), S(R[36 + 4 * leg])};
        }
        const T ss = sp * s;
        rows[0] = ((T(1) - sp) * s) * foot.z;
        rows[1] = ss * (foot.x - prev.x);
        rows[2] = ss * (foot.y - prev.y);
        rows[3] = ss * (foot.z - prev.z);
    }
};

}  // namespace ub
