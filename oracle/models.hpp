// ORACLE — TEST INFRASTRUCTURE ONLY (see ad.hpp header).
//
// CPU restatement, in scalar-generic C++, of the three NMPC models whose values and derivatives the
// reference evaluates through Ungar::Autodiff::Function.  Each function below follows one lambda of
// the reference examples and keeps the reference's flat layout
//     variables = [X | U | parameters],  X = [x_0 .. x_N],  U = [u_0 .. u_{N-1}]
// (example/mpc/quadruped.example.cpp:137-139) and the reference's row order.
// Quaternions are stored (x, y, z, w) (include/ungar/variable_lazy_map.hpp:276-277).
//
// The horizon N is a run-time argument here; the reference fixes N = 30 at compile time
// (quadrotor.example.cpp:52, rc_car.example.cpp:50, quadruped.example.cpp:57) and every size is
// expressed in terms of N, so N = 60 / 100 of BASELINE.json only changes the loop bounds.
#pragma once

#include <vector>

#include "ad.hpp"

namespace oracle {

// Eigen::NumTraits<double>::epsilon(), used by ApproximateNorm (utils/utils.hpp:731-736) and by the
// RC-car slip-angle denominators (rc_car.example.cpp:159-162).
constexpr double kEps = 2.220446049250313e-16;

template <class S> struct V3 { S x, y, z; };
template <class S> struct Q4 { S x, y, z, w; };  // Eigen coefficient order

template <class S> inline V3<S> operator+(const V3<S>& a, const V3<S>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class S> inline V3<S> operator-(const V3<S>& a, const V3<S>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class S> inline V3<S> operator*(const S& s, const V3<S>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <class S> inline V3<S> cross(const V3<S>& a, const V3<S>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class S> inline S sqnorm(const V3<S>& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
template <class S> inline V3<S> load3(const S* p) { return {p[0], p[1], p[2]}; }
template <class S> inline Q4<S> load4(const S* p) { return {p[0], p[1], p[2], p[3]}; }

// Eigen 3.4.0 QuaternionBase::_transformVector (Eigen/src/Geometry/Quaternion.h:531-541):
//   uv = q.vec x v;  uv += uv;  v + q.w * uv + q.vec x uv
template <class S> inline V3<S> rotate(const Q4<S>& q, const V3<S>& v) {
    const V3<S> qv{q.x, q.y, q.z};
    V3<S> uv = cross(qv, v);
    uv       = uv + uv;
    return v + q.w * uv + cross(qv, uv);
}
// Eigen 3.4.0 quaternion product (Eigen/src/Geometry/Quaternion.h:487-498).
template <class S> inline Q4<S> qmul(const Q4<S>& a, const Q4<S>& b) {
    Q4<S> r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
// Utils::ApproximateNorm (include/ungar/utils/utils.hpp:731-736).
template <class S> inline S approx_norm3(const V3<S>& v) { return ad_sqrt(sqnorm(v) + kEps); }
template <class S> inline S approx_norm2(const S& a, const S& b) { return ad_sqrt(a * a + b * b + kEps); }
// Utils::ApproximateExponentialMap (include/ungar/utils/utils.hpp:738-749).
template <class S> inline Q4<S> approx_exp(const V3<S>& v) {
    const S n = approx_norm3(v);
    const S s = ad_sin(0.5 * n);
    Q4<S> q;
    q.x = v.x * s / n;
    q.y = v.y * s / n;
    q.z = v.z * s / n;
    q.w = ad_cos(0.5 * n);
    return q;
}
// Utils::Min for AD scalars = CondExpGt(a, b, b, a) (include/ungar/utils/utils.hpp:969-982).
template <class S> inline S min_ad(const S& a, const S& b) { return cond_gt(a, b, b, a); }

// Tracking term shared by the quadrotor and quadruped objectives:
// Min(|q - qRef|^2, |q + qRef|^2)  (quadrotor.example.cpp:214-216, quadruped.example.cpp:229-231).
template <class S> inline S quat_distance(const Q4<S>& q, const Q4<S>& r) {
    const S dm = (q.x - r.x) * (q.x - r.x) + (q.y - r.y) * (q.y - r.y) + (q.z - r.z) * (q.z - r.z) +
                 (q.w - r.w) * (q.w - r.w);
    const S dp = (q.x + r.x) * (q.x + r.x) + (q.y + r.y) * (q.y + r.y) + (q.z + r.z) * (q.z + r.z) +
                 (q.w + r.w) * (q.w + r.w);
    return min_ad(dm, dp);
}

// Lie-group semi-implicit Euler update shared by quadrotor and quadruped
// (quadrotor.example.cpp:184-187, quadruped.example.cpp:197-200).
template <class S>
inline void lie_euler(const S* x, const S& dt, const V3<S>& pdd, const V3<S>& wd, S* xn) {
    const V3<S> p = load3(x), v = load3(x + 7), w = load3(x + 10);
    const Q4<S> q = load4(x + 3);
    const V3<S> vn = v + dt * pdd;
    const V3<S> wn = w + dt * wd;
    const V3<S> pn = p + dt * vn;
    const Q4<S> qn = qmul(q, approx_exp(dt * wn));
    xn[0] = pn.x; xn[1] = pn.y; xn[2] = pn.z;
    xn[3] = qn.x; xn[4] = qn.y; xn[5] = qn.z; xn[6] = qn.w;
    xn[7] = vn.x; xn[8] = vn.y; xn[9] = vn.z;
    xn[10] = wn.x; xn[11] = wn.y; xn[12] = wn.z;
}

// =============================================================================================
// Layouts
// =============================================================================================
struct Sizes {
    int nx, nu, N, n_dec, n_par, m_eq, m_ineq;
};

struct QuadrotorLayout {  // quadrotor.example.cpp:55-117
    int N;
    static constexpr int nx = 13, nu = 4;
    int X(int k) const { return nx * k; }
    int U(int k) const { return nx * (N + 1) + nu * k; }
    int n_dec() const { return nx * (N + 1) + nu * N; }
    int P() const { return n_dec(); }
    // parameters, in declaration order (quadrotor.example.cpp:104-116)
    int dt() const { return P() + 0; }
    int mass() const { return P() + 1; }
    int moi() const { return P() + 2; }
    int prop(int i) const { return P() + 5 + 3 * i; }
    int g() const { return P() + 17; }
    int b() const { return P() + 18; }
    int d() const { return P() + 19; }
    int rmax() const { return P() + 20; }
    int pref(int k) const { return P() + 21 + 3 * k; }
    int qref(int k) const { return P() + 21 + 3 * (N + 1) + 4 * k; }
    int vref(int k) const { return P() + 21 + 7 * (N + 1) + 3 * k; }
    int wref(int k) const { return P() + 21 + 10 * (N + 1) + 3 * k; }
    int xm() const { return P() + 21 + 13 * (N + 1); }
    int n_par() const { return 21 + 13 * (N + 1) + 13; }
    int m_eq() const { return nx + nx * N; }
    int m_ineq() const { return 2 * nu * N; }
    Sizes sizes() const { return {nx, nu, N, n_dec(), n_par(), m_eq(), m_ineq()}; }
};

struct RcCarLayout {  // rc_car.example.cpp:53-122
    int N;
    static constexpr int nx = 6, nu = 2;
    int X(int k) const { return nx * k; }
    int U(int k) const { return nx * (N + 1) + nu * k; }
    int n_dec() const { return nx * (N + 1) + nu * N; }
    int P() const { return n_dec(); }
    // dt m Iz lf lr Bf Cf Df Br Cr Dr Cm1 Cm2 Cr0 Cr2 (rc_car.example.cpp:105-119)
    int par(int i) const { return P() + i; }
    int pref(int k) const { return P() + 15 + 2 * k; }
    int xm() const { return P() + 15 + 2 * (N + 1); }
    int n_par() const { return 15 + 2 * (N + 1) + 6; }
    int m_eq() const { return nx + nx * N; }
    int m_ineq() const { return 3 * N; }
    Sizes sizes() const { return {nx, nu, N, n_dec(), n_par(), m_eq(), m_ineq()}; }
};

struct QuadrupedLayout {  // quadruped.example.cpp:60-139
    int N;
    static constexpr int nx = 13, nu = 24, np = 29, legs = 4;
    int X(int k) const { return nx * k; }
    int U(int k) const { return nx * (N + 1) + nu * k; }
    int F(int k, int i) const { return U(k) + 6 * i; }
    int R(int k, int i) const { return U(k) + 6 * i + 3; }
    int n_dec() const { return nx * (N + 1) + nu * N; }
    int P(int k) const { return n_dec() + np * k; }
    int S(int k, int i) const { return P(k) + 13 + 4 * i; }
    int Rref(int k, int i) const { return P(k) + 14 + 4 * i; }
    int Rho() const { return n_dec() + np * (N + 1); }
    int dt() const { return Rho() + 0; }
    int mass() const { return Rho() + 1; }
    int moi() const { return Rho() + 2; }
    int hip(int i) const { return Rho() + 5 + 3 * i; }
    int leg_length() const { return Rho() + 17; }
    int g() const { return Rho() + 18; }
    int mu() const { return Rho() + 19; }
    int xm() const { return Rho() + 20; }
    int s_meas(int i) const { return Rho() + 33 + 4 * i; }
    int foot_meas(int i) const { return Rho() + 34 + 4 * i; }
    int n_par() const { return np * (N + 1) + 49; }
    int m_eq() const { return nx + nx * N + 4 * legs * N; }
    int m_ineq() const { return 3 * legs * N; }
    Sizes sizes() const { return {nx, nu, N, n_dec(), n_par(), m_eq(), m_ineq()}; }
};

// =============================================================================================
// Quadrotor  (example/mpc/quadrotor.example.cpp)
// =============================================================================================
// quadrotorDynamics, quadrotor.example.cpp:126-190.
template <class S>
inline void quadrotor_dynamics(const QuadrotorLayout& L, const S* xp, const S* x, const S* u, S* xn) {
    const S& dt = xp[L.dt()];
    const S& g0 = xp[L.g()];
    const S& b  = xp[L.b()];
    const S& d  = xp[L.d()];
    const S& m  = xp[L.mass()];
    const V3<S> moi = load3(xp + L.moi());
    const Q4<S> q   = load4(x + 3);
    const V3<S> w   = load3(x + 10);

    // :157-165  bT_i = b r_i^2 e_z,  bM_i = pP_i x bT_i,  bD_i = d r_i^2 e_z (-1)^i
    V3<S> sumT{S(0.0), S(0.0), S(0.0)}, sumM{S(0.0), S(0.0), S(0.0)}, sumD{S(0.0), S(0.0), S(0.0)};
    for (int i = 0; i < 4; ++i) {
        const S r2 = u[i] * u[i];  // Utils::Pow(r, 2) -> CppAD::pow(x, int) = repeated product
        const V3<S> T{S(0.0), S(0.0), b * r2};
        const V3<S> pP = load3(xp + L.prop(i));
        sumT = sumT + T;
        sumM = sumM + cross(pP, T);
        const double sign = (i % 2 == 0) ? 1.0 : -1.0;  // Utils::Pow(-1.0, i)
        sumD = sumD + V3<S>{S(0.0), S(0.0), d * r2 * sign};
    }
    // :171-174.  The summed thrust is (0, 0, T): CppAD folds the products with the literal zeros of
    // Vector3ad::UnitZ() (:162), so q * sumT is taped as the third column of R(q) times T — restated here in that
    // folded form so that the structural sparsity equals the reference tape's (z row independent of q.z, q.w).
    const S& Tz = sumT.z;
    const V3<S> qT{2.0 * (q.w * q.y + q.z * q.x) * Tz, 2.0 * (q.z * q.y - q.w * q.x) * Tz,
                   Tz - 2.0 * (q.x * q.x + q.y * q.y) * Tz};
    const V3<S> pdd{qT.x / m, qT.y / m, (qT.z - m * g0) / m};
    const V3<S> Iw{moi.x * w.x, moi.y * w.y, moi.z * w.z};
    const V3<S> rhs = sumM + sumD - cross(w, Iw);
    const V3<S> wd{(1.0 / moi.x) * rhs.x, (1.0 / moi.y) * rhs.y, (1.0 / moi.z) * rhs.z};
    lie_euler(x, dt, pdd, wd, xn);
}

// objectiveFunction, quadrotor.example.cpp:196-238.
template <class S> inline void quadrotor_objective(int N, const S* xp, std::vector<S>& y) {
    const QuadrotorLayout L{N};
    S value(0.0);
    for (int k = 0; k <= N; ++k) {
        const S* x = xp + L.X(k);
        const V3<S> dp = load3(x) - load3(xp + L.pref(k));
        const V3<S> dv = load3(x + 7) - load3(xp + L.vref(k));
        const V3<S> dw = load3(x + 10) - load3(xp + L.wref(k));
        value += sqnorm(dp) + quat_distance(load4(x + 3), load4(xp + L.qref(k))) + sqnorm(dv) + sqnorm(dw);
        if (k && k != N) {  // :219-225
            S acc(0.0);
            for (int i = 0; i < 4; ++i) {
                const S du = xp[L.U(k) + i] - xp[L.U(k - 1) + i];
                acc += du * du;
            }
            value += 1e-6 * acc;
        }
        if (k != N) {  // :226-230
            S acc(0.0);
            for (int i = 0; i < 4; ++i) acc += xp[L.U(k) + i] * xp[L.U(k) + i];
            value += 1e-6 * acc;
        }
    }
    y.assign(1, value);
}

// equalityConstraints, quadrotor.example.cpp:243-266.
template <class S> inline void quadrotor_equalities(int N, const S* xp, std::vector<S>& y) {
    const QuadrotorLayout L{N};
    y.clear();
    for (int i = 0; i < 13; ++i) y.push_back(xp[L.X(0) + i] - xp[L.xm() + i]);
    for (int k = 0; k < N; ++k) {
        S xn[13];
        quadrotor_dynamics(L, xp, xp + L.X(k), xp + L.U(k), xn);
        for (int i = 0; i < 13; ++i) y.push_back(xp[L.X(k + 1) + i] - xn[i]);
    }
}

// inequalityConstraints, quadrotor.example.cpp:271-291.
template <class S> inline void quadrotor_inequalities(int N, const S* xp, std::vector<S>& y) {
    const QuadrotorLayout L{N};
    y.clear();
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 4; ++i) {
            const S& r = xp[L.U(k) + i];
            y.push_back(r - xp[L.rmax()]);
            y.push_back(-r);
        }
}

// =============================================================================================
// RC car  (example/mpc/rc_car.example.cpp)
// =============================================================================================
// rcCarDynamics, rc_car.example.cpp:131-185.  State [p(2) phi v(2) omega], input [d delta].
template <class S>
inline void rc_car_dynamics(const RcCarLayout& L, const S* xp, const S* x, const S* u, S* xn) {
    const S &dt = xp[L.par(0)], &m = xp[L.par(1)], &Iz = xp[L.par(2)], &lf = xp[L.par(3)], &lr = xp[L.par(4)];
    const S &Bf = xp[L.par(5)], &Cf = xp[L.par(6)], &Df = xp[L.par(7)];
    const S &Br = xp[L.par(8)], &Cr = xp[L.par(9)], &Dr = xp[L.par(10)];
    const S &Cm1 = xp[L.par(11)], &Cm2 = xp[L.par(12)], &Cr0 = xp[L.par(13)], &Cr2 = xp[L.par(14)];
    const S &px = x[0], &py = x[1], &phi = x[2], &vx = x[3], &vy = x[4], &om = x[5];
    const S &d = u[0], &delta = u[1];

    // :158-165
    const S alphaf = -ad_atan((om * lf + vy) / (vx + kEps)) + delta;
    const S alphar = ad_atan((om * lr - vy) / (vx + kEps));
    const S Ffy    = Df * ad_sin(Cf * ad_atan(Bf * alphaf));
    const S Fry    = Dr * ad_sin(Cr * ad_atan(Br * alphar));
    const S Frx    = (Cm1 - Cm2 * vx) * d - Cr0 - Cr2 * (vx * vx);
    // :168-171
    const S vdx = (Frx - Ffy * ad_sin(delta) + m * vy * om) / m;
    const S vdy = (Fry + Ffy * ad_cos(delta) - m * vx * om) / m;
    const S omd = (Ffy * lf * ad_cos(delta) - Fry * lr) / Iz;
    // :180-184 (position uses the old yaw and the new velocity)
    const S vxn = vx + dt * vdx;
    const S vyn = vy + dt * vdy;
    const S omn = om + dt * omd;
    xn[0] = px + dt * (vxn * ad_cos(phi) - vyn * ad_sin(phi));
    xn[1] = py + dt * (vxn * ad_sin(phi) + vyn * ad_cos(phi));
    xn[2] = phi + dt * omn;
    xn[3] = vxn;
    xn[4] = vyn;
    xn[5] = omn;
}

// objectiveFunction, rc_car.example.cpp:197-231.
template <class S> inline void rc_car_objective(int N, const S* xp, std::vector<S>& y) {
    const RcCarLayout L{N};
    S value(0.0);
    for (int k = 0; k < N; ++k) {
        const S dx = xp[L.X(k)] - xp[L.pref(k)], dy = xp[L.X(k) + 1] - xp[L.pref(k) + 1];
        const S &u0 = xp[L.U(k)], &u1 = xp[L.U(k) + 1];
        value += dx * dx + dy * dy;
        value += 1e-6 * (u0 * u0 + u1 * u1);
        if (k) {
            const S e0 = u0 - xp[L.U(k - 1)], e1 = u1 - xp[L.U(k - 1) + 1];
            value += 1e-6 * (e0 * e0 + e1 * e1);
        }
    }
    const S dx = xp[L.X(N)] - xp[L.pref(N)], dy = xp[L.X(N) + 1] - xp[L.pref(N) + 1];
    value += dx * dx + dy * dy;
    y.assign(1, value);
}

// equalityConstraints, rc_car.example.cpp:236-259.
template <class S> inline void rc_car_equalities(int N, const S* xp, std::vector<S>& y) {
    const RcCarLayout L{N};
    y.clear();
    for (int i = 0; i < 6; ++i) y.push_back(xp[L.X(0) + i] - xp[L.xm() + i]);
    for (int k = 0; k < N; ++k) {
        S xn[6];
        rc_car_dynamics(L, xp, xp + L.X(k), xp + L.U(k), xn);
        for (int i = 0; i < 6; ++i) y.push_back(xp[L.X(k + 1) + i] - xn[i]);
    }
}

// inequalityConstraints, rc_car.example.cpp:264-285.
template <class S> inline void rc_car_inequalities(int N, const S* xp, std::vector<S>& y) {
    const RcCarLayout L{N};
    y.clear();
    for (int k = 0; k < N; ++k) {
        y.push_back(ad_abs(xp[L.U(k)]) - 15.0);
        y.push_back(ad_abs(xp[L.U(k) + 1]) - 15.0);
        y.push_back(0.3 - xp[L.X(k) + 3]);
    }
}

// =============================================================================================
// Quadruped single-rigid-body model  (example/mpc/quadruped.example.cpp)
// =============================================================================================
// quadrupedDynamics, quadruped.example.cpp:148-203.
template <class S>
inline void quadruped_dynamics(const QuadrupedLayout& L, const S* xp, int k, S* xn) {
    const S* x = xp + L.X(k);
    const S& dt = xp[L.dt()];
    const S& g0 = xp[L.g()];
    const S& m  = xp[L.mass()];
    const V3<S> moi = load3(xp + L.moi());
    const Q4<S> q   = load4(x + 3);
    const V3<S> w   = load3(x + 10);

    // :168-178
    V3<S> pdd{S(0.0), S(0.0), -g0};
    const V3<S> Iw{moi.x * w.x, moi.y * w.y, moi.z * w.z};
    const V3<S> wxIw = cross(w, Iw);
    V3<S> wd{-wxIw.x, -wxIw.y, -wxIw.z};
    for (int i = 0; i < 4; ++i) {
        const V3<S> f = load3(xp + L.F(k, i));
        const V3<S> r = load3(xp + L.R(k, i));
        const S& s    = xp[L.S(k, i)];
        pdd = pdd + V3<S>{s * f.x / m, s * f.y / m, s * f.z / m};
        wd  = wd + s * cross(r, rotate(q, f));
    }
    wd = V3<S>{wd.x / moi.x, wd.y / moi.y, wd.z / moi.z};  // :179
    lie_euler(x, dt, pdd, wd, xn);
}

// objectiveFunction, quadruped.example.cpp:209-251.
template <class S> inline void quadruped_objective(int N, const S* xp, std::vector<S>& y) {
    const QuadrupedLayout L{N};
    S value(0.0);
    for (int k = 0; k <= N; ++k) {
        const S* x  = xp + L.X(k);
        const S* pr = xp + L.P(k);
        const V3<S> dp = load3(x) - load3(pr);
        const V3<S> wp{0.1 * dp.x, 0.1 * dp.y, 10.0 * dp.z};  // Vector3r{0.1,0.1,10}.cwiseProduct(p - pRef)
        const V3<S> dv = load3(x + 7) - load3(pr + 7);
        const V3<S> dw = load3(x + 10) - load3(pr + 10);
        value += sqnorm(wp) + quat_distance(load4(x + 3), load4(pr + 3)) + sqnorm(dv) + sqnorm(dw);
        if (k != N) {
            for (int i = 0; i < 4; ++i) {
                const V3<S> f  = load3(xp + L.F(k, i));
                const V3<S> dr = load3(xp + L.R(k, i)) - load3(xp + L.Rref(k, i));
                value += sqnorm(dr);
                value += 1e-8 * sqnorm(f);
            }
        }
    }
    y.assign(1, value);
}

// equalityConstraints, quadruped.example.cpp:256-307.
template <class S> inline void quadruped_equalities(int N, const S* xp, std::vector<S>& y) {
    const QuadrupedLayout L{N};
    y.clear();
    for (int i = 0; i < 13; ++i) y.push_back(xp[L.X(0) + i] - xp[L.xm() + i]);
    for (int k = 0; k < N; ++k) {
        S xn[13];
        quadruped_dynamics(L, xp, k, xn);
        for (int i = 0; i < 13; ++i) y.push_back(xp[L.X(k + 1) + i] - xn[i]);
    }
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < 4; ++i) {
            const S& s     = xp[L.S(k, i)];
            const S& sPrev = k ? xp[L.S(k - 1, i)] : xp[L.s_meas(i)];
            const V3<S> pFoot =
                load3(xp + L.X(k)) + rotate(load4(xp + L.X(k) + 3), load3(xp + L.R(k, i)));
            V3<S> pFootPrev;
            if (k) {
                pFootPrev = load3(xp + L.X(k - 1)) +
                            rotate(load4(xp + L.X(k - 1) + 3), load3(xp + L.R(k - 1, i)));
            } else {
                pFootPrev = load3(xp + L.foot_meas(i));
            }
            y.push_back((1.0 - sPrev) * s * pFoot.z);  // :300
            const S ss = sPrev * s;                    // :301
            y.push_back(ss * (pFoot.x - pFootPrev.x));
            y.push_back(ss * (pFoot.y - pFootPrev.y));
            y.push_back(ss * (pFoot.z - pFootPrev.z));
        }
    }
}

// inequalityConstraints, quadruped.example.cpp:312-338.
template <class S> inline void quadruped_inequalities(int N, const S* xp, std::vector<S>& y) {
    const QuadrupedLayout L{N};
    y.clear();
    const S& mu = xp[L.mu()];
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < 4; ++i) {
            const S& s    = xp[L.S(k, i)];
            const V3<S> f = load3(xp + L.F(k, i));
            const V3<S> r = load3(xp + L.R(k, i));
            y.push_back(-s * f.z);
            y.push_back(s * approx_norm2(f.x, f.y) - mu * f.z);
            y.push_back(s * approx_norm3(r - load3(xp + L.hip(i))) - xp[L.leg_length()]);
        }
    }
}

// =============================================================================================
// Relaxed polynomial barrier  (include/ungar/optimization/soft_inequality_constraint.hpp:131-205),
// as instantiated by SoftSQPOptimizer::MakeSoftInequalityConstraintFunction
// (include/ungar/optimization/soft_sqp.hpp:114-138): Zsoft(z) = sum_i b(-z_i), rhs = 0.
// =============================================================================================
struct PolyBarrier {
    double eps, a1, b1, c1, a2, b2, c2, d2;
    PolyBarrier(double stiffness, double epsilon) : eps(epsilon) {  // :133-145
        a1 = stiffness;
        b1 = -0.5 * a1 * eps;
        c1 = -1.0 / 3.0 * (-b1 - a1 * eps) * eps - 0.5 * a1 * eps * eps - b1 * eps;
        a2 = (-b1 - a1 * eps) / (eps * eps);
        b2 = a1;
        c2 = b1;
        d2 = c1;
    }
    // EvaluateImpl(ad_scalar_t), :181-190, via nested CondExpLt.
    template <class S> S eval(const S& x) const {
        const S quad  = 0.5 * a1 * (x * x) + b1 * x + c1;
        const S cubic = 1.0 / 3.0 * a2 * (x * x * x) + 0.5 * b2 * (x * x) + c2 * x + d2;
        return cond_lt(x, S(0.0), quad, cond_lt(x, S(eps), cubic, S(0.0)));
    }
    template <class S> S soft_constraint(const std::vector<S>& z) const {  // soft_sqp.hpp:116-125
        S sum(0.0);
        for (const S& zi : z) sum += eval(-zi);
        return sum;
    }
};

}  // namespace oracle
