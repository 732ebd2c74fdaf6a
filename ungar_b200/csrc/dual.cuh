// Scalar types the model functors are instantiated with.
//
//   T (float | double)   plain evaluation
//   Dual<T>              forward-mode tangent propagation, ONE tangent direction per thread: the thread
//                        that owns tangent j of a shooting node seeds d z_j = 1 and ends up holding column j
//                        of every per-node Jacobian.
//   Dep                  structural dependency mask (host, model_create time only): which of the node's
//                        local inputs a value depends on.  It defines the CSR patterns exported through the
//                        C ABI the same way CppAD's pattern propagation does for the reference
//                        (include/ungar/autodiff/function.hpp:98-105, :529-574).
//
// Differentiation rules mirror the CppAD semantics the reference relies on: abs'(0) = 0, conditional
// expressions differentiate the selected branch, integer powers are repeated products.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#define UB_HD __host__ __device__ __forceinline__

namespace ub {

// ---------------------------------------------------------------------------------------------
template <class T>
struct Dual {
    T v, d;
    UB_HD Dual() : v(T(0)), d(T(0)) {}
    UB_HD Dual(T value) : v(value), d(T(0)) {}  // NOLINT: constants promote implicitly
    UB_HD Dual(T value, T tangent) : v(value), d(tangent) {}
};

template <class T> UB_HD Dual<T> operator+(const Dual<T>& a, const Dual<T>& b) { return {a.v + b.v, a.d + b.d}; }
template <class T> UB_HD Dual<T> operator-(const Dual<T>& a, const Dual<T>& b) { return {a.v - b.v, a.d - b.d}; }
template <class T> UB_HD Dual<T> operator*(const Dual<T>& a, const Dual<T>& b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
template <class T> UB_HD Dual<T> operator/(const Dual<T>& a, const Dual<T>& b) {
    const T inv = T(1) / b.v;
    const T q   = a.v * inv;
    return {q, (a.d - q * b.d) * inv};
}
template <class T> UB_HD Dual<T> operator-(const Dual<T>& a) { return {-a.v, -a.d}; }
template <class T> UB_HD Dual<T> operator+(const Dual<T>& a, T c) { return {a.v + c, a.d}; }
template <class T> UB_HD Dual<T> operator+(T c, const Dual<T>& a) { return {c + a.v, a.d}; }
template <class T> UB_HD Dual<T> operator-(const Dual<T>& a, T c) { return {a.v - c, a.d}; }
template <class T> UB_HD Dual<T> operator-(T c, const Dual<T>& a) { return {c - a.v, -a.d}; }
template <class T> UB_HD Dual<T> operator*(const Dual<T>& a, T c) { return {a.v * c, a.d * c}; }
template <class T> UB_HD Dual<T> operator*(T c, const Dual<T>& a) { return {c * a.v, c * a.d}; }
template <class T> UB_HD Dual<T> operator/(const Dual<T>& a, T c) { const T inv = T(1) / c; return {a.v * inv, a.d * inv}; }
template <class T> UB_HD Dual<T> operator/(T c, const Dual<T>& a) { const T inv = T(1) / a.v; const T q = c * inv; return {q, -q * a.d * inv}; }
template <class T> UB_HD Dual<T>& operator+=(Dual<T>& a, const Dual<T>& b) { a.v += b.v; a.d += b.d; return a; }
template <class T> UB_HD Dual<T>& operator-=(Dual<T>& a, const Dual<T>& b) { a.v -= b.v; a.d -= b.d; return a; }

// Plain scalars ---------------------------------------------------------------------------------
UB_HD float  m_sqrt(float x) { return sqrtf(x); }
UB_HD double m_sqrt(double x) { return sqrt(x); }
UB_HD float  m_atan(float x) { return atanf(x); }
UB_HD double m_atan(double x) { return atan(x); }
UB_HD float  m_abs(float x) { return fabsf(x); }
UB_HD double m_abs(double x) { return fabs(x); }
UB_HD void m_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
UB_HD void m_sincos(double x, double* s, double* c) { sincos(x, s, c); }
UB_HD float  m_sin(float x) { return sinf(x); }
UB_HD double m_sin(double x) { return sin(x); }
UB_HD float  m_cos(float x) { return cosf(x); }
UB_HD double m_cos(double x) { return cos(x); }
template <class T> UB_HD T val(T x) { return x; }
template <class T> UB_HD T select(bool c, T t, T f) { return c ? t : f; }

// Duals ------------------------------------------------------------------------------------------
template <class T> UB_HD T val(const Dual<T>& x) { return x.v; }
template <class T> UB_HD Dual<T> m_sqrt(const Dual<T>& x) { const T r = m_sqrt(x.v); return {r, x.d * (T(0.5) / r)}; }
template <class T> UB_HD Dual<T> m_atan(const Dual<T>& x) { return {m_atan(x.v), x.d / (T(1) + x.v * x.v)}; }
template <class T> UB_HD Dual<T> m_abs(const Dual<T>& x) {
    const T s = T((x.v > T(0)) - (x.v < T(0)));  // CppAD: sign(0) = 0
    return {m_abs(x.v), s * x.d};
}
template <class T> UB_HD void m_sincos(const Dual<T>& x, Dual<T>* s, Dual<T>* c) {
    T sv, cv;
    m_sincos(x.v, &sv, &cv);
    *s = {sv, cv * x.d};
    *c = {cv, -sv * x.d};
}
template <class T> UB_HD Dual<T> m_sin(const Dual<T>& x) { T s, c; m_sincos(x.v, &s, &c); return {s, c * x.d}; }
template <class T> UB_HD Dual<T> m_cos(const Dual<T>& x) { T s, c; m_sincos(x.v, &s, &c); return {c, -s * x.d}; }
template <class T> UB_HD Dual<T> select(bool c, const Dual<T>& t, const Dual<T>& f) { return c ? t : f; }

// ---------------------------------------------------------------------------------------------
// Structural dependency mask over at most 64 local inputs.  Host only in practice (model_create).
// `v` carries the primal value as a double so that value-dependent selects take a definite branch; the
// mask of a select is the union of both branches (pattern independent of the evaluation point).
// ---------------------------------------------------------------------------------------------
struct Dep {
    double v;
    uint64_t m;
    UB_HD Dep() : v(0.0), m(0) {}
    UB_HD Dep(double value) : v(value), m(0) {}  // NOLINT
    UB_HD Dep(double value, uint64_t mask) : v(value), m(mask) {}
};
UB_HD Dep operator+(const Dep& a, const Dep& b) { return {a.v + b.v, a.m | b.m}; }
UB_HD Dep operator-(const Dep& a, const Dep& b) { return {a.v - b.v, a.m | b.m}; }
UB_HD Dep operator*(const Dep& a, const Dep& b) { return {a.v * b.v, a.m | b.m}; }
UB_HD Dep operator/(const Dep& a, const Dep& b) { return {a.v / b.v, a.m | b.m}; }
UB_HD Dep operator-(const Dep& a) { return {-a.v, a.m}; }
UB_HD Dep operator+(const Dep& a, double c) { return {a.v + c, a.m}; }
UB_HD Dep operator+(double c, const Dep& a) { return {c + a.v, a.m}; }
UB_HD Dep operator-(const Dep& a, double c) { return {a.v - c, a.m}; }
UB_HD Dep operator-(double c, const Dep& a) { return {c - a.v, a.m}; }
UB_HD Dep operator*(const Dep& a, double c) { return {a.v * c, a.m}; }
UB_HD Dep operator*(double c, const Dep& a) { return {c * a.v, a.m}; }
UB_HD Dep operator/(const Dep& a, double c) { return {a.v / c, a.m}; }
UB_HD Dep operator/(double c, const Dep& a) { return {c / a.v, a.m}; }
UB_HD Dep& operator+=(Dep& a, const Dep& b) { a.v += b.v; a.m |= b.m; return a; }
UB_HD Dep& operator-=(Dep& a, const Dep& b) { a.v -= b.v; a.m |= b.m; return a; }
UB_HD double val(const Dep& x) { return x.v; }
UB_HD Dep m_sqrt(const Dep& x) { return {sqrt(x.v), x.m}; }
UB_HD Dep m_atan(const Dep& x) { return {atan(x.v), x.m}; }
UB_HD Dep m_abs(const Dep& x) { return {fabs(x.v), x.m}; }
UB_HD Dep m_sin(const Dep& x) { return {sin(x.v), x.m}; }
UB_HD Dep m_cos(const Dep& x) { return {cos(x.v), x.m}; }
UB_HD void m_sincos(const Dep& x, Dep* s, Dep* c) { *s = {sin(x.v), x.m}; *c = {cos(x.v), x.m}; }
UB_HD Dep select(bool c, const Dep& t, const Dep& f) { return {c ? t.v : f.v, t.m | f.m}; }

// Underlying real type of a scalar (float / double).
template <class S> struct real_of { using type = S; };
template <class T> struct real_of<Dual<T>> { using type = T; };
template <> struct real_of<Dep> { using type = double; };
template <class S> using real_t = typename real_of<S>::type;

// ---------------------------------------------------------------------------------------------
// Small fixed-size algebra shared by the models (Eigen 3.4.0 formulas the reference examples call).
// ---------------------------------------------------------------------------------------------
template <class S> struct Vec3 { S x, y, z; };
template <class S> struct Quat { S x, y, z, w; };  // Eigen coefficient order (variable_lazy_map.hpp:276-277)

template <class S> UB_HD Vec3<S> operator+(const Vec3<S>& a, const Vec3<S>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class S> UB_HD Vec3<S> operator-(const Vec3<S>& a, const Vec3<S>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class S> UB_HD Vec3<S> scale(const S& s, const Vec3<S>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <class S> UB_HD Vec3<S> cross(const Vec3<S>& a, const Vec3<S>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class S> UB_HD S dot(const Vec3<S>& a, const Vec3<S>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// QuaternionBase::_transformVector (Eigen/src/Geometry/Quaternion.h:531-541).
template <class S> UB_HD Vec3<S> rotate(const Quat<S>& q, const Vec3<S>& v) {
    const Vec3<S> qv{q.x, q.y, q.z};
    Vec3<S> uv = cross(qv, v);
    uv         = uv + uv;
    return v + scale(q.w, uv) + cross(qv, uv);
}
// Quaternion product (Eigen/src/Geometry/Quaternion.h:487-498).
template <class S> UB_HD Quat<S> qmul(const Quat<S>& a, const Quat<S>& b) {
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}

// Eigen::NumTraits<double>::epsilon(): the reference is fp64-only, so the fp32 kernels keep the same
// regulariser (it vanishes below fp32 resolution unless |v|^2 < 1e-9, and stays well defined at v = 0).
#define UB_EPS 2.220446049250313e-16

// Utils::ApproximateNorm (include/ungar/utils/utils.hpp:731-736).
template <class S> UB_HD S approx_norm(const Vec3<S>& v) { return m_sqrt(dot(v, v) + real_t<S>(UB_EPS)); }
template <class S> UB_HD S approx_norm2(const S& a, const S& b) { return m_sqrt(a * a + b * b + real_t<S>(UB_EPS)); }
// Utils::ApproximateExponentialMap (include/ungar/utils/utils.hpp:738-749).
template <class S> UB_HD Quat<S> approx_exp(const Vec3<S>& v) {
    const S n = approx_norm(v);
    S s, c;
    m_sincos(real_t<S>(0.5) * n, &s, &c);
    const S k = s / n;
    return {v.x * k, v.y * k, v.z * k, c};
}

}  // namespace ub
