// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link, import or execute this code.  The product (ungar_b200/csrc) never includes it.
//
// Forward-mode automatic differentiation scalars used by the CPU restatement of the reference's
// derivative path.  The reference obtains derivatives from CppAD 20230000.0 + CppADCodeGen
// v2.4.3-ungar (absent from /root/reference); the restated semantics are the ones the
// reference relies on at its call sites:
//   * sparse Jacobian / per-equation sparse Hessian      include/ungar/autodiff/function.hpp:468-483
//   * structural sparsity = dependency sets              include/ungar/autodiff/function.hpp:529-574
//   * CondExpXx(a, b, t, f): value of the selected branch, pattern = union of both branches
//                                                        include/ungar/utils/utils.hpp:969-982
//   * abs'(0) = 0 (CppAD sign convention)                include/ungar/utils/utils.hpp:1001-1015
//   * pow(x, int) by repeated multiplication             include/ungar/utils/utils.hpp:820-837
#pragma once

#include <array>
#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Plain-double overloads so that the same templated model code runs on double and on duals.
// ---------------------------------------------------------------------------------------------
inline double value_of(double x) { return x; }
inline double ad_sqrt(double x) { return std::sqrt(x); }
inline double ad_sin(double x) { return std::sin(x); }
inline double ad_cos(double x) { return std::cos(x); }
inline double ad_atan(double x) { return std::atan(x); }
inline double ad_abs(double x) { return std::fabs(x); }
inline double cond_gt(double a, double b, double t, double f) { return a > b ? t : f; }
inline double cond_lt(double a, double b, double t, double f) { return a < b ? t : f; }
inline double cond_select(bool take_t, double t, double f) { return take_t ? t : f; }

// ---------------------------------------------------------------------------------------------
// Sparse forward dual.  `T` is double (first order) or SDual<double> (second order, nested).
// The tangent is a list of (independent index, partial) pairs sorted by index; an index is present
// iff the value depends *structurally* on that independent (no numeric folding), which is the
// notion of sparsity CppAD's pattern propagation uses.
// ---------------------------------------------------------------------------------------------
template <class T>
struct SDual {
    T v{};
    std::vector<std::pair<int, T>> d;

    SDual() = default;
    SDual(double c) : v(c) {}  // NOLINT: literal constants promote implicitly, like CppAD parameters
    template <class U = T, class = std::enable_if_t<!std::is_same<U, double>::value>>
    SDual(const T& c) : v(c) {}
    SDual(T value, std::vector<std::pair<int, T>> tangent) : v(std::move(value)), d(std::move(tangent)) {}
};

template <class T>
inline double value_of(const SDual<T>& x) { return value_of(x.v); }

// z = a * x.d + b * y.d  (sorted merge).
template <class T>
inline std::vector<std::pair<int, T>> lincomb(const T& a, const std::vector<std::pair<int, T>>& x,
                                              const T& b, const std::vector<std::pair<int, T>>& y) {
    std::vector<std::pair<int, T>> z;
    z.reserve(x.size() + y.size());
    std::size_t i = 0, j = 0;
    while (i < x.size() || j < y.size()) {
        if (j == y.size() || (i < x.size() && x[i].first < y[j].first)) {
            z.emplace_back(x[i].first, a * x[i].second);
            ++i;
        } else if (i == x.size() || y[j].first < x[i].first) {
            z.emplace_back(y[j].first, b * y[j].second);
            ++j;
        } else {
            z.emplace_back(x[i].first, a * x[i].second + b * y[j].second);
            ++i;
            ++j;
        }
    }
    return z;
}

template <class T>
inline std::vector<std::pair<int, T>> scaled(const T& a, const std::vector<std::pair<int, T>>& x) {
    std::vector<std::pair<int, T>> z;
    z.reserve(x.size());
    for (const auto& e : x) z.emplace_back(e.first, a * e.second);
    return z;
}

template <class T> inline SDual<T> operator+(const SDual<T>& x, const SDual<T>& y) {
    return {x.v + y.v, lincomb(T(1.0), x.d, T(1.0), y.d)};
}
template <class T> inline SDual<T> operator-(const SDual<T>& x, const SDual<T>& y) {
    return {x.v - y.v, lincomb(T(1.0), x.d, T(-1.0), y.d)};
}
template <class T> inline SDual<T> operator*(const SDual<T>& x, const SDual<T>& y) {
    return {x.v * y.v, lincomb(y.v, x.d, x.v, y.d)};
}
template <class T> inline SDual<T> operator/(const SDual<T>& x, const SDual<T>& y) {
    const T inv = T(1.0) / y.v;
    const T q   = x.v * inv;
    return {q, lincomb(inv, x.d, T(-1.0) * q * inv, y.d)};
}
template <class T> inline SDual<T> operator-(const SDual<T>& x) { return {T(-1.0) * x.v, scaled(T(-1.0), x.d)}; }

// Mixed with literal doubles.
template <class T> inline SDual<T> operator+(const SDual<T>& x, double c) { return {x.v + T(c), x.d}; }
template <class T> inline SDual<T> operator+(double c, const SDual<T>& x) { return {T(c) + x.v, x.d}; }
template <class T> inline SDual<T> operator-(const SDual<T>& x, double c) { return {x.v - T(c), x.d}; }
template <class T> inline SDual<T> operator-(double c, const SDual<T>& x) { return {T(c) - x.v, scaled(T(-1.0), x.d)}; }
template <class T> inline SDual<T> operator*(const SDual<T>& x, double c) { return {x.v * T(c), scaled(T(c), x.d)}; }
template <class T> inline SDual<T> operator*(double c, const SDual<T>& x) { return {T(c) * x.v, scaled(T(c), x.d)}; }
template <class T> inline SDual<T> operator/(const SDual<T>& x, double c) { return x * (1.0 / c); }
template <class T> inline SDual<T> operator/(double c, const SDual<T>& x) { return SDual<T>(c) / x; }
template <class T> inline SDual<T>& operator+=(SDual<T>& x, const SDual<T>& y) { x = x + y; return x; }
template <class T> inline SDual<T>& operator-=(SDual<T>& x, const SDual<T>& y) { x = x - y; return x; }

template <class T> inline SDual<T> ad_sqrt(const SDual<T>& x) {
    const T r = ad_sqrt(x.v);
    return {r, scaled(T(0.5) / r, x.d)};
}
template <class T> inline SDual<T> ad_sin(const SDual<T>& x) { return {ad_sin(x.v), scaled(ad_cos(x.v), x.d)}; }
template <class T> inline SDual<T> ad_cos(const SDual<T>& x) { return {ad_cos(x.v), scaled(T(-1.0) * ad_sin(x.v), x.d)}; }
template <class T> inline SDual<T> ad_atan(const SDual<T>& x) {
    return {ad_atan(x.v), scaled(T(1.0) / (T(1.0) + x.v * x.v), x.d)};
}
// CppAD: d|x|/dx = sign(x) with sign(0) = 0.
template <class T> inline SDual<T> ad_abs(const SDual<T>& x) {
    const double xv = value_of(x.v);
    const double s  = (xv > 0.0) - (xv < 0.0);
    return {ad_abs(x.v), scaled(T(s), x.d)};
}

// CondExp: value/partials of the selected branch; indices of the other branch are kept (with zero
// partials) so the structural pattern does not depend on the evaluation point.
template <class T>
inline SDual<T> cond_select(bool take_t, const SDual<T>& t, const SDual<T>& f) {
    const SDual<T>& sel = take_t ? t : f;
    const SDual<T>& oth = take_t ? f : t;
    return {sel.v, lincomb(T(1.0), sel.d, T(0.0), oth.d)};
}
template <class T>
inline SDual<T> cond_gt(const SDual<T>& a, const SDual<T>& b, const SDual<T>& t, const SDual<T>& f) {
    return cond_select(value_of(a) > value_of(b), t, f);
}
template <class T>
inline SDual<T> cond_lt(const SDual<T>& a, const SDual<T>& b, const SDual<T>& t, const SDual<T>& f) {
    return cond_select(value_of(a) < value_of(b), t, f);
}

using Dual1 = SDual<double>;
using Dual2 = SDual<Dual1>;

inline Dual1 seed1(double x, int index) { return Dual1{x, {{index, 1.0}}}; }
inline Dual2 seed2(double x, int index) {
    return Dual2{Dual1{x, {{index, 1.0}}}, {{index, Dual1{1.0}}}};
}

// ---------------------------------------------------------------------------------------------
// Dense forward dual of compile-time width K — used by the stage-wise CPU baseline (fast path:
// the tangent loops vectorise).  Same differentiation rules as SDual.
// ---------------------------------------------------------------------------------------------
template <int K>
struct DDual {
    double v = 0.0;
    std::array<double, K> d{};
    DDual() = default;
    DDual(double c) : v(c) {}  // NOLINT
};
template <int K> inline double value_of(const DDual<K>& x) { return x.v; }
#define ORACLE_DD_LOOP for (int i_ = 0; i_ < K; ++i_)
template <int K> inline DDual<K> operator+(const DDual<K>& x, const DDual<K>& y) { DDual<K> z; z.v = x.v + y.v; ORACLE_DD_LOOP z.d[i_] = x.d[i_] + y.d[i_]; return z; }
template <int K> inline DDual<K> operator-(const DDual<K>& x, const DDual<K>& y) { DDual<K> z; z.v = x.v - y.v; ORACLE_DD_LOOP z.d[i_] = x.d[i_] - y.d[i_]; return z; }
template <int K> inline DDual<K> operator*(const DDual<K>& x, const DDual<K>& y) { DDual<K> z; z.v = x.v * y.v; ORACLE_DD_LOOP z.d[i_] = x.d[i_] * y.v + x.v * y.d[i_]; return z; }
template <int K> inline DDual<K> operator/(const DDual<K>& x, const DDual<K>& y) { DDual<K> z; const double inv = 1.0 / y.v; z.v = x.v * inv; const double c = -z.v * inv; ORACLE_DD_LOOP z.d[i_] = x.d[i_] * inv + c * y.d[i_]; return z; }
template <int K> inline DDual<K> operator-(const DDual<K>& x) { DDual<K> z; z.v = -x.v; ORACLE_DD_LOOP z.d[i_] = -x.d[i_]; return z; }
template <int K> inline DDual<K> operator+(const DDual<K>& x, double c) { DDual<K> z = x; z.v += c; return z; }
template <int K> inline DDual<K> operator+(double c, const DDual<K>& x) { return x + c; }
template <int K> inline DDual<K> operator-(const DDual<K>& x, double c) { DDual<K> z = x; z.v -= c; return z; }
template <int K> inline DDual<K> operator-(double c, const DDual<K>& x) { DDual<K> z; z.v = c - x.v; ORACLE_DD_LOOP z.d[i_] = -x.d[i_]; return z; }
template <int K> inline DDual<K> operator*(const DDual<K>& x, double c) { DDual<K> z; z.v = x.v * c; ORACLE_DD_LOOP z.d[i_] = x.d[i_] * c; return z; }
template <int K> inline DDual<K> operator*(double c, const DDual<K>& x) { return x * c; }
template <int K> inline DDual<K> operator/(const DDual<K>& x, double c) { return x * (1.0 / c); }
template <int K> inline DDual<K> operator/(double c, const DDual<K>& x) { return DDual<K>(c) / x; }
template <int K> inline DDual<K>& operator+=(DDual<K>& x, const DDual<K>& y) { x = x + y; return x; }
template <int K> inline DDual<K>& operator-=(DDual<K>& x, const DDual<K>& y) { x = x - y; return x; }
template <int K> inline DDual<K> dd_chain(double f, double df, const DDual<K>& x) { DDual<K> z; z.v = f; ORACLE_DD_LOOP z.d[i_] = df * x.d[i_]; return z; }
template <int K> inline DDual<K> ad_sqrt(const DDual<K>& x) { const double r = std::sqrt(x.v); return dd_chain(r, 0.5 / r, x); }
template <int K> inline DDual<K> ad_sin(const DDual<K>& x) { return dd_chain(std::sin(x.v), std::cos(x.v), x); }
template <int K> inline DDual<K> ad_cos(const DDual<K>& x) { return dd_chain(std::cos(x.v), -std::sin(x.v), x); }
template <int K> inline DDual<K> ad_atan(const DDual<K>& x) { return dd_chain(std::atan(x.v), 1.0 / (1.0 + x.v * x.v), x); }
template <int K> inline DDual<K> ad_abs(const DDual<K>& x) { return dd_chain(std::fabs(x.v), double((x.v > 0.0) - (x.v < 0.0)), x); }
template <int K> inline DDual<K> cond_gt(const DDual<K>& a, const DDual<K>& b, const DDual<K>& t, const DDual<K>& f) { return a.v > b.v ? t : f; }
template <int K> inline DDual<K> cond_lt(const DDual<K>& a, const DDual<K>& b, const DDual<K>& t, const DDual<K>& f) { return a.v < b.v ? t : f; }
#undef ORACLE_DD_LOOP

}  // namespace oracle
