// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// A from-scratch stand-in for the subset of CppAD 20230000.0 + CppADCodeGen v2.4.3-ungar that the reference uses
// (call sites: include/ungar/autodiff/function.hpp:42-613, include/ungar/utils/utils.hpp, include/ungar/autodiff/
// support/quaternion.hpp, include/ungar/optimization/soft_inequality_constraint.hpp).  Neither library is in
// /root/reference nor in this image, so the reference's own headers and example sources are compiled — UNCHANGED, from
// where they lie — against this header instead (recipe: oracle/build_ref.py, outputs only under oracle/_ref/).
//
// What it provides: a tracing scalar CppAD::AD<CppAD::cg::CG<double>> that records a tape, and a
// CppAD::cg::GenericModel<double> that evaluates the tape (forward_zero / sparse Jacobian / sparse Hessian) on the CPU
// instead of generating and compiling C.  A "dynamic library" is a tape file.  Semantics restated from CppAD's
// documented behaviour: operations on parameters are folded; `variable * 0`, `variable + 0`, `variable * 1`,
// `variable / 1`, `0 / variable` are folded ("identical" rules); CondExp differentiates the selected branch and its
// sparsity is the union of both branches; abs'(0) = 0; pow(x, int) is a repeated product.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../ad.hpp"  // oracle::SDual — the same forward-mode duals the restated oracle uses

namespace CppAD {

namespace cg {
template <class Base>
class CG {  // the code-generation scalar: only its value survives here
  public:
    CG() = default;
    CG(Base v) : _v(v) {}  // NOLINT
    Base getValue() const { return _v; }
    bool isParameter() const { return true; }

  private:
    Base _v{};
};
}  // namespace cg

namespace shim {

enum Op : std::uint8_t {
    OP_INDEP, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SQRT, OP_SIN, OP_COS, OP_TAN, OP_ATAN, OP_ACOS, OP_ASIN,
    OP_EXP, OP_LOG, OP_ABS, OP_POW, OP_ATAN2, OP_CLT, OP_CLE, OP_CGT, OP_CGE, OP_CEQ
};

struct Node {
    std::uint8_t op;
    std::int32_t a, b, c, d;
    double k;
};

struct Tape {
    std::vector<Node> nodes;
    std::int32_t n_indep = 0;
    std::vector<std::int32_t> dep_id;  // node id, or -1 for a constant dependent
    std::vector<double> dep_const;
    std::int32_t push(std::uint8_t op, std::int32_t a = -1, std::int32_t b = -1, std::int32_t c = -1, std::int32_t d = -1,
                      double k = 0.0) {
        nodes.push_back(Node{op, a, b, c, d, k});
        return static_cast<std::int32_t>(nodes.size()) - 1;
    }
};

inline thread_local std::shared_ptr<Tape> g_recording;  // active between Independent() and ADFun's constructor

// Scalar-generic replay of a tape (S = double, oracle::Dual1, oracle::Dual2).
template <class S>
void replay(const Tape& t, const S* x, std::vector<S>& y) {
    using namespace oracle;
    std::vector<S> v(t.nodes.size());
    for (std::size_t i = 0; i < t.nodes.size(); ++i) {
        const Node& n = t.nodes[i];
        switch (n.op) {
            case OP_INDEP: v[i] = x[n.a]; break;
            case OP_CONST: v[i] = S(n.k); break;
            case OP_ADD: v[i] = v[n.a] + v[n.b]; break;
            case OP_SUB: v[i] = v[n.a] - v[n.b]; break;
            case OP_MUL: v[i] = v[n.a] * v[n.b]; break;
            case OP_DIV: v[i] = v[n.a] / v[n.b]; break;
            case OP_NEG: v[i] = -v[n.a]; break;
            case OP_SQRT: v[i] = ad_sqrt(v[n.a]); break;
            case OP_SIN: v[i] = ad_sin(v[n.a]); break;
            case OP_COS: v[i] = ad_cos(v[n.a]); break;
            case OP_ATAN: v[i] = ad_atan(v[n.a]); break;
            case OP_ABS: v[i] = ad_abs(v[n.a]); break;
            case OP_CLT: v[i] = cond_select(value_of(v[n.a]) < value_of(v[n.b]), v[n.c], v[n.d]); break;
            case OP_CLE: v[i] = cond_select(value_of(v[n.a]) <= value_of(v[n.b]), v[n.c], v[n.d]); break;
            case OP_CGT: v[i] = cond_select(value_of(v[n.a]) > value_of(v[n.b]), v[n.c], v[n.d]); break;
            case OP_CGE: v[i] = cond_select(value_of(v[n.a]) >= value_of(v[n.b]), v[n.c], v[n.d]); break;
            case OP_CEQ: v[i] = cond_select(value_of(v[n.a]) == value_of(v[n.b]), v[n.c], v[n.d]); break;
            default: throw std::runtime_error("refshim: tape op not supported by the replay (tan/acos/asin/exp/log/pow/atan2 "
                                              "do not occur in the three MPC models)");
        }
    }
    y.resize(t.dep_id.size());
    for (std::size_t i = 0; i < t.dep_id.size(); ++i) y[i] = t.dep_id[i] >= 0 ? v[t.dep_id[i]] : S(t.dep_const[i]);
}

using SparsitySets = std::vector<std::set<std::size_t>>;

struct LibraryImage {  // what a "dynamic library" file holds: the tape and the options of its model
    std::shared_ptr<Tape> tape;
    bool jac = false, hes = false;
    SparsitySets customJac, customHes;
};

inline void write_sets(std::ostream& f, const SparsitySets& s) {
    const std::uint64_t n = s.size();
    f.write(reinterpret_cast<const char*>(&n), 8);
    for (const auto& row : s) {
        const std::uint64_t m = row.size();
        f.write(reinterpret_cast<const char*>(&m), 8);
        for (std::size_t c : row) {
            const std::uint64_t cc = c;
            f.write(reinterpret_cast<const char*>(&cc), 8);
        }
    }
}
inline void read_sets(std::istream& f, SparsitySets& s) {
    std::uint64_t n = 0;
    f.read(reinterpret_cast<char*>(&n), 8);
    s.assign(n, {});
    for (auto& row : s) {
        std::uint64_t m = 0;
        f.read(reinterpret_cast<char*>(&m), 8);
        for (std::uint64_t e = 0; e < m; ++e) {
            std::uint64_t c = 0;
            f.read(reinterpret_cast<char*>(&c), 8);
            row.insert(static_cast<std::size_t>(c));
        }
    }
}

inline void save(const LibraryImage& img, const std::string& path) {
    const Tape& t = *img.tape;
    std::ofstream f(path, std::ios::binary);
    const std::uint64_t magic = 0x32455041545f4255ull, nn = t.nodes.size(), nd = t.dep_id.size();  // "UB_TAPE2"
    const std::int64_t ni = t.n_indep, flags = (img.jac ? 1 : 0) | (img.hes ? 2 : 0);
    f.write(reinterpret_cast<const char*>(&magic), 8);
    f.write(reinterpret_cast<const char*>(&nn), 8);
    f.write(reinterpret_cast<const char*>(&nd), 8);
    f.write(reinterpret_cast<const char*>(&ni), 8);
    f.write(reinterpret_cast<const char*>(&flags), 8);
    f.write(reinterpret_cast<const char*>(t.nodes.data()), nn * sizeof(Node));
    f.write(reinterpret_cast<const char*>(t.dep_id.data()), nd * sizeof(std::int32_t));
    f.write(reinterpret_cast<const char*>(t.dep_const.data()), nd * sizeof(double));
    write_sets(f, img.customJac);
    write_sets(f, img.customHes);
    if (!f) throw std::runtime_error("refshim: cannot write tape " + path);
}

inline LibraryImage load(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    std::uint64_t magic = 0, nn = 0, nd = 0;
    std::int64_t ni = 0, flags = 0;
    f.read(reinterpret_cast<char*>(&magic), 8);
    f.read(reinterpret_cast<char*>(&nn), 8);
    f.read(reinterpret_cast<char*>(&nd), 8);
    f.read(reinterpret_cast<char*>(&ni), 8);
    f.read(reinterpret_cast<char*>(&flags), 8);
    if (!f || magic != 0x32455041545f4255ull) throw std::runtime_error("refshim: not a tape file: " + path);
    LibraryImage img;
    img.tape = std::make_shared<Tape>();
    Tape& t = *img.tape;
    t.nodes.resize(nn);
    t.dep_id.resize(nd);
    t.dep_const.resize(nd);
    t.n_indep = static_cast<std::int32_t>(ni);
    img.jac = flags & 1;
    img.hes = flags & 2;
    f.read(reinterpret_cast<char*>(t.nodes.data()), nn * sizeof(Node));
    f.read(reinterpret_cast<char*>(t.dep_id.data()), nd * sizeof(std::int32_t));
    f.read(reinterpret_cast<char*>(t.dep_const.data()), nd * sizeof(double));
    read_sets(f, img.customJac);
    read_sets(f, img.customHes);
    if (!f) throw std::runtime_error("refshim: truncated tape file: " + path);
    return img;
}

}  // namespace shim

// ------------------------------------------------------------------------------------------------------------------
// The tracing scalar.
// ------------------------------------------------------------------------------------------------------------------
template <class Base>
class AD;

template <>
class AD<cg::CG<double>> {
  public:
    double v = 0.0;
    std::int32_t id = -1;  // tape node, or -1 for a parameter (constant)

    AD() = default;
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
    AD(T x) : v(static_cast<double>(x)) {}  // NOLINT
    AD(const cg::CG<double>& x) : v(x.getValue()) {}  // NOLINT
    AD(double value, std::int32_t node) : v(value), id(node) {}

    bool variable() const { return id >= 0; }
    AD& operator+=(const AD& o);
    AD& operator-=(const AD& o);
    AD& operator*=(const AD& o);
    AD& operator/=(const AD& o);
    AD operator-() const;
    AD operator+() const { return *this; }
    explicit operator double() const { return v; }
};

using ADCGD = AD<cg::CG<double>>;

namespace shim {
inline std::int32_t node_of(const ADCGD& x) {  // node id of an operand, materialising constants
    return x.variable() ? x.id : g_recording->push(OP_CONST, -1, -1, -1, -1, x.v);
}
inline ADCGD unary(std::uint8_t op, const ADCGD& x, double value) {
    if (!x.variable() || !g_recording) return ADCGD(value);
    return ADCGD(value, g_recording->push(op, x.id));
}
inline ADCGD binary(std::uint8_t op, const ADCGD& a, const ADCGD& b, double value) {
    if ((!a.variable() && !b.variable()) || !g_recording) return ADCGD(value);
    const std::int32_t ia = node_of(a), ib = node_of(b);
    return ADCGD(value, g_recording->push(op, ia, ib));
}
}  // namespace shim

// Binary operators with CppAD's "identical" folding rules.
inline ADCGD operator+(const ADCGD& a, const ADCGD& b) {
    if (a.variable() && !b.variable() && b.v == 0.0) return a;
    if (b.variable() && !a.variable() && a.v == 0.0) return b;
    return shim::binary(shim::OP_ADD, a, b, a.v + b.v);
}
inline ADCGD operator-(const ADCGD& a, const ADCGD& b) {
    if (a.variable() && !b.variable() && b.v == 0.0) return a;
    return shim::binary(shim::OP_SUB, a, b, a.v - b.v);
}
inline ADCGD operator*(const ADCGD& a, const ADCGD& b) {
    if (a.variable() && !b.variable()) {
        if (b.v == 0.0) return ADCGD(0.0);
        if (b.v == 1.0) return a;
    }
    if (b.variable() && !a.variable()) {
        if (a.v == 0.0) return ADCGD(0.0);
        if (a.v == 1.0) return b;
    }
    return shim::binary(shim::OP_MUL, a, b, a.v * b.v);
}
inline ADCGD operator/(const ADCGD& a, const ADCGD& b) {
    if (a.variable() && !b.variable() && b.v == 1.0) return a;
    if (b.variable() && !a.variable() && a.v == 0.0) return ADCGD(0.0);
    return shim::binary(shim::OP_DIV, a, b, a.v / b.v);
}
inline ADCGD ADCGD::operator-() const { return shim::unary(shim::OP_NEG, *this, -v); }
inline ADCGD& ADCGD::operator+=(const ADCGD& o) { return *this = *this + o; }
inline ADCGD& ADCGD::operator-=(const ADCGD& o) { return *this = *this - o; }
inline ADCGD& ADCGD::operator*=(const ADCGD& o) { return *this = *this * o; }
inline ADCGD& ADCGD::operator/=(const ADCGD& o) { return *this = *this / o; }

#define UB_SHIM_MIXED(op)                                                                                   \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                              \
    inline ADCGD operator op(const ADCGD& a, T b) { return a op ADCGD(b); }                                 \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                              \
    inline ADCGD operator op(T a, const ADCGD& b) { return ADCGD(a) op b; }
UB_SHIM_MIXED(+)
UB_SHIM_MIXED(-)
UB_SHIM_MIXED(*)
UB_SHIM_MIXED(/)
#undef UB_SHIM_MIXED

// Comparisons act on values (the reference records with "no_compare_op", function.hpp:466).
#define UB_SHIM_CMP(op)                                                                                     \
    inline bool operator op(const ADCGD& a, const ADCGD& b) { return a.v op b.v; }                          \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                              \
    inline bool operator op(const ADCGD& a, T b) { return a.v op static_cast<double>(b); }                  \
    template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>                              \
    inline bool operator op(T a, const ADCGD& b) { return static_cast<double>(a) op b.v; }
UB_SHIM_CMP(<)
UB_SHIM_CMP(<=)
UB_SHIM_CMP(>)
UB_SHIM_CMP(>=)
UB_SHIM_CMP(==)
UB_SHIM_CMP(!=)
#undef UB_SHIM_CMP

// CppAD also defines its math functions for the base types (utils.hpp:958-966 calls CppAD::atan2 / sqrt on doubles).
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline double abs(double x) { return std::fabs(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double atan(double x) { return std::atan(x); }
inline double log(double x) { return std::log(x); }
inline double exp(double x) { return std::exp(x); }
inline double pow(double x, double y) { return std::pow(x, y); }
inline double pow(double x, int n) { return std::pow(x, n); }

inline ADCGD sqrt(const ADCGD& x) { return shim::unary(shim::OP_SQRT, x, std::sqrt(x.v)); }
inline ADCGD sin(const ADCGD& x) { return shim::unary(shim::OP_SIN, x, std::sin(x.v)); }
inline ADCGD cos(const ADCGD& x) { return shim::unary(shim::OP_COS, x, std::cos(x.v)); }
inline ADCGD tan(const ADCGD& x) { return shim::unary(shim::OP_TAN, x, std::tan(x.v)); }
inline ADCGD atan(const ADCGD& x) { return shim::unary(shim::OP_ATAN, x, std::atan(x.v)); }
inline ADCGD acos(const ADCGD& x) { return shim::unary(shim::OP_ACOS, x, std::acos(x.v)); }
inline ADCGD asin(const ADCGD& x) { return shim::unary(shim::OP_ASIN, x, std::asin(x.v)); }
inline ADCGD exp(const ADCGD& x) { return shim::unary(shim::OP_EXP, x, std::exp(x.v)); }
inline ADCGD log(const ADCGD& x) { return shim::unary(shim::OP_LOG, x, std::log(x.v)); }
inline ADCGD abs(const ADCGD& x) { return shim::unary(shim::OP_ABS, x, std::fabs(x.v)); }
inline ADCGD fabs(const ADCGD& x) { return abs(x); }
inline ADCGD atan2(const ADCGD& y, const ADCGD& x) { return shim::binary(shim::OP_ATAN2, y, x, std::atan2(y.v, x.v)); }
inline ADCGD pow(const ADCGD& x, const ADCGD& y) { return shim::binary(shim::OP_POW, x, y, std::pow(x.v, y.v)); }
inline ADCGD pow(const ADCGD& x, double y) { return pow(x, ADCGD(y)); }
inline ADCGD pow(double x, const ADCGD& y) { return pow(ADCGD(x), y); }
inline ADCGD pow(const ADCGD& x, int n) {  // CppAD: integer powers by repeated multiplication
    if (n < 0) return ADCGD(1.0) / pow(x, -n);
    ADCGD p(1.0);
    for (int i = 0; i < n; ++i) p = p * x;
    return p;
}
inline bool isfinite(const ADCGD& x) { return std::isfinite(x.v); }
inline bool isnan(const ADCGD& x) { return std::isnan(x.v); }
inline bool isinf(const ADCGD& x) { return std::isinf(x.v); }
inline ADCGD conj(const ADCGD& x) { return x; }
inline ADCGD real(const ADCGD& x) { return x; }
inline ADCGD imag(const ADCGD&) { return ADCGD(0.0); }
inline ADCGD abs2(const ADCGD& x) { return x * x; }
inline cg::CG<double> Value(const ADCGD& x) {
    if (x.variable() && shim::g_recording) throw std::runtime_error("refshim: Value() of a variable");
    return cg::CG<double>(x.v);
}

namespace shim {
inline ADCGD cond(std::uint8_t op, bool take_t, const ADCGD& a, const ADCGD& b, const ADCGD& t, const ADCGD& f) {
    const ADCGD& sel = take_t ? t : f;
    if (!g_recording || (!a.variable() && !b.variable())) return sel;  // decided by parameters: no operation
    if (!t.variable() && !f.variable() && t.v == f.v) return sel;
    const std::int32_t ia = node_of(a), ib = node_of(b), it = node_of(t), jf = node_of(f);
    return ADCGD(sel.v, g_recording->push(op, ia, ib, it, jf));
}
}  // namespace shim
inline ADCGD CondExpLt(const ADCGD& a, const ADCGD& b, const ADCGD& t, const ADCGD& f) { return shim::cond(shim::OP_CLT, a.v < b.v, a, b, t, f); }
inline ADCGD CondExpLe(const ADCGD& a, const ADCGD& b, const ADCGD& t, const ADCGD& f) { return shim::cond(shim::OP_CLE, a.v <= b.v, a, b, t, f); }
inline ADCGD CondExpGt(const ADCGD& a, const ADCGD& b, const ADCGD& t, const ADCGD& f) { return shim::cond(shim::OP_CGT, a.v > b.v, a, b, t, f); }
inline ADCGD CondExpGe(const ADCGD& a, const ADCGD& b, const ADCGD& t, const ADCGD& f) { return shim::cond(shim::OP_CGE, a.v >= b.v, a, b, t, f); }
inline ADCGD CondExpEq(const ADCGD& a, const ADCGD& b, const ADCGD& t, const ADCGD& f) { return shim::cond(shim::OP_CEQ, a.v == b.v, a, b, t, f); }

// CppAD::Independent(x): start recording with x as the independent variables (function.hpp:456-458).
template <class Vector>
void Independent(Vector& x) {
    shim::g_recording = std::make_shared<shim::Tape>();
    shim::g_recording->n_indep = static_cast<std::int32_t>(x.size());
    for (std::int32_t i = 0; i < static_cast<std::int32_t>(x.size()); ++i) {
        const std::int32_t id = shim::g_recording->push(shim::OP_INDEP, i);
        x[i] = ADCGD(x[i].v, id);
    }
}

// CppAD::ADFun<Base>(x, y): stop recording; y are the dependents (function.hpp:465).
template <class Base>
class ADFun {
  public:
    template <class VectorX, class VectorY>
    ADFun(const VectorX& x, const VectorY& y) : tape(shim::g_recording) {
        if (!tape) throw std::runtime_error("refshim: ADFun without Independent()");
        (void)x;
        for (std::int64_t i = 0; i < static_cast<std::int64_t>(y.size()); ++i) {
            tape->dep_id.push_back(y[i].variable() ? y[i].id : -1);
            tape->dep_const.push_back(y[i].v);
        }
        shim::g_recording.reset();
    }
    void optimize(const std::string& = "") {}
    std::shared_ptr<shim::Tape> tape;
};

// ------------------------------------------------------------------------------------------------------------------
// CppADCodeGen subset: model sources, "compiler", dynamic library = tape file, GenericModel = tape evaluator.
// ------------------------------------------------------------------------------------------------------------------
namespace cg {

template <class T>
class ArrayView {
  public:
    ArrayView(T* data, std::size_t size) : _d(data), _n(size) {}
    template <class U>
    ArrayView(std::vector<U>& v) : _d(v.data()), _n(v.size()) {}  // NOLINT
    T* data() const { return _d; }
    std::size_t size() const { return _n; }
    T& operator[](std::size_t i) const { return _d[i]; }

  private:
    T* _d;
    std::size_t _n;
};

namespace system {
template <class = void>
struct SystemInfo {
    static inline const std::string DYNAMIC_LIB_EXTENSION = ".so";
};
}  // namespace system

using shim::SparsitySets;

template <class Base>
class GenericModel {
  public:
    GenericModel(std::string name, std::shared_ptr<shim::Tape> tape, bool jac, bool hes, SparsitySets customJac,
                 SparsitySets customHes)
        : _name(std::move(name)), _tape(std::move(tape)), _jac(jac), _hes(hes), _customJac(std::move(customJac)),
          _customHes(std::move(customHes)) {
        if (_jac) {
            const SparsitySets use = JacobianSparsitySet();
            for (std::size_t r = 0; r < use.size(); ++r)
                for (std::size_t c : use[r]) { _jrows.push_back(r); _jcols.push_back(c); }
        }
        if (_hes) {
            const SparsitySets use = HessianSparsitySet();
            for (std::size_t r = 0; r < use.size(); ++r)
                for (std::size_t c : use[r]) { _hrows.push_back(r); _hcols.push_back(c); }
        }
    }
    const std::string& getName() const { return _name; }
    bool isForwardZeroAvailable() const { return true; }
    bool isSparseJacobianAvailable() const { return _jac; }
    bool isSparseHessianAvailable() const { return _hes; }
    bool isJacobianSparsityAvailable() const { return _jac; }
    bool isHessianSparsityAvailable() const { return _hes; }
    std::size_t Domain() const { return _tape->n_indep; }
    std::size_t Range() const { return _tape->dep_id.size(); }

    // Structural patterns over ALL independents [x; p] — the reference trims the parameter columns itself through
    // setCustomSparse*Elements (function.hpp:529-574), after which the model reports the custom pattern.
    SparsitySets JacobianSparsitySet() const {
        if (!_customJac.empty()) return _customJac;
        std::vector<oracle::Dual1> y;
        replay1(generic_point(), y);
        SparsitySets s(y.size());
        for (std::size_t r = 0; r < y.size(); ++r)
            for (const auto& e : y[r].d) s[r].insert(static_cast<std::size_t>(e.first));
        return s;
    }
    SparsitySets HessianSparsitySet() const {  // union over the dependents (they are scalar functions in Ungar)
        if (!_customHes.empty()) return _customHes;
        std::vector<oracle::Dual2> y;
        replay2(generic_point(), y);
        SparsitySets s(_tape->n_indep);
        for (const auto& yi : y)
            for (const auto& ri : yi.d)
                for (const auto& cj : ri.second.d) s[ri.first].insert(static_cast<std::size_t>(cj.first));
        return s;
    }
    void JacobianSparsity(std::vector<std::size_t>& rows, std::vector<std::size_t>& cols) const { rows = _jrows; cols = _jcols; }
    void HessianSparsity(std::size_t, std::vector<std::size_t>& rows, std::vector<std::size_t>& cols) const { rows = _hrows; cols = _hcols; }

    void ForwardZero(ArrayView<const Base> x, ArrayView<Base> y) const {
        std::vector<double> out;
        shim::replay<double>(*_tape, x.data(), out);
        for (std::size_t i = 0; i < out.size(); ++i) y[i] = out[i];
    }
    void SparseJacobian(ArrayView<const Base> x, ArrayView<Base> jac, const std::size_t** rows, const std::size_t** cols) const {
        std::vector<oracle::Dual1> y;
        replay1(x.data(), y);
        for (std::size_t e = 0; e < _jrows.size(); ++e) jac[e] = lookup(y[_jrows[e]].d, static_cast<int>(_jcols[e]));
        *rows = _jrows.data();
        *cols = _jcols.data();
    }
    void SparseHessian(ArrayView<const Base> x, ArrayView<const Base> w, ArrayView<Base> hess, const std::size_t** rows,
                       const std::size_t** cols) const {
        std::vector<oracle::Dual2> y;
        replay2(x.data(), y);
        for (std::size_t e = 0; e < _hrows.size(); ++e) {
            double acc = 0.0;
            for (std::size_t i = 0; i < y.size(); ++i) {
                if (w[i] == 0.0) continue;
                for (const auto& ri : y[i].d)
                    if (ri.first == static_cast<int>(_hrows[e])) acc += w[i] * lookup(ri.second.d, static_cast<int>(_hcols[e]));
            }
            hess[e] = acc;
        }
        *rows = _hrows.data();
        *cols = _hcols.data();
    }
    const std::shared_ptr<shim::Tape>& tape() const { return _tape; }

  private:
    static double lookup(const std::vector<std::pair<int, double>>& d, int col) {
        for (const auto& e : d)
            if (e.first == col) return e.second;
        return 0.0;
    }
    const double* generic_point() const {
        if (_generic.empty()) {
            _generic.resize(_tape->n_indep);
            for (std::size_t i = 0; i < _generic.size(); ++i) _generic[i] = 0.37 + 0.001 * static_cast<double>(i % 97);
        }
        return _generic.data();
    }
    void replay1(const double* x, std::vector<oracle::Dual1>& y) const {
        std::vector<oracle::Dual1> in(_tape->n_indep);
        for (std::int32_t i = 0; i < _tape->n_indep; ++i) in[i] = oracle::seed1(x[i], i);
        shim::replay(*_tape, in.data(), y);
    }
    void replay2(const double* x, std::vector<oracle::Dual2>& y) const {
        std::vector<oracle::Dual2> in(_tape->n_indep);
        for (std::int32_t i = 0; i < _tape->n_indep; ++i) in[i] = oracle::seed2(x[i], i);
        shim::replay(*_tape, in.data(), y);
    }
    std::string _name;
    std::shared_ptr<shim::Tape> _tape;
    bool _jac, _hes;
    SparsitySets _customJac, _customHes;
    std::vector<std::size_t> _jrows, _jcols, _hrows, _hcols;
    mutable std::vector<double> _generic;
};

template <class Base>
class ModelCSourceGen {
  public:
    template <class ADFunT>
    ModelCSourceGen(ADFunT& fun, std::string name) : tape(fun.tape), name(std::move(name)) {}
    void setCreateSparseJacobian(bool b) { jac = b; }
    void setCreateSparseHessian(bool b) { hes = b; }
    void setCreateHessianSparsityByEquation(bool) {}
    template <class Pattern>
    void setCustomSparseJacobianElements(const Pattern& p) { customJac = convert(p); }
    template <class Pattern>
    void setCustomSparseHessianElements(const Pattern& p) { customHes = convert(p); }
    std::shared_ptr<shim::Tape> tape;
    std::string name;
    bool jac = false, hes = false;
    SparsitySets customJac, customHes;

  private:
    template <class Pattern>
    static SparsitySets convert(const Pattern& p) {
        SparsitySets s(p.size());
        for (std::size_t r = 0; r < p.size(); ++r) s[r].insert(p[r].begin(), p[r].end());
        return s;
    }
};

template <class Base>
class ModelLibraryCSourceGen {
  public:
    explicit ModelLibraryCSourceGen(ModelCSourceGen<Base>& m) : model(&m) {}
    ModelCSourceGen<Base>* model;
};

template <class Base>
class GccCompiler {
  public:
    void setCompileLibFlags(const std::vector<std::string>&) {}
    void addCompileLibFlag(const std::string&) {}
    void setTemporaryFolder(const std::filesystem::path&) {}
    void setSourcesFolder(const std::filesystem::path&) {}
    void setSaveToDiskFirst(bool) {}
};

template <class Base>
class DynamicLib {
  public:
    virtual ~DynamicLib() = default;
    std::unique_ptr<GenericModel<Base>> model(const std::string& name) const {
        return std::make_unique<GenericModel<Base>>(name, tape, jac, hes, customJac, customHes);
    }
    std::shared_ptr<shim::Tape> tape;
    bool jac = false, hes = false;
    SparsitySets customJac, customHes;
};

// The "shared library" is one file: the tape plus the options of its model.
template <class Base>
class LinuxDynamicLib : public DynamicLib<Base> {
  public:
    explicit LinuxDynamicLib(const std::filesystem::path& path) {
        shim::LibraryImage img = shim::load(path.string());
        this->tape = img.tape; this->jac = img.jac; this->hes = img.hes;
        this->customJac = std::move(img.customJac); this->customHes = std::move(img.customHes);
    }
};

template <class Base>
class DynamicModelLibraryProcessor {
  public:
    DynamicModelLibraryProcessor(ModelLibraryCSourceGen<Base>& lib, std::string nameWithoutExtension)
        : _lib(&lib), _path(std::move(nameWithoutExtension) + system::SystemInfo<>::DYNAMIC_LIB_EXTENSION) {}
    std::unique_ptr<DynamicLib<Base>> createDynamicLibrary(GccCompiler<Base>&) {
        const ModelCSourceGen<Base>& m = *_lib->model;
        shim::save(shim::LibraryImage{m.tape, m.jac, m.hes, m.customJac, m.customHes}, _path);
        auto lib = std::make_unique<DynamicLib<Base>>();
        lib->tape = m.tape; lib->jac = m.jac; lib->hes = m.hes; lib->customJac = m.customJac; lib->customHes = m.customHes;
        return lib;
    }

  private:
    ModelLibraryCSourceGen<Base>* _lib;
    std::string _path;
};

}  // namespace cg
}  // namespace CppAD

namespace std {
template <>
class numeric_limits<CppAD::ADCGD> : public numeric_limits<double> {
  public:
    static CppAD::ADCGD epsilon() { return numeric_limits<double>::epsilon(); }
    static CppAD::ADCGD min() { return numeric_limits<double>::min(); }
    static CppAD::ADCGD max() { return numeric_limits<double>::max(); }
    static CppAD::ADCGD lowest() { return numeric_limits<double>::lowest(); }
    static CppAD::ADCGD quiet_NaN() { return numeric_limits<double>::quiet_NaN(); }
    static CppAD::ADCGD infinity() { return numeric_limits<double>::infinity(); }
};
}  // namespace std
