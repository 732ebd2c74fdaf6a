// ORACLE — TEST INFRASTRUCTURE ONLY (see ad.hpp header).
//
// C interface (loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's CPU legs) over
// the scalar-generic restatement in models.hpp:
//   * oracle_eval / oracle_jacobian / oracle_hessian  restate Function::Evaluate / Jacobian / Hessian
//     (include/ungar/autodiff/function.hpp:180-274): y = f([x; p]); dy/dx with the parameter columns
//     trimmed (:529-550); upper-triangular d2y/dx2 of a scalar function (:552-574).
//   * oracle_kkt_record restates SoftSQPOptimizer::AssembleOSQPInstance
//     (include/ungar/optimization/soft_sqp.hpp:141-158, 245-264) on the monolithic sparse matrices and
//     then cuts the result into the per-shooting-node block record documented in DESIGN.md §3; it
//     fails (returns < 0) if any structural nonzero falls outside the block set, i.e. it proves the
//     block decomposition lossless.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "models.hpp"

namespace oracle {

enum Model { QUADROTOR = 0, RC_CAR = 1, QUADRUPED = 2 };
enum Fn { OBJECTIVE = 0, EQUALITIES = 1, INEQUALITIES = 2 };

static bool get_sizes(int model, int N, Sizes& s) {
    if (N < 1) return false;
    switch (model) {
        case QUADROTOR: s = QuadrotorLayout{N}.sizes(); return true;
        case RC_CAR: s = RcCarLayout{N}.sizes(); return true;
        case QUADRUPED: s = QuadrupedLayout{N}.sizes(); return true;
        default: return false;
    }
}

template <class S>
static bool eval_fn(int model, int fn, int N, const S* xp, std::vector<S>& y) {
    switch (model * 3 + fn) {
        case QUADROTOR * 3 + OBJECTIVE: quadrotor_objective(N, xp, y); return true;
        case QUADROTOR * 3 + EQUALITIES: quadrotor_equalities(N, xp, y); return true;
        case QUADROTOR * 3 + INEQUALITIES: quadrotor_inequalities(N, xp, y); return true;
        case RC_CAR * 3 + OBJECTIVE: rc_car_objective(N, xp, y); return true;
        case RC_CAR * 3 + EQUALITIES: rc_car_equalities(N, xp, y); return true;
        case RC_CAR * 3 + INEQUALITIES: rc_car_inequalities(N, xp, y); return true;
        case QUADRUPED * 3 + OBJECTIVE: quadruped_objective(N, xp, y); return true;
        case QUADRUPED * 3 + EQUALITIES: quadruped_equalities(N, xp, y); return true;
        case QUADRUPED * 3 + INEQUALITIES: quadruped_inequalities(N, xp, y); return true;
        default: return false;
    }
}

// [x; p] with the decision variables seeded (function.hpp:456-458 declares the whole vector
// independent and then trims the parameter columns, :529-550; seeding only x is equivalent).
static std::vector<Dual1> seeded1(const Sizes& s, const double* xp) {
    std::vector<Dual1> v(s.n_dec + s.n_par);
    for (int i = 0; i < s.n_dec; ++i) v[i] = seed1(xp[i], i);
    for (int i = s.n_dec; i < s.n_dec + s.n_par; ++i) v[i] = Dual1(xp[i]);
    return v;
}
static std::vector<Dual2> seeded2(const Sizes& s, const double* xp) {
    std::vector<Dual2> v(s.n_dec + s.n_par);
    for (int i = 0; i < s.n_dec; ++i) v[i] = seed2(xp[i], i);
    for (int i = s.n_dec; i < s.n_dec + s.n_par; ++i) v[i] = Dual2(Dual1(xp[i]));
    return v;
}

struct Triplets {
    std::vector<int> rows, cols;
    std::vector<double> vals;
};

static bool jacobian(int model, int fn, int N, const double* xp, std::vector<double>* y, Triplets& t) {
    Sizes s;
    if (!get_sizes(model, N, s)) return false;
    const auto v = seeded1(s, xp);
    std::vector<Dual1> out;
    if (!eval_fn(model, fn, N, v.data(), out)) return false;
    if (y) y->clear();
    for (std::size_t r = 0; r < out.size(); ++r) {
        if (y) y->push_back(out[r].v);
        for (const auto& e : out[r].d) {  // already sorted by column
            t.rows.push_back(static_cast<int>(r));
            t.cols.push_back(e.first);
            t.vals.push_back(e.second);
        }
    }
    return true;
}

// Upper-triangular Hessian of the scalar objective (function.hpp:232-258, :563-571).
static bool hessian(int model, int N, const double* xp, double* value, std::vector<double>* grad, Triplets& t) {
    Sizes s;
    if (!get_sizes(model, N, s)) return false;
    const auto v = seeded2(s, xp);
    std::vector<Dual2> out;
    if (!eval_fn(model, OBJECTIVE, N, v.data(), out) || out.size() != 1) return false;
    if (value) *value = out[0].v.v;
    if (grad) {
        grad->assign(s.n_dec, 0.0);
        for (const auto& e : out[0].v.d) (*grad)[e.first] = e.second;
    }
    for (const auto& ri : out[0].d)
        for (const auto& cj : ri.second.d)
            if (cj.first >= ri.first) {
                t.rows.push_back(ri.first);
                t.cols.push_back(cj.first);
                t.vals.push_back(cj.second);
            }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Block record layout (element offsets; every array starts on a multiple of 4 elements).
// ---------------------------------------------------------------------------------------------
struct Record {
    int g, A, C, h, cost, grad, H, HN, Hc, size;  // offsets
    int nz, tri, ntri_N, n_legs, hc_per_node;
};

static int round4(int x) { return (x + 3) & ~3; }

static Record make_record(int model, const Sizes& s) {
    Record r{};
    r.nz          = s.nx + s.nu;
    r.tri         = r.nz * (r.nz + 1) / 2;
    r.ntri_N      = s.nx * (s.nx + 1) / 2;
    r.n_legs      = model == QUADRUPED ? 4 : 0;
    r.hc_per_node = model == QUADRUPED ? 0 : s.nu;
    int off       = 0;
    r.g = off;    off = round4(off + s.m_eq);
    r.A = off;    off = round4(off + s.N * s.nx * r.nz);
    r.C = off;    off = round4(off + s.N * r.n_legs * 4 * 20);
    r.h = off;    off = round4(off + s.m_ineq);
    r.cost = off; off = round4(off + 2);
    r.grad = off; off = round4(off + s.n_dec);
    r.H = off;    off = round4(off + s.N * r.tri);
    r.HN = off;   off = round4(off + r.ntri_N);
    r.Hc = off;   off = round4(off + (s.N - 1) * r.hc_per_node);
    r.size = off;
    return r;
}

static int tri_index(int n, int i, int j) { return i * n - i * (i - 1) / 2 + (j - i); }  // i <= j

// Decision-variable index -> (node, local index within z_k = [x_k; u_k]); x_N is node N.
static void locate(const Sizes& s, int idx, int& node, int& local) {
    const int nX = s.nx * (s.N + 1);
    if (idx < nX) {
        node  = idx / s.nx;
        local = idx % s.nx;
    } else {
        node  = (idx - nX) / s.nu;
        local = s.nx + (idx - nX) % s.nu;
    }
}

// Returns 0 on success, a negative code naming the first structural nonzero that does not fit.
static int kkt_record(int model, int N, const double* xp, double stiffness, double epsilon, double* rec) {
    Sizes s;
    if (!get_sizes(model, N, s)) return -1;
    const Record R = make_record(model, s);
    std::fill(rec, rec + R.size, 0.0);

    // --- pieces, exactly the calls of soft_sqp.hpp:143-158 ---------------------------------
    double f = 0.0;
    std::vector<double> grad_f;
    Triplets Hf;
    if (!hessian(model, N, xp, &f, &grad_f, Hf)) return -1;
    std::vector<double> g, h;
    Triplets Jg, Jh;
    if (!jacobian(model, EQUALITIES, N, xp, &g, Jg)) return -1;
    if (!jacobian(model, INEQUALITIES, N, xp, &h, Jh)) return -1;

    // Barrier value / Jacobian / (diagonal) Hessian at Zineq = h (soft_sqp.hpp:236-264).
    const PolyBarrier barrier(stiffness, epsilon);
    std::vector<Dual2> hz(h.size());
    for (std::size_t i = 0; i < h.size(); ++i) hz[i] = seed2(h[i], static_cast<int>(i));
    const Dual2 Z = barrier.soft_constraint(hz);
    std::vector<double> dB(h.size(), 0.0), d2B(h.size(), 0.0);
    for (const auto& e : Z.v.d) dB[e.first] = e.second;
    for (const auto& ri : Z.d)
        for (const auto& cj : ri.second.d) {
            if (cj.first != ri.first && cj.second != 0.0) return -2;  // barrier Hessian must be diagonal
            if (cj.first == ri.first) d2B[ri.first] = cj.second;
        }

    // --- P = triu(Hf) + Jh^T diag(d2B) Jh + 1e-6 I ;  q = grad f + Jh^T dB -----------------
    std::map<std::pair<int, int>, double> P;
    for (std::size_t e = 0; e < Hf.vals.size(); ++e) P[{Hf.rows[e], Hf.cols[e]}] += Hf.vals[e];
    std::vector<double> q = grad_f;
    {
        std::size_t e = 0;
        while (e < Jh.vals.size()) {
            std::size_t e1 = e;
            while (e1 < Jh.vals.size() && Jh.rows[e1] == Jh.rows[e]) ++e1;
            const int row = Jh.rows[e];
            for (std::size_t a = e; a < e1; ++a) {
                q[Jh.cols[a]] += dB[row] * Jh.vals[a];
                for (std::size_t b = e; b < e1; ++b)
                    if (Jh.cols[b] >= Jh.cols[a])
                        P[{Jh.cols[a], Jh.cols[b]}] += d2B[row] * Jh.vals[a] * Jh.vals[b];
            }
            e = e1;
        }
    }
    for (int i = 0; i < s.n_dec; ++i) P[{i, i}] += 1e-6;

    // --- cut into the record ------------------------------------------------------------------
    std::copy(g.begin(), g.end(), rec + R.g);
    std::copy(h.begin(), h.end(), rec + R.h);
    rec[R.cost]     = f;
    rec[R.cost + 1] = Z.v.v;
    std::copy(q.begin(), q.end(), rec + R.grad);

    for (const auto& kv : P) {
        int ni, li, nj, lj;
        locate(s, kv.first.first, ni, li);
        locate(s, kv.first.second, nj, lj);
        if (ni == nj && ni < N && li <= lj) {
            rec[R.H + ni * R.tri + tri_index(R.nz, li, lj)] = kv.second;
        } else if (ni == N && nj == N && li <= lj) {
            rec[R.HN + tri_index(s.nx, li, lj)] = kv.second;
        } else if (R.hc_per_node && nj == ni + 1 && li >= s.nx && lj == li) {
            rec[R.Hc + ni * s.nu + (li - s.nx)] = kv.second;
        } else {
            return -3;
        }
    }

    const int nX = s.nx * (N + 1);
    for (std::size_t e = 0; e < Jg.vals.size(); ++e) {
        const int row = Jg.rows[e], col = Jg.cols[e];
        const double val = Jg.vals[e];
        if (row < s.nx) {  // x_0 - x_measured
            if (col != row || val != 1.0) return -4;
        } else if (row < s.nx + s.nx * N) {  // x_{k+1} - f(x_k, u_k)
            const int k = (row - s.nx) / s.nx, r = (row - s.nx) % s.nx;
            int node, local;
            locate(s, col, node, local);
            if (node == k + 1 && col < nX) {
                if (local != r || val != 1.0) return -5;
            } else if (node == k) {
                rec[R.A + (k * s.nx + r) * R.nz + local] = val;
            } else {
                return -6;
            }
        } else {  // quadruped contact rows
            if (model != QUADRUPED) return -7;
            const QuadrupedLayout L{N};
            const int rr = row - (s.nx + s.nx * N);
            const int k = rr / 16, leg = (rr % 16) / 4, comp = rr % 4;
            double* Crow = rec + R.C + ((k * 4 + leg) * 4 + comp) * 20;
            if (col >= L.X(k) && col < L.X(k) + 7) Crow[col - L.X(k)] = val;
            else if (col >= L.R(k, leg) && col < L.R(k, leg) + 3) Crow[7 + col - L.R(k, leg)] = val;
            else if (k && col >= L.X(k - 1) && col < L.X(k - 1) + 7) Crow[10 + col - L.X(k - 1)] = val;
            else if (k && col >= L.R(k - 1, leg) && col < L.R(k - 1, leg) + 3) Crow[17 + col - L.R(k - 1, leg)] = val;
            else return -8;
        }
    }
    return 0;
}

}  // namespace oracle

// =============================================================================================
// C interface
// =============================================================================================
using namespace oracle;

static int64_t emit(const Triplets& t, int64_t cap, int* rows, int* cols, double* vals) {
    const int64_t nnz = static_cast<int64_t>(t.vals.size());
    if (nnz <= cap && rows && cols && vals) {
        std::copy(t.rows.begin(), t.rows.end(), rows);
        std::copy(t.cols.begin(), t.cols.end(), cols);
        std::copy(t.vals.begin(), t.vals.end(), vals);
    }
    return nnz;
}

extern "C" {

// sizes[7] = nx, nu, N, n_dec, n_par, m_eq, m_ineq
int oracle_sizes(int model, int N, int* sizes) {
    Sizes s;
    if (!get_sizes(model, N, s)) return -1;
    const int v[7] = {s.nx, s.nu, s.N, s.n_dec, s.n_par, s.m_eq, s.m_ineq};
    std::memcpy(sizes, v, sizeof(v));
    return 0;
}

// layout[15] = offsets g A C h cost grad H HN Hc, size, nz, tri, ntri_N, n_legs, hc_per_node
int oracle_record_layout(int model, int N, int* layout) {
    Sizes s;
    if (!get_sizes(model, N, s)) return -1;
    const Record r = make_record(model, s);
    const int v[15] = {r.g, r.A, r.C, r.h, r.cost, r.grad, r.H, r.HN, r.Hc, r.size,
                       r.nz, r.tri, r.ntri_N, r.n_legs, r.hc_per_node};
    std::memcpy(layout, v, sizeof(v));
    return 0;
}

// y = f(xp); returns ny (or < 0).
int oracle_eval(int model, int fn, int N, const double* xp, double* y) {
    std::vector<double> out;
    if (!eval_fn<double>(model, fn, N, xp, out)) return -1;
    std::copy(out.begin(), out.end(), y);
    return static_cast<int>(out.size());
}

// Row-major, columns ascending within a row.  Returns nnz; writes only if nnz <= cap.
int64_t oracle_jacobian(int model, int fn, int N, const double* xp, int64_t cap, int* rows, int* cols, double* vals) {
    Triplets t;
    if (!jacobian(model, fn, N, xp, nullptr, t)) return -1;
    return emit(t, cap, rows, cols, vals);
}

int64_t oracle_hessian(int model, int N, const double* xp, int64_t cap, int* rows, int* cols, double* vals) {
    Triplets t;
    if (!hessian(model, N, xp, nullptr, nullptr, t)) return -1;
    return emit(t, cap, rows, cols, vals);
}

// Barrier Zsoft(z) = sum b(-z_i): value, dZ/dz_i, d2Z/dz_i^2.
int oracle_barrier(double stiffness, double epsilon, int n, const double* z, double* value, double* dz, double* d2z) {
    const PolyBarrier barrier(stiffness, epsilon);
    std::vector<Dual2> v(n);
    for (int i = 0; i < n; ++i) v[i] = seed2(z[i], i);
    const Dual2 Z = barrier.soft_constraint(v);
    *value = Z.v.v;
    std::fill(dz, dz + n, 0.0);
    std::fill(d2z, d2z + n, 0.0);
    for (const auto& e : Z.v.d) dz[e.first] = e.second;
    for (const auto& ri : Z.d)
        for (const auto& cj : ri.second.d)
            if (cj.first == ri.first) d2z[ri.first] = cj.second;
    return 0;
}

int oracle_kkt_record(int model, int N, const double* xp, double stiffness, double epsilon, double* record) {
    return kkt_record(model, N, xp, stiffness, epsilon, record);
}

// One discrete-time step x+ = f(x, u; parameters of xp), node k (quadruped uses p_k).
int oracle_dynamics(int model, int N, const double* xp, int k, double* xnext) {
    switch (model) {
        case QUADROTOR: { const QuadrotorLayout L{N}; quadrotor_dynamics<double>(L, xp, xp + L.X(k), xp + L.U(k), xnext); return 13; }
        case RC_CAR: { const RcCarLayout L{N}; rc_car_dynamics<double>(L, xp, xp + L.X(k), xp + L.U(k), xnext); return 6; }
        case QUADRUPED: { const QuadrupedLayout L{N}; quadruped_dynamics<double>(L, xp, k, xnext); return 13; }
        default: return -1;
    }
}

// Utils::ApproximateExponentialMap(v).coeffs() and its 4x3 Jacobian (row-major) — the function pinned by
// test/autodiff/function.test.cpp:33-59.
int oracle_approx_exp(const double* v, double* q, double* jac) {
    const V3<Dual1> x{seed1(v[0], 0), seed1(v[1], 1), seed1(v[2], 2)};
    const Q4<Dual1> e = approx_exp(x);
    const Dual1* c[4] = {&e.x, &e.y, &e.z, &e.w};
    for (int r = 0; r < 4; ++r) {
        q[r] = c[r]->v;
        for (int j = 0; j < 3; ++j) jac[3 * r + j] = 0.0;
        for (const auto& t : c[r]->d) jac[3 * r + t.first] = t.second;
    }
    return 0;
}

}  // extern "C"

// Value and directional derivative dy/dx . dir of one function (dir has n_dec entries; the parameters carry no tangent):
// the `dwProjection` of BacktrackingLineSearch::Do (backtracking_line_search.hpp:92) without forming the gradient.
extern "C" int oracle_directional(int model, int fn, int N, const double* xp, const double* dir, double* y, double* dy) {
    using namespace oracle;
    Sizes s;
    if (!get_sizes(model, N, s)) return -1;
    std::vector<Dual1> v(s.n_dec + s.n_par);
    for (int i = 0; i < s.n_dec; ++i) v[i] = Dual1(xp[i], {{0, dir[i]}});
    for (int i = s.n_dec; i < s.n_dec + s.n_par; ++i) v[i] = Dual1(xp[i]);
    std::vector<Dual1> out;
    if (!eval_fn(model, fn, N, v.data(), out)) return -1;
    for (std::size_t r = 0; r < out.size(); ++r) {
        y[r]  = out[r].v;
        dy[r] = out[r].d.empty() ? 0.0 : out[r].d[0].second;
    }
    return static_cast<int>(out.size());
}
