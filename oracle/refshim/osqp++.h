// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for osqp-cpp v0.5.1-ungar / OSQP v0.6.3 (external/config/osqp-cpp/
// CMakeLists.txt.in:13), exposing the interface SoftSQPOptimizer uses (include/ungar/optimization/soft_sqp.hpp:160-234).
// The soft SQP only ever poses equality-constrained QPs (l = u = -g(x), soft_sqp.hpp:155-157; inequalities are folded
// into the objective by the barrier), so Solve() is one sparse KKT solve
//     [ P + sigma I   A^T ] [ x ]   [ -q ]
//     [ A            -rho I ] [ y ] = [  b ]        (sigma, rho tiny regularisers as in OSQP's quasi-definite KKT)
// with Eigen::SparseLU — not ADMM.  It exists so that the unchanged reference examples and tests can RUN in oracle/_ref.
#pragma once

#include <cstdlib>
#include <stdexcept>
#include <string>

#include <Eigen/Sparse>
#include <Eigen/SparseLU>

namespace osqp {

using c_int = int;

struct OsqpInstance {
    Eigen::SparseMatrix<double, Eigen::ColMajor, c_int> objective_matrix;  // upper triangle is used
    Eigen::VectorXd objective_vector;
    Eigen::SparseMatrix<double, Eigen::ColMajor, c_int> constraint_matrix;
    Eigen::VectorXd lower_bounds, upper_bounds;
};

struct OsqpSettings {
    bool verbose = false;
    bool polish  = false;
    double eps_abs = 1e-3, eps_rel = 1e-3;
    int max_iter = 4000;
};

enum class OsqpExitCode { kOptimal, kPrimalInfeasible, kDualInfeasible, kOptimalInaccurate, kMaxIterations, kUnknown };
inline std::string ToString(OsqpExitCode c) { return c == OsqpExitCode::kOptimal ? "optimal" : "not optimal"; }

class Status {
  public:
    Status() = default;
    explicit Status(std::string msg) : _ok(false), _msg(std::move(msg)) {}
    bool ok() const { return _ok; }
    const std::string& message() const { return _msg; }

  private:
    bool _ok = true;
    std::string _msg;
};

// Thrown by Solve() once UNGAR_REF_MAX_SOLVES solves have run: lets a driver stop an example's MPC loop after the
// functions have been created (oracle/ref_drivers/example_driver.cpp).
struct StopRequested : std::runtime_error {
    StopRequested() : std::runtime_error("UNGAR_REF_MAX_SOLVES reached") {}
};

class OsqpSolver {
  public:
    Status Init(const OsqpInstance& instance, const OsqpSettings&) {
        _P = instance.objective_matrix;
        _A = instance.constraint_matrix;
        _q = instance.objective_vector;
        _l = instance.lower_bounds;
        _u = instance.upper_bounds;
        if (_P.rows() != _P.cols() || _A.cols() != _P.cols() || _l.size() != _A.rows()) return Status("dimension mismatch");
        _init = true;
        return Status();
    }
    bool IsInitialized() const { return _init; }
    template <class PM, class AM>
    Status UpdateObjectiveAndConstraintMatrices(const PM& P, const AM& A) {
        _P = P;
        _A = A;
        return Status();
    }
    Status SetObjectiveVector(const Eigen::VectorXd& q) { _q = q; return Status(); }
    Status SetBounds(const Eigen::VectorXd& l, const Eigen::VectorXd& u) { _l = l; _u = u; return Status(); }

    OsqpExitCode Solve() {
        static long solves = 0;
        if (const char* cap = std::getenv("UNGAR_REF_MAX_SOLVES"))
            if (solves++ >= std::atol(cap)) throw StopRequested();
        if ((_l - _u).cwiseAbs().maxCoeff() > 1e-12) return OsqpExitCode::kUnknown;  // only equality-constrained QPs
        const int n = static_cast<int>(_P.cols()), m = static_cast<int>(_A.rows());
        std::vector<Eigen::Triplet<double>> t;
        for (int c = 0; c < _P.outerSize(); ++c)
            for (Eigen::SparseMatrix<double, Eigen::ColMajor, c_int>::InnerIterator it(_P, c); it; ++it) {
                if (it.row() > it.col()) continue;  // upper triangle, mirrored
                t.emplace_back(it.row(), it.col(), it.value());
                if (it.row() != it.col()) t.emplace_back(it.col(), it.row(), it.value());
            }
        for (int i = 0; i < n; ++i) t.emplace_back(i, i, 1e-9);
        for (int c = 0; c < _A.outerSize(); ++c)
            for (Eigen::SparseMatrix<double, Eigen::ColMajor, c_int>::InnerIterator it(_A, c); it; ++it) {
                t.emplace_back(n + it.row(), it.col(), it.value());
                t.emplace_back(it.col(), n + it.row(), it.value());
            }
        for (int i = 0; i < m; ++i) t.emplace_back(n + i, n + i, -1e-9);
        Eigen::SparseMatrix<double> K(n + m, n + m);
        K.setFromTriplets(t.begin(), t.end());
        Eigen::VectorXd rhs(n + m);
        rhs.head(n) = -_q;
        rhs.tail(m) = _l;
        Eigen::SparseLU<Eigen::SparseMatrix<double>> lu;
        lu.compute(K);
        if (lu.info() != Eigen::Success) return OsqpExitCode::kUnknown;
        _sol = lu.solve(rhs);
        if (lu.info() != Eigen::Success || !_sol.allFinite()) return OsqpExitCode::kUnknown;
        _x = _sol.head(n);
        return OsqpExitCode::kOptimal;
    }
    Eigen::Map<const Eigen::VectorXd> primal_solution() const { return Eigen::Map<const Eigen::VectorXd>(_x.data(), _x.size()); }

  private:
    Eigen::SparseMatrix<double, Eigen::ColMajor, c_int> _P, _A;
    Eigen::VectorXd _q, _l, _u, _sol, _x;
    bool _init = false;
};

}  // namespace osqp
