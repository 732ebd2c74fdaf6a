// Product header — stand-in for osqp-cpp v0.5.1-ungar / OSQP v0.6.3 (external/config/osqp-cpp/CMakeLists.txt.in:13), exposing the
// interface SoftSQPOptimizer uses (include/ungar/optimization/soft_sqp.hpp:160-234), so that the reference's UNCHANGED optimizer,
// tests and MPC examples link against ungar_b200 and solve their local QPs on the GPU.
//
// The soft SQP only ever poses equality-constrained QPs (l = u = -g(x), soft_sqp.hpp:155-157; inequalities are folded into the
// objective by the relaxed barrier), so Solve() is ONE exact solve of the quasi-definite KKT system on the device through
// ungar_b200_kkt_solve_csc (include/ungar_b200.h) — not ADMM, no tolerances to tune; sigma = rho = 1e-9 play the role of OSQP's
// regularisers.  A QP with l != u is outside what the reference uses and is reported as kUnknown.  There is no CPU solve here.
// The three reference MPC problems additionally have stage-wise solvers that consume the KKT blocks in place
// (ungar_b200_qp_solve / ungar_b200_sqp_solve), 10^2-10^3 x faster than this generic path.
#pragma once

#include <cstdlib>
#include <stdexcept>
#include <string>

#include <Eigen/Sparse>

#include "ungar_b200.h"

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
    double eps_abs = 1e-3, eps_rel = 1e-3;  // accepted for source compatibility: the solve is exact
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

// Thrown by Solve() once UNGAR_B200_MAX_QP_SOLVES solves have run: lets a driver bound the endless receding-horizon loop of an
// unchanged reference example.
struct StopRequested : std::runtime_error {
    StopRequested() : std::runtime_error("UNGAR_B200_MAX_QP_SOLVES reached") {}
};

class OsqpSolver {
  public:
    Status Init(const OsqpInstance& instance, const OsqpSettings&) {
        _P = instance.objective_matrix;
        _A = instance.constraint_matrix;
        _q = instance.objective_vector;
        _l = instance.lower_bounds;
        _u = instance.upper_bounds;
        if (_P.rows() != _P.cols() || (_A.rows() > 0 && _A.cols() != _P.cols()) || _l.size() != _A.rows() || _u.size() != _A.rows())
            return Status("dimension mismatch");
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
        if (const char* cap = std::getenv("UNGAR_B200_MAX_QP_SOLVES"))
            if (solves++ >= std::atol(cap)) throw StopRequested();
        const long n = static_cast<long>(_P.cols()), m = static_cast<long>(_A.rows());
        if (m > 0 && (_l - _u).cwiseAbs().maxCoeff() > 1e-12) return OsqpExitCode::kUnknown;  // only equality-constrained QPs
        _P.makeCompressed();
        _A.makeCompressed();
        _x.resize(n);
        _y.resize(m);
        const int rc = ungar_b200_kkt_solve_csc(n, m, _P.outerIndexPtr(), _P.innerIndexPtr(), _P.valuePtr(), _q.data(),
                                                m > 0 ? _A.outerIndexPtr() : nullptr, m > 0 ? _A.innerIndexPtr() : nullptr,
                                                m > 0 ? _A.valuePtr() : nullptr, m > 0 ? _l.data() : nullptr, 1e-9, 1e-9, _x.data(),
                                                m > 0 ? _y.data() : nullptr, /*device*/ 0);
        if (rc == UNGAR_B200_ECUDA) throw std::runtime_error(std::string("ungar_b200: ") + ungar_b200_last_error());
        if (rc != UNGAR_B200_OK || !_x.allFinite()) return OsqpExitCode::kUnknown;
        return OsqpExitCode::kOptimal;
    }
    Eigen::Map<const Eigen::VectorXd> primal_solution() const { return Eigen::Map<const Eigen::VectorXd>(_x.data(), _x.size()); }
    Eigen::Map<const Eigen::VectorXd> dual_solution() const { return Eigen::Map<const Eigen::VectorXd>(_y.data(), _y.size()); }

  private:
    Eigen::SparseMatrix<double, Eigen::ColMajor, c_int> _P, _A;
    Eigen::VectorXd _q, _l, _u, _x, _y;
    bool _init = false;
};

}  // namespace osqp
