// C++ host mirror of the reference's evaluator interface over the C ABI (include/ungar_b200.h).
//
// `ungar_b200::Function` keeps the method names of `Ungar::Autodiff::Function`
// (include/ungar/autodiff/function.hpp:180-361): Evaluate / operator() / Jacobian / Hessian / Implements* / *Size, with
// host buffers in and out (batch 1 = the reference call; batch B = B reference calls in one launch).  It is header-only,
// needs no Eigen, and throws std::runtime_error where the reference asserts / throws (assert.hpp:97-108,
// function.hpp:531-534).  With Eigen available, `JacobianCsr()` gives exactly the three arrays the reference wraps in its
// `Eigen::Map<const SparseMatrix<real_t, RowMajor>>` (function.hpp:126-133); INTEGRATION.md shows that adaptor.
//
// `ungar_b200::Model` stands where MakeFunction x 3 + MakeNLPProblem stand in the examples
// (quadruped.example.cpp:343-363) and adds the batched KKT sweep that replaces
// SoftSQPOptimizer::AssembleOSQPInstance (optimization/soft_sqp.hpp:141-158).
//
// `ungar_b200::SoftSQPOptimizer` keeps the constructor arguments and the `Optimize(nlpProblem, xp)` call of
// `Ungar::SoftSQPOptimizer` (optimization/soft_sqp.hpp:42-109); the whole loop (KKT sweep, QP solve, backtracking line
// search) runs on the device.
#pragma once

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ungar_b200.h"

namespace ungar_b200 {

using index_t = std::int64_t;

inline void check(int status) {
    if (status != UNGAR_B200_OK) throw std::runtime_error(std::string("ungar_b200: ") + ungar_b200_last_error());
}

template <class Real>
struct dtype_of;
template <>
struct dtype_of<float> {
    static constexpr int value = UNGAR_B200_F32;
};
template <>
struct dtype_of<double> {
    static constexpr int value = UNGAR_B200_F64;
};

struct CsrPattern {  // row-major, rows ascending — what function.hpp:109-124 builds from (rows, cols)
    std::vector<int> innerStarts, outerIndices;
    index_t rows = 0, cols = 0;
};

template <class Real = double>
class Model;

template <class Real = double>
class Function {
  public:
    index_t IndependentVariableSize() const { return _nx; }
    index_t ParameterSize() const { return _np; }
    index_t DependentVariableSize() const { return _ny; }
    bool ImplementsFunction() const { return true; }
    bool ImplementsJacobian() const { return true; }
    bool ImplementsHessian() const { return _which == UNGAR_B200_OBJECTIVE || _which == UNGAR_B200_SOFT_INEQUALITIES; }
    index_t JacobianNonZeros() const { return _nnzJac; }
    index_t HessianNonZeros() const { return _nnzHes; }

    // y = f(xp) for `batch` stacked vectors xp (host memory, row stride = size of xp).
    void Evaluate(const Real* xp, Real* y, index_t batch = 1) const {
        check(ungar_b200_forward_zero(_model, _which, xp, batch, _nx + _np, y, _ny, UNGAR_B200_MEM_HOST, nullptr));
    }
    std::vector<Real> operator()(const std::vector<Real>& xp) const {
        if (static_cast<index_t>(xp.size()) != _nx + _np) throw std::runtime_error("ungar_b200: xp has the wrong size");
        std::vector<Real> y(static_cast<std::size_t>(_ny));
        Evaluate(xp.data(), y.data());
        return y;
    }
    // Nonzeros of the Jacobian in JacobianCsr() order; returns a reference to an internal buffer that the next call
    // overwrites — the ownership rule of the reference (function.hpp:380-383).
    const std::vector<Real>& Jacobian(const std::vector<Real>& xp) const {
        _jacobianData.resize(static_cast<std::size_t>(_nnzJac));
        check(ungar_b200_sparse_jacobian(_model, _which, xp.data(), 1, _nx + _np, _jacobianData.data(), _nnzJac,
                                         UNGAR_B200_MEM_HOST, nullptr));
        return _jacobianData;
    }
    // Upper-triangular Hessian of a scalar function (function.hpp:232-258).
    const std::vector<Real>& Hessian(const std::vector<Real>& xp) const {
        _hessianData.resize(static_cast<std::size_t>(_nnzHes));
        check(ungar_b200_sparse_hessian(_model, _which, xp.data(), 1, _nx + _np, _hessianData.data(), _nnzHes,
                                        UNGAR_B200_MEM_HOST, nullptr));
        return _hessianData;
    }
    const CsrPattern& JacobianCsr() const { return _jacobianCsr; }
    const CsrPattern& HessianCsr() const { return _hessianCsr; }

  private:
    friend class Model<Real>;
    Function(ungar_b200_model* model, int which) : _model(model), _which(which) {
        check(ungar_b200_function_info(model, which, &_nx, &_np, &_ny, &_nnzJac, &_nnzHes));
        const int64_t *rows, *cols;
        int64_t nnz;
        check(ungar_b200_jacobian_sparsity(model, which, &rows, &cols, &nnz));
        _jacobianCsr = MakeCsr(rows, cols, nnz, _ny, _nx);
        if (ImplementsHessian()) {
            check(ungar_b200_hessian_sparsity(model, which, &rows, &cols, &nnz));
            _hessianCsr = MakeCsr(rows, cols, nnz, _nx, _nx);
        }
    }
    static CsrPattern MakeCsr(const int64_t* rows, const int64_t* cols, int64_t nnz, index_t nRows, index_t nCols) {
        CsrPattern p;
        p.rows = nRows;
        p.cols = nCols;
        p.innerStarts.assign(static_cast<std::size_t>(nRows) + 1, 0);
        p.outerIndices.resize(static_cast<std::size_t>(nnz));
        for (int64_t e = 0; e < nnz; ++e) {
            ++p.innerStarts[static_cast<std::size_t>(rows[e]) + 1];
            p.outerIndices[static_cast<std::size_t>(e)] = static_cast<int>(cols[e]);
        }
        for (index_t r = 0; r < nRows; ++r) p.innerStarts[r + 1] += p.innerStarts[r];
        return p;
    }

    ungar_b200_model* _model;
    int _which;
    int64_t _nx = 0, _np = 0, _ny = 0, _nnzJac = 0, _nnzHes = 0;
    CsrPattern _jacobianCsr, _hessianCsr;
    mutable std::vector<Real> _jacobianData, _hessianData;
};

template <class Real>
class Model {
  public:
    Model(int kind, int horizon, double barrierStiffness, double barrierEpsilon, int device = 0) {
        ungar_b200_model_desc desc{kind, horizon, dtype_of<Real>::value, device, barrierStiffness, barrierEpsilon};
        check(ungar_b200_model_create(&desc, &_handle));
        check(ungar_b200_kkt_layout_get(_handle, &_layout));
        _stiffness = barrierStiffness;
        _epsilon   = barrierEpsilon;
    }
    ~Model() { ungar_b200_model_destroy(_handle); }
    Model(const Model&)            = delete;
    Model& operator=(const Model&) = delete;

    Function<Real> objective() const { return Function<Real>(_handle, UNGAR_B200_OBJECTIVE); }
    Function<Real> equalityConstraints() const { return Function<Real>(_handle, UNGAR_B200_EQUALITIES); }
    Function<Real> inequalityConstraints() const { return Function<Real>(_handle, UNGAR_B200_INEQUALITIES); }
    Function<Real> softInequalityConstraints() const { return Function<Real>(_handle, UNGAR_B200_SOFT_INEQUALITIES); }
    const ungar_b200_kkt_layout& layout() const { return _layout; }
    index_t VariableSize() const { return _layout.n_dec + _layout.n_par; }

    // Device-resident sweep (pointers are device pointers; asynchronous on `stream`).
    void KktBlocksDevice(const Real* xp, index_t batch, index_t ldXp, Real* records, index_t ldRec, void* stream = nullptr) {
        check(ungar_b200_kkt_blocks(_handle, xp, batch, ldXp, records, ldRec, UNGAR_B200_MEM_DEVICE, stream));
    }
    // Host buffers in and out (H2D, sweep, D2H inside the call).
    void KktBlocks(const Real* xp, index_t batch, Real* records) {
        check(ungar_b200_kkt_blocks(_handle, xp, batch, VariableSize(), records, _layout.size, UNGAR_B200_MEM_HOST, nullptr));
    }
    // One outer-iteration step: records stay in HBM, [batch][32] summaries come back to the host.
    void Step(const Real* xpHost, index_t batch, Real* summariesHost, Real* recordsDevice = nullptr, void* stream = nullptr) {
        check(ungar_b200_kkt_step(_handle, xpHost, batch, VariableSize(), recordsDevice, _layout.size, summariesHost,
                                  UNGAR_B200_MEM_HOST, stream));
    }
    // The Jacobian sweep alone (g, A and, for the quadruped, C of the record; device pointers).
    void JacobianBlocksDevice(const Real* xp, index_t batch, index_t ldXp, Real* records, index_t ldRec, void* stream = nullptr) {
        check(ungar_b200_jacobian_blocks(_handle, xp, batch, ldXp, records, ldRec, UNGAR_B200_MEM_DEVICE, stream));
    }
    // Exact solve of the QP of the records (device pointers): steps[batch][n_dec], optional multipliers[batch][m_eq].  F64 models.
    void QpSolveDevice(const Real* records, index_t batch, index_t ldRec, Real* steps, Real* multipliers = nullptr, void* stream = nullptr) {
        check(ungar_b200_qp_solve(_handle, records, batch, ldRec, steps, _layout.n_dec, multipliers, _layout.m_eq, stream));
    }
    // BacktrackingLineSearch::Do for the whole batch (device pointers): xp[:, 0:n_dec] += alpha * steps where a step is accepted.
    void LineSearchDevice(Real* xp, index_t batch, index_t ldXp, const Real* steps, const ungar_b200_sqp_options& options,
                          std::int32_t* status = nullptr, Real* info = nullptr, void* stream = nullptr) {
        check(ungar_b200_line_search(_handle, xp, batch, ldXp, steps, _layout.n_dec, &options, status, info, stream));
    }
    ungar_b200_model* handle() const { return _handle; }
    double BarrierStiffness() const { return _stiffness; }
    double BarrierEpsilon() const { return _epsilon; }

  private:
    ungar_b200_model* _handle = nullptr;
    ungar_b200_kkt_layout _layout{};
    double _stiffness = 0.0, _epsilon = 0.0;
};

// Mirror of Ungar::SoftSQPOptimizer (optimization/soft_sqp.hpp:42-61).  The relaxed barrier (stiffness, epsilon) is part of
// the model handle here — the reference JIT-compiles it into a Function of its own on first use (soft_sqp.hpp:114-138, :162) —
// so Optimize checks that both agree.  Only the POLY barrier exists (the reference's default, soft_sqp.hpp:49).
class SoftSQPOptimizer {
  public:
    struct TrajectoryStatus {
        std::int32_t status;      // ungar_b200_sqp_status
        std::int32_t iterations;  // iterations started
    };

    explicit SoftSQPOptimizer(bool verbose, double constraintViolationMultiplier = 1.0, index_t maxIterations = 10,
                              double stiffness = 100.0, double epsilon = 2e-5)
        : _verbose(verbose), _stiffness(stiffness), _epsilon(epsilon) {
        check(ungar_b200_sqp_options_default(&_options));
        _options.max_iterations                  = static_cast<std::int32_t>(maxIterations);
        _options.constraint_violation_multiplier = constraintViolationMultiplier;
    }
    ungar_b200_sqp_options& Options() { return _options; }  // BacktrackingLineSearch::SetParameters equivalent

    // xp: `batch` stacked flat vectors [X | U | parameters] (host).  Returns the optimised decision variables, `batch` rows of
    // n_dec (the reference returns _cache.xp.head(n_dec), soft_sqp.hpp:108).
    std::vector<double> Optimize(Model<double>& nlpProblem, const std::vector<double>& xp, index_t batch = 1) {
        if (nlpProblem.BarrierStiffness() != _stiffness || nlpProblem.BarrierEpsilon() != _epsilon)
            throw std::runtime_error("ungar_b200: the model was created with a different barrier than this optimizer");
        const index_t n = nlpProblem.VariableSize(), nDec = nlpProblem.layout().n_dec;
        if (static_cast<index_t>(xp.size()) != batch * n) throw std::runtime_error("ungar_b200: xp has the wrong size");
        std::vector<double> work(xp);
        _status.assign(static_cast<std::size_t>(batch), TrajectoryStatus{0, 0});
        _info.assign(static_cast<std::size_t>(batch) * UNGAR_B200_LINE_SEARCH_INFO_SIZE, 0.0);
        check(ungar_b200_sqp_solve(nlpProblem.handle(), work.data(), batch, n, &_options, reinterpret_cast<std::int32_t*>(_status.data()),
                                   _info.data(), UNGAR_B200_MEM_HOST, nullptr));
        std::vector<double> x(static_cast<std::size_t>(batch * nDec));
        for (index_t b = 0; b < batch; ++b)
            for (index_t i = 0; i < nDec; ++i) x[static_cast<std::size_t>(b * nDec + i)] = work[static_cast<std::size_t>(b * n + i)];
        if (_verbose)
            for (index_t b = 0; b < batch; ++b)
                std::fprintf(stderr, "soft SQP [%lld]: status %d after %d iteration(s), last step size %g\n", (long long)b,
                             _status[static_cast<std::size_t>(b)].status, _status[static_cast<std::size_t>(b)].iterations,
                             _info[static_cast<std::size_t>(b) * UNGAR_B200_LINE_SEARCH_INFO_SIZE]);
        return x;
    }
    const std::vector<TrajectoryStatus>& Status() const { return _status; }
    // [batch][8] report of the last line search: alpha, theta, phi, f | theta, phi, f before the step, grad f . dw
    const std::vector<double>& LineSearchInfo() const { return _info; }

  private:
    bool _verbose;
    double _stiffness, _epsilon;
    ungar_b200_sqp_options _options{};
    std::vector<TrajectoryStatus> _status;
    std::vector<double> _info;
};

}  // namespace ungar_b200
