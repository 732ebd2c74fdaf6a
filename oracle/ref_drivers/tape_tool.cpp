// ORACLE — TEST INFRASTRUCTURE ONLY.  C interface over the tape "libraries" written by the reference's MakeFunction when
// it runs against oracle/refshim (same GenericModel code path the reference's Function uses).  Loaded with ctypes by
// oracle/make_golden.py to turn the reference's own lambdas into golden vectors.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include <cppad/cg.hpp>

using Model = CppAD::cg::GenericModel<double>;
struct Handle {
    std::unique_ptr<CppAD::cg::LinuxDynamicLib<double>> lib;
    std::unique_ptr<Model> model;
    std::vector<std::size_t> jr, jc, hr, hc;
};

extern "C" {

void* reftape_open(const char* path) {
    try {
        auto h = new Handle;
        h->lib = std::make_unique<CppAD::cg::LinuxDynamicLib<double>>(path);
        h->model = h->lib->model("tape");
        if (h->model->isJacobianSparsityAvailable()) h->model->JacobianSparsity(h->jr, h->jc);
        if (h->model->isHessianSparsityAvailable()) h->model->HessianSparsity(0, h->hr, h->hc);
        return h;
    } catch (...) {
        return nullptr;
    }
}
void reftape_close(void* p) { delete static_cast<Handle*>(p); }
// info[5] = n_indep, n_dep, nnz_jac, nnz_hes, n_nodes
void reftape_info(void* p, int64_t* info) {
    auto* h = static_cast<Handle*>(p);
    info[0] = h->model->Domain(); info[1] = h->model->Range(); info[2] = h->jr.size(); info[3] = h->hr.size();
    info[4] = h->model->tape()->nodes.size();
}
void reftape_eval(void* p, const double* x, double* y) {
    auto* h = static_cast<Handle*>(p);
    h->model->ForwardZero({x, h->model->Domain()}, {y, h->model->Range()});
}
void reftape_jacobian(void* p, const double* x, int64_t* rows, int64_t* cols, double* vals) {
    auto* h = static_cast<Handle*>(p);
    const std::size_t *r, *c;
    h->model->SparseJacobian({x, h->model->Domain()}, {vals, h->jr.size()}, &r, &c);
    for (std::size_t e = 0; e < h->jr.size(); ++e) { rows[e] = r[e]; cols[e] = c[e]; }
}
void reftape_hessian(void* p, const double* x, const double* w, int64_t* rows, int64_t* cols, double* vals) {
    auto* h = static_cast<Handle*>(p);
    const std::size_t *r, *c;
    h->model->SparseHessian({x, h->model->Domain()}, {w, h->model->Range()}, {vals, h->hr.size()}, &r, &c);
    for (std::size_t e = 0; e < h->hr.size(); ++e) { rows[e] = r[e]; cols[e] = c[e]; }
}

}  // extern "C"
