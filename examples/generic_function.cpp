// The generic path without the reference: record an arbitrary lambda with the tracing scalar of ungar_b200/include/cppad/cg.hpp
// (the API Ungar's function.hpp uses: CppAD::Independent, CppAD::ADFun, ModelCSourceGen, GenericModel) and evaluate values, the
// sparse Jacobian and the sparse Hessian on the GPU.  The function is the one the reference pins in test/autodiff/function.test.cpp:
// y = p |x|^2 with x in R^4, p in R^1  ->  dy/dx = 2 p x,  d2y/dx2 = 2 p I.
// Build:  g++ -std=c++17 -Iinclude -Iungar_b200/include examples/generic_function.cpp ungar_b200/libungar_b200.so
#include <cmath>
#include <cstdio>
#include <vector>

#include "cppad/cg.hpp"

int main() {
    using ADCG = CppAD::AD<CppAD::cg::CG<double>>;
    try {
        std::vector<ADCG> xp(5), y(1);
        for (int i = 0; i < 5; ++i) xp[i] = 0.5 + 0.1 * i;  // taping point
        CppAD::Independent(xp);
        ADCG n2 = 0.0;
        for (int i = 0; i < 4; ++i) n2 += xp[i] * xp[i];
        y[0] = xp[4] * n2 + CppAD::CondExpGt(xp[0], ADCG(10.0), CppAD::sin(xp[1]), ADCG(0.0));  // the branch is decided at evaluation time
        CppAD::ADFun<CppAD::cg::CG<double>> fun(xp, y);

        CppAD::cg::ModelCSourceGen<double> gen(fun, "generic_function");
        gen.setCreateSparseJacobian(true);
        gen.setCreateSparseHessian(true);
        CppAD::cg::GenericModel<double> model("generic_function", gen.tape, true, true, {}, {});

        const std::vector<double> x = {0.3, -0.2, 0.5, 0.1, 2.0};
        std::vector<double> value(1);
        model.ForwardZero({x.data(), x.size()}, {value.data(), value.size()});
        std::vector<std::size_t> jr, jc, hr, hc;
        model.JacobianSparsity(jr, jc);
        model.HessianSparsity(0, hr, hc);
        std::vector<double> jac(jr.size()), hes(hr.size()), w = {1.0};
        const std::size_t *rows, *cols;
        model.SparseJacobian({x.data(), x.size()}, {jac.data(), jac.size()}, &rows, &cols);
        model.SparseHessian({x.data(), x.size()}, {w.data(), w.size()}, {hes.data(), hes.size()}, &rows, &cols);

        double err = std::fabs(value[0] - 2.0 * (0.09 + 0.04 + 0.25 + 0.01));
        for (std::size_t e = 0; e < jr.size(); ++e) {
            const double expect = jc[e] < 4 ? 2.0 * x[4] * x[jc[e]] : 0.39;  // the sin(x1) branch is inactive at x0 = 0.3
            err = std::fmax(err, std::fabs(jac[e] - expect));
        }
        for (std::size_t e = 0; e < hr.size(); ++e) {
            const double expect = hr[e] == hc[e] ? (hr[e] < 4 ? 2.0 * x[4] : 0.0) : ((hr[e] == 4 || hc[e] == 4) ? 2.0 * x[hr[e] == 4 ? hc[e] : hr[e]] : 0.0);
            err = std::fmax(err, std::fabs(hes[e] - expect));
        }
        std::printf("generic function on the GPU: y = %.6f, nnz(J) = %zu, nnz(H) = %zu, max error vs closed form = %.3e\n", value[0],
                    jr.size(), hr.size(), err);
        return err < 1e-12 ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
