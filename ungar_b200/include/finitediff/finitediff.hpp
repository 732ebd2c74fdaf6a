// Stand-in for finite-diff v1.0.2-ungar (external/config/finite-diff/
// CMakeLists.txt.in:13), used by Function::TestJacobian / TestHessian (include/ungar/autodiff/function.hpp:285-325):
// central differences of the requested accuracy order (2nd order implemented; higher orders fall back to it).
#pragma once

#include <functional>

#include <Eigen/Core>

namespace fd {

enum AccuracyOrder { SECOND, FOURTH, SIXTH, EIGHTH };

inline void finite_jacobian(const Eigen::VectorXd& x, const std::function<Eigen::VectorXd(const Eigen::VectorXd&)>& f,
                            Eigen::MatrixXd& jac, const AccuracyOrder = SECOND, const double eps = 1e-8) {
    Eigen::VectorXd xp = x;
    const Eigen::VectorXd f0 = f(x);
    jac.resize(f0.size(), x.size());
    for (Eigen::Index j = 0; j < x.size(); ++j) {
        xp[j] = x[j] + eps;
        const Eigen::VectorXd fp = f(xp);
        xp[j] = x[j] - eps;
        const Eigen::VectorXd fm = f(xp);
        xp[j] = x[j];
        jac.col(j) = (fp - fm) / (2.0 * eps);
    }
}

inline void finite_hessian(const Eigen::VectorXd& x, const std::function<double(const Eigen::VectorXd&)>& f,
                           Eigen::MatrixXd& hess, const AccuracyOrder = SECOND, const double eps = 1e-5) {
    const Eigen::Index n = x.size();
    hess.resize(n, n);
    Eigen::VectorXd xp = x;
    for (Eigen::Index i = 0; i < n; ++i)
        for (Eigen::Index j = i; j < n; ++j) {
            auto at = [&](double si, double sj) {
                xp = x;
                xp[i] += si * eps;
                xp[j] += sj * eps;
                return f(xp);
            };
            const double v = (at(1, 1) - at(1, -1) - at(-1, 1) + at(-1, -1)) / (4.0 * eps * eps);
            hess(i, j) = hess(j, i) = v;
        }
}

}  // namespace fd
