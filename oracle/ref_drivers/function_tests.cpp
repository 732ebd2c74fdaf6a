// ORACLE — TEST INFRASTRUCTURE ONLY.  Drives the reference's own Function class (include/ungar/autodiff/function.hpp, used
// unchanged from /root/reference) through the known answers its own tests pin (test/autodiff/function.test.cpp:33-142).
// GoogleTest is absent, so each known answer is a named check in a table; exit code 0 = all reproduced.  Compiled twice: against
// oracle/refshim (CPU tape evaluator: oracle/_ref/function_tests) and against the product's CppAD-compatible header
// (GPU register machine: tests/_ref_gpu/function_tests_gpu).
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "ungar/autodiff/function.hpp"

namespace {

using namespace Ungar;
using namespace Ungar::Autodiff;

int g_failures = 0;

void expect(bool ok, const std::string& what) {
    if (ok) return;
    std::fprintf(stderr, "KNOWN ANSWER NOT REPRODUCED: %s\n", what.c_str());
    ++g_failures;
}

// `count` random points through a self-check of the Function (AD against finite differences or against a plain evaluation).
void sweep(int count, index_t size, const std::string& what, const std::function<bool(const VectorXr&)>& check) {
    bool all = true;
    for (int i = 0; i < count && all; ++i) all = check(VectorXr::Random(size));
    expect(all, what);
}

// y = p |x|^2 (optionally followed by 2 x0^2) on xp = [x (4); p (1)]: the function of the Jacobian and Hessian tests.
template <bool WITH_SECOND_ROW>
struct NormTimesParameter {
    template <typename _Scalar>
    void operator()(const VectorX<_Scalar>& xp, VectorX<_Scalar>& y) const {
        const auto [x, p] = Utils::Decompose<4, 1>(xp);
        if constexpr (WITH_SECOND_ROW) y = VectorX<_Scalar>{{p * x.squaredNorm(), 2.0 * pow(x[0_idx], 2)}};
        else y = VectorX<_Scalar>{{p * x.squaredNorm()}};
    }
};

void exponential_map() {  // function.test.cpp:33-59
    auto approx = []<typename _Scalar>(const VectorX<_Scalar>& x, VectorX<_Scalar>& y) -> void {
        y = Utils::ApproximateExponentialMap(RefToConstVector3<_Scalar>{x}).coeffs();
    };
    const auto exact = [](const Vector3r& v) { return Utils::ExponentialMap(v).coeffs(); };
    Function function = MakeFunction(Function::Blueprint{approx, 3, 0, "exponential_map_test", EnabledDerivatives::JACOBIAN}, true);
    const VectorXr origin = Vector3r::Zero();
    expect(function.TestFunction(origin, exact), "approximate exponential map at the origin");
    expect(function.TestJacobian(origin), "its Jacobian against finite differences at the origin");
    MatrixXr halfIdentity = MatrixXr::Zero(4, 3);
    halfIdentity.topRows(3) = 0.5 * MatrixXr::Identity(3, 3);
    expect((function.Jacobian(origin).toDense() - halfIdentity).cwiseAbs().maxCoeff() < 1e-7, "Jacobian at the origin = [I/2; 0]");
    sweep(1024, 3, "approximate = exact exponential map on [-1, 1]^3", [&](const VectorXr& v) { return function.TestFunction(v, exact); });
}

void jacobian() {  // function.test.cpp:61-109
    const NormTimesParameter<true> f;
    Function function = MakeFunction(Function::Blueprint{f, 4, 1, "jacobian_test", EnabledDerivatives::JACOBIAN}, true);
    const VectorXr x = VectorXr::Random(4), p = VectorXr::Random(1);
    const VectorXr xp = Utils::Compose(x, p).ToDynamic();
    MatrixXr jacobianTruth = MatrixXr::Zero(2, 4);
    jacobianTruth.row(0) = 2.0 * p[0] * x.transpose();
    jacobianTruth(1, 0) = 4.0 * x[0];
    expect(function(xp).isApprox(VectorXr{{p[0] * x.squaredNorm(), 2.0 * pow(x[0], 2)}}), "y = [p |x|^2, 2 x0^2]");
    expect(function.Jacobian(xp).isApprox(jacobianTruth), "J = [[2 p x], [4 x0, 0, 0, 0]]");
    const auto plain = [&](const VectorXr& v) {
        VectorXr y;
        f.template operator()<real_t>(v, y);
        return y;
    };
    sweep(1024, 5, "values and Jacobian at 1024 random points",
          [&](const VectorXr& v) { return function.TestFunction(v, plain) && function.TestJacobian(v); });
}

void hessian() {  // function.test.cpp:111-142
    Function function = MakeFunction(Function::Blueprint{NormTimesParameter<false>{}, 4, 1, "hessian_test", EnabledDerivatives::HESSIAN}, true);
    const VectorXr x = VectorXr::Random(4), p = VectorXr::Random(1);
    expect(function.Hessian(Utils::Compose(x, p).ToDynamic()).isApprox(MatrixXr{2.0 * p[0] * MatrixXr::Identity(4, 4)}), "H = 2 p I");
    sweep(256, 5, "Hessian against finite differences at 256 random points", [&](const VectorXr& v) { return function.TestHessian(v); });
}

}  // namespace

int main() {
    std::srand(7);
    const std::vector<std::pair<const char*, void (*)()>> table = {{"ExponentialMap", exponential_map}, {"Jacobian", jacobian}, {"Hessian", hessian}};
    for (const auto& [name, body] : table) {
        const int before = g_failures;
        body();
        std::printf("FunctionTest.%s: %s\n", name, g_failures == before ? "ok" : "FAILED");
    }
    if (g_failures) return 1;
    std::printf("function_tests: all reference known answers reproduced\n");
    return 0;
}
