// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's own Function class (include/ungar/autodiff/function.hpp, used
// unchanged from /root/reference) driven through the toy functions its own tests pin
// (test/autodiff/function.test.cpp:33-142).  GoogleTest is absent, so the test bodies are restated with plain checks;
// exit code 0 = all known answers reproduced.
#include <cstdio>
#include <cstdlib>

#include "ungar/autodiff/function.hpp"

#define CHECK(cond)                                                                      \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            std::fprintf(stderr, "CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                                \
        }                                                                                \
    } while (0)

int main() {
    using namespace Ungar;
    using namespace Ungar::Autodiff;
    std::srand(7);

    {  // TEST(FunctionTest, ExponentialMap), function.test.cpp:33-59
        auto exp = []<typename _Scalar>(const VectorX<_Scalar>& x, VectorX<_Scalar>& y) -> void {
            y = Utils::ApproximateExponentialMap(RefToConstVector3<_Scalar>{x}).coeffs();
        };
        Function::Blueprint blueprint{exp, 3, 0, "exponential_map_test", EnabledDerivatives::JACOBIAN};
        Function function = MakeFunction(blueprint, true);
        VectorXr x = Vector3r::Zero();
        CHECK(function.TestFunction(x, [&](const Vector3r& v) { return Utils::ExponentialMap(v).coeffs(); }));
        CHECK(function.TestJacobian(x));
        MatrixXr J0 = function.Jacobian(x).toDense();
        MatrixXr expect = MatrixXr::Zero(4, 3);
        expect.topRows(3) = 0.5 * MatrixXr::Identity(3, 3);
        CHECK((J0 - expect).cwiseAbs().maxCoeff() < 1e-7);
        for (int i = 0; i < 1024; ++i) {
            x = Vector3r::Random();
            CHECK(function.TestFunction(x, [&](const Vector3r& v) { return Utils::ExponentialMap(v).coeffs(); }));
        }
    }
    {  // TEST(FunctionTest, Jacobian), function.test.cpp:61-109
        auto f = []<typename _Scalar>(const VectorX<_Scalar>& xp, VectorX<_Scalar>& y) -> void {
            const auto [x, p] = Utils::Decompose<4, 1>(xp);
            y                 = VectorX<_Scalar>{{p * x.squaredNorm(), 2.0 * pow(x[0_idx], 2)}};
        };
        Function::Blueprint blueprint{f, 4, 1, "jacobian_test", EnabledDerivatives::JACOBIAN};
        Function function = MakeFunction(blueprint, true);
        const VectorXr x = VectorXr::Random(4), p = VectorXr::Random(1);
        const VectorXr xp = Utils::Compose(x, p).ToDynamic();
        const VectorXr yGroundTruth = VectorXr{{p[0] * x.squaredNorm(), 2.0 * pow(x[0], 2)}};
        const MatrixXr jacobianGroundTruth =
            MatrixXr{{2.0 * p[0] * x[0], 2.0 * p[0] * x[1], 2.0 * p[0] * x[2], 2.0 * p[0] * x[3]}, {4.0 * x[0], 0.0, 0.0, 0.0}};
        CHECK(function(xp).isApprox(yGroundTruth));
        CHECK(function.Jacobian(xp).isApprox(jacobianGroundTruth));
        auto func = [&](const VectorXr& v) {
            VectorXr y;
            f.template operator()<real_t>(v, y);
            return y;
        };
        for (int i = 0; i < 1024; ++i) {
            VectorXr v = VectorXr::Random(5);
            CHECK(function.TestFunction(v, func));
            CHECK(function.TestJacobian(v));
        }
    }
    {  // TEST(FunctionTest, Hessian), function.test.cpp:111-142
        auto f = []<typename _Scalar>(const VectorX<_Scalar>& xp, VectorX<_Scalar>& y) -> void {
            const auto [x, p] = Utils::Decompose<4, 1>(xp);
            y                 = VectorX<_Scalar>{{p * x.squaredNorm()}};
        };
        Function::Blueprint blueprint{f, 4, 1, "hessian_test", EnabledDerivatives::HESSIAN};
        Function function = MakeFunction(blueprint, true);
        const VectorXr x = VectorXr::Random(4), p = VectorXr::Random(1);
        const VectorXr xp = Utils::Compose(x, p).ToDynamic();
        const MatrixXr hessianGroundTruth = 2.0 * p[0] * MatrixXr::Identity(4, 4);
        CHECK(function.Hessian(xp).isApprox(hessianGroundTruth));
        for (int i = 0; i < 256; ++i) {
            VectorXr v = VectorXr::Random(5);
            CHECK(function.TestHessian(v));
        }
    }
    std::printf("function_tests: all reference known answers reproduced\n");
    return 0;
}
