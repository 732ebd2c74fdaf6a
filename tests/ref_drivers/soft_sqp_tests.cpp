// The reference's own SoftSQPOptimizer (include/ungar/optimization/soft_sqp.hpp, unchanged, from /root/reference) driven through
// the three NLPs its own test pins (test/optimization/soft_sqp.test.cpp:34-111) over the PRODUCT headers: functions taped and
// evaluated on the GPU (cppad/cg.hpp), local QPs solved on the GPU (osqp++.h).  GoogleTest is absent, so the test body is restated
// with plain checks; exit code 0 = the three optima (3, 1), (2, 1), (1, 1) are reached within the test's isApprox(1e-1).
#include <cstdio>
#include <cstdlib>

#include "ungar/autodiff/function.hpp"
#include "ungar/optimization/soft_sqp.hpp"

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
    using namespace std::literals;
    const index_t xSize = 2_idx, pSize = 0_idx;
    const auto close = [](const VectorXr& a, const VectorXr& b) { return a.isApprox(b, 1e-1); };

    const auto objBlueprint = Function::Blueprint{
        [](const VectorXad& xp, VectorXad& y) {
            y = VectorXad{{Utils::Pow(xp.x() - ad_scalar_t{3.0}, 2) + Utils::Pow(xp.y() - ad_scalar_t{2.0}, 2)}};
        },
        xSize, pSize, "obj_soft_sqp_test"sv};
    const auto eqsBlueprint = Function::Blueprint{[](const VectorXad& xp, VectorXad& y) { y = VectorXad{{xp.x() - xp.y()}}; }, xSize,
                                                  pSize, "eqs_soft_sqp_test"sv, EnabledDerivatives::JACOBIAN};
    const auto ineqsBlueprint1 = Function::Blueprint{
        [](const VectorXad& xp, VectorXad& y) { y = VectorXad{{xp.y() - ad_scalar_t{1.0}, -xp.x()}}; }, xSize, pSize,
        "ineqs_1_soft_sqp_test"sv, EnabledDerivatives::JACOBIAN};
    const auto ineqsBlueprint2 = Function::Blueprint{
        [](const VectorXad& xp, VectorXad& y) {
            y = VectorXad{{Utils::Pow(xp.x(), 2) - xp.y() - ad_scalar_t{3.0}, xp.y() - ad_scalar_t{1.0}, -xp.x()}};
        },
        xSize, pSize, "ineqs_2_soft_sqp_test"sv, EnabledDerivatives::JACOBIAN};

    {
        auto nlp = MakeNLPProblem(MakeFunction(objBlueprint, true), hana::nothing, MakeFunction(ineqsBlueprint1, true));
        SoftSQPOptimizer optimizer{false, 1.0, 100_idx, 100.0, 2e-8};
        const VectorXr xOpt = optimizer.Optimize(nlp, VectorXr::Zero(xSize + pSize));
        std::printf("problem 1: (%.4f, %.4f), ground truth (3, 1)\n", xOpt[0], xOpt[1]);
        CHECK(close(xOpt, VectorXr{{3.0, 1.0}}));
    }
    {
        auto nlp = MakeNLPProblem(MakeFunction(objBlueprint, false), hana::nothing, MakeFunction(ineqsBlueprint2, true));
        SoftSQPOptimizer optimizer{false, 1.0, 100_idx, 100.0, 2e-8};
        const VectorXr xOpt = optimizer.Optimize(nlp, VectorXr::Zero(xSize + pSize));
        std::printf("problem 2: (%.4f, %.4f), ground truth (2, 1)\n", xOpt[0], xOpt[1]);
        CHECK(close(xOpt, VectorXr{{2.0, 1.0}}));
    }
    {
        auto nlp = MakeNLPProblem(MakeFunction(objBlueprint, false), MakeFunction(eqsBlueprint, true), MakeFunction(ineqsBlueprint2, false));
        SoftSQPOptimizer optimizer{false, 1.0, 100_idx, 100.0, 2e-8};
        const VectorXr xOpt = optimizer.Optimize(nlp, VectorXr::Zero(xSize + pSize));
        std::printf("problem 3: (%.4f, %.4f), ground truth (1, 1)\n", xOpt[0], xOpt[1]);
        CHECK(close(xOpt, VectorXr{{1.0, 1.0}}));
    }
    std::printf("soft_sqp_tests: all reference optima reproduced\n");
    return 0;
}
