// Test driver (tests/test_tape_host.py): the reference's UNCHANGED Autodiff::MakeFunction (function.hpp:607-613) over the product's
// CppAD-compatible header, with the reference's default recompileLibraries = false, for a lambda whose coefficient is a build-time
// constant.  Two builds (COEFF = 2, COEFF = 3) share one function NAME and one codegen folder: the reference's name-keyed library
// cache (function.hpp:420-451) would hand the second build the first build's function.  Nothing is evaluated (no GPU needed): the
// log says whether the library was "Loading"-ed or "Compiling"-ed, and the tape file shows which coefficient was taped.
#include "ungar/autodiff/function.hpp"

#ifndef COEFF
#define COEFF 2.0
#endif

int main() {
    using namespace Ungar;
    const Autodiff::Function::Blueprint blueprint{
        [](const VectorXad& xp, VectorXad& y) {
            y.resize(1);
            y << COEFF * xp[0] * xp[1];
        },
        2_idx, 0_idx, "stale_probe"sv, EnabledDerivatives::JACOBIAN};
    const auto f = Autodiff::MakeFunction(blueprint);  // recompileLibraries = false: the library is looked up by name
    return f.DependentVariableSize() == 1 ? 0 : 1;
}
