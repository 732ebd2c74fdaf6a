// ORACLE — TEST INFRASTRUCTURE ONLY.  Runs one UNCHANGED reference example (example/mpc/*.example.cpp, compiled from where
// it lies, or a copy under oracle/_ref/src whose ONLY edit is the horizon constant `N`) against the tracing shim: the
// example's own lambdas are taped by the reference's own MakeFunction, the tapes land in UNGAR_CODEGEN_FOLDER, and the
// MPC loop is stopped after UNGAR_REF_MAX_SOLVES QP solves (0 = right after the functions exist).
#include <cstdio>

#include <osqp++.h>

#define main ungar_reference_example_main
#include UNGAR_EXAMPLE_SOURCE
#undef main

int main() {
    try {
        return ungar_reference_example_main();
    } catch (const osqp::StopRequested&) {
        std::printf("example_driver: stopped after the requested number of QP solves\n");
        return 0;
    }
}
