// Runs one UNCHANGED reference example (example/mpc/*.example.cpp, compiled from where it lies) against the PRODUCT headers
// (ungar_b200/include: cppad/cg.hpp, osqp++.h): the example's own lambdas are taped by the reference's own MakeFunction and
// evaluated by the register machine on the GPU, its local QPs are solved on the device, and the endless receding-horizon loop is
// stopped after UNGAR_B200_MAX_QP_SOLVES solves.
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
