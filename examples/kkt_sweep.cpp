// Minimal C++ user of the drop-in boundary: builds the quadruped NMPC problem (example/mpc/quadruped.example.cpp at
// N = 30), evaluates its three functions in the reference's format at the example's initial guess, runs one batched
// KKT sweep and one soft-SQP solve on the device.  Build:  g++ -std=c++17 -Iinclude -Iungar_b200/include examples/kkt_sweep.cpp ungar_b200/libungar_b200.so
#include <cmath>
#include <cstdio>
#include <vector>

#include "ungar_b200/function.hpp"

int main() {
    using namespace ungar_b200;
    constexpr int N = 30;
    try {
        Model<double> model(UNGAR_B200_QUADRUPED, N, /*stiffness*/ 1.0, /*epsilon*/ 1.0);
        const auto& L = model.layout();
        // the example's initial guess (quadruped.example.cpp:378-430): stance, f_i = m g / 4 e_z
        std::vector<double> xp(static_cast<std::size_t>(model.VariableSize()), 0.0);
        const index_t U0 = 13 * (N + 1), P0 = L.n_dec, Rho = P0 + 29 * (N + 1);
        const double hips[4][3] = {{0.2, 0.15, -0.1}, {0.2, -0.15, -0.1}, {-0.2, 0.15, -0.1}, {-0.2, -0.15, -0.1}};
        const double feet[4][3] = {{0.2, 0.1, 0.0}, {0.2, -0.1, 0.0}, {-0.2, 0.1, 0.0}, {-0.2, -0.1, 0.0}};
        xp[Rho] = 1.0 / N; xp[Rho + 1] = 25.0; xp[Rho + 2] = 0.048125; xp[Rho + 3] = 0.093125; xp[Rho + 4] = 0.055625;
        xp[Rho + 17] = 0.42; xp[Rho + 18] = 9.80665; xp[Rho + 19] = 0.7; xp[Rho + 22] = 0.38; xp[Rho + 26] = 1.0;
        for (int i = 0; i < 4; ++i) {
            for (int c = 0; c < 3; ++c) { xp[Rho + 5 + 3 * i + c] = hips[i][c]; xp[Rho + 34 + 4 * i + c] = feet[i][c]; }
            xp[Rho + 33 + 4 * i] = 1.0;
        }
        for (int k = 0; k <= N; ++k) {
            xp[13 * k + 2] = 0.38; xp[13 * k + 6] = 1.0;
            xp[P0 + 29 * k + 2] = 0.38; xp[P0 + 29 * k + 6] = 1.0;
            for (int i = 0; i < 4; ++i) {
                xp[P0 + 29 * k + 13 + 4 * i] = 1.0;
                for (int c = 0; c < 3; ++c) xp[P0 + 29 * k + 14 + 4 * i + c] = hips[i][c] - (c == 2 ? 0.8 * 0.42 : 0.0);
            }
        }
        for (int k = 0; k < N; ++k)
            for (int i = 0; i < 4; ++i) {
                xp[U0 + 24 * k + 6 * i + 2] = 25.0 * 9.80665 / 4.0;
                for (int c = 0; c < 3; ++c) xp[U0 + 24 * k + 6 * i + 3 + c] = feet[i][c];
            }
        Function<double> eqs = model.equalityConstraints();
        const std::vector<double> g = eqs(xp);
        double defect = 0.0;
        for (index_t i = 0; i < 13 + 13 * N; ++i) defect = std::fmax(defect, std::fabs(g[i]));
        const std::vector<double>& J = eqs.Jacobian(xp);
        std::printf("quadruped N=%d: m_eq=%lld nnz(J_g)=%lld  max dynamics defect at the stance guess = %.3e  foot row = %.2f\n", N,
                    (long long)eqs.DependentVariableSize(), (long long)J.size(), defect, g[13 + 13 * N + 3]);
        std::vector<double> rec(static_cast<std::size_t>(L.size));
        model.KktBlocks(xp.data(), 1, rec.data());
        std::printf("KKT record: %lld scalars, objective f = %.6f, barrier = %.6f\n", (long long)L.size, rec[L.cost], rec[L.cost + 1]);
        // one MPC solve like quadruped.example.cpp:444, :512-522: the measured base is 3 cm lower than the stance guess
        std::vector<double> xq(xp);
        xq[Rho + 22] = 0.35;
        SoftSQPOptimizer optimizer{false, 1.0 / N, 4, 1.0, 1.0};
        const std::vector<double> sol = optimizer.Optimize(model, xq);
        const double* info = optimizer.LineSearchInfo().data();
        std::printf("soft SQP: status %d after %d iterations, last step size %.4g, constraint violation %.3e, base height x_0 = %.4f\n",
                    optimizer.Status()[0].status, optimizer.Status()[0].iterations, info[0], info[1], sol[2]);
        const bool sqp_ok = optimizer.Status()[0].iterations >= 1 && sol[2] < 0.38 - 1e-4;  // x_0 moved towards the measurement
        return (defect < 1e-12 && std::fabs(g[13 + 13 * N + 3] - 0.38) < 1e-12 && sqp_ok) ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
