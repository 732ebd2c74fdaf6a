// Lockstep emulation of one warp on 32 host threads (no GPU): runs the REAL qt_cholesky / qt_trsm / qt_outward_solve code of
// ungar_b200/csrc/qp_twisted.cuh (shuffles and __syncwarp through a std::barrier) against plain dense loops.  Built and run by
// tests/test_qp_host.py.
#include <barrier>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
struct double2 { double x, y; };
using std::min; using std::max;
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static std::barrier<> g_bar(32);
static double g_sh[32];
static thread_local int t_lane;
static inline void __syncwarp() { g_bar.arrive_and_wait(); }
static inline double __shfl_sync(unsigned, double v, int src) { g_sh[t_lane] = v; g_bar.arrive_and_wait(); double r = g_sh[src]; g_bar.arrive_and_wait(); return r; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
namespace ub {
inline double2 qp_ld2(const double* p) { return {p[0], p[1]}; }
inline void qp_st2(double* p, double a, double b) { p[0] = a; p[1] = b; }
inline void qp_dmma(double&, double&, double, double) {}
}
#include "compact.cuh"
#define QT_HOST_TEST
#include "qp_twisted.cuh"
using namespace ub;
static double rnd() { return 2.0 * rand() / RAND_MAX - 1.0; }
int main() {
    srand(5);
    const int G = 29;
    double M[29][29], S[29][29], E[29][29], rhs[29], x[29];
    for (auto& r : M) for (auto& v : r) v = rnd();
    for (int i = 0; i < G; ++i) for (int j = 0; j < G; ++j) { S[i][j] = (i == j) * 2.0; for (int k = 0; k < G; ++k) S[i][j] += M[i][k] * M[j][k]; }
    for (auto& r : E) for (auto& v : r) v = rnd();
    for (auto& v : rhs) v = rnd();
    for (auto& v : x) v = rnd();
    // reference
    double L[29][29] = {};
    for (int c = 0; c < G; ++c) { double d = S[c][c]; for (int k = 0; k < c; ++k) d -= L[c][k] * L[c][k]; L[c][c] = sqrt(d); for (int r = c + 1; r < G; ++r) { double v = S[r][c]; for (int k = 0; k < c; ++k) v -= L[r][k] * L[c][k]; L[r][c] = v / L[c][c]; } }
    double y[29]; for (int i = 0; i < G; ++i) { double v = rhs[i]; for (int k = 0; k < i; ++k) v -= L[i][k] * y[k]; y[i] = v / L[i][i]; }
    double Lo[29][29]; for (int i = 0; i < G; ++i) for (int c = 0; c < G; ++c) { double v = E[i][c]; for (int k = 0; k < c; ++k) v -= Lo[i][k] * L[c][k]; Lo[i][c] = v / L[c][c]; }
    double z[29]; for (int i = 0; i < G; ++i) { double v = x[i]; for (int k = 0; k < i; ++k) v -= L[i][k] * z[k]; z[i] = v / L[i][i]; }
    double nu[29]; for (int i = G - 1; i >= 0; --i) { double v = y[i] - z[i]; for (int k = i + 1; k < G; ++k) v -= L[k][i] * nu[k]; nu[i] = v / L[i][i]; }
    std::vector<double> img(1024, 0.0), ws(1024, 0.0), stash(4, 0.0);
    double yo[32], loo[32][29], nuo[32];
    std::vector<std::thread> th;
    for (int lane = 0; lane < 32; ++lane) th.emplace_back([&, lane] {
        t_lane = lane;
        double s[29], e[29];
        for (int c = 0; c < G; ++c) { s[c] = lane < G ? S[lane][c] : 0.0; e[c] = lane < G ? E[lane][c] : 0.0; }
        yo[lane] = qt_cholesky(s, lane < G ? rhs[lane] : 0.0, img.data(), ws.data(), stash.data(), lane);
        if (lane < G) { ws[QpT::fY + lane] = yo[lane]; img[QpT::fY + lane] = yo[lane]; }
        __syncwarp();
        qt_trsm(e, img.data());
        for (int c = 0; c < G; ++c) loo[lane][c] = e[c];
        __syncwarp();
        nuo[lane] = qt_outward_solve(ws.data(), lane < G ? x[lane] : 0.0, lane);
    });
    for (auto& t : th) t.join();
    double wy = 0, wl = 0, wn = 0, wf = 0;
    for (int i = 0; i < G; ++i) { wy = std::max(wy, fabs(yo[i] - y[i])); wn = std::max(wn, fabs(nuo[i] - nu[i])); for (int c = 0; c < G; ++c) wl = std::max(wl, fabs(loo[i][c] - Lo[i][c])); }
    for (int c = 0; c < G; ++c) for (int r = c; r < G; ++r) wf = std::max(wf, fabs(ws[QpT::bc(c) + r] * ws[QpT::fRI + c] - L[r][c]));
    printf("y %.2e  Lo %.2e  nu %.2e  factor %.2e\n", wy, wl, wn, wf);
    return (wy + wl + wn + wf) < 1e-9 ? 0 : 1;
}
