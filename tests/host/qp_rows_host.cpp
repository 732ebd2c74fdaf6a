// Host harness (no GPU): the per-lane row algebra of ungar_b200/csrc/qp_twisted.cuh — U / V rows in their register forms, P^-1, the
// products U P^-1 U^T, V P^-1 U^T, V P^-1 V^T, U P^-1 V^T, the gathers V^T nu / U^T nu — against dense products built independently
// from the same compact chunk (csrc/compact.cuh).  Built and run by tests/test_qp_host.py.
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
#define __global__
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { return {a, b}; }
using std::min; using std::max;
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline double __shfl_sync(unsigned, double v, int) { return v; }
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
static inline void __syncwarp() {}
static inline void __syncthreads() {}
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
struct dim3i { int x; }; static dim3i threadIdx{0}, blockIdx{0};
namespace ub {
inline double2 qp_ld2(const double* p) { return {p[0], p[1]}; }
inline void qp_st2(double* p, double a, double b) { p[0] = a; p[1] = b; }
inline void qp_dmma(double&, double&, double, double) {}
}
#include "compact.cuh"
#define QT_HOST_TEST
#include "qp_twisted.cuh"
using namespace ub;
using K = Compact;
static double rnd() { return 2.0 * rand() / RAND_MAX - 1.0; }
int main() {
    srand(3);
    std::vector<double> sm(K::SMALL), smn(K::SMALL), ab(K::APART);
    for (auto& v : sm) v = rnd();
    for (auto& v : smn) v = rnd();
    for (auto& v : ab) v = rnd();
    for (int qr = 0; qr < 7; ++qr) ab[qr * 32 + 7] = 0.0;
    // SPD P: diag + blocks
    double P[37][37] = {};
    for (int i = 0; i < 13; ++i) { sm[K::oHd + i] = 1.0 + fabs(rnd()); P[i][i] = sm[K::oHd + i]; }
    for (int b = 0; b < 8; ++b) {
        double M[3][3];
        for (auto& r : M) for (auto& v : r) v = rnd();
        double S[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { S[i][j] = (i == j) * 1.0; for (int k = 0; k < 3; ++k) S[i][j] += M[i][k] * M[j][k]; }
        double* h = &sm[K::oHb + 6 * b];
        h[0] = S[0][0]; h[1] = S[0][1]; h[2] = S[0][2]; h[3] = S[1][1]; h[4] = S[1][2]; h[5] = S[2][2];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) P[13 + 3 * b + i][13 + 3 * b + j] = S[i][j];
    }
    double q[37]; for (int i = 0; i < 37; ++i) q[i] = sm[K::oQ + i];
    // dense U (29x37) from sm's Cs, V (29x37) from ab and smn's Cp
    double U[29][37] = {}, V[29][37] = {};
    for (int i = 0; i < 13; ++i) U[i][i] = 1.0;
    for (int l = 0; l < 4; ++l) for (int rr = 0; rr < 4; ++rr) {
        const double* cs = &sm[K::oCs + (l * 4 + rr) * 8];
        const int pc = rr == 0 ? 2 : rr - 1, row = 13 + 4 * l + rr;
        U[row][pc] = cs[0];
        for (int i = 0; i < 4; ++i) U[row][3 + i] = cs[1 + i];
        for (int i = 0; i < 3; ++i) U[row][13 + 6 * l + 3 + i] = cs[5 + i];
        if (rr > 0) {
            const double* cp = &smn[K::oCp + (l * 3 + rr - 1) * 8];
            V[row][pc] = cp[0];
            for (int i = 0; i < 4; ++i) V[row][3 + i] = cp[1 + i];
            for (int i = 0; i < 3; ++i) V[row][13 + 6 * l + 3 + i] = cp[5 + i];
        }
    }
    for (int row = 0; row < 13; ++row) {
        const int qr = K::aq_row(row);
        if (qr >= 0) { for (int z = 0; z < 37; ++z) if (K::aq_col(z) >= 0) V[row][z] = ab[qr * 32 + K::aq_col(z)]; }
        else if (row < 3) { const double* ap = &ab[224 + row * 6]; V[row][row] = ap[0]; V[row][7 + row] = ap[1]; for (int l = 0; l < 4; ++l) V[row][13 + 6 * l + row] = ap[2 + l]; }
        else { const int c = row - 7; const double* ap = &ab[224 + (3 + c) * 6]; V[row][row] = ap[0]; for (int l = 0; l < 4; ++l) V[row][13 + 6 * l + c] = ap[1 + l]; }
    }
    // dense inverse of P by Gauss-Jordan
    double Pi[37][37], A[37][74];
    for (int i = 0; i < 37; ++i) for (int j = 0; j < 37; ++j) { A[i][j] = P[i][j]; A[i][37 + j] = i == j; }
    for (int c = 0; c < 37; ++c) { double p = A[c][c]; for (int j = 0; j < 74; ++j) A[c][j] /= p; for (int r = 0; r < 37; ++r) if (r != c) { double f = A[r][c]; for (int j = 0; j < 74; ++j) A[r][j] -= f * A[c][j]; } }
    for (int i = 0; i < 37; ++i) for (int j = 0; j < 37; ++j) Pi[i][j] = A[i][37 + j];
    double t[37]; for (int i = 0; i < 37; ++i) { t[i] = 0; for (int j = 0; j < 37; ++j) t[i] += Pi[i][j] * q[j]; }
    auto quad = [&](double X[29][37], int i, double Y[29][37], int c) { double a = 0; for (int k = 0; k < 37; ++k) for (int l2 = 0; l2 < 37; ++l2) a += X[i][k] * Pi[k][l2] * Y[c][l2]; return a; };
    // in-place P^-1, t
    for (int lane = 0; lane < 32; ++lane) qt_pinv_t(sm.data(), lane, true);
    double worst[8] = {};
    for (int i = 0; i < 37; ++i) worst[0] = std::max(worst[0], fabs(sm[K::oQ + i] - t[i]));
    for (int lane = 0; lane < 32; ++lane) {
        QtLane L(lane);
        double s[29] = {}, e[29], s2[29], e2[29];
        { QtU u; qt_load_u(sm.data(), lane, L, u); double ut = qt_u_dot(u, sm.data() + K::oQ); qt_u_pinv(u, sm.data()); qt_u_dot_u_rows(u, sm.data(), s); qt_u_dot_v_rows(u, ab.data(), smn.data(), e2);
          if (lane < 29) { double r = 0; for (int k = 0; k < 37; ++k) r += U[lane][k] * t[k]; worst[1] = std::max(worst[1], fabs(ut - r)); } }
        double vt = qt_vpu(ab.data(), smn.data(), sm.data(), L, e);
        double vt2 = qt_vpv<true>(ab.data(), smn.data(), sm.data(), L, s2);
        if (lane < 29) {
            double r = 0; for (int k = 0; k < 37; ++k) r += V[lane][k] * t[k];
            worst[2] = std::max(worst[2], std::max(fabs(vt - r), fabs(vt2 - r)));
            for (int c = 0; c < 29; ++c) {
                if (c <= lane || true) worst[3] = std::max(worst[3], fabs(s[c] - quad(U, lane, U, c)));
                worst[4] = std::max(worst[4], fabs(e[c] - quad(V, lane, U, c)));
                worst[5] = std::max(worst[5], fabs(s2[c] - quad(V, lane, V, c)));
                worst[6] = std::max(worst[6], fabs(e2[c] - quad(U, lane, V, c)));
            }
        } else for (int c = 0; c < 29; ++c) worst[7] = std::max(worst[7], fabs(s[c]) + fabs(e[c]) + fabs(s2[c]) + fabs(e2[c]));
    }
    // gathers
    double nu[32]; for (auto& v : nu) v = rnd();
    double wg = 0;
    for (int k = 0; k < 37; ++k) {
        double a = 0, b = 0; for (int r = 0; r < 29; ++r) { a += V[r][k] * nu[r]; b += U[r][k] * nu[r]; }
        wg = std::max(wg, std::max(fabs(qt_vT_nu(ab.data(), smn.data(), nu, k) - a), fabs(qt_uT_nu(sm.data(), nu, k) - b)));
    }
    double x[37]; for (auto& v : x) v = rnd();
    double wd = 0;
    for (int lane = 0; lane < 29; ++lane) { QtLane L(lane); double r = 0; for (int k = 0; k < 37; ++k) r += V[lane][k] * x[k]; wd = std::max(wd, fabs(qt_v_dot(ab.data(), smn.data(), L, x) - r)); }
    printf("t %.2e | U.t %.2e | V.t %.2e | UPU %.2e | VPU %.2e | VPV %.2e | UPV %.2e | idle lanes %.2e | gathers %.2e | Vx %.2e\n", worst[0], worst[1], worst[2], worst[3], worst[4], worst[5], worst[6], worst[7], wg, wd);
    double tot = 0; for (double w : worst) tot += w; tot += wg + wd;
    return tot < 1e-9 ? 0 : 1;
}
