// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path.
//
// CPU port of the loop that consumes the KKT block records (quadruped: stage-wise Schur complement; quadrotor, RC car: Riccati) — the timed CPU baseline of the `sqp_loop`
// figure of bench.py and a third implementation for the parity tests (next to oracle/sqp_reference.py and the device kernels):
//   * oracle_qp_solve   the equality-constrained QP SoftSQPOptimizer hands to OSQP (include/ungar/optimization/soft_sqp.hpp:141-158,
//                       :193-233), solved exactly by the stage-wise Schur complement of oracle/qp_reference.py::schur_stagewise
//                       (OSQP v0.6.3 itself is absent: external/config/osqp/CMakeLists.txt.in:16);
//   * oracle_sqp_solve  SoftSQPOptimizer::Optimize (soft_sqp.hpp:63-109) with BacktrackingLineSearch::Do
//                       (include/ungar/optimization/backtracking_line_search.hpp:81-165): stage sweep (stage_port.cpp) -> QP ->
//                       line search on the values of the three functions (oracle.cpp), trajectories split over host threads.
// Plain dense block algebra with the structure of the problem used where it is free (P is diagonal plus 3x3 blocks, the state
// rows of U are the identity); it is a straightforward port, not a tuned solver.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {
int oracle_stage_sweep(int model, int N, const double* xp, int64_t batch, int64_t ld_xp, double stiffness, double epsilon,
                       double* records, int64_t ld_rec, int threads);
int oracle_record_layout(int model, int N, int* layout);
int oracle_sizes(int model, int N, int* sizes);
int oracle_eval(int model, int fn, int N, const double* xp, double* y);
int oracle_directional(int model, int fn, int N, const double* xp, const double* dir, double* y, double* dy);
}

namespace {

constexpr int QUADRUPED = 2, NX = 13, NU = 24, NZ = 37, TRI = 703, G = 29;

struct Lay {
    int g, A, C, h, cost, grad, H, HN, Hc, size;
};

Lay layout(int model, int N) {
    int v[15];
    oracle_record_layout(model, N, v);
    return Lay{v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9]};
}

inline int tri(int n, int i, int j) { return i * n - (i * (i - 1)) / 2 + (j - i); }  // i <= j

// y = P^-1 x for the stage Hessian H (packed upper triangle): 13 diagonal state entries + eight 3x3 input blocks; `nz` = 37 or 13.
struct PInv {
    double d[NX];
    double B[8][9];
    int nz;
    void build(const double* Hp, int n, int nzj) {
        nz = nzj;
        for (int i = 0; i < NX; ++i) d[i] = 1.0 / Hp[tri(n, i, i)];
        if (nz == NX) return;
        for (int b = 0; b < 8; ++b) {
            const int a = NX + 3 * b;
            const double m00 = Hp[tri(n, a, a)], m01 = Hp[tri(n, a, a + 1)], m02 = Hp[tri(n, a, a + 2)], m11 = Hp[tri(n, a + 1, a + 1)],
                         m12 = Hp[tri(n, a + 1, a + 2)], m22 = Hp[tri(n, a + 2, a + 2)];
            const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
            const double id = 1.0 / (m00 * c00 + m01 * c01 + m02 * c02);
            double* o = B[b];
            o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
            o[3] = o[1]; o[4] = (m00 * m22 - m02 * m02) * id; o[5] = (m01 * m02 - m00 * m12) * id;
            o[6] = o[2]; o[7] = o[5]; o[8] = (m00 * m11 - m01 * m01) * id;
        }
    }
    void apply(const double* x, double* y) const {
        for (int i = 0; i < NX; ++i) y[i] = d[i] * x[i];
        if (nz == NX) return;
        for (int b = 0; b < 8; ++b) {
            const double* o = B[b];
            const double* xv = x + NX + 3 * b;
            double* yv = y + NX + 3 * b;
            yv[0] = o[0] * xv[0] + o[1] * xv[1] + o[2] * xv[2];
            yv[1] = o[3] * xv[0] + o[4] * xv[1] + o[5] * xv[2];
            yv[2] = o[6] * xv[0] + o[7] * xv[1] + o[8] * xv[2];
        }
    }
};

// Contact rows of stage k on w_k (prev = false) or on w_{k-1} (prev = true) as a dense 16 x 37 matrix.
void contact_matrix(const double* rec, const Lay& L, int k, bool prev, double* M) {
    std::fill(M, M + 16 * NZ, 0.0);
    const double* Ck = rec + L.C + (int64_t)k * 320;
    const int off = prev ? 10 : 0;
    for (int leg = 0; leg < 4; ++leg)
        for (int r = 0; r < 4; ++r) {
            const double* row = Ck + (leg * 4 + r) * 20 + off;
            double* out = M + (leg * 4 + r) * NZ;
            for (int c = 0; c < 7; ++c) out[c] = row[c];
            for (int c = 0; c < 3; ++c) out[NX + 6 * leg + 3 + c] = row[7 + c];
        }
}

// Exact solve of  min 1/2 d^T P d + q^T d  s.t.  A d = -g  for one record (block-tridiagonal Schur complement on the multipliers).
void qp_schur_cpu(const double* rec, const Lay& L, int N, double delta, double* d) {
    const int nX = NX * (N + 1);
    std::vector<double> Ld((size_t)(N + 1) * G * G, 0.0), Lo((size_t)(N + 1) * G * G, 0.0), y((size_t)(N + 1) * G, 0.0),
        nu((size_t)(N + 2) * G, 0.0);
    std::vector<double> U(G * NZ), V(G * NZ), Vprev(G * NZ), XU(G * NZ), XV(G * NZ), XVprev(G * NZ), S(G * G), E(G * G), q(NZ), t(NZ), tprev(NZ);
    std::vector<PInv> pinv(N + 1);
    int rows_prev = 0;
    for (int j = 0; j <= N; ++j) {
        const int nz = j < N ? NZ : NX, rows = j < N ? G : NX;
        pinv[j].build(j < N ? rec + L.H + (int64_t)j * TRI : rec + L.HN, nz, nz);
        // U_j = [I_x 0; Cs_j],  V_j = [A_j; Cp_{j+1}]
        std::fill(U.begin(), U.end(), 0.0);
        for (int i = 0; i < NX; ++i) U[i * NZ + i] = 1.0;
        if (j < N) contact_matrix(rec, L, j, false, U.data() + NX * NZ);
        for (int i = 0; i < nz; ++i) q[i] = i < NX ? rec[L.grad + NX * j + i] : rec[L.grad + nX + NU * j + (i - NX)];
        pinv[j].apply(q.data(), t.data());
        for (int r = 0; r < rows; ++r) pinv[j].apply(U.data() + r * NZ, XU.data() + r * NZ);  // rows of U P^-1
        // S_jj = U P^-1 U^T + V_{j-1} P_{j-1}^-1 V_{j-1}^T + delta I ;  rhs_j = -(b_j + U t_j + V_{j-1} t_{j-1}),  b_j = -g_j
        std::vector<double> rhs(G, 0.0);
        for (int r = 0; r < rows; ++r) {
            for (int c = 0; c <= r; ++c) {
                double acc = r == c ? delta : 0.0;
                for (int k = 0; k < nz; ++k) acc += XU[r * NZ + k] * U[c * NZ + k];
                if (j > 0)
                    for (int k = 0; k < NZ; ++k) acc += XVprev[r * NZ + k] * Vprev[c * NZ + k];
                S[r * G + c] = acc;
            }
            const double gj = r < NX ? rec[L.g + NX * j + r] : rec[L.g + nX + 16 * j + (r - NX)];
            double acc = -gj;
            for (int k = 0; k < nz; ++k) acc += U[r * NZ + k] * t[k];
            if (j > 0)
                for (int k = 0; k < NZ; ++k) acc += Vprev[r * NZ + k] * tprev[k];
            rhs[r] = -acc;
        }
        double* Lj = Ld.data() + (size_t)j * G * G;
        double* Loj = Lo.data() + (size_t)j * G * G;
        if (j > 0) {
            // E = S_{j,j-1} = V_{j-1} P_{j-1}^-1 U_{j-1}^T was prepared below; Lo_j = E L_{j-1}^-T, then S -= Lo Lo^T, rhs -= Lo y_{j-1}
            const double* Lp = Ld.data() + (size_t)(j - 1) * G * G;
            for (int r = 0; r < rows; ++r) {
                for (int c = 0; c < rows_prev; ++c) {
                    double acc = E[r * G + c];
                    for (int k = 0; k < c; ++k) acc -= Loj[r * G + k] * Lp[c * G + k];
                    Loj[r * G + c] = acc / Lp[c * G + c];
                }
                double acc = 0.0;
                for (int c = 0; c < rows_prev; ++c) acc += Loj[r * G + c] * y[(size_t)(j - 1) * G + c];
                rhs[r] -= acc;
            }
            for (int r = 0; r < rows; ++r)
                for (int c = 0; c <= r; ++c) {
                    double acc = 0.0;
                    for (int k = 0; k < rows_prev; ++k) acc += Loj[r * G + k] * Loj[c * G + k];
                    S[r * G + c] -= acc;
                }
        }
        for (int c = 0; c < rows; ++c) {  // Cholesky of S_jj and forward substitution y_j = L_jj^-1 rhs_j
            double acc = S[c * G + c];
            for (int k = 0; k < c; ++k) acc -= Lj[c * G + k] * Lj[c * G + k];
            const double piv = std::sqrt(acc);
            Lj[c * G + c] = piv;
            for (int r = c + 1; r < rows; ++r) {
                double a = S[r * G + c];
                for (int k = 0; k < c; ++k) a -= Lj[r * G + k] * Lj[c * G + k];
                Lj[r * G + c] = a / piv;
            }
        }
        for (int r = 0; r < rows; ++r) {
            double acc = rhs[r];
            for (int k = 0; k < r; ++k) acc -= Lj[r * G + k] * y[(size_t)j * G + k];
            y[(size_t)j * G + r] = acc / Lj[r * G + r];
        }
        if (j < N) {  // prepare stage j + 1: V_j, V_j P_j^-1 and E = V_j P_j^-1 U_j^T
            const int rows_next = j + 1 < N ? G : NX;
            std::fill(V.begin(), V.end(), 0.0);
            for (int r = 0; r < NX; ++r)
                for (int k = 0; k < NZ; ++k) V[r * NZ + k] = rec[L.A + (int64_t)j * NX * NZ + r * NZ + k];
            if (j + 1 < N) contact_matrix(rec, L, j + 1, true, V.data() + NX * NZ);
            for (int r = 0; r < rows_next; ++r) pinv[j].apply(V.data() + r * NZ, XV.data() + r * NZ);
            for (int r = 0; r < rows_next; ++r)
                for (int c = 0; c < rows; ++c) {
                    double acc = 0.0;
                    for (int k = 0; k < NZ; ++k) acc += XV[r * NZ + k] * U[c * NZ + k];
                    E[r * G + c] = acc;
                }
            Vprev = V;
            XVprev = XV;
            tprev = t;
        }
        rows_prev = rows;
    }
    // backward substitution: nu_j = L_jj^-T (y_j - Lo_{j+1}^T nu_{j+1}), then d_j = -P_j^-1 (q_j + U_j^T nu_j + V_j^T nu_{j+1})
    std::vector<double> w(NZ), v(NZ);
    for (int j = N; j >= 0; --j) {
        const int nz = j < N ? NZ : NX, rows = j < N ? G : NX, rows_next = j + 1 < N ? G : (j < N ? NX : 0);
        const double* Lj = Ld.data() + (size_t)j * G * G;
        double* nj = nu.data() + (size_t)j * G;
        const double* nn = nu.data() + (size_t)(j + 1) * G;
        for (int r = rows - 1; r >= 0; --r) {
            double acc = y[(size_t)j * G + r];
            if (j < N) {
                const double* Lon = Lo.data() + (size_t)(j + 1) * G * G;
                for (int k = 0; k < rows_next; ++k) acc -= Lon[k * G + r] * nn[k];
            }
            for (int k = r + 1; k < rows; ++k) acc -= Lj[k * G + r] * nj[k];
            nj[r] = acc / Lj[r * G + r];
        }
        // w = q_j + U_j^T nu_j + V_j^T nu_{j+1}
        for (int i = 0; i < nz; ++i) w[i] = i < NX ? rec[L.grad + NX * j + i] + nj[i] : rec[L.grad + nX + NU * j + (i - NX)];
        if (j < N) {
            std::vector<double> Cs(16 * NZ);
            contact_matrix(rec, L, j, false, Cs.data());
            for (int r = 0; r < 16; ++r)
                for (int k = 0; k < NZ; ++k) w[k] += Cs[r * NZ + k] * nj[NX + r];
            for (int r = 0; r < NX; ++r)
                for (int k = 0; k < NZ; ++k) w[k] += rec[L.A + (int64_t)j * NX * NZ + r * NZ + k] * nn[r];
            if (j + 1 < N) {
                contact_matrix(rec, L, j + 1, true, Cs.data());
                for (int r = 0; r < 16; ++r)
                    for (int k = 0; k < NZ; ++k) w[k] += Cs[r * NZ + k] * nn[NX + r];
            }
        }
        pinv[j].apply(w.data(), v.data());
        for (int i = 0; i < nz; ++i) {
            if (i < NX) d[NX * j + i] = -v[i];
            else d[nX + NU * j + (i - NX)] = -v[i];
        }
    }
}

// Quadrotor / RC car: the only equalities are the initial condition and the defects, and the objective couples u_k with u_{k+1}
// (the Hc blocks).  Riccati recursion on the state augmented with the previous input — the algebra of ungar_b200/csrc/qp_riccati.cuh
// written with plain dense loops.
template <int PX, int PU>
void qp_riccati_cpu(const double* rec, const Lay& L, int N, double* d) {
    constexpr int PZ = PX + PU, PS = PX + PU, PT = PZ * (PZ + 1) / 2;
    const int uoff = PX * (N + 1);
    std::vector<double> Pxx(PX * PX), Pxv(PX * PU, 0.0), Pvv(PU * PU, 0.0), px(PX), pv(PU, 0.0), Y((size_t)N * PU * (PS + 1));
    for (int i = 0; i < PX; ++i) {
        for (int j = 0; j < PX; ++j) Pxx[i * PX + j] = rec[L.HN + (i <= j ? tri(PX, i, j) : tri(PX, j, i))];
        px[i] = rec[L.grad + PX * N + i];
    }
    std::vector<double> W(PX * PZ), T(PZ * PU), M(PZ * PZ), m(PZ), wx(PX), wv(PU), Lc(PU * PU), nPxx(PX * PX), nPxv(PX * PU), nPvv(PU * PU);
    for (int k = N - 1; k >= 0; --k) {
        const double* A = rec + L.A + (int64_t)k * PX * PZ;
        const double* H = rec + L.H + (int64_t)k * PT;
        const double* g = rec + L.g + PX * (k + 1);
        double D[PU];
        for (int a = 0; a < PU; ++a) D[a] = k > 0 ? rec[L.Hc + PU * (k - 1) + a] : 0.0;
        for (int i = 0; i < PX; ++i) {
            double acc = px[i];
            for (int r = 0; r < PX; ++r) acc -= Pxx[i * PX + r] * g[r];
            wx[i] = acc;
        }
        for (int a = 0; a < PU; ++a) {
            double acc = pv[a];
            for (int r = 0; r < PX; ++r) acc -= Pxv[r * PU + a] * g[r];
            wv[a] = acc;
        }
        for (int i = 0; i < PX; ++i)
            for (int j = 0; j < PZ; ++j) {
                double acc = 0.0;
                for (int r = 0; r < PX; ++r) acc += Pxx[i * PX + r] * A[r * PZ + j];
                W[i * PZ + j] = acc;
            }
        for (int i = 0; i < PZ; ++i)
            for (int a = 0; a < PU; ++a) {
                double acc = 0.0;
                for (int r = 0; r < PX; ++r) acc += A[r * PZ + i] * Pxv[r * PU + a];
                T[i * PU + a] = acc;
            }
        for (int i = 0; i < PZ; ++i) {
            for (int j = 0; j < PZ; ++j) {
                double acc = H[i <= j ? tri(PZ, i, j) : tri(PZ, j, i)];
                for (int r = 0; r < PX; ++r) acc += A[r * PZ + i] * W[r * PZ + j];
                if (j >= PX) acc -= T[i * PU + (j - PX)];
                if (i >= PX) acc -= T[j * PU + (i - PX)];
                if (i >= PX && j >= PX) acc += Pvv[(i - PX) * PU + (j - PX)];
                M[i * PZ + j] = acc;
            }
            double acc = i < PX ? rec[L.grad + PX * k + i] : rec[L.grad + uoff + PU * k + (i - PX)];
            for (int r = 0; r < PX; ++r) acc -= A[r * PZ + i] * wx[r];
            if (i >= PX) acc += wv[i - PX];
            m[i] = acc;
        }
        for (int i = 0; i < PU; ++i)  // Cholesky of M_uu
            for (int j = 0; j <= i; ++j) {
                double acc = M[(PX + i) * PZ + PX + j];
                for (int r = 0; r < j; ++r) acc -= Lc[i * PU + r] * Lc[j * PU + r];
                Lc[i * PU + j] = i == j ? std::sqrt(acc) : acc / Lc[j * PU + j];
            }
        double* Yk = Y.data() + (size_t)k * PU * (PS + 1);  // Y = M_uu^-1 [M_ux | D | m_u]
        for (int c = 0; c <= PS; ++c) {
            double y[PU];
            for (int a = 0; a < PU; ++a) y[a] = c < PX ? M[(PX + a) * PZ + c] : c < PS ? (a == c - PX ? D[a] : 0.0) : m[PX + a];
            for (int a = 0; a < PU; ++a) {
                for (int r = 0; r < a; ++r) y[a] -= Lc[a * PU + r] * y[r];
                y[a] /= Lc[a * PU + a];
            }
            for (int a = PU - 1; a >= 0; --a) {
                for (int r = a + 1; r < PU; ++r) y[a] -= Lc[r * PU + a] * y[r];
                y[a] /= Lc[a * PU + a];
            }
            for (int a = 0; a < PU; ++a) Yk[a * (PS + 1) + c] = y[a];
        }
        for (int i = 0; i < PX; ++i) {
            for (int j = 0; j < PX; ++j) {
                double acc = M[i * PZ + j];
                for (int a = 0; a < PU; ++a) acc -= M[i * PZ + PX + a] * Yk[a * (PS + 1) + j];
                nPxx[i * PX + j] = acc;
            }
            for (int b = 0; b < PU; ++b) {
                double acc = 0.0;
                for (int a = 0; a < PU; ++a) acc -= M[i * PZ + PX + a] * Yk[a * (PS + 1) + PX + b];
                nPxv[i * PU + b] = acc;
            }
            double acc = m[i];
            for (int a = 0; a < PU; ++a) acc -= M[i * PZ + PX + a] * Yk[a * (PS + 1) + PS];
            px[i] = acc;
        }
        for (int a = 0; a < PU; ++a) {
            for (int b = 0; b < PU; ++b) nPvv[a * PU + b] = -D[a] * Yk[a * (PS + 1) + PX + b];
            pv[a] = -D[a] * Yk[a * (PS + 1) + PS];
        }
        Pxx = nPxx; Pxv = nPxv; Pvv = nPvv;
    }
    double sv[PS];  // forward rollout from s_0 = [-g_0; 0]
    for (int i = 0; i < PS; ++i) sv[i] = i < PX ? -rec[L.g + i] : 0.0;
    for (int k = 0; k < N; ++k) {
        const double* A = rec + L.A + (int64_t)k * PX * PZ;
        const double* Yk = Y.data() + (size_t)k * PU * (PS + 1);
        double du[PU], nx[PX];
        for (int a = 0; a < PU; ++a) {
            double acc = -Yk[a * (PS + 1) + PS];
            for (int c = 0; c < PS; ++c) acc -= Yk[a * (PS + 1) + c] * sv[c];
            du[a] = acc;
        }
        for (int i = 0; i < PX; ++i) d[PX * k + i] = sv[i];
        for (int a = 0; a < PU; ++a) d[uoff + PU * k + a] = du[a];
        for (int i = 0; i < PX; ++i) {
            double acc = -rec[L.g + PX * (k + 1) + i];
            for (int c = 0; c < PX; ++c) acc -= A[i * PZ + c] * sv[c];
            for (int a = 0; a < PU; ++a) acc -= A[i * PZ + PX + a] * du[a];
            nx[i] = acc;
        }
        for (int i = 0; i < PX; ++i) sv[i] = nx[i];
        for (int a = 0; a < PU; ++a) sv[PX + a] = du[a];
    }
    for (int i = 0; i < PX; ++i) d[PX * N + i] = sv[i];
}

void qp_cpu(int model, const double* rec, const Lay& L, int N, double* d) {
    if (model == 0) qp_riccati_cpu<13, 4>(rec, L, N, d);
    else if (model == 1) qp_riccati_cpu<6, 2>(rec, L, N, d);
    else qp_schur_cpu(rec, L, N, 1e-9, d);
}

// Zsoft(h) = sum_i b(-h_i), RelaxedPolyBarrierFunction{0, stiffness, epsilon} (soft_inequality_constraint.hpp:133-145, :171-179;
// soft_sqp.hpp:116-125), in closed form.
double barrier_sum(const double* h, int n, double stiffness, double eps) {
    const double a1 = stiffness, b1 = -0.5 * a1 * eps, c1 = -1.0 / 3.0 * (-b1 - a1 * eps) * eps - 0.5 * a1 * eps * eps - b1 * eps,
                 a2 = (-b1 - a1 * eps) / (eps * eps), b2 = a1, c2 = b1, d2 = c1;
    double z = 0.0;
    for (int i = 0; i < n; ++i) {
        const double x = -h[i];
        if (x < 0.0) z += 0.5 * a1 * x * x + b1 * x + c1;
        else if (x < eps) z += 1.0 / 3.0 * a2 * x * x * x + 0.5 * b2 * x * x + c2 * x + d2;
    }
    return z;
}

struct Merit {
    int model, N, n_dec, m_eq, m_ineq;
    double stiffness, epsilon, mult;
    std::vector<double> g, h;
    Merit(int model_, int N_, double k, double eps, double m) : model(model_), N(N_), stiffness(k), epsilon(eps), mult(m) {
        int s[7];
        oracle_sizes(model, N, s);
        n_dec = s[3]; m_eq = s[5]; m_ineq = s[6];
        g.resize(m_eq); h.resize(m_ineq);
    }
    double objective(const double* xp) {
        double f;
        oracle_eval(model, 0, N, xp, &f);
        return f;
    }
    double phi(const double* xp) {  // soft_sqp.hpp:85-89
        oracle_eval(model, 2, N, xp, h.data());
        return objective(xp) + barrier_sum(h.data(), m_ineq, stiffness, epsilon);
    }
    double theta(const double* xp) {  // soft_sqp.hpp:90-98
        oracle_eval(model, 1, N, xp, g.data());
        double s = 0.0;
        for (double v : g) s += v * v;
        return mult * std::sqrt(s);
    }
};

// One trajectory of SoftSQPOptimizer::Optimize.  status: 0 max iterations, 1 converged, 2 line search failed.
void sqp_one(int model, int N, double* xp, int64_t n_xp, double stiffness, double epsilon, double mult, int iterations, const Lay& L,
             int32_t* status) {
    Merit m(model, N, stiffness, epsilon, mult);
    std::vector<double> rec(L.size), d(m.n_dec), trial(xp, xp + n_xp);
    status[0] = 0;
    status[1] = 0;
    for (int it = 0; it < iterations; ++it) {
        ++status[1];
        const double objective = m.objective(xp);
        oracle_stage_sweep(model, N, xp, 1, n_xp, stiffness, epsilon, rec.data(), L.size, 1);
        qp_cpu(model, rec.data(), L, N, d.data());
        double f0, proj;
        oracle_directional(model, 0, N, xp, d.data(), &f0, &proj);
        // BacktrackingLineSearch::Do, defaults of backtracking_line_search.hpp:70-76
        const double theta = m.theta(xp), phi = m.phi(xp);
        double alpha = 1.0;
        bool accepted = false;
        while (!accepted && alpha >= 1e-4) {
            for (int i = 0; i < m.n_dec; ++i) trial[i] = xp[i] + alpha * d[i];
            const double thn = m.theta(trial.data()), phn = m.phi(trial.data());
            if (thn > 1e-2) accepted = thn < (1.0 - 1e-6) * theta;
            else if (std::max(theta, thn) < 1e-6 && proj < 0.0) accepted = phn < phi + 1e-4 * alpha * proj;
            else accepted = phn < (1.0 - 1e-6) * phi || thn < (1.0 - 1e-6) * theta;
            if (!accepted) alpha *= 0.5;
        }
        if (!accepted) {
            status[0] = 2;
            return;
        }
        std::copy(trial.begin(), trial.begin() + m.n_dec, xp);
        const double diff = m.objective(xp) - objective;
        if (diff < 0.0 && std::fabs(diff) < 1e-6) {
            status[0] = 1;
            return;
        }
    }
}

template <class F>
void parallel_for(int64_t n, int threads, F&& body) {
    if (threads <= 1) {
        for (int64_t i = 0; i < n; ++i) body(i);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            for (int64_t i = n * t / threads; i < n * (t + 1) / threads; ++i) body(i);
        });
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int oracle_qp_solve_model(int model, int N, const double* records, int64_t batch, int64_t ld_rec, double* steps,
                                     int64_t ld_steps, int threads) {
    if (model < 0 || model > 2 || N < 1 || batch < 0 || threads < 1) return -1;
    const Lay L = layout(model, N);
    if (ld_rec < L.size) return -2;
    parallel_for(batch, threads, [&](int64_t b) { qp_cpu(model, records + b * ld_rec, L, N, steps + b * ld_steps); });
    return 0;
}

extern "C" int oracle_qp_solve(int N, const double* records, int64_t batch, int64_t ld_rec, double* steps, int64_t ld_steps, int threads) {
    return oracle_qp_solve_model(QUADRUPED, N, records, batch, ld_rec, steps, ld_steps, threads);
}

extern "C" int oracle_sqp_solve_model(int model, int N, double* xp, int64_t batch, int64_t ld_xp, double stiffness, double epsilon,
                                      double multiplier, int iterations, int32_t* status, int threads) {
    if (model < 0 || model > 2 || N < 1 || batch < 0 || threads < 1 || iterations < 0) return -1;
    int sizes[7];
    oracle_sizes(model, N, sizes);
    if (ld_xp < sizes[3] + sizes[4]) return -2;  // rows shorter than [X | U | parameters]
    const Lay L = layout(model, N);
    parallel_for(batch, threads,
                 [&](int64_t b) { sqp_one(model, N, xp + b * ld_xp, ld_xp, stiffness, epsilon, multiplier, iterations, L, status + 2 * b); });
    return 0;
}

extern "C" int oracle_sqp_solve(int N, double* xp, int64_t batch, int64_t ld_xp, double stiffness, double epsilon, double multiplier,
                                int iterations, int32_t* status, int threads) {
    return oracle_sqp_solve_model(QUADRUPED, N, xp, batch, ld_xp, stiffness, epsilon, multiplier, iterations, status, threads);
}
