// ORACLE — TEST INFRASTRUCTURE ONLY (see ad.hpp header).
//
// Stage-wise CPU port of the KKT block record: the *timed CPU baseline* (bench.py cpu_baseline,
// kind "port") and a second, structurally different evaluation that tests compare with the
// monolithic assembly of oracle.cpp::kkt_record.
//
// What it restates: one pass of SoftSQPOptimizer::AssembleOSQPInstance
// (include/ungar/optimization/soft_sqp.hpp:141-158, :245-264) — objective Hessian/gradient,
// equality values/Jacobian, inequality values/Jacobian, barrier derivatives — but evaluated node by
// node with dense forward duals (the tangent loops vectorise), which is at least as fast as the
// reference's single-threaded CppADCodeGen straight-line code and, unlike it, can use every core.
// The model functions are the templates of models.hpp (reference line cites there).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "models.hpp"

namespace oracle {

namespace {

struct BarrierD {  // closed-form b', b'' of RelaxedPolyBarrierFunction (soft_inequality_constraint.hpp:171-179)
    PolyBarrier pb;
    BarrierD(double k, double e) : pb(k, e) {}
    double b0(double x) const {
        if (x < 0.0) return 0.5 * pb.a1 * x * x + pb.b1 * x + pb.c1;
        if (x < pb.eps) return 1.0 / 3.0 * pb.a2 * x * x * x + 0.5 * pb.b2 * x * x + pb.c2 * x + pb.d2;
        return 0.0;
    }
    double b1(double x) const {
        if (x < 0.0) return pb.a1 * x + pb.b1;
        if (x < pb.eps) return pb.a2 * x * x + pb.b2 * x + pb.c2;
        return 0.0;
    }
    double b2(double x) const {
        if (x < 0.0) return pb.a1;
        if (x < pb.eps) return 2.0 * pb.a2 * x + pb.b2;
        return 0.0;
    }
};

inline int tri(int n, int i, int j) { return i * n - i * (i - 1) / 2 + (j - i); }

struct Offsets {
    int g, A, C, h, cost, grad, H, HN, Hc, size, nz, tri, ntri_N;
};
inline int round4(int x) { return (x + 3) & ~3; }
Offsets offsets(const Sizes& s, int legs, int hc) {
    Offsets r{};
    r.nz = s.nx + s.nu;
    r.tri = r.nz * (r.nz + 1) / 2;
    r.ntri_N = s.nx * (s.nx + 1) / 2;
    int off = 0;
    r.g = off;    off = round4(off + s.m_eq);
    r.A = off;    off = round4(off + s.N * s.nx * r.nz);
    r.C = off;    off = round4(off + s.N * legs * 80);
    r.h = off;    off = round4(off + s.m_ineq);
    r.cost = off; off = round4(off + 2);
    r.grad = off; off = round4(off + s.n_dec);
    r.H = off;    off = round4(off + s.N * r.tri);
    r.HN = off;   off = round4(off + r.ntri_N);
    r.Hc = off;   off = round4(off + (s.N - 1) * hc);
    r.size = off;
    return r;
}

// Accumulates one inequality row (value hv, gradient row jh over the node's nz locals) into the
// barrier value, the QP gradient and the Gauss-Newton block:  q += dZ/dz * jh,  H += d2Z/dz2 * jh jh^T.
template <int NZ>
inline void add_barrier_row(const BarrierD& bar, double hv, const std::array<double, NZ>& jh, double& B,
                            double* grad_z, double* Hk) {
    B += bar.b0(-hv);
    const double dz  = -bar.b1(-hv);
    const double d2z = bar.b2(-hv);
    int nzidx[NZ];
    int cnt = 0;
    for (int c = 0; c < NZ; ++c)
        if (jh[c] != 0.0) nzidx[cnt++] = c;
    for (int a = 0; a < cnt; ++a) {
        const int ia = nzidx[a];
        grad_z[ia] += dz * jh[ia];
        for (int b = a; b < cnt; ++b) Hk[tri(NZ, ia, nzidx[b])] += d2z * jh[ia] * jh[nzidx[b]];
    }
}

// Squared affine residual c * (w * (z_i - ref))^2 on local index i.
inline void add_square(double c, double w, double diff, int i, int nz, double& cost, double* grad_z, double* Hk) {
    const double r = w * diff;
    cost += c * r * r;
    grad_z[i] += 2.0 * c * w * r;
    Hk[tri(nz, i, i)] += 2.0 * c * w * w;
}

// Min(|q - r|^2, |q + r|^2): CondExpGt(dm, dp, dp, dm) selects '+' only when dm > dp.
inline void add_quat_term(const double* q, const double* r, int i0, int nz, double& cost, double* grad_z, double* Hk) {
    double dm = 0.0, dp = 0.0;
    for (int i = 0; i < 4; ++i) {
        dm += (q[i] - r[i]) * (q[i] - r[i]);
        dp += (q[i] + r[i]) * (q[i] + r[i]);
    }
    const double sgn = dm > dp ? 1.0 : -1.0;
    for (int i = 0; i < 4; ++i) add_square(1.0, 1.0, q[i] + sgn * r[i], i0 + i, nz, cost, grad_z, Hk);
}

// ------------------------------------------------------------------------------------------
void quadrotor_trajectory(int N, const double* xp, const BarrierD& bar, double* rec) {
    const QuadrotorLayout L{N};
    const Sizes s = L.sizes();
    const Offsets R = offsets(s, 0, s.nu);
    constexpr int NZ = 17;
    using D = DDual<NZ>;
    std::fill(rec, rec + R.size, 0.0);
    double cost = 0.0, B = 0.0;
    std::vector<D> v(xp, xp + s.n_dec + s.n_par);  // constants; the node's own z_k is re-seeded below
    for (int i = 0; i < 13; ++i) rec[R.g + i] = xp[L.X(0) + i] - xp[L.xm() + i];
    for (int k = 0; k <= N; ++k) {
        const bool terminal = k == N;
        double* Hk = terminal ? rec + R.HN : rec + R.H + k * R.tri;
        const int nz = terminal ? 13 : NZ;
        double gz[NZ] = {0.0};
        const double* x = xp + L.X(k);
        for (int i = 0; i < 3; ++i) add_square(1.0, 1.0, x[i] - xp[L.pref(k) + i], i, nz, cost, gz, Hk);
        add_quat_term(x + 3, xp + L.qref(k), 3, nz, cost, gz, Hk);
        for (int i = 0; i < 3; ++i) add_square(1.0, 1.0, x[7 + i] - xp[L.vref(k) + i], 7 + i, nz, cost, gz, Hk);
        for (int i = 0; i < 3; ++i) add_square(1.0, 1.0, x[10 + i] - xp[L.wref(k) + i], 10 + i, nz, cost, gz, Hk);
        if (!terminal) {
            const double* u = xp + L.U(k);
            for (int i = 0; i < 4; ++i) {
                add_square(1e-6, 1.0, u[i], 13 + i, nz, cost, gz, Hk);
                if (k) add_square(1e-6, 1.0, u[i] - xp[L.U(k - 1) + i], 13 + i, nz, cost, gz, Hk);
                if (k + 1 < N) {  // the (k+1) term of quadrotor.example.cpp:219-225 seen from u_k
                    const double e = xp[L.U(k + 1) + i] - u[i];
                    gz[13 + i] += -2e-6 * e;
                    Hk[tri(nz, 13 + i, 13 + i)] += 2e-6;
                    rec[R.Hc + k * 4 + i] = -2e-6;
                }
            }
            // dynamics defect and its Jacobian
            for (int i = 0; i < NZ; ++i) {
                D& z = v[i < 13 ? L.X(k) + i : L.U(k) + i - 13];
                z.d.fill(0.0);
                z.d[i] = 1.0;
            }
            D xn[13];
            quadrotor_dynamics<D>(L, v.data(), v.data() + L.X(k), v.data() + L.U(k), xn);
            for (int r = 0; r < 13; ++r) {
                rec[R.g + 13 + 13 * k + r] = xp[L.X(k + 1) + r] - xn[r].v;
                for (int c = 0; c < NZ; ++c) rec[R.A + (k * 13 + r) * NZ + c] = -xn[r].d[c];
            }
            for (int i = 0; i < NZ; ++i) v[i < 13 ? L.X(k) + i : L.U(k) + i - 13].d.fill(0.0);
            // inequality rows r - rmax, -r  (quadrotor.example.cpp:280-288)
            for (int i = 0; i < 4; ++i) {
                std::array<double, NZ> jh{};
                jh[13 + i] = 1.0;
                const double h0 = u[i] - xp[L.rmax()];
                rec[R.h + (k * 4 + i) * 2] = h0;
                add_barrier_row<NZ>(bar, h0, jh, B, gz, Hk);
                jh[13 + i] = -1.0;
                rec[R.h + (k * 4 + i) * 2 + 1] = -u[i];
                add_barrier_row<NZ>(bar, -u[i], jh, B, gz, Hk);
            }
        }
        for (int i = 0; i < nz; ++i) Hk[tri(nz, i, i)] += 1e-6;  // soft_sqp.hpp:148-150
        for (int i = 0; i < 13; ++i) rec[R.grad + L.X(k) + i] = gz[i];
        if (!terminal)
            for (int i = 0; i < 4; ++i) rec[R.grad + L.U(k) + i] = gz[13 + i];
    }
    rec[R.cost] = cost;
    rec[R.cost + 1] = B;
}

// ------------------------------------------------------------------------------------------
void rc_car_trajectory(int N, const double* xp, const BarrierD& bar, double* rec) {
    const RcCarLayout L{N};
    const Sizes s = L.sizes();
    const Offsets R = offsets(s, 0, s.nu);
    constexpr int NZ = 8;
    using D = DDual<NZ>;
    std::fill(rec, rec + R.size, 0.0);
    double cost = 0.0, B = 0.0;
    std::vector<D> v(xp, xp + s.n_dec + s.n_par);
    for (int i = 0; i < 6; ++i) rec[R.g + i] = xp[L.X(0) + i] - xp[L.xm() + i];
    for (int k = 0; k <= N; ++k) {
        const bool terminal = k == N;
        double* Hk = terminal ? rec + R.HN : rec + R.H + k * R.tri;
        const int nz = terminal ? 6 : NZ;
        double gz[NZ] = {0.0};
        const double* x = xp + L.X(k);
        for (int i = 0; i < 2; ++i) add_square(1.0, 1.0, x[i] - xp[L.pref(k) + i], i, nz, cost, gz, Hk);
        if (!terminal) {
            const double* u = xp + L.U(k);
            for (int i = 0; i < 2; ++i) {
                add_square(1e-6, 1.0, u[i], 6 + i, nz, cost, gz, Hk);
                if (k) add_square(1e-6, 1.0, u[i] - xp[L.U(k - 1) + i], 6 + i, nz, cost, gz, Hk);
                if (k + 1 < N) {
                    const double e = xp[L.U(k + 1) + i] - u[i];
                    gz[6 + i] += -2e-6 * e;
                    Hk[tri(nz, 6 + i, 6 + i)] += 2e-6;
                    rec[R.Hc + k * 2 + i] = -2e-6;
                }
            }
            for (int i = 0; i < NZ; ++i) {
                D& z = v[i < 6 ? L.X(k) + i : L.U(k) + i - 6];
                z.d.fill(0.0);
                z.d[i] = 1.0;
            }
            D xn[6];
            rc_car_dynamics<D>(L, v.data(), v.data() + L.X(k), v.data() + L.U(k), xn);
            for (int r = 0; r < 6; ++r) {
                rec[R.g + 6 + 6 * k + r] = xp[L.X(k + 1) + r] - xn[r].v;
                for (int c = 0; c < NZ; ++c) rec[R.A + (k * 6 + r) * NZ + c] = -xn[r].d[c];
            }
            // |d| - 15, |delta| - 15, 0.3 - v_x  (rc_car.example.cpp:271-282)
            const D hrow[3] = {ad_abs(v[L.U(k)]) - 15.0, ad_abs(v[L.U(k) + 1]) - 15.0, 0.3 - v[L.X(k) + 3]};
            for (int i = 0; i < 3; ++i) {
                rec[R.h + 3 * k + i] = hrow[i].v;
                add_barrier_row<NZ>(bar, hrow[i].v, hrow[i].d, B, gz, Hk);
            }
            for (int i = 0; i < NZ; ++i) v[i < 6 ? L.X(k) + i : L.U(k) + i - 6].d.fill(0.0);
        }
        for (int i = 0; i < nz; ++i) Hk[tri(nz, i, i)] += 1e-6;
        for (int i = 0; i < 6; ++i) rec[R.grad + L.X(k) + i] = gz[i];
        if (!terminal)
            for (int i = 0; i < 2; ++i) rec[R.grad + L.U(k) + i] = gz[6 + i];
    }
    rec[R.cost] = cost;
    rec[R.cost + 1] = B;
}

// ------------------------------------------------------------------------------------------
// Foot position p + q * r with the 10 local tangents [p(3) q(4) r(3)] starting at tangent `t0`.
template <int K>
inline V3<DDual<K>> foot_position(const double* pose, const double* r, int t0) {
    using D = DDual<K>;
    D z[10];
    for (int i = 0; i < 10; ++i) {
        z[i] = D(i < 7 ? pose[i] : r[i - 7]);
        z[i].d[t0 + i] = 1.0;
    }
    return V3<D>{z[0], z[1], z[2]} + rotate(Q4<D>{z[3], z[4], z[5], z[6]}, V3<D>{z[7], z[8], z[9]});
}

void quadruped_trajectory(int N, const double* xp, const BarrierD& bar, double* rec) {
    const QuadrupedLayout L{N};
    const Sizes s = L.sizes();
    const Offsets R = offsets(s, 4, 0);
    constexpr int NZ = 37;
    using D = DDual<NZ>;
    std::fill(rec, rec + R.size, 0.0);
    double cost = 0.0, B = 0.0;
    std::vector<D> v(xp, xp + s.n_dec + s.n_par);
    for (int i = 0; i < 13; ++i) rec[R.g + i] = xp[L.X(0) + i] - xp[L.xm() + i];
    const int contact0 = 13 + 13 * N;
    for (int k = 0; k <= N; ++k) {
        const bool terminal = k == N;
        double* Hk = terminal ? rec + R.HN : rec + R.H + k * R.tri;
        const int nz = terminal ? 13 : NZ;
        double gz[NZ] = {0.0};
        const double* x  = xp + L.X(k);
        const double* pr = xp + L.P(k);
        const double wpos[3] = {0.1, 0.1, 10.0};
        for (int i = 0; i < 3; ++i) add_square(1.0, wpos[i], x[i] - pr[i], i, nz, cost, gz, Hk);
        add_quat_term(x + 3, pr + 3, 3, nz, cost, gz, Hk);
        for (int i = 0; i < 6; ++i) add_square(1.0, 1.0, x[7 + i] - pr[7 + i], 7 + i, nz, cost, gz, Hk);
        if (!terminal) {
            for (int leg = 0; leg < 4; ++leg)
                for (int i = 0; i < 3; ++i) {
                    add_square(1.0, 1.0, xp[L.R(k, leg) + i] - xp[L.Rref(k, leg) + i], 13 + 6 * leg + 3 + i, nz, cost, gz, Hk);
                    add_square(1e-8, 1.0, xp[L.F(k, leg) + i], 13 + 6 * leg + i, nz, cost, gz, Hk);
                }
            for (int i = 0; i < NZ; ++i) {
                D& z = v[i < 13 ? L.X(k) + i : L.U(k) + i - 13];
                z.d.fill(0.0);
                z.d[i] = 1.0;
            }
            D xn[13];
            quadruped_dynamics<D>(L, v.data(), k, xn);
            for (int r = 0; r < 13; ++r) {
                rec[R.g + 13 + 13 * k + r] = xp[L.X(k + 1) + r] - xn[r].v;
                for (int c = 0; c < NZ; ++c) rec[R.A + (k * 13 + r) * NZ + c] = -xn[r].d[c];
            }
            // inequality rows (quadruped.example.cpp:330-333)
            const D& mu = v[L.mu()];
            for (int leg = 0; leg < 4; ++leg) {
                const D& sc = v[L.S(k, leg)];
                const V3<D> f = load3(v.data() + L.F(k, leg));
                const V3<D> r = load3(v.data() + L.R(k, leg));
                const D hrow[3] = {-sc * f.z, sc * approx_norm2(f.x, f.y) - mu * f.z,
                                   sc * approx_norm3(r - load3(v.data() + L.hip(leg))) - v[L.leg_length()]};
                for (int i = 0; i < 3; ++i) {
                    rec[R.h + (k * 4 + leg) * 3 + i] = hrow[i].v;
                    add_barrier_row<NZ>(bar, hrow[i].v, hrow[i].d, B, gz, Hk);
                }
            }
            for (int i = 0; i < NZ; ++i) v[i < 13 ? L.X(k) + i : L.U(k) + i - 13].d.fill(0.0);
            // contact rows (quadruped.example.cpp:279-303), 20 local tangents per leg
            for (int leg = 0; leg < 4; ++leg) {
                using E = DDual<20>;
                const double sc = xp[L.S(k, leg)];
                const double sp = k ? xp[L.S(k - 1, leg)] : xp[L.s_meas(leg)];
                const V3<E> foot = foot_position<20>(xp + L.X(k), xp + L.R(k, leg), 0);
                V3<E> prev;
                if (k) prev = foot_position<20>(xp + L.X(k - 1), xp + L.R(k - 1, leg), 10);
                else prev = V3<E>{E(xp[L.foot_meas(leg)]), E(xp[L.foot_meas(leg) + 1]), E(xp[L.foot_meas(leg) + 2])};
                const E rows[4] = {(1.0 - sp) * sc * foot.z, (sp * sc) * (foot.x - prev.x),
                                   (sp * sc) * (foot.y - prev.y), (sp * sc) * (foot.z - prev.z)};
                for (int rr = 0; rr < 4; ++rr) {
                    rec[R.g + contact0 + 16 * k + 4 * leg + rr] = rows[rr].v;
                    for (int c = 0; c < 20; ++c) rec[R.C + ((k * 4 + leg) * 4 + rr) * 20 + c] = rows[rr].d[c];
                }
            }
        }
        for (int i = 0; i < nz; ++i) Hk[tri(nz, i, i)] += 1e-6;
        for (int i = 0; i < 13; ++i) rec[R.grad + L.X(k) + i] = gz[i];
        if (!terminal)
            for (int i = 0; i < 24; ++i) rec[R.grad + L.U(k) + i] = gz[13 + i];
    }
    rec[R.cost] = cost;
    rec[R.cost + 1] = B;
}

}  // namespace

}  // namespace oracle

extern "C" int oracle_stage_sweep(int model, int N, const double* xp, int64_t batch, int64_t ld_xp, double stiffness,
                                  double epsilon, double* records, int64_t ld_rec, int threads) {
    using namespace oracle;
    if (model < 0 || model > 2 || N < 1 || threads < 1) return -1;
    const BarrierD bar(stiffness, epsilon);
    auto work = [&](int64_t b0, int64_t b1) {
        for (int64_t b = b0; b < b1; ++b) {
            const double* x = xp + b * ld_xp;
            double* r = records + b * ld_rec;
            if (model == 0) quadrotor_trajectory(N, x, bar, r);
            else if (model == 1) rc_car_trajectory(N, x, bar, r);
            else quadruped_trajectory(N, x, bar, r);
        }
    };
    if (threads == 1) {
        work(0, batch);
        return 0;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, batch * t / threads, batch * (t + 1) / threads);
    for (auto& th : pool) th.join();
    return 0;
}
