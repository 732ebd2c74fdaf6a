// Batched backtracking line search of the soft SQP, all trial points evaluated on the device.
//
// Replaces BacktrackingLineSearch::Do (include/ungar/optimization/backtracking_line_search.hpp:81-165) called with the two
// merit lambdas of SoftSQPOptimizer::Optimize (include/ungar/optimization/soft_sqp.hpp:81-99):
//     phi(w)   = f(w) + Zsoft(h(w))                       cost function
//     theta(w) = multiplier * sqrt(|g(w)|^2)              constraint violation
// and the bookkeeping around it (soft_sqp.hpp:100-108): the accepted step is applied to the trajectory's flat vector in
// place, the iteration counter advances, and the trajectory stops when no step is accepted or when the objective decreased
// by less than 1e-6.
//
// Mapping.  One CTA per trajectory, one thread per shooting node (strided when N + 1 > blockDim).  The trial point
// w + alpha dw is materialised as a whole flat Ungar vector [X | U | parameters] in shared memory, so the model functors of
// models.cuh — which read neighbouring nodes (x_{k+1} in the defect, u_{k-1} in the input-rate cost, the pose of node k - 1
// in the contact rows) straight from the flat vector — run on it unchanged with plain scalars.  Every trial is one pass:
// per-node partial sums of f, Zsoft and |g|^2, then a fixed-order block reduction (deterministic).  The directional
// derivative grad f . dw (the `dwProjection` of the reference) is one more pass with Dual<T> seeded by dw.
// The whole backtracking loop runs inside the kernel: no host round trip per trial.
#pragma once

#include "sweep.cuh"

namespace ub {

struct LineSearchParams {  // BacktrackingLineSearch::Parameters (backtracking_line_search.hpp:57-76) + soft_sqp.hpp:45
    double alpha_min, theta_min, theta_max, eta, gamma_phi, gamma_theta, gamma_alpha, multiplier, objective_tolerance;
};

enum SqpStatus : int { SQP_RUNNING = 0, SQP_CONVERGED = 1, SQP_LINE_SEARCH_FAILED = 2 };

constexpr int LS_THREADS = 128;
constexpr int LS_INFO    = 8;  // alpha, theta, phi, f (after) | theta0, phi0, f0, grad f . dw

// Fixed-order block reduction of three sums; every thread gets the totals.
template <class T>
__device__ __forceinline__ void ls_reduce3(T& a, T& b, T& c, T* red) {
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();  // red may still be read from the previous reduction
    if ((threadIdx.x & 31) == 0) { red[3 * w] = a; red[3 * w + 1] = b; red[3 * w + 2] = c; }
    __syncthreads();
    a = b = c = T(0);
    for (int i = 0; i < nw; ++i) { a += red[3 * i]; b += red[3 * i + 1]; c += red[3 * i + 2]; }
}

// Per-node contributions to (f, Zsoft, |g|^2) at the flat vector x (shared memory).
template <class Mdl, class T>
__device__ __forceinline__ void ls_node_values(const T* x, int N, int k, const BarrierCoef<T>& bar, T& f, T& zs, T& g2) {
    constexpr int NX = Mdl::NX, NU = Mdl::NU, NZ = Mdl::NZ, NH = Mdl::NH, LEGS = Mdl::LEGS;
    T z[NZ];
#pragma unroll
    for (int i = 0; i < NX; ++i) z[i] = x[Mdl::x_off(N, k) + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) z[NX + i] = k < N ? x[Mdl::u_off(N, k) + i] : T(0);
    Mdl::cost_terms(x, N, k, z, [&](T c, const T& res, bool counts) {
        if (counts) f += c * res * res;
    });
    if (k == 0) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const T r = x[i] - x[Mdl::xm_off(N) + i];
            g2 += r * r;
        }
    }
    if (k == N) return;
    {
        T xn[NX];
        Mdl::dynamics(x, N, k, z, xn);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const T r = x[Mdl::x_off(N, k + 1) + i] - xn[i];
            g2 += r * r;
        }
    }
    {
        T h[NH];
        Mdl::inequalities(x, N, k, z, h);
#pragma unroll
        for (int i = 0; i < NH; ++i) {
            T b0, dz, d2z;
            barrier_eval(bar, h[i], &b0, &dz, &d2z);
            zs += b0;
        }
    }
    if constexpr (LEGS > 0) {
        for (int leg = 0; leg < LEGS; ++leg) {
            T zl[20], rows[4];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                zl[i]      = z[i];
                zl[10 + i] = k ? x[Mdl::x_off(N, k - 1) + i] : T(0);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                zl[7 + i]  = z[NX + 6 * leg + 3 + i];
                zl[17 + i] = k ? x[Mdl::u_off(N, k - 1) + 6 * leg + 3 + i] : T(0);
            }
            Mdl::contact_rows(x, N, k, leg, zl, rows);
#pragma unroll
            for (int r = 0; r < 4; ++r) g2 += rows[r] * rows[r];
        }
    }
}

// Residency: the trial vector takes (n_dec + n_par) * 8 bytes of shared memory per CTA (quadruped N = 100: 53.5 KB -> 4 CTAs per SM);
// left alone ptxas takes 205 registers for the quadruped functors, which caps the SM at 2 CTAs = 8 warps of a latency-bound kernel.
template <class Mdl>
struct LsResidency { static constexpr int MIN_CTAS = Mdl::LEGS > 0 ? 4 : 1; };

template <class Mdl, class T>
__global__ void __launch_bounds__(LS_THREADS, LsResidency<Mdl>::MIN_CTAS)
line_search_kernel(T* __restrict__ xp_all, long long ld_xp, const T* __restrict__ dw_all, long long ld_dw, int N,
                   BarrierCoef<T> bar, LineSearchParams P, int* __restrict__ status_all, T* __restrict__ info_all) {
    constexpr int NX = Mdl::NX, NU = Mdl::NU, NZ = Mdl::NZ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* const sx = reinterpret_cast<T*>(smem_raw);
    __shared__ T red[3 * (LS_THREADS / 32)];
    const long long b = blockIdx.x;
    int* st = status_all ? status_all + 2 * b : nullptr;
    if (st && st[0] != SQP_RUNNING) return;  // uniform per CTA: the trajectory stopped in an earlier iteration
    T* __restrict__ xg        = xp_all + b * ld_xp;
    const T* __restrict__ dw  = dw_all + b * ld_dw;
    const int n_dec = Mdl::n_dec(N), n_in = n_dec + Mdl::n_par(N), t = threadIdx.x;

    for (int e = t; e < n_in; e += blockDim.x) sx[e] = xg[e];
    __syncthreads();

    // ---- grad f . dw  (costFunctionGradient . dw, backtracking_line_search.hpp:92) : one Dual pass seeded with dw -------------
    T proj = T(0), f0 = T(0), z0 = T(0), g0 = T(0);
    for (int k = t; k <= N; k += blockDim.x) {
        Dual<T> z[NZ];
#pragma unroll
        for (int i = 0; i < NX; ++i) z[i] = Dual<T>(sx[Mdl::x_off(N, k) + i], dw[Mdl::x_off(N, k) + i]);
#pragma unroll
        for (int i = 0; i < NU; ++i) z[NX + i] = k < N ? Dual<T>(sx[Mdl::u_off(N, k) + i], dw[Mdl::u_off(N, k) + i]) : Dual<T>();
        Mdl::cost_terms(sx, N, k, z, [&](T c, const Dual<T>& res, bool) { proj += T(2) * c * res.v * res.d; });
        ls_node_values<Mdl, T>(sx, N, k, bar, f0, z0, g0);
    }
    {
        T dummy = T(0), dummy2 = T(0);
        ls_reduce3(proj, dummy, dummy2, red);
    }
    ls_reduce3(f0, z0, g0, red);
    const T mult = T(P.multiplier);
    const T theta = mult * m_sqrt(g0), phi = f0 + z0;

    // ---- backtracking loop (backtracking_line_search.hpp:118-148) -------------------------------------------------------------------
    T alpha = T(1), fn = f0, thn = theta, phn = phi;
    bool accepted = false;
    while (!accepted && alpha >= T(P.alpha_min)) {
        __syncthreads();
        for (int e = t; e < n_dec; e += blockDim.x) sx[e] = xg[e] + alpha * dw[e];
        __syncthreads();
        T f = T(0), zs = T(0), g2 = T(0);
        for (int k = t; k <= N; k += blockDim.x) ls_node_values<Mdl, T>(sx, N, k, bar, f, zs, g2);
        ls_reduce3(f, zs, g2, red);
        fn = f; thn = mult * m_sqrt(g2); phn = f + zs;
        if (thn > T(P.theta_max)) {
            accepted = thn < (T(1) - T(P.gamma_theta)) * theta;
        } else if (fmax(theta, thn) < T(P.theta_min) && proj < T(0)) {
            accepted = phn < phi + T(P.eta) * alpha * proj;
        } else {
            accepted = phn < (T(1) - T(P.gamma_phi)) * phi || thn < (T(1) - T(P.gamma_theta)) * theta;
        }
        if (!(thn == thn) || !(phn == phn)) accepted = false;  // NaN trial: comparisons above are already false; keep it explicit
        if (!accepted) alpha *= T(P.gamma_alpha);
    }
    // ---- apply the step (w += alpha dw, :150-151), advance the SQP bookkeeping (soft_sqp.hpp:100-108) -------------------------------
    if (accepted)
        for (int e = t; e < n_dec; e += blockDim.x) xg[e] = sx[e];
    if (t == 0) {
        if (st) {
            st[1] += 1;
            const T diff = fn - f0;
            if (!accepted) st[0] = SQP_LINE_SEARCH_FAILED;
            else if (diff < T(0) && m_abs(diff) < T(P.objective_tolerance)) st[0] = SQP_CONVERGED;
        }
        if (info_all) {
            T* info = info_all + b * LS_INFO;
            info[0] = accepted ? alpha : T(0);
            info[1] = accepted ? thn : theta;
            info[2] = accepted ? phn : phi;
            info[3] = accepted ? fn : f0;
            info[4] = theta; info[5] = phi; info[6] = f0; info[7] = proj;
        }
    }
}

}  // namespace ub
