// KKT stage sweep, generic version ("v1"): one thread per (shooting node, tangent direction).
//
// A CTA owns a tile of up to M consecutive shooting nodes of one trajectory.  Thread (m, j) seeds
// d z_j = 1 for node k0 + m and, in one forward pass with Dual<T>, obtains column j of
//   - the dynamics Jacobian            -> A_k[:, j]
//   - the inequality Jacobian J_h,k    -> barrier gradient entry j and (after a CTA barrier) column j of
//                                         the Gauss-Newton block J_h^T D J_h
//   - the separable objective terms    -> gradient entry j, diagonal Hessian entry (j, j), coupling entry
// Everything a tile produces is assembled in shared memory in exactly the order it has in the
// trajectory's record, then written to HBM with fully coalesced stores (the tile's slices of A, H, C, g, h,
// grad are contiguous in the record).  Inputs are read straight from the flat Ungar vector
// [X | U | parameters]; they are ~4 % of the traffic and are served from L1/L2.
//
// Replaces: the generated `sparse_jacobian` / `sparse_hessian` / `forward_zero` C code behind
// GenericModel (include/ungar/autodiff/function.hpp:186-189, :224-228, :252-257) for the three example
// problems, and the sparse products of SoftSQPOptimizer::AssembleOSQPInstance
// (include/ungar/optimization/soft_sqp.hpp:141-158, :245-264).
#pragma once

#include "models.cuh"

namespace ub {

struct RecLayout {
    int g, A, C, h, cost, grad, H, HN, Hc, size;
};

// RelaxedPolyBarrierFunction coefficients (optimization/soft_inequality_constraint.hpp:133-145).
template <class T>
struct BarrierCoef {
    T eps, a1, b1, c1, a2, b2, c2, d2;
};

// Zsoft(z) = sum_i b(-z_i) (soft_sqp.hpp:116-125): value b(-h), dZ/dh = -b'(-h), d2Z/dh2 = b''(-h).
template <class T>
UB_HD void barrier_eval(const BarrierCoef<T>& B, T h, T* b0, T* dz, T* d2z) {
    const T x = -h;
    if (x < T(0)) {
        *b0  = T(0.5) * B.a1 * x * x + B.b1 * x + B.c1;
        *dz  = -(B.a1 * x + B.b1);
        *d2z = B.a1;
    } else if (x < B.eps) {
        *b0  = T(1.0 / 3.0) * B.a2 * x * x * x + T(0.5) * B.b2 * x * x + B.c2 * x + B.d2;
        *dz  = -(B.a2 * x * x + B.b2 * x + B.c2);
        *d2z = T(2) * B.a2 * x + B.b2;
    } else {
        *b0 = *dz = *d2z = T(0);
    }
}

UB_HD int tri_index(int n, int i, int j) { return i * n - (i * (i - 1)) / 2 + (j - i); }  // i <= j

template <class Mdl, int M>
struct SweepShape {
    static constexpr int NX = Mdl::NX, NU = Mdl::NU, NZ = Mdl::NZ, NH = Mdl::NH, LEGS = Mdl::LEGS;
    static constexpr int TRI     = NZ * (NZ + 1) / 2;
    static constexpr int THREADS = ((M * NZ + 31) / 32) * 32;
    // shared-memory staging (elements)
    static constexpr int oA = 0, oH = oA + M * NX * NZ, oJ = oH + M * TRI, oC = oJ + M * NH * NZ,
                         oG = oC + M * LEGS * 80, oGc = oG + M * NX, oHv = oGc + M * LEGS * 4, oQ = oHv + M * NH,
                         oHc = oQ + M * NZ, total = oHc + M * NU;
};

template <class T>
__device__ __forceinline__ void copy_out(T* __restrict__ dst, const T* __restrict__ src, int count) {
    for (int e = threadIdx.x; e < count; e += blockDim.x) dst[e] = src[e];
}

template <class Mdl, class T, int M, bool BARRIER>
__global__ void __launch_bounds__(SweepShape<Mdl, M>::THREADS)
kkt_sweep_kernel(const T* __restrict__ xp_all, long long ld_xp, T* __restrict__ rec_all, long long ld_rec,
                 T* __restrict__ stage_cost, int N, int tiles, RecLayout L, BarrierCoef<T> bar) {
    using Sh = SweepShape<Mdl, M>;
    constexpr int NX = Sh::NX, NU = Sh::NU, NZ = Sh::NZ, NH = Sh::NH, LEGS = Sh::LEGS, TRI = Sh::TRI;
    using D = Dual<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* const sm = reinterpret_cast<T*>(smem_raw);
    T *sA = sm + Sh::oA, *sH = sm + Sh::oH, *sJ = sm + Sh::oJ, *sC = sm + Sh::oC, *sG = sm + Sh::oG,
      *sGc = sm + Sh::oGc, *sHv = sm + Sh::oHv, *sQ = sm + Sh::oQ, *sHc = sm + Sh::oHc;

    const int b    = blockIdx.x / tiles;
    const int tile = blockIdx.x - b * tiles;
    const int k0   = tile * M;
    const int mt   = min(M, N - k0);
    const T* __restrict__ x = xp_all + b * ld_xp;
    T* __restrict__ r       = rec_all + b * ld_rec;
    const int t = threadIdx.x, m = t / NZ, j = t - m * NZ;
    const bool active = m < mt;
    const int k = k0 + m;

    T w[NH > 0 ? NH : 1];  // d2Z_i * dh_i/dz_j, kept for the Gauss-Newton column
    T hjj = T(0), cval = T(0), bsum = T(0);
    if (active) {
        D z[NZ];
#pragma unroll
        for (int i = 0; i < NX; ++i) z[i] = D(x[Mdl::x_off(N, k) + i], i == j ? T(1) : T(0));
#pragma unroll
        for (int i = 0; i < NU; ++i) z[NX + i] = D(x[Mdl::u_off(N, k) + i], NX + i == j ? T(1) : T(0));

        // --- dynamics defect x_{k+1} - f(x_k, u_k) and column j of A_k -----------------------------
        {
            D xn[NX];
            Mdl::dynamics(x, N, k, z, xn);
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                sA[(m * NX + i) * NZ + j] = -xn[i].d;
                if (i == j) sG[m * NX + i] = x[Mdl::x_off(N, k + 1) + i] - xn[i].v;
            }
        }
        // --- inequalities, barrier gradient, staging of J_h ------------------------------------------
        T gq = T(0);
        {
            D h[NH];
            Mdl::inequalities(x, N, k, z, h);
#pragma unroll
            for (int i = 0; i < NH; ++i) {
                T b0 = T(0), dz = T(0), d2z = T(0);
                if (BARRIER) barrier_eval(bar, h[i].v, &b0, &dz, &d2z);
                bsum += b0;
                gq += dz * h[i].d;
                w[i] = d2z * h[i].d;
                sJ[(m * NH + i) * NZ + j] = h[i].d;
                if (j == 0) sHv[m * NH + i] = h[i].v;
            }
        }
        // --- separable objective terms -------------------------------------------------------------------
        T hc = T(0);
        Mdl::cost_terms(x, N, k, z, [&](T c, const D& res, bool counts) {
            if (counts) cval += c * res.v * res.v;
            else hc -= T(2) * c * res.d * res.d;  // d2/du_k du_{k+1} of c (u_{k+1} - u_k)^2
            gq += T(2) * c * res.v * res.d;
            hjj += T(2) * c * res.d * res.d;
        });
        sQ[m * NZ + j] = gq;
        if (Mdl::HC && j >= NX) sHc[m * NU + j - NX] = hc;
        // --- contact rows (quadruped), 20 local tangents per leg -----------------------------------------
        if constexpr (LEGS > 0) {
            if (j < 20) {
                for (int leg = 0; leg < LEGS; ++leg) {
                    D zl[20];
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        zl[i]      = D(x[Mdl::x_off(N, k) + i], i == j ? T(1) : T(0));
                        zl[10 + i] = D(k ? x[Mdl::x_off(N, k - 1) + i] : T(0), 10 + i == j ? T(1) : T(0));
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        zl[7 + i]  = D(x[Mdl::u_off(N, k) + 6 * leg + 3 + i], 7 + i == j ? T(1) : T(0));
                        zl[17 + i] = D(k ? x[Mdl::u_off(N, k - 1) + 6 * leg + 3 + i] : T(0), 17 + i == j ? T(1) : T(0));
                    }
                    D rows[4];
                    Mdl::contact_rows(x, N, k, leg, zl, rows);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        sC[((m * LEGS + leg) * 4 + rr) * 20 + j] = rows[rr].d;
                        if (j == 0) sGc[(m * LEGS + leg) * 4 + rr] = rows[rr].v;
                    }
                }
            }
        }
    }
    __syncthreads();
    // --- per-node partials: objective, barrier, |g|_inf, max h (reduced over nodes by finalize_kernel) -----------
    if (active && j == 0) {
        T gmax = T(0), hmax = -INFINITY;
        for (int i = 0; i < NX; ++i) gmax = fmax(gmax, m_abs(sG[m * NX + i]));
        for (int i = 0; i < LEGS * 4; ++i) gmax = fmax(gmax, m_abs(sGc[m * LEGS * 4 + i]));
        for (int i = 0; i < NH; ++i) hmax = fmax(hmax, sHv[m * NH + i]);
        T* pt = stage_cost + ((long long)b * (N + 1) + k) * 4;
        pt[0] = cval; pt[1] = bsum; pt[2] = gmax; pt[3] = hmax;
    }
    // --- column j of the block  H_k = diag(objective) + J_h^T D J_h + 1e-6 I  (upper triangle) ---------------
    if (active) {
        for (int a = 0; a <= j; ++a) {
            T acc = T(0);
#pragma unroll
            for (int i = 0; i < NH; ++i) acc += w[i] * sJ[(m * NH + i) * NZ + a];
            if (a == j) acc += hjj + (BARRIER ? T(1e-6) : T(0));
            sH[m * TRI + tri_index(NZ, a, j)] = acc;
        }
    }
    // --- x_0 - x_measured (first tile) and the terminal node x_N (last tile) -----------------------------
    if (tile == 0 && t < NX) r[L.g + t] = x[t] - x[Mdl::xm_off(N) + t];
    if (tile == tiles - 1 && t < NX) {
        D zN[NZ];
#pragma unroll
        for (int i = 0; i < NX; ++i) zN[i] = D(x[Mdl::x_off(N, N) + i], i == t ? T(1) : T(0));
        T gq = T(0), hnn = T(0), cvalN = T(0);
        Mdl::cost_terms(x, N, N, zN, [&](T c, const D& res, bool) {
            cvalN += c * res.v * res.v;
            gq += T(2) * c * res.v * res.d;
            hnn += T(2) * c * res.d * res.d;
        });
        r[L.grad + Mdl::x_off(N, N) + t] = gq;
        for (int a = 0; a <= t; ++a)
            r[L.HN + tri_index(NX, a, t)] = a == t ? hnn + (BARRIER ? T(1e-6) : T(0)) : T(0);
        if (t == 0) {
            T* pt = stage_cost + ((long long)b * (N + 1) + N) * 4;
            pt[0] = cvalN; pt[1] = T(0); pt[2] = T(0); pt[3] = -INFINITY;
        }
    }
    __syncthreads();
    // --- coalesced write-out of the tile ---------------------------------------------------------------------
    copy_out(r + L.A + k0 * NX * NZ, sA, mt * NX * NZ);
    copy_out(r + L.H + k0 * TRI, sH, mt * TRI);
    copy_out(r + L.g + NX + k0 * NX, sG, mt * NX);
    copy_out(r + L.h + k0 * NH, sHv, mt * NH);
    if constexpr (LEGS > 0) {
        copy_out(r + L.C + k0 * LEGS * 80, sC, mt * LEGS * 80);
        copy_out(r + L.g + NX + N * NX + k0 * LEGS * 4, sGc, mt * LEGS * 4);
    }
    for (int e = t; e < mt * NX; e += blockDim.x) r[L.grad + Mdl::x_off(N, k0) + e] = sQ[(e / NX) * NZ + e % NX];
    for (int e = t; e < mt * NU; e += blockDim.x) r[L.grad + Mdl::u_off(N, k0) + e] = sQ[(e / NU) * NZ + NX + e % NU];
    if constexpr (Mdl::HC != 0) {
        const int cnt = min(mt, N - 1 - k0) * NU;
        for (int e = t; e < cnt; e += blockDim.x) r[L.Hc + k0 * NU + e] = sHc[e];
    }
}

// One warp per trajectory: reduces the sweep's partials [entries][4] = (objective, barrier, |g|_inf, max h) in a fixed
// order (lane-strided, then xor-tree: deterministic), writes cost[0..1] into the record and, when `summaries` is not
// null, the 32-scalar summary: u_0 (nu <= 24), f, Zsoft, |g|_inf, max h, zero padding.
template <class T>
__global__ void finalize_kernel(const T* __restrict__ partials, int entries, const T* __restrict__ xp_all, long long ld_xp,
                                T* __restrict__ rec_all, long long ld_rec, T* __restrict__ summaries, int cost_off, int u0_off,
                                int nu, int nx, int xm_off, long long batch) {
    const long long b = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (b >= batch) return;
    const int lane = threadIdx.x & 31;
    const T* p = partials + b * entries * 4;
    const T* x = xp_all + b * ld_xp;
    T f = T(0), z = T(0), gmax = T(0), hmax = -INFINITY;
    for (int e = lane; e < entries; e += 32) {
        f += p[4 * e];
        z += p[4 * e + 1];
        gmax = fmax(gmax, p[4 * e + 2]);
        hmax = fmax(hmax, p[4 * e + 3]);
    }
    if (lane < nx) gmax = fmax(gmax, m_abs(x[lane] - x[xm_off + lane]));  // rows x_0 - x_measured
    for (int o = 16; o; o >>= 1) {
        f += __shfl_xor_sync(0xffffffffu, f, o);
        z += __shfl_xor_sync(0xffffffffu, z, o);
        gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
    }
    if (lane == 0) {
        rec_all[b * ld_rec + cost_off]     = f;
        rec_all[b * ld_rec + cost_off + 1] = z;
    }
    if (summaries) {
        T v = T(0);
        if (lane < nu) v = x[u0_off + lane];
        else if (lane == 24) v = f;
        else if (lane == 25) v = z;
        else if (lane == 26) v = gmax;
        else if (lane == 27) v = hmax;
        summaries[b * 32 + lane] = v;
    }
}

// out[b, e] = src[e] >= 0 ? rec[b, src[e]] : constant(-src[e]).  Serves the reference-format outputs
// (CSR value arrays / dependent-variable vectors) from the block record.
template <class T>
__global__ void gather_kernel(const T* __restrict__ rec_all, long long ld_rec, const int* __restrict__ src, int count,
                              T* __restrict__ out_all, long long ld_out, long long batch) {
    const long long b = blockIdx.y;
    const T* rec = rec_all + b * ld_rec;
    T* out       = out_all + b * ld_out;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const int s = src[e];
        out[e] = s >= 0 ? rec[s] : (s == -1 ? T(1) : T(0));
    }
}

// Barrier function Zsoft of soft_sqp.hpp:114-138 on z[b, 0:n]: mode 0 value (one CTA per b), 1 dZ/dz, 2 d2Z/dz2.
template <class T>
__global__ void barrier_kernel(const T* __restrict__ z_all, long long ld_z, int n, T* __restrict__ out_all,
                               long long ld_out, BarrierCoef<T> bar, int mode) {
    const long long b = blockIdx.x;
    const T* z = z_all + b * ld_z;
    T* out     = out_all + b * ld_out;
    __shared__ T red[32];
    T acc = T(0);
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        T b0, dz, d2z;
        barrier_eval(bar, z[e], &b0, &dz, &d2z);
        if (mode == 0) acc += b0;
        else out[e] = mode == 1 ? dz : d2z;
    }
    if (mode != 0) return;
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : T(0);
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) out[0] = acc;
    }
}

// 32-scalar per-trajectory summary: u_0 (nu <= 24), f, Zsoft, |g|_inf, max h, zero padding.
template <class T>
__global__ void summary_kernel(const T* __restrict__ xp_all, long long ld_xp, const T* __restrict__ rec_all,
                               long long ld_rec, T* __restrict__ out_all, RecLayout L, int u0_off, int nu, int m_eq,
                               int m_ineq) {
    const long long b = blockIdx.x;
    const T* rec = rec_all + b * ld_rec;
    T* out       = out_all + b * 32;
    T gmax = T(0), hmax = -INFINITY;
    for (int e = threadIdx.x; e < m_eq; e += 32) gmax = fmax(gmax, m_abs(rec[L.g + e]));
    for (int e = threadIdx.x; e < m_ineq; e += 32) hmax = fmax(hmax, rec[L.h + e]);
    for (int o = 16; o; o >>= 1) {
        gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
    }
    const int l = threadIdx.x;
    T v = T(0);
    if (l < nu) v = xp_all[b * ld_xp + u0_off + l];
    else if (l == 24) v = rec[L.cost];
    else if (l == 25) v = rec[L.cost + 1];
    else if (l == 26) v = gmax;
    else if (l == 27) v = hmax;
    out[l] = v;
}

// Reference-format inequality Jacobian: J_h as [N][NH][NZ] written at the record's A slot (the generated
// `sparse_jacobian` of the *_mpc_ineqs library, function.hpp:224-228).  One thread per (trajectory, node, column).
template <class Mdl, class T>
__global__ void jh_kernel(const T* __restrict__ xp_all, long long ld_xp, T* __restrict__ rec_all, long long ld_rec,
                          int N, int a_off, long long total) {
    constexpr int NX = Mdl::NX, NU = Mdl::NU, NZ = Mdl::NZ, NH = Mdl::NH;
    const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (id >= total) return;
    const int j = int(id % NZ);
    const int k = int((id / NZ) % N);
    const long long b = id / ((long long)NZ * N);
    const T* x = xp_all + b * ld_xp;
    Dual<T> z[NZ], h[NH];
#pragma unroll
    for (int i = 0; i < NX; ++i) z[i] = Dual<T>(x[Mdl::x_off(N, k) + i], i == j ? T(1) : T(0));
#pragma unroll
    for (int i = 0; i < NU; ++i) z[NX + i] = Dual<T>(x[Mdl::u_off(N, k) + i], NX + i == j ? T(1) : T(0));
    Mdl::inequalities(x, N, k, z, h);
    T* out = rec_all + b * ld_rec + a_off + (long long)k * NH * NZ + j;
#pragma unroll
    for (int i = 0; i < NH; ++i) out[i * NZ] = h[i].d;
}

template <class Mdl, class T>
int launch_jh(const T* xp, long long batch, long long ld_xp, T* rec, long long ld_rec, int N, const RecLayout& L,
              cudaStream_t stream) {
    const long long total = batch * N * Mdl::NZ;
    jh_kernel<Mdl, T><<<unsigned((total + 255) / 256), 256, 0, stream>>>(xp, ld_xp, rec, ld_rec, N, L.A, total);
    return int(cudaGetLastError());
}

}  // namespace ub
