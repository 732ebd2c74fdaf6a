// Generic equality-constrained QP solve (include/ungar_b200.h: ungar_b200_kkt_solve_csc) — the device back end of the product's
// osqp++.h stand-in (ungar_b200/include/osqp++.h), i.e. of SoftSQPOptimizer::SolveLocalQPProblem
// (include/ungar/optimization/soft_sqp.hpp:193-233) for NLPs that have no stage-wise solver (qp_schur.cuh / qp_riccati.cuh cover
// the three reference MPC problems).  The quasi-definite KKT matrix
//     [ P + sigma I    A^T   ]
//     [ A            -rho I  ]
// is scattered from the CSC inputs into a dense array on the device by a hand-written kernel and factorised with cuSOLVER's dense
// LU (a plain library factorisation on a fallback path; partial pivoting copes with the indefinite system).  OSQP v0.6.3 itself
// (ADMM) is absent from the reference tree (external/config/osqp/CMakeLists.txt.in:16).
#include <cstdarg>
#include <cstdio>
#include <vector>

#include <cusolverDn.h>

#include "../../include/ungar_b200.h"
#include "abi_internal.h"

namespace {

int kfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return ub_set_error(code, buf);
}

#define UBK_CUDA(call)                                                                                                  \
    do {                                                                                                                \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess) { rc = kfail(UNGAR_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); goto done; } \
    } while (0)
#define UBK_SOLVER(call)                                                                                                \
    do {                                                                                                                \
        cusolverStatus_t s_ = (call);                                                                                   \
        if (s_ != CUSOLVER_STATUS_SUCCESS) { rc = kfail(UNGAR_B200_ECUDA, "%s failed with cuSOLVER status %d", #call, int(s_)); goto done; } \
    } while (0)

// K is column-major, ld = n + m.  One thread per stored entry of P (upper triangle, mirrored) or of A (placed twice).
__global__ void scatter_P(const int* __restrict__ colptr, const int* __restrict__ rowidx, const double* __restrict__ vals, int n,
                          double* __restrict__ K, long long ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    for (int e = colptr[c]; e < colptr[c + 1]; ++e) {
        const int r = rowidx[e];
        if (r > c) continue;  // OSQP reads the upper triangle only
        K[r + c * ld] += vals[e];
        if (r != c) K[c + r * ld] += vals[e];
    }
}
__global__ void scatter_A(const int* __restrict__ colptr, const int* __restrict__ rowidx, const double* __restrict__ vals, int n,
                          double* __restrict__ K, long long ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    for (int e = colptr[c]; e < colptr[c + 1]; ++e) {
        const int r = rowidx[e];
        K[(n + r) + c * ld] += vals[e];
        K[c + (n + r) * ld] += vals[e];
    }
}
__global__ void regularise(double* __restrict__ K, long long ld, int n, int m, double sigma, double rho) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) K[i + i * ld] += sigma;
    else if (i < n + m) K[i + i * ld] -= rho;
}

// Per-device context kept across calls (an SQP loop solves a QP of the same size every iteration): the cuSOLVER handle and the
// device buffers.  Like a model handle, it must be used from one thread at a time.
struct Growable {
    void* ptr  = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&ptr, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
};
struct KktContext {
    cusolverDnHandle_t handle = nullptr;
    Growable K, r, vals, work, ptr, idx, piv, info;
};
KktContext g_ctx[64];

}  // namespace

extern "C" int ungar_b200_kkt_solve_csc(int64_t n, int64_t m, const int32_t* P_colptr, const int32_t* P_rowidx, const double* P_vals,
                                        const double* q, const int32_t* A_colptr, const int32_t* A_rowidx, const double* A_vals,
                                        const double* b, double sigma, double rho, double* x, double* y, int32_t device) {
    if (n <= 0 || m < 0 || !P_colptr || !q || !x || (m > 0 && (!A_colptr || !b)))
        return kfail(UNGAR_B200_EINVAL, "bad sizes or null arrays");
    const long long N = n + m;
    if (N > 16384) return kfail(UNGAR_B200_EUNSUPPORTED, "dense KKT fallback limited to n + m <= 16384 (got %lld)", N);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device >= count)
        return kfail(UNGAR_B200_ECUDA, "no usable CUDA device %d (there is no CPU fallback for the QP solve)", device);
    int rc = UNGAR_B200_OK;
    const int nnzP = P_colptr[n], nnzA = m > 0 ? A_colptr[n] : 0;
    KktContext& C = g_ctx[device & 63];
    double *dK = nullptr, *dr = nullptr, *dvals = nullptr, *dwork = nullptr;
    int *dptr = nullptr, *didx = nullptr, *dpiv = nullptr, *dinfo = nullptr;
    std::vector<double> rhs(static_cast<size_t>(N), 0.0), sol(static_cast<size_t>(N), 0.0);
    int lwork = 0, info = 0;
    const int threads = 128;
    const size_t nnz_max = size_t(std::max(std::max(nnzP, nnzA), 1));
    for (int64_t i = 0; i < n; ++i) rhs[size_t(i)] = -q[i];
    for (int64_t i = 0; i < m; ++i) rhs[size_t(n + i)] = b[i];

    UBK_CUDA(cudaSetDevice(device));
    UBK_CUDA(C.K.reserve(size_t(N) * size_t(N) * sizeof(double)));
    UBK_CUDA(C.r.reserve(size_t(N) * sizeof(double)));
    UBK_CUDA(C.ptr.reserve(size_t(n + 1) * sizeof(int)));
    UBK_CUDA(C.idx.reserve(nnz_max * sizeof(int)));
    UBK_CUDA(C.vals.reserve(nnz_max * sizeof(double)));
    UBK_CUDA(C.piv.reserve(size_t(N) * sizeof(int)));
    UBK_CUDA(C.info.reserve(sizeof(int)));
    dK = static_cast<double*>(C.K.ptr); dr = static_cast<double*>(C.r.ptr); dvals = static_cast<double*>(C.vals.ptr);
    dptr = static_cast<int*>(C.ptr.ptr); didx = static_cast<int*>(C.idx.ptr); dpiv = static_cast<int*>(C.piv.ptr);
    dinfo = static_cast<int*>(C.info.ptr);
    UBK_CUDA(cudaMemset(dK, 0, size_t(N) * size_t(N) * sizeof(double)));
    // P (upper triangle mirrored)
    UBK_CUDA(cudaMemcpy(dptr, P_colptr, size_t(n + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (nnzP) {
        UBK_CUDA(cudaMemcpy(didx, P_rowidx, size_t(nnzP) * sizeof(int), cudaMemcpyHostToDevice));
        UBK_CUDA(cudaMemcpy(dvals, P_vals, size_t(nnzP) * sizeof(double), cudaMemcpyHostToDevice));
        scatter_P<<<unsigned((n + threads - 1) / threads), threads>>>(dptr, didx, dvals, int(n), dK, N);
        ub_count_launch();
    }
    if (nnzA) {
        UBK_CUDA(cudaMemcpy(dptr, A_colptr, size_t(n + 1) * sizeof(int), cudaMemcpyHostToDevice));
        UBK_CUDA(cudaMemcpy(didx, A_rowidx, size_t(nnzA) * sizeof(int), cudaMemcpyHostToDevice));
        UBK_CUDA(cudaMemcpy(dvals, A_vals, size_t(nnzA) * sizeof(double), cudaMemcpyHostToDevice));
        scatter_A<<<unsigned((n + threads - 1) / threads), threads>>>(dptr, didx, dvals, int(n), dK, N);
        ub_count_launch();
    }
    regularise<<<unsigned((N + threads - 1) / threads), threads>>>(dK, N, int(n), int(m), sigma, rho);
    ub_count_launch();
    UBK_CUDA(cudaGetLastError());
    UBK_CUDA(cudaMemcpy(dr, rhs.data(), size_t(N) * sizeof(double), cudaMemcpyHostToDevice));
    if (!C.handle) UBK_SOLVER(cusolverDnCreate(&C.handle));
    UBK_SOLVER(cusolverDnDgetrf_bufferSize(C.handle, int(N), int(N), dK, int(N), &lwork));
    UBK_CUDA(C.work.reserve(size_t(std::max(lwork, 1)) * sizeof(double)));
    dwork = static_cast<double*>(C.work.ptr);
    UBK_SOLVER(cusolverDnDgetrf(C.handle, int(N), int(N), dK, int(N), dwork, dpiv, dinfo));
    UBK_CUDA(cudaMemcpy(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) { rc = kfail(UNGAR_B200_EINVAL, "the KKT matrix is singular (LU pivot %d is zero)", info); goto done; }
    UBK_SOLVER(cusolverDnDgetrs(C.handle, CUBLAS_OP_N, int(N), 1, dK, int(N), dpiv, dr, int(N), dinfo));
    UBK_CUDA(cudaMemcpy(sol.data(), dr, size_t(N) * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) x[i] = sol[size_t(i)];
    if (y)
        for (int64_t i = 0; i < m; ++i) y[i] = sol[size_t(n + i)];
done:
    return rc;
}
