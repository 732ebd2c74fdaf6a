// C ABI of ungar_b200 (include/ungar_b200.h): model handles, structural sparsity, kernel launches.
// No CPU fallback lives here: the only host-side arithmetic is the one-time structural analysis
// (dependency masks -> CSR index arrays) at model_create.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <new>
#include <string>
#include <vector>

#include "../../include/ungar_b200.h"
#include "sweep.cuh"
#include "sweep_structured.cuh"
#include "sweep_structured_compact.cuh"
#include "sweep_tpn.cuh"
#include "sweep_small.cuh"
#include "qp_schur.cuh"
#include "qp_twisted.cuh"
#include "qp_riccati.cuh"
#include "line_search.cuh"

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

// Event ring for device-side timing of the sweep kernel (ungar_b200_set_profiling).
constexpr int kRing = 512;
struct EventRing {
    bool enabled = false;
    cudaEvent_t start[kRing] = {}, stop[kRing] = {};
    bool created[kRing] = {};
    int device[kRing] = {};  // events belong to the device they were created on
    int head = 0, count = 0;
} g_ring;

int fail(int code, const char* fmt, ...);
// The timing events of a ring slot, (re)created on the device that is about to record them.
int ring_slot_ready(int slot, int device) {
    if (g_ring.created[slot] && g_ring.device[slot] == device) return 0;
    if (g_ring.created[slot]) {
        cudaEventDestroy(g_ring.start[slot]);
        cudaEventDestroy(g_ring.stop[slot]);
        g_ring.created[slot] = false;
    }
    if (cudaEventCreate(&g_ring.start[slot]) != cudaSuccess || cudaEventCreate(&g_ring.stop[slot]) != cudaSuccess)
        return fail(UNGAR_B200_ECUDA, "cudaEventCreate failed for the profiling ring");
    g_ring.created[slot] = true;
    g_ring.device[slot]  = device;
    return 0;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
}  // namespace

// Shared with the other translation units of the library (abi_internal.h).
int ub_set_error(int code, const char* message) {
    g_last_error = message;
    return code;
}
void ub_count_launch() { ++g_launches; }

namespace {

#define UB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(UNGAR_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// cudaFuncSetAttribute and the SM count are per DEVICE: a process that drives several GPUs through several handles must configure
// each kernel once per device, not once per process.
constexpr int kMaxDevices = 64;
struct PerDevice {
    int v[kMaxDevices] = {};
    int& operator[](int device) { return v[device & (kMaxDevices - 1)]; }
};

struct DeviceBuffer {
    void* ptr  = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return UNGAR_B200_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&ptr, bytes);
        if (e != cudaSuccess) return fail(UNGAR_B200_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return UNGAR_B200_OK;
    }
    ~DeviceBuffer() {
        if (ptr) cudaFree(ptr);
    }
};

struct FunctionTables {
    int64_t nx = 0, np = 0, ny = 0;
    std::vector<int64_t> jac_rows, jac_cols, hes_rows, hes_cols;
    std::vector<int32_t> y_src, jac_src, hes_src;  // record offsets (>= 0) or constants (-1: 1, -2: 0)
    int32_t *d_y_src = nullptr, *d_jac_src = nullptr, *d_hes_src = nullptr;
    bool has_hessian = false;
};

}  // namespace

struct ungar_b200_model {
    ungar_b200_model_desc desc{};
    int N = 0;
    ungar_b200_kkt_layout layout{};
    ub::RecLayout rl{};
    ub::BarrierCoef<double> bar{};
    FunctionTables fn[4];
    DeviceBuffer stage_cost, ws_records, ws_xp, ws_out, ws_qp, ws_steps, ws_status, ws_info, sched, ws_compact, ws_dense, ws_xp_cached, ws_active;
    const int* active = nullptr;              // SQP loop: list of RUNNING trajectories for the next sweep (device), or null
    const unsigned int* n_active = nullptr;   // and its length (device)
    // calls on one handle share its workspaces and scheduler counters: a call on a new stream first waits for the previous call's work
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
    int64_t cached_batch = 0;  // trajectories whose parameter block sits in ws_xp_cached (ungar_b200_set_parameters)
    // F32 handles: the consumers of the records (QP solve, line search, SQP loop) compute in fp64 on a TWIN handle of the same problem
    // (widen on the device -> F64 kernels -> narrow); see f32_consumers below
    ungar_b200_model* twin = nullptr;
    DeviceBuffer wide_a, wide_b, wide_c, wide_d;
    DeviceBuffer ws_stage;  // contiguous landing zone of narrow host rows (ungar_b200_kkt_step_x), see step_impl
    size_t elem = 8;
    // compact record (quadruped): compact slot -> dense offset (or -2: pad), host copy for the ABI and device copy for the gather
    bool compact = false;
    std::vector<int32_t> c2d;
    int32_t* d_c2d = nullptr;
    int32_t* d_d2c = nullptr;  // the inverse: dense offset -> compact slot (or -2: a structural zero of the dense record)
    int64_t dense_size = 0, rec_size = 0;  // rec_size: length of a record in the handle's format
    // host-buffer pipeline of ungar_b200_kkt_step: H2D of chunk c + 1 on `copy_stream` overlaps the sweep of chunk c
    static constexpr int kChunks = 8;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_entry = nullptr, ev_chunk[kChunks] = {};
};

namespace {

// Orders the calls on one handle across streams: the handle's workspaces, partial-sum buffer and work-claim counters are
// shared by every launch, so a call issued on a different stream than the previous one waits for that one's last recorded event.
struct StreamScope {
    ungar_b200_model& m;
    cudaStream_t stream;
    StreamScope(ungar_b200_model& model, cudaStream_t s) : m(model), stream(s) {
        if (!m.ev_last) cudaEventCreateWithFlags(&m.ev_last, cudaEventDisableTiming);
        if (m.has_last && m.last_stream != stream && m.ev_last) cudaStreamWaitEvent(stream, m.ev_last, 0);
    }
    ~StreamScope() {
        if (m.ev_last) {
            cudaEventRecord(m.ev_last, stream);
            m.last_stream = stream;
            m.has_last    = true;
        }
    }
};

using ub::Dep;

// ---------------------------------------------------------------------------------------------
// Layout
// ---------------------------------------------------------------------------------------------
int64_t round4(int64_t x) { return (x + 3) & ~int64_t(3); }

template <class Mdl>
void make_layout(int N, ungar_b200_kkt_layout& L) {
    L.nx = Mdl::NX; L.nu = Mdl::NU; L.nz = Mdl::NZ; L.horizon = N;
    L.n_dec = Mdl::n_dec(N); L.n_par = Mdl::n_par(N); L.m_eq = Mdl::m_eq(N); L.m_ineq = int64_t(Mdl::NH) * N;
    L.tri = int64_t(Mdl::NZ) * (Mdl::NZ + 1) / 2; L.tri_terminal = int64_t(Mdl::NX) * (Mdl::NX + 1) / 2;
    L.legs = Mdl::LEGS; L.hc_per_node = Mdl::HC ? Mdl::NU : 0;
    int64_t off = 0;
    L.g = off;    off = round4(off + L.m_eq);
    L.A = off;    off = round4(off + N * L.nx * L.nz);
    L.C = off;    off = round4(off + N * L.legs * 80);
    L.h = off;    off = round4(off + L.m_ineq);
    L.cost = off; off = round4(off + 2);
    L.grad = off; off = round4(off + L.n_dec);
    L.H = off;    off = round4(off + N * L.tri);
    L.HN = off;   off = round4(off + L.tri_terminal);
    L.Hc = off;   off = round4(off + (N - 1) * L.hc_per_node);
    L.size = off;
    L.dense_size = off;
    L.compact = 0;
    using K = ub::Compact;
    L.node_stride = K::NODE; L.c_Cs = K::oCs; L.c_Cp = K::oCp; L.c_g = K::oG; L.c_q = K::oQ; L.c_Hd = K::oHd; L.c_Hb = K::oHb; L.c_h = K::oHi;
    L.c_AQ = K::oAQ; L.c_AP = K::oAP; L.tail = K::tail(N); L.t_g0 = K::tG0; L.t_qN = K::tQN; L.t_HN = K::tHN; L.t_cost = K::tCost;
}

// ---------------------------------------------------------------------------------------------
// Structural sparsity from dependency masks (what CppAD's pattern propagation gives the reference,
// function.hpp:98-105 / :529-574), and the map from each CSR nonzero to its slot in the block record.
// ---------------------------------------------------------------------------------------------
template <class Mdl>
void seed_locals(const std::vector<double>& xp, int N, int k, Dep* z, int count) {
    for (int i = 0; i < count; ++i) {
        const int off = i < Mdl::NX ? Mdl::x_off(N, k) + i : Mdl::u_off(N, k) + i - Mdl::NX;
        z[i] = Dep(xp[off], uint64_t(1) << i);
    }
}

template <class Mdl>
void build_tables(ungar_b200_model& M) {
    const int N = M.N;
    const ungar_b200_kkt_layout& L = M.layout;
    constexpr int NX = Mdl::NX, NU = Mdl::NU, NZ = Mdl::NZ, NH = Mdl::NH;
    std::vector<double> xp(L.n_dec + L.n_par);
    for (size_t i = 0; i < xp.size(); ++i) xp[i] = 0.37 + 0.001 * double(i % 97);  // any generic point
    auto local_col = [&](int k, int i) { return int64_t(i < NX ? Mdl::x_off(N, k) + i : Mdl::u_off(N, k) + i - NX); };

    // ---------------- equalities
    FunctionTables& E = M.fn[UNGAR_B200_EQUALITIES];
    E.nx = L.n_dec; E.np = L.n_par; E.ny = L.m_eq;
    for (int i = 0; i < NX; ++i) {  // x_0 - x_measured
        E.jac_rows.push_back(i); E.jac_cols.push_back(i); E.jac_src.push_back(-1);
    }
    for (int k = 0; k < N; ++k) {
        Dep z[NZ], xn[NX];
        seed_locals<Mdl>(xp, N, k, z, NZ);
        Mdl::dynamics(xp.data(), N, k, z, xn);
        for (int r = 0; r < NX; ++r) {
            const int64_t row = NX + int64_t(NX) * k + r;
            auto emit = [&](int i) {
                if (xn[r].m >> i & 1) {
                    E.jac_rows.push_back(row); E.jac_cols.push_back(local_col(k, i));
                    E.jac_src.push_back(int32_t(L.A + (int64_t(k) * NX + r) * NZ + i));
                }
            };
            for (int i = 0; i < NX; ++i) emit(i);  // x_k columns
            E.jac_rows.push_back(row); E.jac_cols.push_back(Mdl::x_off(N, k + 1) + r); E.jac_src.push_back(-1);
            for (int i = NX; i < NZ; ++i) emit(i);  // u_k columns
        }
    }
    if constexpr (Mdl::LEGS > 0) {
        for (int k = 0; k < N; ++k)
            for (int leg = 0; leg < Mdl::LEGS; ++leg) {
                Dep zl[20], rows[4];
                for (int i = 0; i < 20; ++i) zl[i] = Dep(0.3 + 0.01 * i, uint64_t(1) << i);
                Mdl::contact_rows(xp.data(), N, k, leg, zl, rows);
                for (int rr = 0; rr < 4; ++rr) {
                    const int64_t row = NX + int64_t(NX) * N + 16 * k + 4 * leg + rr;
                    const int64_t src = L.C + ((int64_t(k) * 4 + leg) * 4 + rr) * 20;
                    auto emit = [&](int i, int64_t col) {
                        if (rows[rr].m >> i & 1) {
                            E.jac_rows.push_back(row); E.jac_cols.push_back(col); E.jac_src.push_back(int32_t(src + i));
                        }
                    };
                    // ascending global column: pose_{k-1}, pose_k, r_{k-1,leg}, r_{k,leg}
                    if (k) for (int i = 0; i < 7; ++i) emit(10 + i, Mdl::x_off(N, k - 1) + i);
                    for (int i = 0; i < 7; ++i) emit(i, Mdl::x_off(N, k) + i);
                    if (k) for (int i = 0; i < 3; ++i) emit(17 + i, Mdl::u_off(N, k - 1) + 6 * leg + 3 + i);
                    for (int i = 0; i < 3; ++i) emit(7 + i, Mdl::u_off(N, k) + 6 * leg + 3 + i);
                }
            }
    }
    for (int64_t i = 0; i < L.m_eq; ++i) E.y_src.push_back(int32_t(L.g + i));

    // ---------------- inequalities (their Jacobian is consumed inside the sweep; for the reference-format
    // call the values come from the Gauss-Newton-free pass: J_h entries are recovered per node below)
    FunctionTables& I = M.fn[UNGAR_B200_INEQUALITIES];
    I.nx = L.n_dec; I.np = L.n_par; I.ny = L.m_ineq;
    for (int64_t i = 0; i < L.m_ineq; ++i) I.y_src.push_back(int32_t(L.h + i));
    for (int k = 0; k < N; ++k) {
        Dep z[NZ], h[NH];
        seed_locals<Mdl>(xp, N, k, z, NZ);
        Mdl::inequalities(xp.data(), N, k, z, h);
        for (int r = 0; r < NH; ++r)
            for (int i = 0; i < NZ; ++i)
                if (h[r].m >> i & 1) {
                    I.jac_rows.push_back(int64_t(NH) * k + r); I.jac_cols.push_back(local_col(k, i));
                    // J_h is written by the sweep into the (otherwise unused in this mode) A area:
                    // slot (k, r, i) of an [N][NH][NZ] array placed at L.A — see launch_sweep(JH_MODE).
                    I.jac_src.push_back(int32_t(L.A + (int64_t(k) * NH + r) * NZ + i));
                }
    }

    // ---------------- objective
    FunctionTables& O = M.fn[UNGAR_B200_OBJECTIVE];
    O.nx = L.n_dec; O.np = L.n_par; O.ny = 1; O.has_hessian = true;
    O.y_src.push_back(int32_t(L.cost));
    std::vector<uint8_t> used(L.n_dec, 0), coupled(L.n_dec, 0);
    for (int k = 0; k <= N; ++k) {
        Dep z[NZ];
        const int nz = k == N ? NX : NZ;
        seed_locals<Mdl>(xp, N, k, z, nz);
        Mdl::cost_terms(xp.data(), N, k, z, [&](double, const Dep& r, bool counts) {
            for (int i = 0; i < nz; ++i)
                if (r.m >> i & 1) {
                    used[local_col(k, i)] = 1;
                    if (!counts) coupled[local_col(k, i)] = 1;
                }
        });
    }
    // Jacobian 1 x n_dec (ascending columns) and upper-triangular Hessian (diagonal + u_k/u_{k+1} coupling)
    auto grad_src = [&](int64_t col) { return int32_t(L.grad + col); };
    auto locate = [&](int64_t col, int& k, int& i) {
        const int64_t nX = int64_t(NX) * (N + 1);
        if (col < nX) { k = int(col / NX); i = int(col % NX); }
        else { k = int((col - nX) / NU); i = NX + int((col - nX) % NU); }
    };
    for (int64_t col = 0; col < L.n_dec; ++col) {
        if (!used[col]) continue;
        O.jac_rows.push_back(0); O.jac_cols.push_back(col); O.jac_src.push_back(grad_src(col));
        int k, i;
        locate(col, k, i);
        O.hes_rows.push_back(col); O.hes_cols.push_back(col);
        O.hes_src.push_back(k == N ? int32_t(L.HN + ub::tri_index(NX, i, i)) : int32_t(L.H + k * L.tri + ub::tri_index(NZ, i, i)));
        if (coupled[col]) {  // (u_k[i], u_{k+1}[i])
            O.hes_rows.push_back(col); O.hes_cols.push_back(col + NU);
            O.hes_src.push_back(int32_t(L.Hc + int64_t(k) * NU + (i - NX)));
        }
    }

    // ---------------- barrier function Zsoft (soft_sqp.hpp:114-138): dense row / diagonal Hessian
    FunctionTables& S = M.fn[UNGAR_B200_SOFT_INEQUALITIES];
    S.nx = L.m_ineq; S.np = 0; S.ny = 1; S.has_hessian = true;
    for (int64_t i = 0; i < L.m_ineq; ++i) {
        S.jac_rows.push_back(0); S.jac_cols.push_back(i);
        S.hes_rows.push_back(i); S.hes_cols.push_back(i);
    }
}

int upload(const std::vector<int32_t>& host, int32_t** dev) {
    *dev = nullptr;
    if (host.empty()) return UNGAR_B200_OK;
    UB_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), host.size() * sizeof(int32_t)));
    UB_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return UNGAR_B200_OK;
}

template <class T>
ub::BarrierCoef<T> cast_barrier(const ub::BarrierCoef<double>& b) {
    return {T(b.eps), T(b.a1), T(b.b1), T(b.c1), T(b.a2), T(b.b2), T(b.c2), T(b.d2)};
}

// ---------------------------------------------------------------------------------------------
// Launches
// ---------------------------------------------------------------------------------------------
enum SweepMode { MODE_KKT = 0, MODE_PLAIN = 1, MODE_JH = 2, MODE_JAC = 3 };  // MODE_JAC: g and A only (the rest of the record unspecified)

template <class Mdl, class T, int M, bool BARRIER>
int launch_generic(ungar_b200_model& mdl, const T* xp, int64_t batch, int64_t ld_xp, T* rec, int64_t ld_rec,
                   cudaStream_t stream) {
    using Sh = ub::SweepShape<Mdl, M>;
    auto kernel = ub::kkt_sweep_kernel<Mdl, T, M, BARRIER>;
    const int smem = Sh::total * int(sizeof(T));
    static PerDevice configured;  // per instantiation and device
    if (!configured[mdl.desc.device]) {
        UB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[mdl.desc.device] = 1;
    }
    const int tiles = (mdl.N + M - 1) / M;
    const long long grid = (long long)batch * tiles;
    if (grid > 2147483647LL) return fail(UNGAR_B200_EINVAL, "batch too large for one launch (%lld CTAs)", grid);
    int slot = -1;
    if (g_ring.enabled) {
        slot = g_ring.head;
        if (int rc = ring_slot_ready(slot, mdl.desc.device)) return rc;
        UB_CUDA(cudaEventRecord(g_ring.start[slot], stream));
    }
    kernel<<<(unsigned)grid, Sh::THREADS, smem, stream>>>(xp, ld_xp, rec, ld_rec, static_cast<T*>(mdl.stage_cost.ptr),
                                                          mdl.N, tiles, mdl.rl, cast_barrier<T>(mdl.bar));
    if (slot >= 0) {
        UB_CUDA(cudaEventRecord(g_ring.stop[slot], stream));
        g_ring.head  = (g_ring.head + 1) % kRing;
        g_ring.count = g_ring.count < kRing ? g_ring.count + 1 : kRing;
    }
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

// Work-claim counters of the persistent grids: zeroed once, re-armed by the kernels themselves (the last CTA / warp to leave).
int ensure_sched(ungar_b200_model& mdl, cudaStream_t stream) {
    if (mdl.sched.ptr) return UNGAR_B200_OK;
    if (int rc = mdl.sched.reserve(2 * sizeof(unsigned int))) return rc;
    UB_CUDA(cudaMemsetAsync(mdl.sched.ptr, 0, 2 * sizeof(unsigned int), stream));
    return UNGAR_B200_OK;
}

// Structured quadruped fp64 sweep (sweep_structured.cuh).  Needs paired nodes (even horizon) and 16-byte aligned
// block slices for the TMA bulk stores; anything else takes the generic kernel.
bool structured_applicable(const ungar_b200_model& mdl, const void* rec, int64_t ld_rec) {
    static const bool forced_generic = [] {
        const char* e = getenv("UNGAR_B200_FORCE_GENERIC");
        return e && e[0] == '1';
    }();
    return !forced_generic && mdl.desc.kind == UNGAR_B200_QUADRUPED && mdl.desc.dtype == UNGAR_B200_F64 &&
           mdl.N % 2 == 0 && ld_rec % 2 == 0 && (reinterpret_cast<uintptr_t>(rec) & 15) == 0;
}

template <bool BARRIER>
int launch_structured(ungar_b200_model& mdl, const double* xp, int64_t batch, int64_t ld_xp, double* rec, int64_t ld_rec,
                      cudaStream_t stream) {
    using Q = ub::QuadrupedStructured;
    auto kernel = ub::quadruped_structured_kernel<BARRIER>;
    static PerDevice configured, sm_counts;
    if (!configured[mdl.desc.device]) {
        UB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Q::SMEM_BYTES));
        UB_CUDA(cudaDeviceGetAttribute(&sm_counts[mdl.desc.device], cudaDevAttrMultiProcessorCount, mdl.desc.device));
        configured[mdl.desc.device] = 1;
    }
    const int sm_count = sm_counts[mdl.desc.device];
    const int runs_per_traj = (mdl.N + 9) / 10;
    const int run_len       = 2 * ((mdl.N + 2 * runs_per_traj - 1) / (2 * runs_per_traj));
    const long long total_runs = (long long)batch * runs_per_traj;
    const unsigned grid = unsigned(std::min<long long>(total_runs, (long long)sm_count * 6));  // persistent: 6 teams / SM
    if (int rc = ensure_sched(mdl, stream)) return rc;
    int slot = -1;
    if (g_ring.enabled) {
        slot = g_ring.head;
        if (int rc = ring_slot_ready(slot, mdl.desc.device)) return rc;
        UB_CUDA(cudaEventRecord(g_ring.start[slot], stream));
    }
    kernel<<<grid, Q::WARPS * 32, Q::SMEM_BYTES, stream>>>(xp, ld_xp, rec, ld_rec, static_cast<double*>(mdl.stage_cost.ptr),
                                                           mdl.N, run_len, runs_per_traj, total_runs, mdl.rl, mdl.bar, static_cast<unsigned int*>(mdl.sched.ptr));
    if (slot >= 0) {
        UB_CUDA(cudaEventRecord(g_ring.stop[slot], stream));
        g_ring.head  = (g_ring.head + 1) % kRing;
        g_ring.count = g_ring.count < kRing ? g_ring.count + 1 : kRing;
    }
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

// Structured quadruped fp64 sweep writing the COMPACT record (sweep_structured_compact.cuh): any horizon; the record must be
// 16-byte aligned with an even stride (one TMA bulk store per node).
template <bool BARRIER>
int launch_structured_compact(ungar_b200_model& mdl, const double* xp, int64_t batch, int64_t ld_xp, double* rec, int64_t ld_rec,
                              cudaStream_t stream) {
    using Q = ub::QuadrupedCompactSweep;
    auto kernel = ub::quadruped_compact_kernel<BARRIER>;
    static PerDevice configured, sm_counts;
    if (!configured[mdl.desc.device]) {
        UB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Q::SMEM_BYTES));
        UB_CUDA(cudaDeviceGetAttribute(&sm_counts[mdl.desc.device], cudaDevAttrMultiProcessorCount, mdl.desc.device));
        configured[mdl.desc.device] = 1;
    }
    if ((reinterpret_cast<uintptr_t>(rec) & 15) != 0 || (ld_rec & 1) != 0)
        return fail(UNGAR_B200_EINVAL, "compact records must be 16-byte aligned with an even stride (TMA bulk stores)");
    const int sm_count = sm_counts[mdl.desc.device];
    const int runs_per_traj = (mdl.N + 9) / 10;
    const int run_len       = 2 * ((mdl.N + 2 * runs_per_traj - 1) / (2 * runs_per_traj));
    const long long total_runs = (long long)batch * runs_per_traj;
    const unsigned grid = unsigned(std::min<long long>(total_runs, (long long)sm_count * Q::TEAMS_PER_SM));  // persistent
    if (int rc = ensure_sched(mdl, stream)) return rc;
    int slot = -1;
    if (g_ring.enabled) {
        slot = g_ring.head;
        if (int rc = ring_slot_ready(slot, mdl.desc.device)) return rc;
        UB_CUDA(cudaEventRecord(g_ring.start[slot], stream));
    }
    kernel<<<grid, Q::WARPS * 32, Q::SMEM_BYTES, stream>>>(xp, ld_xp, rec, ld_rec, static_cast<double*>(mdl.stage_cost.ptr), mdl.N, run_len,
                                                           runs_per_traj, total_runs, mdl.bar, static_cast<unsigned int*>(mdl.sched.ptr), mdl.active, mdl.n_active);
    if (slot >= 0) {
        UB_CUDA(cudaEventRecord(g_ring.stop[slot], stream));
        g_ring.head  = (g_ring.head + 1) % kRing;
        g_ring.count = g_ring.count < kRing ? g_ring.count + 1 : kRing;
    }
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

// Thread-per-node sweep (sweep_tpn.cuh) for the models without contact rows; one CTA per trajectory, persistent.
template <class Mdl, class T>
bool tpn_applicable(const ungar_b200_model& mdl) {
    const char* e = getenv("UNGAR_B200_FORCE_GENERIC");
    if (e && e[0] == '1') return false;
    const char* t = getenv("UNGAR_B200_FORCE_TPN");
    if (!Mdl::TPN_DEFAULT && !(t && t[0] == '1')) return false;
    const size_t smem = size_t(ub::TpnShape<Mdl>::offsets(mdl.N, int(mdl.layout.n_dec + mdl.layout.n_par)).total) * sizeof(T);
    return Mdl::LEGS == 0 && mdl.N + 1 <= 1024 && smem <= 200 * 1024;
}

template <class Mdl, class T, bool BARRIER, bool STRUCT>
int launch_tpn_s(ungar_b200_model& mdl, const T* xp, int64_t batch, int64_t ld_xp, T* rec, int64_t ld_rec, cudaStream_t stream) {
    auto kernel = ub::tpn_sweep_kernel<Mdl, T, BARRIER, STRUCT>;
    const int n_xp = int(mdl.layout.n_dec + mdl.layout.n_par);
    const auto offs = ub::TpnShape<Mdl>::offsets(mdl.N, n_xp);
    const int smem = offs.total * int(sizeof(T));
    const int threads = ((mdl.N + 1 + 31) / 32) * 32;
    // TMA bulk stores need 16-byte aligned global rows: record base and stride
    const int bulk = ((reinterpret_cast<uintptr_t>(rec) & 15) == 0 && (ld_rec * sizeof(T)) % 16 == 0) ? 1 : 0;
    static PerDevice configured_smem, sm_counts;
    if (configured_smem[mdl.desc.device] < smem) {
        UB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        UB_CUDA(cudaDeviceGetAttribute(&sm_counts[mdl.desc.device], cudaDevAttrMultiProcessorCount, mdl.desc.device));
        configured_smem[mdl.desc.device] = smem;
    }
    const int sm_count = sm_counts[mdl.desc.device];
    int per_sm = 1;
    UB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    const unsigned grid = unsigned(std::min<long long>(batch, (long long)sm_count * std::max(per_sm, 1)));
    int slot = -1;
    if (g_ring.enabled) {
        slot = g_ring.head;
        if (int rc = ring_slot_ready(slot, mdl.desc.device)) return rc;
        UB_CUDA(cudaEventRecord(g_ring.start[slot], stream));
    }
    kernel<<<grid, threads, smem, stream>>>(xp, ld_xp, rec, ld_rec, static_cast<T*>(mdl.stage_cost.ptr), mdl.N, n_xp, batch,
                                            mdl.rl, cast_barrier<T>(mdl.bar), offs, bulk);
    if (slot >= 0) {
        UB_CUDA(cudaEventRecord(g_ring.stop[slot], stream));
        g_ring.head  = (g_ring.head + 1) % kRing;
        g_ring.count = g_ring.count < kRing ? g_ring.count + 1 : kRing;
    }
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

// Warp-team sweep for the small models (sweep_small.cuh): 8 lanes per node, 4-node staging image, TMA bulk stores.
template <class Mdl, class T>
bool small_applicable(const ungar_b200_model& mdl, const void* rec, int64_t ld_rec) {
    const char* e = getenv("UNGAR_B200_FORCE_GENERIC");
    const char* t = getenv("UNGAR_B200_FORCE_TPN");
    if ((e && e[0] == '1') || (t && t[0] == '1')) return false;
    return Mdl::LEGS == 0 && (reinterpret_cast<uintptr_t>(rec) & 15) == 0 && (ld_rec * sizeof(T)) % 16 == 0;
}

template <class Mdl, class T, bool BARRIER, bool JAC = false>
int launch_small(ungar_b200_model& mdl, const T* xp, int64_t batch, int64_t ld_xp, T* rec, int64_t ld_rec, cudaStream_t stream) {
    constexpr int WARPS = 4;
    auto kernel = ub::small_team_kernel<Mdl, T, BARRIER, WARPS, JAC>;
    const int n_xp = int(mdl.layout.n_dec + mdl.layout.n_par);
    const auto offs = ub::SmallShape<Mdl, T>::offsets(mdl.N, n_xp);
    const int smem = WARPS * offs.total * int(sizeof(T));
    static PerDevice configured_smem, sm_counts;
    if (configured_smem[mdl.desc.device] < smem) {
        UB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        UB_CUDA(cudaDeviceGetAttribute(&sm_counts[mdl.desc.device], cudaDevAttrMultiProcessorCount, mdl.desc.device));
        configured_smem[mdl.desc.device] = smem;
    }
    const int sm_count = sm_counts[mdl.desc.device];
    int per_sm = 1;
    UB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem));
    const long long want = (batch + WARPS - 1) / WARPS;
    const unsigned grid = unsigned(std::min<long long>(want, (long long)sm_count * std::max(per_sm, 1)));
    if (int rc = ensure_sched(mdl, stream)) return rc;
    int slot = -1;
    if (g_ring.enabled) {
        slot = g_ring.head;
        if (int rc = ring_slot_ready(slot, mdl.desc.device)) return rc;
        UB_CUDA(cudaEventRecord(g_ring.start[slot], stream));
    }
    kernel<<<grid, WARPS * 32, smem, stream>>>(xp, ld_xp, rec, ld_rec, static_cast<T*>(mdl.stage_cost.ptr), mdl.N, n_xp, batch,
                                               mdl.rl, cast_barrier<T>(mdl.bar), offs, static_cast<unsigned int*>(mdl.sched.ptr), mdl.active, mdl.n_active);
    if (slot >= 0) {
        UB_CUDA(cudaEventRecord(g_ring.stop[slot], stream));
        g_ring.head  = (g_ring.head + 1) % kRing;
        g_ring.count = g_ring.count < kRing ? g_ring.count + 1 : kRing;
    }
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

// UNGAR_B200_TPN_STRUCT=0/1 overrides the per-model default (hand-structured node Jacobian vs in-register vector duals).
template <class Mdl, class T, bool BARRIER>
int launch_tpn(ungar_b200_model& mdl, const T* xp, int64_t batch, int64_t ld_xp, T* rec, int64_t ld_rec, cudaStream_t stream) {
    static const int forced = [] {
        const char* e = getenv("UNGAR_B200_TPN_STRUCT");
        return e ? (e[0] == '1' ? 1 : 0) : -1;
    }();
    const bool structured = forced >= 0 ? forced == 1 : Mdl::TPN_STRUCTURED_DEFAULT;
    return structured ? launch_tpn_s<Mdl, T, BARRIER, true>(mdl, xp, batch, ld_xp, rec, ld_rec, stream)
                      : launch_tpn_s<Mdl, T, BARRIER, false>(mdl, xp, batch, ld_xp, rec, ld_rec, stream);
}

template <class Mdl, class T, int M>
int launch_sweep_t(ungar_b200_model& mdl, const void* xp, int64_t batch, int64_t ld_xp, void* rec, int64_t ld_rec,
                   int mode, void* summaries, cudaStream_t stream, bool compact_out) {
    int entries = mdl.N + 1;  // partial entries per trajectory (generic sweep: one per node)
    if (int rc = mdl.stage_cost.reserve(size_t(batch) * (mdl.N + 1) * 4 * sizeof(T))) return rc;
    const T* x = static_cast<const T*>(xp);
    T* r       = static_cast<T*>(rec);
    int rc;
    if (mode == MODE_JH && !compact_out) {
        rc = ub::launch_jh<Mdl, T>(x, batch, ld_xp, r, ld_rec, mdl.N, mdl.rl, stream);
        if (rc == 0) ++g_launches;
        else return fail(UNGAR_B200_ECUDA, "J_h kernel launch failed: %s", cudaGetErrorString(cudaError_t(rc)));
        return UNGAR_B200_OK;
    }
    if (compact_out) {
        if constexpr (std::is_same<Mdl, ub::Quadruped>::value && std::is_same<T, double>::value) {
            if (mode == MODE_JH) return fail(UNGAR_B200_EINVAL, "internal: J_h is served from dense records");
            entries = 2 * ((mdl.N + 9) / 10);
            rc = mode == MODE_KKT ? launch_structured_compact<true>(mdl, x, batch, ld_xp, r, ld_rec, stream)
                                  : launch_structured_compact<false>(mdl, x, batch, ld_xp, r, ld_rec, stream);
        } else
            return fail(UNGAR_B200_EUNSUPPORTED, "compact records exist for the quadruped in F64 only");
    } else if constexpr (std::is_same<Mdl, ub::Quadruped>::value && std::is_same<T, double>::value) {
        if (structured_applicable(mdl, rec, ld_rec)) {
            entries = 2 * ((mdl.N + 9) / 10);  // structured sweep: one per (run, warp)
            rc = mode == MODE_KKT ? launch_structured<true>(mdl, x, batch, ld_xp, r, ld_rec, stream)
                                  : launch_structured<false>(mdl, x, batch, ld_xp, r, ld_rec, stream);
        } else
            rc = mode == MODE_KKT ? launch_generic<Mdl, T, M, true>(mdl, x, batch, ld_xp, r, ld_rec, stream)
                                  : launch_generic<Mdl, T, M, false>(mdl, x, batch, ld_xp, r, ld_rec, stream);
    } else if constexpr (Mdl::LEGS == 0) {
        const size_t small_smem = 4 * size_t(ub::SmallShape<Mdl, T>::offsets(mdl.N, int(mdl.layout.n_dec + mdl.layout.n_par)).total) * sizeof(T);
        if (small_applicable<Mdl, T>(mdl, rec, ld_rec) && small_smem <= 200 * 1024) {
            entries = 1;  // one partial per trajectory
            rc = mode == MODE_KKT   ? launch_small<Mdl, T, true>(mdl, x, batch, ld_xp, r, ld_rec, stream)
                 : mode == MODE_JAC ? launch_small<Mdl, T, false, true>(mdl, x, batch, ld_xp, r, ld_rec, stream)
                                    : launch_small<Mdl, T, false>(mdl, x, batch, ld_xp, r, ld_rec, stream);
        } else if (tpn_applicable<Mdl, T>(mdl))
            rc = mode == MODE_KKT ? launch_tpn<Mdl, T, true>(mdl, x, batch, ld_xp, r, ld_rec, stream)
                                  : launch_tpn<Mdl, T, false>(mdl, x, batch, ld_xp, r, ld_rec, stream);
        else
            rc = mode == MODE_KKT ? launch_generic<Mdl, T, M, true>(mdl, x, batch, ld_xp, r, ld_rec, stream)
                                  : launch_generic<Mdl, T, M, false>(mdl, x, batch, ld_xp, r, ld_rec, stream);
    } else if (mode == MODE_KKT) rc = launch_generic<Mdl, T, M, true>(mdl, x, batch, ld_xp, r, ld_rec, stream);
    else rc = launch_generic<Mdl, T, M, false>(mdl, x, batch, ld_xp, r, ld_rec, stream);
    if (rc) return rc;
    if (mode == MODE_JAC) return UNGAR_B200_OK;  // no objective / barrier partials to reduce
    const ungar_b200_kkt_layout& L = mdl.layout;
    const int threads = 128;  // 4 trajectories per CTA
    ub::finalize_kernel<T><<<unsigned((batch * 32 + threads - 1) / threads), threads, 0, stream>>>(
        static_cast<const T*>(mdl.stage_cost.ptr), entries, x, ld_xp, r, ld_rec, static_cast<T*>(summaries),
        compact_out ? int(ub::Compact::tail(mdl.N) + ub::Compact::tCost) : mdl.rl.cost,
        int(L.nx * (L.horizon + 1)), int(L.nu), int(L.nx), int(Mdl::xm_off(mdl.N)), batch);
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

template <class T>
int launch_gather_t(const void* rec, int64_t ld_rec, const int32_t* src, int64_t count, void* out, int64_t ld_out, int64_t batch, cudaStream_t stream);
int ensure_d2c(ungar_b200_model& mdl);

int launch_sweep(ungar_b200_model& mdl, const void* xp, int64_t batch, int64_t ld_xp, void* rec, int64_t ld_rec,
                 int mode, void* summaries, cudaStream_t stream, bool compact_out = false) {
    const bool f64 = mdl.desc.dtype == UNGAR_B200_F64;
    // Quadruped fp64 into a DENSE record the structured kernel cannot take (odd horizon: its TMA stores pair the nodes; records that are
    // not 16-byte aligned): the compact sweep — one chunk per node, any horizon — into a workspace, then one expansion pass that writes
    // every dense slot once (value or structural zero).  ~0.5 ms per 1024 x N=100 instead of the generic kernel's 2.4 ms.
    static const bool forced_generic = [] {
        const char* e = getenv("UNGAR_B200_FORCE_GENERIC");
        return e && e[0] == '1';
    }();
    if (mdl.desc.kind == UNGAR_B200_QUADRUPED && f64 && !compact_out && !forced_generic && (mode == MODE_KKT || mode == MODE_PLAIN) &&
        !structured_applicable(mdl, rec, ld_rec)) {
        const int64_t csize = ub::Compact::size(mdl.N);
        if (int rc = ensure_d2c(mdl)) return rc;
        if (int rc = mdl.ws_compact.reserve(size_t(batch) * csize * sizeof(double))) return rc;
        if (int rc = launch_sweep_t<ub::Quadruped, double, 4>(mdl, xp, batch, ld_xp, mdl.ws_compact.ptr, csize, mode, summaries, stream, true)) return rc;
        for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
            const int64_t nb = std::min<int64_t>(65535, batch - b0);
            if (int rc = launch_gather_t<double>(static_cast<const double*>(mdl.ws_compact.ptr) + b0 * csize, csize, mdl.d_d2c, mdl.dense_size,
                                                 static_cast<double*>(rec) + b0 * ld_rec, ld_rec, nb, stream)) return rc;
        }
        return UNGAR_B200_OK;
    }
    switch (mdl.desc.kind) {
        case UNGAR_B200_QUADROTOR:
            return f64 ? launch_sweep_t<ub::Quadrotor, double, 15>(mdl, xp, batch, ld_xp, rec, ld_rec, mode, summaries, stream, compact_out)
                       : launch_sweep_t<ub::Quadrotor, float, 15>(mdl, xp, batch, ld_xp, rec, ld_rec, mode, summaries, stream, compact_out);
        case UNGAR_B200_RC_CAR:
            return f64 ? launch_sweep_t<ub::RcCar, double, 30>(mdl, xp, batch, ld_xp, rec, ld_rec, mode, summaries, stream, compact_out)
                       : launch_sweep_t<ub::RcCar, float, 30>(mdl, xp, batch, ld_xp, rec, ld_rec, mode, summaries, stream, compact_out);
        case UNGAR_B200_QUADRUPED:
            return f64 ? launch_sweep_t<ub::Quadruped, double, 4>(mdl, xp, batch, ld_xp, rec, ld_rec, mode, summaries, stream, compact_out)
                       : launch_sweep_t<ub::Quadruped, float, 4>(mdl, xp, batch, ld_xp, rec, ld_rec, mode, summaries, stream, compact_out);
    }
    return fail(UNGAR_B200_EINVAL, "unknown model kind %d", mdl.desc.kind);
}

template <class T>
int launch_gather_t(const void* rec, int64_t ld_rec, const int32_t* src, int64_t count, void* out, int64_t ld_out,
                    int64_t batch, cudaStream_t stream) {
    if (count == 0 || batch == 0) return UNGAR_B200_OK;
    const int threads = 256;
    dim3 grid(unsigned(std::min<int64_t>((count + threads - 1) / threads, 64)), unsigned(batch));
    if (batch > 65535) return fail(UNGAR_B200_EINVAL, "reference-format calls support batch <= 65535 (got %lld)", (long long)batch);
    ub::gather_kernel<T><<<grid, threads, 0, stream>>>(static_cast<const T*>(rec), ld_rec, src, int(count),
                                                      static_cast<T*>(out), ld_out, batch);
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

int launch_gather(ungar_b200_model& mdl, const void* rec, int64_t ld_rec, const int32_t* src, int64_t count, void* out,
                  int64_t ld_out, int64_t batch, cudaStream_t stream) {
    return mdl.desc.dtype == UNGAR_B200_F64
               ? launch_gather_t<double>(rec, ld_rec, src, count, out, ld_out, batch, stream)
               : launch_gather_t<float>(rec, ld_rec, src, count, out, ld_out, batch, stream);
}

template <class T>
int launch_barrier_t(ungar_b200_model& mdl, const void* z, int64_t ld_z, void* out, int64_t ld_out, int64_t batch,
                     int mode, cudaStream_t stream) {
    ub::barrier_kernel<T><<<unsigned(batch), 256, 0, stream>>>(static_cast<const T*>(z), ld_z, int(mdl.layout.m_ineq),
                                                              static_cast<T*>(out), ld_out, cast_barrier<T>(mdl.bar), mode);
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

// QP solve launches.  `skip_status` (may be null): trajectories whose status is not RUNNING are skipped.
// Quadruped: twisted stage-wise Schur complement on the COMPACT record (qp_twisted.cuh; the contact rows are extra equalities).
// `rec_is_compact` false: the caller's records are dense -> one gather pass into the handle's compact workspace first.
int ensure_c2d(ungar_b200_model& mdl) {
    if (mdl.d_c2d) return UNGAR_B200_OK;
    if (mdl.c2d.empty()) {
        const ungar_b200_kkt_layout& L = mdl.layout;
        mdl.c2d = ub::compact_to_dense_map(mdl.N, ub::DenseOffsets{L.g, L.A, L.C, L.h, L.cost, L.grad, L.H, L.HN});
    }
    UB_CUDA(cudaMalloc(reinterpret_cast<void**>(&mdl.d_c2d), mdl.c2d.size() * sizeof(int32_t)));
    UB_CUDA(cudaMemcpy(mdl.d_c2d, mdl.c2d.data(), mdl.c2d.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return UNGAR_B200_OK;
}

int ensure_d2c(ungar_b200_model& mdl) {
    if (mdl.d_d2c) return UNGAR_B200_OK;
    if (mdl.c2d.empty()) {
        const ungar_b200_kkt_layout& L = mdl.layout;
        mdl.c2d = ub::compact_to_dense_map(mdl.N, ub::DenseOffsets{L.g, L.A, L.C, L.h, L.cost, L.grad, L.H, L.HN});
    }
    std::vector<int32_t> d2c(size_t(mdl.dense_size), -2);
    for (size_t i = 0; i < mdl.c2d.size(); ++i)
        if (mdl.c2d[i] >= 0) d2c[size_t(mdl.c2d[i])] = int32_t(i);
    UB_CUDA(cudaMalloc(reinterpret_cast<void**>(&mdl.d_d2c), d2c.size() * sizeof(int32_t)));
    UB_CUDA(cudaMemcpy(mdl.d_d2c, d2c.data(), d2c.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return UNGAR_B200_OK;
}

int launch_qp_twisted(ungar_b200_model& mdl, const void* rec, bool rec_is_compact, int64_t batch, int64_t ld_rec, void* steps, int64_t ld_steps,
                      void* mult, int64_t ld_mult, const int32_t* skip_status, cudaStream_t stream) {
    using Q = ub::QpT;
    const int64_t csize = ub::Compact::size(mdl.N);
    if (!rec_is_compact) {
        if (int rc = ensure_c2d(mdl)) return rc;
        if (int rc = mdl.ws_compact.reserve(size_t(batch) * csize * sizeof(double))) return rc;
        for (int64_t b0 = 0; b0 < batch; b0 += 65535) {  // gather grid: blockIdx.y = trajectory
            const int64_t nb = std::min<int64_t>(65535, batch - b0);
            if (int rc = launch_gather_t<double>(static_cast<const double*>(rec) + b0 * ld_rec, ld_rec, mdl.d_c2d, csize,
                                                 static_cast<double*>(mdl.ws_compact.ptr) + b0 * csize, csize, nb, stream)) return rc;
        }
        rec = mdl.ws_compact.ptr;
        ld_rec = csize;
    } else if ((reinterpret_cast<uintptr_t>(rec) & 15) != 0 || (ld_rec & 1) != 0) {
        return fail(UNGAR_B200_EINVAL, "compact records must be 16-byte aligned with an even stride (TMA bulk loads)");
    }
    if (int rc = mdl.ws_qp.reserve(size_t(batch) * (mdl.N + 1) * Q::WS_GROUP * sizeof(double))) return rc;
    static PerDevice configured;
    if (!configured[mdl.desc.device]) {
        UB_CUDA(cudaFuncSetAttribute(ub::qp_twisted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Q::SMEM_BYTES));
        // one CTA of 7 trajectories x 29.4 KB per SM needs the largest shared-memory carveout
        UB_CUDA(cudaFuncSetAttribute(ub::qp_twisted_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured[mdl.desc.device] = 1;
    }
    if (batch > 2147483647LL) return fail(UNGAR_B200_EINVAL, "batch too large for one launch");
    ub::qp_twisted_kernel<<<unsigned((batch + Q::SLOTS - 1) / Q::SLOTS), Q::THREADS, Q::SMEM_BYTES, stream>>>(
        static_cast<const double*>(rec), ld_rec, static_cast<double*>(mdl.ws_qp.ptr), static_cast<double*>(steps), ld_steps,
        static_cast<double*>(mult), ld_mult, mdl.N, batch, 1e-9, skip_status);
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

// Quadrotor, RC car: Riccati recursion (qp_riccati.cuh; only the initial condition and the defects are equalities).
template <class Mdl>
int launch_qp_riccati(ungar_b200_model& mdl, const void* rec, int64_t batch, int64_t ld_rec, void* steps, int64_t ld_steps, void* mult,
                      int64_t ld_mult, const int32_t* skip_status, cudaStream_t stream) {
    using R = ub::RiccatiShape<Mdl>;
    if (int rc = mdl.ws_qp.reserve(size_t(batch) * mdl.N * R::WS_STAGE * sizeof(double))) return rc;
    auto kernel = ub::qp_riccati_kernel<Mdl>;
    static PerDevice configured;
    if (!configured[mdl.desc.device]) {
        UB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, R::SMEM_BYTES));
        configured[mdl.desc.device] = 1;
    }
    const unsigned grid = unsigned((batch + R::WARPS - 1) / R::WARPS);
    kernel<<<grid, R::WARPS * 32, R::SMEM_BYTES, stream>>>(static_cast<const double*>(rec), ld_rec, static_cast<double*>(mdl.ws_qp.ptr),
                                                           static_cast<double*>(steps), ld_steps, static_cast<double*>(mult), ld_mult,
                                                           mdl.N, batch, mdl.rl, skip_status);
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

int launch_qp(ungar_b200_model& mdl, const void* rec, bool rec_is_compact, int64_t batch, int64_t ld_rec, void* steps, int64_t ld_steps, void* mult,
              int64_t ld_mult, const int32_t* skip_status, cudaStream_t stream) {
    switch (mdl.desc.kind) {
        case UNGAR_B200_QUADROTOR: return launch_qp_riccati<ub::Quadrotor>(mdl, rec, batch, ld_rec, steps, ld_steps, mult, ld_mult, skip_status, stream);
        case UNGAR_B200_RC_CAR: return launch_qp_riccati<ub::RcCar>(mdl, rec, batch, ld_rec, steps, ld_steps, mult, ld_mult, skip_status, stream);
        case UNGAR_B200_QUADRUPED: return launch_qp_twisted(mdl, rec, rec_is_compact, batch, ld_rec, steps, ld_steps, mult, ld_mult, skip_status, stream);
    }
    return fail(UNGAR_B200_EINVAL, "unknown model kind %d", mdl.desc.kind);
}

int check_options(const ungar_b200_sqp_options& o) {
    if (o.max_iterations < 0) return fail(UNGAR_B200_EINVAL, "max_iterations %d < 0", o.max_iterations);
    if (!(o.gamma_alpha > 0.0 && o.gamma_alpha < 1.0)) return fail(UNGAR_B200_EINVAL, "gamma_alpha must lie in (0, 1)");
    if (!(o.alpha_min > 0.0)) return fail(UNGAR_B200_EINVAL, "alpha_min must be positive (the backtracking loop would not end)");
    return UNGAR_B200_OK;
}

template <class Mdl>
int launch_line_search_t(ungar_b200_model& mdl, double* xp, int64_t batch, int64_t ld_xp, const double* steps, int64_t ld_steps,
                         const ungar_b200_sqp_options& o, int32_t* status, double* info, cudaStream_t stream) {
    auto kernel = ub::line_search_kernel<Mdl, double>;
    const int smem = int((mdl.layout.n_dec + mdl.layout.n_par) * sizeof(double));
    if (smem > 220 * 1024) return fail(UNGAR_B200_EUNSUPPORTED, "horizon too long for the shared-memory trial point (%d bytes)", smem);
    static PerDevice configured;  // per instantiation and device: largest size configured so far
    if (smem > configured[mdl.desc.device]) {
        UB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[mdl.desc.device] = smem;
    }
    const ub::LineSearchParams P{o.alpha_min, o.theta_min, o.theta_max, o.eta, o.gamma_phi, o.gamma_theta, o.gamma_alpha,
                                 o.constraint_violation_multiplier, o.objective_tolerance};
    kernel<<<unsigned(batch), ub::LS_THREADS, smem, stream>>>(xp, ld_xp, steps, ld_steps, mdl.N, mdl.bar, P, status, info);
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

int launch_line_search(ungar_b200_model& mdl, double* xp, int64_t batch, int64_t ld_xp, const double* steps, int64_t ld_steps,
                       const ungar_b200_sqp_options& o, int32_t* status, double* info, cudaStream_t stream) {
    switch (mdl.desc.kind) {
        case UNGAR_B200_QUADROTOR: return launch_line_search_t<ub::Quadrotor>(mdl, xp, batch, ld_xp, steps, ld_steps, o, status, info, stream);
        case UNGAR_B200_RC_CAR: return launch_line_search_t<ub::RcCar>(mdl, xp, batch, ld_xp, steps, ld_steps, o, status, info, stream);
        case UNGAR_B200_QUADRUPED: return launch_line_search_t<ub::Quadruped>(mdl, xp, batch, ld_xp, steps, ld_steps, o, status, info, stream);
    }
    return fail(UNGAR_B200_EINVAL, "unknown model kind %d", mdl.desc.kind);
}

bool valid_fn(int32_t f) { return f >= 0 && f <= 3; }

// What a reference-format call needs: which sweep mode fills the slots its gather map points at.
enum Want { WANT_Y = 0, WANT_JAC = 1, WANT_HES = 2 };

int reference_call(ungar_b200_model* mdl, int32_t function, int want, const void* xp, int64_t batch, int64_t ld_xp,
                   void* out, int64_t ld_out, int32_t mem, void* stream_) {
    if (!mdl) return fail(UNGAR_B200_EINVAL, "null model");
    if (!valid_fn(function)) return fail(UNGAR_B200_EINVAL, "unknown function %d", function);
    if (batch < 0 || (batch > 0 && (!xp || !out))) return fail(UNGAR_B200_EINVAL, "null buffer");
    if (mem != UNGAR_B200_MEM_DEVICE && mem != UNGAR_B200_MEM_HOST) return fail(UNGAR_B200_EINVAL, "unknown mem %d", mem);
    FunctionTables& F = mdl->fn[function];
    if (want == WANT_HES && !F.has_hessian)
        return fail(UNGAR_B200_EUNSUPPORTED, "the Hessian is implemented only for scalar functions (function.hpp:136-137)");
    const int64_t n_in  = F.nx + F.np;
    const int64_t n_out = want == WANT_Y ? F.ny : want == WANT_JAC ? int64_t(F.jac_rows.size()) : int64_t(F.hes_rows.size());
    if (ld_xp < n_in) return fail(UNGAR_B200_EINVAL, "ld_xp %lld < %lld", (long long)ld_xp, (long long)n_in);
    if (ld_out < n_out) return fail(UNGAR_B200_EINVAL, "output stride %lld < %lld", (long long)ld_out, (long long)n_out);
    if (batch == 0) return UNGAR_B200_OK;
    UB_CUDA(cudaSetDevice(mdl->desc.device));
    StreamScope scope_(*mdl, static_cast<cudaStream_t>(stream_));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t es = mdl->elem;

    const void* d_xp = xp;
    void* d_out      = out;
    int64_t d_ld_xp = ld_xp, d_ld_out = ld_out;
    if (mem == UNGAR_B200_MEM_HOST) {
        if (int rc = mdl->ws_xp.reserve(size_t(batch) * n_in * es)) return rc;
        if (int rc = mdl->ws_out.reserve(size_t(batch) * std::max<int64_t>(n_out, 1) * es)) return rc;
        UB_CUDA(cudaMemcpy2DAsync(mdl->ws_xp.ptr, n_in * es, xp, ld_xp * es, n_in * es, batch, cudaMemcpyHostToDevice, stream));
        d_xp = mdl->ws_xp.ptr; d_ld_xp = n_in;
        d_out = mdl->ws_out.ptr; d_ld_out = std::max<int64_t>(n_out, 1);
    }

    int rc = UNGAR_B200_OK;
    if (function == UNGAR_B200_SOFT_INEQUALITIES) {
        rc = mdl->desc.dtype == UNGAR_B200_F64 ? launch_barrier_t<double>(*mdl, d_xp, d_ld_xp, d_out, d_ld_out, batch, want, stream)
                                               : launch_barrier_t<float>(*mdl, d_xp, d_ld_xp, d_out, d_ld_out, batch, want, stream);
    } else {
        // the reference-format calls are served from a DENSE internal record whatever the handle's record format is
        if ((rc = mdl->ws_dense.reserve(size_t(batch) * mdl->dense_size * es))) return rc;
        const int mode = (function == UNGAR_B200_INEQUALITIES && want == WANT_JAC) ? MODE_JH : MODE_PLAIN;
        rc = launch_sweep(*mdl, d_xp, batch, d_ld_xp, mdl->ws_dense.ptr, mdl->dense_size, mode, nullptr, stream, false);
        if (rc) return rc;
        const int32_t* src = want == WANT_Y ? F.d_y_src : want == WANT_JAC ? F.d_jac_src : F.d_hes_src;
        rc = launch_gather(*mdl, mdl->ws_dense.ptr, mdl->dense_size, src, n_out, d_out, d_ld_out, batch, stream);
    }
    if (rc) return rc;
    if (mem == UNGAR_B200_MEM_HOST) {
        if (n_out)
            UB_CUDA(cudaMemcpy2DAsync(out, ld_out * es, d_out, d_ld_out * es, n_out * es, batch, cudaMemcpyDeviceToHost, stream));
        UB_CUDA(cudaStreamSynchronize(stream));
    }
    return UNGAR_B200_OK;
}

}  // namespace

// ---- F32 handles and the consumers of the records ---------------------------------------------------------------------------------
// BASELINE configs 2 and 3 are fp32.  Their sweeps run in fp32; the QP factorisation and the line search do not: the reference computes
// them in double (data_types.hpp:89), the Schur / Riccati pivots span 1e16 (weights of 1e-8 against a 1e-9 regularisation) and the
// acceptance tests of the line search compare relative changes of 1e-6.  An F32 handle therefore keeps a TWIN F64 handle of the same
// problem: arguments are widened on the device, the F64 kernels run, results are narrowed back.  qp_solve consumes the fp32 record as
// it is (fp32 sweep, fp64 factorisation); line_search and sqp_solve evaluate the model in fp64 at the widened iterate.
namespace {

template <class A, class B>
__global__ void convert_kernel(const A* __restrict__ src, long long ld_src, B* __restrict__ dst, long long ld_dst, long long n, long long batch) {
    const long long b = blockIdx.y;
    for (long long b0 = b; b0 < batch; b0 += gridDim.y)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            dst[b0 * ld_dst + i] = static_cast<B>(src[b0 * ld_src + i]);
}

template <class A, class B>
int convert(const A* src, int64_t ld_src, B* dst, int64_t ld_dst, int64_t n, int64_t batch, cudaStream_t stream) {
    if (n <= 0 || batch <= 0) return UNGAR_B200_OK;
    const dim3 grid(unsigned(std::min<int64_t>((n + 255) / 256, 1024)), unsigned(std::min<int64_t>(batch, 65535)));
    convert_kernel<A, B><<<grid, 256, 0, stream>>>(src, ld_src, dst, ld_dst, n, batch);
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

int ensure_twin(ungar_b200_model& m) {
    if (m.twin) return UNGAR_B200_OK;
    ungar_b200_model_desc d = m.desc;
    d.dtype = UNGAR_B200_F64;
    return ungar_b200_model_create(&d, &m.twin);
}

}  // namespace

// =============================================================================================
extern "C" {

int32_t ungar_b200_abi_version(void) { return UNGAR_B200_ABI_VERSION; }
const char* ungar_b200_last_error(void) { return g_last_error.c_str(); }
int64_t ungar_b200_launch_count(void) { return g_launches.load(); }

int ungar_b200_model_create(const ungar_b200_model_desc* desc, ungar_b200_model** out) {
    if (!desc || !out) return fail(UNGAR_B200_EINVAL, "null argument");
    *out = nullptr;
    if (desc->kind < 0 || desc->kind > 2) return fail(UNGAR_B200_EINVAL, "unknown model kind %d", desc->kind);
    if (desc->dtype != UNGAR_B200_F32 && desc->dtype != UNGAR_B200_F64) return fail(UNGAR_B200_EINVAL, "unknown dtype %d", desc->dtype);
    if (desc->horizon < 2 || desc->horizon > 4096) return fail(UNGAR_B200_EINVAL, "horizon %d out of range [2, 4096]", desc->horizon);
    if (!(desc->barrier_stiffness > 0.0) || !(desc->barrier_epsilon > 0.0))
        return fail(UNGAR_B200_EINVAL, "barrier stiffness and epsilon must be positive");
    if (desc->record_format != UNGAR_B200_RECORD_DENSE && desc->record_format != UNGAR_B200_RECORD_COMPACT)
        return fail(UNGAR_B200_EINVAL, "unknown record format %d", desc->record_format);
    if (desc->record_format == UNGAR_B200_RECORD_COMPACT && (desc->kind != UNGAR_B200_QUADRUPED || desc->dtype != UNGAR_B200_F64))
        return fail(UNGAR_B200_EUNSUPPORTED, "compact records exist for the quadruped in F64 only (the other two problems have 2-7x smaller, denser blocks)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(UNGAR_B200_ECUDA, "no usable CUDA device (%s); ungar_b200 has no CPU fallback", cudaGetErrorString(e));
    if (desc->device < 0 || desc->device >= ndev) return fail(UNGAR_B200_EINVAL, "device %d out of range (%d devices)", desc->device, ndev);
    UB_CUDA(cudaSetDevice(desc->device));

    ungar_b200_model* M = new (std::nothrow) ungar_b200_model;
    if (!M) return fail(UNGAR_B200_ENOMEM, "out of host memory");
    M->desc = *desc;
    M->N    = desc->horizon;
    M->elem = desc->dtype == UNGAR_B200_F64 ? 8 : 4;
    switch (desc->kind) {
        case UNGAR_B200_QUADROTOR: make_layout<ub::Quadrotor>(M->N, M->layout); build_tables<ub::Quadrotor>(*M); break;
        case UNGAR_B200_RC_CAR: make_layout<ub::RcCar>(M->N, M->layout); build_tables<ub::RcCar>(*M); break;
        default: make_layout<ub::Quadruped>(M->N, M->layout); build_tables<ub::Quadruped>(*M); break;
    }
    ungar_b200_kkt_layout& L = M->layout;
    if (L.size > 2147483647LL) { delete M; return fail(UNGAR_B200_EINVAL, "record too large"); }
    M->rl = {int(L.g), int(L.A), int(L.C), int(L.h), int(L.cost), int(L.grad), int(L.H), int(L.HN), int(L.Hc), int(L.size)};
    M->dense_size = L.dense_size;
    M->rec_size   = L.dense_size;
    if (desc->record_format == UNGAR_B200_RECORD_COMPACT) {
        M->compact  = true;
        M->rec_size = ub::Compact::size(M->N);
        L.compact   = 1;
        L.size      = M->rec_size;
        M->c2d      = ub::compact_to_dense_map(M->N, ub::DenseOffsets{L.g, L.A, L.C, L.h, L.cost, L.grad, L.H, L.HN});
    }
    {  // soft_inequality_constraint.hpp:133-145
        const double a1 = desc->barrier_stiffness, eps = desc->barrier_epsilon;
        const double b1 = -0.5 * a1 * eps;
        const double c1 = -1.0 / 3.0 * (-b1 - a1 * eps) * eps - 0.5 * a1 * eps * eps - b1 * eps;
        M->bar = {eps, a1, b1, c1, (-b1 - a1 * eps) / (eps * eps), a1, b1, c1};
    }
    for (int f = 0; f < 3; ++f) {
        int rc = upload(M->fn[f].y_src, &M->fn[f].d_y_src);
        if (!rc) rc = upload(M->fn[f].jac_src, &M->fn[f].d_jac_src);
        if (!rc) rc = upload(M->fn[f].hes_src, &M->fn[f].d_hes_src);
        if (rc) { ungar_b200_model_destroy(M); return rc; }
    }
    *out = M;
    return UNGAR_B200_OK;
}

int ungar_b200_model_destroy(ungar_b200_model* model) {
    if (!model) return UNGAR_B200_OK;
    cudaSetDevice(model->desc.device);
    for (auto& f : model->fn) {
        if (f.d_y_src) cudaFree(f.d_y_src);
        if (f.d_jac_src) cudaFree(f.d_jac_src);
        if (f.d_hes_src) cudaFree(f.d_hes_src);
    }
    if (model->d_c2d) cudaFree(model->d_c2d);
    if (model->d_d2c) cudaFree(model->d_d2c);
    if (model->twin) ungar_b200_model_destroy(model->twin);
    if (model->ev_last) cudaEventDestroy(model->ev_last);
    if (model->copy_stream) cudaStreamDestroy(model->copy_stream);
    if (model->ev_entry) cudaEventDestroy(model->ev_entry);
    for (auto& e : model->ev_chunk)
        if (e) cudaEventDestroy(e);
    delete model;
    return UNGAR_B200_OK;
}

int ungar_b200_function_info(const ungar_b200_model* model, int32_t function, int64_t* independent_size,
                             int64_t* parameter_size, int64_t* dependent_size, int64_t* nnz_jacobian,
                             int64_t* nnz_hessian) {
    if (!model) return fail(UNGAR_B200_EINVAL, "null model");
    if (!valid_fn(function)) return fail(UNGAR_B200_EINVAL, "unknown function %d", function);
    const FunctionTables& F = model->fn[function];
    if (independent_size) *independent_size = F.nx;
    if (parameter_size) *parameter_size = F.np;
    if (dependent_size) *dependent_size = F.ny;
    if (nnz_jacobian) *nnz_jacobian = int64_t(F.jac_rows.size());
    if (nnz_hessian) *nnz_hessian = int64_t(F.hes_rows.size());
    return UNGAR_B200_OK;
}

int ungar_b200_jacobian_sparsity(const ungar_b200_model* model, int32_t function, const int64_t** rows,
                                 const int64_t** cols, int64_t* nnz) {
    if (!model || !rows || !cols || !nnz) return fail(UNGAR_B200_EINVAL, "null argument");
    if (!valid_fn(function)) return fail(UNGAR_B200_EINVAL, "unknown function %d", function);
    const FunctionTables& F = model->fn[function];
    *rows = F.jac_rows.data(); *cols = F.jac_cols.data(); *nnz = int64_t(F.jac_rows.size());
    return UNGAR_B200_OK;
}

int ungar_b200_hessian_sparsity(const ungar_b200_model* model, int32_t function, const int64_t** rows,
                                const int64_t** cols, int64_t* nnz) {
    if (!model || !rows || !cols || !nnz) return fail(UNGAR_B200_EINVAL, "null argument");
    if (!valid_fn(function)) return fail(UNGAR_B200_EINVAL, "unknown function %d", function);
    const FunctionTables& F = model->fn[function];
    if (!F.has_hessian)
        return fail(UNGAR_B200_EUNSUPPORTED, "the Hessian is implemented only for scalar functions (function.hpp:136-137)");
    *rows = F.hes_rows.data(); *cols = F.hes_cols.data(); *nnz = int64_t(F.hes_rows.size());
    return UNGAR_B200_OK;
}

int ungar_b200_forward_zero(ungar_b200_model* model, int32_t function, const void* xp, int64_t batch, int64_t ld_xp,
                            void* y, int64_t ld_y, int32_t mem, void* stream) {
    return reference_call(model, function, WANT_Y, xp, batch, ld_xp, y, ld_y, mem, stream);
}

int ungar_b200_sparse_jacobian(ungar_b200_model* model, int32_t function, const void* xp, int64_t batch,
                               int64_t ld_xp, void* vals, int64_t ld_vals, int32_t mem, void* stream) {
    return reference_call(model, function, WANT_JAC, xp, batch, ld_xp, vals, ld_vals, mem, stream);
}

int ungar_b200_sparse_hessian(ungar_b200_model* model, int32_t function, const void* xp, int64_t batch, int64_t ld_xp,
                              void* vals, int64_t ld_vals, int32_t mem, void* stream) {
    return reference_call(model, function, WANT_HES, xp, batch, ld_xp, vals, ld_vals, mem, stream);
}

int ungar_b200_kkt_layout_get(const ungar_b200_model* model, ungar_b200_kkt_layout* out) {
    if (!model || !out) return fail(UNGAR_B200_EINVAL, "null argument");
    *out = model->layout;
    return UNGAR_B200_OK;
}

int ungar_b200_kkt_compact_map(const ungar_b200_model* model, const int32_t** map, int64_t* count) {
    if (!model || !map || !count) return fail(UNGAR_B200_EINVAL, "null argument");
    if (!model->compact) return fail(UNGAR_B200_EUNSUPPORTED, "the handle's records are dense");
    *map = model->c2d.data();
    *count = int64_t(model->c2d.size());
    return UNGAR_B200_OK;
}

static int blocks_call(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp, void* records, int64_t ld_rec,
                       int32_t mem, void* stream_, int mode) {
    if (!model) return fail(UNGAR_B200_EINVAL, "null model");
    if (batch < 0 || (batch > 0 && (!xp || !records))) return fail(UNGAR_B200_EINVAL, "null buffer");
    const ungar_b200_kkt_layout& L = model->layout;
    const int64_t n_in = L.n_dec + L.n_par;
    if (ld_xp < n_in) return fail(UNGAR_B200_EINVAL, "ld_xp %lld < %lld", (long long)ld_xp, (long long)n_in);
    if (ld_rec < L.size) return fail(UNGAR_B200_EINVAL, "ld_rec %lld < record size %lld", (long long)ld_rec, (long long)L.size);
    if (mem != UNGAR_B200_MEM_DEVICE && mem != UNGAR_B200_MEM_HOST) return fail(UNGAR_B200_EINVAL, "unknown mem %d", mem);
    if (batch == 0) return UNGAR_B200_OK;
    UB_CUDA(cudaSetDevice(model->desc.device));
    StreamScope scope_(*model, static_cast<cudaStream_t>(stream_));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t es = model->elem;
    if (mem == UNGAR_B200_MEM_DEVICE) return launch_sweep(*model, xp, batch, ld_xp, records, ld_rec, mode, nullptr, stream, model->compact);

    if (int rc = model->ws_xp.reserve(size_t(batch) * n_in * es)) return rc;
    if (int rc = model->ws_records.reserve(size_t(batch) * L.size * es)) return rc;
    // Chunked pipeline: the record is 10-25x the input, so the D2H transfer dominates; PCIe is full duplex, so the upload and the sweep
    // of chunk c + 1 run while chunk c's record goes back on the copy stream — all but the first chunk's upload and sweep are hidden.
    const int chunks = batch >= 64 * ungar_b200_model::kChunks ? ungar_b200_model::kChunks : 1;
    if (chunks > 1 && !model->copy_stream) {
        UB_CUDA(cudaStreamCreateWithFlags(&model->copy_stream, cudaStreamNonBlocking));
        UB_CUDA(cudaEventCreateWithFlags(&model->ev_entry, cudaEventDisableTiming));
        for (auto& e : model->ev_chunk) UB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    char* const w_xp  = static_cast<char*>(model->ws_xp.ptr);
    char* const w_rec = static_cast<char*>(model->ws_records.ptr);
    const int64_t per = (batch + chunks - 1) / chunks;
    for (int c = 0; c < chunks; ++c) {
        const int64_t b0 = c * per, nb = std::min<int64_t>(per, batch - b0);
        if (nb <= 0) break;
        const char* from = static_cast<const char*>(xp) + size_t(b0) * ld_xp * es;
        char* to         = static_cast<char*>(records) + size_t(b0) * ld_rec * es;
        UB_CUDA(cudaMemcpy2DAsync(w_xp + size_t(b0) * n_in * es, n_in * es, from, ld_xp * es, n_in * es, nb, cudaMemcpyHostToDevice, stream));
        if (int rc = launch_sweep(*model, w_xp + size_t(b0) * n_in * es, nb, n_in, w_rec + size_t(b0) * L.size * es, L.size, mode, nullptr, stream,
                                  model->compact)) return rc;
        cudaStream_t ds = stream;
        if (chunks > 1) {
            UB_CUDA(cudaEventRecord(model->ev_chunk[c], stream));
            UB_CUDA(cudaStreamWaitEvent(model->copy_stream, model->ev_chunk[c], 0));
            ds = model->copy_stream;
        }
        if (ld_rec == L.size) UB_CUDA(cudaMemcpyAsync(to, w_rec + size_t(b0) * L.size * es, size_t(nb) * L.size * es, cudaMemcpyDeviceToHost, ds));
        else UB_CUDA(cudaMemcpy2DAsync(to, ld_rec * es, w_rec + size_t(b0) * L.size * es, L.size * es, L.size * es, nb, cudaMemcpyDeviceToHost, ds));
    }
    if (chunks > 1) {
        UB_CUDA(cudaEventRecord(model->ev_entry, model->copy_stream));
        UB_CUDA(cudaStreamWaitEvent(stream, model->ev_entry, 0));
    }
    UB_CUDA(cudaStreamSynchronize(stream));
    return UNGAR_B200_OK;
}

int ungar_b200_kkt_blocks(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp, void* records,
                          int64_t ld_rec, int32_t mem, void* stream_) {
    return blocks_call(model, xp, batch, ld_xp, records, ld_rec, mem, stream_, MODE_KKT);
}

int ungar_b200_jacobian_blocks(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp, void* records,
                               int64_t ld_rec, int32_t mem, void* stream_) {
    return blocks_call(model, xp, batch, ld_xp, records, ld_rec, mem, stream_, MODE_JAC);
}

int ungar_b200_qp_solve(ungar_b200_model* model, const void* records_device, int64_t batch, int64_t ld_rec, void* steps,
                        int64_t ld_steps, void* multipliers, int64_t ld_multipliers, void* stream_) {
    if (!model) return fail(UNGAR_B200_EINVAL, "null model");
    if (model->desc.dtype != UNGAR_B200_F64) {  // fp32 record in, fp64 factorisation, fp32 step out
        if (batch < 0 || (batch > 0 && (!records_device || !steps))) return fail(UNGAR_B200_EINVAL, "null buffer");
        const ungar_b200_kkt_layout& L = model->layout;
        if (ld_rec < L.size || ld_steps < L.n_dec || (multipliers && ld_multipliers < L.m_eq))
            return fail(UNGAR_B200_EINVAL, "stride smaller than the row it holds");
        if (batch == 0) return UNGAR_B200_OK;
        UB_CUDA(cudaSetDevice(model->desc.device));
        if (int rc = ensure_twin(*model)) return rc;
        cudaStream_t stream = static_cast<cudaStream_t>(stream_);
        StreamScope scope_(*model, stream);
        if (int rc = model->wide_a.reserve(size_t(batch) * L.size * 8)) return rc;
        if (int rc = model->wide_b.reserve(size_t(batch) * L.n_dec * 8)) return rc;
        if (multipliers) { if (int rc = model->wide_c.reserve(size_t(batch) * L.m_eq * 8)) return rc; }
        double* wrec = static_cast<double*>(model->wide_a.ptr);
        double* wstep = static_cast<double*>(model->wide_b.ptr);
        double* wmult = multipliers ? static_cast<double*>(model->wide_c.ptr) : nullptr;
        if (int rc = convert(static_cast<const float*>(records_device), ld_rec, wrec, L.size, L.size, batch, stream)) return rc;
        if (int rc = ungar_b200_qp_solve(model->twin, wrec, batch, L.size, wstep, L.n_dec, wmult, L.m_eq, stream_)) return rc;
        if (int rc = convert(wstep, L.n_dec, static_cast<float*>(steps), ld_steps, L.n_dec, batch, stream)) return rc;
        if (multipliers) return convert(wmult, L.m_eq, static_cast<float*>(multipliers), ld_multipliers, L.m_eq, batch, stream);
        return UNGAR_B200_OK;
    }
    if (batch < 0 || (batch > 0 && (!records_device || !steps))) return fail(UNGAR_B200_EINVAL, "null buffer");
    const ungar_b200_kkt_layout& L = model->layout;
    if (ld_rec < L.size || ld_steps < L.n_dec || (multipliers && ld_multipliers < L.m_eq))
        return fail(UNGAR_B200_EINVAL, "stride smaller than the row it holds");
    if (batch == 0) return UNGAR_B200_OK;
    UB_CUDA(cudaSetDevice(model->desc.device));
    StreamScope scope_(*model, static_cast<cudaStream_t>(stream_));
    return launch_qp(*model, records_device, model->compact, batch, ld_rec, steps, ld_steps, multipliers, ld_multipliers, nullptr,
                     static_cast<cudaStream_t>(stream_));
}

// Test hook (not part of include/ungar_b200.h): copies the first `count` doubles of the QP workspace (per trajectory and group: the
// factor image of qp_twisted.cuh) to the host, so that tests can compare the kernel's intermediate blocks with oracle/qp_reference.py.
int ungar_b200_debug_qp_workspace(ungar_b200_model* model, double* host_out, int64_t count) {
    if (!model || !host_out || count < 0) return fail(UNGAR_B200_EINVAL, "null argument");
    if (size_t(count) * sizeof(double) > model->ws_qp.cap) return fail(UNGAR_B200_EINVAL, "workspace holds fewer than %lld doubles", (long long)count);
    UB_CUDA(cudaSetDevice(model->desc.device));
    UB_CUDA(cudaDeviceSynchronize());
    UB_CUDA(cudaMemcpy(host_out, model->ws_qp.ptr, size_t(count) * sizeof(double), cudaMemcpyDeviceToHost));
    return UNGAR_B200_OK;
}

int ungar_b200_sqp_options_default(ungar_b200_sqp_options* out) {
    if (!out) return fail(UNGAR_B200_EINVAL, "null argument");
    *out = ungar_b200_sqp_options{10, 0, 1.0, 1e-4, 1e-6, 1e-2, 1e-4, 1e-6, 1e-6, 0.5, 1e-6};
    return UNGAR_B200_OK;
}

int ungar_b200_line_search(ungar_b200_model* model, void* xp, int64_t batch, int64_t ld_xp, const void* steps, int64_t ld_steps,
                           const ungar_b200_sqp_options* options, int32_t* status, void* info, void* stream_) {
    if (!model || !options) return fail(UNGAR_B200_EINVAL, "null argument");
    if (batch < 0 || (batch > 0 && (!xp || !steps))) return fail(UNGAR_B200_EINVAL, "null buffer");
    const ungar_b200_kkt_layout& L = model->layout;
    if (ld_xp < L.n_dec + L.n_par || ld_steps < L.n_dec) return fail(UNGAR_B200_EINVAL, "stride smaller than the row it holds");
    if (int rc = check_options(*options)) return rc;
    if (batch == 0) return UNGAR_B200_OK;
    UB_CUDA(cudaSetDevice(model->desc.device));
    if (model->desc.dtype != UNGAR_B200_F64) {  // widen xp and the step, search in fp64, narrow the accepted iterate and the record
        if (int rc = ensure_twin(*model)) return rc;
        cudaStream_t stream = static_cast<cudaStream_t>(stream_);
        StreamScope scope_(*model, stream);
        const int64_t n_in = L.n_dec + L.n_par;
        if (int rc = model->wide_a.reserve(size_t(batch) * n_in * 8)) return rc;
        if (int rc = model->wide_b.reserve(size_t(batch) * L.n_dec * 8)) return rc;
        if (info) { if (int rc = model->wide_d.reserve(size_t(batch) * UNGAR_B200_LINE_SEARCH_INFO_SIZE * 8)) return rc; }
        double* wxp = static_cast<double*>(model->wide_a.ptr);
        double* wdw = static_cast<double*>(model->wide_b.ptr);
        double* winfo = info ? static_cast<double*>(model->wide_d.ptr) : nullptr;
        if (int rc = convert(static_cast<const float*>(xp), ld_xp, wxp, n_in, n_in, batch, stream)) return rc;
        if (int rc = convert(static_cast<const float*>(steps), ld_steps, wdw, L.n_dec, L.n_dec, batch, stream)) return rc;
        if (int rc = ungar_b200_line_search(model->twin, wxp, batch, n_in, wdw, L.n_dec, options, status, winfo, stream_)) return rc;
        if (int rc = convert(wxp, n_in, static_cast<float*>(xp), ld_xp, L.n_dec, batch, stream)) return rc;
        if (info) return convert(winfo, UNGAR_B200_LINE_SEARCH_INFO_SIZE, static_cast<float*>(info), UNGAR_B200_LINE_SEARCH_INFO_SIZE,
                                 UNGAR_B200_LINE_SEARCH_INFO_SIZE, batch, stream);
        return UNGAR_B200_OK;
    }
    StreamScope scope_(*model, static_cast<cudaStream_t>(stream_));
    return launch_line_search(*model, static_cast<double*>(xp), batch, ld_xp, static_cast<const double*>(steps), ld_steps, *options,
                              status, static_cast<double*>(info), static_cast<cudaStream_t>(stream_));
}

int ungar_b200_sqp_solve(ungar_b200_model* model, void* xp, int64_t batch, int64_t ld_xp, const ungar_b200_sqp_options* options,
                         int32_t* status, void* info, int32_t mem, void* stream_) {
    if (!model || !options) return fail(UNGAR_B200_EINVAL, "null argument");
    if (batch < 0 || (batch > 0 && (!xp || !status))) return fail(UNGAR_B200_EINVAL, "null buffer");
    if (mem != UNGAR_B200_MEM_DEVICE && mem != UNGAR_B200_MEM_HOST) return fail(UNGAR_B200_EINVAL, "unknown mem %d", mem);
    const ungar_b200_kkt_layout& L = model->layout;
    const int64_t n_in = L.n_dec + L.n_par;
    if (ld_xp < n_in) return fail(UNGAR_B200_EINVAL, "ld_xp %lld < %lld", (long long)ld_xp, (long long)n_in);
    if (int rc = check_options(*options)) return rc;
    if (batch == 0) return UNGAR_B200_OK;
    UB_CUDA(cudaSetDevice(model->desc.device));
    if (model->desc.dtype != UNGAR_B200_F64) {  // the whole loop in fp64 on the twin, the iterate widened / narrowed on the device
        if (int rc = ensure_twin(*model)) return rc;
        cudaStream_t stream = static_cast<cudaStream_t>(stream_);
        StreamScope scope_(*model, stream);
        if (int rc = model->wide_a.reserve(size_t(batch) * n_in * 8)) return rc;
        if (info) { if (int rc = model->wide_d.reserve(size_t(batch) * UNGAR_B200_LINE_SEARCH_INFO_SIZE * 8)) return rc; }
        double* wxp = static_cast<double*>(model->wide_a.ptr);
        double* winfo = info ? static_cast<double*>(model->wide_d.ptr) : nullptr;
        float* d_xp32 = static_cast<float*>(xp);
        int64_t d_ld = ld_xp;
        int32_t* d_status = status;
        if (mem == UNGAR_B200_MEM_HOST) {
            if (int rc = model->ws_xp.reserve(size_t(batch) * n_in * 4)) return rc;
            if (int rc = model->ws_status.reserve(size_t(batch) * 2 * sizeof(int32_t))) return rc;
            UB_CUDA(cudaMemcpy2DAsync(model->ws_xp.ptr, n_in * 4, xp, ld_xp * 4, n_in * 4, batch, cudaMemcpyHostToDevice, stream));
            d_xp32 = static_cast<float*>(model->ws_xp.ptr); d_ld = n_in;
            d_status = static_cast<int32_t*>(model->ws_status.ptr);
        }
        if (int rc = convert(d_xp32, d_ld, wxp, n_in, n_in, batch, stream)) return rc;
        if (int rc = ungar_b200_sqp_solve(model->twin, wxp, batch, n_in, options, d_status, winfo, UNGAR_B200_MEM_DEVICE, stream_)) return rc;
        if (int rc = convert(wxp, n_in, d_xp32, d_ld, L.n_dec, batch, stream)) return rc;
        if (mem == UNGAR_B200_MEM_HOST) {
            if (info) {
                if (int rc = model->ws_info.reserve(size_t(batch) * UNGAR_B200_LINE_SEARCH_INFO_SIZE * 4)) return rc;
                if (int rc = convert(winfo, UNGAR_B200_LINE_SEARCH_INFO_SIZE, static_cast<float*>(model->ws_info.ptr), UNGAR_B200_LINE_SEARCH_INFO_SIZE,
                                     UNGAR_B200_LINE_SEARCH_INFO_SIZE, batch, stream)) return rc;
                UB_CUDA(cudaMemcpyAsync(info, model->ws_info.ptr, size_t(batch) * UNGAR_B200_LINE_SEARCH_INFO_SIZE * 4, cudaMemcpyDeviceToHost, stream));
            }
            UB_CUDA(cudaMemcpy2DAsync(xp, ld_xp * 4, d_xp32, n_in * 4, L.n_dec * 4, batch, cudaMemcpyDeviceToHost, stream));
            UB_CUDA(cudaMemcpyAsync(status, d_status, size_t(batch) * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            UB_CUDA(cudaStreamSynchronize(stream));
            return UNGAR_B200_OK;
        }
        if (info) return convert(winfo, UNGAR_B200_LINE_SEARCH_INFO_SIZE, static_cast<float*>(info), UNGAR_B200_LINE_SEARCH_INFO_SIZE,
                                 UNGAR_B200_LINE_SEARCH_INFO_SIZE, batch, stream);
        return UNGAR_B200_OK;
    }
    StreamScope scope_(*model, static_cast<cudaStream_t>(stream_));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t es = sizeof(double);
    // the records never leave the device here: the quadruped loop runs on the compact format whatever the handle's format is
    const bool crec = model->desc.kind == UNGAR_B200_QUADRUPED;
    const int64_t rsize = crec ? ub::Compact::size(model->N) : model->dense_size;
    if (int rc = model->ws_records.reserve(size_t(batch) * rsize * es)) return rc;
    if (int rc = model->ws_steps.reserve(size_t(batch) * L.n_dec * es)) return rc;
    double* d_xp    = static_cast<double*>(xp);
    int64_t d_ld_xp = ld_xp;
    int32_t* d_status = status;
    double* d_info    = static_cast<double*>(info);
    if (mem == UNGAR_B200_MEM_HOST) {
        if (int rc = model->ws_xp.reserve(size_t(batch) * n_in * es)) return rc;
        if (int rc = model->ws_status.reserve(size_t(batch) * 2 * sizeof(int32_t))) return rc;
        if (info) {
            if (int rc = model->ws_info.reserve(size_t(batch) * UNGAR_B200_LINE_SEARCH_INFO_SIZE * es)) return rc;
        }
        if (ld_xp == n_in) UB_CUDA(cudaMemcpyAsync(model->ws_xp.ptr, xp, size_t(batch) * n_in * es, cudaMemcpyHostToDevice, stream));  // one linear transfer
        else UB_CUDA(cudaMemcpy2DAsync(model->ws_xp.ptr, n_in * es, xp, ld_xp * es, n_in * es, batch, cudaMemcpyHostToDevice, stream));
        d_xp = static_cast<double*>(model->ws_xp.ptr); d_ld_xp = n_in;
        d_status = static_cast<int32_t*>(model->ws_status.ptr);
        d_info   = info ? static_cast<double*>(model->ws_info.ptr) : nullptr;
    }
    UB_CUDA(cudaMemsetAsync(d_status, 0, size_t(batch) * 2 * sizeof(int32_t), stream));
    if (d_info) UB_CUDA(cudaMemsetAsync(d_info, 0, size_t(batch) * UNGAR_B200_LINE_SEARCH_INFO_SIZE * es, stream));
    // Trajectories that stopped (soft_sqp.hpp:100-108 breaks out per problem) are skipped by every kernel: the QP solve and the line search
    // read the status, the sweep takes the list of RUNNING trajectories rebuilt after each line search (no host round trip: the list and
    // its length stay on the device).  Only the sweeps that take a work list use it (quadruped compact, small-team); others sweep all.
    if (int rc = model->ws_active.reserve(size_t(batch) * sizeof(int) + 16)) return rc;
    int* const d_active = static_cast<int*>(model->ws_active.ptr) + 4;
    unsigned int* const d_nact = static_cast<unsigned int*>(model->ws_active.ptr);
    struct ActiveReset { ungar_b200_model* m; ~ActiveReset() { m->active = nullptr; m->n_active = nullptr; } } active_reset{model};
    for (int it = 0; it < options->max_iterations; ++it) {
        // AssembleOSQPInstance -> Solve -> BacktrackingLineSearch::Do (soft_sqp.hpp:76-99), stream-ordered
        if (it > 0) {
            UB_CUDA(cudaMemsetAsync(d_nact, 0, sizeof(unsigned int), stream));
            ub::build_active_kernel<<<unsigned((batch + 255) / 256), 256, 0, stream>>>(d_status, batch, d_active, d_nact);
            ++g_launches;
            model->active = d_active;
            model->n_active = d_nact;
        }
        if (int rc = launch_sweep(*model, d_xp, batch, d_ld_xp, model->ws_records.ptr, rsize, MODE_KKT, nullptr, stream, crec)) return rc;
        if (int rc = launch_qp(*model, model->ws_records.ptr, crec, batch, rsize, model->ws_steps.ptr, L.n_dec, nullptr, 0, d_status, stream)) return rc;
        if (int rc = launch_line_search(*model, d_xp, batch, d_ld_xp, static_cast<const double*>(model->ws_steps.ptr), L.n_dec, *options,
                                        d_status, d_info, stream)) return rc;
    }
    if (mem == UNGAR_B200_MEM_HOST) {
        UB_CUDA(cudaMemcpy2DAsync(xp, ld_xp * es, d_xp, n_in * es, L.n_dec * es, batch, cudaMemcpyDeviceToHost, stream));
        UB_CUDA(cudaMemcpyAsync(status, d_status, size_t(batch) * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        if (info) UB_CUDA(cudaMemcpyAsync(info, d_info, size_t(batch) * UNGAR_B200_LINE_SEARCH_INFO_SIZE * es, cudaMemcpyDeviceToHost, stream));
        UB_CUDA(cudaStreamSynchronize(stream));
    }
    return UNGAR_B200_OK;
}

int ungar_b200_set_profiling(int32_t enabled) {
    g_ring.enabled = enabled != 0;
    g_ring.head = g_ring.count = 0;
    return UNGAR_B200_OK;
}

int ungar_b200_sweep_times(float* ms, int32_t cap, int32_t* count) {
    if (!ms || !count || cap < 0) return fail(UNGAR_B200_EINVAL, "null argument");
    const int n = g_ring.count < cap ? g_ring.count : cap;
    for (int i = 0; i < n; ++i) {
        const int slot = ((g_ring.head - n + i) % kRing + kRing) % kRing;
        UB_CUDA(cudaEventSynchronize(g_ring.stop[slot]));
        UB_CUDA(cudaEventElapsedTime(&ms[i], g_ring.start[slot], g_ring.stop[slot]));
    }
    *count = n;
    g_ring.head = g_ring.count = 0;
    return UNGAR_B200_OK;
}

// One outer-iteration step.  `width` leading scalars of every row of `src` are copied into columns [0, width) of the device workspace
// `w_xp` (rows of n_in scalars) — the whole flat vector, or only the decision variables when the parameter block is cached —
// then the sweep runs on the workspace.  Host sources: chunked H2D on a copy stream overlapped with the sweep of the previous chunk.
static int step_impl(ungar_b200_model* model, const void* src, int64_t batch, int64_t ld_src, int64_t width, char* w_xp, void* records_device,
                     int64_t ld_rec, void* summaries, int32_t mem, cudaStream_t stream) {
    const ungar_b200_kkt_layout& L = model->layout;
    const int64_t n_in = L.n_dec + L.n_par;
    const size_t es = model->elem;
    if (mem == UNGAR_B200_MEM_DEVICE) {
        UB_CUDA(cudaMemcpy2DAsync(w_xp, n_in * es, src, ld_src * es, width * es, batch, cudaMemcpyDeviceToDevice, stream));
        return launch_sweep(*model, w_xp, batch, n_in, records_device, ld_rec, MODE_KKT, summaries, stream, model->compact);
    }
    if (int rc = model->ws_out.reserve(size_t(batch) * UNGAR_B200_SUMMARY_SIZE * es)) return rc;
    char* const w_sum = static_cast<char*>(model->ws_out.ptr);
    // Chunked pipeline: the inputs are 4 % of the traffic of the sweep but PCIe is ~100x slower than HBM, so the transfer
    // dominates; copying chunk c + 1 while chunk c is swept hides all but the last chunk's kernel time.
    const int chunks = batch >= 64 * ungar_b200_model::kChunks ? ungar_b200_model::kChunks : 1;
    if (chunks > 1 && !model->copy_stream) {
        UB_CUDA(cudaStreamCreateWithFlags(&model->copy_stream, cudaStreamNonBlocking));
        UB_CUDA(cudaEventCreateWithFlags(&model->ev_entry, cudaEventDisableTiming));
        for (auto& e : model->ev_chunk) UB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (chunks > 1) {  // the copies start after whatever the caller already queued on `stream`
        UB_CUDA(cudaEventRecord(model->ev_entry, stream));
        UB_CUDA(cudaStreamWaitEvent(model->copy_stream, model->ev_entry, 0));
    }
    const int64_t per = (batch + chunks - 1) / chunks;
    char* stage = nullptr;
    {
        const char* flat = getenv("UNGAR_B200_H2D_PITCHED");  // measurement switch: the former pitched H2D copy
        // A pitched H2D copy pays a DMA descriptor per row: rows of 1-2 KB (quadrotor, RC car) crawl (0.18 / 0.37 G nodes/s against
        // 0.40 / 1.05 staged), and even the quadruped's 29.9 KB rows swing between 26 and 44 GB/s from box to box and run to run while a
        // linear transfer of the same bytes holds 42-49 GB/s in the same process.
        if (width < n_in && ld_src == width && !(flat && flat[0] == '1')) {
            if (int rc = model->ws_stage.reserve(size_t(batch) * width * es)) return rc;
            stage = static_cast<char*>(model->ws_stage.ptr);
        }
    }
    for (int c = 0; c < chunks; ++c) {
        const int64_t b0 = c * per, nb = std::min<int64_t>(per, batch - b0);
        if (nb <= 0) break;
        cudaStream_t cs = chunks > 1 ? model->copy_stream : stream;
        const char* from = static_cast<const char*>(src) + size_t(b0) * ld_src * es;
        if (ld_src == n_in && width == n_in) {
            UB_CUDA(cudaMemcpyAsync(w_xp + size_t(b0) * n_in * es, from, size_t(nb) * n_in * es, cudaMemcpyHostToDevice, cs));
        } else if (ld_src == width && stage) {
            // contiguous host rows narrower than the device rows (decision variables only): ONE linear transfer over PCIe into a
            // landing zone on the copy stream; the rows are scattered into place on the COMPUTE stream (below, at HBM speed), so the
            // copy stream goes straight on to the next chunk
            UB_CUDA(cudaMemcpyAsync(stage + size_t(b0) * width * es, from, size_t(nb) * width * es, cudaMemcpyHostToDevice, cs));
        } else {
            UB_CUDA(cudaMemcpy2DAsync(w_xp + size_t(b0) * n_in * es, n_in * es, from, ld_src * es, width * es, nb, cudaMemcpyHostToDevice, cs));
        }
        if (chunks > 1) {
            UB_CUDA(cudaEventRecord(model->ev_chunk[c], cs));
            UB_CUDA(cudaStreamWaitEvent(stream, model->ev_chunk[c], 0));
        }
        if (ld_src == width && stage && !(ld_src == n_in && width == n_in)) {
            const int rc = es == 8 ? convert(reinterpret_cast<const double*>(stage + size_t(b0) * width * es), width,
                                             reinterpret_cast<double*>(w_xp + size_t(b0) * n_in * es), n_in, width, nb, stream)
                                   : convert(reinterpret_cast<const float*>(stage + size_t(b0) * width * es), width,
                                             reinterpret_cast<float*>(w_xp + size_t(b0) * n_in * es), n_in, width, nb, stream);
            if (rc) return rc;
        }
        if (int rc = launch_sweep(*model, w_xp + size_t(b0) * n_in * es, nb, n_in, static_cast<char*>(records_device) + size_t(b0) * ld_rec * es,
                                  ld_rec, MODE_KKT, w_sum + size_t(b0) * UNGAR_B200_SUMMARY_SIZE * es, stream, model->compact)) return rc;
    }
    UB_CUDA(cudaMemcpyAsync(summaries, w_sum, size_t(batch) * UNGAR_B200_SUMMARY_SIZE * es, cudaMemcpyDeviceToHost, stream));
    UB_CUDA(cudaStreamSynchronize(stream));
    return UNGAR_B200_OK;
}

static int step_check(ungar_b200_model* model, const void* in, int64_t batch, int64_t ld_in, int64_t need, void*& records_device, int64_t& ld_rec,
                      void* summaries, int32_t mem) {
    if (!model) return fail(UNGAR_B200_EINVAL, "null model");
    if (batch < 0 || (batch > 0 && (!in || !summaries))) return fail(UNGAR_B200_EINVAL, "null buffer");
    if (ld_in < need) return fail(UNGAR_B200_EINVAL, "input stride %lld < %lld", (long long)ld_in, (long long)need);
    if (mem != UNGAR_B200_MEM_DEVICE && mem != UNGAR_B200_MEM_HOST) return fail(UNGAR_B200_EINVAL, "unknown mem %d", mem);
    if (batch == 0) return UNGAR_B200_OK;
    UB_CUDA(cudaSetDevice(model->desc.device));
    const ungar_b200_kkt_layout& L = model->layout;
    if (!records_device) {
        if (int rc = model->ws_records.reserve(size_t(batch) * L.size * model->elem)) return rc;
        records_device = model->ws_records.ptr;
        ld_rec = L.size;
    } else if (ld_rec < L.size) {
        return fail(UNGAR_B200_EINVAL, "ld_rec %lld < record size %lld", (long long)ld_rec, (long long)L.size);
    }
    return UNGAR_B200_OK;
}

int ungar_b200_kkt_step(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp, void* records_device,
                        int64_t ld_rec, void* summaries, int32_t mem, void* stream_) {
    const int64_t n_in = model ? model->layout.n_dec + model->layout.n_par : 0;
    if (int rc = step_check(model, xp, batch, ld_xp, n_in, records_device, ld_rec, summaries, mem)) return rc;
    if (batch == 0) return UNGAR_B200_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    StreamScope scope_(*model, stream);
    if (mem == UNGAR_B200_MEM_DEVICE)  // device inputs are swept in place: no staging copy
        return launch_sweep(*model, xp, batch, ld_xp, records_device, ld_rec, MODE_KKT, summaries, stream, model->compact);
    if (int rc = model->ws_xp.reserve(size_t(batch) * n_in * model->elem)) return rc;
    return step_impl(model, xp, batch, ld_xp, n_in, static_cast<char*>(model->ws_xp.ptr), records_device, ld_rec, summaries, mem, stream);
}

int ungar_b200_set_parameters(ungar_b200_model* model, const void* parameters, int64_t batch, int64_t ld_par, int32_t mem, void* stream_) {
    if (!model) return fail(UNGAR_B200_EINVAL, "null model");
    const ungar_b200_kkt_layout& L = model->layout;
    if (batch < 0 || (batch > 0 && !parameters)) return fail(UNGAR_B200_EINVAL, "null buffer");
    if (ld_par < L.n_par) return fail(UNGAR_B200_EINVAL, "ld_par %lld < %lld", (long long)ld_par, (long long)L.n_par);
    if (mem != UNGAR_B200_MEM_DEVICE && mem != UNGAR_B200_MEM_HOST) return fail(UNGAR_B200_EINVAL, "unknown mem %d", mem);
    UB_CUDA(cudaSetDevice(model->desc.device));
    StreamScope scope_(*model, static_cast<cudaStream_t>(stream_));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t es = model->elem;
    const int64_t n_in = L.n_dec + L.n_par;
    model->cached_batch = 0;
    if (batch == 0) return UNGAR_B200_OK;
    if (int rc = model->ws_xp_cached.reserve(size_t(batch) * n_in * es)) return rc;
    UB_CUDA(cudaMemcpy2DAsync(static_cast<char*>(model->ws_xp_cached.ptr) + L.n_dec * es, n_in * es, parameters, ld_par * es, L.n_par * es, batch,
                              mem == UNGAR_B200_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, stream));
    if (mem == UNGAR_B200_MEM_HOST) UB_CUDA(cudaStreamSynchronize(stream));
    model->cached_batch = batch;
    return UNGAR_B200_OK;
}

int ungar_b200_kkt_step_x(ungar_b200_model* model, const void* x, int64_t batch, int64_t ld_x, void* records_device, int64_t ld_rec,
                          void* summaries, int32_t mem, void* stream_) {
    const int64_t n_dec = model ? model->layout.n_dec : 0;
    if (int rc = step_check(model, x, batch, ld_x, n_dec, records_device, ld_rec, summaries, mem)) return rc;
    if (batch == 0) return UNGAR_B200_OK;
    StreamScope scope_(*model, static_cast<cudaStream_t>(stream_));
    if (batch != model->cached_batch)
        return fail(UNGAR_B200_EINVAL, "ungar_b200_set_parameters holds the parameters of %lld trajectories, this call has %lld",
                    (long long)model->cached_batch, (long long)batch);
    return step_impl(model, x, batch, ld_x, n_dec, static_cast<char*>(model->ws_xp_cached.ptr), records_device, ld_rec, summaries, mem,
                     static_cast<cudaStream_t>(stream_));
}

int ungar_b200_summaries(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp, const void* records,
                         int64_t ld_rec, void* summaries, void* stream_) {
    if (!model || (batch > 0 && (!xp || !records || !summaries))) return fail(UNGAR_B200_EINVAL, "null argument");
    if (batch <= 0) return batch == 0 ? UNGAR_B200_OK : fail(UNGAR_B200_EINVAL, "negative batch");
    UB_CUDA(cudaSetDevice(model->desc.device));
    StreamScope scope_(*model, static_cast<cudaStream_t>(stream_));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const ungar_b200_kkt_layout& L = model->layout;
    const int u0 = int(L.nx * (L.horizon + 1));
    if (model->compact)
        ub::summary_compact_kernel<<<unsigned(batch), 32, 0, stream>>>(static_cast<const double*>(xp), ld_xp, static_cast<const double*>(records), ld_rec,
                                                                        static_cast<double*>(summaries), model->N, u0);
    else if (model->desc.dtype == UNGAR_B200_F64)
        ub::summary_kernel<double><<<unsigned(batch), 32, 0, stream>>>(static_cast<const double*>(xp), ld_xp,
            static_cast<const double*>(records), ld_rec, static_cast<double*>(summaries), model->rl, u0, int(L.nu), int(L.m_eq), int(L.m_ineq));
    else
        ub::summary_kernel<float><<<unsigned(batch), 32, 0, stream>>>(static_cast<const float*>(xp), ld_xp,
            static_cast<const float*>(records), ld_rec, static_cast<float*>(summaries), model->rl, u0, int(L.nu), int(L.m_eq), int(L.m_ineq));
    ++g_launches;
    UB_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

}  // extern "C"
