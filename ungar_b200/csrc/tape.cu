// Generic tape path of the C ABI (include/ungar_b200.h, "Generic path"): host-side analysis of a recorded tape and the
// launches of the register-machine kernels (tape_machine.cuh).  The only host arithmetic here is structural (liveness,
// dependency sets, colouring); every value and derivative is computed on the device.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <dlfcn.h>
#include <unistd.h>
#include <fstream>
#include <sstream>
#include <sys/stat.h>
#include <array>
#include <atomic>
#include <thread>
#include <type_traits>
#include <cuda.h>
#include <nvrtc.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <set>
#include <string>
#include <vector>

#include "../../include/ungar_b200.h"
#include "abi_internal.h"
#include "tape_machine.cuh"

namespace {

using ub::tape::Instr;
static_assert(int(UNGAR_B200_OP_CEQ) == int(ub::tape::T_CEQ) && int(UNGAR_B200_OP_COUNT) == int(ub::tape::T_OUTPUT) &&
                  int(UNGAR_B200_OP_SQRT) == int(ub::tape::T_SQRT),
              "the ABI op codes are the machine's op codes");

int tfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return ub_set_error(code, buf);
}

#define UBT_CUDA(call)                                                                                                  \
    do {                                                                                                                \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess) return tfail(UNGAR_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {
    void* ptr  = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return UNGAR_B200_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&ptr, bytes);
        if (e != cudaSuccess) return tfail(UNGAR_B200_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return UNGAR_B200_OK;
    }
    template <class T>
    int upload(const std::vector<T>& host) {
        if (int rc = reserve(std::max<size_t>(host.size(), 1) * sizeof(T))) return rc;
        if (!host.empty()) {
            cudaError_t e = cudaMemcpy(ptr, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return tfail(UNGAR_B200_ECUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
        }
        return UNGAR_B200_OK;
    }
    ~DevBuf() {
        if (ptr) cudaFree(ptr);
    }
};

bool is_unary(int op) {
    return op == UNGAR_B200_OP_NEG || (op >= UNGAR_B200_OP_SQRT && op <= UNGAR_B200_OP_ABS);
}
bool is_binary(int op) {
    return (op >= UNGAR_B200_OP_ADD && op <= UNGAR_B200_OP_DIV) || op == UNGAR_B200_OP_POW || op == UNGAR_B200_OP_ATAN2;
}
bool is_cond(int op) { return op >= UNGAR_B200_OP_CLT && op <= UNGAR_B200_OP_CEQ; }
// d2/d(operands)2 != 0: the operation makes its operands interact in the Hessian.
bool is_nonlinear_unary(int op) { return is_unary(op) && op != UNGAR_B200_OP_NEG && op != UNGAR_B200_OP_ABS; }

using IndexSet = std::vector<int32_t>;  // sorted, unique
IndexSet set_union(const IndexSet& a, const IndexSet& b) {
    IndexSet r;
    r.reserve(a.size() + b.size());
    std::set_union(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r));
    return r;
}

}  // namespace

struct ungar_b200_tape {
    int device = 0;
    int64_t n_indep = 0, n_dep = 0;
    std::vector<ungar_b200_tape_node> nodes;
    std::vector<int32_t> dep_id;
    std::vector<double> dep_const;
    std::vector<char> live;  // node reaches a dependent
    int64_t n_live = 0;

    // program
    std::vector<Instr> code;
    std::vector<double> consts;
    int n_slots = 0;

    // structural patterns (lazy)
    bool jp_done = false, hp_done = false;
    std::vector<int64_t> jp_rows, jp_cols, hp_rows, hp_cols;

    // selected elements
    bool j_set = false, h_set = false;
    std::vector<int64_t> j_rows, j_cols, h_rows, h_cols;
    std::vector<int> color;     // [n_indep], -1: no selected element in that column
    int n_colors = 0;
    std::vector<int> jac_slot;  // [n_dep * n_colors]
    std::vector<int> h_pi, h_pj;               // Hessian directions
    std::vector<int> h_di, h_dj, h_pr;         // per element: direction of e_i, of e_j, of e_i + e_j (-1 on the diagonal)

    // NVRTC-specialised straight-line kernels, one per ORDER (0 values, 1 Jacobian, 2 Hessian jets): tried once each
    struct Special {
        // 0: not tried, 1: ready, -1: unavailable (too large, NVRTC missing, compile error: the interpreter serves), 2: a worker thread is
        // compiling the segmented kernels of a long tape (tens of seconds) while the interpreter keeps serving the calls
        std::atomic<int> state{0};
        std::thread worker;
        CUmodule module = nullptr;
        CUfunction fn = nullptr;
        std::vector<CUfunction> parts;  // segmented kernels of a tape beyond kSpecializeMax (run back to back, values cross through the scratch)
        uint64_t key = 0;
        bool from_cache = false;
        double compile_seconds = 0.0;
    } special[4];  // orders 0, 1, 2 and [3]: the reverse sweep of a scalar function (gradient in one pass)
    int64_t calls[4] = {0, 0, 0, 0};
    std::vector<int> elem_of_indep;  // scalar functions: selected Jacobian element of independent i, or -1
    bool elem_uploaded = false;

    // device state
    bool uploaded = false, j_uploaded = false, h_uploaded = false;
    DevBuf d_elem, d_code, d_consts, d_color, d_jac_slot, d_pi, d_pj, d_di, d_dj, d_pr, d_w, scratch, ws_x, ws_out, ws_q;
};

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// Analysis
// ---------------------------------------------------------------------------------------------------------------------
int validate(const ungar_b200_tape& T) {
    const int64_t n = int64_t(T.nodes.size());
    for (int64_t i = 0; i < n; ++i) {
        const ungar_b200_tape_node& nd = T.nodes[i];
        auto ok = [&](int32_t r) { return r >= 0 && r < i; };
        if (nd.op >= UNGAR_B200_OP_COUNT) return tfail(UNGAR_B200_EINVAL, "node %lld: unknown op %d", (long long)i, int(nd.op));
        if (nd.op == UNGAR_B200_OP_INDEP) {
            if (nd.a < 0 || nd.a >= T.n_indep) return tfail(UNGAR_B200_EINVAL, "node %lld: independent %d out of range", (long long)i, nd.a);
        } else if (nd.op == UNGAR_B200_OP_CONST) {
        } else if (is_unary(nd.op)) {
            if (!ok(nd.a)) return tfail(UNGAR_B200_EINVAL, "node %lld: operand out of range", (long long)i);
        } else if (is_binary(nd.op)) {
            if (!ok(nd.a) || !ok(nd.b)) return tfail(UNGAR_B200_EINVAL, "node %lld: operand out of range", (long long)i);
        } else if (!ok(nd.a) || !ok(nd.b) || !ok(nd.c) || !ok(nd.d)) {
            return tfail(UNGAR_B200_EINVAL, "node %lld: operand out of range", (long long)i);
        }
    }
    for (int64_t r = 0; r < T.n_dep; ++r)
        if (T.dep_id[r] >= n) return tfail(UNGAR_B200_EINVAL, "dependent %lld: node %d out of range", (long long)r, T.dep_id[r]);
    return UNGAR_B200_OK;
}

template <class F>
void for_operands(const ungar_b200_tape_node& nd, F&& f) {
    if (nd.op == UNGAR_B200_OP_INDEP || nd.op == UNGAR_B200_OP_CONST) return;
    f(nd.a);
    if (is_unary(nd.op)) return;
    f(nd.b);
    if (is_binary(nd.op)) return;
    f(nd.c);
    f(nd.d);
}

// Dead-node elimination, slot allocation by liveness (free list), instruction stream with the dependents written out right
// after the node that defines them.
void build_program(ungar_b200_tape& T) {
    const int64_t n = int64_t(T.nodes.size());
    T.live.assign(size_t(n), 0);
    for (int32_t id : T.dep_id)
        if (id >= 0) T.live[size_t(id)] = 1;
    for (int64_t i = n - 1; i >= 0; --i)
        if (T.live[size_t(i)]) for_operands(T.nodes[size_t(i)], [&](int32_t o) { T.live[size_t(o)] = 1; });
    std::vector<int32_t> uses(size_t(n), 0);
    for (int64_t i = 0; i < n; ++i)
        if (T.live[size_t(i)]) for_operands(T.nodes[size_t(i)], [&](int32_t o) { ++uses[size_t(o)]; });
    std::multimap<int32_t, int64_t> outputs_of;  // node -> dependents it defines
    for (int64_t r = 0; r < T.n_dep; ++r)
        if (T.dep_id[size_t(r)] >= 0) { outputs_of.emplace(T.dep_id[size_t(r)], r); ++uses[size_t(T.dep_id[size_t(r)])]; }

    std::vector<int32_t> slot(size_t(n), -1), free_slots;
    int next_slot = 0;
    T.code.clear();
    T.consts.clear();
    T.n_live = 0;
    auto release = [&](int32_t o) {
        if (--uses[size_t(o)] == 0) free_slots.push_back(slot[size_t(o)]);
    };
    auto take_slot = [&]() {
        if (!free_slots.empty()) { const int32_t s = free_slots.back(); free_slots.pop_back(); return s; }
        return int32_t(next_slot++);
    };
    // Independents are loaded right before their first use, not where Independent() recorded them (all at the start): otherwise every
    // one of them would hold a slot from the beginning and the scratch would be as large as the input vector.
    auto materialize = [&](int32_t o) {
        if (slot[size_t(o)] >= 0) return;
        slot[size_t(o)] = take_slot();
        T.code.push_back(Instr{ub::tape::T_INDEP, slot[size_t(o)], T.nodes[size_t(o)].a, 0, 0, 0});
    };
    auto emit_outputs = [&](int64_t i) {
        auto range = outputs_of.equal_range(int32_t(i));
        for (auto it = range.first; it != range.second; ++it) {
            T.code.push_back(Instr{ub::tape::T_OUTPUT, 0, slot[size_t(i)], int(it->second), 0, 0});
            release(int32_t(i));
        }
    };
    for (int64_t i = 0; i < n; ++i) {
        if (!T.live[size_t(i)]) continue;
        ++T.n_live;
        const ungar_b200_tape_node& nd = T.nodes[size_t(i)];
        if (nd.op == UNGAR_B200_OP_INDEP) {
            if (outputs_of.count(int32_t(i))) { materialize(int32_t(i)); emit_outputs(i); }  // a dependent that IS an independent
            continue;
        }
        Instr in{int(nd.op), 0, 0, 0, 0, 0};
        if (nd.op == UNGAR_B200_OP_CONST) { in.a = int(T.consts.size()); T.consts.push_back(nd.k); }
        else {
            for_operands(nd, [&](int32_t o) { materialize(o); });
            in.a = slot[size_t(nd.a)];
            if (!is_unary(nd.op)) in.b = slot[size_t(nd.b)];
            if (is_cond(nd.op)) { in.c = slot[size_t(nd.c)]; in.d = slot[size_t(nd.d)]; }
            if (nd.op == UNGAR_B200_OP_POW) {  // constant exponent: closed-form derivatives (finite for a <= 0), see jet_pow_const
                in.c = -1;
                if (T.nodes[size_t(nd.b)].op == UNGAR_B200_OP_CONST) { in.c = int(T.consts.size()); T.consts.push_back(T.nodes[size_t(nd.b)].k); }
            }
        }
        // operands may be released before the destination is chosen: every instruction reads all operands before it writes
        for_operands(nd, [&](int32_t o) { release(o); });
        slot[size_t(i)] = take_slot();
        in.dst = slot[size_t(i)];
        T.code.push_back(in);
        emit_outputs(i);
    }
    for (int64_t r = 0; r < T.n_dep; ++r)
        if (T.dep_id[size_t(r)] < 0) {
            T.code.push_back(Instr{ub::tape::T_OUTPUT_CONST, 0, int(T.consts.size()), int(r), 0, 0});
            T.consts.push_back(T.dep_const[size_t(r)]);
        }
    T.n_slots = std::max(next_slot, 1);
}

// Dependency sets of the live nodes (freed at last use); calls visit(i, node, sets) for every live node after its set is built.
template <class Visit>
void propagate_sets(const ungar_b200_tape& T, Visit&& visit) {
    const int64_t n = int64_t(T.nodes.size());
    std::vector<int32_t> uses(size_t(n), 0);
    for (int64_t i = 0; i < n; ++i)
        if (T.live[size_t(i)]) for_operands(T.nodes[size_t(i)], [&](int32_t o) { ++uses[size_t(o)]; });
    for (int32_t id : T.dep_id)
        if (id >= 0) ++uses[size_t(id)];  // kept until the end
    std::vector<IndexSet> sets(static_cast<size_t>(n), IndexSet());
    for (int64_t i = 0; i < n; ++i) {
        if (!T.live[size_t(i)]) continue;
        const ungar_b200_tape_node& nd = T.nodes[size_t(i)];
        IndexSet* const base = sets.data();
        IndexSet cur;
        if (nd.op == UNGAR_B200_OP_INDEP) cur.push_back(nd.a);
        else if (nd.op == UNGAR_B200_OP_CONST) cur.clear();
        else if (is_unary(nd.op)) cur = base[nd.a];
        else if (is_binary(nd.op)) cur = set_union(base[nd.a], base[nd.b]);
        else cur = set_union(base[nd.c], base[nd.d]);  // CondExp: the union of both branches; the compared values carry no derivative
        base[i].swap(cur);
        visit(i, nd, sets);
        for_operands(nd, [&](int32_t o) {
            if (--uses[size_t(o)] == 0) IndexSet().swap(sets[size_t(o)]);
        });
    }
}

void jacobian_pattern(ungar_b200_tape& T) {
    if (T.jp_done) return;
    std::vector<IndexSet> rows(size_t(T.n_dep));
    std::multimap<int32_t, int64_t> outputs_of;
    for (int64_t r = 0; r < T.n_dep; ++r)
        if (T.dep_id[size_t(r)] >= 0) outputs_of.emplace(T.dep_id[size_t(r)], r);
    propagate_sets(T, [&](int64_t i, const ungar_b200_tape_node&, const std::vector<IndexSet>& sets) {
        auto range = outputs_of.equal_range(int32_t(i));
        for (auto it = range.first; it != range.second; ++it) rows[size_t(it->second)] = sets[size_t(i)];
    });
    T.jp_rows.clear();
    T.jp_cols.clear();
    for (int64_t r = 0; r < T.n_dep; ++r)
        for (int32_t c : rows[size_t(r)]) { T.jp_rows.push_back(r); T.jp_cols.push_back(c); }
    T.jp_done = true;
}

// Index pairs that interact through a nonlinear operation on the way to a dependent (full symmetric pattern).
void hessian_pattern(ungar_b200_tape& T) {
    if (T.hp_done) return;
    std::vector<std::set<int32_t>> H(size_t(T.n_indep));
    auto cross = [&](const IndexSet& a, const IndexSet& b) {
        for (int32_t i : a)
            for (int32_t j : b) { H[size_t(i)].insert(j); H[size_t(j)].insert(i); }
    };
    propagate_sets(T, [&](int64_t, const ungar_b200_tape_node& nd, const std::vector<IndexSet>& sets) {
        const IndexSet& a = nd.op == UNGAR_B200_OP_INDEP || nd.op == UNGAR_B200_OP_CONST ? sets[0] : sets[size_t(nd.a)];
        if (is_nonlinear_unary(nd.op)) cross(a, a);
        else if (nd.op == UNGAR_B200_OP_MUL) cross(a, sets[size_t(nd.b)]);
        else if (nd.op == UNGAR_B200_OP_DIV) { cross(a, sets[size_t(nd.b)]); cross(sets[size_t(nd.b)], sets[size_t(nd.b)]); }
        else if (nd.op == UNGAR_B200_OP_POW || nd.op == UNGAR_B200_OP_ATAN2) {
            const IndexSet u = set_union(a, sets[size_t(nd.b)]);
            cross(u, u);
        }
    });
    T.hp_rows.clear();
    T.hp_cols.clear();
    for (int64_t i = 0; i < T.n_indep; ++i)
        for (int32_t j : H[size_t(i)]) { T.hp_rows.push_back(i); T.hp_cols.push_back(j); }
    T.hp_done = true;
}

// Column compression: columns that share no row of the STRUCTURAL pattern get the same colour (greedy, column order), so
// each selected element is read from exactly one direction.
int choose_jacobian(ungar_b200_tape& T, const int64_t* rows, const int64_t* cols, int64_t nnz) {
    jacobian_pattern(T);
    std::vector<std::vector<int32_t>> row_cols(size_t(T.n_dep)), col_rows(size_t(T.n_indep));
    for (size_t e = 0; e < T.jp_rows.size(); ++e) {
        row_cols[size_t(T.jp_rows[e])].push_back(int32_t(T.jp_cols[e]));
        col_rows[size_t(T.jp_cols[e])].push_back(int32_t(T.jp_rows[e]));
    }
    std::vector<int64_t> sel_rows, sel_cols;
    if (rows) {
        for (int64_t e = 0; e < nnz; ++e) {
            if (rows[e] < 0 || rows[e] >= T.n_dep || cols[e] < 0 || cols[e] >= T.n_indep)
                return tfail(UNGAR_B200_EINVAL, "Jacobian element %lld (%lld, %lld) out of range", (long long)e, (long long)rows[e], (long long)cols[e]);
            const auto& rc = row_cols[size_t(rows[e])];
            if (!std::binary_search(rc.begin(), rc.end(), int32_t(cols[e])))
                return tfail(UNGAR_B200_EINVAL, "Jacobian element (%lld, %lld) is structurally zero", (long long)rows[e], (long long)cols[e]);
        }
        sel_rows.assign(rows, rows + nnz);
        sel_cols.assign(cols, cols + nnz);
    } else {
        sel_rows = T.jp_rows;
        sel_cols = T.jp_cols;
    }
    std::vector<char> wanted(size_t(T.n_indep), 0);
    for (int64_t c : sel_cols) wanted[size_t(c)] = 1;
    T.color.assign(size_t(T.n_indep), -1);
    T.n_colors = 0;
    std::vector<int> stamp;
    for (int64_t j = 0; j < T.n_indep; ++j) {
        if (!wanted[size_t(j)]) continue;
        stamp.assign(size_t(T.n_colors) + 1, 0);
        for (int32_t r : col_rows[size_t(j)])
            for (int32_t c : row_cols[size_t(r)])
                if (T.color[size_t(c)] >= 0) stamp[size_t(T.color[size_t(c)])] = 1;
        int col = 0;
        while (stamp[size_t(col)]) ++col;
        T.color[size_t(j)] = col;
        T.n_colors = std::max(T.n_colors, col + 1);
    }
    T.n_colors = std::max(T.n_colors, 1);
    T.jac_slot.assign(size_t(T.n_dep) * size_t(T.n_colors), -1);
    for (size_t e = 0; e < sel_rows.size(); ++e) {
        int& s = T.jac_slot[size_t(sel_rows[e]) * size_t(T.n_colors) + size_t(T.color[size_t(sel_cols[e])])];
        if (s >= 0) return tfail(UNGAR_B200_EINVAL, "Jacobian element (%lld, %lld) listed twice", (long long)sel_rows[e], (long long)sel_cols[e]);
        s = int(e);
    }
    T.j_rows.swap(sel_rows);
    T.j_cols.swap(sel_cols);
    T.j_set = true;
    T.j_uploaded = false;
    // scalar functions: where the reverse sweep puts d y / d x_i
    T.elem_of_indep.clear();
    T.elem_uploaded = false;
    if (T.n_dep == 1) {
        T.elem_of_indep.assign(size_t(T.n_indep), -1);
        for (size_t e = 0; e < T.j_cols.size(); ++e) T.elem_of_indep[size_t(T.j_cols[e])] = int(e);
    }
    return UNGAR_B200_OK;
}

int choose_hessian(ungar_b200_tape& T, const int64_t* rows, const int64_t* cols, int64_t nnz) {
    hessian_pattern(T);
    std::vector<int64_t> sel_rows, sel_cols;
    if (rows) {
        std::set<std::pair<int64_t, int64_t>> pattern;
        for (size_t e = 0; e < T.hp_rows.size(); ++e) pattern.emplace(T.hp_rows[e], T.hp_cols[e]);
        for (int64_t e = 0; e < nnz; ++e) {
            if (rows[e] < 0 || rows[e] >= T.n_indep || cols[e] < 0 || cols[e] >= T.n_indep)
                return tfail(UNGAR_B200_EINVAL, "Hessian element %lld out of range", (long long)e);
            if (!pattern.count({rows[e], cols[e]}))
                return tfail(UNGAR_B200_EINVAL, "Hessian element (%lld, %lld) is structurally zero", (long long)rows[e], (long long)cols[e]);
        }
        {
            std::set<std::pair<int64_t, int64_t>> seen;
            for (int64_t e = 0; e < nnz; ++e)
                if (!seen.emplace(rows[e], cols[e]).second)
                    return tfail(UNGAR_B200_EINVAL, "Hessian element (%lld, %lld) is listed twice", (long long)rows[e], (long long)cols[e]);
        }
        sel_rows.assign(rows, rows + nnz);
        sel_cols.assign(cols, cols + nnz);
    } else {
        sel_rows = T.hp_rows;
        sel_cols = T.hp_cols;
    }
    T.h_pi.clear(); T.h_pj.clear();
    std::map<int64_t, int> diag_dir;
    std::map<std::pair<int64_t, int64_t>, int> pair_dir;
    auto diag = [&](int64_t i) {
        auto it = diag_dir.find(i);
        if (it != diag_dir.end()) return it->second;
        const int d = int(T.h_pi.size());
        T.h_pi.push_back(int(i)); T.h_pj.push_back(int(i));
        diag_dir.emplace(i, d);
        return d;
    };
    T.h_di.assign(sel_rows.size(), 0); T.h_dj.assign(sel_rows.size(), 0); T.h_pr.assign(sel_rows.size(), -1);
    for (size_t e = 0; e < sel_rows.size(); ++e) {
        const int64_t i = std::min(sel_rows[e], sel_cols[e]), j = std::max(sel_rows[e], sel_cols[e]);
        T.h_di[e] = diag(i);
        T.h_dj[e] = diag(j);
        if (i != j) {
            auto it = pair_dir.find({i, j});
            if (it == pair_dir.end()) {
                it = pair_dir.emplace(std::make_pair(i, j), int(T.h_pi.size())).first;
                T.h_pi.push_back(int(i)); T.h_pj.push_back(int(j));
            }
            T.h_pr[e] = it->second;
        }
    }
    if (T.h_pi.empty()) { T.h_pi.push_back(0); T.h_pj.push_back(0); }  // keep the launch geometry non-empty
    T.h_rows.swap(sel_rows);
    T.h_cols.swap(sel_cols);
    T.h_set = true;
    T.h_uploaded = false;
    return UNGAR_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Device side
// ---------------------------------------------------------------------------------------------------------------------
int ensure_device(ungar_b200_tape& T) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || T.device >= count)
        return tfail(UNGAR_B200_ECUDA, "no usable CUDA device %d (there is no CPU fallback for tape evaluation)", T.device);
    UBT_CUDA(cudaSetDevice(T.device));
    if (!T.uploaded) {
        if (int rc = T.d_code.upload(T.code)) return rc;
        if (int rc = T.d_consts.upload(T.consts)) return rc;
        T.uploaded = true;
    }
    return UNGAR_B200_OK;
}

struct Staged {
    const double* d_x;
    int64_t ld_x;
    double* d_out;
    int64_t ld_out;
};

int stage_in(ungar_b200_tape& T, const double* x, int64_t batch, int64_t ld_x, double* out, int64_t ld_out, int64_t n_out, int32_t mem,
             cudaStream_t stream, Staged& s) {
    s = Staged{x, ld_x, out, ld_out};
    if (mem == UNGAR_B200_MEM_HOST) {
        if (int rc = T.ws_x.reserve(size_t(batch) * T.n_indep * sizeof(double))) return rc;
        if (int rc = T.ws_out.reserve(size_t(batch) * std::max<int64_t>(n_out, 1) * sizeof(double))) return rc;
        UBT_CUDA(cudaMemcpy2DAsync(T.ws_x.ptr, T.n_indep * sizeof(double), x, ld_x * sizeof(double), T.n_indep * sizeof(double), batch,
                                   cudaMemcpyHostToDevice, stream));
        s = Staged{static_cast<const double*>(T.ws_x.ptr), T.n_indep, static_cast<double*>(T.ws_out.ptr), std::max<int64_t>(n_out, 1)};
    }
    return UNGAR_B200_OK;
}

int stage_out(const Staged& s, double* out, int64_t ld_out, int64_t n_out, int64_t batch, int32_t mem, cudaStream_t stream) {
    if (mem == UNGAR_B200_MEM_HOST) {
        if (n_out)
            UBT_CUDA(cudaMemcpy2DAsync(out, ld_out * sizeof(double), s.d_out, s.ld_out * sizeof(double), n_out * sizeof(double), batch,
                                       cudaMemcpyDeviceToHost, stream));
        UBT_CUDA(cudaStreamSynchronize(stream));
    }
    return UNGAR_B200_OK;
}

int check_call(const ungar_b200_tape* T, const void* x, int64_t batch, int64_t ld_x, const void* out, int64_t ld_out, int64_t n_out, int32_t mem) {
    if (!T) return tfail(UNGAR_B200_EINVAL, "null tape");
    if (batch < 0 || (batch > 0 && (!x || (!out && n_out > 0)))) return tfail(UNGAR_B200_EINVAL, "null buffer");
    if (ld_x < T->n_indep) return tfail(UNGAR_B200_EINVAL, "ld_x %lld < %lld", (long long)ld_x, (long long)T->n_indep);
    if (ld_out < n_out) return tfail(UNGAR_B200_EINVAL, "output stride %lld < %lld", (long long)ld_out, (long long)n_out);
    if (mem != UNGAR_B200_MEM_DEVICE && mem != UNGAR_B200_MEM_HOST) return tfail(UNGAR_B200_EINVAL, "unknown mem %d", mem);
    return UNGAR_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// NVRTC specialisation.  The register machine pays ~0.3 us per tape instruction regardless of the batch
// (fetch-decode of the instruction record, jets through HBM/L2).  For tapes up to kSpecializeMax instructions the same program
// is emitted as ONE straight-line CUDA kernel — a named Jet<ORDER> variable per slot, the very op functions of tape_machine.cuh —
// compiled with NVRTC for sm_100a and loaded with the driver API.  Where the reference runs gcc on generated C and caches the
// library under its NAME only (function.hpp:420-451: a changed lambda under an old name silently reuses the stale library), the
// compiled kernel is cached under a CONTENT hash: FNV-1a over the instruction stream, the constants, ORDER, the target arch and
// the text of tape_machine.cuh.  UNGAR_B200_KERNEL_CACHE (default: $UNGAR_CODEGEN_FOLDER or /tmp/ungar_b200_kernels) holds
// <hash>.cubin; UNGAR_B200_NO_NVRTC=1 disables the path.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSpecializeMax = 12000;   // instructions per KERNEL; NVRTC + ptxas time grows superlinearly (8 k: seconds, 40 k: minutes)
constexpr int kSegment       = 6000;    // longer tapes are cut into kernels of this many instructions (compile time stays linear)
constexpr int kSegmentedMax  = 100000;
constexpr int kLiveWindow    = 64;      // instructions a value may wait in a register for its next use inside a segmented kernel  // beyond this even the segmented module takes minutes to compile: the interpreter keeps serving
constexpr int64_t kSpecializeAfter = 1; // specialise on the second call of an ORDER: a function evaluated once never pays the compile

// The driver API and NVRTC are bound at run time (dlopen), not at link time: the library must load — and export its ABI — on hosts
// without a GPU driver (libcuda.so.1 is part of the driver, not of the toolkit), where only the host-side analysis is usable.
struct LazyApi {
    bool tried = false, ok = false;
    CUresult (*moduleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*moduleUnload)(CUmodule) = nullptr;
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    nvrtcResult (*createProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*compileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*getProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*getProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*getCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*getCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*destroyProgram)(nvrtcProgram*) = nullptr;
};
LazyApi& lazy_api() {
    static LazyApi api;
    if (api.tried) return api;
    api.tried = true;
    void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    void* rt = dlopen("libnvrtc.so.12", RTLD_NOW | RTLD_GLOBAL);
    if (!rt) rt = dlopen("libnvrtc.so", RTLD_NOW | RTLD_GLOBAL);
    if (!rt) rt = dlopen("/usr/local/cuda/lib64/libnvrtc.so", RTLD_NOW | RTLD_GLOBAL);
    if (!cu || !rt) return api;
    bool all = true;
    auto bind = [&](auto& fn, void* lib, const char* name) {
        fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(lib, name));
        all = all && fn != nullptr;
    };
    bind(api.moduleLoadData, cu, "cuModuleLoadData");
    bind(api.moduleGetFunction, cu, "cuModuleGetFunction");
    bind(api.moduleUnload, cu, "cuModuleUnload");
    bind(api.launchKernel, cu, "cuLaunchKernel");
    bind(api.createProgram, rt, "nvrtcCreateProgram");
    bind(api.compileProgram, rt, "nvrtcCompileProgram");
    bind(api.getProgramLogSize, rt, "nvrtcGetProgramLogSize");
    bind(api.getProgramLog, rt, "nvrtcGetProgramLog");
    bind(api.getCUBINSize, rt, "nvrtcGetCUBINSize");
    bind(api.getCUBIN, rt, "nvrtcGetCUBIN");
    bind(api.destroyProgram, rt, "nvrtcDestroyProgram");
    api.ok = all;
    return api;
}

uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

std::string machine_header_dir() {
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&fnv1a), &info) && info.dli_fname) {
        std::string lib = info.dli_fname;  // .../ungar_b200/libungar_b200.so -> .../ungar_b200/csrc
        const size_t cut = lib.find_last_of('/');
        return (cut == std::string::npos ? std::string(".") : lib.substr(0, cut)) + "/csrc";
    }
    return "csrc";
}

std::string cache_dir() {
    if (const char* e = getenv("UNGAR_B200_KERNEL_CACHE")) return e;
    if (const char* e = getenv("UNGAR_CODEGEN_FOLDER")) return std::string(e) + "/ungar_b200_kernels";
    return "/tmp/ungar_b200_kernels";
}

std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// Straight-line kernel text of the program for one ORDER (same thread mapping and outputs as ub::tape::tape_kernel<ORDER>).
std::string hexd(double v) {  // exact round trip of a double constant
    char buf[64];
    snprintf(buf, sizeof(buf), "%a", v);
    return std::string(buf);
}

// One statement per instruction; slots are the local variables r<slot>.
void emit_instruction(std::ostringstream& o, const ungar_b200_tape& T, const ub::tape::Instr& in, int order) {
    using namespace ub::tape;
    switch (in.op) {
        case T_INDEP:
            o << "  r" << in.dst << " = jet_const<ORDER>(x[" << in.a << "]);";
            if (order >= 1) o << " r" << in.dst << ".d = (S.kind == 0 ? S.color[" << in.a << "] == dir : (" << in.a << " == s0 || " << in.a << " == s1)) ? 1.0 : 0.0;";
            o << "\n";
            break;
        case T_CONST: o << "  r" << in.dst << " = jet_const<ORDER>(" << hexd(T.consts[size_t(in.a)]) << ");\n"; break;
        case T_ADD: case T_SUB: case T_MUL: case T_DIV: case T_ATAN2:
            o << "  r" << in.dst << " = jet_binary<ORDER>(" << in.op << ", r" << in.a << ", r" << in.b << ");\n";
            break;
        case T_POW:
            if (in.c >= 0) o << "  r" << in.dst << " = jet_pow_const<ORDER>(r" << in.a << ", " << hexd(T.consts[size_t(in.c)]) << ");\n";
            else o << "  r" << in.dst << " = jet_binary<ORDER>(" << int(T_POW) << ", r" << in.a << ", r" << in.b << ");\n";
            break;
        case T_CLT: case T_CLE: case T_CGT: case T_CGE: case T_CEQ:
            o << "  r" << in.dst << " = jet_compare(" << in.op << ", r" << in.a << ".v, r" << in.b << ".v) ? r" << in.c << " : r" << in.d << ";\n";
            break;
        case T_OUTPUT: case T_OUTPUT_CONST: {
            const std::string a = in.op == T_OUTPUT ? "r" + std::to_string(in.a) : "jet_const<ORDER>(" + hexd(T.consts[size_t(in.a)]) + ")";
            if (order == 0) o << "  out[" << in.b << "] = " << a << ".v;\n";
            if (order == 1) o << "  { const int e = out_slot[(long long)" << in.b << " * ndir + dir]; if (e >= 0) out[e] = " << a << ".d; }\n";
            if (order == 2) o << "  acc += weights[" << in.b << "] * " << a << ".dd;\n";
            break;
        }
        default:  // unary
            o << "  r" << in.dst << " = jet_unary<ORDER>(" << in.op << ", r" << in.a << ");\n";
    }
}

const char* kKernelProlog =
    "  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;\n  if (t >= batch * ndir) return;\n"
    "  const long long b = t / ndir;\n  const int dir = int(t - b * ndir);\n"
    "  const double* __restrict__ x = x_all + b * ld_x;\n  double* __restrict__ out = out_all + b * ld_out;\n"
    "  int s0 = -1, s1 = -1;\n  if (ORDER >= 1 && S.kind == 1) { s0 = S.pi[dir]; s1 = S.pj[dir]; }\n  double acc = 0.0;\n  (void)s0; (void)s1; (void)acc; (void)x; (void)out;\n";

std::string generate_kernel_source(const ungar_b200_tape& T, int order) {
    std::ostringstream o;
    o.precision(17);
    o << "#include \"tape_machine.cuh\"\nusing namespace ub::tape;\n"
      << "extern \"C\" __global__ void __launch_bounds__(128) tape_special(Seeds S, const double* __restrict__ x_all, long long ld_x, long long batch, int ndir,\n"
      << "    double* __restrict__ out_all, long long ld_out, const int* __restrict__ out_slot, const double* __restrict__ weights) {\n"
      << "  constexpr int ORDER = " << order << ";\n"
      << kKernelProlog;
    for (int k = 0; k < T.n_slots; ++k) o << "  Jet<ORDER> r" << k << ";\n";
    for (const ub::tape::Instr& in : T.code) emit_instruction(o, T, in, order);
    if (order == 2) o << "  out[dir] = acc;\n";
    o << "}\n";
    return o.str();
}

// Reverse sweep of a SCALAR function: one thread per xp vector computes the whole gradient in two passes over the tape instead of one
// forward pass per colour (an objective has one dense row: as many colours as independents it depends on).  Forward: the value
// program as in the single-kernel form, every instruction's value also written to the scratch array V[instruction][thread].
// Backward: adjoints live in registers named after the SLOT of the value they belong to — the adjoint of a value is live exactly where
// the value was (from its last use back to its definition), so the liveness-based slot assignment of the forward program is a valid
// register assignment for the adjoints too; an instruction takes the adjoint of its result, clears it (the slot's previous occupant
// starts from zero) and adds its partial derivatives, read from V, to the adjoints of its operands' slots.
void instr_uses(const ub::tape::Instr& in, int reads[4], int& write);

std::string generate_reverse_source(const ungar_b200_tape& T) {
    using namespace ub::tape;
    const int n = int(T.code.size());
    std::vector<int> prod(size_t(T.n_slots), -1);               // instruction that produced the slot's current value
    std::vector<std::array<int, 4>> from(size_t(n), std::array<int, 4>{-1, -1, -1, -1});
    for (int i = 0; i < n; ++i) {
        int rd[4], wr;
        instr_uses(T.code[size_t(i)], rd, wr);
        for (int q = 0; q < 4; ++q)
            if (rd[q] >= 0) from[size_t(i)][size_t(q)] = prod[size_t(rd[q])];
        if (wr >= 0) prod[size_t(wr)] = i;
    }
    std::ostringstream o;
    o.precision(17);
    o << "#include \"tape_machine.cuh\"\nusing namespace ub::tape;\n#define V(i) scratch[(long long)(i) * stride + t]\n"
      << "extern \"C\" __global__ void __launch_bounds__(128) tape_reverse(const double* __restrict__ x_all, long long ld_x, long long batch,\n"
      << "    double* __restrict__ out_all, long long ld_out, const int* __restrict__ elem, int nnz, double* __restrict__ scratch, long long stride) {\n"
      << "  constexpr int ORDER = 0;\n"
      << "  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;\n  if (t >= batch) return;\n"
      << "  const double* __restrict__ x = x_all + t * ld_x;\n  double* __restrict__ out = out_all + t * ld_out;\n"
      << "  for (int e = 0; e < nnz; ++e) out[e] = 0.0;\n";
    for (int k = 0; k < T.n_slots; ++k) o << "  Jet<ORDER> r" << k << "; double a" << k << " = 0.0;\n";
    for (int i = 0; i < n; ++i) {
        const Instr& in = T.code[size_t(i)];
        if (in.op == T_OUTPUT || in.op == T_OUTPUT_CONST) continue;
        emit_instruction(o, T, in, 0);
        o << "  V(" << i << ") = r" << in.dst << ".v;\n";
    }
    auto A = [](int slot) { return "a" + std::to_string(slot); };
    for (int i = n - 1; i >= 0; --i) {
        const Instr& in = T.code[size_t(i)];
        const std::array<int, 4>& pr = from[size_t(i)];
        const std::string va = "V(" + std::to_string(pr[0]) + ")", vb = "V(" + std::to_string(pr[1]) + ")", vi = "V(" + std::to_string(i) + ")";
        if (in.op == T_OUTPUT) { o << "  " << A(in.a) << " += 1.0;\n"; continue; }
        if (in.op == T_OUTPUT_CONST) continue;
        if (in.op == T_INDEP) { o << "  { const int e = elem[" << in.a << "]; if (e >= 0) out[e] += " << A(in.dst) << "; " << A(in.dst) << " = 0.0; }\n"; continue; }
        if (in.op == T_CONST) { o << "  " << A(in.dst) << " = 0.0;\n"; continue; }
        o << "  { const double g = " << A(in.dst) << "; " << A(in.dst) << " = 0.0; ";
        switch (in.op) {
            case T_ADD: o << A(in.a) << " += g; " << A(in.b) << " += g;"; break;
            case T_SUB: o << A(in.a) << " += g; " << A(in.b) << " -= g;"; break;
            case T_MUL: o << "const double pa = " << va << ", pb = " << vb << "; " << A(in.a) << " += g * pb; " << A(in.b) << " += g * pa;"; break;
            case T_DIV: o << "const double ib = 1.0 / " << vb << "; " << A(in.a) << " += g * ib; " << A(in.b) << " -= g * " << vi << " * ib;"; break;
            case T_ATAN2: o << "const double pa = " << va << ", pb = " << vb << ", inv = g / (pa * pa + pb * pb); " << A(in.a) << " += pb * inv; " << A(in.b) << " -= pa * inv;"; break;
            case T_POW:
                if (in.c >= 0) o << "const double k = " << hexd(T.consts[size_t(in.c)]) << "; " << A(in.a) << " += g * k * pow(" << va << ", k - 1.0);";
                else o << "const double pa = " << va << ", pb = " << vb << "; " << A(in.a) << " += g * pb * pow(pa, pb - 1.0); " << A(in.b) << " += g * " << vi << " * log(pa);";
                break;
            case T_CLT: case T_CLE: case T_CGT: case T_CGE: case T_CEQ:
                o << "if (jet_compare(" << in.op << ", " << va << ", " << vb << ")) " << A(in.c) << " += g; else " << A(in.d) << " += g;";
                break;
            case T_NEG: o << A(in.a) << " -= g;"; break;
            case T_SQRT: o << A(in.a) << " += g * 0.5 / " << vi << ";"; break;
            case T_SIN: o << A(in.a) << " += g * cos(" << va << ");"; break;
            case T_COS: o << A(in.a) << " -= g * sin(" << va << ");"; break;
            case T_TAN: o << "const double y = " << vi << "; " << A(in.a) << " += g * (1.0 + y * y);"; break;
            case T_ATAN: o << "const double pa = " << va << "; " << A(in.a) << " += g / (1.0 + pa * pa);"; break;
            case T_ACOS: o << "const double pa = " << va << "; " << A(in.a) << " -= g * rsqrt(1.0 - pa * pa);"; break;
            case T_ASIN: o << "const double pa = " << va << "; " << A(in.a) << " += g * rsqrt(1.0 - pa * pa);"; break;
            case T_EXP: o << A(in.a) << " += g * " << vi << ";"; break;
            case T_LOG: o << A(in.a) << " += g / " << va << ";"; break;
            default: o << "const double pa = " << va << "; " << A(in.a) << " += g * double((pa > 0.0) - (pa < 0.0));";  // T_ABS
        }
        o << " }\n";
    }
    o << "}\n";
    return o.str();
}

// Which slots an instruction reads / writes (-1: none).
void instr_uses(const ub::tape::Instr& in, int reads[4], int& write) {
    using namespace ub::tape;
    reads[0] = reads[1] = reads[2] = reads[3] = -1;
    write = -1;
    switch (in.op) {
        case T_INDEP: case T_CONST: write = in.dst; break;
        case T_ADD: case T_SUB: case T_MUL: case T_DIV: case T_ATAN2: reads[0] = in.a; reads[1] = in.b; write = in.dst; break;
        case T_POW: reads[0] = in.a; if (in.c < 0) reads[1] = in.b; write = in.dst; break;
        case T_CLT: case T_CLE: case T_CGT: case T_CGE: case T_CEQ: reads[0] = in.a; reads[1] = in.b; reads[2] = in.c; reads[3] = in.d; write = in.dst; break;
        case T_OUTPUT: reads[0] = in.a; break;
        case T_OUTPUT_CONST: break;
        default: reads[0] = in.a; write = in.dst;  // unary
    }
}

// A tape beyond kSpecializeMax as a sequence of kernels of kSegment instructions each.  Inside a kernel the slots are registers, as in
// the single-kernel form; a value that crosses a cut travels through the scratch array of the interpreter ([slot][component][thread],
// coalesced): kernel k loads the slots it reads before writing them and stores the slots it wrote that a later kernel still reads
// (backward liveness over the cuts).  The Hessian order's accumulator crosses the cuts in one extra scratch row.
std::string generate_segmented_source(const ungar_b200_tape& T, int order, int& n_parts) {
    const int n = int(T.code.size());
    n_parts = (n + kSegment - 1) / kSegment;
    const size_t np = size_t(n_parts), ns = size_t(T.n_slots);
    std::vector<std::vector<int>> loads(np), stores(np), touched(np);
    std::vector<char> live(ns, 0);  // live at the END of the segment being processed (backward)
    for (int k = n_parts - 1; k >= 0; --k) {
        const int i0 = k * kSegment, i1 = std::min(n, i0 + kSegment);
        std::vector<char> written(ns, 0), exposed(ns, 0), seen(ns, 0);
        for (int i = i0; i < i1; ++i) {
            int rd[4], wr;
            instr_uses(T.code[size_t(i)], rd, wr);
            for (int r : rd)
                if (r >= 0) {
                    if (!written[size_t(r)]) exposed[size_t(r)] = 1;  // read before any write of this segment: comes from an earlier one
                    seen[size_t(r)] = 1;
                }
            if (wr >= 0) { written[size_t(wr)] = 1; seen[size_t(wr)] = 1; }
        }
        for (int sl = 0; sl < T.n_slots; ++sl) {
            if (seen[size_t(sl)]) touched[size_t(k)].push_back(sl);
            if (exposed[size_t(sl)]) loads[size_t(k)].push_back(sl);
            if (written[size_t(sl)] && live[size_t(sl)]) stores[size_t(k)].push_back(sl);
        }
        for (int sl = 0; sl < T.n_slots; ++sl) live[size_t(sl)] = exposed[size_t(sl)] || (live[size_t(sl)] && !written[size_t(sl)]);
    }
    std::ostringstream o;
    o.precision(17);
    o << "#include \"tape_machine.cuh\"\nusing namespace ub::tape;\n";
    for (int k = 0; k < n_parts; ++k) {
        const int i0 = k * kSegment, i1 = std::min(n, i0 + kSegment);
        o << "extern \"C\" __global__ void __launch_bounds__(128) tape_part_" << k
          << "(Seeds S, const double* __restrict__ x_all, long long ld_x, long long batch, int ndir,\n"
          << "    double* __restrict__ out_all, long long ld_out, const int* __restrict__ out_slot, const double* __restrict__ weights,\n"
          << "    double* __restrict__ scratch, long long stride) {\n"
          << "  constexpr int ORDER = " << order << ";\n"
          << kKernelProlog;
        // a value that comes from an earlier kernel is loaded right before its first use, a value a later kernel needs is stored right
        // after its last write: live ranges stay as short as in the tape itself (loading every live-in at the top made hundreds of jets
        // live at once: kilobytes of spills per thread, and a stack size that slowed every later launch of the process)
        std::vector<char> need_load(ns, 0), need_store(ns, 0);
        for (int sl : loads[size_t(k)]) need_load[size_t(sl)] = 1;
        for (int sl : stores[size_t(k)]) need_store[size_t(sl)] = 1;
        std::vector<int> last_write(ns, -1);
        for (int i = i0; i < i1; ++i) {
            int rd[4], wr;
            instr_uses(T.code[size_t(i)], rd, wr);
            if (wr >= 0) last_write[size_t(wr)] = i;
        }
        // ... and inside a kernel a value whose next use lies more than kLiveWindow instructions ahead is parked in the scratch array
        // and re-loaded there: the quadruped tapes keep 700-2300 values alive at once, which ptxas would spill to thread-local memory
        // (a stack frame of kilobytes, and launches of every other kernel of the process slowed down by the local-memory pool it forces)
        std::set<std::pair<int, int>> park_after, reload_before;  // (instruction, slot)
        {
            std::vector<int> last_pos(ns, -1);     // last access of the slot's CURRENT value in this kernel
            std::vector<char> in_scratch(ns, 0);   // the current value is in the scratch array
            for (int sl : loads[size_t(k)]) in_scratch[size_t(sl)] = 1;
            for (int i = i0; i < i1; ++i) {
                int rd[4], wr;
                instr_uses(T.code[size_t(i)], rd, wr);
                for (int r : rd) {
                    if (r < 0) continue;
                    const int lp = last_pos[size_t(r)];
                    if (lp >= 0 && lp != i && i - lp > kLiveWindow) {
                        if (!in_scratch[size_t(r)]) { park_after.insert({lp, r}); in_scratch[size_t(r)] = 1; }
                        reload_before.insert({i, r});
                    }
                    last_pos[size_t(r)] = i;
                }
                if (wr >= 0) { last_pos[size_t(wr)] = i; in_scratch[size_t(wr)] = 0; }
            }
        }
        for (int sl : touched[size_t(k)]) o << "  Jet<ORDER> r" << sl << ";\n";
        // the Hessian order accumulates sum_i w_i y_i'' over the outputs: the partial sum crosses the cuts in one extra scratch row
        if (order == 2 && k > 0) o << "  acc = scratch[(long long)" << T.n_slots << " * (ORDER + 1) * stride + t];\n";
        for (int i = i0; i < i1; ++i) {
            int rd[4], wr;
            instr_uses(T.code[size_t(i)], rd, wr);
            for (int r : rd)
                if (r >= 0 && (need_load[size_t(r)] || reload_before.count({i, r}))) {
                    o << "  r" << r << " = load_slot<ORDER>(scratch, stride, t, " << r << ");\n";
                    need_load[size_t(r)] = 0;
                    reload_before.erase({i, r});
                }
            if (wr >= 0) need_load[size_t(wr)] = 0;  // (an exposed read always precedes the first write, so this never drops a load)
            emit_instruction(o, T, T.code[size_t(i)], order);
            bool stored = false;
            if (wr >= 0 && need_store[size_t(wr)] && last_write[size_t(wr)] == i) {
                o << "  store_slot<ORDER>(scratch, stride, t, " << wr << ", r" << wr << ");\n";
                stored = true;
            }
            // park values whose next use is far away (the definition itself, or an operand last touched here)
            if (wr >= 0 && !stored && park_after.count({i, wr})) o << "  store_slot<ORDER>(scratch, stride, t, " << wr << ", r" << wr << ");\n";
            for (int r : rd)
                if (r >= 0 && r != wr && park_after.count({i, r})) {
                    o << "  store_slot<ORDER>(scratch, stride, t, " << r << ", r" << r << ");\n";
                    park_after.erase({i, r});
                }
        }
        if (order == 2) {
            if (k + 1 < n_parts) o << "  scratch[(long long)" << T.n_slots << " * (ORDER + 1) * stride + t] = acc;\n";
            else o << "  out[dir] = acc;\n";
        }
        o << "}\n";
    }
    return o.str();
}

bool driver_ok(CUresult r) { return r == CUDA_SUCCESS; }

// Compiles (or loads from the content-hashed cache) the specialised kernel of `order`.  Never fails the call: on any problem the
// state becomes -1 and the interpreter keeps serving.
int specialize_impl(ungar_b200_tape& T, int order) {
    ungar_b200_tape::Special& S = T.special[order];
    const char* off = getenv("UNGAR_B200_NO_NVRTC");
    if ((off && off[0] == '1') || T.code.empty()) return -1;
    if (order == 3 && (int(T.code.size()) > kSpecializeMax || T.n_dep != 1)) return -1;  // the reverse sweep: one kernel, scalar functions
    const bool segmented = order < 3 && int(T.code.size()) > kSpecializeMax;
    if (segmented && int(T.code.size()) > kSegmentedMax) return -1;
    if (segmented) {
        const char* seg_off = getenv("UNGAR_B200_NO_SEGMENTS");  // measurement switch: long tapes stay on the interpreter
        if (seg_off && seg_off[0] == '1') return -1;
    }
    LazyApi& api = lazy_api();
    if (!api.ok) return -1;
    const std::string dir = machine_header_dir();
    const std::string header = slurp(dir + "/tape_machine.cuh");
    if (header.empty()) return -1;
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, T.device);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, T.device);
    if (major != 10) return -1;  // sm_100a only
    const std::string arch = "sm_100a";
    uint64_t h = fnv1a(T.code.data(), T.code.size() * sizeof(Instr));
    h = fnv1a(T.consts.data(), T.consts.size() * sizeof(double), h);
    h = fnv1a(&order, sizeof(order), h);
    const int dims[6] = {T.n_slots, int(T.n_indep), int(T.n_dep), 5 /* generator version */, segmented ? kSegment : 0, segmented ? 1 : 0};
    h = fnv1a(dims, sizeof(dims), h);
    h = fnv1a(arch.data(), arch.size(), h);
    h = fnv1a(header.data(), header.size(), h);
    S.key = h;
    char name[64];
    snprintf(name, sizeof(name), "%016llx.cubin", (unsigned long long)h);
    const std::string cdir = cache_dir(), path = cdir + "/" + name;
    std::string cubin = slurp(path);
    S.from_cache = !cubin.empty();
    if (cubin.empty()) {
        int n_parts_gen = 0;
        const std::string src = order == 3 ? generate_reverse_source(T) : segmented ? generate_segmented_source(T, order, n_parts_gen) : generate_kernel_source(T, order);
        nvrtcProgram prog;
        if (api.createProgram(&prog, src.c_str(), "tape_special.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) return -1;
        const std::string inc1 = "-I" + dir, inc2 = "-I/usr/local/cuda/include", a = "--gpu-architecture=" + arch;
        const char* opts[] = {a.c_str(), inc1.c_str(), inc2.c_str(), "--std=c++17", "-default-device", "--fmad=true"};
        const auto t0 = std::chrono::steady_clock::now();
        const nvrtcResult rc = api.compileProgram(prog, 6, opts);
        S.compile_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rc != NVRTC_SUCCESS) {
            size_t n = 0;
            api.getProgramLogSize(prog, &n);
            std::string log(n, '\0');
            api.getProgramLog(prog, log.data());
            if (getenv("UNGAR_B200_NVRTC_VERBOSE")) fprintf(stderr, "ungar_b200: NVRTC failed, the interpreter serves this tape:\n%s\n", log.c_str());
            api.destroyProgram(&prog);
            return -1;
        }
        size_t n = 0;
        if (api.getCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) { api.destroyProgram(&prog); return -1; }
        cubin.resize(n);
        api.getCUBIN(prog, cubin.data());
        api.destroyProgram(&prog);
        mkdir(cdir.c_str(), 0755);
        const std::string tmp = path + ".tmp" + std::to_string(getpid());
        std::ofstream f(tmp, std::ios::binary);
        f.write(cubin.data(), std::streamsize(cubin.size()));
        f.close();
        if (f) rename(tmp.c_str(), path.c_str());
    }
    cudaFree(nullptr);  // make sure the runtime's primary context is current for the driver API
    if (!driver_ok(api.moduleLoadData(&S.module, cubin.data()))) return -1;
    if (segmented) {
        const int n_parts = (int(T.code.size()) + kSegment - 1) / kSegment;
        S.parts.assign(size_t(n_parts), nullptr);
        for (int k = 0; k < n_parts; ++k) {
            const std::string fname = "tape_part_" + std::to_string(k);
            if (!driver_ok(api.moduleGetFunction(&S.parts[size_t(k)], S.module, fname.c_str()))) { S.parts.clear(); return -1; }
        }
    } else if (!driver_ok(api.moduleGetFunction(&S.fn, S.module, order == 3 ? "tape_reverse" : "tape_special"))) {
        return -1;
    }
    return 1;
}


// Short tapes compile in seconds, in the calling thread.  The segmented kernels of a long tape take tens of seconds: a worker thread
// compiles them while the interpreter keeps serving the calls; the specialised kernels take over from the first call after the
// compile has finished (ungar_b200_tape_special_wait blocks until then).  UNGAR_B200_NVRTC_SYNC=1 compiles everything in the caller.
void specialize(ungar_b200_tape& T, int order) {
    ungar_b200_tape::Special& S = T.special[order];
    const char* sync = getenv("UNGAR_B200_NVRTC_SYNC");
    lazy_api();  // bind NVRTC / the driver API in this thread (the first use is not thread-safe)
    if (order < 3 && int(T.code.size()) > kSpecializeMax && !(sync && sync[0] == '1')) {
        S.state.store(2, std::memory_order_release);
        ungar_b200_tape* tp = &T;
        S.worker = std::thread([tp, order] {
            cudaSetDevice(tp->device);
            tp->special[order].state.store(specialize_impl(*tp, order), std::memory_order_release);
        });
    } else {
        S.state.store(specialize_impl(T, order), std::memory_order_release);
    }
}

template <int ORDER>
int launch(ungar_b200_tape& T, const ub::tape::Seeds& seeds, const double* d_x, int64_t ld_x, int64_t batch, int ndir, double* d_out,
           int64_t ld_out, const int* out_slot, const double* weights, cudaStream_t stream) {
    const long long threads = (long long)batch * ndir;
    const long long stride  = (threads + 31) & ~31LL;
    const long long blocks = (threads + 127) / 128;
    if (blocks > 2147483647LL) return tfail(UNGAR_B200_EINVAL, "batch x directions too large for one launch");
    // the straight-line kernel from the second call of this ORDER on (a function evaluated once never pays the compile)
    if (T.calls[ORDER]++ >= kSpecializeAfter && T.special[ORDER].state.load(std::memory_order_acquire) == 0) specialize(T, ORDER);
    const int special_state = T.special[ORDER].state.load(std::memory_order_acquire);
    if (special_state == 1 && !T.special[ORDER].parts.empty()) {  // a long tape: its kernels back to back, values cross through the scratch
        if (int rc = T.scratch.reserve(size_t(T.n_slots + 1) * (ORDER + 1) * size_t(stride) * sizeof(double))) return rc;  // + the Hessian accumulator's row
        ub::tape::Seeds sd = seeds;
        long long ldx = ld_x, b64 = batch, ldo = ld_out, st = stride;
        int nd = ndir;
        double* scr = static_cast<double*>(T.scratch.ptr);
        void* args[] = {&sd, &d_x, &ldx, &b64, &nd, &d_out, &ldo, &out_slot, &weights, &scr, &st};
        for (CUfunction fn : T.special[ORDER].parts) {
            const CUresult r = lazy_api().launchKernel(fn, unsigned(blocks), 1, 1, 128, 1, 1, 0, reinterpret_cast<CUstream>(stream), args, nullptr);
            if (r != CUDA_SUCCESS) return tfail(UNGAR_B200_ECUDA, "cuLaunchKernel of a segment of the specialised tape kernel failed (%d)", int(r));
            ub_count_launch();
        }
        return UNGAR_B200_OK;
    }
    if (special_state == 1) {
        ub::tape::Seeds sd = seeds;
        long long ldx = ld_x, b64 = batch, ldo = ld_out;
        int nd = ndir;
        void* args[] = {&sd, &d_x, &ldx, &b64, &nd, &d_out, &ldo, &out_slot, &weights};
        const CUresult r = lazy_api().launchKernel(T.special[ORDER].fn, unsigned(blocks), 1, 1, 128, 1, 1, 0, reinterpret_cast<CUstream>(stream), args, nullptr);
        if (r != CUDA_SUCCESS) return tfail(UNGAR_B200_ECUDA, "cuLaunchKernel of the specialised tape kernel failed (%d)", int(r));
        ub_count_launch();
        return UNGAR_B200_OK;
    }
    if (int rc = T.scratch.reserve(size_t(T.n_slots) * (ORDER + 1) * size_t(stride) * sizeof(double))) return rc;
    const ub::tape::Program P{static_cast<const Instr*>(T.d_code.ptr), static_cast<const double*>(T.d_consts.ptr), int(T.code.size()),
                              T.n_slots, int(T.n_indep), int(T.n_dep)};
    ub::tape::tape_kernel<ORDER><<<unsigned(blocks), 128, 0, stream>>>(P, seeds, d_x, ld_x, batch, ndir, static_cast<double*>(T.scratch.ptr),
                                                                         stride, d_out, ld_out, out_slot, weights);
    ub_count_launch();
    UBT_CUDA(cudaGetLastError());
    return UNGAR_B200_OK;
}

// Gradient of a scalar function by the generated reverse sweep (generate_reverse_source): one thread per xp vector.  Returns 1 when the
// call was served, 0 when the caller should use the forward path (not a scalar function, too few colours to pay, kernel unavailable).
constexpr int kReverseMinColors = 8;
int launch_reverse(ungar_b200_tape& T, const double* d_x, int64_t ld_x, int64_t batch, double* d_out, int64_t ld_out, int nnz, cudaStream_t stream, int& served) {
    served = 0;
    // OPT-IN (UNGAR_B200_REVERSE=1): the generator is validated on the CPU (tests/test_tape_host.py replays the generated text against
    // differences and compiles it with NVRTC) but has not run on a GPU yet, and its compile time grows to minutes for tapes of
    // several thousand instructions — too long for a synchronous first use.
    const char* on = getenv("UNGAR_B200_REVERSE");
    if (!(on && on[0] == '1')) return UNGAR_B200_OK;
    if (T.n_dep != 1 || T.n_colors < kReverseMinColors || int(T.code.size()) > kSpecializeMax) return UNGAR_B200_OK;
    if (T.calls[3]++ >= kSpecializeAfter && T.special[3].state.load(std::memory_order_acquire) == 0) specialize(T, 3);
    if (T.special[3].state.load(std::memory_order_acquire) != 1) return UNGAR_B200_OK;
    if (T.elem_of_indep.empty()) return UNGAR_B200_OK;
    if (!T.elem_uploaded) {
        if (int rc = T.d_elem.upload(T.elem_of_indep)) return rc;
        T.elem_uploaded = true;
    }
    const long long stride = (batch + 31) & ~31LL;
    const long long blocks = (batch + 127) / 128;
    if (int rc = T.scratch.reserve(size_t(T.code.size()) * size_t(stride) * sizeof(double))) return rc;
    long long ldx = ld_x, b64 = batch, ldo = ld_out, st = stride;
    int nz = nnz;
    const int* elem = static_cast<const int*>(T.d_elem.ptr);
    double* scr = static_cast<double*>(T.scratch.ptr);
    void* args[] = {&d_x, &ldx, &b64, &d_out, &ldo, &elem, &nz, &scr, &st};
    const CUresult r = lazy_api().launchKernel(T.special[3].fn, unsigned(blocks), 1, 1, 128, 1, 1, 0, reinterpret_cast<CUstream>(stream), args, nullptr);
    if (r != CUDA_SUCCESS) return tfail(UNGAR_B200_ECUDA, "cuLaunchKernel of the reverse-sweep kernel failed (%d)", int(r));
    ub_count_launch();
    served = 1;
    return UNGAR_B200_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

int ungar_b200_tape_create(const ungar_b200_tape_node* nodes, int64_t n_nodes, int64_t n_independent, const int32_t* dependents,
                           const double* dependent_constants, int64_t n_dependent, int32_t device, ungar_b200_tape** out) {
    if (!out) return tfail(UNGAR_B200_EINVAL, "null output handle");
    *out = nullptr;
    if (n_nodes < 0 || n_independent < 0 || n_dependent < 0 || (n_nodes > 0 && !nodes) || (n_dependent > 0 && !dependents))
        return tfail(UNGAR_B200_EINVAL, "bad tape sizes or null arrays");
    if (n_nodes > 2000000000LL || n_independent > 2000000000LL) return tfail(UNGAR_B200_EINVAL, "tape too large");
    ungar_b200_tape* T = new (std::nothrow) ungar_b200_tape;
    if (!T) return tfail(UNGAR_B200_ENOMEM, "out of host memory");
    T->device  = device;
    T->n_indep = n_independent;
    T->n_dep   = n_dependent;
    T->nodes.assign(nodes, nodes + n_nodes);
    T->dep_id.assign(dependents, dependents + n_dependent);
    T->dep_const.assign(size_t(n_dependent), 0.0);
    if (dependent_constants) T->dep_const.assign(dependent_constants, dependent_constants + n_dependent);
    if (int rc = validate(*T)) {
        delete T;
        return rc;
    }
    build_program(*T);
    *out = T;
    return UNGAR_B200_OK;
}

int ungar_b200_tape_kernel_source(const ungar_b200_tape* tape, int32_t order, char* buffer, int64_t capacity, int64_t* required, int32_t* n_kernels) {
    if (!tape || order < 0 || order > 3 || capacity < 0 || (capacity > 0 && !buffer)) return tfail(UNGAR_B200_EINVAL, "bad argument");
    if (order == 3 && (tape->n_dep != 1 || int(tape->code.size()) > kSpecializeMax))
        return tfail(UNGAR_B200_EUNSUPPORTED, "the reverse sweep serves scalar functions of up to %d instructions", kSpecializeMax);
    int parts = 1;
    const std::string src = order == 3 ? generate_reverse_source(*tape)
                            : int(tape->code.size()) > kSpecializeMax ? generate_segmented_source(*tape, order, parts) : generate_kernel_source(*tape, order);
    if (required) *required = int64_t(src.size()) + 1;
    if (n_kernels) *n_kernels = parts;
    if (capacity > 0) {
        const size_t n = std::min<size_t>(src.size(), size_t(capacity) - 1);
        memcpy(buffer, src.data(), n);
        buffer[n] = 0;
    }
    return UNGAR_B200_OK;
}

// info[4 * order + {0, 1, 2, 3}] for order 0..2: state (0 not tried, 1 specialised, -1 interpreter), served from the kernel cache (0 / 1),
// low and high 32 bits of the content hash.  Test / diagnostics hook of the NVRTC path.
int ungar_b200_tape_special_info(const ungar_b200_tape* tape, int64_t* info) {
    if (!tape || !info) return tfail(UNGAR_B200_EINVAL, "null argument");
    for (int o = 0; o < 4; ++o) {
        const int st = tape->special[o].state.load(std::memory_order_acquire);
        info[4 * o + 0] = st;
        info[4 * o + 1] = st != 2 && tape->special[o].from_cache ? 1 : 0;  // (a worker may still be writing these)
        info[4 * o + 2] = st != 2 ? int64_t(tape->special[o].key & 0xffffffffull) : 0;
        info[4 * o + 3] = st != 2 ? int64_t(tape->special[o].key >> 32) : 0;
    }
    return UNGAR_B200_OK;
}

int ungar_b200_tape_special_wait(ungar_b200_tape* tape) {
    if (!tape) return tfail(UNGAR_B200_EINVAL, "null argument");
    for (auto& sp : tape->special)
        if (sp.worker.joinable()) sp.worker.join();
    return UNGAR_B200_OK;
}

int ungar_b200_tape_destroy(ungar_b200_tape* tape) {
    if (tape)
        for (auto& sp : tape->special) {
            if (sp.worker.joinable()) sp.worker.join();
            if (sp.module && lazy_api().ok) lazy_api().moduleUnload(sp.module);
        }
    if (!tape) return UNGAR_B200_OK;
    int count = 0;
    if (cudaGetDeviceCount(&count) == cudaSuccess && tape->device < count) cudaSetDevice(tape->device);
    delete tape;
    return UNGAR_B200_OK;
}

int ungar_b200_tape_info(const ungar_b200_tape* tape, int64_t* info) {
    if (!tape || !info) return tfail(UNGAR_B200_EINVAL, "null argument");
    info[0] = tape->n_indep; info[1] = tape->n_dep; info[2] = tape->n_live; info[3] = tape->n_slots;
    info[4] = tape->j_set ? tape->n_colors : 0;
    info[5] = tape->h_set ? int64_t(tape->h_pi.size()) : 0;
    return UNGAR_B200_OK;
}

int ungar_b200_tape_jacobian_pattern(ungar_b200_tape* tape, const int64_t** rows, const int64_t** cols, int64_t* nnz) {
    if (!tape || !rows || !cols || !nnz) return tfail(UNGAR_B200_EINVAL, "null argument");
    jacobian_pattern(*tape);
    *rows = tape->jp_rows.data(); *cols = tape->jp_cols.data(); *nnz = int64_t(tape->jp_rows.size());
    return UNGAR_B200_OK;
}

int ungar_b200_tape_hessian_pattern(ungar_b200_tape* tape, const int64_t** rows, const int64_t** cols, int64_t* nnz) {
    if (!tape || !rows || !cols || !nnz) return tfail(UNGAR_B200_EINVAL, "null argument");
    hessian_pattern(*tape);
    *rows = tape->hp_rows.data(); *cols = tape->hp_cols.data(); *nnz = int64_t(tape->hp_rows.size());
    return UNGAR_B200_OK;
}

int ungar_b200_tape_set_jacobian_elements(ungar_b200_tape* tape, const int64_t* rows, const int64_t* cols, int64_t nnz) {
    if (!tape || (rows && !cols) || nnz < 0) return tfail(UNGAR_B200_EINVAL, "bad argument");
    return choose_jacobian(*tape, rows, cols, nnz);
}

int ungar_b200_tape_set_hessian_elements(ungar_b200_tape* tape, const int64_t* rows, const int64_t* cols, int64_t nnz) {
    if (!tape || (rows && !cols) || nnz < 0) return tfail(UNGAR_B200_EINVAL, "bad argument");
    return choose_hessian(*tape, rows, cols, nnz);
}

int ungar_b200_tape_forward_zero(ungar_b200_tape* tape, const double* x, int64_t batch, int64_t ld_x, double* y, int64_t ld_y, int32_t mem,
                                 void* stream_) {
    if (int rc = check_call(tape, x, batch, ld_x, y, ld_y, tape ? tape->n_dep : 0, mem)) return rc;
    if (batch == 0) return UNGAR_B200_OK;
    if (int rc = ensure_device(*tape)) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Staged s;
    if (int rc = stage_in(*tape, x, batch, ld_x, y, ld_y, tape->n_dep, mem, stream, s)) return rc;
    if (int rc = launch<0>(*tape, ub::tape::Seeds{0, nullptr, nullptr, nullptr}, s.d_x, s.ld_x, batch, 1, s.d_out, s.ld_out, nullptr, nullptr, stream))
        return rc;
    return stage_out(s, y, ld_y, tape->n_dep, batch, mem, stream);
}

int ungar_b200_tape_sparse_jacobian(ungar_b200_tape* tape, const double* x, int64_t batch, int64_t ld_x, double* vals, int64_t ld_vals,
                                    int32_t mem, void* stream_) {
    if (!tape) return tfail(UNGAR_B200_EINVAL, "null tape");
    if (!tape->j_set)
        if (int rc = choose_jacobian(*tape, nullptr, nullptr, 0)) return rc;
    const int64_t nnz = int64_t(tape->j_rows.size());
    if (int rc = check_call(tape, x, batch, ld_x, vals, ld_vals, nnz, mem)) return rc;
    if (batch == 0 || nnz == 0) return UNGAR_B200_OK;
    if (int rc = ensure_device(*tape)) return rc;
    if (!tape->j_uploaded) {
        if (int rc = tape->d_color.upload(tape->color)) return rc;
        if (int rc = tape->d_jac_slot.upload(tape->jac_slot)) return rc;
        tape->j_uploaded = true;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Staged s;
    if (int rc = stage_in(*tape, x, batch, ld_x, vals, ld_vals, nnz, mem, stream, s)) return rc;
    int served = 0;  // a scalar function with many colours: the whole gradient in one reverse sweep per vector (from the second call on)
    if (int rc = launch_reverse(*tape, s.d_x, s.ld_x, batch, s.d_out, s.ld_out, int(nnz), stream, served)) return rc;
    if (!served) {
        const ub::tape::Seeds seeds{0, static_cast<const int*>(tape->d_color.ptr), nullptr, nullptr};
        if (int rc = launch<1>(*tape, seeds, s.d_x, s.ld_x, batch, tape->n_colors, s.d_out, s.ld_out, static_cast<const int*>(tape->d_jac_slot.ptr),
                               nullptr, stream))
            return rc;
    }
    return stage_out(s, vals, ld_vals, nnz, batch, mem, stream);
}

int ungar_b200_tape_sparse_hessian(ungar_b200_tape* tape, const double* x, const double* weights, int64_t batch, int64_t ld_x, double* vals,
                                   int64_t ld_vals, int32_t mem, void* stream_) {
    if (!tape) return tfail(UNGAR_B200_EINVAL, "null tape");
    if (!tape->h_set)
        if (int rc = choose_hessian(*tape, nullptr, nullptr, 0)) return rc;
    const int64_t nnz = int64_t(tape->h_rows.size());
    if (int rc = check_call(tape, x, batch, ld_x, vals, ld_vals, nnz, mem)) return rc;
    if (batch == 0 || nnz == 0) return UNGAR_B200_OK;
    if (int rc = ensure_device(*tape)) return rc;
    if (!tape->h_uploaded) {
        if (int rc = tape->d_pi.upload(tape->h_pi)) return rc;
        if (int rc = tape->d_pj.upload(tape->h_pj)) return rc;
        if (int rc = tape->d_di.upload(tape->h_di)) return rc;
        if (int rc = tape->d_dj.upload(tape->h_dj)) return rc;
        if (int rc = tape->d_pr.upload(tape->h_pr)) return rc;
        tape->h_uploaded = true;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    std::vector<double> w(size_t(tape->n_dep), 1.0);
    if (weights) w.assign(weights, weights + tape->n_dep);
    if (int rc = tape->d_w.reserve(std::max<size_t>(w.size(), 1) * sizeof(double))) return rc;
    UBT_CUDA(cudaMemcpyAsync(tape->d_w.ptr, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    UBT_CUDA(cudaStreamSynchronize(stream));  // `w` is a stack-lifetime staging buffer
    const int ndir = int(tape->h_pi.size());
    if (int rc = tape->ws_q.reserve(size_t(batch) * ndir * sizeof(double))) return rc;
    Staged s;
    if (int rc = stage_in(*tape, x, batch, ld_x, vals, ld_vals, nnz, mem, stream, s)) return rc;
    const ub::tape::Seeds seeds{1, nullptr, static_cast<const int*>(tape->d_pi.ptr), static_cast<const int*>(tape->d_pj.ptr)};
    if (int rc = launch<2>(*tape, seeds, s.d_x, s.ld_x, batch, ndir, static_cast<double*>(tape->ws_q.ptr), ndir, nullptr,
                           static_cast<const double*>(tape->d_w.ptr), stream))
        return rc;
    const long long total = (long long)batch * nnz;
    ub::tape::hessian_combine_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(
        static_cast<const double*>(tape->ws_q.ptr), ndir, static_cast<const int*>(tape->d_di.ptr), static_cast<const int*>(tape->d_dj.ptr),
        static_cast<const int*>(tape->d_pr.ptr), int(nnz), s.d_out, s.ld_out, batch);
    ub_count_launch();
    UBT_CUDA(cudaGetLastError());
    return stage_out(s, vals, ld_vals, nnz, batch, mem, stream);
}

}  // extern "C"
