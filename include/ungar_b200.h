/*
 * ungar_b200 — C ABI of the B200-native batched derivative-evaluation engine for Ungar NMPC problems.
 *
 * This is the drop-in boundary.  In the reference (fdevinc/ungar @ db9cc70) the derivative path sits
 * behind `CppAD::cg::GenericModel<double>` over a JIT-compiled, dlopen'ed C library
 * (include/ungar/autodiff/function.hpp:364-365, :433-438, :505-514).  Each entry point below names the
 * reference call it replaces.  The reference evaluates ONE flat vector xp = [x; p] per call on one CPU
 * core; every entry point here takes `batch` such vectors (stride `ld_xp` scalars) and evaluates all of
 * them on the GPU.  `batch = 1` with `UNGAR_B200_MEM_HOST` reproduces the reference call exactly.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - every function returns a status code (0 = success) and never throws; the message of the last
 *     failure on the calling thread is available from ungar_b200_last_error()
 *     (the reference aborts via UNGAR_ASSERT, include/ungar/assert.hpp:97-108, or throws
 *     std::runtime_error, function.hpp:531-534);
 *   - buffers are in the model's dtype (float for F32, double for F64), owned by the caller;
 *   - with UNGAR_B200_MEM_DEVICE the pointers are device pointers and the call is asynchronous on
 *     `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *     with UNGAR_B200_MEM_HOST the pointers are host pointers, the library stages them through its own
 *     device buffers (H2D, kernels, D2H on `stream`) and returns after the results have landed;
 *   - a model handle is immutable after creation except for its internal workspaces: calls on one handle
 *     must be issued from one thread at a time (the reference's Function is not re-entrant either:
 *     function.hpp:380-383).  Calls on one handle are ordered even across streams: its workspaces and work-claim
 *     counters are shared by every launch, so a call on another stream than the previous call's first waits (on the
 *     device, cudaStreamWaitEvent) for that call's work.  Use one handle per stream for concurrent pipelines.
 *   - there is NO CPU fallback: every compute entry point fails with UNGAR_B200_ECUDA when no CUDA
 *     device is usable.
 */
#ifndef UNGAR_B200_H_
#define UNGAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNGAR_B200_ABI_VERSION 2

typedef struct ungar_b200_model ungar_b200_model;

enum ungar_b200_status {
    UNGAR_B200_OK           = 0,
    UNGAR_B200_EINVAL       = 1, /* bad argument (sizes, null pointers, unknown enum) */
    UNGAR_B200_ECUDA        = 2, /* CUDA runtime error or no device */
    UNGAR_B200_ENOMEM       = 3,
    UNGAR_B200_EUNSUPPORTED = 4  /* e.g. Hessian of a vector-valued function (function.hpp:136-137) */
};

/* The three NMPC problems of example/mpc/{quadrotor,rc_car,quadruped}.example.cpp. */
enum ungar_b200_model_kind { UNGAR_B200_QUADROTOR = 0, UNGAR_B200_RC_CAR = 1, UNGAR_B200_QUADRUPED = 2 };

/* The reference computes in double only (include/ungar/data_types.hpp:89); F32 is an addition. */
enum ungar_b200_dtype { UNGAR_B200_F32 = 0, UNGAR_B200_F64 = 1 };

/* The functions an Ungar NLP problem is made of (include/ungar/optimization/concepts.hpp:153-161) plus
 * the barrier function SoftSQPOptimizer JIT-compiles for itself (optimization/soft_sqp.hpp:114-138):
 * Zsoft(z) = sum_i b(-z_i), independent size m_ineq, no parameters. */
enum ungar_b200_function {
    UNGAR_B200_OBJECTIVE         = 0,
    UNGAR_B200_EQUALITIES        = 1,
    UNGAR_B200_INEQUALITIES      = 2,
    UNGAR_B200_SOFT_INEQUALITIES = 3
};

enum ungar_b200_mem { UNGAR_B200_MEM_DEVICE = 0, UNGAR_B200_MEM_HOST = 1 };

/* Format of the per-trajectory KKT block record (ungar_b200_kkt_layout below).
 *   DENSE    every block dense, as SURVEY.md §8d counts them (A_k 13x37, H_k packed 37x37, C_k 4x4x20 ...): ~65 % of the bytes are
 *            structural zeros for the quadruped.
 *   COMPACT  only the structurally non-zero slots (quadruped.example.cpp:162-200, :216-244, :269-303), one contiguous 16-byte aligned
 *            chunk per shooting node: 2.5x fewer bytes to write, to read back in the QP solve and to copy to the host.  Quadruped
 *            only (UNGAR_B200_EUNSUPPORTED otherwise).  ungar_b200_kkt_compact_map gives, for every compact slot, its offset in the
 *            dense record, so dense-from-compact is one scatter.  Records passed in must be 16-byte aligned with an even stride. */
enum ungar_b200_record_format { UNGAR_B200_RECORD_DENSE = 0, UNGAR_B200_RECORD_COMPACT = 1 };

typedef struct ungar_b200_model_desc {
    int32_t kind;    /* ungar_b200_model_kind */
    int32_t horizon; /* N; the reference hard-codes 30 (quadrotor.example.cpp:52) */
    int32_t dtype;   /* ungar_b200_dtype */
    int32_t device;  /* CUDA device ordinal */
    /* RelaxedPolyBarrierFunction{0, stiffness, epsilon} (optimization/soft_inequality_constraint.hpp:133-145);
     * per example: quadrotor (100, 2e-5), rc_car (100, 1e-2), quadruped (1, 1). */
    double barrier_stiffness;
    double barrier_epsilon;
    int32_t record_format; /* ungar_b200_record_format: the format of every `records` argument of this handle */
    int32_t reserved;      /* 0 */
} ungar_b200_model_desc;

/* Replaces FunctionFactory::Make / MakeFunction (function.hpp:589-613): where the reference tapes,
 * generates C, runs gcc and dlopens, this selects the hand-written sm_100a kernels of the model,
 * derives the structural sparsity of its three functions and allocates device-side index tables. */
int ungar_b200_model_create(const ungar_b200_model_desc* desc, ungar_b200_model** out);
int ungar_b200_model_destroy(ungar_b200_model* model);

/* Function::IndependentVariableSize / ParameterSize / DependentVariableSize (function.hpp:351-361) and
 * the nnz counts the Function constructor derives (function.hpp:98-101, :138-140). */
int ungar_b200_function_info(const ungar_b200_model* model, int32_t function, int64_t* independent_size,
                             int64_t* parameter_size, int64_t* dependent_size, int64_t* nnz_jacobian,
                             int64_t* nnz_hessian);

/* GenericModel::JacobianSparsity(rows, cols) (function.hpp:103-105): row-major, rows ascending, columns
 * ascending within a row, parameter columns trimmed (function.hpp:529-550).  The arrays belong to the
 * handle. */
int ungar_b200_jacobian_sparsity(const ungar_b200_model* model, int32_t function, const int64_t** rows,
                                 const int64_t** cols, int64_t* nnz);

/* GenericModel::HessianSparsity(0, rows, cols) (function.hpp:142-144): upper triangle of the x-x block
 * (function.hpp:552-574); scalar functions only. */
int ungar_b200_hessian_sparsity(const ungar_b200_model* model, int32_t function, const int64_t** rows,
                                const int64_t** cols, int64_t* nnz);

/* GenericModel::ForwardZero (function.hpp:186-189): y[b, :] = f(xp[b, :]). */
int ungar_b200_forward_zero(ungar_b200_model* model, int32_t function, const void* xp, int64_t batch,
                            int64_t ld_xp, void* y, int64_t ld_y, int32_t mem, void* stream);

/* GenericModel::SparseJacobian (function.hpp:224-228): vals[b, :] = nonzeros of df/dx at xp[b, :] in the
 * order of ungar_b200_jacobian_sparsity. */
int ungar_b200_sparse_jacobian(ungar_b200_model* model, int32_t function, const void* xp, int64_t batch,
                               int64_t ld_xp, void* vals, int64_t ld_vals, int32_t mem, void* stream);

/* GenericModel::SparseHessian with a unit weight on dependent 0 (function.hpp:249-257): vals[b, :] =
 * nonzeros of the upper-triangular Hessian in the order of ungar_b200_hessian_sparsity. */
int ungar_b200_sparse_hessian(ungar_b200_model* model, int32_t function, const void* xp, int64_t batch,
                              int64_t ld_xp, void* vals, int64_t ld_vals, int32_t mem, void* stream);

/* Element offsets of the per-trajectory KKT block record written by ungar_b200_kkt_blocks
 * (every array starts on a multiple of 4 elements; `size` is the record length = minimum ld_rec).
 *   g    [m_eq]                 equality residuals, reference row order
 *   A    [N][nx][nz]            A_k = d g_{dyn,k} / d [x_k; u_k]  (= -df/dz; d/dx_{k+1} = I is implicit)
 *   C    [N][legs][4][20]       quadruped contact rows wrt [p_k q_k r_{k,i} | p_{k-1} q_{k-1} r_{k-1,i}]
 *   h    [m_ineq]               inequality values
 *   cost [2]                    objective f, barrier Zsoft(h)
 *   grad [n_dec]                QP vector  q = grad f + J_h^T dZsoft        (soft_sqp.hpp:151-153)
 *   H    [N][nz(nz+1)/2]        upper triangle (row-major packed) of the z_k-z_k block of
 *                               P = grad^2 f + J_h^T d2Zsoft J_h + 1e-6 I   (soft_sqp.hpp:145-150)
 *   HN   [nx(nx+1)/2]           the same for the terminal state x_N
 *   Hc   [N-1][nu]              diagonal of the u_{k}-u_{k+1} coupling block (quadrotor, rc_car) */
typedef struct ungar_b200_kkt_layout {
    int64_t g, A, C, h, cost, grad, H, HN, Hc, size; /* block offsets of the DENSE arrangement; `size` = record length in the handle's format */
    int64_t nx, nu, nz, horizon, n_dec, n_par, m_eq, m_ineq, tri, tri_terminal, legs, hc_per_node;
    /* round 2 (ABI 2) */
    int64_t compact;    /* 1 if the handle's records are COMPACT */
    int64_t dense_size; /* length of the dense arrangement (= size for DENSE handles) */
    /* COMPACT records: chunk k starts at k * node_stride; offsets inside a chunk; `tail` after the N chunks (csrc/compact.cuh):
     *   Cs [legs][4][8]  contact rows wrt (p_c, q, r_leg) of node k        Cp [legs][3][8]  rows 1..3 wrt the same columns of node k-1
     *   g  [29] defects | contact values     q [37] QP vector of [x_k; u_k]     Hd [13] state diagonal    Hb [8][6] 3x3 input blocks
     *   h  [12]                              AQ [7][32] q+ / w+ rows wrt (q, w, u)                  AP [6][6] p+ / v+ rows
     *   tail: g0 [13] (x_0 - x_measured) | qN [13] | HN diagonal [13] | cost [2]   (at tail + t_*) */
    int64_t node_stride, c_Cs, c_Cp, c_g, c_q, c_Hd, c_Hb, c_h, c_AQ, c_AP, tail, t_g0, t_qN, t_HN, t_cost;
} ungar_b200_kkt_layout;

int ungar_b200_kkt_layout_get(const ungar_b200_model* model, ungar_b200_kkt_layout* out);

/* COMPACT handles: map[e] = offset in the dense arrangement of compact slot e (e < layout.size), or -2 for a pad slot (always 0).
 * The array belongs to the handle.  UNGAR_B200_EUNSUPPORTED for DENSE handles. */
int ungar_b200_kkt_compact_map(const ungar_b200_model* model, const int32_t** map, int64_t* count);

/* Replaces one pass of SoftSQPOptimizer::AssembleOSQPInstance (soft_sqp.hpp:141-158, :236-264): every
 * value, Jacobian block and Gauss-Newton Hessian block of every shooting node of every trajectory, in one
 * sweep: records[b, :] laid out as ungar_b200_kkt_layout. */
int ungar_b200_kkt_blocks(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp,
                          void* records, int64_t ld_rec, int32_t mem, void* stream);

/* The "Jacobian sweep" alone (BASELINE.json configs[1]): after the call the blocks `g` (equality residuals) and `A`
 * (dynamics Jacobians; for the quadruped also `C`) of every record are those ungar_b200_kkt_blocks would write; the other
 * blocks of the record are unspecified.  Quadrotor and RC car skip the objective / inequality / Gauss-Newton work and
 * the `H` stores (about 40 % of the bytes); the quadruped runs the full sweep.  Replaces Function::Evaluate + Function::Jacobian
 * of the equality constraints (function.hpp:180-230) for every shooting interval. */
int ungar_b200_jacobian_blocks(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp, void* records,
                               int64_t ld_rec, int32_t mem, void* stream);

/* Per-trajectory summary (32 scalars: u_0 [nu<=24], cost f, barrier, |g|_inf, max h, pad) read back from
 * the records — the payload of the one all-gather per outer iteration. */
#define UNGAR_B200_SUMMARY_SIZE 32
int ungar_b200_summaries(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp,
                         const void* records, int64_t ld_rec, void* summaries, void* stream);

/* One outer-iteration step of the B200-native data flow: inputs `xp` (host or device per `mem`), the KKT
 * records stay resident in HBM (`records_device`, a DEVICE pointer, or NULL to use the handle's workspace)
 * for the on-device consumers, and only the per-trajectory summaries go back (`summaries`, host or device per
 * `mem`).  With UNGAR_B200_MEM_HOST the call returns after the summaries have landed. */
int ungar_b200_kkt_step(ungar_b200_model* model, const void* xp, int64_t batch, int64_t ld_xp, void* records_device,
                        int64_t ld_rec, void* summaries, int32_t mem, void* stream);

/* The MPC data flow splits the flat vector: the parameter block of xp (references, constants, measured state — Ungar's `parameters`,
 * quadruped.example.cpp:94-139) changes once per control cycle, the decision variables [X | U] every outer iteration.
 * ungar_b200_set_parameters copies parameters[b, 0:n_par] (host or device per `mem`) into a device-resident copy of xp owned by the
 * handle; ungar_b200_kkt_step_x is ungar_b200_kkt_step taking only the decision variables x[b, 0:n_dec] (45 % fewer bytes over PCIe
 * for the quadruped).  `batch` must equal the batch of the last ungar_b200_set_parameters. */
int ungar_b200_set_parameters(ungar_b200_model* model, const void* parameters, int64_t batch, int64_t ld_par, int32_t mem, void* stream);
int ungar_b200_kkt_step_x(ungar_b200_model* model, const void* x, int64_t batch, int64_t ld_x, void* records_device, int64_t ld_rec,
                          void* summaries, int32_t mem, void* stream);

/* Replaces the QP solve SoftSQPOptimizer delegates to OSQP (optimization/soft_sqp.hpp:193-233) for the equality-
 * constrained QP assembled by ungar_b200_kkt_blocks:  min 1/2 d^T P d + q^T d  s.t.  A d = -g.  Consumes the records in
 * place (DEVICE pointers), writes the step `steps[b, 0:n_dec]` (the reference's `primal_solution()`, [X | U] order) and,
 * if `multipliers` is not NULL, the equality multipliers `multipliers[b, 0:m_eq]` in the reference's row order.
 * Exact stage-wise factorisations, not ADMM: quadrotor and RC car by a Riccati recursion (the input-rate coupling of
 * the objective is carried by augmenting the state with the previous input), quadruped by a block-tridiagonal Schur
 * complement on the multipliers (its contact rows are extra equalities).  The factorisation is always fp64 (the
 * reference computes in double, data_types.hpp:89): an F32 model hands its fp32 record to it through a twin F64 handle
 * (record widened, step and multipliers narrowed, on the device). */
int ungar_b200_qp_solve(ungar_b200_model* model, const void* records_device, int64_t batch, int64_t ld_rec, void* steps,
                        int64_t ld_steps, void* multipliers, int64_t ld_multipliers, void* stream);

/* Options of the outer loop: SoftSQPOptimizer's constructor arguments (optimization/soft_sqp.hpp:44-50; the barrier
 * stiffness / epsilon belong to the model descriptor) and BacktrackingLineSearch::Parameters
 * (optimization/backtracking_line_search.hpp:57-76).  ungar_b200_sqp_options_default fills the reference defaults:
 * max_iterations 10, multiplier 1, alpha_min 1e-4, theta_min 1e-6, theta_max 1e-2, eta 1e-4, gamma_phi 1e-6,
 * gamma_theta 1e-6, gamma_alpha 0.5, objective_tolerance 1e-6 (the literal of soft_sqp.hpp:103). */
typedef struct ungar_b200_sqp_options {
    int32_t max_iterations;
    int32_t reserved;
    double constraint_violation_multiplier;
    double alpha_min, theta_min, theta_max, eta, gamma_phi, gamma_theta, gamma_alpha;
    double objective_tolerance;
} ungar_b200_sqp_options;

int ungar_b200_sqp_options_default(ungar_b200_sqp_options* out);

/* Per-trajectory state of the outer loop, two int32 per trajectory: { status, iterations started }. */
enum ungar_b200_sqp_status {
    UNGAR_B200_SQP_RUNNING            = 0, /* after ungar_b200_sqp_solve: stopped by max_iterations */
    UNGAR_B200_SQP_CONVERGED          = 1, /* objective decreased by less than objective_tolerance (soft_sqp.hpp:101-108) */
    UNGAR_B200_SQP_LINE_SEARCH_FAILED = 2  /* no step size >= alpha_min accepted; the iterate is unchanged (:100-102) */
};
/* Per-trajectory line-search report, 8 scalars of the model dtype:
 * { alpha (0 = rejected), theta, phi, f after the step | theta, phi, f before it, grad f . dw }. */
#define UNGAR_B200_LINE_SEARCH_INFO_SIZE 8

/* Replaces BacktrackingLineSearch::Do (optimization/backtracking_line_search.hpp:81-165) with the merit functions
 * SoftSQPOptimizer::Optimize passes (optimization/soft_sqp.hpp:81-99): cost phi = f + Zsoft(h), constraint violation
 * theta = multiplier * |g|_2, both evaluated on the device at every trial point w + alpha dw.  All DEVICE pointers.
 * On acceptance xp[b, 0:n_dec] += alpha * steps[b, :] in place.  `status` (int32 [batch][2], may be NULL) skips the
 * trajectories whose status is not RUNNING and receives the bookkeeping of soft_sqp.hpp:100-108; `info`
 * ([batch][8], may be NULL) receives the report above.  The search always evaluates in fp64 (its acceptance tests compare
 * relative changes of 1e-6): an F32 model widens `xp` and `steps` on the device, searches on a twin F64 handle and narrows the
 * accepted iterate and the report. */
int ungar_b200_line_search(ungar_b200_model* model, void* xp, int64_t batch, int64_t ld_xp, const void* steps,
                           int64_t ld_steps, const ungar_b200_sqp_options* options, int32_t* status, void* info,
                           void* stream);

/* Replaces SoftSQPOptimizer::Optimize (optimization/soft_sqp.hpp:63-109) for a batch of independent problems:
 * max_iterations times { KKT sweep -> QP solve -> line search }, entirely on the device, no host round trip inside
 * the loop.  `xp` is updated in place (the reference returns _cache.xp.head(n_dec)); `status` (int32 [batch][2])
 * and `info` ([batch][8], may be NULL; report of the last line search) are host or device buffers per `mem`.
 * Trajectories that stop early are frozen exactly where the reference's loop breaks.  The loop computes in
 * fp64; an F32 model (BASELINE configs 2 and 3) widens the iterate on the device, runs it on a twin F64 handle and narrows the result. */
int ungar_b200_sqp_solve(ungar_b200_model* model, void* xp, int64_t batch, int64_t ld_xp,
                         const ungar_b200_sqp_options* options, int32_t* status, void* info, int32_t mem, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Generic path: a recorded operation tape of an arbitrary user lambda, evaluated on the GPU.
 *
 * This is the seam of the reference's own plugin boundary: Ungar::Autodiff::Function calls a
 * CppAD::cg::GenericModel<double> loaded from a JIT-compiled library (autodiff/function.hpp:364-365, :433-438,
 * :505-514).  ungar_b200/include/cppad/cg.hpp supplies that class (and the tracing AD scalar that records the
 * tape) on top of the entry points below, so the reference's unchanged function.hpp / tests / examples evaluate
 * their lambdas on the device.  Where the reference tapes -> generates C -> runs gcc -> dlopens
 * (function.hpp:453-503), a tape handle is analysed once on the host (dead-node elimination, slot allocation by
 * liveness, structural sparsity by dependency propagation, column colouring) and then interpreted by a register
 * machine kernel, one thread per (xp vector, seed direction).
 *
 * Node encoding: node i may only reference nodes < i.  INDEP: a = index of the independent.  CONST: k = value.
 * Unary ops: a.  Binary ops: a, b (POW: a^b, ATAN2: atan2(a, b)).  Conditionals CppAD::CondExp{Lt,Le,Gt,Ge,Eq}(a, b,
 * c, d): c if a OP b else d.  `dependents[r]` = node id of dependent r, or -1 for a constant dependent with value
 * `dependent_constants[r]`. */
enum ungar_b200_tape_op {
    UNGAR_B200_OP_INDEP = 0, UNGAR_B200_OP_CONST, UNGAR_B200_OP_ADD, UNGAR_B200_OP_SUB, UNGAR_B200_OP_MUL, UNGAR_B200_OP_DIV,
    UNGAR_B200_OP_NEG, UNGAR_B200_OP_SQRT, UNGAR_B200_OP_SIN, UNGAR_B200_OP_COS, UNGAR_B200_OP_TAN, UNGAR_B200_OP_ATAN,
    UNGAR_B200_OP_ACOS, UNGAR_B200_OP_ASIN, UNGAR_B200_OP_EXP, UNGAR_B200_OP_LOG, UNGAR_B200_OP_ABS, UNGAR_B200_OP_POW,
    UNGAR_B200_OP_ATAN2, UNGAR_B200_OP_CLT, UNGAR_B200_OP_CLE, UNGAR_B200_OP_CGT, UNGAR_B200_OP_CGE, UNGAR_B200_OP_CEQ,
    UNGAR_B200_OP_COUNT
};

typedef struct ungar_b200_tape_node {
    uint8_t op; /* ungar_b200_tape_op */
    int32_t a, b, c, d;
    double k;
} ungar_b200_tape_node;

typedef struct ungar_b200_tape ungar_b200_tape;

/* CppAD::ADFun + ModelCSourceGen + compile + dlopen (function.hpp:456-503): host-side analysis only, no device work
 * (the program is uploaded on first evaluation), so sizes and sparsity are available without a GPU. */
int ungar_b200_tape_create(const ungar_b200_tape_node* nodes, int64_t n_nodes, int64_t n_independent,
                           const int32_t* dependents, const double* dependent_constants, int64_t n_dependent,
                           int32_t device, ungar_b200_tape** out);
int ungar_b200_tape_destroy(ungar_b200_tape* tape);
/* Diagnostics of the NVRTC-specialised kernels (csrc/tape.cu): from the second evaluation of an order on, a tape of up to 12 000
 * instructions runs as ONE straight-line sm_100a kernel (a longer one, up to 100 000, as a sequence of kernels of 6 000 instructions)
 * compiled with NVRTC and cached on disk under a CONTENT hash (instruction
 * stream, constants, order, arch, the text of csrc/tape_machine.cuh) — the reference caches its generated library by NAME only
 * (function.hpp:420-451).  info[4 * order + {0, 1, 2, 3}] (16 entries), order 0..2 and 3 = the reverse sweep that serves the gradient
 * of a SCALAR function with 8 or more colours in one pass per vector = {state: 0 not tried / 1 specialised / -1 interpreter / 2 compiling,
 * served from the cache, low 32 bits of the hash, high 32 bits}.  UNGAR_B200_KERNEL_CACHE names the cache directory,
 * UNGAR_B200_NO_NVRTC=1 keeps the interpreter.  The reverse sweep is opt-in (UNGAR_B200_REVERSE=1): validated on the CPU by replaying its
 * generated text, not yet run on a GPU. */
int ungar_b200_tape_special_info(const ungar_b200_tape* tape, int64_t* info);
/* The kernels of a tape beyond 12 000 instructions take tens of seconds to compile: a worker thread does it (state 2 in
 * ungar_b200_tape_special_info) while the interpreter keeps serving the calls.  This call blocks until no compile is in flight.
 * UNGAR_B200_NVRTC_SYNC=1 compiles in the calling thread instead. */
int ungar_b200_tape_special_wait(ungar_b200_tape* tape);
/* The CUDA source the NVRTC path generates for `order` (0 values, 1 Jacobian, 2 Hessian jets, 3 the reverse sweep of a scalar
 * function, kernel `tape_reverse`): ONE kernel `tape_special` for tapes of
 * up to 12 000 instructions, otherwise kernels `tape_part_<k>` of 6 000 instructions each whose cross-kernel values travel through the
 * scratch array.  Host-only (no device, no compile): `buffer` receives at most `capacity` bytes including the terminating 0,
 * `*required` the size of the whole text, `*n_kernels` the number of kernels.  Diagnostics / tests. */
int ungar_b200_tape_kernel_source(const ungar_b200_tape* tape, int32_t order, char* buffer, int64_t capacity, int64_t* required,
                                  int32_t* n_kernels);
/* info[6] = independents, dependents, nodes kept after dead-node elimination, scratch slots, Jacobian colours,
 * Hessian directions (the last two are 0 until the element sets below are chosen). */
int ungar_b200_tape_info(const ungar_b200_tape* tape, int64_t* info);

/* GenericModel::JacobianSparsitySet (function.hpp:531): structural pattern over ALL independents [x; p], row-major,
 * columns ascending.  GenericModel::HessianSparsitySet (:559): full symmetric pattern, union over the dependents. */
int ungar_b200_tape_jacobian_pattern(ungar_b200_tape* tape, const int64_t** rows, const int64_t** cols, int64_t* nnz);
int ungar_b200_tape_hessian_pattern(ungar_b200_tape* tape, const int64_t** rows, const int64_t** cols, int64_t* nnz);

/* ModelCSourceGen::setCustomSparseJacobianElements / setCustomSparseHessianElements (function.hpp:549, :573): the
 * elements the evaluation calls return, in this order (the reference passes the structural pattern with the
 * parameter columns trimmed, and for the Hessian only the upper triangle).  Every element must be structurally
 * non-zero.  NULL rows selects the whole structural pattern. */
int ungar_b200_tape_set_jacobian_elements(ungar_b200_tape* tape, const int64_t* rows, const int64_t* cols, int64_t nnz);
int ungar_b200_tape_set_hessian_elements(ungar_b200_tape* tape, const int64_t* rows, const int64_t* cols, int64_t nnz);

/* GenericModel::ForwardZero / SparseJacobian / SparseHessian (function.hpp:186-189, :224-228, :252-257) for `batch`
 * vectors x[b, 0:n_independent] (stride ld_x); F64; host or device buffers per `mem` like every other entry point.
 * `weights` (HOST pointer, n_dependent values; NULL = all ones): the Hessian is that of sum_r weights[r] * y_r. */
int ungar_b200_tape_forward_zero(ungar_b200_tape* tape, const double* x, int64_t batch, int64_t ld_x, double* y,
                                 int64_t ld_y, int32_t mem, void* stream);
int ungar_b200_tape_sparse_jacobian(ungar_b200_tape* tape, const double* x, int64_t batch, int64_t ld_x, double* vals,
                                    int64_t ld_vals, int32_t mem, void* stream);
int ungar_b200_tape_sparse_hessian(ungar_b200_tape* tape, const double* x, const double* weights, int64_t batch,
                                   int64_t ld_x, double* vals, int64_t ld_vals, int32_t mem, void* stream);

/* Replaces osqp::OsqpSolver::Solve (optimization/soft_sqp.hpp:226) for ANY equality-constrained QP — the only kind
 * SoftSQPOptimizer poses (l = u = -g, soft_sqp.hpp:155-157) — when the problem has no stage-wise solver:
 *     min 1/2 x^T P x + q^T x   s.t.  A x = b
 * P (n x n, upper triangle used, like OSQP) and A (m x n) in compressed-sparse-column form, HOST arrays; x[n] and y[m]
 * (multipliers, may be NULL) HOST outputs.  One dense LU of the quasi-definite KKT matrix [P + sigma I, A^T; A, -rho I]
 * on the device (scatter kernel + cuSOLVER getrf/getrs); n + m <= 16384.  This is the back end of the osqp++.h
 * stand-in in ungar_b200/include, which lets the reference's unchanged SoftSQPOptimizer run on the GPU. */
int ungar_b200_kkt_solve_csc(int64_t n, int64_t m, const int32_t* P_colptr, const int32_t* P_rowidx, const double* P_vals,
                             const double* q, const int32_t* A_colptr, const int32_t* A_rowidx, const double* A_vals,
                             const double* b, double sigma, double rho, double* x, double* y, int32_t device);

/* Device-side timing of the dominant kernel (the KKT sweep): when enabled, every sweep launch is bracketed by
 * CUDA events on the launching stream; ungar_b200_sweep_times synchronises and returns up to `cap` most recent
 * durations in milliseconds (oldest first) and clears the ring. */
int ungar_b200_set_profiling(int32_t enabled);
int ungar_b200_sweep_times(float* ms, int32_t cap, int32_t* count);

/* Number of kernel launches this library has issued in the calling process (bench accounting). */
int64_t ungar_b200_launch_count(void);

const char* ungar_b200_last_error(void);
int32_t ungar_b200_abi_version(void);

#ifdef __cplusplus
}
#endif

#endif /* UNGAR_B200_H_ */
