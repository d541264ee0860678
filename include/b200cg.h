/* b200cg.h - C ABI of libb200cg.so: the B200-native (sm_100a, fp64) conjugate-gradient Dirichlet-Poisson
 * hot path of Ruslan361/iterative_solvers.
 *
 * The reference has no FFI: its boundary is the public C++ surface of the static library `dirichlet_solver`
 * (solver/CMakeLists.txt:65). The drop-in C++ classes in iterative_solvers_b200/dropin/ keep those
 * signatures and route their bodies through the entry points below; each entry point names the reference
 * interface it replaces (paths relative to the reference root). INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes only; every function returns a b200cg_status (0 = ok) and never
 * throws; b200cg_last_error() returns a thread-local message for the last failure on the calling thread.
 * All host vectors are fp64 in the reference's compact unknown ordering (grid_system.cpp:84-111: block B =
 * bottom-right rows, then block U = upper rows; RECT domain: row-major). For a sharded plan (world > 1) host
 * vectors hold only this rank's contiguous index range [lo, hi) (b200cg_local_range).
 * There is no CPU fallback: without a usable CUDA device every compute call fails with B200CG_ERR_NO_DEVICE.
 */
#ifndef B200CG_H
#define B200CG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CG_VERSION 100

typedef enum {
  B200CG_OK = 0,
  B200CG_ERR_INVALID_ARG = 1,
  B200CG_ERR_NO_DEVICE = 2,
  B200CG_ERR_CUDA = 3,
  B200CG_ERR_STATE = 4,    /* call order: e.g. solve with rhs_on_device before any rhs was set */
  B200CG_ERR_COMM = 5,     /* NCCL bootstrap / collective failure */
  B200CG_ERR_UNSUPPORTED = 6
} b200cg_status;

typedef enum {
  B200CG_DOMAIN_LSHAPE = 0, /* the reference's L-shaped region: rectangle minus its lower-left quadrant
                               (grid_system.cpp:17-43); requires even n == m (the reference numbering is
                               only self-consistent there, grid_system.cpp:103-111) */
  B200CG_DOMAIN_RECT = 1,   /* full rectangle, any n, m >= 2 (no reference counterpart; synthetic configs) */
  B200CG_DOMAIN_GENERIC = 2, /* no geometry: a plan of generic_rows unknowns that only serves the CSR entry points
                               (MSGSolver receives just a matrix and a rhs, msg_solver.hpp:50-53) */
  B200CG_DOMAIN_LSHAPE_ANY = 3 /* the L-shaped region for any n, m >= 4 with the reference's numbering defects repaired
                               (block B is n-1-n/2 wide, block U counted from m/2): identical to LSHAPE for even
                               n == m, defined where the reference builds a malformed system */
} b200cg_domain;

typedef enum {
  B200CG_OP_MATRIX_FREE = 0, /* on-the-fly 5-point stencil: MatrixFreeSystem::apply, matrix_free_system.cpp:203-340 */
  B200CG_OP_CSR = 1          /* assembled CSR SpMV: KokkosSparse::spmv at msg_solver.cpp:93 */
} b200cg_operator;

typedef enum {
  B200CG_RULE_REL_L2 = 0,  /* MatrixFreeSolver: loop while it < max_it && ||r||_2 > eps_rel*||r0||_2
                              (matrix_free_system.cpp:409); alpha = r.r/p.Ap, beta = r'.r'/r.r */
  B200CG_RULE_MAXNORM = 1  /* MSGSolver: max-norm rules checked every iteration in the order precision,
                              residual, exact error (msg_solver.cpp:144-162); alpha = r.z/Az.z,
                              beta = (||r'||_2)^2 / r.z (msg_solver.cpp:96-102,165) */
} b200cg_rule;

typedef enum {
  B200CG_PRECOND_NONE = 0,      /* the reference's iteration */
  B200CG_PRECOND_MULTIGRID = 1  /* V(2,2) cycle, damped Jacobi, full weighting / bilinear, rediscretised coarse operators */
} b200cg_preconditioner;

/* Same values as the reference's enum class StopCriterion (msg_solver.hpp:9-15). */
typedef enum {
  B200CG_STOP_ITERATIONS = 0,
  B200CG_STOP_PRECISION = 1,
  B200CG_STOP_RESIDUAL = 2,
  B200CG_STOP_EXACT_ERROR = 3,
  B200CG_STOP_INTERRUPTED = 4
} b200cg_stop;

typedef struct b200cg_plan_s* b200cg_plan_t;

/* Geometry + placement. (n, m, a..d) are GridSystem/MatrixFreeSystem's constructor arguments
 * (grid_system.cpp:301-318; note those constructors take (m, n, ...)): n, m = numbers of intervals. */
typedef struct {
  int n, m;
  double a, b, c, d;
  int domain;            /* b200cg_domain */
  int device;            /* CUDA device ordinal of this process */
  int rank, world;       /* row-slab sharding, one process per GPU; world <= 1: single GPU */
  const void* comm_id;   /* world > 1: the 128-byte id from b200cg_comm_unique_id() of rank 0 */
  int tile_rows;         /* rows a CTA marches per tile; 0 = default */
  int reserved0;
  int64_t generic_rows;  /* B200CG_DOMAIN_GENERIC: number of unknowns (n, m, a..d are ignored) */
  int reserved[4];
} b200cg_plan_desc;

typedef struct {
  int op;                  /* b200cg_operator */
  int rule;                /* b200cg_rule */
  double eps_rel;          /* RULE_REL_L2 (MatrixFreeSolver's eps, matrix_free_system.hpp:100) */
  double eps_p, eps_r, eps_e; /* RULE_MAXNORM; <= 0 disables a rule (dirichlet_solver.cpp:71-87 passes -1) */
  int max_it;
  int callback_every;      /* RULE_MAXNORM: callbacks at it 0, 1, every callback_every (reference: 100,
                              msg_solver.cpp:172) and the final one. RULE_REL_L2: a registered callback fires
                              every iteration with (||dx||_2, recomputed ||b-Ax||_2, ||x-u||_2)
                              (matrix_free_system.cpp:444-468). Ignored when cb == NULL */
  int rhs_on_device;       /* 1: use the rhs already resident in the plan (b200cg_build_rhs / _set_rhs);
                              b_host is ignored and no H2D copy happens */
  int keep_x_on_device;    /* 1: skip the D2H copy of the solution (x_host may be NULL) */
  int iters_per_graph;     /* CG iterations captured per CUDA-graph launch; 0 = default */
  int small_grid_path;     /* grids that fit the shared memory of one thread-block cluster (about 300^2) are solved by
                              a single cluster-resident kernel instead of the graph loop: 0 = automatic, 1 = never,
                              2 = require it (B200CG_ERR_UNSUPPORTED if the grid does not fit) */
  int single_sweep;        /* matrix-free operator: each iteration runs as ONE sweep by forming alpha from the single-reduction CG
                              recurrence (Chronopoulos-Gear) instead of p.Ap - the same iterates in exact arithmetic, <= 4e-14
                              relative apart in fp64 on the reference's grids (tests/studies/single_reduction_cg.py), same
                              iteration counts. RULE_REL_L2 without a callback: 40 instead of 56 bytes per unknown-iteration.
                              RULE_MAXNORM (MSGSolver's rules, callbacks included): 48 (56 with u) instead of 64 (72) bytes, r.z
                              replaced by r.r of the same residual. On sharded plans it needs the peer-memory exchange and >= 4
                              rows per rank. 0 = the plan's default (on unless B200CG_SINGLE_SWEEP=0), 1 = on, 2 = never (dot
                              sweep + update sweep with alpha = r.r / p.Ap resp. r.z / Az.z). Ignored where it does not apply
                              (assembled operator, per-iteration report, NCCL-fallback exchange) */
  int preconditioner;      /* b200cg_preconditioner. B200CG_PRECOND_MULTIGRID (opt-in; the reference has no preconditioner,
                              solver.hpp:17-66 is the base class kept for one): CG preconditioned by a geometric-multigrid
                              V-cycle - matrix-free operator, RULE_REL_L2, no callback, single-GPU plan; the iteration
                              count no longer grows with n (7 instead of ~2.7 n). Same x0, same stop rule; the iterates
                              are NOT the reference's (a different Krylov space), only the solution agrees */
  int reserved[4];
} b200cg_params;

typedef struct {
  int iterations;          /* completed iterations (getIterations()) */
  int converged;
  int stop_reason;         /* b200cg_stop */
  double r0_l2, r_l2;      /* ||r0||_2 and the recurrence ||r||_2 at exit */
  /* the three max-norms feed the MAXNORM stop rules; under B200CG_RULE_REL_L2 r_max and dx_max may be left 0 */
  double r_max;            /* ||r||_inf (recurrence) - MSGSolver::getFinalResidualNorm */
  double dx_max;           /* ||x_n - x_{n-1}||_inf  - MSGSolver::getFinalPrecision */
  double err_max;          /* ||x - u||_inf          - MSGSolver::getFinalErrorNorm (DBL_MAX without u, and under
                              B200CG_RULE_REL_L2 without a callback, where nothing tracks it) */
  double total_ms;         /* host wall time of the whole call, copies included */
  double solve_ms;         /* device time of the iterations (CUDA events on the solve stream) */
  double device_ms;        /* device time of the whole call on the solve stream: H2D, init, iterations, D2H */
  double h2d_ms, d2h_ms;
  int64_t h2d_bytes, d2h_bytes;
  int64_t kernel_launches; /* kernels of this library launched by the call */
  double dot_kernel_ms;    /* average duration of the sampled dot-phase / update-phase kernels of this call */
  double upd_kernel_ms;    /* (CUDA event nodes around the first iteration of every graph launch) */
  int kernel_samples;
  int64_t local_unknowns;  /* unknowns owned by this rank */
  double upd_even_ms;      /* sampled update-phase kernel of even / odd iterations: with x-deferral (REL_L2 rule, */
  double upd_odd_ms;       /* no report) even iterations skip x (32 B/unknown) and odd ones carry both (48 B)    */
  int x_deferral;          /* 1 if this solve touched x only every other iteration */
  int cluster_path;        /* 1 if this solve ran as one cluster-resident kernel (small_grid_path) */
  int peer_exchange;       /* sharded plans: 1 if halo rows and reductions went over NVLink peer memory (CUDA IPC),
                              0 if over NCCL send/recv + all-reduce */
  int single_sweep;        /* 1 if this solve ran the single-sweep iteration (b200cg_params.single_sweep), under either rule */
  int preconditioner;      /* b200cg_preconditioner this solve ran with */
  int mg_levels;           /* multigrid levels used (0 without the preconditioner) */
} b200cg_info;

/* (iteration, precision, residual, error) - Solver::setIterationCallback, solver.hpp:46-50. Called on the
 * solving thread (mainwindow.cpp:49-52 re-emits it as a queued signal). */
typedef void (*b200cg_iter_cb)(void* user, int iteration, double precision, double residual, double error);

/* ------------------------------------------------------------------ library */
const char* b200cg_last_error(void);
int b200cg_version(void);
int b200cg_device_count(int* count);
/* Pinned host memory for callers that want full-rate H2D/D2H (plain malloc'd pointers work too). */
int b200cg_alloc_pinned(void** ptr, size_t bytes);
int b200cg_free_pinned(void* ptr);
/* world > 1 bootstrap: rank 0 obtains 128 opaque bytes and ships them to every rank out of band. */
int b200cg_comm_unique_id(void* id128);

/* ------------------------------------------------------------------ plan = geometry + device buffers
 * replaces: GridSystem::GridSystem (grid_system.cpp:301-322), MatrixFreeSystem::MatrixFreeSystem
 * (matrix_free_system.cpp:144-159) */
int b200cg_plan_create(b200cg_plan_t* plan, const b200cg_plan_desc* desc);
int b200cg_plan_destroy(b200cg_plan_t plan);
/* number of unknowns of the whole system: MatrixFreeSystem::size (matrix_free_system.hpp:66),
 * calculate_position_in_template(n-1, m-1) + 1 (grid_system.cpp:162) */
int b200cg_size(b200cg_plan_t plan, int64_t* n_unknowns);
int b200cg_local_range(b200cg_plan_t plan, int64_t* lo, int64_t* hi);
/* The row-slab partition a plan with this descriptor would use (pure geometry, needs no device): unknown rows
 * [y_lo, y_hi) and compact index range [lo, hi) of desc->rank among desc->world ranks, balanced by unknowns. */
int b200cg_partition(const b200cg_plan_desc* desc, int* y_lo, int* y_hi, int64_t* lo, int64_t* hi,
                     int64_t* n_unknowns);
/* The tile table the sweep kernels of such a plan would walk on a GPU with `sms` SMs and `ctas_per_sm` resident
 * CTAs per SM (pure host logic, needs no device; for tests and tools). A tile is four ints {col0, ya, yb, xlo}: the
 * 512-column strip starting at storage column col0, emit rows [ya, yb), first unknown x of these rows; it writes the
 * unknowns x in [max(col0, xlo), min(col0 + 503, n - 1)]. CTA c walks tiles [cta_begin[c], cta_begin[c + 1]).
 * weights (n_weights entries, or NULL for the initial equal split) are the per-CTA shares feedback balancing
 * converges to. tiles holds `capacity` quadruples, cta_begin sms * ctas_per_sm + 1 ints. desc->reserved0 = 1 asks
 * for the single-sweep kernel's strips instead: 424 staged columns from storage column col0 = strip * 420 + 2, writing
 * x in [max(col0 - 2, xlo), min(col0 + 417, n - 1)]; reserved0 = 2 for its wide geometry (slabs of >= 4 M unknowns: one
 * 15-warp CTA per SM, 844 staged columns from col0 = strip * 840 + 2; call with ctas_per_sm = 1). */
int b200cg_work_split(const b200cg_plan_desc* desc, int sms, int ctas_per_sm, const double* weights, int n_weights,
                      int* tiles, int64_t capacity, int64_t* n_tiles, int* cta_begin, int* grid);

/* ------------------------------------------------------------------ setup data (kernel K0)
 * rhs b = f - Dirichlet neighbour terms: calculate_value (grid_system.cpp:45-67),
 * MatrixFreeSystem::initialize_rhs (matrix_free_system.cpp:104-141) */
int b200cg_build_rhs(b200cg_plan_t plan);
int b200cg_set_rhs(b200cg_plan_t plan, const double* b_host);
int b200cg_get_rhs(b200cg_plan_t plan, double* b_host);
/* get_true_solution_vector (grid_system.cpp:276-299, matrix_free_system.cpp:162-199) */
int b200cg_get_true_solution(b200cg_plan_t plan, double* u_host);
/* node_x_coords / node_y_coords (grid_system.cpp:188-190,234-236) */
int b200cg_get_coords(b200cg_plan_t plan, double* xs_host, double* ys_host);

/* ------------------------------------------------------------------ operator
 * y = A x: MatrixFreeSystem::apply / operator* (matrix_free_system.cpp:203-340, .hpp:59-63) */
int b200cg_apply(b200cg_plan_t plan, const double* x_host, double* y_host);
/* CSR of the assembled path. set: upload a caller-owned matrix (GridSystem::get_matrix as MSGSolver receives
 * it, msg_solver.hpp:50-53). assemble: build it on the device from the geometry (GridSystem::initiate_matrix,
 * grid_system.cpp:157-274; per-row order diag, left, right, top, bottom). get: copy it out. */
int b200cg_set_csr(b200cg_plan_t plan, int64_t nrows, int64_t nnz, const int* row_map, const int* entries,
                   const double* values);
int b200cg_assemble_csr(b200cg_plan_t plan, int64_t* nnz);
int b200cg_get_csr(b200cg_plan_t plan, int* row_map, int* entries, double* values);
/* y = A x through the CSR matrix: KokkosSparse::spmv("N", 1, A, x, 0, y) (msg_solver.cpp:93, dirichlet_solver.cpp:153) */
int b200cg_csr_apply(b200cg_plan_t plan, const double* x_host, double* y_host);

/* ------------------------------------------------------------------ solve
 * replaces MatrixFreeSolver::solve (matrix_free_system.cpp:383-482) and MSGSolver::solve (msg_solver.cpp:10-212).
 * b_host: rhs (ignored when params->rhs_on_device); u_host: true solution or NULL (empty true_solution);
 * x_host: receives the solution; stop_flag: non-zero -> the solve ends with INTERRUPTED (MSGSolver::requestStop,
 * msg_solver.hpp:76; the reference polls every iteration, msg_solver.cpp:82): the loop kernels look at the flag every
 * 16th iteration (one kernel covers the whole solve on the small-grid path: every 128th). On a sharded plan raising it
 * on ONE rank is enough: the request travels with the per-iteration reductions and every rank stops at the same
 * iteration. params (max_it, iters_per_graph, rule, eps) must be identical on all ranks of a sharded plan. */
int b200cg_solve(b200cg_plan_t plan, const b200cg_params* params, const double* b_host, const double* u_host,
                 double* x_host, b200cg_info* info, b200cg_iter_cb cb, void* user, const volatile int* stop_flag);

/* A queue of right-hand sides through one plan (time steps, parameter sweeps; bench.py's end-to-end leg): `count`
 * back-to-back b200cg_solve calls on the matrix-free operator (MatrixFreeSolver::solve, matrix_free_system.cpp:383-482,
 * once per rhs) whose host copies overlap the iterations - the H2D copy of b_hosts[i+1] and the D2H copy of x_hosts[i-1]
 * run on two copy streams (second staging buffer) while solve i iterates, so only the first upload and the last download
 * are exposed. Same arithmetic, same results as `count` separate calls. No true solution, no iteration callback;
 * params->rhs_on_device and keep_x_on_device must be 0; pinned host buffers (b200cg_alloc_pinned) are what makes the
 * copies asynchronous. x_hosts[i] may be read (and b_hosts[i] / x_hosts[i] reused) once done(user, i, &infos[i]) has
 * been called - on the calling thread, in order, at the latest before the function returns. A stop request ends the
 * queue with the solve it interrupts; the infos of the solves never started carry B200CG_STOP_INTERRUPTED and
 * iterations = 0. Inside a batch info.h2d_ms / d2h_ms / device_ms cover the scatter / gather kernels only. On a sharded
 * plan every rank calls it with the same count and params. */
typedef void (*b200cg_batch_cb)(void* user, int index, const b200cg_info* info);
int b200cg_solve_batch(b200cg_plan_t plan, const b200cg_params* params, int count, const double* const* b_hosts,
                       double* const* x_hosts, b200cg_info* infos, b200cg_batch_cb done, void* user,
                       const volatile int* stop_flag);

/* ------------------------------------------------------------------ post-processing on the last solution
 * residual = A x - b and error = x - u: DirichletSolver::computeResidual / computeError
 * (dirichlet_solver.cpp:147-180). Either output may be NULL. op selects the stencil or the CSR matrix. */
int b200cg_postprocess(b200cg_plan_t plan, int op, double* residual_host, double* error_host);
/* Diagnostics: start/end time stamps (ns, device global timer) of every persistent CTA in the last launch of a sweep
 * kernel flavour (0 = dot phase, 1 = update phase without x, 2 = update phase with x, 3 = single-sweep iteration). out receives 2 * (*n_ctas)
 * values; capacity is the number of pairs out can hold. */
int b200cg_cta_times(b200cg_plan_t plan, int flavour, uint64_t* out, int capacity, int* n_ctas);
/* Diagnostics of the sharded single-sweep iteration (plans created with B200CG_PEER_TRACE=1 in the environment): four
 * device global-timer stamps (ns) per iteration, ring of 4096 iterations indexed by iteration % 4096 - this rank's sums
 * ready (its sweep is over), published to every rank, every rank's flag seen, scalars formed. out receives 4 * (*n)
 * values. The timers of different GPUs are not synchronised: scripts/peer_trace.py bounds their offsets from the
 * stamps themselves. */
int b200cg_peer_trace(b200cg_plan_t plan, uint64_t* out, int capacity, int* n_iterations);
/* copy the device-resident solution of the last solve (keep_x_on_device) to the host */
int b200cg_get_solution(b200cg_plan_t plan, double* x_host);

#ifdef __cplusplus
}
#endif
#endif /* B200CG_H */
