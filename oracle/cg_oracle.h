/* TEST INFRASTRUCTURE ONLY (oracle/): plain-C CPU restatement of the reference's conjugate-gradient
 * Dirichlet-Poisson path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it - as the checker, never as the thing measured or shipped. The product library
 * (iterative_solvers_b200/csrc) has no CPU fallback and never links this.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle.py) against
 *   - the reference's golden vectors check.py:4-19, check_debug.py:36, py_debug.txt:5-15, and
 *   - the UNMODIFIED reference sources compiled here into oracle/_ref/libref_cg.so (oracle/Makefile).
 * The RECT (full-rectangle) and LSHAPE_ANY (L-shape for any n, m with the reference's numbering defects repaired)
 * domain kinds have NO reference counterpart: "parity unpinned" for them; they are only checked against an
 * independent scipy sparse solve, the analytic solution, and (LSHAPE_ANY at even n == m) the LSHAPE kind itself.
 */
#ifndef CG_ORACLE_H
#define CG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { CGO_LSHAPE = 0, CGO_RECT = 1, CGO_LSHAPE_ANY = 3 };

typedef struct {
  int n, m;           /* numbers of intervals in x and y (grid_system.cpp:314-315) */
  double a, b, c, d;  /* domain [a,b] x [c,d] */
  double hx, hy;      /* steps */
  double xk, yk, A;   /* 1/hx^2, 1/hy^2, -2(xk+yk) (grid_system.cpp:316-318) */
  int kind;           /* CGO_LSHAPE (reference geometry) or CGO_RECT */
} cgo_grid;

/* Argument order (m, n, ...) follows GridSystem / MatrixFreeSystem (grid_system.cpp:301).
 * Returns 0, or -1 if the reference numbering is not self-consistent for (n, m) (LSHAPE needs even n == m >= 4). */
int cgo_grid_init(cgo_grid* g, int m, int n, double a, double b, double c, double d, int kind);
long cgo_size(const cgo_grid* g);
int cgo_is_unknown(const cgo_grid* g, int x, int y);
long cgo_index(const cgo_grid* g, int x, int y); /* compact index in the reference order, -1 if not an unknown */
void cgo_node(const cgo_grid* g, long idx, int* x, int* y);

void cgo_rhs(const cgo_grid* g, double* b);                      /* N */
void cgo_true_solution(const cgo_grid* g, double* u);            /* N */
void cgo_node_coords(const cgo_grid* g, double* xs, double* ys); /* N each */
void cgo_apply(const cgo_grid* g, const double* x, double* y);   /* y = A x (row-wise walk of the form below) */
void cgo_apply_nodewise(const cgo_grid* g, const double* x, double* y); /* the reference's node-by-node form */

typedef struct {
  int iterations;
  int converged;
  double r0_norm; /* ||r_0||_2 */
  double r_norm;  /* recurrence ||r||_2 at exit */
  double seconds;
} cgo_mf_info;

/* MatrixFreeSolver::solve (matrix_free_system.cpp:383-482). hist (nullable): 3 doubles per iteration
 * (||x_k - x_{k-1}||_2, ||b - A x_k||_2 recomputed, ||x_k - u||_2) as the iteration callback reports them;
 * u may be NULL when hist is NULL. snapshot_r / snapshot_p (nullable): r and p after the last iteration. */
void cgo_mf_solve(const cgo_grid* g, const double* b, const double* u, double eps, int max_it, double* x,
                  cgo_mf_info* info, double* hist, int hist_cap, double* snapshot_r, double* snapshot_p);

/* Not a reference function: MatrixFreeSolver::solve with alpha from the single-reduction CG recurrence, in the
 * reference's arithmetic - the CPU statement of the product's opt-in single-sweep iteration. */
void cgo_mf_solve_single(const cgo_grid* g, const double* b, double eps, int max_it, double* x, cgo_mf_info* info);

/* 0 = reference summation order (default), 1 = long-double accumulation (diagnostic, see cg_oracle.c) */
void cgo_set_dot_mode(int mode);

long cgo_csr_nnz(const cgo_grid* g);
/* GridSystem::initiate_matrix (grid_system.cpp:157-274): per row diag, left, right, top, bottom. */
void cgo_csr_assemble(const cgo_grid* g, int* row_map, int* entries, double* values);
void cgo_spmv(long nrows, const int* row_map, const int* entries, const double* values, const double* x,
              double* y);

enum { CGO_STOP_ITERATIONS = 0, CGO_STOP_PRECISION = 1, CGO_STOP_RESIDUAL = 2, CGO_STOP_EXACT_ERROR = 3,
       CGO_STOP_INTERRUPTED = 4 };

typedef struct {
  int iterations;
  int converged;
  int stop_reason;
  double r_max;   /* final ||r||_inf (recurrence)      msg_solver.cpp:188 */
  double dx_max;  /* final ||x_n - x_{n-1}||_inf        msg_solver.cpp:189 */
  double err_max; /* final ||x - u||_inf                msg_solver.cpp:190 */
  double r_l2;    /* final ||r||_2 */
  double seconds;
  int n_callbacks;
} cgo_msg_info;

/* MSGSolver::solve (msg_solver.cpp:10-212) on a CSR matrix. u may be NULL (empty true_solution).
 * eps_* <= 0 disables the rule. cb_log (nullable): rows of (iteration, dx_max, r_max, err_max) at the
 * reference's callback cadence (it 0, 1, every 100, final). */
void cgo_msg_solve(long nrows, const int* row_map, const int* entries, const double* values, const double* b,
                   const double* u, double eps_p, double eps_r, double eps_e, int max_it, double* x,
                   cgo_msg_info* info, double* cb_log, int cb_cap);
/* not a reference function: the same rules with alpha from the single-reduction recurrence, matrix-free operator */
void cgo_msg_solve_single(const cgo_grid* g, const double* b, const double* u, double eps_p, double eps_r, double eps_e,
                          int max_it, double* x, cgo_msg_info* info, double* cb_log, int cb_cap);

#ifdef __cplusplus
}
#endif
#endif
