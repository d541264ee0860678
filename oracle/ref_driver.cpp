// TEST INFRASTRUCTURE ONLY (oracle/). extern "C" handles around the UNMODIFIED reference classes,
// compiled together with the reference's own sources where they lie under /root/reference/solver/
// (recipe: oracle/Makefile, outputs only into oracle/_ref/). Used by tests/ to pin the C restatement
// (oracle/cg_oracle.c) and by `bench.py --impl reference` / the cpu_baseline leg as the timed CPU baseline.
// Nothing in the product library links or loads this.
//
// Reference entry points wrapped:
//   MatrixFreeSystem / MatrixFreeSolver   solver/matrix_free_system.hpp:12-69, :72-126
//   GridSystem                            solver/grid_system.h:16-87
//   MSGSolver                             solver/msg_solver.hpp:17-121
//   DirichletSolver / SolverResults       solver/dirichlet_solver.hpp:11-24, :79-184
#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "matrix_free_system.hpp"
#include "dirichlet_solver.hpp"

namespace {
// MSGSolver::solve prints a progress log to std::cout (msg_solver.cpp:174-177,202-208); keep test output clean.
struct CoutSilencer {
  std::ostringstream sink;
  std::streambuf* old;
  CoutSilencer() : sink(), old(std::cout.rdbuf(sink.rdbuf())) {}
  ~CoutSilencer() { std::cout.rdbuf(old); }
};
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- matrix-free path
void* ref_mf_create(int m, int n, double a, double b, double c, double d) {
  return new MatrixFreeSystem(m, n, a, b, c, d);
}
void ref_mf_destroy(void* h) { delete static_cast<MatrixFreeSystem*>(h); }
int ref_mf_size(void* h) { return static_cast<MatrixFreeSystem*>(h)->size(); }
void ref_mf_rhs(void* h, double* out) {
  const std::vector<double>& r = static_cast<MatrixFreeSystem*>(h)->get_rhs();
  std::memcpy(out, r.data(), r.size() * sizeof(double));
}
void ref_mf_true_solution(void* h, double* out) {
  std::vector<double> u = static_cast<MatrixFreeSystem*>(h)->get_true_solution_vector();
  std::memcpy(out, u.data(), u.size() * sizeof(double));
}
void ref_mf_apply(void* h, const double* x, double* y) {
  MatrixFreeSystem* s = static_cast<MatrixFreeSystem*>(h);
  std::vector<double> xv(x, x + s->size()), yv(s->size());
  s->apply(xv, yv);
  std::memcpy(y, yv.data(), yv.size() * sizeof(double));
}
// b == nullptr -> the system's own rhs. hist (nullable) receives (precision, residual_norm, error_norm)
// per iteration as the iteration callback reports them (matrix_free_system.cpp:466-468).
// Returns the iteration count; *seconds is the wall time of solve() alone.
int ref_mf_solve(void* h, const double* b, double eps, int max_it, double* x_out, int* converged,
                 double* hist, int hist_cap, double* seconds) {
  MatrixFreeSystem* s = static_cast<MatrixFreeSystem*>(h);
  std::vector<double> bv = b ? std::vector<double>(b, b + s->size()) : s->get_rhs();
  std::vector<double> u = s->get_true_solution_vector();
  MatrixFreeSolver solver(*s, bv, eps, max_it);
  bool conv = false;
  solver.setCompletionCallback([&](bool ok, const std::string&) { conv = ok; });
  if (hist) {
    solver.setIterationCallback([&](int it, double prec, double res, double err) {
      if (it < hist_cap) {
        hist[3 * it + 0] = prec;
        hist[3 * it + 1] = res;
        hist[3 * it + 2] = err;
      }
    });
  }
  double t0 = now_s();
  std::vector<double> x = solver.solve(u);
  double t1 = now_s();
  if (seconds) *seconds = t1 - t0;
  if (x_out) std::memcpy(x_out, x.data(), x.size() * sizeof(double));
  if (converged) *converged = conv ? 1 : 0;
  return solver.getIterations();
}

// ---------------------------------------------------------------- assembled path
void* ref_grid_create(int m, int n, double a, double b, double c, double d) {
  return new GridSystem(m, n, a, b, c, d);
}
void ref_grid_destroy(void* h) { delete static_cast<GridSystem*>(h); }
int ref_grid_rows(void* h) { return static_cast<GridSystem*>(h)->get_matrix().numRows(); }
int ref_grid_nnz(void* h) { return static_cast<GridSystem*>(h)->get_matrix().nnz(); }
void ref_grid_csr(void* h, int* row_map, int* entries, double* values) {
  const KokkosCrsMatrix& A = static_cast<GridSystem*>(h)->get_matrix();
  for (int i = 0; i <= A.numRows(); ++i) row_map[i] = A.graph.row_map(i);
  for (int k = 0; k < A.nnz(); ++k) {
    entries[k] = A.graph.entries(k);
    values[k] = A.values(k);
  }
}
void ref_grid_rhs(void* h, double* out) {
  const KokkosVector& r = static_cast<GridSystem*>(h)->get_rhs();
  for (size_t i = 0; i < r.extent(0); ++i) out[i] = r(i);
}
void ref_grid_true_solution(void* h, double* out) {
  KokkosVector u = static_cast<GridSystem*>(h)->get_true_solution_vector();
  for (size_t i = 0; i < u.extent(0); ++i) out[i] = u(i);
}
void ref_grid_coords(void* h, double* xs, double* ys) {
  GridSystem* g = static_cast<GridSystem*>(h);
  std::memcpy(xs, g->get_x_coords().data(), g->get_x_coords().size() * sizeof(double));
  std::memcpy(ys, g->get_y_coords().data(), g->get_y_coords().size() * sizeof(double));
}

struct ref_msg_info {
  int iterations;
  int converged;
  int stop_reason;  // StopCriterion as int (msg_solver.hpp:9-15)
  double final_residual_norm, final_error_norm, final_precision;
  double seconds;
  int n_callbacks;
};

// MSGSolver on the GridSystem's own matrix and rhs. eps_* <= 0 disables a rule (msg_solver.cpp:144-162).
// cb_log (nullable, capacity cb_cap rows of 4): (iteration, precision_max, r_max, err_max) per callback.
void ref_msg_solve(void* grid, double eps_p, double eps_r, double eps_e, int max_it, int with_true,
                   double* x_out, ref_msg_info* info, double* cb_log, int cb_cap) {
  GridSystem* g = static_cast<GridSystem*>(grid);
  MSGSolver solver(g->get_matrix(), g->get_rhs(), 1e-6, max_it);
  solver.setPrecisionEps(eps_p);
  solver.setResidualEps(eps_r);
  solver.setExactErrorEps(eps_e);
  int ncb = 0;
  if (cb_log) {
    solver.setIterationCallback([&](int it, double p, double r, double e) {
      if (ncb < cb_cap) {
        cb_log[4 * ncb + 0] = it;
        cb_log[4 * ncb + 1] = p;
        cb_log[4 * ncb + 2] = r;
        cb_log[4 * ncb + 3] = e;
      }
      ++ncb;
    });
  }
  KokkosVector u;
  if (with_true) u = g->get_true_solution_vector();
  CoutSilencer quiet;
  double t0 = now_s();
  KokkosVector x = solver.solve(u);
  double t1 = now_s();
  if (x_out)
    for (size_t i = 0; i < x.extent(0); ++i) x_out[i] = x(i);
  info->iterations = solver.getIterations();
  info->converged = solver.hasConverged() ? 1 : 0;
  info->stop_reason = static_cast<int>(solver.getStopReason());
  info->final_residual_norm = solver.getFinalResidualNorm();
  info->final_error_norm = solver.getFinalErrorNorm();
  info->final_precision = solver.getFinalPrecision();
  info->seconds = t1 - t0;
  info->n_callbacks = ncb;
}

// DirichletSolver facade (note its (n, m) order, dirichlet_solver.cpp:24). Output arrays hold N doubles each.
struct ref_dirichlet_out {
  int iterations;
  int converged;
  double residual_norm, error_norm;
  char stop_reason[256];
  int size;
  double seconds;
};
void ref_dirichlet_solve(int n, int m, double a, double b, double c, double d, double eps_p, double eps_r,
                         double eps_e, int max_iter, int use_p, int use_r, int use_e, double* solution,
                         double* true_solution, double* residual, double* error, double* x_coords,
                         double* y_coords, ref_dirichlet_out* out) {
  CoutSilencer quiet;
  DirichletSolver ds(n, m, a, b, c, d);
  ds.setSolverParameters(eps_p, eps_r, eps_e, max_iter);
  ds.enablePrecisionStopping(use_p != 0);
  ds.enableResidualStopping(use_r != 0);
  ds.enableErrorStopping(use_e != 0);
  double t0 = now_s();
  SolverResults res = ds.solve();
  double t1 = now_s();
  auto put = [](double* dst, const std::vector<double>& v) {
    if (dst) std::memcpy(dst, v.data(), v.size() * sizeof(double));
  };
  put(solution, res.solution);
  put(true_solution, res.true_solution);
  put(residual, res.residual);
  put(error, res.error);
  put(x_coords, res.x_coords);
  put(y_coords, res.y_coords);
  out->iterations = res.iterations;
  out->converged = res.converged ? 1 : 0;
  out->residual_norm = res.residual_norm;
  out->error_norm = res.error_norm;
  std::strncpy(out->stop_reason, res.stop_reason.c_str(), sizeof(out->stop_reason) - 1);
  out->stop_reason[sizeof(out->stop_reason) - 1] = 0;
  out->size = static_cast<int>(res.solution.size());
  out->seconds = t1 - t0;
}

}  // extern "C"
