// Drop-in host classes for the reference's public C++ surface (static library `dirichlet_solver`,
// solver/CMakeLists.txt:65): same class names, constructor and method signatures, argument meaning, callback
// cadence, ownership and error conventions - with every numerical body routed to libb200cg (include/b200cg.h).
// A consumer such as qt_gui/src/mainwindow.cpp keeps its `#include "dirichlet_solver.hpp"` and links
// -lb200_dropin -lb200cg instead of the Kokkos build.
//
// Interface map (reference file:line -> here):
//   Solver                 solver/solver.hpp:17-66
//   StopCriterion          solver/msg_solver.hpp:9-15
//   MSGSolver              solver/msg_solver.hpp:17-121          (solve: msg_solver.cpp:10-212)
//   GridSystem             solver/grid_system.h:16-87            (assembly: grid_system.cpp:157-274)
//   MatrixFreeSystem       solver/matrix_free_system.hpp:12-69   (apply: matrix_free_system.cpp:203-340)
//   MatrixFreeSolver       solver/matrix_free_system.hpp:72-126  (solve: matrix_free_system.cpp:383-482)
//   SolverResults          solver/dirichlet_solver.hpp:11-24
//   ResultsIO              solver/dirichlet_solver.hpp:27-77     (dirichlet_solver.cpp:255-457)
//   DirichletSolver        solver/dirichlet_solver.hpp:79-184    (solve: dirichlet_solver.cpp:61-131)
// Errors: C-ABI status codes become std::runtime_error (std::invalid_argument for rejected geometry), as the
// reference throws from its constructors and solve(); the GUI wraps solve() in try/catch(std::exception).
#pragma once

#include <atomic>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

#include "kokkos_compat.hpp"

using execution_space = Kokkos::DefaultExecutionSpace;
using memory_space = Kokkos::HostSpace;
using KokkosVector = Kokkos::View<double*, memory_space>;
using KokkosCrsMatrix = KokkosSparse::CrsMatrix<double, int, execution_space, void, int>;

namespace KokkosSparse {
// y = beta*y + alpha*A*x with A*x computed on the GPU (b200cg_csr_apply); mode must be "N".
void spmv(const char mode[], double alpha, const KokkosCrsMatrix& A, const KokkosVector& x, double beta,
          const KokkosVector& y);
}  // namespace KokkosSparse

namespace b200 {
// Owns one b200cg_plan_t; shared between a system object and the solvers created on it.
class PlanHandle;
using PlanPtr = std::shared_ptr<PlanHandle>;
}  // namespace b200

// ------------------------------------------------------------------------------------------------- Solver
class Solver {
 protected:
  const KokkosCrsMatrix& a;
  const KokkosVector& b;
  double eps;
  int maxIterations;
  int iterations;
  std::string name;
  std::function<void(int, double, double, double)> iteration_callback;
  std::function<void(bool, const std::string&)> completion_callback;

 public:
  Solver(const KokkosCrsMatrix& a, const KokkosVector& b, double eps = 1e-6, int maxIterations = 10000,
         const std::string& name = "Базовый решатель")
      : a(a), b(b), eps(eps), maxIterations(maxIterations), iterations(0), name(name) {}
  virtual ~Solver() = default;

  virtual KokkosVector solve(const KokkosVector& true_solution) = 0;

  // (iteration, ||x_n - x_{n-1}||, ||A x - b||, ||u - x||)
  void setIterationCallback(std::function<void(int, double, double, double)> callback) { iteration_callback = callback; }
  void setCompletionCallback(std::function<void(bool, const std::string&)> callback) { completion_callback = callback; }
  int getIterations() const { return iterations; }
  std::string getName() const { return name; }
};

// ------------------------------------------------------------------------------------------------- MSGSolver
enum class StopCriterion { ITERATIONS, PRECISION, RESIDUAL, EXACT_ERROR, INTERRUPTED };

class MSGSolver : public Solver {
 public:
  MSGSolver(const KokkosCrsMatrix& a, const KokkosVector& b, double eps = 1e-6, int maxIterations = 10000);
  ~MSGSolver() override;

  void setPrecisionEps(double e) { eps_precision = e; }
  void setResidualEps(double e) { eps_residual = e; }
  void setExactErrorEps(double e) { eps_exact_error = e; }

  bool hasConverged() const { return converged; }
  StopCriterion getStopReason() const { return stop_reason; }
  std::string getStopReasonText() const;

  void requestStop() { stop_requested.store(1); }
  void resetStop() { stop_requested.store(0); }
  bool isStopRequested() const { return stop_requested.load() != 0; }

  double getFinalResidualNorm() const { return final_residual_norm; }
  double getFinalErrorNorm() const { return final_error_norm; }
  double getFinalPrecision() const { return final_precision; }

  // hides Solver::setIterationCallback exactly as the reference does (msg_solver.hpp:32,112)
  void setIterationCallback(std::function<void(int, double, double, double)> callback) { iteration_callback = callback; }

  KokkosVector solve(const KokkosVector& true_solution) override;
  std::string generateReport(int n, int m, double a, double b, double c, double d) const;

  // B200 additions (not in the reference): reuse an existing geometric plan (DirichletSolver passes its GridSystem's)
  // and expose the device timing of the last solve. matrix_free = true: the plan's geometry IS this matrix
  // (GridSystem assembled it from that geometry), so the iteration applies the 5-point operator on the fly instead
  // of reading the CSR arrays - bit-identical operator (tests: csr_apply == apply), ~2.5x less memory traffic.
  void attachPlan(const b200::PlanPtr& plan, bool matrix_free = false) {
    plan_ = plan;
    matrix_free_ = matrix_free;
  }
  bool usesMatrixFreeOperator() const { return matrix_free_; }
  const b200::PlanPtr& plan() const { return plan_; }
  double lastSolveMilliseconds() const { return last_solve_ms; }

 private:
  double eps_precision, eps_residual, eps_exact_error;
  bool converged;
  StopCriterion stop_reason;
  double final_residual_norm, final_error_norm, final_precision;
  std::function<void(int, double, double, double)> iteration_callback;
  std::atomic<int> stop_requested;  // polled by b200cg_solve between graph launches
  b200::PlanPtr plan_;
  bool matrix_free_ = false;
  double last_solve_ms = 0.0;
};

// ------------------------------------------------------------------------------------------------- GridSystem
class GridSystem {
 public:
  struct NodeCoordinates {
    double x;
    double y;
  };

  GridSystem(int m, int n, double a, double b, double c, double d);
  ~GridSystem();

  // The CSR arrays are assembled on the device and mirrored into the host containers the first time someone asks for
  // them (3 GB at 8192^2): a DirichletSolver on the matrix-free operator never does.
  const KokkosCrsMatrix& get_matrix() const {
    ensure_matrix();
    return matrix;
  }
  const KokkosVector& get_rhs() const { return rhs; }
  KokkosVector get_true_solution_vector();
  const std::vector<double>& get_x_coords() const { return node_x_coords; }
  const std::vector<double>& get_y_coords() const { return node_y_coords; }
  NodeCoordinates get_node_coordinates(int solution_index) const;

  friend std::ostream& operator<<(std::ostream& os, const GridSystem& grid);

  // B200 additions
  const b200::PlanPtr& plan() const { return plan_; }
  const KokkosCrsMatrix& matrix_handle() const { return matrix; }  // the member itself, without forcing the assembly
  void ensure_device_matrix() const;                               // CSR resident on the device (no host mirror)
  long long rows() const { return rows_; }

 private:
  void ensure_matrix() const;
  int n, m;
  double a, b, c, d;
  long long rows_ = 0;
  mutable long long nnz_ = -1;  // -1: not assembled yet
  mutable KokkosCrsMatrix matrix;
  KokkosVector rhs;
  std::vector<double> node_x_coords, node_y_coords;
  b200::PlanPtr plan_;
};

// ------------------------------------------------------------------------------------------------- preconditioned CG
// B200 addition behind the reference's abstract base (solver.hpp:17-66 exists for further solvers; the reference ships
// only MSGSolver): CG preconditioned by a geometric-multigrid V-cycle on the grid's geometry
// (b200cg_params.preconditioner = B200CG_PRECOND_MULTIGRID). Same x0 = 0 and the relative-residual stop rule of
// MatrixFreeSolver (matrix_free_system.cpp:409): ||r||_2 <= eps ||r0||_2. The iteration count does not grow with the
// grid (7 at any n, against ~2.7 n for plain CG); the iterates are not the reference's, the solution is.
class MultigridPCGSolver : public Solver {
 public:
  explicit MultigridPCGSolver(const GridSystem& grid, double eps = 1e-6, int maxIterations = 10000);
  ~MultigridPCGSolver() override;

  KokkosVector solve(const KokkosVector& true_solution) override;  // true_solution is not used (may be empty)

  bool hasConverged() const { return converged; }
  double getInitialResidualNorm() const { return r0_l2; }  // ||r0||_2
  double getFinalResidualNorm() const { return r_l2; }     // recurrence ||r||_2 at exit
  int multigridLevels() const { return levels; }
  double lastSolveMilliseconds() const { return last_solve_ms; }

 private:
  b200::PlanPtr plan_;
  bool converged = false;
  double r0_l2 = 0.0, r_l2 = 0.0, last_solve_ms = 0.0;
  int levels = 0;
};

// ------------------------------------------------------------------------------------------------- matrix-free
class MatrixFreeSystem {
 public:
  MatrixFreeSystem(int m, int n, double a, double b, double c, double d);
  ~MatrixFreeSystem() = default;

  const std::vector<double>& get_rhs() const { return rhs; }
  std::vector<double> get_true_solution_vector();
  void apply(const std::vector<double>& x, std::vector<double>& y) const;
  std::vector<double> operator*(const std::vector<double>& x) const {
    std::vector<double> result(size());
    apply(x, result);
    return result;
  }
  int size() const { return static_cast<int>(rhs.size()); }

  friend std::ostream& operator<<(std::ostream& os, const MatrixFreeSystem& grid);

  const b200::PlanPtr& plan() const { return plan_; }  // B200 addition

 private:
  int n, m;
  double a, b, c, d;
  std::vector<double> rhs;
  b200::PlanPtr plan_;
};

class MatrixFreeSolver {
 public:
  MatrixFreeSolver(const MatrixFreeSystem& system, const std::vector<double>& b, double eps = 1e-6,
                   int maxIterations = 10000, const std::string& name = "Matrix-free solver");
  virtual ~MatrixFreeSolver() = default;

  std::vector<double> solve(const std::vector<double>& true_solution);

  void setIterationCallback(std::function<void(int, double, double, double)> callback) { iteration_callback = callback; }
  void setCompletionCallback(std::function<void(bool, const std::string&)> callback) { completion_callback = callback; }
  int getIterations() const { return iterations; }
  std::string getName() const { return name; }

  // B200 additions: device timing of the last solve; opt-in multigrid preconditioner (no per-iteration callback then)
  double lastSolveMilliseconds() const { return last_solve_ms; }
  void enableMultigridPreconditioner(bool enable) { multigrid = enable; }
  // B200 addition: solve() once per entry of right_hand_sides (the constructor's b is not used), same results, as ONE
  // queue whose host copies overlap the iterations (b200cg_solve_batch). No per-iteration callback; the completion
  // callback fires once per solve; getIterations() reports the last solve, iterations_out every one.
  std::vector<std::vector<double>> solveBatch(const std::vector<std::vector<double>>& right_hand_sides,
                                              std::vector<int>* iterations_out = nullptr);

 private:
  bool multigrid = false;
  const MatrixFreeSystem& system;
  const std::vector<double>& b;
  double eps;
  int maxIterations;
  int iterations;
  std::string name;
  std::function<void(int, double, double, double)> iteration_callback;
  std::function<void(bool, const std::string&)> completion_callback;
  double last_solve_ms = 0.0;
};

// ------------------------------------------------------------------------------------------------- facade
struct SolverResults {
  std::vector<double> solution;
  std::vector<double> true_solution;
  std::vector<double> residual;  // A x - b
  std::vector<double> error;     // x - u
  std::vector<double> x_coords;
  std::vector<double> y_coords;
  double residual_norm = 0.0;
  double error_norm = 0.0;
  int iterations = 0;
  double precision = 0.0;  // the reference never assigns it; here: ||x_n - x_{n-1}||_inf at exit
  bool converged = false;
  std::string stop_reason;
};

class ResultsIO {
 public:
  static bool saveResults(const std::string& filename, const SolverResults& results, int n, int m, double a, double b,
                          double c, double d, const std::string& solver_name);
  // reads the sections by their true length (the reference assumes n*m entries, which overruns on the L-shape)
  static bool loadResults(const std::string& filename, SolverResults& results, int& n, int& m, double& a, double& b,
                          double& c, double& d, std::string& solver_name);
  static bool saveMatrixAndRhs(const std::string& filename, const KokkosCrsMatrix& A, const KokkosVector& b, int n,
                               int m);
  static bool saveSolutionFor3D(const std::string& filename, const std::vector<std::vector<double>>& solution,
                                double a_bound, double b_bound, double c_bound, double d_bound);
};

class DirichletSolver {
 public:
  // note the (n, m) order, swapped relative to GridSystem (dirichlet_solver.cpp:24)
  DirichletSolver(int n = 10, int m = 10, double a = 0.0, double b = 1.0, double c = 0.0, double d = 1.0);
  ~DirichletSolver();

  void setGridParameters(int n, int m, double a, double b, double c, double d);
  void setSolverParameters(double eps_p, double eps_r, double eps_e, int max_iter);

  void enablePrecisionStopping(bool enable) { use_precision_stopping = enable; }
  void enableResidualStopping(bool enable) { use_residual_stopping = enable; }
  void enableErrorStopping(bool enable) { use_error_stopping = enable; }
  void enableMaxIterationsStopping(bool enable) { use_max_iterations_stopping = enable; }

  void requestStop() {
    stop_requested = true;
    if (solver) solver->requestStop();
  }
  std::string getMethodName() const { return solver ? solver->getName() : "МСГ"; }

  void setIterationCallback(std::function<void(int, double, double, double)> callback);
  void setCompletionCallback(std::function<void(const SolverResults&)> callback) { completion_callback = callback; }

  SolverResults solve();

  std::vector<double> getSolution() const;
  std::vector<double> getTrueSolution() const;
  std::vector<std::vector<double>> solutionToMatrix() const;

  std::string generateReport() const;
  bool saveResultsToFile(const std::string& filename) const;
  bool saveMatrixAndRhsToFile(const std::string& filename) const;

  const GridSystem* getGridSystem() const { return grid.get(); }

 private:
  int n_internal, m_internal;
  double a_bound, b_bound, c_bound, d_bound;
  double eps_precision, eps_residual, eps_exact_error;
  int max_iterations;
  bool use_precision_stopping, use_residual_stopping, use_error_stopping, use_max_iterations_stopping;
  bool stop_requested = false;
  std::function<void(int, double, double, double)> iteration_callback;
  std::function<void(const SolverResults&)> completion_callback;
  std::unique_ptr<GridSystem> grid;
  std::unique_ptr<MSGSolver> solver;
  KokkosVector solution;
  KokkosVector true_solution;
  SolverResults last_results;
};
