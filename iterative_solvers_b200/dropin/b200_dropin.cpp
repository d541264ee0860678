// Bodies of the drop-in classes: thin marshalling onto the C ABI of libb200cg. No numerical work happens here.
#include "b200_dropin.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <limits>
#include <sstream>
#include <stdexcept>

#include "b200cg.h"

namespace b200 {

static void check(int status) {
  if (status == B200CG_OK) return;
  const std::string msg = std::string("b200cg: ") + b200cg_last_error();
  if (status == B200CG_ERR_INVALID_ARG) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

static int default_device() {
  const char* env = std::getenv("B200CG_DEVICE");
  return env ? std::atoi(env) : 0;
}

class PlanHandle {
 public:
  explicit PlanHandle(const b200cg_plan_desc& desc) { check(b200cg_plan_create(&plan_, &desc)); }
  ~PlanHandle() { b200cg_plan_destroy(plan_); }
  PlanHandle(const PlanHandle&) = delete;
  PlanHandle& operator=(const PlanHandle&) = delete;
  b200cg_plan_t get() const { return plan_; }
  bool has_matrix = false;

 private:
  b200cg_plan_t plan_ = nullptr;
};

static PlanPtr make_geometric_plan(int m, int n, double a, double b, double c, double d) {
  b200cg_plan_desc desc = {};
  desc.n = n;
  desc.m = m;
  desc.a = a;
  desc.b = b;
  desc.c = c;
  desc.d = d;
  desc.domain = B200CG_DOMAIN_LSHAPE;  // the reference's region (grid_system.cpp:17-43)
  desc.device = default_device();
  desc.world = 1;
  return std::make_shared<PlanHandle>(desc);
}

static PlanPtr make_generic_plan(long long rows) {
  b200cg_plan_desc desc = {};
  desc.domain = B200CG_DOMAIN_GENERIC;
  desc.generic_rows = rows;
  desc.device = default_device();
  desc.world = 1;
  return std::make_shared<PlanHandle>(desc);
}

static void upload_matrix(const PlanPtr& plan, const KokkosCrsMatrix& A) {
  check(b200cg_set_csr(plan->get(), A.numRows(), A.nnz(), A.graph.row_map.data(), A.graph.entries.data(),
                       A.values.data()));
  plan->has_matrix = true;
}

struct CallbackBox {
  const std::function<void(int, double, double, double)>* fn;
};
static void callback_trampoline(void* user, int it, double precision, double residual, double error) {
  const CallbackBox* box = static_cast<const CallbackBox*>(user);
  if (box->fn && *box->fn) (*box->fn)(it, precision, residual, error);
}

}  // namespace b200

using b200::check;

// ------------------------------------------------------------------------------------------------- spmv
void KokkosSparse::spmv(const char mode[], double alpha, const KokkosCrsMatrix& A, const KokkosVector& x, double beta,
                        const KokkosVector& y) {
  if (!mode || mode[0] != 'N') throw std::invalid_argument("KokkosSparse::spmv: only mode \"N\" is supported");
  // One plan per calling thread and matrix size is kept between calls (stream, events, device buffers; b200cg_set_csr
  // reuses buffers of the right size). The matrix CONTENTS are uploaded on every call: the Views are plain host
  // memory the caller may have rewritten since.
  static thread_local b200::PlanPtr cached;
  static thread_local long long cached_rows = -1;
  if (!cached || cached_rows != A.numRows()) {
    cached = b200::make_generic_plan(A.numRows());
    cached_rows = A.numRows();
  }
  b200::PlanPtr plan = cached;
  b200::upload_matrix(plan, A);
  std::vector<double> ax(static_cast<size_t>(A.numRows()));
  check(b200cg_csr_apply(plan->get(), x.data(), ax.data()));
  for (size_t i = 0; i < ax.size(); ++i) y(i) = (beta == 0.0 ? 0.0 : beta * y(i)) + alpha * ax[i];
}

// ------------------------------------------------------------------------------------------------- MSGSolver
MSGSolver::MSGSolver(const KokkosCrsMatrix& a, const KokkosVector& b, double eps, int maxIterations)
    : Solver(a, b, eps, maxIterations, "Метод серединных градиентов"),
      eps_precision(eps), eps_residual(eps), eps_exact_error(eps),
      converged(false), stop_reason(StopCriterion::ITERATIONS),
      final_residual_norm(0.0), final_error_norm(0.0), final_precision(0.0), stop_requested(0) {}

MSGSolver::~MSGSolver() = default;

std::string MSGSolver::getStopReasonText() const {
  switch (stop_reason) {
    case StopCriterion::ITERATIONS: return "Достигнуто максимальное число итераций";
    case StopCriterion::PRECISION: return "Достигнута требуемая точность по норме разности xn и xn-1";
    case StopCriterion::RESIDUAL: return "Достигнута требуемая точность по норме невязки";
    case StopCriterion::EXACT_ERROR: return "Достигнута требуемая точность по норме разности с истинным решением";
    case StopCriterion::INTERRUPTED: return "Прервано пользователем";
  }
  return "Неизвестная причина остановки";
}

KokkosVector MSGSolver::solve(const KokkosVector& true_solution) {
  converged = false;  // msg_solver.cpp:12-13
  stop_requested.store(0);
  const long long rows = static_cast<long long>(b.extent(0));
  if (!plan_) plan_ = b200::make_generic_plan(rows);
  if (!matrix_free_ && !plan_->has_matrix) b200::upload_matrix(plan_, a);

  b200cg_params prm = {};
  prm.op = matrix_free_ ? B200CG_OP_MATRIX_FREE : B200CG_OP_CSR;
  prm.rule = B200CG_RULE_MAXNORM;
  prm.eps_p = eps_precision;
  prm.eps_r = eps_residual;
  prm.eps_e = eps_exact_error;
  prm.max_it = maxIterations;
  prm.callback_every = 100;  // msg_solver.cpp:172
  b200cg_info info = {};
  KokkosVector x("x", static_cast<size_t>(rows));
  b200::CallbackBox box{&iteration_callback};
  const bool want_cb = static_cast<bool>(iteration_callback);
  static_assert(sizeof(std::atomic<int>) == sizeof(int), "stop flag is polled through a plain int view");
  const volatile int* flag = reinterpret_cast<const volatile int*>(&stop_requested);
  check(b200cg_solve(plan_->get(), &prm, b.data(), true_solution.extent(0) > 0 ? true_solution.data() : nullptr,
                     x.data(), &info, want_cb ? b200::callback_trampoline : nullptr, &box, flag));
  iterations = info.iterations;
  converged = info.converged != 0;
  stop_reason = static_cast<StopCriterion>(info.stop_reason);
  final_residual_norm = info.r_max;  // msg_solver.cpp:188-190
  final_precision = info.dx_max;
  final_error_norm = info.err_max;
  last_solve_ms = info.solve_ms;
  return x;
}

std::string MSGSolver::generateReport(int n, int m, double a_, double b_, double c_, double d_) const {
  std::ostringstream out;
  out << "ОТЧЕТ О РЕШЕНИИ ЗАДАЧИ ДИРИХЛЕ\n===========================\n\n";
  out << "ПАРАМЕТРЫ ЗАДАЧИ:\n----------------\n"
      << "Размер сетки: " << n << "x" << m << " внутренних узлов\n"
      << "Область: [" << a_ << ", " << b_ << "] x [" << c_ << ", " << d_ << "]\n"
      << "Шаг по x: " << (b_ - a_) / (n + 1) << "\n"
      << "Шаг по y: " << (d_ - c_) / (m + 1) << "\n"
      << "Общее количество неизвестных: " << n * m << "\n\n";
  out << "МЕТОД РЕШЕНИЯ:\n-------------\n"
      << "Название метода: " << name << "\n"
      << "Максимальное число итераций: " << maxIterations << "\n"
      << "Критерии остановки:\n"
      << "  - Точность ||xn-x(n-1)||: " << eps_precision << "\n"
      << "  - Норма невязки ||Ax-b||: " << eps_residual << "\n"
      << "  - Норма ошибки ||u-x||: " << eps_exact_error << "\n\n";
  out << "РЕЗУЛЬТАТЫ РЕШЕНИЯ:\n-----------------\n"
      << "Выполнено итераций: " << iterations << "\n"
      << "Сходимость: " << (converged ? "Да" : "Нет") << "\n"
      << "Причина остановки: " << getStopReasonText() << "\n"
      << "Достигнутые величины:\n"
      << std::scientific
      << "  - Точность ||xn-x(n-1)||: " << final_precision << "\n"
      << "  - Норма невязки ||Ax-b||: " << final_residual_norm << "\n"
      << "  - Норма ошибки ||u-x||: " << final_error_norm << "\n\n";
  out << "ПРИМЕЧАНИЯ:\n----------\n"
      << "- Все нормы вычислены как maximum-norm (максимальный модуль элемента)\n"
      << "- Для сравнения с истинным решением используется функция u(x,y) = exp(x^2 - y^2)\n";
  return out.str();
}

// ------------------------------------------------------------------------------------------------- GridSystem
GridSystem::GridSystem(int m_, int n_, double a_, double b_, double c_, double d_)
    : n(n_), m(m_), a(a_), b(b_), c(c_), d(d_) {
  if (!Kokkos::is_initialized()) Kokkos::initialize();  // grid_system.cpp:304-306
  plan_ = b200::make_geometric_plan(m, n, a, b, c, d);
  int64_t rows = 0;
  check(b200cg_size(plan_->get(), &rows));
  rows_ = rows;
  rhs = KokkosVector("rhs", static_cast<size_t>(rows));
  check(b200cg_build_rhs(plan_->get()));
  check(b200cg_get_rhs(plan_->get(), rhs.data()));
  node_x_coords.resize(static_cast<size_t>(rows));
  node_y_coords.resize(static_cast<size_t>(rows));
  check(b200cg_get_coords(plan_->get(), node_x_coords.data(), node_y_coords.data()));
}

GridSystem::~GridSystem() = default;

// assembly on the device (grid_system.cpp:157-274), on first use
void GridSystem::ensure_device_matrix() const {
  if (nnz_ >= 0) return;
  int64_t nnz = 0;
  check(b200cg_assemble_csr(plan_->get(), &nnz));
  plan_->has_matrix = true;
  nnz_ = nnz;
}

// ... and its mirror in the host containers the API exposes, on first use
void GridSystem::ensure_matrix() const {
  if (matrix.numRows() > 0 || rows_ == 0) return;
  ensure_device_matrix();
  Kokkos::View<int*, memory_space> row_map("row_map", static_cast<size_t>(rows_ + 1));
  Kokkos::View<int*, memory_space> entries("entries", static_cast<size_t>(nnz_));
  Kokkos::View<double*, memory_space> values("values", static_cast<size_t>(nnz_));
  check(b200cg_get_csr(plan_->get(), row_map.data(), entries.data(), values.data()));
  matrix = KokkosCrsMatrix("A", static_cast<int>(rows_), static_cast<int>(rows_), static_cast<int>(nnz_), values, row_map,
                           entries);
}

KokkosVector GridSystem::get_true_solution_vector() {
  if (rows_ == 0)
    throw std::runtime_error("Matrix not initialized, cannot determine size for true solution vector.");
  KokkosVector u("true_u", static_cast<size_t>(rows_));
  check(b200cg_get_true_solution(plan_->get(), u.data()));
  return u;
}

GridSystem::NodeCoordinates GridSystem::get_node_coordinates(int solution_index) const {
  NodeCoordinates at{0.0, 0.0};  // zero coordinates for an index outside the system (grid_system.cpp:339-341)
  if (solution_index >= 0 && solution_index < rows_) {
    at.x = node_x_coords[static_cast<size_t>(solution_index)];
    at.y = node_y_coords[static_cast<size_t>(solution_index)];
  }
  return at;
}

std::ostream& operator<<(std::ostream& os, const GridSystem& grid) {
  grid.ensure_matrix();
  const double rows = grid.matrix.numRows(), cols = grid.matrix.numCols();
  os << "GridSystem Matrix Information:" << std::endl
     << "  Dimensions: " << grid.n << "x" << grid.m << std::endl
     << "  Domain: [" << grid.a << ", " << grid.b << "] x [" << grid.c << ", " << grid.d << "]" << std::endl
     << "  Matrix size: " << grid.matrix.numRows() << " rows x " << grid.matrix.numCols() << " columns" << std::endl
     << "  Non-zero elements: " << grid.matrix.nnz() << std::endl
     << "  Sparsity: " << (1.0 - grid.matrix.nnz() / (rows * cols)) * 100.0 << "%" << std::endl;
  return os;
}

// ------------------------------------------------------------------------------------------------- preconditioned CG
MultigridPCGSolver::MultigridPCGSolver(const GridSystem& grid, double eps_, int maxIterations_)
    : Solver(grid.matrix_handle(), grid.get_rhs(), eps_, maxIterations_, "МСГ с многосеточным предобуславливателем"),
      plan_(grid.plan()) {}

MultigridPCGSolver::~MultigridPCGSolver() = default;

KokkosVector MultigridPCGSolver::solve(const KokkosVector& /*true_solution*/) {
  b200cg_params prm = {};
  prm.op = B200CG_OP_MATRIX_FREE;
  prm.rule = B200CG_RULE_REL_L2;
  prm.eps_rel = eps;
  prm.max_it = maxIterations;
  prm.preconditioner = B200CG_PRECOND_MULTIGRID;
  b200cg_info info = {};
  KokkosVector x("x", b.extent(0));
  check(b200cg_solve(plan_->get(), &prm, b.data(), nullptr, x.data(), &info, nullptr, nullptr, nullptr));
  iterations = info.iterations;
  converged = info.converged != 0;
  r0_l2 = info.r0_l2;
  r_l2 = info.r_l2;
  levels = info.mg_levels;
  last_solve_ms = info.solve_ms;
  if (completion_callback)
    completion_callback(converged, converged ? "Converged successfully" : "Failed to converge within maximum iterations");
  return x;
}

// ------------------------------------------------------------------------------------------------- matrix-free
MatrixFreeSystem::MatrixFreeSystem(int m_, int n_, double a_, double b_, double c_, double d_)
    : n(n_), m(m_), a(a_), b(b_), c(c_), d(d_) {
  plan_ = b200::make_geometric_plan(m, n, a, b, c, d);
  int64_t rows = 0;
  check(b200cg_size(plan_->get(), &rows));
  rhs.resize(static_cast<size_t>(rows));
  check(b200cg_build_rhs(plan_->get()));  // matrix_free_system.cpp:104-141 on the device
  check(b200cg_get_rhs(plan_->get(), rhs.data()));
}

std::vector<double> MatrixFreeSystem::get_true_solution_vector() {
  std::vector<double> u(rhs.size());
  check(b200cg_get_true_solution(plan_->get(), u.data()));
  return u;
}

void MatrixFreeSystem::apply(const std::vector<double>& x, std::vector<double>& y) const {
  if (x.size() != rhs.size()) throw std::invalid_argument("MatrixFreeSystem::apply: x has the wrong length");
  y.resize(rhs.size());
  check(b200cg_apply(plan_->get(), x.data(), y.data()));
}

std::ostream& operator<<(std::ostream& os, const MatrixFreeSystem& grid) {
  os << "MatrixFreeSystem Information:" << std::endl
     << "  Dimensions: " << grid.n << "x" << grid.m << std::endl
     << "  Domain: [" << grid.a << ", " << grid.b << "] x [" << grid.c << ", " << grid.d << "]" << std::endl
     << "  System size: " << grid.size() << std::endl
     << "  Memory savings: Matrix-free approach does not store matrix elements" << std::endl;
  return os;
}

MatrixFreeSolver::MatrixFreeSolver(const MatrixFreeSystem& system_, const std::vector<double>& b_, double eps_,
                                   int maxIterations_, const std::string& name_)
    : system(system_), b(b_), eps(eps_), maxIterations(maxIterations_), iterations(0), name(name_) {}

std::vector<double> MatrixFreeSolver::solve(const std::vector<double>& true_solution) {
  const size_t rows = static_cast<size_t>(system.size());
  if (b.size() != rows) throw std::invalid_argument("MatrixFreeSolver::solve: b has the wrong length");
  b200cg_params prm = {};
  prm.op = B200CG_OP_MATRIX_FREE;
  prm.rule = B200CG_RULE_REL_L2;
  prm.eps_rel = eps;
  prm.max_it = maxIterations;
  prm.preconditioner = multigrid ? B200CG_PRECOND_MULTIGRID : B200CG_PRECOND_NONE;
  b200cg_info info = {};
  std::vector<double> x(rows);
  b200::CallbackBox box{&iteration_callback};
  // A registered callback asks for the reference's per-iteration report (||dx||_2, recomputed ||b-Ax||_2,
  // ||x-u||_2, matrix_free_system.cpp:444-468); without one that extra work is skipped.
  const bool want_cb = static_cast<bool>(iteration_callback);
  const double* u = true_solution.size() == rows ? true_solution.data() : nullptr;
  check(b200cg_solve(system.plan()->get(), &prm, b.data(), u, x.data(), &info,
                     want_cb ? b200::callback_trampoline : nullptr, &box, nullptr));
  iterations = info.iterations;
  last_solve_ms = info.solve_ms;
  const bool ok = info.converged != 0;
  if (completion_callback)  // matrix_free_system.cpp:472-479
    completion_callback(ok, ok ? "Converged successfully" : "Failed to converge within maximum iterations");
  return x;
}

std::vector<std::vector<double>> MatrixFreeSolver::solveBatch(const std::vector<std::vector<double>>& right_hand_sides,
                                                             std::vector<int>* iterations_out) {
  const size_t rows = static_cast<size_t>(system.size()), count = right_hand_sides.size();
  std::vector<std::vector<double>> xs(count, std::vector<double>(rows));
  std::vector<const double*> b_ptrs(count);
  std::vector<double*> x_ptrs(count);
  for (size_t i = 0; i < count; ++i) {
    if (right_hand_sides[i].size() != rows) throw std::invalid_argument("MatrixFreeSolver::solveBatch: a right-hand side has the wrong length");
    b_ptrs[i] = right_hand_sides[i].data();
    x_ptrs[i] = xs[i].data();
  }
  b200cg_params prm = {};
  prm.op = B200CG_OP_MATRIX_FREE;
  prm.rule = B200CG_RULE_REL_L2;
  prm.eps_rel = eps;
  prm.max_it = maxIterations;
  prm.preconditioner = multigrid ? B200CG_PRECOND_MULTIGRID : B200CG_PRECOND_NONE;
  std::vector<b200cg_info> infos(count);
  check(b200cg_solve_batch(system.plan()->get(), &prm, static_cast<int>(count), b_ptrs.data(), x_ptrs.data(), infos.data(),
                           nullptr, nullptr, nullptr));
  if (iterations_out) iterations_out->assign(count, 0);
  for (size_t i = 0; i < count; ++i) {
    iterations = infos[i].iterations;
    last_solve_ms = infos[i].solve_ms;
    if (iterations_out) (*iterations_out)[i] = infos[i].iterations;
    const bool ok = infos[i].converged != 0;
    if (completion_callback)
      completion_callback(ok, ok ? "Converged successfully" : "Failed to converge within maximum iterations");
  }
  return xs;
}

// ------------------------------------------------------------------------------------------------- facade
DirichletSolver::DirichletSolver(int n, int m, double a, double b, double c, double d)
    : n_internal(n), m_internal(m), a_bound(a), b_bound(b), c_bound(c), d_bound(d),
      eps_precision(1e-6), eps_residual(1e-6), eps_exact_error(1e-6), max_iterations(10000),
      use_precision_stopping(true), use_residual_stopping(true), use_error_stopping(false),
      use_max_iterations_stopping(true) {
  if (!Kokkos::is_initialized()) Kokkos::initialize();
  grid = std::make_unique<GridSystem>(m_internal, n_internal, a_bound, b_bound, c_bound, d_bound);
}

DirichletSolver::~DirichletSolver() {
  solver.reset();
  grid.reset();
}

void DirichletSolver::setGridParameters(int n, int m, double a, double b, double c, double d) {
  n_internal = n;
  m_internal = m;
  a_bound = a;
  b_bound = b;
  c_bound = c;
  d_bound = d;
  solver.reset();  // it refers to the old grid's matrix
  grid = std::make_unique<GridSystem>(m_internal, n_internal, a_bound, b_bound, c_bound, d_bound);
}

void DirichletSolver::setSolverParameters(double eps_p, double eps_r, double eps_e, int max_iter) {
  eps_precision = eps_p;
  eps_residual = eps_r;
  eps_exact_error = eps_e;
  max_iterations = max_iter;
}

void DirichletSolver::setIterationCallback(std::function<void(int, double, double, double)> callback) {
  iteration_callback = callback;
}

SolverResults DirichletSolver::solve() {
  if (!grid) throw std::runtime_error("Сетка не инициализирована");
  // The operator: the grid's plan applies the 5-point stencil on the fly - the same matrix GridSystem assembles, bit for
  // bit, at 40 % of the traffic - unless B200CG_DIRICHLET_OPERATOR=csr asks for the assembled arrays (then the matrix
  // is assembled on the device; neither way needs its host mirror).
  const char* op_env = std::getenv("B200CG_DIRICHLET_OPERATOR");
  const bool matrix_free = !(op_env && std::string(op_env) == "csr");
  if (!matrix_free) grid->ensure_device_matrix();
  solver = std::make_unique<MSGSolver>(grid->matrix_handle(), grid->get_rhs(),
                                       std::min({eps_precision, eps_residual, eps_exact_error}), max_iterations);
  solver->attachPlan(grid->plan(), matrix_free);
  // a disabled rule is passed as -1 (dirichlet_solver.cpp:71-87)
  solver->setPrecisionEps(use_precision_stopping ? eps_precision : -1.0);
  solver->setResidualEps(use_residual_stopping ? eps_residual : -1.0);
  solver->setExactErrorEps(use_error_stopping ? eps_exact_error : -1.0);
  if (iteration_callback) solver->setIterationCallback(iteration_callback);
  if (stop_requested) stop_requested = false;

  true_solution = grid->get_true_solution_vector();
  solution = solver->solve(true_solution);

  SolverResults results;
  const size_t rows = solution.extent(0);
  results.solution.assign(solution.data(), solution.data() + rows);
  results.true_solution.assign(true_solution.data(), true_solution.data() + rows);
  results.residual.resize(rows);
  results.error.resize(rows);
  // A x - b and x - u on the device (dirichlet_solver.cpp:147-180)
  check(b200cg_postprocess(grid->plan()->get(), matrix_free ? B200CG_OP_MATRIX_FREE : B200CG_OP_CSR, results.residual.data(),
                           results.error.data()));
  results.x_coords = grid->get_x_coords();
  results.y_coords = grid->get_y_coords();
  results.iterations = solver->getIterations();
  results.converged = solver->hasConverged();
  results.stop_reason = solver->getStopReasonText();
  results.residual_norm = solver->getFinalResidualNorm();
  results.error_norm = solver->getFinalErrorNorm();
  results.precision = solver->getFinalPrecision();
  last_results = results;
  if (completion_callback) completion_callback(results);
  return results;
}

std::vector<double> DirichletSolver::getSolution() const {
  return std::vector<double>(solution.data(), solution.data() + solution.extent(0));
}

std::vector<double> DirichletSolver::getTrueSolution() const {
  return std::vector<double>(true_solution.data(), true_solution.data() + true_solution.extent(0));
}

std::vector<std::vector<double>> DirichletSolver::solutionToMatrix() const {
  // Node-grid view of the solution: m-1 rows of n-1 unknown columns, zeros where the L-shaped region has no
  // unknown. (The reference indexes an n*m array, which overruns on the L-shape: dirichlet_solver.cpp:193-205.)
  const int cols = std::max(n_internal - 1, 0), rows = std::max(m_internal - 1, 0);
  std::vector<std::vector<double>> out(static_cast<size_t>(rows), std::vector<double>(static_cast<size_t>(cols), 0.0));
  if (!grid || solution.extent(0) == 0) return out;
  const std::vector<double>& xs = grid->get_x_coords();
  const std::vector<double>& ys = grid->get_y_coords();
  const double hx = (b_bound - a_bound) / n_internal, hy = (d_bound - c_bound) / m_internal;
  for (size_t k = 0; k < solution.extent(0); ++k) {
    const long i = std::lround((xs[k] - a_bound) / hx) - 1, j = std::lround((ys[k] - c_bound) / hy) - 1;
    if (i >= 0 && i < cols && j >= 0 && j < rows) out[static_cast<size_t>(j)][static_cast<size_t>(i)] = solution(k);
  }
  return out;
}

std::string DirichletSolver::generateReport() const {
  if (!solver) return "Решение еще не выполнено";
  return solver->generateReport(n_internal, m_internal, a_bound, b_bound, c_bound, d_bound);
}

bool DirichletSolver::saveResultsToFile(const std::string& filename) const {
  if (!solver) return false;
  return ResultsIO::saveResults(filename, last_results, n_internal, m_internal, a_bound, b_bound, c_bound, d_bound,
                                solver->getName());
}

bool DirichletSolver::saveMatrixAndRhsToFile(const std::string& filename) const {
  if (!grid) return false;
  return ResultsIO::saveMatrixAndRhs(filename, grid->get_matrix(), grid->get_rhs(), n_internal, m_internal);
}

// ------------------------------------------------------------------------------------------------- ResultsIO
namespace {
void write_section(std::ofstream& f, const char* title, const std::vector<double>& v) {
  f << title << "\n" << std::scientific;
  for (double value : v) f << value << "\n";
}

// Reads doubles until the next section title (or EOF); returns the title it stopped at ("" at EOF).
std::string read_section(std::ifstream& f, std::vector<double>& v) {
  v.clear();
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    const char first = line[0];
    if ((first >= '0' && first <= '9') || first == '-' || first == '+' || first == '.' || first == 'n' || first == 'i') {
      v.push_back(std::strtod(line.c_str(), nullptr));
    } else {
      return line;
    }
  }
  return std::string();
}
}  // namespace

bool ResultsIO::saveResults(const std::string& filename, const SolverResults& r, int n, int m, double a, double b,
                            double c, double d, const std::string& solver_name) {
  std::ofstream f(filename);
  if (!f) return false;
  f << "PARAMETERS\n" << n << " " << m << "\n" << a << " " << b << " " << c << " " << d << "\n" << solver_name << "\n";
  f << "CONVERGENCE\n" << r.iterations << "\n" << (r.converged ? "1" : "0") << "\n" << r.stop_reason << "\n"
    << std::scientific << r.residual_norm << " " << r.error_norm << "\n";
  write_section(f, "SOLUTION", r.solution);
  write_section(f, "TRUE_SOLUTION", r.true_solution);
  write_section(f, "RESIDUAL", r.residual);
  write_section(f, "ERROR", r.error);
  write_section(f, "X_COORDS", r.x_coords);
  write_section(f, "Y_COORDS", r.y_coords);
  return static_cast<bool>(f);
}

bool ResultsIO::loadResults(const std::string& filename, SolverResults& r, int& n, int& m, double& a, double& b,
                            double& c, double& d, std::string& solver_name) {
  std::ifstream f(filename);
  if (!f) return false;
  std::string line;
  if (!std::getline(f, line) || line != "PARAMETERS") return false;
  f >> n >> m >> a >> b >> c >> d;
  f.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
  std::getline(f, solver_name);
  if (!std::getline(f, line) || line != "CONVERGENCE") return false;
  int conv = 0;
  f >> r.iterations >> conv;
  r.converged = (conv == 1);
  f.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
  std::getline(f, r.stop_reason);
  f >> r.residual_norm >> r.error_norm;
  f.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
  if (!std::getline(f, line) || line != "SOLUTION") return false;
  std::string next = read_section(f, r.solution);
  if (next != "TRUE_SOLUTION") return false;
  next = read_section(f, r.true_solution);
  if (next != "RESIDUAL") return false;
  next = read_section(f, r.residual);
  if (next != "ERROR") return false;
  next = read_section(f, r.error);
  if (next == "X_COORDS") next = read_section(f, r.x_coords);  // optional sections (dirichlet_solver.cpp:388-402)
  if (next == "Y_COORDS") read_section(f, r.y_coords);
  return true;
}

bool ResultsIO::saveMatrixAndRhs(const std::string& filename, const KokkosCrsMatrix& A, const KokkosVector& b, int n,
                                 int m) {
  std::ofstream f(filename);
  if (!f) return false;
  const int rows = A.numRows(), nnz = A.nnz();
  f << "MATRIX_INFO\n" << n << " " << m << "\n" << rows << " " << nnz << "\n";
  f << "MATRIX\n";
  for (int i = 0; i <= rows; ++i) f << A.graph.row_map(static_cast<size_t>(i)) << "\n";
  for (int k = 0; k < nnz; ++k) f << A.graph.entries(static_cast<size_t>(k)) << "\n";
  f << std::scientific;
  for (int k = 0; k < nnz; ++k) f << A.values(static_cast<size_t>(k)) << "\n";
  f << "RHS\n";
  for (int i = 0; i < rows; ++i) f << b(static_cast<size_t>(i)) << "\n";
  return static_cast<bool>(f);
}

bool ResultsIO::saveSolutionFor3D(const std::string& filename, const std::vector<std::vector<double>>& solution,
                                  double a_bound, double b_bound, double c_bound, double d_bound) {
  std::ofstream f(filename);
  if (!f.is_open()) return false;
  const size_t rows = solution.size();
  if (rows == 0) return false;
  const size_t cols = solution[0].size();
  if (cols == 0) return false;
  // gnuplot "x y z" blocks, one blank line between grid rows (dirichlet_solver.hpp:59-72)
  const double hx = (b_bound - a_bound) / (cols + 1), hy = (d_bound - c_bound) / (rows + 1);
  for (size_t j = 0; j < rows; ++j) {
    for (size_t i = 0; i < cols; ++i)
      f << a_bound + (i + 1) * hx << " " << c_bound + (j + 1) * hy << " " << solution[j][i] << std::endl;
    f << std::endl;
  }
  return true;
}
