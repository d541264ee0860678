// The binding INTEGRATION.md section 2 shows a maintainer (kept compiling by tests/test_integration_snippets.py):
// the reference's MatrixFreeSystem / MatrixFreeSolver / MSGSolver keep their interfaces and forward the hot calls to
// the C ABI. Minimal stand-ins for the reference's own types are declared here; only the b200cg_* usage matters.
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200cg.h"

class MatrixFreeSystem {
 public:
  MatrixFreeSystem(int m, int n, double a, double b, double c, double d) {  // matrix_free_system.cpp:144-159
    b200cg_plan_desc desc = {};
    desc.n = n; desc.m = m; desc.a = a; desc.b = b; desc.c = c; desc.d = d;
    desc.domain = B200CG_DOMAIN_LSHAPE; desc.device = 0; desc.world = 1;
    if (b200cg_plan_create(&plan_, &desc)) throw std::runtime_error(b200cg_last_error());
    int64_t N = 0;
    b200cg_size(plan_, &N);
    rhs.resize(static_cast<size_t>(N));
    b200cg_build_rhs(plan_);
    b200cg_get_rhs(plan_, rhs.data());  // replaces initialize_rhs :104-141
  }
  ~MatrixFreeSystem() { b200cg_plan_destroy(plan_); }
  void apply(const std::vector<double>& x, std::vector<double>& y) const {  // :203-340
    y.resize(x.size());
    if (b200cg_apply(plan_, x.data(), y.data())) throw std::runtime_error(b200cg_last_error());
  }
  size_t size() const { return rhs.size(); }
  std::vector<double> rhs;
  b200cg_plan_t plan_ = nullptr;
};

class MatrixFreeSolver {
 public:
  using IterationCallback = std::function<void(int, double, double, double)>;
  double eps = 1e-6;
  int maxIterations = 10000, iterations = 0;
  IterationCallback iteration_callback;
  std::function<void(bool, const std::string&)> completion_callback;

  std::vector<double> solve(const MatrixFreeSystem& system, const std::vector<double>& b,
                            const std::vector<double>& true_solution) {  // :383-482
    b200cg_params prm = {};
    prm.op = B200CG_OP_MATRIX_FREE;
    prm.rule = B200CG_RULE_REL_L2;
    prm.eps_rel = eps;
    prm.max_it = maxIterations;
    prm.single_sweep = 0;  // the plan's default: one sweep per iteration (2: the two-sweep iteration)
    b200cg_info info;
    std::vector<double> x(system.size());
    auto tramp = [](void* u, int it, double p, double r, double e) { (*static_cast<IterationCallback*>(u))(it, p, r, e); };
    if (b200cg_solve(system.plan_, &prm, b.data(), true_solution.empty() ? nullptr : true_solution.data(), x.data(), &info,
                     iteration_callback ? +tramp : nullptr, &iteration_callback, nullptr))
      throw std::runtime_error(b200cg_last_error());
    iterations = info.iterations;
    if (completion_callback)
      completion_callback(info.converged != 0, info.converged ? "Converged successfully"
                                                              : "Failed to converge within maximum iterations");
    return x;
  }

  // many right-hand sides (time steps, parameter sweeps): solve() once per entry, host copies under the iterations
  std::vector<std::vector<double>> solveBatch(const MatrixFreeSystem& system, const std::vector<std::vector<double>>& rhs) {
    b200cg_params prm = {};
    prm.op = B200CG_OP_MATRIX_FREE;
    prm.rule = B200CG_RULE_REL_L2;
    prm.eps_rel = eps;
    prm.max_it = maxIterations;
    std::vector<std::vector<double>> xs(rhs.size(), std::vector<double>(system.size()));
    std::vector<const double*> bp;
    std::vector<double*> xp;
    for (size_t i = 0; i < rhs.size(); ++i) {
      bp.push_back(rhs[i].data());
      xp.push_back(xs[i].data());
    }
    std::vector<b200cg_info> infos(rhs.size());
    auto done = [](void* u, int i, const b200cg_info* info) { static_cast<MatrixFreeSolver*>(u)->iterations = info->iterations + 0 * i; };
    if (b200cg_solve_batch(system.plan_, &prm, static_cast<int>(rhs.size()), bp.data(), xp.data(), infos.data(), +done, this,
                           nullptr))
      throw std::runtime_error(b200cg_last_error());
    return xs;
  }
};

// MSGSolver::solve on an assembled matrix (msg_solver.cpp:10-212)
int msg_solve_csr(int64_t rows, int64_t nnz, const int* row_map, const int* entries, const double* values, const double* b,
                  const double* u, double* x, double eps_p, double eps_r, double eps_e, int max_it,
                  const volatile int* stop_requested, b200cg_info* info) {
  b200cg_plan_desc desc = {};
  desc.domain = B200CG_DOMAIN_GENERIC;
  desc.generic_rows = rows;
  desc.world = 1;
  b200cg_plan_t plan = nullptr;
  if (int rc = b200cg_plan_create(&plan, &desc)) return rc;
  int rc = b200cg_set_csr(plan, rows, nnz, row_map, entries, values);
  if (!rc) {
    b200cg_params prm = {};
    prm.op = B200CG_OP_CSR;
    prm.rule = B200CG_RULE_MAXNORM;
    prm.eps_p = eps_p; prm.eps_r = eps_r; prm.eps_e = eps_e;  // <= 0 disables a rule (dirichlet_solver.cpp:71-87)
    prm.max_it = max_it;
    prm.callback_every = 100;
    rc = b200cg_solve(plan, &prm, b, u, x, info, nullptr, nullptr, stop_requested);
  }
  b200cg_plan_destroy(plan);
  return rc;
}

int main() { return b200cg_version() > 0 ? 0 : 1; }
