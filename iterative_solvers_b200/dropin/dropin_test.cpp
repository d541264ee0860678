// Test driver for the drop-in classes: exercises the reference's public surface exactly as its callers do
// (qt_gui/src/mainwindow.cpp:185-266,290-308; solver/main.cpp:596-712) and dumps the results for pytest
// (tests/test_dropin_gpu.py) to compare with the golden fixtures. Usage: dropin_test <command> <args...> <outdir>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "dirichlet_solver.hpp"
#include "matrix_free_system.hpp"

namespace {
std::string g_out;

template <class T>
void dump(const std::string& name, const T* data, size_t count) {
  std::ofstream f(g_out + "/" + name, std::ios::binary);
  f.write(reinterpret_cast<const char*>(data), static_cast<std::streamsize>(count * sizeof(T)));
}
void dump(const std::string& name, const std::vector<double>& v) { dump(name, v.data(), v.size()); }

std::vector<double> slurp(const std::string& name) {
  std::ifstream f(g_out + "/" + name, std::ios::binary | std::ios::ate);
  std::vector<double> v;
  if (!f) return v;
  v.resize(static_cast<size_t>(f.tellg()) / sizeof(double));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(double)));
  return v;
}

int cmd_dirichlet(char** a) {
  const int n = std::atoi(a[0]), m = std::atoi(a[1]);
  const double lo_x = std::atof(a[2]), hi_x = std::atof(a[3]), lo_y = std::atof(a[4]), hi_y = std::atof(a[5]);
  DirichletSolver solver(n, m, lo_x, hi_x, lo_y, hi_y);
  solver.setSolverParameters(std::atof(a[6]), std::atof(a[7]), std::atof(a[8]), std::atoi(a[9]));
  solver.enablePrecisionStopping(std::atoi(a[10]) != 0);
  solver.enableResidualStopping(std::atoi(a[11]) != 0);
  solver.enableErrorStopping(std::atoi(a[12]) != 0);
  std::ofstream cb(g_out + "/callbacks.txt");
  cb.precision(17);
  solver.setIterationCallback([&](int it, double p, double r, double e) { cb << it << " " << p << " " << r << " " << e << "\n"; });
  int completions = 0;
  solver.setCompletionCallback([&](const SolverResults&) { ++completions; });
  SolverResults res = solver.solve();
  dump("solution.bin", res.solution);
  dump("true_solution.bin", res.true_solution);
  dump("residual.bin", res.residual);
  dump("error.bin", res.error);
  dump("x_coords.bin", res.x_coords);
  dump("y_coords.bin", res.y_coords);
  std::ofstream(g_out + "/report.txt") << solver.generateReport();
  const bool saved = solver.saveResultsToFile(g_out + "/results.txt");
  const bool saved_matrix = solver.saveMatrixAndRhsToFile(g_out + "/matrix.txt");
  SolverResults back;
  int n2 = 0, m2 = 0;
  double p2[4] = {0, 0, 0, 0};
  std::string name2;
  const bool loaded = ResultsIO::loadResults(g_out + "/results.txt", back, n2, m2, p2[0], p2[1], p2[2], p2[3], name2);
  const bool roundtrip = loaded && n2 == n && m2 == m && back.solution.size() == res.solution.size() &&
                         back.iterations == res.iterations && back.converged == res.converged &&
                         back.stop_reason == res.stop_reason && back.y_coords.size() == res.y_coords.size();
  double worst = 0.0;
  if (roundtrip)
    for (size_t i = 0; i < res.solution.size(); ++i)
      worst = std::max(worst, std::abs(back.solution[i] - res.solution[i]) / (std::abs(res.solution[i]) + 1e-300));
  std::vector<std::vector<double>> mat = solver.solutionToMatrix();
  const bool saved3d = ResultsIO::saveSolutionFor3D(g_out + "/surface.txt", mat, lo_x, hi_x, lo_y, hi_y);
  std::ofstream info(g_out + "/info.txt");
  info.precision(17);
  info << "iterations=" << res.iterations << "\nconverged=" << res.converged << "\nresidual_norm=" << res.residual_norm
       << "\nerror_norm=" << res.error_norm << "\nprecision=" << res.precision << "\nstop_reason=" << res.stop_reason
       << "\nmethod=" << solver.getMethodName() << "\ncompletions=" << completions << "\nsaved=" << saved
       << "\nsaved_matrix=" << saved_matrix << "\nio_roundtrip=" << roundtrip << "\nio_worst_rel=" << worst
       << "\nsaved3d=" << saved3d << "\nmatrix_rows=" << mat.size() << "\nmatrix_cols=" << (mat.empty() ? 0 : mat[0].size())
       << "\nsize=" << res.solution.size() << "\n";
  return 0;
}

int cmd_mf(char** a) {
  const int n = std::atoi(a[0]);
  const double lo = std::atof(a[1]), hi = std::atof(a[2]), eps = std::atof(a[3]);
  const int max_it = std::atoi(a[4]), with_cb = std::atoi(a[5]);
  MatrixFreeSystem system(n, n, lo, hi, lo, hi);
  dump("rhs.bin", system.get_rhs());
  std::vector<double> u = system.get_true_solution_vector();
  dump("true.bin", u);
  std::vector<double> vin = slurp("apply_in.bin");
  if (static_cast<int>(vin.size()) == system.size()) {
    dump("apply_out.bin", system * vin);  // operator* -> apply
  }
  MatrixFreeSolver solver(system, system.get_rhs(), eps, max_it);
  std::ofstream hist(g_out + "/hist.txt");
  hist.precision(17);
  if (with_cb)
    solver.setIterationCallback([&](int it, double p, double r, double e) { hist << it << " " << p << " " << r << " " << e << "\n"; });
  bool converged = false;
  std::string message;
  solver.setCompletionCallback([&](bool ok, const std::string& msg) { converged = ok; message = msg; });
  std::vector<double> x = solver.solve(u);
  dump("x.bin", x);
  std::ostringstream desc;
  desc << system;
  std::ofstream info(g_out + "/info.txt");
  info.precision(17);
  info << "iterations=" << solver.getIterations() << "\nconverged=" << converged << "\nmessage=" << message
       << "\nname=" << solver.getName() << "\nsize=" << system.size() << "\nsolve_ms=" << solver.lastSolveMilliseconds()
       << "\ndescribes_size=" << (desc.str().find("System size: " + std::to_string(system.size())) != std::string::npos) << "\n";
  return 0;
}

// MatrixFreeSolver::solveBatch (B200 addition): three scaled copies of the system's rhs through the queue, next to solve()
int cmd_mfbatch(char** a) {
  const int n = std::atoi(a[0]);
  const double lo = std::atof(a[1]), hi = std::atof(a[2]), eps = std::atof(a[3]);
  const int max_it = std::atoi(a[4]);
  MatrixFreeSystem system(n, n, lo, hi, lo, hi);
  std::vector<std::vector<double>> rhs(3, system.get_rhs());
  for (double& v : rhs[1]) v *= -2.0;
  for (double& v : rhs[2]) v *= 0.5;
  MatrixFreeSolver solver(system, system.get_rhs(), eps, max_it);
  int completions = 0;
  solver.setCompletionCallback([&](bool ok, const std::string&) { completions += ok ? 1 : 100; });
  std::vector<int> its;
  std::vector<std::vector<double>> xs = solver.solveBatch(rhs, &its);
  std::vector<double> single = solver.solve(std::vector<double>());
  dump("x0.bin", xs[0]);
  dump("x1.bin", xs[1]);
  dump("x2.bin", xs[2]);
  dump("x_single.bin", single);
  std::ofstream info(g_out + "/info.txt");
  info << "iterations=" << its[0] << " " << its[1] << " " << its[2] << "\nsingle_iterations=" << solver.getIterations()
       << "\ncompletions=" << completions << "\n";
  return 0;
}

// the opt-in multigrid-preconditioned CG through both doors: the Solver subclass on a GridSystem and the switch of
// MatrixFreeSolver
int cmd_mgpcg(char** a) {
  const int n = std::atoi(a[0]);
  const double lo = std::atof(a[1]), hi = std::atof(a[2]), eps = std::atof(a[3]);
  const int max_it = std::atoi(a[4]);
  GridSystem grid(n, n, lo, hi, lo, hi);
  MultigridPCGSolver solver(grid, eps, max_it);
  int completions = 0;
  solver.setCompletionCallback([&](bool, const std::string&) { ++completions; });
  Solver& base = solver;  // used through the reference's abstract interface
  KokkosVector x = base.solve(KokkosVector());
  dump("x.bin", x.data(), x.extent(0));
  dump("rhs.bin", grid.get_rhs().data(), grid.get_rhs().extent(0));
  MatrixFreeSystem system(n, n, lo, hi, lo, hi);
  MatrixFreeSolver mf(system, system.get_rhs(), eps, max_it);
  mf.enableMultigridPreconditioner(true);
  std::vector<double> x2 = mf.solve(std::vector<double>());
  dump("x_mf.bin", x2);
  std::ofstream info(g_out + "/info.txt");
  info.precision(17);
  info << "iterations=" << base.getIterations() << "\nconverged=" << solver.hasConverged() << "\nlevels=" << solver.multigridLevels()
       << "\nr0=" << solver.getInitialResidualNorm() << "\nr=" << solver.getFinalResidualNorm() << "\nname=" << base.getName()
       << "\ncompletions=" << completions << "\nmf_iterations=" << mf.getIterations() << "\n";
  return 0;
}

int cmd_grid(char** a) {
  const int n = std::atoi(a[0]);
  const double lo = std::atof(a[1]), hi = std::atof(a[2]);
  GridSystem grid(n, n, lo, hi, lo, hi);
  const KokkosCrsMatrix& A = grid.get_matrix();
  dump("row_map.bin", A.graph.row_map.data(), static_cast<size_t>(A.numRows()) + 1);
  dump("entries.bin", A.graph.entries.data(), static_cast<size_t>(A.nnz()));
  dump("values.bin", A.values.data(), static_cast<size_t>(A.nnz()));
  dump("rhs.bin", grid.get_rhs().data(), grid.get_rhs().extent(0));
  dump("xs.bin", grid.get_x_coords());
  dump("ys.bin", grid.get_y_coords());
  KokkosVector u = grid.get_true_solution_vector();
  dump("true.bin", u.data(), u.extent(0));
  // MSGSolver as a stand-alone solver: only the matrix and the rhs, no geometry (generic plan)
  MSGSolver solver(A, grid.get_rhs(), 1e-6, std::atoi(a[6]));
  solver.setPrecisionEps(std::atof(a[3]));
  solver.setResidualEps(std::atof(a[4]));
  solver.setExactErrorEps(std::atof(a[5]));
  KokkosVector x = solver.solve(u);
  dump("x.bin", x.data(), x.extent(0));
  // KokkosSparse::spmv through the device
  KokkosVector y("y", x.extent(0));
  KokkosSparse::spmv("N", 1.0, A, x, 0.0, y);
  dump("Ax.bin", y.data(), y.extent(0));
  GridSystem::NodeCoordinates first = grid.get_node_coordinates(0), bad = grid.get_node_coordinates(-5);
  std::ostringstream desc;
  desc << grid;
  std::ofstream info(g_out + "/info.txt");
  info.precision(17);
  info << "rows=" << A.numRows() << "\nnnz=" << A.nnz() << "\niterations=" << solver.getIterations()
       << "\nconverged=" << solver.hasConverged() << "\nstop=" << static_cast<int>(solver.getStopReason())
       << "\nr_max=" << solver.getFinalResidualNorm() << "\ndx_max=" << solver.getFinalPrecision()
       << "\nerr_max=" << solver.getFinalErrorNorm() << "\nfirst_x=" << first.x << "\nfirst_y=" << first.y
       << "\nbad_x=" << bad.x << "\nbad_y=" << bad.y
       << "\ndescribes_nnz=" << (desc.str().find("Non-zero elements: " + std::to_string(A.nnz())) != std::string::npos) << "\n";
  return 0;
}

int cmd_errors(char**) {
  int caught = 0;
  try {
    GridSystem bad(6, 8, 0, 1, 0, 1);  // n != m: the reference numbering is malformed there
  } catch (const std::invalid_argument&) {
    ++caught;
  }
  try {
    MatrixFreeSystem odd(7, 7, 0, 1, 0, 1);
  } catch (const std::exception&) {
    ++caught;
  }
  try {
    MatrixFreeSystem ok(6, 6, 1, 2, 1, 2);
    std::vector<double> wrong(3), y;
    ok.apply(wrong, y);
  } catch (const std::invalid_argument&) {
    ++caught;
  }
  std::ofstream(g_out + "/info.txt") << "caught=" << caught << "\n";
  return caught == 3 ? 0 : 1;
}

int cmd_io(char**) {
  // ResultsIO without a device: save / load round trip of an L-shaped result (section lengths != n*m), a file without
  // the optional coordinate sections (dirichlet_solver.cpp:388-402), a truncated file, the gnuplot surface format
  SolverResults r;
  const int n = 6, m = 6, count = 16;  // 16 unknowns on the 6x6 L-shaped grid
  for (int i = 0; i < count; ++i) {
    r.solution.push_back(1.0 + 0.125 * i);
    r.true_solution.push_back(1.0 + 0.125 * i + 1e-7);
    r.residual.push_back((i % 2 ? -1.0 : 1.0) * 3.5e-9 * (i + 1));
    r.error.push_back(-1e-7);
    r.x_coords.push_back(0.5 + i / 12.0);
    r.y_coords.push_back(1.0 / 6.0 * (1 + i / 4));
  }
  r.residual_norm = 5.6e-8;
  r.error_norm = 1e-7;
  r.iterations = 13;
  r.converged = true;
  r.stop_reason = "Converged by residual";
  const std::string path = g_out + "/roundtrip.txt";
  int failures = 0;
  failures += ResultsIO::saveResults(path, r, n, m, 0.0, 1.0, 2.0, 3.0, "MSG Solver") ? 0 : 1;
  SolverResults back;
  int n2 = 0, m2 = 0;
  double p[4] = {0, 0, 0, 0};
  std::string name;
  failures += ResultsIO::loadResults(path, back, n2, m2, p[0], p[1], p[2], p[3], name) ? 0 : 1;
  auto close = [](const std::vector<double>& x, const std::vector<double>& y) {
    if (x.size() != y.size()) return false;
    for (size_t i = 0; i < x.size(); ++i)
      if (std::abs(x[i] - y[i]) > 1e-6 * std::abs(y[i])) return false;  // the text format keeps 7 digits
    return true;
  };
  failures += (n2 == n && m2 == m && p[0] == 0.0 && p[1] == 1.0 && p[2] == 2.0 && p[3] == 3.0) ? 0 : 1;
  failures += (name == "MSG Solver" && back.stop_reason == r.stop_reason && back.iterations == 13 && back.converged) ? 0 : 1;
  failures += (close(back.solution, r.solution) && close(back.true_solution, r.true_solution) &&
               close(back.residual, r.residual) && close(back.error, r.error) && close(back.x_coords, r.x_coords) &&
               close(back.y_coords, r.y_coords)) ? 0 : 1;
  failures += (std::abs(back.residual_norm - r.residual_norm) <= 1e-6 * r.residual_norm) ? 0 : 1;

  // a file written by an older writer: no coordinate sections
  {
    std::ifstream in(path);
    std::ofstream out(g_out + "/no_coords.txt");
    std::string line;
    while (std::getline(in, line)) {
      if (line == "X_COORDS") break;
      out << line << "\n";
    }
  }
  SolverResults short_back;
  failures += ResultsIO::loadResults(g_out + "/no_coords.txt", short_back, n2, m2, p[0], p[1], p[2], p[3], name) ? 0 : 1;
  failures += (short_back.solution.size() == (size_t)count && short_back.error.size() == (size_t)count &&
               short_back.x_coords.empty() && short_back.y_coords.empty()) ? 0 : 1;

  // broken inputs are refused, not half-read
  {
    std::ofstream(g_out + "/truncated.txt") << "PARAMETERS\n6 6\n0 1 2 3\nMSG Solver\nCONVERGENCE\n13\n1\nok\n1e-8 1e-7\nSOLUTION\n1.0\n";
    std::ofstream(g_out + "/garbage.txt") << "hello\n";
  }
  SolverResults bad;
  failures += ResultsIO::loadResults(g_out + "/truncated.txt", bad, n2, m2, p[0], p[1], p[2], p[3], name) ? 1 : 0;
  failures += ResultsIO::loadResults(g_out + "/garbage.txt", bad, n2, m2, p[0], p[1], p[2], p[3], name) ? 1 : 0;
  failures += ResultsIO::loadResults(g_out + "/does_not_exist.txt", bad, n2, m2, p[0], p[1], p[2], p[3], name) ? 1 : 0;

  // gnuplot surface: "x y z" per node, blank line between grid rows, interior nodes of [0,1]x[2,3]
  std::vector<std::vector<double>> grid = {{1.0, 2.0, 3.0}, {4.0, 5.0, 6.0}};
  failures += ResultsIO::saveSolutionFor3D(g_out + "/surface.txt", grid, 0.0, 1.0, 2.0, 3.0) ? 0 : 1;
  failures += ResultsIO::saveSolutionFor3D(g_out + "/empty.txt", {}, 0.0, 1.0, 2.0, 3.0) ? 1 : 0;
  std::ofstream(g_out + "/info.txt") << "failures=" << failures << "\n";
  return failures == 0 ? 0 : 1;
}

int cmd_stop(char** a) {
  // requestStop() from another thread while solve() runs on this one (mainwindow.cpp:268-288)
  const int n = std::atoi(a[0]);
  DirichletSolver solver(n, n, 0, 1, 0, 1);
  solver.setSolverParameters(1e-300, 1e-300, 1e-300, 2000000000);
  solver.enablePrecisionStopping(false);  // only the iteration cap remains: the solve runs until it is stopped
  solver.enableResidualStopping(false);
  solver.enableErrorStopping(false);
  std::thread stopper([&] {
    std::this_thread::sleep_for(std::chrono::milliseconds(300));
    solver.requestStop();
  });
  SolverResults res = solver.solve();
  stopper.join();
  std::ofstream(g_out + "/info.txt") << "iterations=" << res.iterations << "\nconverged=" << res.converged
                                     << "\nstop_reason=" << res.stop_reason << "\n";
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: dropin_test <dirichlet|mf|mfbatch|grid|mgpcg|errors|io|stop> <args...> <outdir>\n");
    return 2;
  }
  g_out = argv[argc - 1];
  const std::string cmd = argv[1];
  try {
    if (cmd == "dirichlet" && argc == 16) return cmd_dirichlet(argv + 2);
    if (cmd == "mf" && argc == 9) return cmd_mf(argv + 2);
    if (cmd == "grid" && argc == 10) return cmd_grid(argv + 2);
    if (cmd == "mgpcg" && argc == 8) return cmd_mgpcg(argv + 2);
    if (cmd == "mfbatch" && argc == 8) return cmd_mfbatch(argv + 2);
    if (cmd == "errors") return cmd_errors(argv + 2);
    if (cmd == "io" && argc == 3) return cmd_io(argv + 2);
    if (cmd == "stop" && argc == 4) return cmd_stop(argv + 2);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "dropin_test: %s\n", e.what());
    return 3;
  }
  std::fprintf(stderr, "dropin_test: bad command line\n");
  return 2;
}
