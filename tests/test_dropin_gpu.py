"""The C++ drop-in classes (iterative_solvers_b200/dropin: DirichletSolver, GridSystem, MSGSolver,
MatrixFreeSystem, MatrixFreeSolver, ResultsIO) driven exactly as the reference's callers drive them, through the
compiled test driver; results are compared with the fixtures produced by the unmodified reference classes."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL = 1e-10


@pytest.fixture(scope="module")
def driver():
    from iterative_solvers_b200 import build

    build.build_all()
    assert os.path.exists(build.DROPIN_TEST)
    return build.DROPIN_TEST


def run(driver, *args, expect=0, env=None):
    proc = subprocess.run([driver, *map(str, args)], capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, **env) if env else None)
    assert proc.returncode == expect, f"{args}: rc={proc.returncode}\n{proc.stdout}\n{proc.stderr}"
    return proc


def info(outdir):
    out = {}
    for line in open(os.path.join(outdir, "info.txt"), encoding="utf-8"):
        if "=" in line:
            k, v = line.rstrip("\n").split("=", 1)
            out[k] = v
    return out


def f64(outdir, name):
    return np.fromfile(os.path.join(outdir, name), dtype=np.float64)


def relmax(x, ref):
    return np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300)


# ---------------------------------------------------------------- CPU: the boundary exists and fails loudly
def test_driver_builds_and_rejects_bad_command_lines(driver, tmp_path):
    run(driver, "nonsense", tmp_path, expect=2)


def test_cpp_surface_fails_loudly_without_a_gpu(driver, tmp_path):
    from iterative_solvers_b200 import capi

    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    proc = run(driver, "mf", 6, 1, 2, 1e-8, 100, 0, tmp_path, expect=3)
    assert "no" in proc.stderr.lower() and "device" in proc.stderr.lower()


def test_results_io_formats_on_the_host(driver, tmp_path):
    """ResultsIO (dirichlet_solver.cpp:330-470) needs no device: save/load round trip with L-shape section lengths,
    files without the optional coordinate sections, refused broken files, the gnuplot surface layout."""
    run(driver, "io", tmp_path)
    assert info(tmp_path)["failures"] == "0"
    text = open(os.path.join(tmp_path, "roundtrip.txt"), encoding="utf-8").read().split("\n")
    assert text[0] == "PARAMETERS" and text[1] == "6 6" and text[3] == "MSG Solver" and text[4] == "CONVERGENCE"
    for title in ("SOLUTION", "TRUE_SOLUTION", "RESIDUAL", "ERROR", "X_COORDS", "Y_COORDS"):
        i = text.index(title)
        float(text[i + 1]) and float(text[i + 16])  # 16 values per section, one per line
    blocks = open(os.path.join(tmp_path, "surface.txt"), encoding="utf-8").read().strip().split("\n\n")
    rows = [np.array([[float(v) for v in line.split()] for line in b.strip().split("\n")]) for b in blocks]
    assert len(rows) == 2 and all(r.shape == (3, 3) for r in rows)
    assert np.allclose(rows[0][:, 0], [0.25, 0.5, 0.75]) and np.allclose(rows[0][:, 1], 2 + 1 / 3)
    assert np.allclose(rows[1][:, 1], 2 + 2 / 3) and np.array_equal(rows[1][:, 2], [4.0, 5.0, 6.0])


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_dirichlet_solver_gui_defaults(driver, tmp_path, golden_ref):
    """GUI defaults (mainwindow.cpp:112-125): [1,2]^2, 30x30, eps 1e-6, precision+residual rules on."""
    run(driver, "dirichlet", 30, 30, 1, 2, 1, 2, 1e-6, 1e-6, 1e-6, 10000, 1, 1, 0, tmp_path)
    i = info(tmp_path)
    ref = golden_ref["dirichlet_n30_info"]
    assert int(i["iterations"]) == int(ref[0]) == 79 and int(i["converged"]) == 1
    assert i["stop_reason"] == bytes(golden_ref["dirichlet_n30_stop_reason"]).decode("utf-8")
    assert i["method"] == "Метод серединных градиентов"
    sol = f64(tmp_path, "solution.bin")
    assert relmax(sol, golden_ref["dirichlet_n30_solution"]) < REL
    assert np.array_equal(f64(tmp_path, "x_coords.bin"), golden_ref["dirichlet_n30_x_coords"])
    assert np.array_equal(f64(tmp_path, "y_coords.bin"), golden_ref["dirichlet_n30_y_coords"])
    assert np.max(np.abs(f64(tmp_path, "true_solution.bin") / golden_ref["dirichlet_n30_true_solution"] - 1)) < 1e-15
    bscale = np.max(np.abs(golden_ref["grid_n30_a1_rhs"]))
    assert np.max(np.abs(f64(tmp_path, "residual.bin") - golden_ref["dirichlet_n30_residual"])) <= REL * bscale
    assert np.max(np.abs(f64(tmp_path, "error.bin") - golden_ref["dirichlet_n30_error"])) <= REL * np.max(np.abs(sol))
    assert abs(float(i["residual_norm"]) - ref[2]) <= 1e-6 * ref[2]
    assert abs(float(i["error_norm"]) - ref[3]) <= 1e-9 * ref[3]
    # `precision` is finally filled (the reference leaves it uninitialised): ||x_n - x_{n-1}||_inf < eps_p
    assert 0 < float(i["precision"]) < 1e-6
    # callbacks: it 0, 1, final (79 < 100) on the solving thread; one completion callback
    cbs = np.loadtxt(os.path.join(tmp_path, "callbacks.txt"))
    assert list(cbs[:, 0]) == list(golden_ref["grid_n30_a1_msg_pr_cb"][:, 0]) == [0, 1, 79]
    assert int(i["completions"]) == 1
    # report + file formats
    report = open(os.path.join(tmp_path, "report.txt"), encoding="utf-8").read()
    assert "Выполнено итераций: 79" in report and "Сходимость: Да" in report
    assert int(i["saved"]) == 1 and int(i["saved_matrix"]) == 1 and int(i["io_roundtrip"]) == 1 and int(i["saved3d"]) == 1
    assert float(i["io_worst_rel"]) < 1e-6  # std::scientific keeps 7 significant digits
    results_txt = open(os.path.join(tmp_path, "results.txt"), encoding="utf-8").read().split("\n")
    assert results_txt[0] == "PARAMETERS" and results_txt[1] == "30 30" and "SOLUTION" in results_txt
    matrix_txt = open(os.path.join(tmp_path, "matrix.txt")).read().split("\n")
    assert matrix_txt[0] == "MATRIX_INFO" and matrix_txt[2] == "616 2964"
    assert (int(i["matrix_rows"]), int(i["matrix_cols"])) == (29, 29)


@pytest.mark.gpu
@pytest.mark.parametrize("operator", ["matrix-free (default)", "csr"])
def test_dirichlet_solver_residual_rule_only(driver, tmp_path, golden_ref, operator):
    """DirichletSolver iterates on the matrix-free operator by default (the same matrix, bit for bit) and on the assembled
    CSR arrays with B200CG_DIRICHLET_OPERATOR=csr: the reference's fixtures hold for both."""
    run(driver, "dirichlet", 128, 128, 0, 1, 0, 1, 1e-8, 1e-8, 1e-8, 10000, 0, 1, 0, tmp_path,
        env={"B200CG_DIRICHLET_OPERATOR": "csr"} if operator == "csr" else None)
    i = info(tmp_path)
    ref = golden_ref["grid_n128_a0_msg_r_info"]
    assert abs(int(i["iterations"]) - int(ref[0])) <= 1 and int(ref[0]) == 482
    assert relmax(f64(tmp_path, "solution.bin"), golden_ref["grid_n128_a0_msg_r_x"]) < REL
    cbs = np.loadtxt(os.path.join(tmp_path, "callbacks.txt"))
    assert list(cbs[:-1, 0]) == [0, 1, 100, 200, 300, 400]  # cadence of msg_solver.cpp:75,172


@pytest.mark.gpu
@pytest.mark.parametrize("n,a_tag,iters", [(6, 1, 13), (30, 1, 88), (128, 0, 352)])
def test_matrix_free_classes(driver, tmp_path, golden_ref, n, a_tag, iters):
    tag = f"mf_n{n}_a{a_tag}"
    lo, hi = (0, 1) if a_tag == 0 else (1, 2)
    golden_ref[tag + "_apply_in"].tofile(os.path.join(tmp_path, "apply_in.bin"))
    with_cb = 1 if n <= 30 else 0
    run(driver, "mf", n, lo, hi, 1e-8, 10000, with_cb, tmp_path)
    i = info(tmp_path)
    assert abs(int(i["iterations"]) - iters) <= 1 and int(i["converged"]) == 1
    assert i["message"] == "Converged successfully" and i["name"] == "Matrix-free solver"
    assert int(i["size"]) == len(golden_ref[tag + "_rhs"]) and int(i["describes_size"]) == 1
    assert relmax(f64(tmp_path, "rhs.bin"), golden_ref[tag + "_rhs"]) < 1e-14  # device exp() vs glibc
    assert np.array_equal(f64(tmp_path, "apply_out.bin"), golden_ref[tag + "_apply_out"])
    assert relmax(f64(tmp_path, "x.bin"), golden_ref[tag + "_x"]) < REL
    if with_cb:
        hist = np.loadtxt(os.path.join(tmp_path, "hist.txt")).reshape(-1, 4)
        ref = golden_ref[tag + "_hist"]
        assert len(hist) == len(ref) and list(hist[:, 0]) == list(range(len(ref)))
        assert np.all(np.abs(hist[:, 1:] - ref) <= 1e-9 * np.max(np.abs(ref), axis=0) + 1e-9 * np.abs(ref))


@pytest.mark.gpu
def test_matrix_free_solver_iteration_cap(driver, tmp_path):
    run(driver, "mf", 30, 1, 2, 1e-12, 5, 0, tmp_path)
    i = info(tmp_path)
    assert int(i["iterations"]) == 5 and int(i["converged"]) == 0
    assert i["message"] == "Failed to converge within maximum iterations"


@pytest.mark.gpu
@pytest.mark.parametrize("n,a_tag", [(6, 1), (30, 1)])
def test_grid_system_and_standalone_msg_solver(driver, tmp_path, golden_ref, n, a_tag):
    tag = f"grid_n{n}_a{a_tag}"
    lo, hi = (0, 1) if a_tag == 0 else (1, 2)
    run(driver, "grid", n, lo, hi, 1e-6, 1e-6, -1.0, 10000, tmp_path)
    i = info(tmp_path)
    assert [int(i["rows"]), int(i["nnz"])] == list(golden_ref[tag + "_shape"]) and int(i["describes_nnz"]) == 1
    assert np.array_equal(np.fromfile(os.path.join(tmp_path, "row_map.bin"), dtype=np.int32), golden_ref[tag + "_row_map"])
    assert np.array_equal(np.fromfile(os.path.join(tmp_path, "entries.bin"), dtype=np.int32), golden_ref[tag + "_entries"])
    assert np.array_equal(f64(tmp_path, "values.bin"), golden_ref[tag + "_values"])
    assert relmax(f64(tmp_path, "rhs.bin"), golden_ref[tag + "_rhs"]) < 1e-14
    assert np.array_equal(f64(tmp_path, "xs.bin"), golden_ref[tag + "_xs"])
    ref = golden_ref[f"{tag}_msg_pr_info"]
    assert int(i["iterations"]) == int(ref[0]) and int(i["stop"]) == int(ref[2])
    x = f64(tmp_path, "x.bin")
    assert relmax(x, golden_ref[f"{tag}_msg_pr_x"]) < REL
    # KokkosSparse::spmv stand-in: A x through the device equals a host CSR product in stored order
    rm, en, va = golden_ref[tag + "_row_map"], golden_ref[tag + "_entries"], golden_ref[tag + "_values"]
    ax = np.array([sum(va[k] * x[en[k]] for k in range(rm[r], rm[r + 1])) for r in range(len(rm) - 1)])
    assert np.max(np.abs(f64(tmp_path, "Ax.bin") - ax)) <= 1e-12 * np.max(np.abs(ax))
    assert (float(i["first_x"]), float(i["first_y"])) == (golden_ref[tag + "_xs"][0], golden_ref[tag + "_ys"][0])
    assert (float(i["bad_x"]), float(i["bad_y"])) == (0.0, 0.0)


@pytest.mark.gpu
def test_error_conventions(driver, tmp_path):
    run(driver, "errors", tmp_path)
    assert info(tmp_path)["caught"] == "3"


@pytest.mark.gpu
def test_request_stop_from_another_thread(driver, tmp_path):
    run(driver, "stop", 1024, tmp_path)
    i = info(tmp_path)
    assert i["stop_reason"] == "Прервано пользователем" and int(i["converged"]) == 0 and int(i["iterations"]) > 0


@pytest.mark.gpu
def test_multigrid_pcg_behind_the_solver_interface(driver, tmp_path, oracle_mod):
    """The opt-in preconditioned CG as a Solver subclass on a GridSystem (solver.hpp:17-66 is the base class kept for
    further solvers) and as a switch of MatrixFreeSolver: same solution as the reference-order plain CG, 7 iterations."""
    from oracle import mg_oracle as mg

    n = 128
    run(driver, "mgpcg", n, 0, 1, 1e-8, 1000, tmp_path)
    i = info(tmp_path)
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, 0)
    b = o.rhs()
    assert np.array_equal(f64(tmp_path, "rhs.bin"), b) or relmax(f64(tmp_path, "rhs.bin"), b) < 1e-14
    ref = mg.MgPcg(n, n).solve(mg.to_grid(f64(tmp_path, "rhs.bin"), n, n, True), eps=1e-8, max_it=1000)
    assert int(i["iterations"]) == ref["iterations"] == int(i["mf_iterations"]) and int(i["converged"]) == 1
    assert int(i["levels"]) == ref["levels"] == 6 and int(i["completions"]) == 1
    x = f64(tmp_path, "x.bin")
    assert relmax(x, mg.from_grid(ref["x"], n, n, True)) < 1e-10
    assert np.array_equal(x, f64(tmp_path, "x_mf.bin"))
    plain = o.mf_solve(b=b, eps=1e-10, max_it=20000)
    assert relmax(x, plain["x"]) < 1e-7
    assert float(i["r"]) <= 1e-8 * float(i["r0"])


@pytest.mark.gpu
def test_matrix_free_solver_batch(driver, tmp_path, golden_ref):
    """MatrixFreeSolver::solveBatch (B200 addition, b200cg_solve_batch): three scaled right-hand sides through the queue.
    The scales are powers of two, under which every operation of CG scales exactly: all three take the reference's
    iteration count for this grid and give the scaled solution; the first equals solve() on the same system."""
    n = 64
    run(driver, "mfbatch", n, 0, 1, 1e-8, 10000, tmp_path)
    i = info(tmp_path)
    ref_x = golden_ref[f"mf_n{n}_a0_x"]
    ref_it = int(golden_ref[f"mf_n{n}_a0_iters"][0])
    its = [int(t) for t in i["iterations"].split()]
    assert its == [ref_it] * 3 and int(i["single_iterations"]) == ref_it
    assert int(i["completions"]) == 4  # three from the queue, one from solve(), all converged
    assert relmax(f64(tmp_path, "x0.bin"), ref_x) < REL
    assert relmax(f64(tmp_path, "x_single.bin"), ref_x) < REL
    assert np.array_equal(f64(tmp_path, "x0.bin"), f64(tmp_path, "x_single.bin"))
    assert relmax(f64(tmp_path, "x1.bin"), -2.0 * f64(tmp_path, "x0.bin")) < 1e-12
    assert relmax(f64(tmp_path, "x2.bin"), 0.5 * f64(tmp_path, "x0.bin")) < 1e-12
