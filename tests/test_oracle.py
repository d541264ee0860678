"""CPU tests: pin the C restatement (oracle/cg_oracle.c) against
(1) the reference's own known-answer data, (2) fixtures produced by the unmodified reference classes,
(3) the compiled reference itself when oracle/_ref is present, (4) an independent scipy solve."""
import numpy as np
import pytest

DOMAINS = {0: (0.0, 1.0), 1: (1.0, 2.0)}


def mk(oracle_mod, n, a_tag, kind=0):
    a, b = DOMAINS[a_tag]
    return oracle_mod.Oracle(n, n, a, b, a, b, kind)


# ---------------------------------------------------------------- (1) reference scripts
def test_matrix_matches_check_py(oracle_mod, golden_scripts):
    """apply(e_j) reproduces the 16x16 matrix of check.py:4-19 exactly; so does the CSR assembly."""
    o = mk(oracle_mod, 6, 1)
    assert o.N == 16
    A = golden_scripts["check_matrix"]
    cols = np.stack([o.apply(np.eye(16)[j]) for j in range(16)], axis=1)
    assert np.array_equal(cols, A)
    row_map, entries, values = o.csr()
    dense = np.zeros((16, 16))
    for i in range(16):
        for k in range(row_map[i], row_map[i + 1]):
            dense[i, entries[k]] += values[k]
    assert np.array_equal(dense, A)
    assert row_map[-1] == 60


def test_rhs_matches_check_debug_py(oracle_mod, golden_scripts):
    """GridSystem(6,6,1,2,1,2) rhs equals check_debug.py:36 to its 8 printed decimals."""
    b = mk(oracle_mod, 6, 1).rhs()
    assert np.max(np.abs(b - golden_scripts["check_debug_rhs"])) < 5.1e-9


def test_two_cg_steps_match_py_debug(oracle_mod, golden_scripts):
    """x2 after two iterations equals py_debug.txt:14 to 3e-11 (the script's RHS is rounded to 8 decimals);
    with the script's own rounded RHS the agreement is at rounding level. Sign conventions: SURVEY 8c."""
    o = mk(oracle_mod, 6, 1)
    s = o.mf_solve(eps=1e-9, max_it=2)
    assert s["iterations"] == 2
    assert np.max(np.abs(s["x"] - golden_scripts["py_x2"])) < 1e-10
    brounded = golden_scripts["check_debug_rhs"]
    s2 = o.mf_solve(b=brounded, eps=1e-9, max_it=2, snapshots=True)
    assert np.max(np.abs(s2["x"] - golden_scripts["py_x2"])) < 1e-13
    assert np.max(np.abs(-s2["r"] - golden_scripts["py_r2"])) < 2e-11  # r_cpp = -r_py
    s1 = o.mf_solve(b=brounded, eps=1e-9, max_it=1, snapshots=True)
    assert np.max(np.abs(s1["x"] - golden_scripts["py_x1"])) < 1e-13
    assert np.max(np.abs(s1["p"] - golden_scripts["py_h1"])) < 2e-11   # h1 = +z1
    assert np.max(np.abs(o.apply(s1["p"]) - golden_scripts["py_A_h1"])) < 1e-8
    assert np.max(np.abs(o.apply(-brounded) - golden_scripts["py_A_h0"])) < 1e-8
    # MSG flavour, 2 iterations (solver/main.cpp:601-602): r2 max-norm = |py_debug r2|_inf
    m = o.msg_solve(b=brounded, eps_p=1e-9, eps_r=1e-9, max_it=2)
    assert m["iterations"] == 2 and m["stop_reason"] == "ITERATIONS"
    assert abs(m["r_max"] - np.max(np.abs(golden_scripts["py_r2"]))) < 2e-11
    assert np.max(np.abs(m["x"] - golden_scripts["py_x2"])) < 1e-13


# ---------------------------------------------------------------- (2) fixtures from the unmodified reference
@pytest.mark.parametrize("n,a_tag,iters", [(6, 1, 13), (30, 1, 88), (64, 0, 178), (128, 0, 352), (128, 1, 362)])
def test_matrix_free_path_bit_exact(oracle_mod, golden_ref, n, a_tag, iters):
    o = mk(oracle_mod, n, a_tag)
    tag = f"mf_n{n}_a{a_tag}"
    assert np.array_equal(o.rhs(), golden_ref[tag + "_rhs"])
    assert np.array_equal(o.true_solution(), golden_ref[tag + "_true"])
    assert np.array_equal(o.apply(golden_ref[tag + "_apply_in"]), golden_ref[tag + "_apply_out"])
    s = o.mf_solve(eps=1e-8, max_it=10000, with_hist=(n <= 30))
    assert s["iterations"] == iters == int(golden_ref[tag + "_iters"][0])
    assert s["converged"]
    assert np.array_equal(s["x"], golden_ref[tag + "_x"])  # same operations in the same order
    assert np.array_equal(o.mf_solve(eps=1e-8, max_it=2)["x"], golden_ref[tag + "_x2"])
    if n <= 30:
        assert np.array_equal(s["hist"], golden_ref[tag + "_hist"])


@pytest.mark.parametrize("n,a_tag", [(6, 1), (30, 1), (128, 0)])
def test_assembled_path_bit_exact(oracle_mod, golden_ref, n, a_tag):
    o = mk(oracle_mod, n, a_tag)
    tag = f"grid_n{n}_a{a_tag}"
    row_map, entries, values = o.csr()
    assert [o.N, len(values)] == list(golden_ref[tag + "_shape"])
    if n <= 30:
        assert np.array_equal(row_map, golden_ref[tag + "_row_map"])
        assert np.array_equal(entries, golden_ref[tag + "_entries"])
        assert np.array_equal(values, golden_ref[tag + "_values"])
    assert np.array_equal(o.rhs(), golden_ref[tag + "_rhs"])
    xs, ys = o.node_coords()
    assert np.array_equal(xs, golden_ref[tag + "_xs"]) and np.array_equal(ys, golden_ref[tag + "_ys"])
    u = o.true_solution()
    assert np.array_equal(u, golden_ref[tag + "_true"])
    eps = 1e-6 if n <= 30 else 1e-8
    for cname, kw in {"pr": dict(eps_p=eps, eps_r=eps), "r": dict(eps_p=-1.0, eps_r=eps)}.items():
        s = o.msg_solve(u=u, eps_e=-1.0, max_it=10000, cb_cap=256, **kw)
        info = golden_ref[f"{tag}_msg_{cname}_info"]
        assert s["iterations"] == int(info[0]) and s["converged"] == bool(info[1])
        assert oracle_mod.STOP_NAMES.index(s["stop_reason"]) == int(info[2])
        assert (s["r_max"], s["dx_max"], s["err_max"]) == tuple(info[3:6])
        assert np.array_equal(s["x"], golden_ref[f"{tag}_msg_{cname}_x"])
        assert np.array_equal(s["callbacks"], golden_ref[f"{tag}_msg_{cname}_cb"])


MSG_CASES = ["n30_a1_exact", "n30_a1_exact_first", "n30_a1_nou", "n64_a0_pr", "n64_a0_r", "n64_a0_cap"]


@pytest.mark.parametrize("tag", MSG_CASES)
def test_msg_rules_bit_exact_on_the_extra_fixtures(oracle_mod, golden_msg, tag):
    """MSGSolver branches beyond the first fixture file, run by the unmodified reference: the exact-error rule firing
    (alone, and ahead of the other two), a solve without a true solution (error stays DBL_MAX, its rule is skipped), the
    64 x 64 grid on [0,1]^2, the iteration cap (ITERATIONS, not converged). Bit for bit, callbacks included."""
    n, a, b, eps_p, eps_r, eps_e, max_it, with_true = golden_msg[f"msg_{tag}_params"]
    o = oracle_mod.Oracle(int(n), int(n), a, b, a, b)
    u = o.true_solution() if with_true else None
    s = o.msg_solve(u=u, eps_p=eps_p, eps_r=eps_r, eps_e=eps_e, max_it=int(max_it), cb_cap=256)
    info = golden_msg[f"msg_{tag}_info"]
    assert s["iterations"] == int(info[0]) and s["converged"] == bool(info[1])
    assert oracle_mod.STOP_NAMES.index(s["stop_reason"]) == int(info[2])
    assert (s["r_max"], s["dx_max"], s["err_max"]) == tuple(info[3:6])
    assert np.array_equal(s["x"], golden_msg[f"msg_{tag}_x"])
    assert np.array_equal(s["callbacks"], golden_msg[f"msg_{tag}_cb"])


def test_extra_fixture_stop_reasons(golden_msg):
    got = {tag: (int(golden_msg[f"msg_{tag}_info"][0]), int(golden_msg[f"msg_{tag}_info"][2])) for tag in MSG_CASES}
    assert got == {"n30_a1_exact": (50, 3), "n30_a1_exact_first": (48, 3), "n30_a1_nou": (88, 1), "n64_a0_pr": (180, 1),
                   "n64_a0_r": (238, 2), "n64_a0_cap": (150, 0)}


def test_pinned_iteration_counts(golden_ref):
    """SURVEY 8c pins, re-derived from the unmodified reference when the fixture was made."""
    assert int(golden_ref["grid_n30_a1_msg_pr_info"][0]) == 79
    assert int(golden_ref["grid_n30_a1_msg_r_info"][0]) == 102
    assert int(golden_ref["grid_n128_a0_msg_r_info"][0]) == 482
    assert int(golden_ref["grid_n128_a0_msg_pr_info"][0]) == 355
    assert int(golden_ref["dirichlet_n30_info"][0]) == 79


def test_facade_postprocessing(oracle_mod, golden_ref):
    """DirichletSolver::solve results (dirichlet_solver.cpp:101-123,147-180): residual = A x - b, error = x - u."""
    o = mk(oracle_mod, 30, 1)
    u = o.true_solution()
    s = o.msg_solve(u=u, eps_p=1e-6, eps_r=1e-6, eps_e=-1.0, max_it=10000)
    assert np.array_equal(s["x"], golden_ref["dirichlet_n30_solution"])
    assert np.array_equal(o.spmv(o.csr(), s["x"]) - o.rhs(), golden_ref["dirichlet_n30_residual"])
    assert np.array_equal(s["x"] - u, golden_ref["dirichlet_n30_error"])
    assert s["r_max"] == golden_ref["dirichlet_n30_info"][2] and s["err_max"] == golden_ref["dirichlet_n30_info"][3]


# ---------------------------------------------------------------- (3) live reference, when it was built here
def test_against_live_reference(oracle_mod):
    if not oracle_mod.Reference.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    for n, (a, b) in [(8, (0.0, 1.0)), (10, (1.0, 2.0)), (48, (0.0, 1.0))]:
        ref = oracle_mod.Reference.MatrixFree(n, n, a, b, a, b)
        o = oracle_mod.Oracle(n, n, a, b, a, b)
        assert ref.N == o.N
        assert np.array_equal(ref.rhs(), o.rhs())
        x = np.random.default_rng(n).standard_normal(o.N)
        assert np.array_equal(ref.apply(x), o.apply(x))
        rs, os_ = ref.solve(eps=1e-10, max_it=500), o.mf_solve(eps=1e-10, max_it=500)
        assert rs["iterations"] == os_["iterations"] and np.array_equal(rs["x"], os_["x"])


def test_port_equals_the_compiled_reference_at_config2_size(oracle_mod):
    """BASELINE.json configs[1] (4096^2, 12.6 M unknowns), 5 iterations - the case tests/test_gpu_parity.py holds the GPU
    to: the C restatement reproduces the unmodified reference's rhs and its fifth iterate bit for bit there too (same
    operations, same sequential sums), so a comparison with the restatement IS a comparison with the reference."""
    if not oracle_mod.Reference.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    n = 4096
    ref = oracle_mod.Reference.MatrixFree(n, n, 0.0, 1.0, 0.0, 1.0)
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0)
    b = o.rhs()
    assert ref.N == o.N == 12574721 and np.array_equal(ref.rhs(), b)
    rs, ps = ref.solve(eps=1e-8, max_it=5), o.mf_solve(b=b, eps=1e-8, max_it=5)
    assert rs["iterations"] == ps["iterations"] == 5 and np.array_equal(rs["x"], ps["x"])


@pytest.mark.parametrize("m,n,kind", [(4, 4, 0), (6, 6, 0), (30, 30, 0), (128, 128, 0), (2, 2, 1), (2, 5, 1), (7, 9, 1),
                                      (37, 1200, 1), (4, 4, 3), (5, 4, 3), (4, 5, 3), (9, 12, 3), (12, 9, 3), (33, 20, 3),
                                      (481, 333, 3)])
def test_rowwise_apply_is_the_nodewise_apply(oracle_mod, m, n, kind):
    """cgo_apply walks the operator row by row for speed; cgo_apply_nodewise is the reference's node-by-node form
    (matrix_free_system.cpp:203-340: boundary predicates and index formulas per neighbour). Same products in the same order:
    bit-identical on every domain kind, square or not, even or odd."""
    o = oracle_mod.Oracle(m, n, 0.0, 1.0, -1.0, 0.5, kind) if kind else oracle_mod.Oracle(m, n, 0.0, 1.0, 0.0, 1.0)
    rng = np.random.default_rng(m * 1000 + n)
    for _ in range(2):
        v = rng.standard_normal(o.N)
        assert np.array_equal(o.apply(v), o.apply_nodewise(v))


def test_reference_rejects_what_oracle_rejects(oracle_mod):
    """The reference numbering is only self-consistent for even n == m; the oracle refuses the rest."""
    for n, m in [(8, 6), (7, 7), (2, 2)]:
        with pytest.raises(ValueError):
            oracle_mod.Oracle(m, n, 0, 1, 0, 1)


# ---------------------------------------------------------------- (4) independent checks
@pytest.mark.parametrize("kind", [0, 1])
def test_against_scipy_direct_solve(oracle_mod, kind):
    sp = pytest.importorskip("scipy.sparse")
    from scipy.sparse.linalg import spsolve

    n = 32 if kind == 0 else 33
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, kind)
    # independent matrix from apply() on unit vectors is O(N^2); use the geometry instead
    rows, cols, vals = [], [], []
    for i in range(o.N):
        x, y = o.node(i)
        rows.append(i); cols.append(i); vals.append(o.g.A)
        for (dx, dy, c) in ((-1, 0, o.g.xk), (1, 0, o.g.xk), (0, 1, o.g.yk), (0, -1, o.g.yk)):
            j = o.index(x + dx, y + dy)
            if j >= 0:
                rows.append(i); cols.append(j); vals.append(c)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(o.N, o.N))
    xin = np.random.default_rng(3).standard_normal(o.N)
    assert np.max(np.abs(A @ xin - o.apply(xin))) < 1e-9 * np.max(np.abs(A @ xin))
    b = o.rhs()
    xd = spsolve(A.tocsc(), b)
    s = o.mf_solve(eps=1e-12, max_it=5000)
    assert s["converged"]
    assert np.max(np.abs(s["x"] - xd)) < 1e-9 * np.max(np.abs(xd))
    # O(h^2) agreement with the analytic solution u = exp(x^2 - y^2)
    assert np.max(np.abs(s["x"] - o.true_solution())) < 5e-4


def test_dof_counts(oracle_mod):
    """(n/2-1)(m/2) + (n-1)(m/2-1)."""
    for n, N in [(6, 16), (30, 616), (128, 12033)]:
        assert oracle_mod.Oracle(n, n).N == N
    assert oracle_mod.Oracle(128, 128, kind=1).N == 127 * 127


# ---------------------------------------------------------------- repaired general L-shape (no reference counterpart)
def test_lshape_any_equals_reference_geometry_where_the_reference_is_valid(oracle_mod):
    for n in (6, 30, 64):
        ref = oracle_mod.Oracle(n, n, 1.0, 2.0, 1.0, 2.0, oracle_mod.LSHAPE)
        gen = oracle_mod.Oracle(n, n, 1.0, 2.0, 1.0, 2.0, oracle_mod.LSHAPE_ANY)
        assert ref.N == gen.N and np.array_equal(ref.rhs(), gen.rhs())
        x = np.random.default_rng(n).standard_normal(ref.N)
        assert np.array_equal(ref.apply(x), gen.apply(x))
        for a, b in zip(ref.csr(), gen.csr()):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("n,m", [(7, 7), (9, 12), (33, 20)])
def test_lshape_any_is_a_well_formed_system(oracle_mod, n, m):
    """Odd and non-square grids, where the reference builds a malformed system or crashes: the repaired
    numbering is a bijection, the operator is symmetric, CG converges to the analytic solution at O(h^2)."""
    o = oracle_mod.Oracle(m, n, 0.0, 1.0, 0.0, 1.0, oracle_mod.LSHAPE_ANY)
    seen = set()
    for y in range(m + 1):
        for x in range(n + 1):
            i = o.index(x, y)
            if i >= 0:
                assert o.node(i) == (x, y)
                seen.add(i)
    assert seen == set(range(o.N))
    row_map, entries, values = o.csr()
    assert np.all(np.diff(row_map) <= 5) and entries.min() >= 0 and entries.max() < o.N
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(o.N), rng.standard_normal(o.N)
    assert abs(np.dot(x, o.apply(y)) - np.dot(y, o.apply(x))) <= 1e-9 * abs(np.dot(x, o.apply(y))) + 1e-9
    s = o.mf_solve(eps=1e-12, max_it=5000)
    assert s["converged"] and np.max(np.abs(s["x"] - o.true_solution())) < 5e-3
