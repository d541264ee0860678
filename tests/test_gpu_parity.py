"""GPU parity tests (run on the B200 box: pytest -m gpu). Every call goes through the C ABI (libb200cg.so via
the ctypes mirror); the CPU oracle and the committed golden fixtures are the checkers.

Bars (BASELINE.json north_star): iteration count within +-1 of the reference, final residual and solution
max-abs difference within 1e-10 relative in fp64. Element-wise operations (apply, CSR assembly, axpys) are
expected to be numerically identical; only dot-product summation order differs.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-10  # the north-star tolerance
# 16384^2, 2 iterations: GPU vs the oracle in the REFERENCE's summation order. The reference's own sequential-sum error is
# 6.7e-13 there (oracle reference order vs oracle long-double sums, measured on the CPU; it grows with the iteration
# count, not with N: 2e-10 after 5 iterations at 4096^2), so the north-star bar itself holds
HEADLINE_REF_ORDER_BAR = 1e-10
DOMAINS = {0: (0.0, 1.0), 1: (1.0, 2.0)}


@pytest.fixture(scope="module")
def capi():
    from iterative_solvers_b200 import capi as c

    c.lib()  # raises if the CUDA library was not built: no fallback
    assert c.device_count() >= 1, "these tests need a CUDA device"
    return c


def plan_for(capi, n, a_tag=0, domain=0, **kw):
    a, b = DOMAINS[a_tag]
    return capi.Plan(n, n, a, b, a, b, domain=domain, **kw)


def oracle_for(oracle_mod, n, a_tag=0, kind=0, m=None):
    a, b = DOMAINS[a_tag]
    return oracle_mod.Oracle(m or n, n, a, b, a, b, kind)


def relmax(x, ref):
    return np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300)


def ulp_diff(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.spacing(np.abs(b)), 1e-300))


# ---------------------------------------------------------------- K0: rhs, true solution, coordinates
@pytest.mark.parametrize("n,a_tag", [(6, 1), (30, 1), (128, 0), (128, 1)])
def test_setup_vectors(capi, golden_ref, n, a_tag):
    with plan_for(capi, n, a_tag) as p:
        tag = f"mf_n{n}_a{a_tag}"
        assert p.N == len(golden_ref[tag + "_rhs"])
        p.build_rhs()
        b = p.get_rhs()
        ref = golden_ref[tag + "_rhs"]
        # device exp() vs glibc exp(): a few ulp of the largest term; boundary terms are O(1/h^2) * u
        assert np.max(np.abs(b - ref)) <= 8 * np.spacing(np.max(np.abs(ref)))
        assert relmax(b, ref) < 1e-14
        u = p.true_solution()
        assert ulp_diff(u, golden_ref[tag + "_true"]) <= 4
        if (n, a_tag) in ((6, 1), (30, 1), (128, 0)):
            xs, ys = p.coords()
            assert np.array_equal(xs, golden_ref[f"grid_n{n}_a{a_tag}_xs"])
            assert np.array_equal(ys, golden_ref[f"grid_n{n}_a{a_tag}_ys"])


def test_set_get_rhs_roundtrip(capi):
    with plan_for(capi, 30) as p:
        v = np.random.default_rng(1).standard_normal(p.N)
        p.set_rhs(v)
        assert np.array_equal(p.get_rhs(), v)


# ---------------------------------------------------------------- apply
@pytest.mark.parametrize("n,a_tag", [(6, 1), (30, 1), (64, 0), (128, 0), (128, 1)])
def test_apply_matches_reference_fixture(capi, golden_ref, n, a_tag):
    with plan_for(capi, n, a_tag) as p:
        tag = f"mf_n{n}_a{a_tag}"
        y = p.apply(golden_ref[tag + "_apply_in"])
        assert np.array_equal(y, golden_ref[tag + "_apply_out"])  # same operations in the same order


def test_apply_reproduces_check_py_matrix(capi, golden_scripts):
    with plan_for(capi, 6, 1) as p:
        cols = np.stack([p.apply(np.eye(16)[j]) for j in range(16)], axis=1)
        assert np.array_equal(cols, golden_scripts["check_matrix"])


@pytest.mark.parametrize("n,domain,tile_rows", [(600, 0, 0), (1030, 0, 7), (1030, 0, 1), (512, 0, 64), (1009, 1, 5),
                                                (505, 1, 0), (4, 0, 0), (2, 1, 0), (3, 1, 0)])
def test_apply_vs_oracle_across_strips_and_tiles(capi, oracle_mod, n, domain, tile_rows):
    """Grids wider than one 504-column strip, ragged tile heights, both domain kinds, smallest grids."""
    o = oracle_for(oracle_mod, n, 0, domain)
    with plan_for(capi, n, 0, domain, tile_rows=tile_rows) as p:
        assert p.N == o.N
        x = np.random.default_rng(n).standard_normal(o.N)
        assert np.array_equal(p.apply(x), o.apply(x))


def test_apply_rect_nonsquare(capi, oracle_mod):
    o = oracle_mod.Oracle(37, 1200, 0.0, 2.0, -1.0, 0.5, 1)  # m=37, n=1200
    with capi.Plan(37, 1200, 0.0, 2.0, -1.0, 0.5, domain=1) as p:
        x = np.random.default_rng(5).standard_normal(o.N)
        assert np.array_equal(p.apply(x), o.apply(x))


# ---------------------------------------------------------------- MatrixFreeSolver path
def test_two_iterations_match_py_debug(capi, golden_scripts, golden_ref):
    """The reference's own known answer: x2 of py_debug.txt:14 (script RHS rounded to 8 decimals)."""
    with plan_for(capi, 6, 1) as p:
        x, info = p.solve(b=golden_scripts["check_debug_rhs"], eps_rel=1e-9, max_it=2)
        assert info["iterations"] == 2 and not info["converged"]
        assert np.max(np.abs(x - golden_scripts["py_x2"])) < 1e-13
        x, info = p.solve(b=golden_ref["mf_n6_a1_rhs"], eps_rel=1e-8, max_it=2)
        assert relmax(x, golden_ref["mf_n6_a1_x2"]) < 1e-14


@pytest.mark.parametrize("path", [0, 1], ids=["auto(cluster)", "graph"])
@pytest.mark.parametrize("n,a_tag,iters", [(6, 1, 13), (30, 1, 88), (64, 0, 178), (128, 0, 352), (128, 1, 362)])
def test_matrix_free_solve_parity(capi, golden_ref, n, a_tag, iters, path):
    """Both execution paths: grids this small default to the single cluster-resident kernel (small_grid_path 0);
    small_grid_path 1 forces the CUDA-graph loop of sweep kernels that large grids use."""
    tag = f"mf_n{n}_a{a_tag}"
    with plan_for(capi, n, a_tag) as p:
        x, info = p.solve(b=golden_ref[tag + "_rhs"], eps_rel=1e-8, max_it=10000, small_grid_path=path)
        assert info["cluster_path"] == (1 if path == 0 else 0)
        assert abs(info["iterations"] - iters) <= 1
        assert info["converged"]
        assert relmax(x, golden_ref[tag + "_x"]) < REL
        assert info["r_l2"] <= 1e-8 * info["r0_l2"]
        # config 1 pins
        if (n, a_tag) == (128, 0):
            assert info["iterations"] == 352
            assert abs(info["r0_l2"] - 5.187466787388469e5) < 1e-6
        # device-built rhs instead of the host one: still inside the bar
        p.build_rhs()
        x2, info2 = p.solve(rhs_on_device=True, eps_rel=1e-8, max_it=10000, small_grid_path=path)
        assert abs(info2["iterations"] - iters) <= 1
        assert relmax(x2, golden_ref[tag + "_x"]) < REL


def test_matrix_free_solver_callback_history(capi, golden_ref):
    """Registered callback: (it, ||dx||_2, recomputed ||b-Ax||_2, ||x-u||_2) every iteration
    (matrix_free_system.cpp:444-468)."""
    for n, a_tag in [(6, 1), (30, 1)]:
        tag = f"mf_n{n}_a{a_tag}"
        hist = golden_ref[tag + "_hist"]
        got = []
        with plan_for(capi, n, a_tag) as p:
            x, info = p.solve(b=golden_ref[tag + "_rhs"], u=golden_ref[tag + "_true"], eps_rel=1e-8,
                              max_it=10000, callback=lambda it, pr, rs, er: got.append((it, pr, rs, er)))
        assert info["iterations"] == len(hist) == len(got)
        got = np.array(got)
        assert np.array_equal(got[:, 0], np.arange(len(hist)))
        scale = np.max(np.abs(hist), axis=0)
        assert np.all(np.abs(got[:, 1:] - hist) <= 1e-9 * scale + 1e-9 * np.abs(hist))
        assert relmax(x, golden_ref[tag + "_x"]) < REL


def test_small_grid_path_limits(capi):
    with plan_for(capi, 1024) as p:  # too large for one cluster's shared memory
        p.build_rhs()
        x, info = p.solve(rhs_on_device=True, eps_rel=1e-8, max_it=4)
        assert info["cluster_path"] == 0 and info["iterations"] == 4
        with pytest.raises(capi.B200CGError) as e:
            p.solve(rhs_on_device=True, eps_rel=1e-8, max_it=4, small_grid_path=2)
        assert e.value.status == 6
    for n, domain in [(250, 0), (301, 1), (4, 0), (2, 1)]:  # largest sizes that still fit, smallest grids
        with plan_for(capi, n, 0, domain) as p:
            p.build_rhs()
            xa, ia = p.solve(rhs_on_device=True, eps_rel=1e-9, max_it=5000, small_grid_path=0)
            xb, ib = p.solve(rhs_on_device=True, eps_rel=1e-9, max_it=5000, small_grid_path=1)
            assert ia["cluster_path"] == 1 and ib["cluster_path"] == 0
            assert abs(ia["iterations"] - ib["iterations"]) <= 1
            assert relmax(xa, xb) < REL


@pytest.mark.parametrize("path", [0, 1], ids=["auto(cluster)", "graph"])
def test_solve_edge_cases(capi, path):
    with plan_for(capi, 30) as p:
        import functools
        p.solve = functools.partial(p.solve, small_grid_path=path)
        zero = np.zeros(p.N)
        x, info = p.solve(b=zero, eps_rel=1e-8, max_it=100)  # r0 = 0: loop never entered, "converged"
        assert info["iterations"] == 0 and info["converged"] and np.all(x == 0)
        p.build_rhs()
        x, info = p.solve(rhs_on_device=True, eps_rel=1e-8, max_it=0)
        assert info["iterations"] == 0 and not info["converged"] and np.all(x == 0)
        x, info = p.solve(rhs_on_device=True, eps_rel=2.0, max_it=50)  # eps >= 1: satisfied at once
        assert info["iterations"] == 0 and info["converged"]
        x, info = p.solve(rhs_on_device=True, eps_rel=1e-8, max_it=7)  # odd cap inside one graph launch
        assert info["iterations"] == 7 and not info["converged"] and info["stop_reason"] == "ITERATIONS"
        x, info = p.solve(rhs_on_device=True, eps_rel=1e-8, max_it=7, iters_per_graph=2)
        assert info["iterations"] == 7


def test_tile_height_and_graph_length_do_not_change_the_answer(capi, oracle_mod):
    o = oracle_for(oracle_mod, 64)
    ref = o.mf_solve(eps=1e-9, max_it=10000)
    for tile_rows, k in [(1, 2), (5, 6), (64, 100), (0, 0)]:
        with plan_for(capi, 64, tile_rows=tile_rows) as p:
            x, info = p.solve(b=o.rhs(), eps_rel=1e-9, max_it=10000, iters_per_graph=k)
            assert abs(info["iterations"] - ref["iterations"]) <= 1
            assert relmax(x, ref["x"]) < REL


def test_interrupt_flag(capi):
    with plan_for(capi, 128) as p:
        p.build_rhs()
        flag = ctypes.c_int(1)
        x, info = p.solve(rhs_on_device=True, eps_rel=1e-30, max_it=100000, iters_per_graph=10, stop_flag=flag,
                          small_grid_path=1)
        assert info["stop_reason"] == "INTERRUPTED" and not info["converged"]
        assert 0 < info["iterations"] <= 10
        # inside a long graph launch the loop kernels themselves see the flag: every 16th iteration (both iteration schemes)
        for ss in (1, 2):
            x, info = p.solve(rhs_on_device=True, eps_rel=1e-30, max_it=100000, iters_per_graph=100, stop_flag=flag,
                              small_grid_path=1, single_sweep=ss)
            assert info["stop_reason"] == "INTERRUPTED" and not info["converged"] and info["iterations"] == 16
        x, info = p.solve(rhs_on_device=True, rule=capi.RULE_MAXNORM, eps_r=1e-30, max_it=100000, iters_per_graph=100,
                          stop_flag=flag, small_grid_path=1)
        assert info["stop_reason"] == "INTERRUPTED" and info["iterations"] == 16
        p.assemble_csr()
        x, info = p.solve(rhs_on_device=True, op=capi.OP_CSR, rule=capi.RULE_MAXNORM, eps_r=1e-30, max_it=100000,
                          iters_per_graph=100, stop_flag=flag)
        assert info["stop_reason"] == "INTERRUPTED" and info["iterations"] == 16
        # cluster-resident kernel: the mapped flag is polled every 128 iterations
        x, info = p.solve(rhs_on_device=True, eps_rel=1e-30, max_it=100000, stop_flag=flag, small_grid_path=2)
        assert info["stop_reason"] == "INTERRUPTED" and not info["converged"] and info["cluster_path"] == 1
        assert 0 < info["iterations"] <= 128


def test_dense_callback_cadence_does_not_lap_the_ring(capi):
    """callback_every = 1 on a long small-grid solve appends more records than the device ring holds (1024): such a solve
    must stay on the graph path, whose launches are sized to the ring, and deliver every record exactly once."""
    n = 128
    with plan_for(capi, n) as p:
        p.build_rhs()
        runs = {}
        for every in (1, 100):
            got = []
            x, info = p.solve(rhs_on_device=True, rule=capi.RULE_MAXNORM, eps_p=-1.0, eps_r=-1.0, max_it=1200,
                              callback_every=every, callback=lambda it, pr, rs, er: got.append((it, pr, rs, er)))
            runs[every] = (x, info, np.array(got))
        x, info = p.solve(rhs_on_device=True, rule=capi.RULE_MAXNORM, eps_p=-1.0, eps_r=-1.0, max_it=1200,
                          callback_every=100, small_grid_path=1, callback=lambda it, pr, rs, er: got.append((it, pr, rs, er)))
        (x1, i1, g1), (x100, i100, g100) = runs[1], runs[100]
        assert i1["iterations"] == i100["iterations"] == 1200  # no rule armed: both run to the cap
        assert i1["cluster_path"] == 0 and i100["cluster_path"] == 1  # 1203 records do not fit one launch's ring, 15 do
        assert np.array_equal(g1[:, 0], np.concatenate([np.arange(0, 1201), [1200]]))  # it 0 .. 1200, then the final call
        # the same kernels (graph path) with the sparse cadence: bit-identical iterates, the records are a subset
        assert np.array_equal(x1, x) and info["cluster_path"] == 0
        sel = np.isin(g1[:-1, 0], g100[:-1, 0])
        gs = np.array(got[-len(g100):])
        assert np.array_equal(g1[:-1][sel], gs[:-1]) and np.array_equal(g1[-1], gs[-1])
        # the cluster-resident kernel sums in another order: same records to rounding while the residual is not yet noise
        early = g100[:-1, 0] <= 300
        assert np.allclose(g1[:-1][sel][early][:, 2], g100[:-1][early][:, 2], rtol=1e-6, atol=0)


def test_csr_reupload_invalidates_cached_graphs(capi, oracle_mod):
    """b200cg_set_csr / _assemble_csr free and reallocate the matrix: graphs captured with the old pointers must go."""
    n = 30
    o = oracle_for(oracle_mod, n)
    b = o.rhs()
    with plan_for(capi, n) as p:
        nnz = p.assemble_csr()
        kw = dict(b=b, op=capi.OP_CSR, rule=capi.RULE_MAXNORM, eps_p=-1.0, eps_r=1e-9, max_it=5000)
        x1, i1 = p.solve(**kw)
        row_map, entries, values = p.get_csr(nnz)
        junk = [np.zeros(1 << 20) for _ in range(4)]  # churn the allocator a little
        p.set_csr(row_map, entries, 2.0 * values)     # A -> 2 A: the solution halves
        del junk
        x2, i2 = p.solve(**kw)
        assert relmax(2.0 * x2, x1) < 1e-8
        p.assemble_csr()
        x3, i3 = p.solve(**kw)
        assert np.array_equal(x3, x1) and i3["iterations"] == i1["iterations"]


# ---------------------------------------------------------------- MSGSolver rules (max-norm), both operators
@pytest.mark.parametrize("op", [0, 1, 2], ids=["matrix-free(cluster)", "csr", "matrix-free(graph)"])
@pytest.mark.parametrize("n,a_tag", [(6, 1), (30, 1), (128, 0)])
def test_maxnorm_rules_parity(capi, golden_ref, n, a_tag, op):
    path = 1 if op == 2 else 0
    op = 0 if op == 2 else op
    tag = f"grid_n{n}_a{a_tag}"
    eps = 1e-6 if n <= 30 else 1e-8
    with plan_for(capi, n, a_tag) as p:
        if op == 1:
            p.assemble_csr()
        for cname, kw in {"pr": dict(eps_p=eps, eps_r=eps), "r": dict(eps_p=-1.0, eps_r=eps)}.items():
            info_ref = golden_ref[f"{tag}_msg_{cname}_info"]
            cb_ref = golden_ref[f"{tag}_msg_{cname}_cb"]
            got = []
            x, info = p.solve(b=golden_ref[tag + "_rhs"], u=golden_ref[tag + "_true"], op=op,
                              rule=capi.RULE_MAXNORM, max_it=10000, small_grid_path=path,
                              callback=lambda it, pr, rs, er: got.append((it, pr, rs, er)), **kw)
            assert abs(info["iterations"] - int(info_ref[0])) <= 1
            assert info["converged"] == bool(info_ref[1])
            assert info["stop_reason"] == capi.STOP_NAMES[int(info_ref[2])]
            assert relmax(x, golden_ref[f"{tag}_msg_{cname}_x"]) < REL
            if info["iterations"] == int(info_ref[0]):
                # the recurrence residual has dropped ~14 orders of magnitude: it agrees to 1e-10 of the problem
                # scale |b|_inf (the north-star bar) and to a few digits of its own tiny value
                bscale = np.max(np.abs(golden_ref[tag + "_rhs"]))
                assert abs(info["r_max"] - info_ref[3]) <= REL * bscale
                assert abs(info["r_max"] - info_ref[3]) <= 1e-3 * abs(info_ref[3])
                assert abs(info["err_max"] - info_ref[5]) <= 1e-9 * abs(info_ref[5])
                # callback cadence: it 0, 1, every 100, final (msg_solver.cpp:75,172,193)
                got = np.array(got)
                assert np.array_equal(got[:, 0], cb_ref[:, 0])
                assert got[0, 1] == cb_ref[0, 1] == np.finfo(np.float64).max
                assert np.allclose(got[1:, 1:], cb_ref[1:, 1:], rtol=1e-3, atol=0)
                assert np.all(np.abs(got[1:, 2] - cb_ref[1:, 2]) <= REL * bscale)


def test_maxnorm_without_true_solution(capi, golden_ref, golden_msg):
    """Against a run of the unmodified reference (tests/golden/reference_outputs_msg.npz): without a true solution the
    error stays DBL_MAX and its rule is skipped (msg_solver.cpp:64-72,158)."""
    info_ref, x_ref = golden_msg["msg_n30_a1_nou_info"], golden_msg["msg_n30_a1_nou_x"]
    with plan_for(capi, 30, 1) as p:
        x, info = p.solve(b=golden_ref["grid_n30_a1_rhs"], rule=capi.RULE_MAXNORM, eps_p=1e-7, eps_r=-1.0, eps_e=1e-3,
                          max_it=10000)
        assert info["iterations"] == int(info_ref[0]) and info["stop_reason"] == capi.STOP_NAMES[int(info_ref[2])]
        assert info["err_max"] == np.finfo(np.float64).max == info_ref[5]
        assert relmax(x, x_ref) < REL


def test_exact_error_rule(capi, golden_ref, golden_msg):
    """The exact-error rule firing, against a run of the unmodified reference (msg_solver.cpp:158-162)."""
    info_ref, x_ref = golden_msg["msg_n30_a1_exact_info"], golden_msg["msg_n30_a1_exact_x"]
    assert capi.STOP_NAMES[int(info_ref[2])] == "EXACT_ERROR"
    b, u = golden_ref["grid_n30_a1_rhs"], golden_ref["grid_n30_a1_true"]
    for op in (0, 1):
        with plan_for(capi, 30, 1) as p:
            if op:
                p.assemble_csr()
            x, info = p.solve(b=b, u=u, op=op, rule=capi.RULE_MAXNORM, eps_e=5e-3, max_it=10000)
            assert info["stop_reason"] == "EXACT_ERROR" and abs(info["iterations"] - int(info_ref[0])) <= 1
            assert relmax(x, x_ref) < 1e-9


# ---------------------------------------------------------------- CSR path
def test_general_csr_with_long_and_ragged_rows(capi, oracle_mod):
    """b200cg_set_csr takes any matrix (MSGSolver receives just a matrix and a rhs): long, short and empty rows - all
    bit-equal to the stored-order row sums of the serial restatement."""
    rng = np.random.default_rng(3)
    nrows = 1000
    lens = rng.integers(0, 40, nrows)
    lens[::97] = 300
    lens[5:37] = 0            # a whole warp of empty rows
    row_map = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    entries = rng.integers(0, nrows, row_map[-1]).astype(np.int32)
    values = rng.standard_normal(row_map[-1])
    v = rng.standard_normal(nrows)
    o = oracle_mod.Oracle(6, 6, 1.0, 2.0, 1.0, 2.0)  # (any grid: spmv only needs the library)
    with capi.Plan(6, 6, domain=capi.DOMAIN_GENERIC, generic_rows=nrows) as p:
        p.set_csr(row_map, entries, values)
        assert np.array_equal(p.csr_apply(v), o.spmv((row_map, entries, values), v))


@pytest.mark.parametrize("n,a_tag", [(6, 1), (30, 1)])
def test_csr_assembly_matches_reference(capi, golden_ref, n, a_tag):
    tag = f"grid_n{n}_a{a_tag}"
    with plan_for(capi, n, a_tag) as p:
        nnz = p.assemble_csr()
        assert [p.N, nnz] == list(golden_ref[tag + "_shape"])
        row_map, entries, values = p.get_csr(nnz)
        assert np.array_equal(row_map, golden_ref[tag + "_row_map"])
        assert np.array_equal(entries, golden_ref[tag + "_entries"])  # per-row order diag, L, R, T, B
        assert np.array_equal(values, golden_ref[tag + "_values"])


@pytest.mark.parametrize("n,domain", [(128, 0), (700, 0), (333, 1)])
def test_csr_assembly_and_spmv_vs_oracle(capi, oracle_mod, n, domain):
    o = oracle_for(oracle_mod, n, 0, domain)
    csr = o.csr()
    with plan_for(capi, n, 0, domain) as p:
        nnz = p.assemble_csr()
        got = p.get_csr(nnz)
        for a, b in zip(got, csr):
            assert np.array_equal(a, b)
        x = np.random.default_rng(n).standard_normal(o.N)
        y = p.csr_apply(x)
        assert np.array_equal(y, o.spmv(csr, x))
        assert np.array_equal(y, p.apply(x))  # both operators accumulate in the same order
        # caller-supplied matrix (what MSGSolver receives)
        p.set_csr(*csr)
        assert np.array_equal(p.csr_apply(x), y)


# ---------------------------------------------------------------- DirichletSolver post-processing
def test_postprocess_matches_facade(capi, golden_ref):
    for op in (0, 1):
        with plan_for(capi, 30, 1) as p:
            if op:
                p.assemble_csr()
            x, info = p.solve(b=golden_ref["grid_n30_a1_rhs"], u=golden_ref["grid_n30_a1_true"], op=op,
                              rule=capi.RULE_MAXNORM, eps_p=1e-6, eps_r=1e-6, max_it=10000)
            res, err = p.postprocess(op=op)
            assert info["iterations"] == int(golden_ref["dirichlet_n30_info"][0])
            assert relmax(x, golden_ref["dirichlet_n30_solution"]) < REL
            ref_res = golden_ref["dirichlet_n30_residual"]
            # residual = A x - b: differences in x are amplified by |A| ~ 1e4
            assert np.max(np.abs(res - ref_res)) <= 1e-10 * np.max(np.abs(golden_ref["grid_n30_a1_rhs"]))
            assert np.max(np.abs(err - golden_ref["dirichlet_n30_error"])) <= REL * np.max(np.abs(x))
            assert np.array_equal(p.get_solution(), x)


# ---------------------------------------------------------------- RECT domain (parity unpinned: oracle only)
@pytest.mark.parametrize("n", [33, 200])
def test_rect_solve_vs_oracle(capi, oracle_mod, n):
    o = oracle_for(oracle_mod, n, 0, 1)
    ref = o.mf_solve(eps=1e-9, max_it=10000)
    with plan_for(capi, n, 0, 1) as p:
        x, info = p.solve(b=o.rhs(), eps_rel=1e-9, max_it=10000)
        assert abs(info["iterations"] - ref["iterations"]) <= 1
        assert relmax(x, ref["x"]) < REL
        assert np.max(np.abs(x - o.true_solution())) < 1e-3  # O(h^2) against the analytic solution


@pytest.mark.parametrize("n,m", [(7, 7), (9, 12), (33, 20), (1201, 777), (640, 640)])
def test_general_lshape_vs_oracle(capi, oracle_mod, n, m):
    """B200CG_DOMAIN_LSHAPE_ANY: the L-shaped region for odd / non-square grids (parity unpinned: the reference has no
    valid counterpart); identical to the reference geometry for even n == m."""
    o = oracle_mod.Oracle(m, n, 0.0, 1.0, 0.0, 1.0, oracle_mod.LSHAPE_ANY)
    with capi.Plan(m, n, 0.0, 1.0, 0.0, 1.0, domain=capi.DOMAIN_LSHAPE_ANY) as p:
        assert p.N == o.N
        v = np.random.default_rng(n).standard_normal(o.N)
        assert np.array_equal(p.apply(v), o.apply(v))
        p.build_rhs()
        assert relmax(p.get_rhs(), o.rhs()) < 1e-14
        nnz = p.assemble_csr()
        for a, b in zip(p.get_csr(nnz), o.csr()):
            assert np.array_equal(a, b)
        if n <= 64:
            ref = o.mf_solve(eps=1e-10, max_it=5000)
            for path in (0, 1):
                x, info = p.solve(b=o.rhs(), eps_rel=1e-10, max_it=5000, small_grid_path=path)
                assert abs(info["iterations"] - ref["iterations"]) <= 1 and relmax(x, ref["x"]) < REL
        if n == m and n % 2 == 0:
            with capi.Plan(m, n, 0.0, 1.0, 0.0, 1.0, domain=capi.DOMAIN_LSHAPE) as q:
                assert np.array_equal(q.apply(v), p.apply(v))


# ---------------------------------------------------------------- sizes of BASELINE.json
def test_config2_fixed_iterations_vs_oracle(capi, oracle_mod):
    """4096^2 (12.6 M unknowns): 5 iterations, x / ||r|| against the oracle at the same count.

    At this size the REFERENCE's sequential fp64 dot products are themselves only ~1e-10 accurate: the rhs has
    O(1/h^2) ~ 1e7 boundary entries, so the running sum reaches ~1e18 (ulp 128) while 12 M interior products are
    O(1e2). The GPU's tree sums do not share that error. The test pins this down: the GPU result matches the oracle
    run with long-double sums to ~1e-13, the reference-order oracle is the one that sits ~1e-10 away from both,
    and GPU vs reference-order stays inside 1e-9."""
    n = 4096
    o = oracle_for(oracle_mod, n)
    b = o.rhs()
    ref = o.mf_solve(b=b, eps=1e-8, max_it=5)
    acc = o.mf_solve(b=b, eps=1e-8, max_it=5, accurate_dots=True)
    with plan_for(capi, n) as p:
        x, info = p.solve(b=b, eps_rel=1e-8, max_it=5)
        assert info["iterations"] == 5
        d_gpu_acc, d_ref_acc, d_gpu_ref = relmax(x, acc["x"]), relmax(ref["x"], acc["x"]), relmax(x, ref["x"])
        assert d_gpu_acc < 1e-12
        assert d_gpu_ref < 1e-9
        assert d_gpu_acc <= d_ref_acc  # the deviation from the reference is the reference's own summation error
        assert abs(info["r_l2"] - acc["r_norm"]) <= 1e-12 * acc["r_norm"]
        assert abs(info["r_l2"] - ref["r_norm"]) <= 1e-9 * ref["r_norm"]
        assert abs(info["r0_l2"] - acc["r0_norm"]) <= 1e-13 * acc["r0_norm"]
        v = np.random.default_rng(0).standard_normal(o.N)
        assert np.array_equal(p.apply(v), o.apply(v))


def test_headline_size_vs_oracle(capi, oracle_mod):
    """16384^2 (BASELINE.json's headline grid, 201 M unknowns): the operator bit for bit and a 2-iteration solve of both
    iteration schemes against the oracle at the same count - in the reference's summation order and with long-double
    sums (see test_config2_fixed_iterations_vs_oracle for why the reference-order bar is wider at these sizes)."""
    n = 16384
    o = oracle_for(oracle_mod, n)
    b = o.rhs()
    with plan_for(capi, n) as p:
        v = np.random.default_rng(5).standard_normal(o.N)
        y = p.apply(v)
        assert np.array_equal(y, o.apply(v))
        del v, y
        xs, i_s = p.solve(b=b, eps_rel=1e-8, max_it=2)                  # default: single sweep per iteration
        xd, i_d = p.solve(b=b, eps_rel=1e-8, max_it=2, single_sweep=2)  # two sweeps per iteration
        assert i_s["single_sweep"] == 1 and i_d["single_sweep"] == 0
        assert i_s["iterations"] == i_d["iterations"] == 2
    acc = o.mf_solve(b=b, eps=1e-8, max_it=2, accurate_dots=True)
    xacc, acc_r, acc_r0 = acc["x"], acc["r_norm"], acc["r0_norm"]
    del acc
    ref = o.mf_solve(b=b, eps=1e-8, max_it=2)
    d_ref_acc = relmax(ref["x"], xacc)
    report = {}
    for name, x, info in (("single sweep", xs, i_s), ("two sweeps", xd, i_d)):
        report[name] = dict(x_vs_acc=relmax(x, xacc), x_vs_ref=relmax(x, ref["x"]), r_vs_acc=abs(info["r_l2"] - acc_r) / acc_r,
                            r_vs_ref=abs(info["r_l2"] - ref["r_norm"]) / ref["r_norm"], r0_vs_acc=abs(info["r0_l2"] - acc_r0) / acc_r0)
    report["reference order vs long-double sums"] = dict(x=d_ref_acc, r=abs(ref["r_norm"] - acc_r) / acc_r)
    print(report)
    for name in ("single sweep", "two sweeps"):
        d = report[name]
        # against long-double sums: the GPU's fp64 tree sums over 2e8 terms with entries up to 1e8 are good to ~1e-12
        # (measured: x 7.6e-13, ||r|| 1.5e-12, ||r0|| 2.2e-13) - the same size as the reference's own sequential-sum
        # error there (x 6.7e-13, ||r|| 2.6e-12); GPU vs reference order: x 1.4e-12, ||r|| 4.2e-12
        assert d["x_vs_acc"] < 5e-12, report
        assert d["r_vs_acc"] < 1e-11 and d["r0_vs_acc"] < 1e-12, report
        assert d["x_vs_ref"] < HEADLINE_REF_ORDER_BAR and d["r_vs_ref"] < HEADLINE_REF_ORDER_BAR, report


def test_config4_size_csr_vs_oracle(capi, oracle_mod):
    """8192^2 through the assembled operator (BASELINE.json configs[3], 50 M unknowns, 251 M non-zeros): device assembly
    against the oracle's, SpMV bit for bit (and equal to the matrix-free apply), 2 MSGSolver iterations at the same count."""
    n = 8192
    o = oracle_for(oracle_mod, n)
    b = o.rhs()
    csr = o.csr()
    with plan_for(capi, n) as p:
        nnz = p.assemble_csr()
        assert nnz == len(csr[2])
        row_map, entries, values = p.get_csr(nnz)
        assert np.array_equal(row_map, csr[0]) and np.array_equal(entries, csr[1]) and np.array_equal(values, csr[2])
        del row_map, entries, values
        v = np.random.default_rng(6).standard_normal(o.N)
        y = p.csr_apply(v)
        assert np.array_equal(y, o.spmv(csr, v)) and np.array_equal(y, p.apply(v))
        del v, y
        x, info = p.solve(b=b, op=capi.OP_CSR, rule=capi.RULE_MAXNORM, eps_p=-1.0, eps_r=1e-30, max_it=2)
        xm, info_m = p.solve(b=b, rule=capi.RULE_MAXNORM, eps_p=-1.0, eps_r=1e-30, max_it=2)
        assert info["iterations"] == info_m["iterations"] == 2
    ref = o.msg_solve(csr=csr, b=b, eps_p=-1.0, eps_r=1e-30, max_it=2)
    assert ref["iterations"] == 2
    assert relmax(x, ref["x"]) < 1e-9 and relmax(xm, ref["x"]) < 1e-9  # reference-order sums at 50 M terms
    assert relmax(x, xm) < 1e-13                                       # the two operators agree bit for bit; same sums
    assert abs(info["r_max"] - ref["r_max"]) <= 1e-9 * ref["r_max"]


def test_full_size_operator_properties(capi):
    """16384^2-class properties that need no oracle: linearity, symmetry, negative definiteness, and a CG
    run whose recurrence residual matches the recomputed one."""
    n = 8192
    with plan_for(capi, n) as p:
        rng = np.random.default_rng(11)
        x, y = rng.standard_normal(p.N), rng.standard_normal(p.N)
        Ax, Ay = p.apply(x), p.apply(y)
        assert abs(np.dot(x, Ay) - np.dot(y, Ax)) <= 1e-9 * abs(np.dot(x, Ay)) + 1e-3  # symmetric
        assert np.dot(x, Ax) < 0  # negative definite Laplacian
        Axy = p.apply(x + 2.0 * y)
        assert np.max(np.abs(Axy - (Ax + 2.0 * Ay))) <= 1e-12 * np.max(np.abs(Axy))
        p.build_rhs()
        b = p.get_rhs()
        xs, info = p.solve(rhs_on_device=True, eps_rel=1e-30, max_it=200)
        assert info["iterations"] == 200 and info["r_l2"] < info["r0_l2"]
        res, _ = p.postprocess(want_error=False)  # A x - b recomputed
        assert abs(np.linalg.norm(res) - info["r_l2"]) <= 1e-8 * info["r0_l2"]
        assert np.max(np.abs(res + (b - p.apply(xs)))) <= 1e-12 * np.max(np.abs(b))


# ---------------------------------------------------------------- execution-strategy knobs must not change the answer
def _solve_with_env(capi, env, n, iters, **solve_kw):
    import os

    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        with plan_for(capi, n) as p:  # the knobs are read when the plan is created
            p.build_rhs()
            return p.solve(rhs_on_device=True, eps_rel=0.0, max_it=iters, **solve_kw)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_x_deferral_is_bit_identical(capi):
    """Two-sweep iteration: x touched every other iteration (its default under the relative-residual rule) vs every iteration: the same additions in
    the same order, so with the same fixed work split (same launch shape for both update flavours, no balancing) the
    solutions are bit-identical - for odd and even iteration counts."""
    common = {"B200CG_BALANCE": "0", "B200CG_SHAPE_NOX": "2", "B200CG_SHAPE_UPD": "2"}
    for iters in (7, 40):
        xa, ia = _solve_with_env(capi, dict(common, B200CG_XDEFER="1"), 2048, iters, single_sweep=2)
        xb, ib = _solve_with_env(capi, dict(common, B200CG_XDEFER="0"), 2048, iters, single_sweep=2)
        assert ia["x_deferral"] == 1 and ib["x_deferral"] == 0
        assert ia["iterations"] == ib["iterations"] == iters
        assert np.array_equal(xa, xb)
        assert ia["r_l2"] == ib["r_l2"]


def test_feedback_balancing_changes_only_the_summation_order(capi):
    """The work split is re-cut from measured per-CTA times during the first graph launches; iterates may then differ
    from the fixed split only by dot-product rounding."""
    for ss in (0, 2):  # the default single-sweep iteration and the two-sweep one
        xa, ia = _solve_with_env(capi, {"B200CG_BALANCE": "4"}, 4096, 120, iters_per_graph=20, single_sweep=ss)
        xb, ib = _solve_with_env(capi, {"B200CG_BALANCE": "0"}, 4096, 120, iters_per_graph=20, single_sweep=ss)
        assert ia["iterations"] == ib["iterations"] == 120 and ia["single_sweep"] == ib["single_sweep"] == (1 if ss == 0 else 0)
        assert relmax(xa, xb) < 1e-12
        assert abs(ia["r_l2"] - ib["r_l2"]) <= 1e-12 * ib["r_l2"]
        # fixed split: reproducible bit for bit from run to run
        xc, ic = _solve_with_env(capi, {"B200CG_BALANCE": "0"}, 4096, 120, iters_per_graph=20, single_sweep=ss)
        assert np.array_equal(xb, xc) and ib["r_l2"] == ic["r_l2"]


def test_launch_shapes_do_not_change_the_answer(capi):
    ref, _ = _solve_with_env(capi, {"B200CG_BALANCE": "0"}, 1536, 30, single_sweep=2)
    for dot, upd, nox in [(0, 0, 0), (1, 0, 1), (2, 2, 2), (3, 2, 3)]:
        x, info = _solve_with_env(capi, {"B200CG_BALANCE": "0", "B200CG_SHAPE_DOT": str(dot), "B200CG_SHAPE_UPD": str(upd),
                                         "B200CG_SHAPE_NOX": str(nox)}, 1536, 30, single_sweep=2)
        assert info["iterations"] == 30 and relmax(x, ref) < 1e-12
