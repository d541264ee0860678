"""CPU: the numpy restatement of the opt-in multigrid-preconditioned CG (oracle/mg_oracle.py). The reference has no
preconditioner, so parity is unpinned for this path by construction; these tests anchor the restatement to what the
reference does define: the operator (bit for bit against the reference-pinned oracle), the stop rule, and the solution
(against the reference-order plain CG solve and the analytic solution)."""
import numpy as np
import pytest

from oracle import mg_oracle as mg


@pytest.mark.parametrize("n,lshape", [(8, True), (16, True), (64, True), (20, False), (33, False)])
def test_level_operator_is_the_reference_operator(oracle_mod, n, lshape):
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, oracle_mod.LSHAPE if lshape else oracle_mod.RECT)
    S = mg.MgPcg(n, n, lshape=lshape)
    v = np.random.default_rng(n).standard_normal(o.N)
    got = mg.from_grid(S.levels[0].apply(mg.to_grid(v, n, n, lshape)), n, n, lshape)
    assert np.array_equal(got, o.apply(v))


def test_hierarchy_rules():
    assert [(L.n, L.m) for L in mg.MgPcg(128, 128).levels] == [(128, 128), (64, 64), (32, 32), (16, 16), (8, 8), (4, 4)]
    assert len(mg.MgPcg(30, 30).levels) == 1          # n/2 = 15: the re-entrant corner would leave the coarse grid lines
    assert [L.n for L in mg.MgPcg(24, 24).levels] == [24, 12, 6]  # 6 < 8: coarsest
    assert [L.n for L in mg.MgPcg(40, 40).levels] == [40, 20, 10]  # 10 / 2 = 5 is odd: the corner would leave the grid lines
    assert [(L.n, L.m) for L in mg.MgPcg(50, 200, lshape=False).levels] == [(200, 50), (100, 25)]


@pytest.mark.parametrize("n", [16, 64, 128, 256])
def test_iteration_count_is_independent_of_n_and_the_solution_is_the_cg_solution(oracle_mod, n):
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, 0)
    b, u = o.rhs(), o.true_solution()
    out = mg.MgPcg(n, n).solve(mg.to_grid(b, n, n, True), eps=1e-8, max_it=100)
    assert out["converged"] and out["iterations"] <= 8
    assert out["r_norm"] <= 1e-8 * out["r0_norm"]
    x = mg.from_grid(out["x"], n, n, True)
    # true residual of the returned x: the recurrence residual did not drift
    assert np.linalg.norm(b - o.apply(x)) <= 2e-8 * np.linalg.norm(b)
    ref = o.mf_solve(b=b, eps=1e-10, max_it=20000)  # the reference-order plain CG, converged two digits further
    assert np.max(np.abs(x - ref["x"])) <= 1e-7 * np.max(np.abs(ref["x"]))
    # discretisation error O(h^2) against the analytic solution (1.15e-5 at n = 128)
    assert np.max(np.abs(x - u)) <= 0.25 * (128.0 / n) ** 2 * 1e-4


def test_rect_and_uncoarsenable_grids_still_converge(oracle_mod):
    for n, lshape in [(30, True), (33, False), (100, False)]:
        o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, oracle_mod.LSHAPE if lshape else oracle_mod.RECT)
        b = o.rhs()
        out = mg.MgPcg(n, n, lshape=lshape).solve(mg.to_grid(b, n, n, lshape), eps=1e-8, max_it=2000)
        assert out["converged"]
        x = mg.from_grid(out["x"], n, n, lshape)
        assert np.linalg.norm(b - o.apply(x)) <= 2e-8 * np.linalg.norm(b)


def test_edge_cases():
    S = mg.MgPcg(16, 16)
    z = np.zeros((17, 17))
    out = S.solve(z, eps=1e-8, max_it=10)
    assert out["iterations"] == 0 and out["converged"] and not np.any(out["x"])
    b = np.where(S.levels[0].mask, 1.0, 0.0)
    out = S.solve(b, eps=1e-8, max_it=0)
    assert out["iterations"] == 0 and not out["converged"]
    out = S.solve(b, eps=1e-30, max_it=3)
    assert out["iterations"] == 3 and not out["converged"]
