"""GPU: the opt-in multigrid-preconditioned CG (b200cg_params.preconditioner = B200CG_PRECOND_MULTIGRID, csrc/mg.cu)
through the C ABI against its numpy restatement (oracle/mg_oracle.py: same operations in the same order, so the same
iteration counts and iterates up to dot-product rounding), against the plain CG solve and the analytic solution."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from iterative_solvers_b200 import capi as c

    c.lib()
    assert c.device_count() >= 1, "these tests need a CUDA device"
    return c


def relmax(x, ref):
    return np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300)


@pytest.mark.parametrize("n,domain,eps", [(8, 0, 1e-8), (16, 0, 1e-10), (64, 0, 1e-8), (128, 0, 1e-8), (256, 0, 1e-8),
                                          (600, 0, 1e-8), (30, 0, 1e-8), (24, 0, 1e-8), (200, 1, 1e-8), (33, 1, 1e-8),
                                          (1024, 0, 1e-9)])
def test_matches_the_numpy_restatement(capi, oracle_mod, n, domain, eps):
    from oracle import mg_oracle as mg

    lshape = domain == 0
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, domain)
    b = o.rhs()
    S = mg.MgPcg(n, n, lshape=lshape)
    ref = S.solve(mg.to_grid(b, n, n, lshape), eps=eps, max_it=3000)
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, domain=domain) as p:
        x, info = p.solve(b=b, eps_rel=eps, max_it=3000, preconditioner=capi.PRECOND_MULTIGRID)
        assert info["preconditioner"] == 1 and info["mg_levels"] == ref["levels"]
        assert info["single_sweep"] == 0 and info["cluster_path"] == 0
        assert abs(info["iterations"] - ref["iterations"]) <= 1 and info["converged"]
        if info["iterations"] == ref["iterations"]:
            assert relmax(x, mg.from_grid(ref["x"], n, n, lshape)) < 1e-10
            assert abs(info["r_l2"] - ref["r_norm"]) <= 1e-6 * ref["r_norm"] + 1e-14 * ref["r0_norm"]
        assert abs(info["r0_l2"] - ref["r0_norm"]) <= 1e-13 * ref["r0_norm"]
        res, _ = p.postprocess(want_error=False)
        assert np.linalg.norm(res) <= 2.0 * eps * np.linalg.norm(b)
        assert np.array_equal(res, o.apply(x) - b)


def test_same_solution_as_plain_cg_in_a_fraction_of_the_iterations(capi):
    """4096^2 (BASELINE.json configs[1]): plain CG to 1e-9 needs ~12 000 iterations, the preconditioned one 8."""
    n = 4096
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        p.build_rhs()
        u = p.true_solution()
        p.solve(rhs_on_device=True, eps_rel=1e-9, max_it=1, preconditioner=capi.PRECOND_MULTIGRID)  # builds the hierarchy
        xm, im = p.solve(rhs_on_device=True, eps_rel=1e-9, max_it=200, preconditioner=capi.PRECOND_MULTIGRID)
        assert im["converged"] and im["iterations"] <= 10
        res, _ = p.postprocess(want_error=False)
        b = p.get_rhs()
        assert np.linalg.norm(res) <= 2e-9 * np.linalg.norm(b)
        xc, ic = p.solve(rhs_on_device=True, eps_rel=1e-9, max_it=40000)
        assert ic["converged"] and ic["iterations"] > 100 * im["iterations"]
        # Both stop at the same relative residual, but the plain iteration's ALGEBRAIC error at that point is the larger
        # one (kappa ~ n^2): at eps = 1e-8 it sits 4.8e-6 from the analytic solution, the preconditioned solve 1.4e-7
        # (profiles/r2_converged_runs.md). The two solutions agree to the plain solve's error level.
        e_m, e_c = np.max(np.abs(xm - u)), np.max(np.abs(xc - u))
        assert relmax(xm, xc) < 1e-6
        assert e_m <= e_c and e_m < 5e-8
        assert im["solve_ms"] < 0.1 * ic["solve_ms"]


def test_edge_cases_and_refusals(capi, oracle_mod):
    o = oracle_mod.Oracle(64, 64, 0.0, 1.0, 0.0, 1.0, 0)
    b, u = o.rhs(), o.true_solution()
    with capi.Plan(64, 64, 0.0, 1.0, 0.0, 1.0) as p:
        x, info = p.solve(b=np.zeros_like(b), eps_rel=1e-8, max_it=10, preconditioner=1)
        assert info["iterations"] == 0 and not np.any(x)
        x, info = p.solve(b=b, eps_rel=1e-8, max_it=0, preconditioner=1)
        assert info["iterations"] == 0 and not info["converged"]
        x, info = p.solve(b=b, eps_rel=1e-30, max_it=3, preconditioner=1)
        assert info["iterations"] == 3 and not info["converged"] and info["stop_reason"] == "ITERATIONS"
        # a second solve on the same plan reuses the hierarchy
        x1, i1 = p.solve(b=b, eps_rel=1e-8, max_it=100, preconditioner=1)
        x2, i2 = p.solve(b=b, eps_rel=1e-8, max_it=100, preconditioner=1)
        assert np.array_equal(x1, x2) and i1["iterations"] == i2["iterations"]
        # and the plain solve after it is untouched by the scratch use
        ref = o.mf_solve(b=b, eps=1e-8, max_it=10000)
        x3, i3 = p.solve(b=b, eps_rel=1e-8, max_it=10000)
        assert i3["iterations"] == ref["iterations"] and relmax(x3, ref["x"]) < 1e-10
        with pytest.raises(capi.B200CGError):
            p.solve(b=b, rule=capi.RULE_MAXNORM, eps_r=1e-8, max_it=100, preconditioner=1)
        with pytest.raises(capi.B200CGError):
            p.solve(b=b, eps_rel=1e-8, max_it=100, preconditioner=1, callback=lambda *a: None)
        with pytest.raises(capi.B200CGError):
            p.solve(b=b, eps_rel=1e-8, max_it=100, preconditioner=7)
        p.assemble_csr()
        with pytest.raises(capi.B200CGError):
            p.solve(b=b, op=capi.OP_CSR, eps_rel=1e-8, max_it=100, preconditioner=1)
