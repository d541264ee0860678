"""CPU check behind the opt-in single-sweep iteration (csrc/fused_kernel.cuh): forming alpha from the single-reduction
CG recurrence (Chronopoulos-Gear) instead of p.Ap leaves the reference's iteration counts and its solution (to
rounding) unchanged. tests/studies/single_reduction_cg.py runs the same comparison up to 1100^2."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def study(oracle_mod):  # oracle_mod builds the oracle library
    spec = importlib.util.spec_from_file_location("single_reduction_cg", os.path.join(HERE, "studies", "single_reduction_cg.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("n,iters", [(64, 178), (128, 352)])
def test_single_reduction_cg_follows_the_reference(study, n, iters):
    it_oracle, it_numpy, it_single, d_numpy, d_single = study.compare(n, 1e-8)
    assert it_oracle == it_numpy == it_single == iters  # pinned counts
    assert d_numpy < 1e-12 and d_single < 1e-12


@pytest.mark.parametrize("n,a_tag,iters", [(6, 1, 13), (30, 1, 88), (64, 0, 178), (128, 0, 352), (128, 1, 362), (256, 0, 690),
                                           (512, 0, 1342)])
def test_c_oracle_single_reduction_variant(oracle_mod, n, a_tag, iters):
    """The same statement in the reference's own arithmetic (oracle/cg_oracle.c: cgo_mf_solve_single vs cgo_mf_solve,
    sequential dots, reference apply order) on every golden grid: identical iteration counts, x at rounding distance."""
    a, b = {0: (0.0, 1.0), 1: (1.0, 2.0)}[a_tag]
    o = oracle_mod.Oracle(n, n, a, b, a, b)
    ref = o.mf_solve(eps=1e-8, max_it=20000)
    one = o.mf_solve_single(eps=1e-8, max_it=20000)
    assert ref["iterations"] == one["iterations"] == iters
    assert one["converged"]
    assert np.max(np.abs(one["x"] - ref["x"])) <= 1e-12 * np.max(np.abs(ref["x"]))
    assert abs(one["r_norm"] - ref["r_norm"]) <= 1e-6 * ref["r_norm"]


@pytest.mark.parametrize("n,a_tag", [(6, 1), (30, 1), (128, 0)])
def test_c_oracle_single_reduction_variant_of_the_maxnorm_rules(oracle_mod, golden_ref, n, a_tag):
    """MSGSolver's rules with alpha from the recurrence (cgo_msg_solve_single - what the product's single sweep computes
    under the max-norm rules) against the fixtures the unmodified reference produced: the pinned 14 / 79 / 102 / 355 / 482
    iterations, the same stop reasons, x and the reported norms at rounding distance, the same callback cadence."""
    a, b = {0: (0.0, 1.0), 1: (1.0, 2.0)}[a_tag]
    tag = f"grid_n{n}_a{a_tag}"
    eps = 1e-6 if n <= 30 else 1e-8
    o = oracle_mod.Oracle(n, n, a, b, a, b)
    rhs, u = golden_ref[tag + "_rhs"], golden_ref[tag + "_true"]
    for cname, kw in {"pr": dict(eps_p=eps, eps_r=eps), "r": dict(eps_p=-1.0, eps_r=eps)}.items():
        info_ref = golden_ref[f"{tag}_msg_{cname}_info"]  # iterations, converged, stop reason, r_max, dx_max, err_max
        cb_ref = golden_ref[f"{tag}_msg_{cname}_cb"]
        one = o.msg_solve_single(b=rhs, u=u, max_it=10000, cb_cap=64, **kw)
        assert one["iterations"] == int(info_ref[0]) and one["converged"] == bool(info_ref[1])
        assert one["stop_reason"] == oracle_mod.STOP_NAMES[int(info_ref[2])]
        assert np.max(np.abs(one["x"] - golden_ref[f"{tag}_msg_{cname}_x"])) <= 1e-12 * np.max(np.abs(u))
        assert abs(one["dx_max"] - info_ref[4]) <= 1e-6 * info_ref[4]
        assert abs(one["err_max"] - info_ref[5]) <= 1e-9 * info_ref[5]
        assert abs(one["r_max"] - info_ref[3]) <= 1e-3 * info_ref[3] + 1e-12 * np.max(np.abs(rhs))
        assert np.array_equal(one["callbacks"][:, 0], cb_ref[:, 0])
        assert np.allclose(one["callbacks"][1:, 1:], cb_ref[1:, 1:], rtol=1e-3, atol=0)


@pytest.mark.parametrize("n,with_u", [(64, True), (200, True), (256, False)])
def test_maxnorm_single_reduction_vs_reference_order_on_larger_grids(oracle_mod, n, with_u):
    """Beyond the fixtures: cgo_msg_solve (the reference's arithmetic on the assembled matrix) and the single-reduction variant
    stop at the same iteration for the same reason, also under the exact-error rule."""
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0)
    u = o.true_solution() if with_u else None
    rules = [dict(eps_p=1e-8, eps_r=1e-8), dict(eps_p=-1.0, eps_r=1e-7)]
    if with_u:
        rules.append(dict(eps_p=-1.0, eps_r=-1.0, eps_e=2e-4 if n == 64 else 2e-5))
    for kw in rules:
        ref = o.msg_solve(u=u, max_it=20000, **kw)
        one = o.msg_solve_single(u=u, max_it=20000, **kw)
        assert one["iterations"] == ref["iterations"] and one["stop_reason"] == ref["stop_reason"], (kw, ref["iterations"])
        assert np.max(np.abs(one["x"] - ref["x"])) <= 1e-12 * np.max(np.abs(ref["x"]))


@pytest.mark.parametrize("tag", ["n30_a1_exact", "n30_a1_exact_first", "n30_a1_nou", "n64_a0_pr", "n64_a0_r", "n64_a0_cap"])
def test_maxnorm_single_reduction_variant_on_the_extra_fixtures(oracle_mod, golden_msg, tag):
    """The single-reduction statement of MSGSolver's rules against more runs of the unmodified reference: exact-error rule,
    no true solution, 64 x 64, iteration cap - same iteration, same reason, x at rounding distance, same callback cadence."""
    n, a, b, eps_p, eps_r, eps_e, max_it, with_true = golden_msg[f"msg_{tag}_params"]
    o = oracle_mod.Oracle(int(n), int(n), a, b, a, b)
    u = o.true_solution() if with_true else None
    one = o.msg_solve_single(u=u, eps_p=eps_p, eps_r=eps_r, eps_e=eps_e, max_it=int(max_it), cb_cap=256)
    info, x_ref, cb_ref = golden_msg[f"msg_{tag}_info"], golden_msg[f"msg_{tag}_x"], golden_msg[f"msg_{tag}_cb"]
    assert one["iterations"] == int(info[0]) and one["converged"] == bool(info[1])
    assert oracle_mod.STOP_NAMES.index(one["stop_reason"]) == int(info[2])
    assert np.max(np.abs(one["x"] - x_ref)) <= 1e-12 * np.max(np.abs(x_ref))
    assert abs(one["dx_max"] - info[4]) <= 1e-6 * info[4]
    assert one["err_max"] == info[5] if not with_true else abs(one["err_max"] - info[5]) <= 1e-9 * info[5] + 1e-12
    assert np.array_equal(one["callbacks"][:, 0], cb_ref[:, 0])
