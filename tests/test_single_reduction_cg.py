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
