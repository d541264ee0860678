"""CPU check behind the opt-in single-sweep iteration (csrc/fused_kernel.cuh): forming alpha from the single-reduction
CG recurrence (Chronopoulos-Gear) instead of p.Ap leaves the reference's iteration counts and its solution (to
rounding) unchanged. tests/studies/single_reduction_cg.py runs the same comparison up to 1100^2."""
import importlib.util
import os

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
