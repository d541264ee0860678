"""Property tests (hypothesis, CPU): the row-slab partition of the C ABI and the oracle's operator, over random grid
shapes, domains and rank counts - the size-independent facts the GPU parity tests lean on."""
import numpy as np
from hypothesis import given, settings, strategies as st

from iterative_solvers_b200 import capi


def unknowns(n, m, domain):
    total = (n - 1) * (m - 1)
    if domain != capi.DOMAIN_RECT:
        total -= (n // 2) * (m // 2)  # the lower-left quadrant x <= n/2, y <= m/2 (interior nodes)
    return total


@settings(max_examples=150, deadline=None)
@given(st.integers(4, 3000), st.integers(4, 3000), st.sampled_from([capi.DOMAIN_RECT, capi.DOMAIN_LSHAPE_ANY]),
       st.integers(1, 16))
def test_partition_tiles_the_rows_and_balances_unknowns(n, m, domain, world):
    if world > m - 1:
        world = m - 1
    prev_hi, prev_yhi, counts = 0, 1, []
    for rank in range(world):
        ylo, yhi, lo, hi, N = capi.partition(m, n, domain=domain, rank=rank, world=world)
        assert N == unknowns(n, m, domain)
        assert ylo == prev_yhi and lo == prev_hi and yhi > ylo  # contiguous, at least one row each
        prev_hi, prev_yhi = hi, yhi
        counts.append(hi - lo)
    assert prev_yhi == m and prev_hi == unknowns(n, m, domain)
    # balanced by unknowns up to a couple of rows (a row is at most n - 1 unknowns wide)
    assert max(counts) - min(counts) <= 2 * (n - 1) + (N // world if world > m // 4 else 0)


@settings(max_examples=40, deadline=None)
@given(st.integers(4, 40), st.integers(4, 40), st.sampled_from([1, 3]), st.integers(0, 2**31 - 1))
def test_oracle_operator_is_symmetric_negative_definite(oracle_mod, n, m, kind, seed):
    """kind 1 = RECT, 3 = LSHAPE_ANY (oracle/cg_oracle.h). CG needs A = A^T < 0 (the reference solves A x = b with the
    negative Laplacian, grid_system.cpp:316-318); the stencil couples only nodes that are both unknowns."""
    o = oracle_mod.Oracle(m, n, 0.0, 1.0, 0.0, 2.0, kind)
    rng = np.random.default_rng(seed)
    x, y = rng.standard_normal(o.N), rng.standard_normal(o.N)
    ax, ay = o.apply(x), o.apply(y)
    assert abs(np.dot(y, ax) - np.dot(x, ay)) <= 1e-10 * (np.linalg.norm(ax) * np.linalg.norm(y) + 1e-300)
    assert np.dot(x, ax) < 0
    # linearity
    assert np.max(np.abs(o.apply(2.0 * x - 3.0 * y) - (2.0 * ax - 3.0 * ay))) <= 1e-9 * (np.max(np.abs(ax)) + np.max(np.abs(ay)))
    # the assembled matrix is the same operator
    csr = o.csr()
    assert np.array_equal(o.spmv(csr, x), ax)
