"""Host logic of the sweep kernels' work split (b200cg_work_split, csrc/plan.cu: build_tiles): whatever the grid,
the slab, the launch shape and the balancing weights, the tiles dealt to the CTAs must cover every unknown of the
rank exactly once and nothing else. Runs without a GPU.

The unknown set restates the reference's domain (grid_system.cpp:17-43, :79-111): interior nodes of the rectangle
minus its lower-left quadrant."""
import numpy as np
import pytest

from iterative_solvers_b200 import capi

STRIP_OUT = 504  # csrc/common.cuh


def unknown_mask(n, m, domain):
    mask = np.zeros((m + 1, n + 1), dtype=np.int32)
    mask[1:m, 1:n] = 1
    if domain in (capi.DOMAIN_LSHAPE, capi.DOMAIN_LSHAPE_ANY):
        mask[1:m // 2 + 1, 1:n // 2 + 1] = 0
    return mask


FUSED_STRIP_OUT, FUSED_SHIFT = 420, 2  # csrc/fused_kernel.cuh: 7 consumer warps x 60 columns


def coverage(n, m, tiles, strip_out=STRIP_OUT, shift=0):
    cov = np.zeros((m + 1, n + 1), dtype=np.int32)
    for col0, ya, yb, xlo in tiles:
        x0, x1 = max(col0 - shift, xlo), min(col0 - shift + strip_out - 1, n - 1)
        assert (col0 - shift) % strip_out == 0 and ya < yb
        if x0 <= x1:
            cov[ya:yb, x0:x1 + 1] += 1
    return cov


def check_split(n, m, domain, world, sms, ctas, weights=None, tile_rows=0, fused=False):
    total = np.zeros((m + 1, n + 1), dtype=np.int32)
    for rank in range(world):
        ylo, yhi, lo, hi, N = capi.partition(m, n, domain=domain, rank=rank, world=world)
        tiles, cta_begin = capi.work_split(m, n, domain=domain, rank=rank, world=world, sms=sms, ctas_per_sm=ctas,
                                           weights=weights, tile_rows=tile_rows, fused=fused)
        grid = len(cta_begin) - 1
        assert 1 <= grid <= sms * ctas
        assert cta_begin[0] == 0 and cta_begin[-1] == len(tiles) and np.all(np.diff(cta_begin) >= 0)
        assert np.all(tiles[:, 1] >= ylo) and np.all(tiles[:, 2] <= yhi)
        if domain != capi.DOMAIN_RECT:  # a tile never straddles the two blocks of the L
            lower = tiles[:, 2] <= m // 2 + 1
            assert np.all(lower | (tiles[:, 1] > m // 2))
            assert np.all(tiles[lower & (tiles[:, 1] <= m // 2), 3] == n // 2 + 1)
            assert np.all(tiles[tiles[:, 1] > m // 2, 3] == 1)
        cov = coverage(n, m, tiles, FUSED_STRIP_OUT * int(fused), FUSED_SHIFT) if fused else coverage(n, m, tiles)
        assert cov.sum() == hi - lo
        total += cov
    assert np.array_equal(total, unknown_mask(n, m, domain))


@pytest.mark.parametrize("n", [4, 6, 30, 128, 500, 504, 506, 1010, 2048])
@pytest.mark.parametrize("domain", [capi.DOMAIN_LSHAPE, capi.DOMAIN_RECT])
def test_equal_split_covers_every_unknown_once(n, domain):
    for ctas in (2, 3):
        check_split(n, n, domain, 1, 148, ctas)


@pytest.mark.parametrize("n,m", [(5, 7), (31, 64), (777, 333), (1200, 90), (90, 1200)])
def test_general_l_shape_and_rectangles(n, m):
    check_split(n, m, capi.DOMAIN_LSHAPE_ANY, 1, 148, 2)
    check_split(n, m, capi.DOMAIN_RECT, 1, 148, 3)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_row_slabs(world):
    check_split(1024, 1024, capi.DOMAIN_LSHAPE, world, 148, 2)
    check_split(300, 300, capi.DOMAIN_RECT, world, 148, 3)
    check_split(40, 40, capi.DOMAIN_LSHAPE, world, 148, 2)


def test_fixed_tile_rows():
    for tr in (1, 3, 7, 64):
        check_split(256, 256, capi.DOMAIN_LSHAPE, 1, 148, 2, tile_rows=tr)
        check_split(256, 256, capi.DOMAIN_LSHAPE, 2, 148, 3, tile_rows=tr)


def test_balancing_weights_keep_the_cover_and_the_table_bound():
    rng = np.random.default_rng(7)
    n, sms, ctas = 2048, 148, 2
    strips = (n - 1) // STRIP_OUT + 1
    base, _ = capi.work_split(n, n, sms=sms, ctas_per_sm=ctas)
    bound = len(base) + 4 * (strips + 2) + 64  # upload_tiles allocates the device table once with this slack
    for _ in range(6):
        w = rng.uniform(0.8, 1.25, size=sms * ctas)
        tiles, cta_begin = capi.work_split(n, n, sms=sms, ctas_per_sm=ctas, weights=w)
        assert len(tiles) <= bound
        check_split(n, n, capi.DOMAIN_LSHAPE, 1, sms, ctas, weights=w)
    # shares follow the weights: a CTA with twice the weight gets about twice the rows
    w = np.ones(sms * ctas)
    w[: sms] = 2.0
    tiles, cta_begin = capi.work_split(n, n, sms=sms, ctas_per_sm=ctas, weights=w)
    rows = np.array([sum(int(t[2] - t[1]) for t in tiles[cta_begin[c]:cta_begin[c + 1]]) for c in range(sms * ctas)])
    ratio = rows[:sms].mean() / rows[sms:].mean()
    assert 1.8 < ratio < 2.2


def test_small_machines_and_errors():
    check_split(128, 128, capi.DOMAIN_LSHAPE, 1, 1, 1)
    check_split(128, 128, capi.DOMAIN_LSHAPE, 1, 4, 3)
    with pytest.raises(capi.B200CGError):
        capi.work_split(128, 128, sms=0)
    with pytest.raises(capi.B200CGError):
        capi.work_split(129, 129)  # odd n: not a reference grid


@pytest.mark.parametrize("n", [6, 30, 418, 420, 422, 842, 2048])
def test_single_sweep_strip_geometry(n):
    """The single-sweep kernel cuts strips of 420 written columns (desc.reserved0 = 1)."""
    check_split(n, n, capi.DOMAIN_LSHAPE, 1, 148, 2, fused=True)
    check_split(n, n + 3, capi.DOMAIN_RECT, 1, 148, 2, fused=True, tile_rows=3)
    check_split(n + 1, n, capi.DOMAIN_LSHAPE_ANY, 1, 4, 2, fused=True)
    check_split(n, n + 1, capi.DOMAIN_RECT, 2, 16, 2, fused=True) if n >= 8 else None
    # the wide geometry of large slabs (desc.reserved0 = 2): 840 written columns, one CTA per SM
    check_split(n, n, capi.DOMAIN_LSHAPE, 1, 148, 1, fused=2)
    check_split(2 * n + 2, n + 3, capi.DOMAIN_RECT, 3, 16, 1, fused=2) if n >= 8 else None
