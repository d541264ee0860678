"""The lane-level CPU model of the single-sweep kernel's data flow (scripts/model_single_sweep.py) on the real tile
tables of b200cg_work_split: strips of 420 written columns (7 consumer warps), 64-column warp windows, the two-deep row pipeline, per-row
masks across the two blocks of the L, stale shared-memory columns as NaN. It pins the index arithmetic the CUDA kernel
(csrc/fused_kernel.cuh) implements, including the unmasked inputs of FULL stages. No GPU needed."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    spec = importlib.util.spec_from_file_location("model_single_sweep", os.path.join(ROOT, "scripts", "model_single_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("n,m,lshape,iters,tile_rows", [(30, 30, True, 5, 0), (64, 64, True, 4, 5), (70, 46, True, 4, 0),
                                                        (77, 33, False, 4, 3), (500, 24, True, 3, 0)])
def test_model_matches_plain_single_reduction_cg(model, n, m, lshape, iters, tile_rows):
    worst, dx, dr, ntiles = model.run(n, m, lshape, iters, tile_rows)
    assert ntiles >= 1
    assert worst < 1e-12 and dx < 1e-10 and dr < 1e-10


@pytest.mark.parametrize("n,m,lshape,iters,world,tile_rows", [(64, 64, True, 5, 2, 0), (70, 60, True, 4, 3, 0),
                                                              (77, 60, False, 4, 3, 5), (96, 96, True, 4, 8, 0)])
def test_sharded_model(model, n, m, lshape, iters, world, tile_rows):
    """F_SHARD: row slabs with two halo rows per side (the neighbours' halo row + one of the two extra rows behind the
    stored rows), filled by the neighbours' sweeps; sums added over the ranks."""
    worst, dx, dr = model.run_sharded(n, m, lshape, iters, world, tile_rows=tile_rows)
    assert worst < 1e-12 and dx < 1e-12 and dr < 1e-12


@pytest.mark.parametrize("n,m,lshape,iters,tile_rows", [(64, 64, True, 4, 0), (900, 30, True, 3, 0), (430, 26, False, 3, 4),
                                                        (845, 64, True, 3, 0)])
def test_strips_with_stale_columns_and_full_stages(model, n, m, lshape, iters, tile_rows):
    """Several strips, the last one mostly beyond the row pitch (stale shared memory = NaN in the model), tiles tall enough
    for FULL stages, strips crossing the re-entrant edge of the L."""
    worst, dx, dr, ntiles = model.run(n, m, lshape, iters, tile_rows)
    assert worst < 1e-12 and dx < 1e-12 and dr < 1e-12


@pytest.mark.parametrize("n,m,lshape,iters,tile_rows", [(64, 64, True, 4, 0), (900, 30, True, 3, 0), (1700, 26, False, 3, 4),
                                                        (1690, 64, True, 3, 0)])
def test_wide_geometry_model(model, n, m, lshape, iters, tile_rows):
    """Slabs of >= 4 M unknowns run one 15-warp CTA per SM on 840-column strips (14 consumer warps): same data flow."""
    worst, dx, dr, ntiles = model.run(n, m, lshape, iters, tile_rows, warps=14)
    assert worst < 1e-12 and dx < 1e-12 and dr < 1e-12


@pytest.mark.parametrize("n,m,lshape,iters,tile_rows,warps,with_u", [(30, 30, True, 5, 0, 7, True), (64, 64, True, 5, 5, 7, True),
                                                                    (130, 90, True, 4, 0, 7, False), (430, 26, False, 3, 4, 7, True),
                                                                    (845, 64, True, 3, 0, 7, True), (900, 30, True, 3, 0, 14, True),
                                                                    (1700, 26, False, 3, 4, 14, False)])
def test_maxnorm_flavour_model(model, n, m, lshape, iters, tile_rows, warps, with_u):
    """F_MAXN (MSGSolver's rules in one sweep): x every iteration, |r'|_inf, |x' - x|_inf, |x' - u|_inf under the kernel's
    selects - FULL stages hand x and u on unmasked, so a stale (NaN) column would surface in a maximum."""
    worst, dx, ntiles = model.run_maxn(n, m, lshape, iters, tile_rows, warps=warps, with_u=with_u)
    assert ntiles >= 1 and worst < 1e-12 and dx < 1e-12


@pytest.mark.parametrize("n,m,lshape,iters,world,tile_rows,with_u", [(64, 64, True, 5, 2, 0, True), (70, 60, True, 4, 3, 0, True),
                                                                    (77, 60, False, 4, 3, 5, False), (96, 96, True, 4, 8, 0, True)])
def test_sharded_maxnorm_flavour_model(model, n, m, lshape, iters, world, tile_rows, with_u):
    """F_SHARD | F_MAXN: every slab streams its own rows of x and u; sums are added, maxima maximised over the ranks."""
    worst, dx = model.run_sharded_maxn(n, m, lshape, iters, world, tile_rows=tile_rows, with_u=with_u)
    assert worst < 1e-12 and dx < 1e-12
