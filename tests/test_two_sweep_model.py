"""The thread-level CPU model of the default two-sweep kernel's data flow (scripts/model_two_sweep.py) on the real tile
tables: strips of 504 written columns with halo threads and the warp-edge neighbour, 2- and 4-row stages with the
unrolled FULL path, x-deferral, row slabs with the peer-memory halo stores. It pins the index arithmetic of
csrc/stream_kernel.cuh - e.g. that a FULL stage never holds the row a neighbour's halo needs. No GPU needed."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    spec = importlib.util.spec_from_file_location("model_two_sweep", os.path.join(ROOT, "scripts", "model_two_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("n,m,lshape,iters,world,hs,tile_rows,xdefer", [
    (30, 30, True, 5, 1, 4, 0, True), (64, 64, True, 4, 1, 2, 0, True), (64, 64, True, 5, 2, 2, 0, True),
    (64, 64, True, 4, 3, 4, 0, False), (77, 60, False, 4, 2, 4, 5, True), (70, 46, True, 3, 2, 4, 1, True),
    (1030, 24, True, 3, 2, 2, 0, True)])
def test_model_matches_reference_form_cg(model, n, m, lshape, iters, world, hs, tile_rows, xdefer):
    dx, dr = model.run(n, m, lshape, iters, world, hs, tile_rows, xdefer)
    assert dx < 1e-12 and dr < 1e-12
