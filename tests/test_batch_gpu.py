"""b200cg_solve_batch (a queue of right-hand sides whose host copies overlap the iterations) on the GPU: same results as
one b200cg_solve call per right-hand side (MatrixFreeSolver::solve, matrix_free_system.cpp:383-482), checked against the
CPU oracle as well; completion callbacks, buffer reuse, interrupt, argument errors."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from iterative_solvers_b200 import capi as c

    c.lib()
    assert c.device_count() >= 1, "these tests need a CUDA device"
    return c


@pytest.fixture()
def fixed_split():
    """No feedback balancing: the work split, hence the summation order, is the same in every solve of a plan."""
    saved = os.environ.get("B200CG_BALANCE")
    os.environ["B200CG_BALANCE"] = "0"
    yield
    if saved is None:
        os.environ.pop("B200CG_BALANCE", None)
    else:
        os.environ["B200CG_BALANCE"] = saved


def relmax(x, ref):
    return np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300)


@pytest.mark.parametrize("n,small_grid_path", [(128, 0), (128, 1), (400, 0)])
def test_batch_equals_separate_solves_and_the_oracle(capi, oracle_mod, fixed_split, n, small_grid_path):
    """Five different right-hand sides (pinned buffers) through the queue: bit-equal to five separate calls on the same plan,
    and the reference's iteration count / solution for each (oracle)."""
    rng = np.random.default_rng(5)
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0)
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        p.build_rhs()
        b0 = p.get_rhs()
        scales = [1.0, -0.5, 2.0, 0.25, 3.0]
        pins_b = [capi.PinnedArray(p.N) for _ in scales]
        pins_x = [capi.PinnedArray(p.N) for _ in scales]
        for k, (pb, sc) in enumerate(zip(pins_b, scales)):
            pb.array[:] = sc * b0 + (0.0 if k == 0 else 1e-3 * np.max(np.abs(b0))) * rng.standard_normal(p.N)
        kw = dict(eps_rel=1e-8, max_it=5000, small_grid_path=small_grid_path)
        singles = [p.solve(b=pb.array.copy(), **kw) for pb in pins_b]
        order = []
        infos = p.solve_batch([pb.array for pb in pins_b], [px.array for px in pins_x],
                              done=lambda i, info: order.append((i, info["iterations"])), **kw)
        assert [i for i, _ in order] == list(range(len(scales)))
        for k, (info, (xs, si)) in enumerate(zip(infos, singles)):
            assert info["iterations"] == si["iterations"] == order[k][1] and info["converged"]
            assert info["h2d_bytes"] == info["d2h_bytes"] == 8 * p.N
            assert np.array_equal(pins_x[k].array, xs), k
            assert info["r_l2"] == si["r_l2"]
        for k in (0, 3):  # the reference's answer (k = 0: the reference's own rhs; k = 3: a perturbed one)
            ref = o.mf_solve(b=pins_b[k].array.copy(), eps=1e-8, max_it=5000)
            assert abs(infos[k]["iterations"] - ref["iterations"]) <= (0 if k == 0 else 1)
            if infos[k]["iterations"] == ref["iterations"]:
                assert relmax(pins_x[k].array, ref["x"]) < 1e-10
        assert infos[0]["cluster_path"] == (1 if (n == 128 and small_grid_path == 0) else 0)
        for h in pins_b + pins_x:
            h.free()


def test_batch_reuses_two_output_buffers(capi, fixed_split):
    """The queue of bench.py: one right-hand side buffer, two alternating solution buffers; each solution is read inside its
    completion callback, before the buffer is written again two solves later. Pageable buffers work too."""
    n = 1024
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        p.build_rhs()
        hb = capi.PinnedArray(p.N)
        hx = [capi.PinnedArray(p.N), capi.PinnedArray(p.N)]
        hb.array[:] = p.get_rhs()
        kw = dict(eps_rel=0.0, max_it=60)
        ref, rinfo = p.solve(b=hb.array, **kw)
        seen = []
        outs = [hx[i & 1].array for i in range(6)]
        infos = p.solve_batch([hb.array] * 6, outs, done=lambda i, info: seen.append(np.array_equal(outs[i], ref)), **kw)
        assert seen == [True] * 6
        assert all(i["iterations"] == 60 and i["single_sweep"] == rinfo["single_sweep"] == 1 for i in infos)
        assert p.solve_batch([], []) == []
        # pageable host memory: the copies are then staged by the driver, the results are the same
        xs = [np.empty(p.N) for _ in range(3)]
        p.solve_batch([hb.array.copy() for _ in range(3)], xs, **kw)
        assert all(np.array_equal(x, ref) for x in xs)
        # the plan still serves the one-call path afterwards (staging buffers back in their roles)
        again, _ = p.solve(b=hb.array, **kw)
        assert np.array_equal(again, ref)
        res, _ = p.postprocess(want_error=False)
        assert np.all(np.isfinite(res))
        hb.free()
        for h in hx:
            h.free()


def test_batch_interrupt_and_argument_errors(capi):
    n = 512
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        p.build_rhs()
        b = p.get_rhs()
        xs = [np.empty(p.N) for _ in range(3)]
        flag = ctypes.c_int(1)  # requestStop before the first solve: it is interrupted, the rest never start
        done = []
        infos = p.solve_batch([b] * 3, xs, done=lambda i, info: done.append(i), eps_rel=1e-30, max_it=100000,
                              stop_flag=flag)
        assert infos[0]["stop_reason"] == "INTERRUPTED" and infos[0]["iterations"] <= 16
        assert [i["stop_reason"] for i in infos[1:]] == ["INTERRUPTED"] * 2
        assert [i["iterations"] for i in infos[1:]] == [0, 0] and done == [0]
        with pytest.raises(capi.B200CGError) as e:
            p.solve_batch([b], [xs[0]], op=capi.OP_CSR)
        assert e.value.status == 6  # B200CG_ERR_UNSUPPORTED
        prm = capi.Params(op=capi.OP_MATRIX_FREE, rule=capi.RULE_REL_L2, eps_rel=1e-8, max_it=10, rhs_on_device=1)
        bp = (ctypes.c_void_p * 1)(b.ctypes.data)
        xp = (ctypes.c_void_p * 1)(xs[0].ctypes.data)
        info = (capi.SolveInfo * 1)()
        assert p.L.b200cg_solve_batch(p.h, ctypes.byref(prm), 1, bp, xp, info, None, None, None) == 1  # INVALID_ARG
        bp[0] = None
        prm.rhs_on_device = 0
        assert p.L.b200cg_solve_batch(p.h, ctypes.byref(prm), 1, bp, xp, info, None, None, None) == 1
        # and the plan is still usable
        x, i2 = p.solve(b=b, eps_rel=1e-8, max_it=20000)
        assert i2["converged"]
