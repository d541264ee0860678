"""world_size-2 gloo test (CPU) of the N > 1 host-side logic: the row-slab partition the C ABI reports, the
128-byte bootstrap blob broadcast, and the one-row halo protocol - each rank applies the operator on its slab
with the oracle as the compute stand-in and the assembled result must equal the unsharded apply."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, domain, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from iterative_solvers_b200 import capi
    from oracle.oracle import Oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # bootstrap blob: rank 0 -> everyone (bench.py ships b200cg_comm_unique_id() the same way)
        blob = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            blob = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(blob, src=0)
        assert blob.tolist() == list(range(128))

        ylo, yhi, lo, hi, N = capi.partition(n, n, domain, rank, world)
        o = Oracle(n, n, 0.0, 1.0, 0.0, 1.0, domain)
        assert N == o.N
        ranges = [None] * world
        dist.all_gather_object(ranges, (ylo, yhi, lo, hi))
        for a, b in zip(ranges, ranges[1:]):
            assert a[1] == b[0] and a[3] == b[2]

        # global vector, deterministic on every rank; this rank keeps only its slab + halos
        xg = np.random.default_rng(7).standard_normal(N)
        owned = np.zeros(N)
        owned[lo:hi] = xg[lo:hi]

        def row_range(y):  # compact range of grid row y
            first = o.index(1 if (domain == 1 or y > n // 2) else n // 2 + 1, y)
            return first, o.index(n - 1, y) + 1

        # halo exchange: first/last owned rows go to the neighbours (comm_halo in csrc/comm.cu)
        reqs = []
        if rank > 0:
            a, b = row_range(ylo)
            reqs.append(dist.isend(torch.from_numpy(owned[a:b].copy()), rank - 1))
            ha, hb = row_range(ylo - 1)
            below = torch.zeros(hb - ha, dtype=torch.float64)
            reqs.append(dist.irecv(below, rank - 1))
        if rank < world - 1:
            a, b = row_range(yhi - 1)
            reqs.append(dist.isend(torch.from_numpy(owned[a:b].copy()), rank + 1))
            ha2, hb2 = row_range(yhi)
            above = torch.zeros(hb2 - ha2, dtype=torch.float64)
            reqs.append(dist.irecv(above, rank + 1))
        for r in reqs:
            r.wait()
        if rank > 0:
            owned[ha:hb] = below.numpy()
        if rank < world - 1:
            owned[ha2:hb2] = above.numpy()

        y_local = o.apply(owned)[lo:hi]  # rows of the slab only need the slab and its two halo rows
        y_ref = o.apply(xg)[lo:hi]
        assert np.array_equal(y_local, y_ref)

        # the two scalar all-reduces of an iteration: sum of partial dots equals the global dot to rounding
        part = torch.tensor([float(np.dot(xg[lo:hi], y_ref))], dtype=torch.float64)
        dist.all_reduce(part)
        full = float(np.dot(xg, o.apply(xg)))
        assert abs(part.item() - full) <= 1e-12 * abs(full)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,domain", [(64, 0), (51, 1)])
def test_two_rank_slab_protocol(tmp_path, n, domain):
    import torch.multiprocessing as mp

    from iterative_solvers_b200 import build

    build.build_library()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, domain, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


# ---------------------------------------------------------------- the sharded single-sweep iteration (F_SHARD), 2 processes
def _single_sweep_worker(rank, world, port, n, iters, out_dir, maxn=False):
    """Each process owns one slab of the lane-level kernel model (scripts/model_single_sweep.py) and runs the single-sweep
    iteration on it; what the CUDA kernel stores into its neighbours over NVLink travels here as gloo messages, the
    PeerSync sums as an all-reduce. The assembled iterate must equal a global single-reduction CG. maxn: the max-norm
    flavour (MSGSolver's rules) - x every iteration, the three maxima cross the ranks as a MAX all-reduce, and every rank
    must reach the same stop verdict at the same iteration."""
    sys.path.insert(0, ROOT)
    import importlib.util

    import torch
    import torch.distributed as dist

    from iterative_solvers_b200 import capi

    spec = importlib.util.spec_from_file_location("model_single_sweep", os.path.join(ROOT, "scripts", "model_single_sweep.py"))
    model = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(model)

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        G = model.Grid(n, n, True)
        rng = np.random.default_rng(n)
        b = np.where(G.mask, rng.standard_normal(G.mask.shape), 0.0)
        bp = G.to_pitched(b)
        ylo, yhi, _lo, _hi, _N = capi.partition(n, n, capi.DOMAIN_LSHAPE, rank, world)
        sl = model.Slab(G, ylo, yhi, rank > 0, rank + 1 < world)
        for y in range(ylo - 2, yhi + 2):  # r0 = b with both halo rows: what the init exchanges deliver
            ri = sl.row_index(y)
            if ri >= 0:
                sl.r[0][ri] = bp[y]
        tiles = capi.work_split(n, n, domain=capi.DOMAIN_LSHAPE, rank=rank, world=world, sms=4, ctas_per_sm=2, fused=True)[0]

        def exchange(buf):
            """rows ylo, ylo+1 -> the rank below (its halo row yhi, its extra row yhi+1); rows yhi-1, yhi-2 -> the rank
            above (its halo row ylo-1, its extra row ylo-2): fused_kernel.cuh, F_SHARD stores."""
            reqs, inbox = [], []
            for arr in (sl.r[buf], sl.p[buf]):
                if rank > 0:
                    reqs.append(dist.isend(torch.from_numpy(arr[[1, 2]].copy()), rank - 1))
                    t = torch.zeros(2, G.pitch, dtype=torch.float64)
                    reqs.append(dist.irecv(t, rank - 1))
                    inbox.append((arr, [0, sl.yrows], t))          # their yhi-1 -> our row ylo-1; their yhi-2 -> extra row 0
                if rank + 1 < world:
                    reqs.append(dist.isend(torch.from_numpy(arr[[sl.yrows - 2, sl.yrows - 3]].copy()), rank + 1))
                    t = torch.zeros(2, G.pitch, dtype=torch.float64)
                    reqs.append(dist.irecv(t, rank + 1))
                    inbox.append((arr, [sl.yrows - 1, sl.yrows + 1], t))  # their ylo -> our row yhi; their ylo+1 -> extra row 1
            for q in reqs:
                q.wait()
            for arr, rows, t in inbox:
                arr[rows] = t.numpy()

        gamma = float(np.sum(b * b))
        alpha = gamma / float(np.sum(b * G.apply(b)))
        beta = alpha_prev = 0.0
        ut = np.where(G.mask, np.random.default_rng(n + 1).standard_normal(G.mask.shape), 0.0)
        u_slab = np.zeros_like(sl.x)
        for y in range(ylo - 1, yhi + 1):
            u_slab[sl.row_index(y)] = G.to_pitched(ut)[y]
        norms, stop_at = [], None
        for k in range(iters):
            par = k & 1
            out = model.sweep(G, tiles, sl.r[par], sl.p[par], sl.x, sl.r[par ^ 1], sl.p[par ^ 1], alpha, beta,
                              0.0 if maxn else alpha_prev, x2=False if maxn else bool(k & 1), slab=sl, maxn=maxn,
                              u=u_slab if maxn else None)
            exchange(par ^ 1)
            sums = torch.tensor(out[:2], dtype=torch.float64)
            dist.all_reduce(sums)
            g2, d2 = float(sums[0]), float(sums[1])
            if maxn:
                mx = torch.tensor(out[2], dtype=torch.float64)
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                norms.append([float(v) for v in mx])
                if stop_at is None and float(mx[1]) < 0.05:  # a precision rule both ranks must trip in the same iteration
                    stop_at = k + 1
                beta = (np.sqrt(g2) * np.sqrt(g2)) / gamma
            else:
                alpha_prev = alpha if not (k & 1) else 0.0
                beta = g2 / gamma
            alpha, gamma = g2 / (d2 - beta * g2 / alpha), g2
        x_own = sl.x + (alpha_prev * sl.p[iters & 1] if (iters & 1 and not maxn) else 0.0)
        if maxn:
            seen = [None] * world
            dist.all_gather_object(seen, (stop_at, norms))
            assert all(v == seen[0] for v in seen)  # identical maxima, hence identical verdicts, on every rank
        parts = [None] * world
        dist.all_gather_object(parts, (ylo, yhi, x_own[1:1 + yhi - ylo]))
        if rank == 0:
            xg = np.zeros((n + 1, G.pitch))
            for a, c, rows in parts:
                xg[a:c] = rows
            # global single-reduction CG
            r = b.copy(); p = np.zeros_like(b); xs = np.zeros_like(b)
            ga = float(np.sum(r * r)); al = ga / float(np.sum(r * G.apply(r))); be = 0.0
            for _ in range(iters):
                p = r + be * p
                xs = xs + al * p
                r = r - al * G.apply(p)
                g2 = float(np.sum(r * r)); d2 = float(np.sum(r * G.apply(r)))
                be = g2 / ga
                al, ga = g2 / (d2 - be * g2 / al), g2
            assert np.max(np.abs(G.from_pitched(xg) - xs)) <= 1e-12 * np.max(np.abs(xs))
            if maxn:  # the maxima of the last iteration against the global iterate
                assert abs(norms[-1][0] - np.max(np.abs(r))) <= 1e-12 * np.max(np.abs(r))
                assert abs(norms[-1][2] - np.max(np.abs(xs - ut))) <= 1e-12 * np.max(np.abs(xs - ut))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,iters", [(64, 5), (96, 4)])
def test_two_rank_single_sweep_protocol(tmp_path, n, iters):
    import torch.multiprocessing as mp

    from iterative_solvers_b200 import build

    build.build_library()
    port = _free_port()
    mp.spawn(_single_sweep_worker, args=(2, port, n, iters, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


@pytest.mark.parametrize("n,iters", [(64, 5)])
def test_two_rank_maxnorm_single_sweep_protocol(tmp_path, n, iters):
    """F_SHARD | F_MAXN with gloo standing in for the peer-memory slots: sums added, maxima maximised, same verdict everywhere."""
    import torch.multiprocessing as mp

    from iterative_solvers_b200 import build

    build.build_library()
    port = _free_port()
    mp.spawn(_single_sweep_worker, args=(2, port, n, iters, str(tmp_path), True), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
