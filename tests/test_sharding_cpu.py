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
