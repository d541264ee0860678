"""Launched under torchrun (one process per GPU) by tests/test_multigpu.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multigpu.py
Row-slab sharded solve through the C ABI on every rank; rank 0 assembles the slabs and checks them against the
CPU oracle (same grid, rhs, x0, tolerance): iterations +-1, solution within 1e-10 relative."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterative_solvers_b200 import capi  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_comm_id():  # an NCCL unique id bootstraps exactly one communicator: one per plan
        blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            blob = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(blob, src=0)
        return bytes(blob.cpu().tolist())

    failures = []
    # "mf-hs2": every sweep flavour with 2-row stages (the launch shapes are read when the plan is created);
    # "msg0": the max-norm rules without a true solution
    hs2 = {"B200CG_SHAPE_DOT": "0", "B200CG_SHAPE_UPD": "0", "B200CG_SHAPE_NOX": "0"}
    cases = [(64, 0, 1e-8, "mf"), (256, 0, 1e-8, "mf"), (1100, 0, 1e-6, "mf"), (333, 1, 1e-8, "mf"),
             (256, 0, 1e-8, "mf-hs2"), (128, 0, 1e-8, "msg"), (128, 0, 1e-8, "msg0"), (128, 0, 1e-8, "cb")]
    # "mf": the default iteration = single sweep on a sharded plan (two halo rows per side over peer memory, one
    # publish-and-wait per iteration); "mf-2s": the two-sweep iteration
    cases += [(256, 0, 1e-8, "mf-2s"), (333, 1, 1e-8, "mf-2s"), (600, 0, 1e-6, "mf-2s")]
    for n, domain, eps, kind in cases:
        o = Oracle(n, n, 0.0, 1.0, 0.0, 1.0, domain)
        b, u = o.rhs(), o.true_solution()
        if kind == "mf-hs2":
            os.environ.update(hs2)
        plan = capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, domain=domain, device=local, rank=rank, world=world,
                         comm_id=fresh_comm_id())
        for k in list(hs2):
            os.environ.pop(k, None)
        lo, hi = plan.lo, plan.hi
        got_cb = []
        if kind in ("mf", "mf-hs2", "mf-2s"):
            ref = o.mf_solve(b=b, eps=eps, max_it=20000)
            two = kind != "mf"
            x, info = plan.solve(b=b[lo:hi], eps_rel=eps, max_it=20000, single_sweep=2 if two else 0)
            if info["single_sweep"] != (0 if two else 1):
                failures.append((n, domain, kind, "wrong iteration scheme"))
        elif kind == "cb":
            ref = o.mf_solve(b=b, eps=eps, max_it=20000, with_hist=True)
            x, info = plan.solve(b=b[lo:hi], u=u[lo:hi], eps_rel=eps, max_it=20000,
                                 callback=lambda it, p, r, e: got_cb.append((it, p, r, e)))
        elif kind == "msg0":
            ref = o.msg_solve(b=b, eps_p=eps, eps_r=eps, max_it=20000)
            x, info = plan.solve(b=b[lo:hi], rule=capi.RULE_MAXNORM, eps_p=eps, eps_r=eps, max_it=20000)
        else:
            ref = o.msg_solve(b=b, u=u, eps_p=eps, eps_r=eps, max_it=20000)
            x, info = plan.solve(b=b[lo:hi], u=u[lo:hi], rule=capi.RULE_MAXNORM, eps_p=eps, eps_r=eps, max_it=20000)
        v = np.random.default_rng(n).standard_normal(o.N)
        y = plan.apply(v[lo:hi])
        res, _ = plan.postprocess(want_error=False)
        plan.close()
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, x, y, res))
        if rank == 0:
            xg, yg, rg = np.empty(o.N), np.empty(o.N), np.empty(o.N)
            for plo, phi, px, py, pr in parts:
                xg[plo:phi], yg[plo:phi], rg[plo:phi] = px, py, pr
            rel = np.max(np.abs(xg - ref["x"])) / np.max(np.abs(ref["x"]))
            ok = abs(info["iterations"] - ref["iterations"]) <= 1 and rel < 1e-10
            ok = ok and np.array_equal(yg, o.apply(v))
            ok = ok and np.max(np.abs(rg - (o.apply(xg) - b))) <= 1e-12 * np.max(np.abs(b))
            if kind in ("msg", "msg0"):
                ok = ok and info["stop_reason"] == ref["stop_reason"]
            if kind == "cb":
                got = np.array(got_cb)
                ok = ok and len(got) == len(ref["hist"]) and np.all(
                    np.abs(got[:, 1:] - ref["hist"]) <= 1e-9 * np.max(np.abs(ref["hist"]), axis=0) + 1e-9 * np.abs(ref["hist"]))
            print(f"[multigpu] n={n} domain={domain} {kind} (peer_exchange={info['peer_exchange']}): "
                  f"iterations {info['iterations']} vs {ref['iterations']}, "
                  f"max rel diff {rel:.2e} -> {'ok' if ok else 'FAIL'}", flush=True)
            if not ok:
                failures.append((n, domain, kind))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU_OK" if not failures else f"MULTIGPU_FAIL {failures}", flush=True)
        sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
