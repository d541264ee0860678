#!/usr/bin/env python
"""Where the per-iteration cross-GPU cost of the sharded single-sweep iteration goes: a device global-timer trace of
publish -> all flags seen -> scalars formed on every rank (b200cg_peer_trace), split into
  sweep    : from the previous iteration's scalars to this rank's sums being ready (the kernel itself + launch gap)
  publish  : remote stores of the sums, system-scope fence, remote flag stores
  skew     : waiting for the LAST rank to publish (load imbalance / launch skew between GPUs - not a latency)
  latency  : from the last publication to this rank seeing every flag (NVLink flag round trip + polling)
  finalize : summing the slots and forming alpha / beta
The GPUs' timers are not synchronised; their offsets are bounded from the stamps themselves (rank r sees rank 0's flag
after it was published and vice versa) and the midpoint is used; the half-width is reported as the uncertainty.
Run: B200CG_PEER_TRACE=1 python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/peer_trace.py
     [--grid-n 16384] [--iters 2000]   (strong scaling: the n x n grid sharded over all ranks). Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200CG_PEER_TRACE"] = "1"
from iterative_solvers_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid-n", dest="n", type=int, default=16384)
    ap.add_argument("--iters", type=int, default=2000)
    args = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        blob = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(blob, src=0)
    plan = capi.Plan(args.n, args.n, 0.0, 1.0, 0.0, 1.0, device=local, rank=rank, world=world,
                     comm_id=bytes(blob.cpu().tolist()))
    plan.build_rhs()
    kw = dict(rhs_on_device=True, keep_x_on_device=True, eps_rel=0.0, iters_per_graph=100)
    for _ in range(2):
        plan.solve(max_it=500, **kw)  # graphs, work-split feedback
    iters = min(args.iters, 4000)
    _, info = plan.solve(max_it=iters, **kw)
    assert info["single_sweep"] == 1 and info["peer_exchange"] == 1, info
    tr = plan.peer_trace()[:iters]
    parts = [None] * world
    dist.all_gather_object(parts, tr)
    if rank == 0:
        T = [p.astype(np.float64) for p in parts]
        off, unc = [0.0], [0.0]
        for r in range(1, world):
            lo = np.max(T[r][:, 1] - T[0][:, 2])
            hi = np.min(T[r][:, 2] - T[0][:, 1])
            off.append(0.5 * (lo + hi))
            unc.append(0.5 * (hi - lo))
        C = [T[r] - off[r] for r in range(world)]  # every rank's stamps in rank 0's clock
        last_pub = np.max(np.stack([c[:, 1] for c in C]), axis=0)
        first_ready = np.min(np.stack([c[:, 0] for c in C]), axis=0)
        last_ready = np.max(np.stack([c[:, 0] for c in C]), axis=0)
        rows = []
        for r in range(world):
            c = C[r]
            rows.append({"rank": r,
                         "sweep_us": float(np.mean(c[1:, 0] - c[:-1, 3])) * 1e-3,
                         "publish_us": float(np.mean(c[:, 1] - c[:, 0])) * 1e-3,
                         "skew_wait_us": float(np.mean(last_pub - c[:, 1])) * 1e-3,
                         "flag_latency_us": float(np.mean(c[:, 2] - last_pub)) * 1e-3,
                         "finalize_us": float(np.mean(c[:, 3] - c[:, 2])) * 1e-3,
                         "clock_offset_us": off[r] * 1e-3, "offset_uncertainty_us": unc[r] * 1e-3})
        period = float(np.mean(C[0][1:, 3] - C[0][:-1, 3])) * 1e-3
        mean = {k: float(np.mean([row[k] for row in rows])) for k in rows[0] if k.endswith("_us") and "offset" not in k}
        print(json.dumps({"grid_n": args.n, "gpus": world, "iterations_traced": iters, "unknowns_per_gpu": plan.N / world,
                          "iteration_period_us": period, "mean_over_ranks": mean,
                          "ready_spread_us": float(np.mean(last_ready - first_ready)) * 1e-3,
                          "solve_ms_per_iteration": info["solve_ms"] / info["iterations"], "per_rank": rows}), flush=True)
    plan.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
