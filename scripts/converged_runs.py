#!/usr/bin/env python
"""Converged solves on record (loop condition of MatrixFreeSolver::solve, matrix_free_system.cpp:409: relative
recurrence residual <= eps): iterations, seconds, final ||r||/||r0||, the TRUE residual ||b - A x||/||b|| recomputed by
b200cg_postprocess, max|x - u| against the analytic solution (expected O(h^2)), for the plain CG iteration(s) and the
opt-in multigrid-preconditioned one. One GPU:  python scripts/converged_runs.py --grid-n 4096 [--eps 1e-8]
N GPUs (row slabs): python -m torch.distributed.run --nproc-per-node N ... scripts/converged_runs.py --grid-n 16384
Prints one JSON line per solve."""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterative_solvers_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid-n", dest="n", type=int, default=4096)
    ap.add_argument("--eps", type=float, default=1e-8)
    ap.add_argument("--max-it", type=int, default=200000)
    ap.add_argument("--modes", default="single_sweep,two_sweep,multigrid")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    comm_id = None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            blob = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(blob, src=0)
        comm_id = bytes(blob.cpu().tolist())

    def allreduce(v, op):
        if world == 1:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    n = args.n
    plan = capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, device=local, rank=rank, world=world, comm_id=comm_id)
    plan.build_rhs()
    b = plan.get_rhs()
    u = plan.true_solution()
    b_norm = math.sqrt(allreduce(float(np.dot(b, b)), "SUM"))
    for mode in args.modes.split(","):
        if mode == "multigrid" and world > 1:
            continue  # the preconditioner is single-GPU
        kw = dict(rhs_on_device=True, eps_rel=args.eps, max_it=args.max_it, iters_per_graph=100)
        kw.update({"single_sweep": dict(single_sweep=1), "two_sweep": dict(single_sweep=2),
                   "multigrid": dict(preconditioner=capi.PRECOND_MULTIGRID)}[mode])
        plan.solve(**dict(kw, max_it=4))  # warm-up: graphs, work-split feedback, multigrid hierarchy
        x, info = plan.solve(**kw)
        res, _ = plan.postprocess(want_error=False)
        true_res = math.sqrt(allreduce(float(np.dot(res, res)), "SUM")) / b_norm
        err = allreduce(float(np.max(np.abs(x - u))), "MAX")
        if rank == 0:
            print(json.dumps({
                "grid_n": n, "unknowns": plan.N, "gpus": world, "iteration": mode, "eps_rel": args.eps,
                "iterations": info["iterations"], "converged": info["converged"], "stop_reason": info["stop_reason"],
                "solve_seconds": info["solve_ms"] * 1e-3, "ms_per_iteration": info["solve_ms"] / max(info["iterations"], 1),
                "final_rel_recurrence_residual": info["r_l2"] / info["r0_l2"], "true_rel_residual": true_res,
                "max_abs_error_vs_analytic": err, "h2": (1.0 / n) ** 2, "error_over_h2": err * n * n,
                "mg_levels": info["mg_levels"], "single_sweep": info["single_sweep"],
                "peer_exchange": info["peer_exchange"]}), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
