#!/usr/bin/env python
"""MSGSolver's iteration (max-norm rules, msg_solver.cpp:80-183) on the matrix-free operator: single sweep (default) against
dot sweep + update sweep, with and without the true solution, device-timed through the C ABI. One JSON line per case.
  python scripts/maxnorm_bench.py [--grid-n 16384] [--iters 200] [--reps 3]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_solvers_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid-n", type=int, default=16384)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    n = args.grid_n
    peak = 6544.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        p.build_rhs()
        u = p.true_solution()
        for with_u in (True, False):
            for ss, name in ((0, "single sweep"), (2, "dot sweep + update sweep")):
                kw = dict(rhs_on_device=True, keep_x_on_device=True, u=u if with_u else None, rule=capi.RULE_MAXNORM,
                          eps_p=-1.0, eps_r=1e-300, max_it=args.iters, single_sweep=ss)
                p.solve(**kw)
                ms, its, info = 0.0, 0, None
                for _ in range(args.reps):
                    _, info = p.solve(**kw)
                    ms += info["solve_ms"]
                    its += info["iterations"]
                bytes_it = (48.0 if info["single_sweep"] else 64.0) + (8.0 if with_u else 0.0)
                value = p.N * its / (ms * 1e-3) / 1e9
                print(json.dumps({"workload": f"{n}x{n} L-shaped grid, MSGSolver rules, matrix-free", "iteration": name,
                                  "single_sweep": info["single_sweep"], "true_solution": with_u, "iterations": its,
                                  "gdof_it_per_s": value, "ms_per_iteration": ms / its,
                                  "algorithmic_bytes_per_dof_iter": bytes_it, "hbm_gbs": value * bytes_it,
                                  "frac_of_measured_peak": value * bytes_it / peak, "kernel_ms": [info["upd_even_ms"], info["upd_odd_ms"]],
                                  "r_max": info["r_max"], "dx_max": info["dx_max"], "err_max": info["err_max"]}), flush=True)


if __name__ == "__main__":
    main()
