#!/usr/bin/env python
"""Smallest program that launches the default single-sweep kernels at 16384^2 (no torch import): a target for
`ncu --set full -k regex:cg_fused_kernel -s 8 -c 2` - two 6-iteration solves; launches 9 and 10 are the even (x untouched) and
the odd (x touched) flavour of the second solve."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_solvers_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
    p.build_rhs()
    for _ in range(2):
        _, info = p.solve(rhs_on_device=True, keep_x_on_device=True, eps_rel=0.0, max_it=6, iters_per_graph=8)
    print(info["iterations"], info["single_sweep"], info["kernel_launches"])
