"""Numerical study (CPU, numpy): r.Ar formed from edge sums,  A_diag * sum r^2 + 2 xk * sum_h r_i r_j + 2 yk * sum_v r_i r_j,
against the stencil form sum r * (A r), along a CG run on the reference's L-shaped grid. The edge form subtracts two
numbers of size |A_diag| * r.r to get r.Ar; the question is how many digits of alpha survive as the grid grows
(|A_diag| = 4 n^2 on the unit square, r.Ar / r.r is the Rayleigh quotient of the residual).

    python tests/studies/edge_sum_delta.py [n ...]
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("single_reduction_cg", os.path.join(HERE, "single_reduction_cg.py"))
base = importlib.util.module_from_spec(spec)
spec.loader.exec_module(base)


def run(n, iters):
    o = base.om.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, base.om.LSHAPE)
    A, xk, yk, mask = base.make_apply(o, n, n)
    b = base.to_grid(o, o.rhs(), mask)
    ap = lambda p: base.apply_grid(p, A, xk, yk, mask)
    r = b.copy(); p = np.zeros_like(b)
    gamma = float(np.sum(r * r)); alpha = gamma / float(np.sum(r * ap(r))); beta = 0.0
    worst = 0.0
    for it in range(iters):
        p = r + beta * p
        r = r - alpha * ap(p)
        g2 = float(np.sum(r * r))
        d_st = float(np.sum(r * ap(r)))
        d_ed = A * g2 + 2 * xk * float(np.sum(r[:, :-1] * r[:, 1:])) + 2 * yk * float(np.sum(r[:-1, :] * r[1:, :]))
        # exact-ish reference in extended precision
        rl = r.astype(np.longdouble)
        d_ex = float(np.sum(rl * ap(rl)))
        worst = max(worst, abs(d_ed - d_ex) / abs(d_ex))
        if it in (0, iters // 2, iters - 1):
            print(f"  n={n} it={it + 1}: Rayleigh quotient / A_diag = {d_ex / g2 / A:.3e}, rel err stencil {abs(d_st - d_ex) / abs(d_ex):.1e}, "
                  f"edge {abs(d_ed - d_ex) / abs(d_ex):.1e}")
        beta = g2 / gamma
        alpha = g2 / (d_st - beta * g2 / alpha)
        gamma = g2
    return worst


if __name__ == "__main__":
    for n in [int(a) for a in sys.argv[1:]] or [128, 512, 1024]:
        print(f"n={n}: worst relative error of the edge form over the run: {run(n, 300):.1e}")
