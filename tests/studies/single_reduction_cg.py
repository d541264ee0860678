"""Numerical study (CPU, numpy): does a single-reduction CG (Chronopoulos-Gear recurrence for alpha) stay within the
parity bar (iterations +-1, x within 1e-10 relative) of the reference's CG on the reference grids?"""
import sys, time
import numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as om

def make_apply(o, n, m):
    hx, hy = 1.0 / n, 1.0 / m
    xk, yk = 1 / (hx * hx), 1 / (hy * hy)
    A = -2 * (xk + yk)
    mask = np.zeros((m + 1, n + 1), dtype=bool)
    mask[1:m, 1:n] = True
    mask[1:m // 2 + 1, 1:n // 2 + 1] = False
    idx = np.argwhere(mask)  # row-major (y, x): need the reference's compact order
    return A, xk, yk, mask

def to_grid(o, v, mask):
    # compact order: block B rows y<=m/2 (x>n/2) then block U rows -> row-major over mask rows works since each row is contiguous
    g = np.zeros(mask.shape)
    g[mask] = v  # row-major over (y, x) equals the reference order: rows ascending, x ascending within a row
    return g

def apply_grid(p, A, xk, yk, mask):
    out = np.zeros_like(p)
    c = p[1:-1, 1:-1]
    t = A * c
    t = t + xk * p[1:-1, :-2]
    t = t + xk * p[1:-1, 2:]
    t = t + yk * p[2:, 1:-1]
    t = t + yk * p[:-2, 1:-1]
    out[1:-1, 1:-1] = t
    out[~mask] = 0.0
    return out

def cg_ref(b, apply, eps, max_it):
    x = np.zeros_like(b); r = b.copy(); p = r.copy()
    rr = float(np.sum(r * r)); r0 = np.sqrt(rr); it = 0
    while it < max_it:
        Ap = apply(p)
        alpha = rr / float(np.sum(p * Ap))
        x += alpha * p; r -= alpha * Ap
        rr_new = float(np.sum(r * r)); it += 1
        if np.sqrt(rr_new) <= eps * r0: break
        beta = rr_new / rr; rr = rr_new
        p = r + beta * p
    return x, it

def cg_single(b, apply, eps, max_it):
    """Chronopoulos-Gear: one reduction (gamma = r.r, delta = r.Ar) per iteration; Ap recomputed from p."""
    x = np.zeros_like(b); r = b.copy(); p = np.zeros_like(b)
    w = apply(r)
    gamma = float(np.sum(r * r)); delta = float(np.sum(r * w)); r0 = np.sqrt(gamma)
    alpha = gamma / delta; beta = 0.0; it = 0
    while it < max_it:
        p = r + beta * p
        Ap = apply(p)
        x += alpha * p
        r = r - alpha * Ap
        w = apply(r)
        gamma_new = float(np.sum(r * r)); delta = float(np.sum(r * w)); it += 1
        if np.sqrt(gamma_new) <= eps * r0: break
        beta = gamma_new / gamma
        alpha = gamma_new / (delta - beta * gamma_new / alpha)
        gamma = gamma_new
    return x, it

def compare(n, eps):
    """(oracle iterations, single-reduction iterations, relative max difference of x) on the n x n L-shaped grid."""
    o = om.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, om.LSHAPE)
    A, xk, yk, mask = make_apply(o, n, n)
    b = to_grid(o, o.rhs(), mask)
    ap = lambda p: apply_grid(p, A, xk, yk, mask)
    ref = o.mf_solve(b=o.rhs(), eps=eps, max_it=50000)
    xr, itr = cg_ref(b, ap, eps, 50000)
    xs, its = cg_single(b, ap, eps, 50000)
    xo = to_grid(o, ref["x"], mask)
    sc = np.max(np.abs(xo))
    return ref["iterations"], itr, its, np.max(np.abs(xr - xo)) / sc, np.max(np.abs(xs - xo)) / sc


if __name__ == "__main__":
    for n, eps in [(64, 1e-8), (128, 1e-8), (256, 1e-8), (512, 1e-8), (1100, 1e-6), (1024, 1e-8)]:
        t = time.time()
        it_o, it_r, it_s, d_r, d_s = compare(n, eps)
        print(f"n={n}: oracle it {it_o}, numpy-CG it {it_r} (diff {d_r:.2e}), "
              f"single-reduction it {it_s} (diff vs oracle {d_s:.2e}) [{time.time()-t:.1f}s]", flush=True)
