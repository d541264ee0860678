"""TEST INFRASTRUCTURE ONLY: numpy restatement of the opt-in multigrid-preconditioned CG (b200cg_params.preconditioner = 1,
iterative_solvers_b200/csrc/mg.cu). PARITY UNPINNED: the reference has no preconditioner (solver/solver.hpp:17-66 is the
abstract base kept for one), so this oracle restates OUR algorithm operation by operation; what anchors it to the
reference is the operator (the 5-point stencil in the accumulation order of matrix_free_system.cpp:216-266, checked
against oracle.Oracle.apply in tests/test_mg_oracle.py), the stop rule of MatrixFreeSolver::solve
(matrix_free_system.cpp:409) and the solution, which must agree with the reference-order plain CG solve.

Algorithm (every elementwise operation separately rounded, in this order):
  levels   : n, m halved while both are even, >= 8, and (L-shape) the re-entrant lines x = n/2, y = m/2 stay on coarse
             grid lines (n/2, m/2 even); coarse operators by rediscretisation (h doubled, grid_system.cpp:314-318)
  smoother : damped Jacobi, omega = 0.8: x <- x + (omega / A_diag) * (b - A x); the first sweep from x = 0 is x = (omega/A_diag) b
  V-cycle  : 2 pre-sweeps, r = b - A x, full-weighting restriction, recursion, bilinear prolongation, 2 post-sweeps;
             coarsest level: 16 sweeps from zero. Symmetric, so it is a valid CG preconditioner.
  PCG      : x0 = 0, r0 = b, z = M r, p = z; alpha = r.z / p.Ap; x += alpha p; r -= alpha Ap; stop when
             ||r||_2 <= eps ||r0||_2 (or it = max_it); z = M r; beta = r'.z' / r.z; p = z + beta p.
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

OMEGA = 0.8
NU_PRE = NU_POST = 2
NU_COARSEST = 16


class Level:
    def __init__(self, n, m, a, b, c, d, lshape):
        self.n, self.m = n, m
        hx, hy = (b - a) / n, (d - c) / m  # grid_system.cpp:314-318
        self.A = -2 * (1 / (hx * hx) + 1 / (hy * hy))
        self.xk, self.yk = 1 / (hx * hx), 1 / (hy * hy)
        self.w = OMEGA / self.A
        self.xsplit, self.ysplit = (n // 2, m // 2) if lshape else (0, 0)
        self.mask = np.zeros((m + 1, n + 1), dtype=bool)
        self.mask[1:m, 1:n] = True
        if lshape:
            self.mask[1:m // 2 + 1, 1:n // 2 + 1] = False

    def can_coarsen(self):
        ok = self.n % 2 == 0 and self.m % 2 == 0 and self.n >= 8 and self.m >= 8
        if self.ysplit:
            ok = ok and self.xsplit % 2 == 0 and self.ysplit % 2 == 0
        return ok

    def apply(self, v):
        """A v on the node grid, accumulation order diag, left, right, top, bottom (matrix_free_system.cpp:216-266)."""
        out = np.zeros_like(v)
        t = self.A * v[1:-1, 1:-1]
        t = t + self.xk * v[1:-1, :-2]
        t = t + self.xk * v[1:-1, 2:]
        t = t + self.yk * v[2:, 1:-1]
        t = t + self.yk * v[:-2, 1:-1]
        out[1:-1, 1:-1] = t
        out[~self.mask] = 0.0
        return out

    def jacobi_first(self, b):
        return np.where(self.mask, self.w * b, 0.0)

    def jacobi(self, x, b):
        return np.where(self.mask, x + self.w * (b - self.apply(x)), 0.0)

    def residual(self, x, b):
        return np.where(self.mask, b - self.apply(x), 0.0)


def restrict(r, coarse):
    """Full weighting onto the coarse unknowns: (4 c + 2 (W + E + N + S) + (NW + NE + SW + SE)) / 16."""
    c = r[2:-1:2, 2:-1:2]
    s1 = ((r[2:-1:2, 1:-2:2] + r[2:-1:2, 3::2]) + r[3::2, 2:-1:2]) + r[1:-2:2, 2:-1:2]
    s2 = ((r[3::2, 1:-2:2] + r[3::2, 3::2]) + r[1:-2:2, 1:-2:2]) + r[1:-2:2, 3::2]
    out = np.zeros((coarse.m + 1, coarse.n + 1))
    out[1:-1, 1:-1] = 0.0625 * ((4.0 * c + 2.0 * s1) + s2)
    out[~coarse.mask] = 0.0
    return out


def prolong_add(x, e, fine):
    """x += bilinear interpolation of the coarse correction e (zero on the coarse boundary), on the fine unknowns."""
    p = np.zeros_like(x)
    p[::2, ::2] = e
    p[::2, 1::2] = 0.5 * (e[:, :-1] + e[:, 1:])
    p[1::2, ::2] = 0.5 * (e[:-1, :] + e[1:, :])
    p[1::2, 1::2] = 0.25 * (((e[:-1, :-1] + e[:-1, 1:]) + e[1:, :-1]) + e[1:, 1:])
    return np.where(fine.mask, x + p, 0.0)


class MgPcg:
    def __init__(self, m, n, a=0.0, b=1.0, c=0.0, d=1.0, lshape=True):
        self.levels = [Level(n, m, a, b, c, d, lshape)]
        while self.levels[-1].can_coarsen():
            L = self.levels[-1]
            self.levels.append(Level(L.n // 2, L.m // 2, a, b, c, d, lshape))

    def vcycle(self, l, b):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            x = L.jacobi_first(b)
            for _ in range(NU_COARSEST - 1):
                x = L.jacobi(x, b)
            return x
        x = L.jacobi_first(b)
        for _ in range(NU_PRE - 1):
            x = L.jacobi(x, b)
        r = L.residual(x, b)
        e = self.vcycle(l + 1, restrict(r, self.levels[l + 1]))
        x = prolong_add(x, e, L)
        for _ in range(NU_POST):
            x = L.jacobi(x, b)
        return x

    def solve(self, b_grid, eps=1e-8, max_it=1000):
        """b_grid: rhs on the node grid ((m+1) x (n+1), zero outside the unknowns). Returns dict(x, iterations, ...)."""
        L = self.levels[0]
        r = np.where(L.mask, b_grid, 0.0)
        x = np.zeros_like(r)
        r0 = float(np.sqrt(np.sum(r * r)))
        hist = []
        it, converged = 0, False
        if max_it > 0 and r0 > eps * r0:
            z = self.vcycle(0, r)
            p = z.copy()
            rz = float(np.sum(r * z))
            while True:
                Ap = L.apply(p)
                alpha = rz / float(np.sum(p * Ap))
                x = x + alpha * p
                r = r - alpha * Ap
                rn = float(np.sqrt(np.sum(r * r)))
                it += 1
                hist.append(rn)
                if not (it < max_it and rn > eps * r0):
                    converged = rn <= eps * r0
                    break
                z = self.vcycle(0, r)
                rz_new = float(np.sum(r * z))
                beta = rz_new / rz
                rz = rz_new
                p = z + beta * p
        else:
            converged = r0 <= eps * r0
        return dict(x=x, iterations=it, converged=converged, r0_norm=r0, r_norm=hist[-1] if hist else r0, hist=hist,
                    levels=len(self.levels))


def compact_nodes(n, m, lshape):
    """(ys, xs) of the unknowns in the reference's compact order (grid_system.cpp:84-111): block B (rows 1..m/2, columns
    n/2+1..n-1), then block U (rows m/2+1..m-1, columns 1..n-1); RECT: row-major."""
    if not lshape:
        ys, xs = np.meshgrid(np.arange(1, m), np.arange(1, n), indexing="ij")
        return ys.ravel(), xs.ravel()
    yb, xb = np.meshgrid(np.arange(1, m // 2 + 1), np.arange(n // 2 + 1, n), indexing="ij")
    yu, xu = np.meshgrid(np.arange(m // 2 + 1, m), np.arange(1, n), indexing="ij")
    return np.concatenate([yb.ravel(), yu.ravel()]), np.concatenate([xb.ravel(), xu.ravel()])


def to_grid(v, n, m, lshape):
    g = np.zeros((m + 1, n + 1))
    ys, xs = compact_nodes(n, m, lshape)
    g[ys, xs] = v
    return g


def from_grid(g, n, m, lshape):
    ys, xs = compact_nodes(n, m, lshape)
    return g[ys, xs].copy()
