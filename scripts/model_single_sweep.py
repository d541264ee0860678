"""CPU model of the single-sweep kernel's data flow (csrc/fused_kernel.cuh), lane by lane: the 484/480-column strips,
the 64-column warp windows with shuffle neighbours, the two-deep row pipeline, the per-row unknown masks and the
rows a tile streams (ya-2 .. yb+1). It runs whole CG iterations on small grids from the real tile tables
(b200cg_work_split, no GPU) and compares them with a plain numpy single-reduction CG. Columns the bulk copies would
leave stale in shared memory are NaN here, so a missing mask shows up at once.

This is a design check of index arithmetic, not a product path:  python scripts/model_single_sweep.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_solvers_b200 import capi  # noqa: E402

XOFF, STRIP_COLS, WARP_STEP, STRIP_LOAD = 4, 484, 60, 512


def shfl_up(v):
    out = v.copy()
    out[:, 1:] = v[:, :-1]
    return out


def shfl_down(v):
    out = v.copy()
    out[:, :-1] = v[:, 1:]
    return out


class Grid:
    def __init__(self, n, m, lshape=True):
        self.n, self.m = n, m
        self.xsplit, self.ysplit = (n // 2, m // 2) if lshape else (0, 0)
        self.pitch = (n + 1 + XOFF + 15) // 16 * 16
        self.ybase, self.yrows = 0, m + 1
        hx, hy = 1.0 / n, 1.0 / m
        self.xk, self.yk = 1 / (hx * hx), 1 / (hy * hy)
        self.A = -2 * (self.xk + self.yk)
        self.mask = np.zeros((m + 1, n + 1), dtype=bool)
        self.mask[1:m, 1:n] = True
        if lshape:
            self.mask[1:m // 2 + 1, 1:n // 2 + 1] = False

    def to_pitched(self, g):
        out = np.zeros((self.yrows, self.pitch))
        out[:, XOFF:XOFF + self.n + 1] = g
        return out

    def from_pitched(self, p):
        return p[:, XOFF:XOFF + self.n + 1].copy()

    def apply(self, p):
        out = np.zeros_like(p)
        t = self.A * p[1:-1, 1:-1]
        t = t + self.xk * p[1:-1, :-2]
        t = t + self.xk * p[1:-1, 2:]
        t = t + self.yk * p[2:, 1:-1]
        t = t + self.yk * p[:-2, 1:-1]
        out[1:-1, 1:-1] = t
        out[~self.mask] = 0.0
        return out


def sweep(G, tiles, r_in, p_in, x, r_out, p_out, alpha, beta, alpha_prev, x2, edge=False):
    """One launch of the kernel over all tiles; returns (gamma', delta'). edge: the F_EDGE variant (r'.A r' from
    edge sums, rows streamed from ya-1)."""
    warp = np.arange(8)[:, None]
    lane = np.arange(32)[None, :]
    sc = WARP_STEP * warp + 2 * lane
    writer = (lane >= 1) & (lane <= 30) & (warp >= 0)
    gam = dlt = dh = dv = 0.0

    def stencil(c, l, r, t, b):
        v = G.A * c
        v = v + G.xk * l
        v = v + G.xk * r
        v = v + G.yk * t
        v = v + G.yk * b
        return v

    for col0, ya, yb, _xlo in tiles:
        row_doubles = min(STRIP_COLS, G.pitch - col0)
        x0 = col0 + sc - XOFF
        z = np.zeros((8, 32))
        P1x, P1y, P2x, P2y, LP1, RP1 = z, z, z, z, z, z
        R1x, R1y, R2x, R2y, LR1, RR1 = z, z, z, z, z, z
        r1x, r1y, x1x, x1y, q1x, q1y = z, z, z, z, z, z
        k1a = k1b = np.zeros((8, 32), dtype=bool)
        for y in range(ya - (1 if edge else 2), yb + 2):
            stored = G.ybase <= y < G.ybase + G.yrows
            row_ok = 1 <= y <= G.m - 1
            xlo = G.xsplit + 1 if (G.ysplit != 0 and y <= G.ysplit) else 1
            k0a = row_ok & (x0 >= xlo) & (x0 <= G.n - 1)
            k0b = row_ok & (x0 + 1 >= xlo) & (x0 + 1 <= G.n - 1)
            if stored:
                def stage(src):
                    buf = np.full(STRIP_LOAD, np.nan)  # what the bulk copy does not write stays stale
                    buf[:row_doubles] = src[y - G.ybase, col0:col0 + row_doubles]
                    return buf
                sp, sr = stage(p_in), stage(r_in)
                cpx, cpy, crx, cry = sp[sc], sp[sc + 1], sr[sc], sr[sc + 1]
                if x2 and ya <= y < yb:
                    sx = stage(x)
                    cxx, cxy = sx[sc], sx[sc + 1]
                else:
                    cxx = cxy = z
            else:
                cpx = cpy = crx = cry = cxx = cxy = z
            cpx, cpy = np.where(k0a, cpx, 0.0), np.where(k0b, cpy, 0.0)
            crx, cry = np.where(k0a, crx, 0.0), np.where(k0b, cry, 0.0)
            cxx, cxy = np.where(k0a, cxx, 0.0), np.where(k0b, cxy, 0.0)
            P0x, P0y = crx + beta * cpx, cry + beta * cpy
            LP0, RP0 = shfl_up(P0y), shfl_down(P0x)
            ap0 = stencil(P1x, LP1, P1y, P0x, P2x)
            ap1 = stencil(P1y, P1x, RP1, P0y, P2y)
            R0x = np.where(k1a, r1x - alpha * ap0, 0.0)
            R0y = np.where(k1b, r1y - alpha * ap1, 0.0)
            if ya <= y - 1 < yb:
                st = writer & (k1a | k1b)
                cols = (col0 + sc)[st]
                yy = y - 1 - G.ybase
                r_out[yy, cols], r_out[yy, cols + 1] = R0x[st], R0y[st]
                p_out[yy, cols], p_out[yy, cols + 1] = P1x[st], P1y[st]
                if x2:
                    x[yy, cols] = ((x1x + alpha_prev * q1x) + alpha * P1x)[st]
                    x[yy, cols + 1] = ((x1y + alpha_prev * q1y) + alpha * P1y)[st]
                w = writer & np.ones((8, 32), dtype=bool)
                gam += float(np.sum(R0x[w] * R0x[w]) + np.sum(R0y[w] * R0y[w]))
            LR0, RR0 = shfl_up(R0y), shfl_down(R0x)
            w = writer & np.ones((8, 32), dtype=bool)
            if edge:
                if ya <= y - 1 < yb:
                    dh += float(np.sum(R0x[w] * R0y[w]) + np.sum(R0y[w] * RR0[w]))
                if ya <= y - 2 < yb:
                    dv += float(np.sum(R1x[w] * R0x[w]) + np.sum(R1y[w] * R0y[w]))
            elif ya <= y - 2 < yb:
                w0 = stencil(R1x, LR1, R1y, R0x, R2x)
                w1 = stencil(R1y, R1x, RR1, R0y, R2y)
                dlt += float(np.sum(R1x[w] * w0[w]) + np.sum(R1y[w] * w1[w]))
            P2x, P2y, P1x, P1y, LP1, RP1 = P1x, P1y, P0x, P0y, LP0, RP0
            R2x, R2y, R1x, R1y, LR1, RR1 = R1x, R1y, R0x, R0y, LR0, RR0
            r1x, r1y, x1x, x1y, q1x, q1y = crx, cry, cxx, cxy, cpx, cpy
            k1a, k1b = k0a, k0b
    if edge:
        dlt = G.A * gam + 2.0 * G.xk * dh + 2.0 * G.yk * dv
    return gam, dlt


def run(n, m, lshape, iters, tile_rows=0, sms=4, edge=False):
    G = Grid(n, m, lshape)
    domain = (capi.DOMAIN_LSHAPE if n == m and n % 2 == 0 else capi.DOMAIN_LSHAPE_ANY) if lshape else capi.DOMAIN_RECT
    tiles, _ = capi.work_split(m, n, domain=domain, sms=sms, ctas_per_sm=2, tile_rows=tile_rows, fused=True)
    rng = np.random.default_rng(n * 1000 + m)
    b = np.where(G.mask, rng.standard_normal(G.mask.shape), 0.0)

    # reference: plain single-reduction CG on the node grid
    r = b.copy(); p = np.zeros_like(b); xs = np.zeros_like(b)
    w = G.apply(r)
    gamma = float(np.sum(r * r)); alpha = gamma / float(np.sum(r * w)); beta = 0.0
    hist = []
    for _ in range(iters):
        p = r + beta * p
        xs = xs + alpha * p
        r = r - alpha * G.apply(p)
        g2 = float(np.sum(r * r)); d2 = float(np.sum(r * G.apply(r)))
        hist.append((g2, d2))
        beta = g2 / gamma
        alpha = g2 / (d2 - beta * g2 / alpha)
        gamma = g2

    # the kernel's data flow
    rb = [G.to_pitched(b), np.zeros((G.yrows, G.pitch))]
    pb = [np.zeros((G.yrows, G.pitch)), np.zeros((G.yrows, G.pitch))]
    x = np.zeros((G.yrows, G.pitch))
    gamma = float(np.sum(b * b)); alpha = gamma / float(np.sum(b * G.apply(b))); beta = 0.0; alpha_prev = 0.0
    worst = 0.0
    for k in range(iters):
        par = k & 1
        g2, d2 = sweep(G, tiles, rb[par], pb[par], x, rb[par ^ 1], pb[par ^ 1], alpha, beta, alpha_prev, x2=bool(k & 1),
                       edge=edge)
        worst = max(worst, abs(g2 - hist[k][0]) / hist[k][0], abs(d2 - hist[k][1]) / abs(hist[k][1]))
        alpha_prev = alpha if not (k & 1) else 0.0
        beta = g2 / gamma
        alpha, gamma = g2 / (d2 - beta * g2 / alpha), g2
    if iters & 1:  # pending x update of a final even pass (x_flush_kernel)
        x = x + alpha_prev * pb[iters & 1]
    xr = G.from_pitched(x)
    assert np.all(np.isfinite(xr))
    assert not np.any(xr[~G.mask]) and not np.any(G.from_pitched(rb[iters & 1])[~G.mask])
    dx = np.max(np.abs(xr - xs)) / np.max(np.abs(xs))
    dr = np.max(np.abs(G.from_pitched(rb[iters & 1]) - r)) / np.max(np.abs(r))
    return worst, dx, dr, len(tiles)


if __name__ == "__main__":
    for n, m, lshape, iters, tr in [(30, 30, True, 7, 0), (64, 64, True, 6, 0), (64, 64, True, 5, 5), (130, 90, True, 6, 0),
                                    (77, 33, False, 6, 3), (1000, 40, True, 4, 0), (970, 24, False, 3, 7)]:
        for edge in (False, True):
            worst, dx, dr, nt = run(n, m, lshape, iters, tr, edge=edge)
            print(f"n={n} m={m} {'L' if lshape else 'rect'} tile_rows={tr} tiles={nt} {'edge sums' if edge else 'stencil'}: "
                  f"dots {worst:.1e}, x {dx:.1e}, r {dr:.1e}")
            # the edge-sum form of r'.A r' cancels A_diag * gamma' against the edge terms: its dots carry ~1e-11
            assert worst < (1e-9 if edge else 1e-12) and dx < 1e-10 and dr < 1e-10
    print("MODEL_OK")
