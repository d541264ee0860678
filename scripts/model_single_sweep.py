"""CPU model of the single-sweep kernel's data flow (csrc/fused_kernel.cuh), lane by lane: the 424/420-column strips,
the 64-column warp windows with shuffle neighbours, the two-deep row pipeline, the per-row unknown masks (and their
absence on the inputs of FULL stages) and the rows a tile streams (ya-2 .. yb+1). It runs whole CG iterations on small grids from the real tile tables
(b200cg_work_split, no GPU) and compares them with a plain numpy single-reduction CG. Columns the bulk copies would
leave stale in shared memory are NaN here, so a missing mask shows up at once.

This is a design check of index arithmetic, not a product path:  python scripts/model_single_sweep.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_solvers_b200 import capi  # noqa: E402

XOFF, WARPS, WARP_STEP, STRIP_LOAD, HS = 4, 7, 60, 512, 4


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


class Slab:
    """One rank of a sharded plan: rows ylo-1 .. yhi stored at row index y - ybase, plus the two extra rows behind them
    (index yrows: row ylo-2, index yrows+1: row yhi+1) that the single-sweep kernel uses as second halo rows."""

    def __init__(self, G, ylo, yhi, below, above):
        self.ylo, self.yhi, self.ybase, self.yrows = ylo, yhi, ylo - 1, yhi - ylo + 2
        self.has_below, self.has_above = below, above
        shape = (self.yrows + 2, G.pitch)
        self.r = [np.zeros(shape), np.zeros(shape)]
        self.p = [np.zeros(shape), np.zeros(shape)]
        self.x = np.zeros(shape)

    def row_index(self, y):
        if self.ybase <= y < self.ybase + self.yrows:
            return y - self.ybase
        if y == self.ylo - 2 and self.has_below:
            return self.yrows
        if y == self.yhi + 1 and self.has_above:
            return self.yrows + 1
        return -1


def sweep(G, tiles, r_in, p_in, x, r_out, p_out, alpha, beta, alpha_prev, x2, slab=None, nb_below=None,
          nb_above=None, warps=WARPS, maxn=False, u=None):
    """One launch of the kernel over all tiles; returns (gamma', delta'). slab / nb_*: the F_SHARD variant - this
    rank's Slab and the (r_out, p_out) arrays of the neighbour ranks, whose halo rows receive this slab's two first /
    last rows. maxn: the F_MAXN flavour (MSGSolver's rules) - x += alpha * p every iteration and the maxima |r'|_inf,
    |x' - x|_inf, |x' - u|_inf (u: pitched true solution or None) come back as a third value; FULL stages read x and u
    unmasked too, so the maxima are taken under selects."""
    shard = nb_below is not None or nb_above is not None or slab is not None
    if slab is None:
        slab = Slab(G, 1, G.m, False, False)  # one rank owning every row: rows 0 .. m stored, no extra rows in use
    warp = np.arange(warps)[:, None]  # consumer warps
    lane = np.arange(32)[None, :]
    sc = WARP_STEP * warp + 2 * lane
    writer = (lane >= 1) & (lane <= 30) & (warp >= 0)
    strip_cols = WARP_STEP * warps + 4
    gam = dlt = 0.0
    mx = [0.0, 0.0, 0.0]  # |r'|_inf, |dx|_inf, |x' - u|_inf

    def stencil(c, l, r, t, b):
        v = G.A * c
        v = v + G.xk * l
        v = v + G.xk * r
        v = v + G.yk * t
        v = v + G.yk * b
        return v

    for col0, ya, yb, _xlo in tiles:
        row_doubles = min(strip_cols, G.pitch - col0)
        x0 = col0 + sc - XOFF
        z = np.zeros((warps, 32))
        P1x, P1y, P2x, P2y, LP1, RP1 = z, z, z, z, z, z
        R1x, R1y, R2x, R2y, LR1, RR1 = z, z, z, z, z, z
        r1x, r1y, x1x, x1y, q1x, q1y = z, z, z, z, z, z
        u1x, u1y = z, z
        k1a = k1b = np.zeros((warps, 32), dtype=bool)
        for y in range(ya - 2, yb + 2):
            # the stage this row arrives in (HS rows from ya-2 on) and whether the kernel takes its FULL path there
            y0 = ya - 2 + (y - (ya - 2)) // HS * HS
            full = y0 + HS <= yb + 2 and y0 >= ya + (3 if shard else 2) and y0 + HS <= yb - (1 if shard else 0)
            ri = slab.row_index(y)
            stored = ri >= 0
            row_ok = 1 <= y <= G.m - 1
            xlo = G.xsplit + 1 if (G.ysplit != 0 and y <= G.ysplit) else 1
            k0a = row_ok & (x0 >= xlo) & (x0 <= G.n - 1)
            k0b = row_ok & (x0 + 1 >= xlo) & (x0 + 1 <= G.n - 1)
            if stored:
                def stage(src):
                    buf = np.full(strip_cols, np.nan)  # what the bulk copy does not write stays stale
                    buf[:row_doubles] = src[ri, col0:col0 + row_doubles]
                    return buf
                sp, sr = stage(p_in), stage(r_in)
                cpx, cpy, crx, cry = sp[sc], sp[sc + 1], sr[sc], sr[sc + 1]
                if (x2 or maxn) and ya <= y < yb:
                    sx = stage(x)
                    cxx, cxy = sx[sc], sx[sc + 1]
                else:
                    cxx = cxy = z
                if maxn and u is not None and ya <= y < yb:
                    su = stage(u)
                    cux, cuy = su[sc], su[sc + 1]
                else:
                    cux = cuy = z
            else:
                cpx = cpy = crx = cry = cxx = cxy = cux = cuy = z
            if not full:  # FULL stages read their inputs unmasked (zeros outside the unknowns, stale data only where masked later)
                cpx, cpy = np.where(k0a, cpx, 0.0), np.where(k0b, cpy, 0.0)
                crx, cry = np.where(k0a, crx, 0.0), np.where(k0b, cry, 0.0)
                cxx, cxy = np.where(k0a, cxx, 0.0), np.where(k0b, cxy, 0.0)
                cux, cuy = np.where(k0a, cux, 0.0), np.where(k0b, cuy, 0.0)
            P0x, P0y = crx + beta * cpx, cry + beta * cpy
            LP0, RP0 = shfl_up(P0y), shfl_down(P0x)
            ap0 = stencil(P1x, LP1, P1y, P0x, P2x)
            ap1 = stencil(P1y, P1x, RP1, P0y, P2y)
            R0x = np.where(k1a, r1x - alpha * ap0, 0.0)
            R0y = np.where(k1b, r1y - alpha * ap1, 0.0)
            if ya <= y - 1 < yb:
                st = writer & (k1a | k1b)
                cols = (col0 + sc)[st]
                yy = y - 1 - slab.ybase
                r_out[yy, cols], r_out[yy, cols + 1] = R0x[st], R0y[st]
                p_out[yy, cols], p_out[yy, cols + 1] = P1x[st], P1y[st]
                ye = y - 1
                if nb_below is not None and ye in (slab.ylo, slab.ylo + 1):  # its top halo row / its second extra row
                    nb, rows = nb_below
                    row = rows - 1 if ye == slab.ylo else rows + 1
                    for arr, vx, vy in ((nb[0], R0x, R0y), (nb[1], P1x, P1y)):
                        arr[row, cols], arr[row, cols + 1] = vx[st], vy[st]
                if nb_above is not None and ye in (slab.yhi - 1, slab.yhi - 2):  # its bottom halo row / first extra row
                    nb, rows = nb_above
                    row = 0 if ye == slab.yhi - 1 else rows
                    for arr, vx, vy in ((nb[0], R0x, R0y), (nb[1], P1x, P1y)):
                        arr[row, cols], arr[row, cols + 1] = vx[st], vy[st]
                if x2:
                    x[yy, cols] = ((x1x + alpha_prev * q1x) + alpha * P1x)[st]
                    x[yy, cols + 1] = ((x1y + alpha_prev * q1y) + alpha * P1y)[st]
                w = writer & np.ones((warps, 32), dtype=bool)
                gam += float(np.sum(R0x[w] * R0x[w]) + np.sum(R0y[w] * R0y[w]))
                if maxn:
                    with np.errstate(invalid="ignore"):
                        xmx, xmy = x1x + alpha * P1x, x1y + alpha * P1y
                        x[yy, cols], x[yy, cols + 1] = xmx[st], xmy[st]
                        # the kernel's selects: writer lanes, unknown columns (generic rows: k1a / k1b; FULL rows: ma / mb,
                        # the same sets there); np.max would hand a NaN from a stale column through
                        d0 = np.where(writer & k1a, np.abs(xmx - x1x), 0.0)
                        d1 = np.where(writer & k1b, np.abs(xmy - x1y), 0.0)
                        mx[0] = max(mx[0], float(np.max(np.abs(R0x[w]))), float(np.max(np.abs(R0y[w]))))
                        mx[1] = max(mx[1], float(np.max(d0)), float(np.max(d1)))
                        if u is not None:
                            e0 = np.where(writer & k1a, np.abs(xmx - u1x), 0.0)
                            e1 = np.where(writer & k1b, np.abs(xmy - u1y), 0.0)
                            mx[2] = max(mx[2], float(np.max(e0)), float(np.max(e1)))
                    assert np.all(np.isfinite(mx))
            LR0, RR0 = shfl_up(R0y), shfl_down(R0x)
            w = writer & np.ones((warps, 32), dtype=bool)
            if ya <= y - 2 < yb:
                w0 = stencil(R1x, LR1, R1y, R0x, R2x)
                w1 = stencil(R1y, R1x, RR1, R0y, R2y)
                dlt += float(np.sum(R1x[w] * w0[w]) + np.sum(R1y[w] * w1[w]))
            P2x, P2y, P1x, P1y, LP1, RP1 = P1x, P1y, P0x, P0y, LP0, RP0
            R2x, R2y, R1x, R1y, LR1, RR1 = R1x, R1y, R0x, R0y, LR0, RR0
            r1x, r1y, x1x, x1y, q1x, q1y = crx, cry, cxx, cxy, cpx, cpy
            u1x, u1y = cux, cuy
            k1a, k1b = k0a, k0b
    if maxn:
        return gam, dlt, mx
    return gam, dlt


def run_maxn(n, m, lshape, iters, tile_rows=0, sms=4, warps=WARPS, with_u=True):
    """The F_MAXN flavour (MSGSolver's rules in one sweep): x every iteration and the three maxima, against a plain numpy
    statement of the same recurrence on the node grid (oracle/cg_oracle.c: cgo_msg_solve_single)."""
    G = Grid(n, m, lshape)
    domain = (capi.DOMAIN_LSHAPE if n == m and n % 2 == 0 else capi.DOMAIN_LSHAPE_ANY) if lshape else capi.DOMAIN_RECT
    tiles, _ = capi.work_split(m, n, domain=domain, sms=sms, ctas_per_sm=2 if warps == 7 else 1, tile_rows=tile_rows,
                               fused=2 if warps == 14 else 1)
    rng = np.random.default_rng(n * 1000 + m + 7)
    b = np.where(G.mask, rng.standard_normal(G.mask.shape), 0.0)
    ut = np.where(G.mask, rng.standard_normal(G.mask.shape), 0.0)

    r = b.copy(); z = np.zeros_like(b); xs = np.zeros_like(b)
    gamma = float(np.sum(r * r)); alpha = gamma / float(np.sum(r * G.apply(r))); beta = 0.0; rz = gamma
    hist = []
    for _ in range(iters):
        z = r + beta * z
        xn = xs + alpha * z
        dmax = float(np.max(np.abs(xn - xs))); xs = xn
        r = r - alpha * G.apply(z)
        g2 = float(np.sum(r * r)); d2 = float(np.sum(r * G.apply(r)))
        hist.append((g2, d2, float(np.max(np.abs(r))), dmax, float(np.max(np.abs(xs - ut)))))
        beta = (np.sqrt(g2) * np.sqrt(g2)) / rz
        alpha = g2 / (d2 - beta * g2 / alpha)
        rz = g2

    rb = [G.to_pitched(b), np.zeros((G.yrows, G.pitch))]
    pb = [np.zeros((G.yrows, G.pitch)), np.zeros((G.yrows, G.pitch))]
    x = np.zeros((G.yrows, G.pitch))
    up = G.to_pitched(ut) if with_u else None
    gamma = float(np.sum(b * b)); alpha = gamma / float(np.sum(b * G.apply(b))); beta = 0.0; rz = gamma
    worst = 0.0
    for k in range(iters):
        par = k & 1
        g2, d2, mx = sweep(G, tiles, rb[par], pb[par], x, rb[par ^ 1], pb[par ^ 1], alpha, beta, 0.0, x2=False,
                           warps=warps, maxn=True, u=up)
        ref = hist[k]
        worst = max(worst, abs(g2 - ref[0]) / ref[0], abs(d2 - ref[1]) / abs(ref[1]), abs(mx[0] - ref[2]) / ref[2],
                    abs(mx[1] - ref[3]) / ref[3], abs(mx[2] - ref[4]) / ref[4] if with_u else 0.0)
        beta = (np.sqrt(g2) * np.sqrt(g2)) / rz
        alpha = g2 / (d2 - beta * g2 / alpha)
        rz = g2
    xr = G.from_pitched(x)
    assert np.all(np.isfinite(xr)) and not np.any(xr[~G.mask])
    return worst, np.max(np.abs(xr - xs)) / np.max(np.abs(xs)), len(tiles)


def run(n, m, lshape, iters, tile_rows=0, sms=4, warps=WARPS):
    """warps: 7 (420-column strips, two CTAs per SM) or 14 (the wide geometry of large slabs: 840-column strips)."""
    G = Grid(n, m, lshape)
    domain = (capi.DOMAIN_LSHAPE if n == m and n % 2 == 0 else capi.DOMAIN_LSHAPE_ANY) if lshape else capi.DOMAIN_RECT
    tiles, _ = capi.work_split(m, n, domain=domain, sms=sms, ctas_per_sm=2 if warps == 7 else 1, tile_rows=tile_rows,
                               fused=2 if warps == 14 else 1)
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
                       warps=warps)
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


def run_sharded(n, m, lshape, iters, world, sms=4, tile_rows=0):
    """The F_SHARD data flow: `world` row slabs, each sweeping its own tiles, the two first / last rows of r' and p stored
    into the neighbours' halo row and extra row, the sums added over the ranks. Compared with one global CG."""
    G = Grid(n, m, lshape)
    domain = (capi.DOMAIN_LSHAPE if n == m and n % 2 == 0 else capi.DOMAIN_LSHAPE_ANY) if lshape else capi.DOMAIN_RECT
    rng = np.random.default_rng(n * 1000 + m)
    b = np.where(G.mask, rng.standard_normal(G.mask.shape), 0.0)
    bp = G.to_pitched(b)
    slabs, tiles = [], []
    for rank in range(world):
        ylo, yhi, _lo, _hi, _n = capi.partition(m, n, domain=domain, rank=rank, world=world)
        assert yhi - ylo >= 4
        sl = Slab(G, ylo, yhi, rank > 0, rank + 1 < world)
        for y in range(ylo - 2, yhi + 2):  # r0 = b incl. both halo rows (what the initial exchanges deliver)
            ri = sl.row_index(y)
            if ri >= 0:
                sl.r[0][ri] = bp[y]
        slabs.append(sl)
        tiles.append(capi.work_split(m, n, domain=domain, rank=rank, world=world, sms=sms, ctas_per_sm=2,
                                     tile_rows=tile_rows, fused=True)[0])

    r = b.copy(); p = np.zeros_like(b); xs = np.zeros_like(b)
    gamma = float(np.sum(r * r)); alpha = gamma / float(np.sum(r * G.apply(r))); beta = 0.0
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

    gamma = float(np.sum(b * b)); alpha = gamma / float(np.sum(b * G.apply(b))); beta = 0.0; alpha_prev = 0.0
    worst = 0.0
    for k in range(iters):
        par = k & 1
        g2 = d2 = 0.0
        for rank, sl in enumerate(slabs):
            below = slabs[rank - 1] if rank > 0 else None
            above = slabs[rank + 1] if rank + 1 < world else None
            gg, dd = sweep(G, tiles[rank], sl.r[par], sl.p[par], sl.x, sl.r[par ^ 1], sl.p[par ^ 1], alpha, beta, alpha_prev,
                           x2=bool(k & 1), slab=sl,
                           nb_below=((below.r[par ^ 1], below.p[par ^ 1]), below.yrows) if below else None,
                           nb_above=((above.r[par ^ 1], above.p[par ^ 1]), above.yrows) if above else None)
            g2 += gg
            d2 += dd
        worst = max(worst, abs(g2 - hist[k][0]) / hist[k][0], abs(d2 - hist[k][1]) / abs(hist[k][1]))
        alpha_prev = alpha if not (k & 1) else 0.0
        beta = g2 / gamma
        alpha, gamma = g2 / (d2 - beta * g2 / alpha), g2
    xg = np.zeros((m + 1, G.pitch)); rg = np.zeros((m + 1, G.pitch))
    for sl in slabs:
        xo = sl.x + (alpha_prev * sl.p[iters & 1] if iters & 1 else 0.0)  # x_flush_kernel
        xg[sl.ylo:sl.yhi] = xo[1:1 + sl.yhi - sl.ylo]
        rg[sl.ylo:sl.yhi] = sl.r[iters & 1][1:1 + sl.yhi - sl.ylo]
    xr, rr = G.from_pitched(xg), G.from_pitched(rg)
    assert np.all(np.isfinite(xr)) and not np.any(xr[~G.mask]) and not np.any(rr[~G.mask])
    return worst, np.max(np.abs(xr - xs)) / np.max(np.abs(xs)), np.max(np.abs(rr - r)) / np.max(np.abs(r))


def run_sharded_maxn(n, m, lshape, iters, world, sms=4, tile_rows=0, with_u=True):
    """F_SHARD | F_MAXN: the max-norm flavour on row slabs - every slab streams its own rows of x and u, r' and p cross the
    slab edges as in run_sharded, the sums are added and the maxima maximised over the ranks."""
    G = Grid(n, m, lshape)
    domain = (capi.DOMAIN_LSHAPE if n == m and n % 2 == 0 else capi.DOMAIN_LSHAPE_ANY) if lshape else capi.DOMAIN_RECT
    rng = np.random.default_rng(n * 1000 + m + 11)
    b = np.where(G.mask, rng.standard_normal(G.mask.shape), 0.0)
    ut = np.where(G.mask, rng.standard_normal(G.mask.shape), 0.0)
    bp, up = G.to_pitched(b), G.to_pitched(ut)
    slabs, tiles, us = [], [], []
    for rank in range(world):
        ylo, yhi, _lo, _hi, _n = capi.partition(m, n, domain=domain, rank=rank, world=world)
        assert yhi - ylo >= 4
        sl = Slab(G, ylo, yhi, rank > 0, rank + 1 < world)
        uu = np.zeros_like(sl.x)
        for y in range(ylo - 2, yhi + 2):
            ri = sl.row_index(y)
            if ri >= 0:
                sl.r[0][ri] = bp[y]
                uu[ri] = up[y]
        slabs.append(sl)
        us.append(uu if with_u else None)
        tiles.append(capi.work_split(m, n, domain=domain, rank=rank, world=world, sms=sms, ctas_per_sm=2,
                                     tile_rows=tile_rows, fused=True)[0])

    r = b.copy(); z = np.zeros_like(b); xs = np.zeros_like(b)
    gamma = float(np.sum(r * r)); alpha = gamma / float(np.sum(r * G.apply(r))); beta = 0.0; rz = gamma
    hist = []
    for _ in range(iters):
        z = r + beta * z
        xn = xs + alpha * z
        dmax = float(np.max(np.abs(xn - xs))); xs = xn
        r = r - alpha * G.apply(z)
        g2 = float(np.sum(r * r)); d2 = float(np.sum(r * G.apply(r)))
        hist.append((g2, d2, float(np.max(np.abs(r))), dmax, float(np.max(np.abs(xs - ut)))))
        beta = (np.sqrt(g2) * np.sqrt(g2)) / rz
        alpha = g2 / (d2 - beta * g2 / alpha)
        rz = g2

    gamma = float(np.sum(b * b)); alpha = gamma / float(np.sum(b * G.apply(b))); beta = 0.0; rz = gamma
    worst = 0.0
    for k in range(iters):
        par = k & 1
        g2 = d2 = 0.0
        mx = [0.0, 0.0, 0.0]
        for rank, sl in enumerate(slabs):
            below = slabs[rank - 1] if rank > 0 else None
            above = slabs[rank + 1] if rank + 1 < world else None
            gg, dd, mm = sweep(G, tiles[rank], sl.r[par], sl.p[par], sl.x, sl.r[par ^ 1], sl.p[par ^ 1], alpha, beta, 0.0,
                               x2=False, slab=sl,
                               nb_below=((below.r[par ^ 1], below.p[par ^ 1]), below.yrows) if below else None,
                               nb_above=((above.r[par ^ 1], above.p[par ^ 1]), above.yrows) if above else None,
                               maxn=True, u=us[rank])
            g2 += gg
            d2 += dd
            mx = [max(a, c) for a, c in zip(mx, mm)]
        ref = hist[k]
        worst = max(worst, abs(g2 - ref[0]) / ref[0], abs(d2 - ref[1]) / abs(ref[1]), abs(mx[0] - ref[2]) / ref[2],
                    abs(mx[1] - ref[3]) / ref[3], abs(mx[2] - ref[4]) / ref[4] if with_u else 0.0)
        beta = (np.sqrt(g2) * np.sqrt(g2)) / rz
        alpha = g2 / (d2 - beta * g2 / alpha)
        rz = g2
    xg = np.zeros((m + 1, G.pitch))
    for sl in slabs:
        xg[sl.ylo:sl.yhi] = sl.x[1:1 + sl.yhi - sl.ylo]
    xr = G.from_pitched(xg)
    assert np.all(np.isfinite(xr)) and not np.any(xr[~G.mask])
    return worst, np.max(np.abs(xr - xs)) / np.max(np.abs(xs))


if __name__ == "__main__":
    for n, m, lshape, iters, tr in [(30, 30, True, 7, 0), (64, 64, True, 6, 0), (64, 64, True, 5, 5), (130, 90, True, 6, 0),
                                    (77, 33, False, 6, 3), (1000, 40, True, 4, 0), (970, 24, False, 3, 7),
                                    (900, 30, True, 3, 0), (430, 26, False, 3, 4), (845, 64, True, 3, 0)]:
        worst, dx, dr, nt = run(n, m, lshape, iters, tr)
        print(f"n={n} m={m} {'L' if lshape else 'rect'} tile_rows={tr} tiles={nt}: dots {worst:.1e}, x {dx:.1e}, r {dr:.1e}")
        assert worst < 1e-12 and dx < 1e-10 and dr < 1e-10
    for n, m, lshape, iters, tr in [(64, 64, True, 4, 0), (900, 30, True, 3, 0), (1700, 26, False, 3, 4), (1690, 64, True, 3, 0)]:
        worst, dx, dr, nt = run(n, m, lshape, iters, tr, warps=14)
        print(f"n={n} m={m} {'L' if lshape else 'rect'} tile_rows={tr} tiles={nt} wide geometry: dots {worst:.1e}, x {dx:.1e}, r {dr:.1e}")
        assert worst < 1e-12 and dx < 1e-10 and dr < 1e-10
    for n, m, lshape, iters, world, tr in [(64, 64, True, 6, 2, 0), (64, 64, True, 5, 3, 0), (130, 90, True, 5, 4, 0),
                                           (77, 60, False, 5, 3, 5), (1000, 40, True, 4, 2, 0), (96, 96, True, 7, 8, 0)]:
        worst, dx, dr = run_sharded(n, m, lshape, iters, world, tile_rows=tr)
        print(f"n={n} m={m} {'L' if lshape else 'rect'} {world} slabs tile_rows={tr}: dots {worst:.1e}, x {dx:.1e}, r {dr:.1e}")
        assert worst < 1e-12 and dx < 1e-12 and dr < 1e-12
    for n, m, lshape, iters, tr, warps, with_u in [(30, 30, True, 5, 0, 7, True), (64, 64, True, 5, 5, 7, True),
                                                   (130, 90, True, 4, 0, 7, False), (77, 33, False, 4, 3, 7, True),
                                                   (1000, 40, True, 3, 0, 7, True), (430, 26, False, 3, 4, 7, True),
                                                   (900, 30, True, 3, 0, 14, True), (1700, 26, False, 3, 4, 14, False)]:
        worst, dx, nt = run_maxn(n, m, lshape, iters, tr, warps=warps, with_u=with_u)
        print(f"n={n} m={m} {'L' if lshape else 'rect'} tile_rows={tr} tiles={nt} warps={warps} max-norm flavour"
              f"{'' if with_u else ' (no u)'}: sums and maxima {worst:.1e}, x {dx:.1e}")
        assert worst < 1e-12 and dx < 1e-12
    for n, m, lshape, iters, world, tr, with_u in [(64, 64, True, 5, 2, 0, True), (130, 90, True, 4, 4, 0, True),
                                                   (77, 60, False, 4, 3, 5, False), (1000, 40, True, 3, 2, 0, True),
                                                   (96, 96, True, 5, 8, 0, True)]:
        worst, dx = run_sharded_maxn(n, m, lshape, iters, world, tile_rows=tr, with_u=with_u)
        print(f"n={n} m={m} {'L' if lshape else 'rect'} {world} slabs tile_rows={tr} max-norm flavour"
              f"{'' if with_u else ' (no u)'}: sums and maxima {worst:.1e}, x {dx:.1e}")
        assert worst < 1e-12 and dx < 1e-12
    print("MODEL_OK")
