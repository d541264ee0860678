"""CPU model of the default two-sweep kernel's data flow (csrc/stream_kernel.cuh), thread by thread: 512-column strips
writing 504, the halo threads, the warp-edge neighbour (c_edge), HS-row stages with the unrolled FULL path and its
conditions, x-deferral (F_NOX / F_X2), row slabs with one halo row per side and the peer-memory halo stores. It runs
whole CG iterations from the real tile tables (b200cg_work_split, no GPU) and compares them with a plain numpy CG in
the reference's form. Stale shared-memory columns are NaN.

A design check of index arithmetic (it reproduces, e.g., the rule that a FULL stage must not hold a tile's first emit
row - with 2-row stages that row would otherwise never reach the neighbour's halo):  python scripts/model_two_sweep.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_solvers_b200 import capi  # noqa: E402

XOFF, STRIP_LOAD, CONS_THREADS = 4, 512, 256


class Grid:
    def __init__(self, n, m, lshape=True):
        self.n, self.m = n, m
        self.xsplit, self.ysplit = (n // 2, m // 2) if lshape else (0, 0)
        self.pitch = (n + 1 + XOFF + 15) // 16 * 16
        hx, hy = 1.0 / n, 1.0 / m
        self.xk, self.yk = 1 / (hx * hx), 1 / (hy * hy)
        self.A = -2 * (self.xk + self.yk)
        self.mask = np.zeros((m + 1, n + 1), dtype=bool)
        self.mask[1:m, 1:n] = True
        if lshape:
            self.mask[1:m // 2 + 1, 1:n // 2 + 1] = False

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
    """Rows ylo-1 .. yhi of one rank, row y at index y - ybase (one halo row per side)."""

    def __init__(self, G, ylo, yhi):
        self.ylo, self.yhi, self.ybase, self.yrows = ylo, yhi, ylo - 1, yhi - ylo + 2
        shape = (self.yrows, G.pitch)
        self.r = [np.zeros(shape), np.zeros(shape)]
        self.p = [np.zeros(shape), np.zeros(shape)]
        self.x = np.zeros(shape)


def sweep(G, sl, tiles, mode, par, alpha, beta, alpha_prev, flags, hs, nb_below=None, nb_above=None):
    """mode 'dot': returns (p.Ap, r.p). mode 'upd': writes r', p (and x unless flags == 'nox'), returns r'.r'.
    flags: None (x every iteration), 'nox', 'x2'. nb_*: (r_out, p_out, yrows) of the neighbour slabs (peer stores)."""
    tid = np.arange(CONS_THREADS)
    lane = tid & 31
    c2 = 2 * tid
    is_out = (tid >= 2) & (tid < CONS_THREADS - 2)
    c_edge = np.where(lane == 0, np.maximum(c2 - 1, 0), np.where(lane == 31, np.minimum(c2 + 2, STRIP_LOAD - 1), c2))
    r_in, p_in, r_out, p_out, x = sl.r[par], sl.p[par], sl.r[par ^ 1], sl.p[par ^ 1], sl.x
    s0 = s1 = 0.0

    def shfl_up(v):
        out = v.copy()
        keep = lane >= 1
        out[keep] = v[np.where(keep)[0] - 1]
        return out

    def shfl_down(v):
        out = v.copy()
        keep = lane <= 30
        out[keep] = v[np.where(keep)[0] + 1]
        return out

    for col0, ya, yb, xlo in tiles:
        row_doubles = min(STRIP_LOAD, G.pitch - col0)
        x0 = col0 + c2 - XOFF
        v0 = is_out & (x0 >= xlo) & (x0 <= G.n - 1)
        v1 = is_out & (x0 + 1 >= xlo) & (x0 + 1 <= G.n - 1)
        z = np.zeros(CONS_THREADS)
        pmx, pmy, pcx, pcy, Lp, Rp = z, z, z, z, z, z
        rpx, rpy, xpx, xpy, qpx, qpy = z, z, z, z, z, z
        S = yb - ya + 2  # rows ya-1 .. yb
        for st0 in range(0, S, hs):
            nrows = min(hs, S - st0)
            y0 = ya - 1 + st0
            first = st0 == 0
            full = nrows == hs and not first and (y0 - 1 > ya) and (y0 + hs <= yb)
            for j in range(nrows):
                y = y0 + j

                def stage(src):
                    buf = np.full(STRIP_LOAD, np.nan)
                    buf[:row_doubles] = src[y - sl.ybase, col0:col0 + row_doubles]
                    return buf
                sp, sr = stage(p_in), stage(r_in)
                cpx, cpy, crx, cry = sp[c2], sp[c2 + 1], sr[c2], sr[c2 + 1]
                inner = full or (ya <= y < yb)
                if mode == "upd" and flags != "nox" and inner:
                    sx = stage(x)
                    cxx, cxy = sx[c2], sx[c2 + 1]
                else:
                    cxx = cxy = z
                pnx, pny = crx + beta * cpx, cry + beta * cpy
                pe = sr[c_edge] + beta * sp[c_edge]
                L = np.where(lane == 0, pe, shfl_up(pny))
                R = np.where(lane == 31, pe, shfl_down(pnx))
                if full or y > ya:
                    ap0 = G.A * pcx
                    ap0 = ap0 + G.xk * Lp
                    ap0 = ap0 + G.xk * pcy
                    ap0 = ap0 + G.yk * pnx
                    ap0 = ap0 + G.yk * pmx
                    ap1 = G.A * pcy
                    ap1 = ap1 + G.xk * pcx
                    ap1 = ap1 + G.xk * Rp
                    ap1 = ap1 + G.yk * pny
                    ap1 = ap1 + G.yk * pmy
                    ap0, ap1 = np.where(v0, ap0, 0.0), np.where(v1, ap1, 0.0)
                    p0, p1 = np.where(v0, pcx, 0.0), np.where(v1, pcy, 0.0)
                    r0, r1 = np.where(v0, rpx, 0.0), np.where(v1, rpy, 0.0)
                    st_ok = v0 | v1
                    if mode == "dot":
                        s0 += float(np.sum(p0 * ap0) + np.sum(p1 * ap1))
                        s1 += float(np.sum(r0 * p0) + np.sum(r1 * p1))
                    else:
                        xo0, xo1 = np.where(v0, xpx, 0.0), np.where(v1, xpy, 0.0)
                        if flags == "x2":
                            xo0 = xo0 + alpha_prev * np.where(v0, qpx, 0.0)
                            xo1 = xo1 + alpha_prev * np.where(v1, qpy, 0.0)
                        xn0, xn1 = xo0 + alpha * p0, xo1 + alpha * p1
                        rn0, rn1 = r0 - alpha * ap0, r1 - alpha * ap1
                        cols = (col0 + c2)[st_ok]
                        yy = y - 1 - sl.ybase
                        if flags != "nox":
                            x[yy, cols], x[yy, cols + 1] = xn0[st_ok], xn1[st_ok]
                        r_out[yy, cols], r_out[yy, cols + 1] = rn0[st_ok], rn1[st_ok]
                        p_out[yy, cols], p_out[yy, cols + 1] = p0[st_ok], p1[st_ok]
                        ye = y - 1
                        if not full and nb_below is not None and ye == sl.ylo:  # -> its top halo row
                            nr, npp, rows = nb_below
                            nr[rows - 1, cols], nr[rows - 1, cols + 1] = rn0[st_ok], rn1[st_ok]
                            npp[rows - 1, cols], npp[rows - 1, cols + 1] = p0[st_ok], p1[st_ok]
                        if not full and nb_above is not None and ye == sl.yhi - 1:  # -> its bottom halo row
                            nr, npp, _rows = nb_above
                            nr[0, cols], nr[0, cols + 1] = rn0[st_ok], rn1[st_ok]
                            npp[0, cols], npp[0, cols + 1] = p0[st_ok], p1[st_ok]
                        s0 += float(np.sum(rn0 * rn0) + np.sum(rn1 * rn1))
                pmx, pmy, pcx, pcy, Lp, Rp = pcx, pcy, pnx, pny, L, R
                rpx, rpy, xpx, xpy, qpx, qpy = crx, cry, cxx, cxy, cpx, cpy
    return s0, s1


def run(n, m, lshape, iters, world=1, hs=4, tile_rows=0, xdefer=True, sms=4):
    G = Grid(n, m, lshape)
    domain = (capi.DOMAIN_LSHAPE if n == m and n % 2 == 0 else capi.DOMAIN_LSHAPE_ANY) if lshape else capi.DOMAIN_RECT
    rng = np.random.default_rng(n * 1000 + m)
    b = np.where(G.mask, rng.standard_normal(G.mask.shape), 0.0)
    bp = np.zeros((m + 1, G.pitch))
    bp[:, XOFF:XOFF + n + 1] = b
    slabs, tiles = [], []
    for rank in range(world):
        ylo, yhi, _lo, _hi, _n = capi.partition(m, n, domain=domain, rank=rank, world=world)
        sl = Slab(G, ylo, yhi)
        sl.r[0][:] = bp[ylo - 1:yhi + 1]  # r0 = b with both halo rows (the init exchange)
        slabs.append(sl)
        tiles.append(capi.work_split(m, n, domain=domain, rank=rank, world=world, sms=sms, ctas_per_sm=2,
                                     tile_rows=tile_rows)[0])

    # reference-form CG (matrix_free_system.cpp:409-441)
    r = b.copy(); p = r.copy(); xs = np.zeros_like(b); rr = float(np.sum(r * r))
    for _ in range(iters):
        Ap = G.apply(p)
        al = rr / float(np.sum(p * Ap))
        xs = xs + al * p
        r = r - al * Ap
        rr_new = float(np.sum(r * r))
        p = r + (rr_new / rr) * p
        rr = rr_new

    rr = float(np.sum(b * b)); beta = 0.0; alpha_prev = 0.0
    for k in range(iters):
        par = k & 1
        pAp = sum(sweep(G, sl, tiles[i], "dot", par, 0.0, beta, 0.0, None, hs)[0] for i, sl in enumerate(slabs))
        alpha = rr / pAp
        flags = ("x2" if (k & 1) else "nox") if xdefer else None
        rr_new = 0.0
        for i, sl in enumerate(slabs):
            below = slabs[i - 1] if i > 0 else None
            above = slabs[i + 1] if i + 1 < world else None
            rr_new += sweep(G, sl, tiles[i], "upd", par, alpha, beta, alpha_prev, flags, hs,
                            nb_below=(below.r[par ^ 1], below.p[par ^ 1], below.yrows) if below else None,
                            nb_above=(above.r[par ^ 1], above.p[par ^ 1], above.yrows) if above else None)[0]
        alpha_prev = alpha if flags == "nox" else 0.0
        beta = rr_new / rr
        rr = rr_new
    xg = np.zeros((m + 1, G.pitch)); rg = np.zeros((m + 1, G.pitch))
    for sl in slabs:
        xo = sl.x + (alpha_prev * sl.p[iters & 1] if (xdefer and iters & 1) else 0.0)  # x_flush_kernel
        xg[sl.ylo:sl.yhi] = xo[1:1 + sl.yhi - sl.ylo]
        rg[sl.ylo:sl.yhi] = sl.r[iters & 1][1:1 + sl.yhi - sl.ylo]
    xr, rres = xg[:, XOFF:XOFF + n + 1], rg[:, XOFF:XOFF + n + 1]
    assert np.all(np.isfinite(xr)) and not np.any(xr[~G.mask]) and not np.any(rres[~G.mask])
    return np.max(np.abs(xr - xs)) / np.max(np.abs(xs)), np.max(np.abs(rres - r)) / np.max(np.abs(r))


if __name__ == "__main__":
    for n, m, lshape, iters, world, hs, tr, xd in [(30, 30, True, 5, 1, 4, 0, True), (64, 64, True, 4, 1, 2, 0, True),
                                                   (64, 64, True, 5, 2, 2, 0, True), (64, 64, True, 4, 3, 4, 0, False),
                                                   (77, 60, False, 4, 2, 4, 5, True), (1030, 24, True, 3, 2, 2, 0, True)]:
        dx, dr = run(n, m, lshape, iters, world, hs, tr, xd)
        print(f"n={n} m={m} {'L' if lshape else 'rect'} ranks={world} HS={hs} tile_rows={tr} xdefer={xd}: x {dx:.1e}, r {dr:.1e}")
        assert dx < 1e-12 and dr < 1e-12
    print("MODEL_OK")
