"""ctypes mirror of include/b200cg.h. Thin by design: argument marshalling and status -> exception only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200CG_LIB", os.path.join(PKG, "libb200cg.so"))

DOMAIN_LSHAPE, DOMAIN_RECT, DOMAIN_GENERIC, DOMAIN_LSHAPE_ANY = 0, 1, 2, 3
OP_MATRIX_FREE, OP_CSR = 0, 1
RULE_REL_L2, RULE_MAXNORM = 0, 1
PRECOND_NONE, PRECOND_MULTIGRID = 0, 1
STOP_NAMES = ["ITERATIONS", "PRECISION", "RESIDUAL", "EXACT_ERROR", "INTERRUPTED"]
ERR_NO_DEVICE = 2

# every symbol include/b200cg.h declares (tests/test_abi.py checks the two lists against each other)
EXPORTS = [
    "b200cg_last_error", "b200cg_version", "b200cg_device_count", "b200cg_alloc_pinned", "b200cg_free_pinned",
    "b200cg_comm_unique_id", "b200cg_plan_create", "b200cg_plan_destroy", "b200cg_size", "b200cg_local_range",
    "b200cg_partition", "b200cg_work_split",
    "b200cg_build_rhs", "b200cg_set_rhs", "b200cg_get_rhs", "b200cg_get_true_solution", "b200cg_get_coords",
    "b200cg_apply", "b200cg_set_csr", "b200cg_assemble_csr", "b200cg_get_csr", "b200cg_csr_apply",
    "b200cg_solve", "b200cg_solve_batch", "b200cg_postprocess", "b200cg_get_solution", "b200cg_cta_times", "b200cg_peer_trace",
]


class B200CGError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"b200cg status {status}: {message}")
        self.status = status


class PlanDesc(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
                ("d", C.c_double), ("domain", C.c_int), ("device", C.c_int), ("rank", C.c_int),
                ("world", C.c_int), ("comm_id", C.c_void_p), ("tile_rows", C.c_int), ("reserved0", C.c_int),
                ("generic_rows", C.c_int64), ("reserved", C.c_int * 4)]


class Params(C.Structure):
    _fields_ = [("op", C.c_int), ("rule", C.c_int), ("eps_rel", C.c_double), ("eps_p", C.c_double),
                ("eps_r", C.c_double), ("eps_e", C.c_double), ("max_it", C.c_int), ("callback_every", C.c_int),
                ("rhs_on_device", C.c_int), ("keep_x_on_device", C.c_int), ("iters_per_graph", C.c_int),
                ("small_grid_path", C.c_int), ("single_sweep", C.c_int), ("preconditioner", C.c_int),
                ("reserved", C.c_int * 4)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("stop_reason", C.c_int),
                ("r0_l2", C.c_double), ("r_l2", C.c_double), ("r_max", C.c_double), ("dx_max", C.c_double),
                ("err_max", C.c_double), ("total_ms", C.c_double), ("solve_ms", C.c_double), ("device_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("kernel_launches", C.c_int64), ("dot_kernel_ms", C.c_double),
                ("upd_kernel_ms", C.c_double), ("kernel_samples", C.c_int), ("local_unknowns", C.c_int64),
                ("upd_even_ms", C.c_double), ("upd_odd_ms", C.c_double), ("x_deferral", C.c_int),
                ("cluster_path", C.c_int), ("peer_exchange", C.c_int), ("single_sweep", C.c_int),
                ("preconditioner", C.c_int), ("mg_levels", C.c_int)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["stop_reason"] = STOP_NAMES[self.stop_reason]
        d["converged"] = bool(self.converged)
        return d


ITER_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double)
BATCH_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(SolveInfo))

_lib = None


def lib():
    """Loads libb200cg.so. Raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200CGError(-1, f"{LIB_PATH} is missing: run `python -m iterative_solvers_b200.build` "
                                  "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.b200cg_last_error.restype = C.c_char_p
        L.b200cg_plan_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(PlanDesc)]
        L.b200cg_solve.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(SolveInfo), C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200cg_solve_batch.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_void_p), C.POINTER(SolveInfo), C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("b200cg_plan_destroy", "b200cg_build_rhs"):
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("b200cg_set_rhs", "b200cg_get_rhs", "b200cg_get_true_solution", "b200cg_get_solution"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        for name in ("b200cg_get_coords", "b200cg_apply", "b200cg_csr_apply"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200cg_work_split.argtypes = [C.POINTER(PlanDesc), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64,
                                        C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int)]
        L.b200cg_size.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.b200cg_local_range.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.b200cg_set_csr.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200cg_assemble_csr.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.b200cg_get_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200cg_postprocess.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b200cg_alloc_pinned.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.b200cg_free_pinned.argtypes = [C.c_void_p]
        L.b200cg_comm_unique_id.argtypes = [C.c_void_p]
        L.b200cg_device_count.argtypes = [C.POINTER(C.c_int)]
        L.b200cg_cta_times.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.b200cg_peer_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise B200CGError(status, lib().b200cg_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    status = lib().b200cg_device_count(C.byref(n))
    return n.value if status == 0 else 0


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().b200cg_comm_unique_id(buf))
    return buf.raw


def partition(m, n, domain=DOMAIN_LSHAPE, rank=0, world=1):
    """Row slab of `rank` among `world`: (y_lo, y_hi, lo, hi, N). Pure geometry - works without a GPU."""
    desc = PlanDesc(n=int(n), m=int(m), a=0.0, b=1.0, c=0.0, d=1.0, domain=int(domain), device=0, rank=int(rank),
                    world=int(world))
    ylo, yhi, lo, hi, N = C.c_int(), C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
    check(lib().b200cg_partition(C.byref(desc), C.byref(ylo), C.byref(yhi), C.byref(lo), C.byref(hi), C.byref(N)))
    return ylo.value, yhi.value, lo.value, hi.value, N.value


def work_split(m, n, domain=DOMAIN_LSHAPE, rank=0, world=1, sms=148, ctas_per_sm=2, weights=None, tile_rows=0,
               fused=False):
    """The sweep kernels' tile table for such a plan: (tiles[k, 4] = col0, ya, yb, xlo; cta_begin[grid + 1]).
    fused: the single-sweep kernel's strip geometry (1: 420 written columns from storage column strip * 420 + 2; 2: the wide
    one of large slabs, 840 written columns).
    Pure host logic (b200cg_work_split) - works without a GPU."""
    desc = PlanDesc(n=int(n), m=int(m), a=0.0, b=1.0, c=0.0, d=1.0, domain=int(domain), device=0, rank=int(rank),
                    world=int(world), tile_rows=int(tile_rows), reserved0=int(fused))
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    nw = 0 if w is None else int(w.size)
    count, grid = C.c_int64(), C.c_int()
    check(lib().b200cg_work_split(C.byref(desc), int(sms), int(ctas_per_sm), _ptr(w), nw, None, C.c_int64(0),
                                  C.byref(count), None, C.byref(grid)))
    tiles = np.zeros((max(count.value, 1), 4), dtype=np.int32)
    cta_begin = np.zeros(int(sms) * int(ctas_per_sm) + 1, dtype=np.int32)
    check(lib().b200cg_work_split(C.byref(desc), int(sms), int(ctas_per_sm), _ptr(w), nw, _ptr(tiles),
                                  C.c_int64(tiles.shape[0]), C.byref(count), _ptr(cta_begin), C.byref(grid)))
    return tiles[:count.value], cta_begin[:grid.value + 1]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PinnedArray:
    """fp64 host array in page-locked memory (b200cg_alloc_pinned) exposed as a numpy view."""

    def __init__(self, n):
        self.ptr = C.c_void_p()
        check(lib().b200cg_alloc_pinned(C.byref(self.ptr), C.c_size_t(max(int(n), 1) * 8)))
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_double)), shape=(int(n),))

    def free(self):
        if self.ptr:
            self.array = None
            lib().b200cg_free_pinned(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Plan:
    """b200cg_plan_t. Constructor arguments follow GridSystem / MatrixFreeSystem: (m, n, a, b, c, d)."""

    def __init__(self, m, n, a=0.0, b=1.0, c=0.0, d=1.0, domain=DOMAIN_LSHAPE, device=0, rank=0, world=1,
                 comm_id: bytes | None = None, tile_rows=0, generic_rows=0):
        self.L = lib()
        self.h = C.c_void_p()
        self._id = C.create_string_buffer(comm_id, 128) if comm_id else None
        desc = PlanDesc(n=int(n), m=int(m), a=a, b=b, c=c, d=d, domain=int(domain), device=int(device),
                        rank=int(rank), world=int(world),
                        comm_id=C.cast(self._id, C.c_void_p) if self._id else None, tile_rows=int(tile_rows),
                        generic_rows=int(generic_rows))
        check(self.L.b200cg_plan_create(C.byref(self.h), C.byref(desc)))
        nn = C.c_int64()
        check(self.L.b200cg_size(self.h, C.byref(nn)))
        self.N = nn.value
        lo, hi = C.c_int64(), C.c_int64()
        check(self.L.b200cg_local_range(self.h, C.byref(lo), C.byref(hi)))
        self.lo, self.hi = lo.value, hi.value
        self.n_local = self.hi - self.lo

    def close(self):
        if getattr(self, "h", None):
            self.L.b200cg_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- setup data
    def build_rhs(self):
        check(self.L.b200cg_build_rhs(self.h))

    def set_rhs(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert b.shape == (self.n_local,)
        check(self.L.b200cg_set_rhs(self.h, _ptr(b)))

    def _out(self, fn):
        out = np.empty(self.n_local)
        check(fn(self.h, _ptr(out)))
        return out

    def get_rhs(self):
        return self._out(self.L.b200cg_get_rhs)

    def true_solution(self):
        return self._out(self.L.b200cg_get_true_solution)

    def get_solution(self):
        return self._out(self.L.b200cg_get_solution)

    def coords(self):
        xs, ys = np.empty(self.n_local), np.empty(self.n_local)
        check(self.L.b200cg_get_coords(self.h, _ptr(xs), _ptr(ys)))
        return xs, ys

    # ---- operator
    def apply(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.n_local,)
        y = np.empty(self.n_local)
        check(self.L.b200cg_apply(self.h, _ptr(x), _ptr(y)))
        return y

    def set_csr(self, row_map, entries, values):
        row_map = np.ascontiguousarray(row_map, dtype=np.int32)
        entries = np.ascontiguousarray(entries, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        check(self.L.b200cg_set_csr(self.h, len(row_map) - 1, len(values), _ptr(row_map), _ptr(entries),
                                    _ptr(values)))

    def assemble_csr(self):
        nnz = C.c_int64()
        check(self.L.b200cg_assemble_csr(self.h, C.byref(nnz)))
        return nnz.value

    def get_csr(self, nnz):
        row_map = np.empty(self.N + 1, dtype=np.int32)
        entries = np.empty(nnz, dtype=np.int32)
        values = np.empty(nnz)
        check(self.L.b200cg_get_csr(self.h, _ptr(row_map), _ptr(entries), _ptr(values)))
        return row_map, entries, values

    def csr_apply(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.N)
        check(self.L.b200cg_csr_apply(self.h, _ptr(x), _ptr(y)))
        return y

    # ---- solve
    def solve(self, b=None, u=None, x_out=None, op=OP_MATRIX_FREE, rule=RULE_REL_L2, eps_rel=1e-6, eps_p=-1.0,
              eps_r=-1.0, eps_e=-1.0, max_it=10000, callback=None, callback_every=100, rhs_on_device=False,
              keep_x_on_device=False, iters_per_graph=0, stop_flag=None, small_grid_path=0, single_sweep=0,
              preconditioner=0):
        """Returns (x, info dict). b / u / x_out may be numpy arrays or PinnedArray.array views."""
        prm = Params(op=op, rule=rule, eps_rel=eps_rel, eps_p=eps_p, eps_r=eps_r, eps_e=eps_e, max_it=int(max_it),
                     callback_every=int(callback_every), rhs_on_device=int(bool(rhs_on_device)),
                     keep_x_on_device=int(bool(keep_x_on_device)), iters_per_graph=int(iters_per_graph),
                     small_grid_path=int(small_grid_path), single_sweep=int(single_sweep),
                     preconditioner=int(preconditioner))
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.float64)
            assert b.shape == (self.n_local,)
        if u is not None:
            u = np.ascontiguousarray(u, dtype=np.float64)
        if x_out is None and not keep_x_on_device:
            x_out = np.empty(self.n_local)
        info = SolveInfo()
        cb = None
        if callback is not None:
            cb = ITER_CB(lambda _user, it, p, r, e: callback(it, p, r, e))
        flag_ptr = None if stop_flag is None else C.cast(C.byref(stop_flag), C.c_void_p)
        check(self.L.b200cg_solve(self.h, C.byref(prm), _ptr(b), _ptr(u), _ptr(x_out), C.byref(info),
                                  C.cast(cb, C.c_void_p) if cb else None, None, flag_ptr))
        return x_out, info.as_dict()

    def solve_batch(self, bs, xs, done=None, rule=RULE_REL_L2, eps_rel=1e-6, max_it=10000, iters_per_graph=0,
                    stop_flag=None, small_grid_path=0, single_sweep=0, preconditioner=0, op=OP_MATRIX_FREE):
        """b200cg_solve_batch: one matrix-free solve per right-hand side of `bs` into the arrays of `xs` (numpy arrays or
        PinnedArray.array views; the lists may repeat buffers), copies overlapped with the iterations. done(i, info dict)
        fires once xs[i] is complete. Returns the list of info dicts."""
        count = len(bs)
        assert len(xs) == count
        prm = Params(op=op, rule=rule, eps_rel=eps_rel, eps_p=-1.0, eps_r=-1.0, eps_e=-1.0, max_it=int(max_it),
                     callback_every=0, rhs_on_device=0, keep_x_on_device=0, iters_per_graph=int(iters_per_graph),
                     small_grid_path=int(small_grid_path), single_sweep=int(single_sweep),
                     preconditioner=int(preconditioner))
        for a in list(bs) + list(xs):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape == (self.n_local,)
        bp = (C.c_void_p * max(count, 1))(*[a.ctypes.data for a in bs])
        xp = (C.c_void_p * max(count, 1))(*[a.ctypes.data for a in xs])
        infos = (SolveInfo * max(count, 1))()
        cb = BATCH_CB(lambda _user, i, info: done(i, info.contents.as_dict())) if done is not None else None
        flag_ptr = None if stop_flag is None else C.cast(C.byref(stop_flag), C.c_void_p)
        check(self.L.b200cg_solve_batch(self.h, C.byref(prm), count, bp, xp, infos,
                                        C.cast(cb, C.c_void_p) if cb else None, None, flag_ptr))
        return [infos[i].as_dict() for i in range(count)]

    def cta_times(self, flavour):
        """(start_ns, end_ns) of every persistent CTA in the last launch of a sweep-kernel flavour (diagnostics)."""
        cap = 1024
        buf = np.zeros(2 * cap, dtype=np.uint64)
        n = C.c_int()
        check(self.L.b200cg_cta_times(self.h, int(flavour), _ptr(buf), cap, C.byref(n)))
        return buf[: 2 * n.value].reshape(-1, 2).astype(np.int64)

    def peer_trace(self):
        """[4096, 4] global-timer stamps of the sharded single sweep's cross-rank step (plans created with
        B200CG_PEER_TRACE=1): sums ready, published, all flags seen, scalars formed; row = iteration % 4096."""
        buf = np.zeros(4 * 4096, dtype=np.uint64)
        n = C.c_int()
        check(self.L.b200cg_peer_trace(self.h, _ptr(buf), 4096, C.byref(n)))
        return buf[: 4 * n.value].reshape(-1, 4).astype(np.int64)

    def postprocess(self, op=OP_MATRIX_FREE, want_residual=True, want_error=True):
        res = np.empty(self.n_local) if want_residual else None
        err = np.empty(self.n_local) if want_error else None
        check(self.L.b200cg_postprocess(self.h, int(op), _ptr(res), _ptr(err)))
        return res, err
