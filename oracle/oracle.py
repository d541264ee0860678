"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the CPU oracle.

* ``Oracle``    - the plain-C restatement (oracle/cg_oracle.c -> oracle/_build/libcg_oracle.so).
* ``Reference`` - the UNMODIFIED reference sources compiled in place (oracle/_ref/libref_cg.so,
  recipe in oracle/Makefile); only available where it was built (this container) or travelled to.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this module - as the checker or the timed CPU baseline, never as a product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libcg_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_cg.so")

LSHAPE, RECT, LSHAPE_ANY = 0, 1, 3
STOP_NAMES = ["ITERATIONS", "PRECISION", "RESIDUAL", "EXACT_ERROR", "INTERRUPTED"]

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(reference_root: str = "/root/reference") -> None:
    """Compile the restatement and, when the reference sources are present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir(os.path.join(reference_root, "solver")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={reference_root}"])


class _Grid(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
                ("d", C.c_double), ("hx", C.c_double), ("hy", C.c_double), ("xk", C.c_double),
                ("yk", C.c_double), ("A", C.c_double), ("kind", C.c_int)]


class MfInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("r0_norm", C.c_double),
                ("r_norm", C.c_double), ("seconds", C.c_double)]


class MsgInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("stop_reason", C.c_int),
                ("r_max", C.c_double), ("dx_max", C.c_double), ("err_max", C.c_double),
                ("r_l2", C.c_double), ("seconds", C.c_double), ("n_callbacks", C.c_int)]


def _opt(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


class Oracle:
    """Plain-C restatement bound to one grid. Argument order (m, n, a, b, c, d) as GridSystem."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build()
            L = C.CDLL(ORACLE_SO)
            L.cgo_grid_init.restype = C.c_int
            L.cgo_size.restype = C.c_long
            L.cgo_index.restype = C.c_long
            L.cgo_csr_nnz.restype = C.c_long
            cls._lib = L
        return cls._lib

    def __init__(self, m, n, a=0.0, b=1.0, c=0.0, d=1.0, kind=LSHAPE):
        self.L = self.lib()
        self.g = _Grid()
        rc = self.L.cgo_grid_init(C.byref(self.g), int(m), int(n), C.c_double(a), C.c_double(b),
                                  C.c_double(c), C.c_double(d), int(kind))
        if rc != 0:
            raise ValueError(f"grid (n={n}, m={m}, kind={kind}) is outside the reference's valid domain")
        self.N = int(self.L.cgo_size(C.byref(self.g)))

    def index(self, x, y):
        return int(self.L.cgo_index(C.byref(self.g), int(x), int(y)))

    def node(self, idx):
        x, y = C.c_int(), C.c_int()
        self.L.cgo_node(C.byref(self.g), C.c_long(idx), C.byref(x), C.byref(y))
        return x.value, y.value

    def rhs(self):
        out = np.empty(self.N)
        self.L.cgo_rhs(C.byref(self.g), _opt(out))
        return out

    def true_solution(self):
        out = np.empty(self.N)
        self.L.cgo_true_solution(C.byref(self.g), _opt(out))
        return out

    def node_coords(self):
        xs, ys = np.empty(self.N), np.empty(self.N)
        self.L.cgo_node_coords(C.byref(self.g), _opt(xs), _opt(ys))
        return xs, ys

    def apply(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.N)
        self.L.cgo_apply(C.byref(self.g), _opt(x), _opt(y))
        return y

    def apply_nodewise(self, x):
        """The same operator in the reference's node-by-node form (the definition cgo_apply is checked against)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.N)
        self.L.cgo_apply_nodewise(C.byref(self.g), _opt(x), _opt(y))
        return y

    def mf_solve(self, b=None, eps=1e-6, max_it=10000, with_hist=False, snapshots=False, accurate_dots=False):
        """accurate_dots=True: diagnostic long-double summation instead of the reference's sequential fp64."""
        self.L.cgo_set_dot_mode(1 if accurate_dots else 0)
        try:
            return self._mf_solve(b, eps, max_it, with_hist, snapshots)
        finally:
            self.L.cgo_set_dot_mode(0)

    def _mf_solve(self, b, eps, max_it, with_hist, snapshots):
        b = self.rhs() if b is None else np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.N)
        info = MfInfo()
        hist = np.zeros((max_it, 3)) if with_hist else None
        u = self.true_solution() if with_hist else None
        r = np.empty(self.N) if snapshots else None
        p = np.empty(self.N) if snapshots else None
        self.L.cgo_mf_solve(C.byref(self.g), _opt(b), _opt(u), C.c_double(eps), int(max_it), _opt(x),
                            C.byref(info), _opt(hist), int(max_it if with_hist else 0), _opt(r), _opt(p))
        out = dict(x=x, iterations=info.iterations, converged=bool(info.converged), r0_norm=info.r0_norm,
                   r_norm=info.r_norm, seconds=info.seconds)
        if with_hist:
            out["hist"] = hist[: info.iterations]
        if snapshots:
            out["r"], out["p"] = r, p
        return out

    def mf_solve_single(self, b=None, eps=1e-6, max_it=10000):
        """The same solve with alpha from the single-reduction CG recurrence (not a reference function; see cg_oracle.c)."""
        b = self.rhs() if b is None else np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.N)
        info = MfInfo()
        self.L.cgo_mf_solve_single(C.byref(self.g), _opt(b), C.c_double(eps), int(max_it), _opt(x), C.byref(info))
        return dict(x=x, iterations=info.iterations, converged=bool(info.converged), r0_norm=info.r0_norm,
                    r_norm=info.r_norm, seconds=info.seconds)

    def csr(self):
        nnz = int(self.L.cgo_csr_nnz(C.byref(self.g)))
        row_map = np.empty(self.N + 1, dtype=np.int32)
        entries = np.empty(nnz, dtype=np.int32)
        values = np.empty(nnz)
        self.L.cgo_csr_assemble(C.byref(self.g), _opt(row_map), _opt(entries), _opt(values))
        return row_map, entries, values

    def spmv(self, csr, x):
        row_map, entries, values = csr
        y = np.empty(len(row_map) - 1)
        self.L.cgo_spmv(C.c_long(len(row_map) - 1), _opt(row_map), _opt(entries), _opt(values),
                        _opt(np.ascontiguousarray(x, dtype=np.float64)), _opt(y))
        return y

    def msg_solve(self, csr=None, b=None, u=None, eps_p=1e-6, eps_r=1e-6, eps_e=-1.0, max_it=10000,
                  cb_cap=0, accurate_dots=False):
        """accurate_dots=True: diagnostic long-double summation instead of the reference's sequential fp64."""
        self.L.cgo_set_dot_mode(1 if accurate_dots else 0)
        try:
            return self._msg_solve(csr, b, u, eps_p, eps_r, eps_e, max_it, cb_cap)
        finally:
            self.L.cgo_set_dot_mode(0)

    def _msg_solve(self, csr, b, u, eps_p, eps_r, eps_e, max_it, cb_cap):
        row_map, entries, values = self.csr() if csr is None else csr
        b = self.rhs() if b is None else np.ascontiguousarray(b, dtype=np.float64)
        nrows = len(row_map) - 1
        x = np.empty(nrows)
        info = MsgInfo()
        log = np.zeros((cb_cap, 4)) if cb_cap else None
        self.L.cgo_msg_solve(C.c_long(nrows), _opt(row_map), _opt(entries), _opt(values), _opt(b), _opt(u),
                             C.c_double(eps_p), C.c_double(eps_r), C.c_double(eps_e), int(max_it), _opt(x),
                             C.byref(info), _opt(log), int(cb_cap))
        out = dict(x=x, iterations=info.iterations, converged=bool(info.converged),
                   stop_reason=STOP_NAMES[info.stop_reason], r_max=info.r_max, dx_max=info.dx_max,
                   err_max=info.err_max, r_l2=info.r_l2, seconds=info.seconds, n_callbacks=info.n_callbacks)
        if cb_cap:
            out["callbacks"] = log[: min(cb_cap, info.n_callbacks)]
        return out

    def msg_solve_single(self, b=None, u=None, eps_p=1e-6, eps_r=1e-6, eps_e=-1.0, max_it=10000, cb_cap=0):
        """MSGSolver's rules with alpha from the single-reduction recurrence on the matrix-free operator (not a reference
        function; see cg_oracle.c)."""
        b = self.rhs() if b is None else np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.N)
        info = MsgInfo()
        log = np.zeros((cb_cap, 4)) if cb_cap else None
        self.L.cgo_msg_solve_single(C.byref(self.g), _opt(b), _opt(u), C.c_double(eps_p), C.c_double(eps_r),
                                    C.c_double(eps_e), int(max_it), _opt(x), C.byref(info), _opt(log), int(cb_cap))
        out = dict(x=x, iterations=info.iterations, converged=bool(info.converged),
                   stop_reason=STOP_NAMES[info.stop_reason], r_max=info.r_max, dx_max=info.dx_max,
                   err_max=info.err_max, r_l2=info.r_l2, seconds=info.seconds, n_callbacks=info.n_callbacks)
        if cb_cap:
            out["callbacks"] = log[: min(cb_cap, info.n_callbacks)]
        return out


class _RefMsgInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("stop_reason", C.c_int),
                ("final_residual_norm", C.c_double), ("final_error_norm", C.c_double),
                ("final_precision", C.c_double), ("seconds", C.c_double), ("n_callbacks", C.c_int)]


class _RefDirichletOut(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("residual_norm", C.c_double),
                ("error_norm", C.c_double), ("stop_reason", C.c_char * 256), ("size", C.c_int),
                ("seconds", C.c_double)]


class Reference:
    """The unmodified reference classes behind oracle/ref_driver.cpp."""

    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(REF_SO)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.ref_mf_create.restype = C.c_void_p
            L.ref_grid_create.restype = C.c_void_p
            cls._lib = L
        return cls._lib

    # ---- matrix-free
    class MatrixFree:
        def __init__(self, m, n, a=0.0, b=1.0, c=0.0, d=1.0):
            self.L = Reference.lib()
            self.h = C.c_void_p(self.L.ref_mf_create(int(m), int(n), C.c_double(a), C.c_double(b),
                                                     C.c_double(c), C.c_double(d)))
            self.N = int(self.L.ref_mf_size(self.h))

        def __del__(self):
            if getattr(self, "h", None):
                self.L.ref_mf_destroy(self.h)
                self.h = None

        def rhs(self):
            out = np.empty(self.N)
            self.L.ref_mf_rhs(self.h, _opt(out))
            return out

        def true_solution(self):
            out = np.empty(self.N)
            self.L.ref_mf_true_solution(self.h, _opt(out))
            return out

        def apply(self, x):
            x = np.ascontiguousarray(x, dtype=np.float64)
            y = np.empty(self.N)
            self.L.ref_mf_apply(self.h, _opt(x), _opt(y))
            return y

        def solve(self, b=None, eps=1e-6, max_it=10000, with_hist=False):
            x = np.empty(self.N)
            conv, secs = C.c_int(), C.c_double()
            hist = np.zeros((max_it, 3)) if with_hist else None
            bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
            its = self.L.ref_mf_solve(self.h, _opt(bb), C.c_double(eps), int(max_it), _opt(x), C.byref(conv),
                                      _opt(hist), int(max_it if with_hist else 0), C.byref(secs))
            out = dict(x=x, iterations=int(its), converged=bool(conv.value), seconds=secs.value)
            if with_hist:
                out["hist"] = hist[:its]
            return out

    # ---- assembled
    class Grid:
        def __init__(self, m, n, a=0.0, b=1.0, c=0.0, d=1.0):
            self.L = Reference.lib()
            self.h = C.c_void_p(self.L.ref_grid_create(int(m), int(n), C.c_double(a), C.c_double(b),
                                                       C.c_double(c), C.c_double(d)))
            self.N = int(self.L.ref_grid_rows(self.h))
            self.nnz = int(self.L.ref_grid_nnz(self.h))

        def __del__(self):
            if getattr(self, "h", None):
                self.L.ref_grid_destroy(self.h)
                self.h = None

        def csr(self):
            row_map = np.empty(self.N + 1, dtype=np.int32)
            entries = np.empty(self.nnz, dtype=np.int32)
            values = np.empty(self.nnz)
            self.L.ref_grid_csr(self.h, _opt(row_map), _opt(entries), _opt(values))
            return row_map, entries, values

        def rhs(self):
            out = np.empty(self.N)
            self.L.ref_grid_rhs(self.h, _opt(out))
            return out

        def true_solution(self):
            out = np.empty(self.N)
            self.L.ref_grid_true_solution(self.h, _opt(out))
            return out

        def coords(self):
            xs, ys = np.empty(self.N), np.empty(self.N)
            self.L.ref_grid_coords(self.h, _opt(xs), _opt(ys))
            return xs, ys

        def msg_solve(self, eps_p=1e-6, eps_r=1e-6, eps_e=-1.0, max_it=10000, with_true=True, cb_cap=0):
            x = np.empty(self.N)
            info = _RefMsgInfo()
            log = np.zeros((cb_cap, 4)) if cb_cap else None
            self.L.ref_msg_solve(self.h, C.c_double(eps_p), C.c_double(eps_r), C.c_double(eps_e), int(max_it),
                                 int(bool(with_true)), _opt(x), C.byref(info), _opt(log), int(cb_cap))
            out = dict(x=x, iterations=info.iterations, converged=bool(info.converged),
                       stop_reason=STOP_NAMES[info.stop_reason], r_max=info.final_residual_norm,
                       err_max=info.final_error_norm, dx_max=info.final_precision, seconds=info.seconds,
                       n_callbacks=info.n_callbacks)
            if cb_cap:
                out["callbacks"] = log[: min(cb_cap, info.n_callbacks)]
            return out

    @staticmethod
    def dirichlet_solve(n, m, a=0.0, b=1.0, c=0.0, d=1.0, eps_p=1e-6, eps_r=1e-6, eps_e=1e-6, max_iter=10000,
                        use_p=True, use_r=True, use_e=False):
        """DirichletSolver(n, m, ...).solve() -> SolverResults as a dict (dirichlet_solver.cpp:61-131)."""
        L = Reference.lib()
        g = Reference.Grid(m, n, a, b, c, d)
        N = g.N
        arrs = {k: np.empty(N) for k in ("solution", "true_solution", "residual", "error", "x_coords",
                                         "y_coords")}
        out = _RefDirichletOut()
        L.ref_dirichlet_solve(int(n), int(m), C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(d),
                              C.c_double(eps_p), C.c_double(eps_r), C.c_double(eps_e), int(max_iter),
                              int(use_p), int(use_r), int(use_e), *[_opt(arrs[k]) for k in
                                                                    ("solution", "true_solution", "residual",
                                                                     "error", "x_coords", "y_coords")],
                              C.byref(out))
        arrs.update(iterations=out.iterations, converged=bool(out.converged), residual_norm=out.residual_norm,
                    error_norm=out.error_norm, stop_reason=out.stop_reason.decode("utf-8"), size=out.size,
                    seconds=out.seconds)
        return arrs
