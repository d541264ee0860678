"""Regenerates tests/golden/*.npz. Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Three fixture files:
* reference_scripts.npz - the known-answer data the reference itself ships: the 16x16 matrix of
  check.py:4-19, the RHS of check_debug.py:36 and every vector/scalar of py_debug.txt:5-15, parsed
  from the files where they lie (nothing is retyped by hand).
* reference_outputs.npz - outputs of the UNMODIFIED reference classes (oracle/_ref/libref_cg.so, built by
  oracle/Makefile from /root/reference/solver/*.cpp) on small grids: rhs, true solution, apply() on a
  seeded random vector, MatrixFreeSolver / MSGSolver / DirichletSolver solves.
* reference_outputs_msg.npz - more MSGSolver runs of the unmodified reference: the exact-error rule (stop reason
  EXACT_ERROR), a solve without a true solution, the 64 x 64 grid on [0,1]^2 and an iteration cap.
The GPU box has no /root/reference; the -m gpu tests read only these fixtures.
"""
import ast
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference, build  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def parse_scripts():
    src = open(os.path.join(REF, "check.py"), encoding="utf-8").read()
    mat_txt = re.search(r"A = np\.array\(\s*(\[.*?\])\s*\)\s*\n", src, re.S).group(1).replace("\\\n", "")
    A = np.array(ast.literal_eval(mat_txt), dtype=np.float64)
    dbg = open(os.path.join(REF, "check_debug.py"), encoding="utf-8").read()
    b_txt = re.search(r"^b = (\[.*?\])$", dbg, re.M).group(1)
    b = np.array(ast.literal_eval(b_txt), dtype=np.float64)
    out = dict(check_matrix=A, check_debug_rhs=b)
    names = {"A @ h0": "A_h0", "alpha0": "alpha0", "x1": "x1", "r1": "r1", "beta0": "beta0", "h1": "h1",
             "A @ h1": "A_h1", "alpha1": "alpha1", "x2": "x2", "r2": "r2", "h0": "h0"}
    for line in open(os.path.join(REF, "py_debug.txt"), encoding="utf-8"):
        m = re.match(r"^(A @ h0|A @ h1|alpha0|alpha1|beta0|x1|x2|r1|r2|h0|h1) = (.*)$", line.strip())
        if m:
            out["py_" + names[m.group(1)]] = np.array(ast.literal_eval(m.group(2)), dtype=np.float64)
    return out


def reference_outputs():
    out = {}
    rng = np.random.default_rng(0)
    # --- matrix-free path: (n, domain) cases; SURVEY 8c pins 13 / 88 / 352 / 362 iterations
    for n, (a, b) in [(6, (1.0, 2.0)), (30, (1.0, 2.0)), (128, (0.0, 1.0)), (128, (1.0, 2.0)), (64, (0.0, 1.0))]:
        tag = f"mf_n{n}_a{int(a)}"
        mf = Reference.MatrixFree(n, n, a, b, a, b)
        out[tag + "_rhs"] = mf.rhs()
        out[tag + "_true"] = mf.true_solution()
        xin = rng.standard_normal(mf.N)
        out[tag + "_apply_in"] = xin
        out[tag + "_apply_out"] = mf.apply(xin)
        s = mf.solve(eps=1e-8, max_it=10000, with_hist=(n <= 30))
        out[tag + "_x"] = s["x"]
        out[tag + "_iters"] = np.array([s["iterations"], int(s["converged"])])
        if n <= 30:
            out[tag + "_hist"] = s["hist"]
        s2 = mf.solve(eps=1e-8, max_it=2)
        out[tag + "_x2"] = s2["x"]
    # --- assembled path
    for n, (a, b) in [(6, (1.0, 2.0)), (30, (1.0, 2.0)), (128, (0.0, 1.0))]:
        tag = f"grid_n{n}_a{int(a)}"
        g = Reference.Grid(n, n, a, b, a, b)
        row_map, entries, values = g.csr()
        xs, ys = g.coords()
        out[tag + "_rhs"] = g.rhs()
        out[tag + "_true"] = g.true_solution()
        out[tag + "_xs"], out[tag + "_ys"] = xs, ys
        if n <= 30:
            out[tag + "_row_map"], out[tag + "_entries"], out[tag + "_values"] = row_map, entries, values
        out[tag + "_shape"] = np.array([g.N, g.nnz])
        cases = {"pr": dict(eps_p=1e-6, eps_r=1e-6), "r": dict(eps_p=-1.0, eps_r=1e-6)} if n <= 30 else \
                {"pr": dict(eps_p=1e-8, eps_r=1e-8), "r": dict(eps_p=-1.0, eps_r=1e-8)}
        for cname, kw in cases.items():
            s = g.msg_solve(eps_e=-1.0, max_it=10000, with_true=True, cb_cap=256, **kw)
            out[f"{tag}_msg_{cname}_x"] = s["x"]
            out[f"{tag}_msg_{cname}_info"] = np.array(
                [s["iterations"], int(s["converged"]), ["ITERATIONS", "PRECISION", "RESIDUAL", "EXACT_ERROR",
                                                        "INTERRUPTED"].index(s["stop_reason"]),
                 s["r_max"], s["dx_max"], s["err_max"]])
            out[f"{tag}_msg_{cname}_cb"] = s["callbacks"]
        if n == 6:
            s = g.msg_solve(eps_p=1e-9, eps_r=1e-9, eps_e=1e-9, max_it=2, with_true=True)  # solver/main.cpp:601-602
            out[tag + "_msg_2it_x"] = s["x"]
            out[tag + "_msg_2it_rmax"] = np.array([s["r_max"]])
    # --- facade
    d = Reference.dirichlet_solve(30, 30, 1.0, 2.0, 1.0, 2.0)  # GUI defaults, mainwindow.cpp:112-125
    for k in ("solution", "true_solution", "residual", "error", "x_coords", "y_coords"):
        out["dirichlet_n30_" + k] = d[k]
    out["dirichlet_n30_info"] = np.array([d["iterations"], int(d["converged"]), d["residual_norm"], d["error_norm"]])
    out["dirichlet_n30_stop_reason"] = np.frombuffer(d["stop_reason"].encode("utf-8"), dtype=np.uint8)
    return out


STOP = ["ITERATIONS", "PRECISION", "RESIDUAL", "EXACT_ERROR", "INTERRUPTED"]


def reference_outputs_msg():
    """MSGSolver::solve branches the first file does not reach (msg_solver.cpp:64-72,132-139,158-162,80)."""
    out = {}
    cases = [("n30_a1_exact", 30, (1.0, 2.0), dict(eps_p=-1.0, eps_r=-1.0, eps_e=5e-3, max_it=10000, with_true=True)),
             ("n30_a1_exact_first", 30, (1.0, 2.0), dict(eps_p=1e-12, eps_r=1e-12, eps_e=1e-2, max_it=10000, with_true=True)),
             ("n30_a1_nou", 30, (1.0, 2.0), dict(eps_p=1e-7, eps_r=-1.0, eps_e=1e-3, max_it=10000, with_true=False)),
             ("n64_a0_pr", 64, (0.0, 1.0), dict(eps_p=1e-8, eps_r=1e-8, eps_e=-1.0, max_it=10000, with_true=True)),
             ("n64_a0_r", 64, (0.0, 1.0), dict(eps_p=-1.0, eps_r=1e-8, eps_e=-1.0, max_it=10000, with_true=True)),
             ("n64_a0_cap", 64, (0.0, 1.0), dict(eps_p=1e-30, eps_r=1e-30, eps_e=-1.0, max_it=150, with_true=True))]
    for tag, n, (a, b), kw in cases:
        g = Reference.Grid(n, n, a, b, a, b)
        s = g.msg_solve(cb_cap=256, **kw)
        out[f"msg_{tag}_x"] = s["x"]
        out[f"msg_{tag}_info"] = np.array([s["iterations"], int(s["converged"]), STOP.index(s["stop_reason"]), s["r_max"],
                                           s["dx_max"], s["err_max"]])
        out[f"msg_{tag}_cb"] = s["callbacks"]
        out[f"msg_{tag}_params"] = np.array([n, a, b, kw["eps_p"], kw["eps_r"], kw["eps_e"], kw["max_it"], int(kw["with_true"])])
    return out


if __name__ == "__main__":
    build(REF)
    np.savez_compressed(os.path.join(OUT, "reference_outputs_msg.npz"), **reference_outputs_msg())
    np.savez_compressed(os.path.join(OUT, "reference_scripts.npz"), **parse_scripts())
    np.savez_compressed(os.path.join(OUT, "reference_outputs.npz"), **reference_outputs())
    for f in ("reference_scripts.npz", "reference_outputs.npz", "reference_outputs_msg.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
