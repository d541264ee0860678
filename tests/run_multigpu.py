"""Launched under torchrun (one process per GPU) by tests/test_multigpu.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multigpu.py
Row-slab sharded solve through the C ABI on every rank; rank 0 assembles the slabs and checks them against the
CPU oracle (same grid, rhs, x0, tolerance): iterations +-1, solution within 1e-10 relative."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterative_solvers_b200 import capi  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


FULL_CASES = [(64, 0, 1e-8, "mf"), (256, 0, 1e-8, "mf"), (1100, 0, 1e-6, "mf"), (333, 1, 1e-8, "mf"),
              (256, 0, 1e-8, "mf-hs2"), (128, 0, 1e-8, "msg"), (128, 0, 1e-8, "msg0"), (128, 0, 1e-8, "cb"),
              (256, 0, 1e-8, "mf-2s"), (333, 1, 1e-8, "mf-2s"), (600, 0, 1e-6, "mf-2s"), (256, 0, 1e-8, "mf-w"),
              (1100, 0, 1e-6, "mf-w"), (256, 0, 1e-8, "stop"), (256, 0, 1e-8, "stop-2s"), (128, 0, 1e-8, "msg-2s"),
              (128, 0, 1e-8, "msg0-2s"), (256, 0, 1e-7, "msg-w"), (500, 0, 1e-5, "msg"), (333, 1, 1e-7, "msg0"),
              (256, 0, 1e-8, "stop-msg")]
# bench.py runs these in-process before its timed loop at N > 1 (the oracle solves take ~20 s of rank 0's host time)
BENCH_CASES = [(64, 0, 1e-8, "mf"), (256, 0, 1e-8, "mf"), (500, 0, 1e-5, "mf"), (333, 1, 1e-8, "mf"),
               (256, 0, 1e-8, "mf-hs2"), (128, 0, 1e-8, "msg"), (128, 0, 1e-8, "msg0"), (128, 0, 1e-8, "cb"),
               (256, 0, 1e-8, "mf-2s"), (333, 1, 1e-8, "mf-2s"), (256, 0, 1e-8, "mf-w"), (500, 0, 1e-5, "mf-w"),
               (256, 0, 1e-8, "stop"), (256, 0, 1e-8, "stop-2s"), (128, 0, 1e-8, "msg-2s"), (256, 0, 1e-7, "msg-w"),
               (500, 0, 1e-5, "msg"), (256, 0, 1e-8, "stop-msg")]


def run_cases(rank, world, local, cases, log=print):
    """Every rank calls this inside an initialised NCCL process group. Kinds: "mf" = the default iteration (single sweep:
    two halo rows per side over peer memory, one publish-and-wait per iteration), "mf-2s" = the two-sweep iteration,
    "mf-hs2" = two sweeps with 2-row stages (the launch shapes are read when the plan is created), "mf-w" = the single
    sweep in the wide geometry of large slabs (840-column strips, one CTA per SM) forced onto a small grid, "stop" /
    "stop-2s" / "stop-msg" = an interrupt raised on rank 0 ONLY must end the solve on every rank at the same iteration (single
    sweep, two sweeps, single sweep under the max-norm rules), "msg" / "msg0" = the max-norm rules with / without a true
    solution (single sweep on peer-memory plans: sums AND maxima cross the ranks in one publish-and-wait; the callback
    records are compared with the oracle's), "msg-2s" / "msg0-2s" = the same as dot sweep + update sweep, "msg-w" = in the wide
    geometry, "cb" = the per-iteration report callback.
    Returns {"cases", "ok", "max_rel", "iterations_equal", "failed"} (meaningful on rank 0)."""

    def fresh_comm_id():  # an NCCL unique id bootstraps exactly one communicator: one per plan
        blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            blob = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(blob, src=0)
        return bytes(blob.cpu().tolist())

    failures, max_rel, its_equal, refs = [], 0.0, True, {}
    hs2 = {"B200CG_SHAPE_DOT": "0", "B200CG_SHAPE_UPD": "0", "B200CG_SHAPE_NOX": "0"}
    for n, domain, eps, kind in cases:
        o = Oracle(n, n, 0.0, 1.0, 0.0, 1.0, domain)
        b, u = o.rhs(), o.true_solution()
        if kind == "mf-hs2":
            os.environ.update(hs2)
        if kind in ("mf-w", "msg-w"):
            os.environ["B200CG_FUSED_CW"] = "14"
        plan = capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, domain=domain, device=local, rank=rank, world=world,
                         comm_id=fresh_comm_id())
        for k in list(hs2) + ["B200CG_FUSED_CW"]:
            os.environ.pop(k, None)
        lo, hi = plan.lo, plan.hi
        got_cb = []

        def reference(key, fn):  # the oracle is serial host code: only rank 0 needs it, once per distinct solve
            if rank != 0:
                return None
            if key not in refs:
                refs[key] = fn()
            return refs[key]

        if kind in ("stop", "stop-2s", "stop-msg"):
            # requestStop on one rank: the request travels with the per-iteration reductions (kernels poll the mapped flag
            # every 16th iteration), so all ranks report INTERRUPTED after the same 16 iterations
            import ctypes

            flag = ctypes.c_int(1 if rank == 0 else 0)
            rules = dict(rule=capi.RULE_MAXNORM, eps_p=-1.0, eps_r=1e-300) if kind == "stop-msg" else dict(eps_rel=1e-30)
            x, info = plan.solve(b=b[lo:hi], max_it=100000, iters_per_graph=100, stop_flag=flag,
                                 single_sweep=2 if kind == "stop-2s" else 0, **rules)
            plan.close()
            seen = [None] * world
            dist.all_gather_object(seen, (info["iterations"], info["stop_reason"], info["converged"]))
            ok = all(s == (16, "INTERRUPTED", False) for s in seen)
            if rank == 0:
                log(f"[multigpu] n={n} domain={domain} {kind}: per-rank (iterations, stop reason, converged) = {seen} -> "
                    f"{'ok' if ok else 'FAIL'}")
                if not ok:
                    failures.append((n, domain, kind))
            continue
        # the GPU solve first, on all ranks together; rank 0 computes the reference afterwards while the others wait in
        # the gather below (inside a solve a rank waits at most 20 s for its peers)
        if kind in ("mf", "mf-hs2", "mf-2s", "mf-w"):
            two = kind in ("mf-hs2", "mf-2s")
            x, info = plan.solve(b=b[lo:hi], eps_rel=eps, max_it=20000, single_sweep=2 if two else 0)
            if info["single_sweep"] != (0 if two else 1):
                failures.append((n, domain, kind, "wrong iteration scheme"))
            ref = reference((n, domain, eps, "mf"), lambda: o.mf_solve(b=b, eps=eps, max_it=20000))
        elif kind == "cb":
            x, info = plan.solve(b=b[lo:hi], u=u[lo:hi], eps_rel=eps, max_it=20000,
                                 callback=lambda it, p, r, e: got_cb.append((it, p, r, e)))
            ref = reference((n, domain, eps, "cb"), lambda: o.mf_solve(b=b, eps=eps, max_it=20000, with_hist=True))
        else:  # the max-norm rules: "msg", "msg-2s", "msg-w" with the true solution, "msg0", "msg0-2s" without
            two, with_u = kind.endswith("-2s"), not kind.startswith("msg0")
            x, info = plan.solve(b=b[lo:hi], u=u[lo:hi] if with_u else None, rule=capi.RULE_MAXNORM, eps_p=eps, eps_r=eps,
                                 max_it=20000, single_sweep=2 if two else 0,
                                 callback=lambda it, p, r, e: got_cb.append((it, p, r, e)))
            if info["single_sweep"] != (0 if two or not info["peer_exchange"] else 1):
                failures.append((n, domain, kind, "wrong iteration scheme"))
            ref = reference((n, domain, eps, "msg", with_u),
                            lambda: o.msg_solve(b=b, u=u if with_u else None, eps_p=eps, eps_r=eps, max_it=20000, cb_cap=512))
        v = np.random.default_rng(n).standard_normal(o.N)
        y = plan.apply(v[lo:hi])
        res, _ = plan.postprocess(want_error=False)
        plan.close()
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, x, y, res))
        if rank == 0:
            xg, yg, rg = np.empty(o.N), np.empty(o.N), np.empty(o.N)
            for plo, phi, px, py, pr in parts:
                xg[plo:phi], yg[plo:phi], rg[plo:phi] = px, py, pr
            rel = float(np.max(np.abs(xg - ref["x"])) / np.max(np.abs(ref["x"])))
            max_rel = max(max_rel, rel)
            its_equal = its_equal and info["iterations"] == ref["iterations"]
            ok = abs(info["iterations"] - ref["iterations"]) <= 1 and rel < 1e-10
            ok = ok and np.array_equal(yg, o.apply(v))
            ok = ok and np.max(np.abs(rg - (o.apply(xg) - b))) <= 1e-12 * np.max(np.abs(b))
            if kind.startswith("msg"):
                ok = ok and info["stop_reason"] == ref["stop_reason"]
                if info["iterations"] == ref["iterations"]:  # callback cadence and values (it 0, 1, every 100, final)
                    got, cbr = np.array(got_cb), ref["callbacks"]
                    ok = ok and got.shape == cbr.shape and np.array_equal(got[:, 0], cbr[:, 0])
                    ok = ok and np.allclose(got[1:, 1:3], cbr[1:, 1:3], rtol=1e-3, atol=1e-10 * np.max(np.abs(b)))
                    ok = ok and abs(info["dx_max"] - ref["dx_max"]) <= 1e-6 * ref["dx_max"]
                    if not kind.startswith("msg0"):
                        ok = ok and abs(info["err_max"] - ref["err_max"]) <= 1e-9 * ref["err_max"] + 1e-10 * np.max(np.abs(u))
            if kind == "cb":
                got = np.array(got_cb)
                ok = ok and len(got) == len(ref["hist"]) and np.all(
                    np.abs(got[:, 1:] - ref["hist"]) <= 1e-9 * np.max(np.abs(ref["hist"]), axis=0) + 1e-9 * np.abs(ref["hist"]))
            log(f"[multigpu] n={n} domain={domain} {kind} (peer_exchange={info['peer_exchange']}, "
                f"single_sweep={info['single_sweep']}): iterations {info['iterations']} vs {ref['iterations']}, "
                f"max rel diff {rel:.2e} -> {'ok' if ok else 'FAIL'}")
            if not ok:
                failures.append((n, domain, kind))
    dist.barrier()
    return {"cases": len(cases), "ok": not failures, "max_rel": max_rel, "iterations_equal": bool(its_equal),
            "failed": [list(map(str, f)) for f in failures], "world": world,
            "bar": "iterations within +-1 of the CPU oracle, solution within 1e-10 relative, apply bit-equal"}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run_cases(rank, world, local, FULL_CASES, log=lambda m: print(m, flush=True))
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU_OK" if out["ok"] else f"MULTIGPU_FAIL {out['failed']}", flush=True)
        sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
