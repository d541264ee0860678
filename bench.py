#!/usr/bin/env python
"""bench.py - CG GDOF-iter/s of the matrix-free Dirichlet-Poisson solve (BASELINE.json metric).

A "step" is one fixed-iteration solve (default 500 CG iterations) of the L-shaped
Dirichlet system on an n x n grid through the C ABI (b200cg_solve), i.e. what MatrixFreeSolver::solve does.
  value : whole-job DOF-iterations/s with the right-hand side already resident in HBM and the solution left
          there; timed on the device (CUDA events of the library's solve stream), max over ranks.
  e2e   : the same solve with HOST buffers (pinned): H2D of b and D2H of x inside the timed region.
  roofline / hbm_gbs_actual : REAL traffic (algorithmic bytes of the kernels that ran: 40 B per unknown-iteration for the
          default single-sweep iteration) against the measured copy peak; model_80B: the same work expressed in SURVEY
          8d's 80-byte store-Ap model (a work figure - it may exceed the physical bandwidth).
N > 1 (torchrun, one process per GPU): row slabs of one larger grid, n_G = even(round(n * sqrt(G))), so the
unknowns per GPU stay fixed (weak scaling, BASELINE.json configs[4]); halo rows and scalar reductions go over NVLink
peer memory (NCCL as fallback) inside the library. --scaling strong shards the fixed --grid-n grid instead (configs[2]).
Before the timed loop every N > 1 run checks 10 sharded solves against the CPU oracle (multi_gpu_parity).
Workload at N = 1: the 16384^2 grid - the configuration BASELINE.json quotes its metric and target on ("a 16384^2 fp64
Dirichlet CG solve at >= 70 % of B200 HBM bandwidth per GPU"; configs[2] and [4] at one GPU); it fits one GPU (17 GB).
Extras of the default N = 1 line: two_sweep_extra (the two-sweep iteration on the same grid) and csr_extra
(configs[3]: assembled CSR CG at 8192^2 beside the matrix-free iteration). configs[1] (4096^2) and configs[3] as lines
of their own: --grid-n 4096 / --op csr --grid-n 8192.

--impl reference times the reference's own CPU solver (unmodified sources compiled into oracle/_ref, else the
C port in oracle/) on a bounded sample of the same workload, on rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cg_gdof_iter_per_s"
UNIT = "GDOF-iter/s"
BYTES_MODEL = 80.0        # algorithmic bytes per DOF-iteration of the store-Ap formulation (a work MODEL: the
                          # implementation moves 40 - 56 B, so "GB/s at 80 B" can exceed the physical bandwidth)
BYTES_CSR = 152.0         # SURVEY 8d, assembled path: 12 nnz + 4 (N+1) + 88 N per iteration with nnz -> 5 N
BYTES_UPD = 48.0          # update-phase kernel touching x: r, p_old, x in; x, r, p out
BYTES_UPD_NOX = 32.0      # update-phase kernel of an even iteration under x-deferral: r, p_old in; r, p out
BYTES_DOT = 16.0          # dot-phase kernel: r, p_old in
NOMINAL_HBM_GBS = 8000.0  # BASELINE.json "vs 8 TB/s peak"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--grid-n", dest="n", type=int, default=16384, help="grid intervals per side at 1 GPU")
    ap.add_argument("--iters", type=int, default=500, help="CG iterations per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): unknowns per GPU fixed; strong: the --n grid sharded over all GPUs "
                         "(BASELINE.json configs[2])")
    ap.add_argument("--domain", default="lshape", choices=["lshape", "rect"])
    ap.add_argument("--op", default="mf", choices=["mf", "csr"])
    ap.add_argument("--single-sweep", type=int, default=0, choices=[0, 1, 2],
                    help="b200cg_params.single_sweep: 0 = plan default (one sweep per iteration, Chronopoulos-Gear alpha), "
                         "2 = the two-sweep iteration (alpha = r.r / p.Ap)")
    ap.add_argument("--no-extras", "--no-single-sweep-extra", dest="no_extras", action="store_true",
                    help="skip the extra legs (two-sweep iteration on the same grid, assembled-CSR CG at 8192^2)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the in-run sharded parity cases")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the process to its GPU's NUMA node")
    ap.add_argument("--tile-rows", type=int, default=0)
    ap.add_argument("--iters-per-graph", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-n", type=int, default=4096)
    ap.add_argument("--cpu-sample-iters", type=int, default=5)
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(single_sweep):
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    ncu --set full capture of the 16384^2 workload, if any."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        return json.load(open(path)).get("fused_x2_dram_bytes_per_launch" if single_sweep else "upd_kernel_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.proc = None
        self.path = None
        self.idx = device_index
        if shutil.which("nvidia-smi"):
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(device_index)], stdout=self.out,
                                         stderr=subprocess.DEVNULL)

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(device_index):
    """Host side of the end-to-end number: run this process - and first-touch the pinned buffers it allocates next - on
    the NUMA node its GPU hangs off, so that at N > 1 the ranks' H2D / D2H copies do not all cross the socket link.
    Returns a short description (or why nothing was done)."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return {"bound": False, "why": "the platform reports no NUMA node for the GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": f"no allowed CPU on NUMA node {node}"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "gpu_pci": bus, "numa_node": node, "cpus": len(cpus)}
    except Exception as exc:
        return {"bound": False, "why": repr(exc)}


def grid_side(n1, gpus, domain):
    if gpus == 1:
        return n1
    n = int(round(n1 * math.sqrt(gpus)))
    if domain == "lshape" and n % 2:
        n += 1
    return n


def workload_name(n, iters, domain, op):
    return (f"{n}x{n} grid, {'L-shaped (reference) domain' if domain == 'lshape' else 'full rectangle'}, "
            f"{'matrix-free' if op == 'mf' else 'assembled CSR'} CG fp64, fixed {iters} iterations per solve, unit square")


# --------------------------------------------------------------------------------------------- reference arm
def time_reference(n, iters):
    """One bounded sample of the workload on the host: MatrixFreeSolver::solve for `iters` iterations."""
    from oracle import oracle

    if oracle.Reference.available():
        mf = oracle.Reference.MatrixFree(n, n, 0.0, 1.0, 0.0, 1.0)
        s = mf.solve(eps=0.0, max_it=iters)
        return mf.N, s["iterations"], s["seconds"], "reference"
    o = oracle.Oracle(n, n, 0.0, 1.0, 0.0, 1.0)
    b = o.rhs()
    s = o.mf_solve(b=b, eps=0.0, max_it=iters, with_hist=True)  # with_hist: the reference's per-iteration reporting work
    return o.N, s["iterations"], s["seconds"], "port"


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    n, iters = min(args.cpu_sample_n, 2048), 3
    for _ in range(args.warmup):
        time_reference(n, iters)
    secs, dofit = 0.0, 0.0
    kind = "reference"
    for _ in range(args.steps):
        N, its, s, kind = time_reference(n, iters)
        secs += s
        dofit += N * its
    value = dofit / secs / 1e9
    sample = (f"MatrixFreeSolver::solve on the {n}x{n} L-shaped grid ({N} unknowns), {iters} iterations per step "
              f"(same operator, rhs and x0 as the {args.n}^2 workload; the reference code is serial)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(grid_side(args.n, args.gpus, args.domain), args.iters, args.domain, args.op),
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- B200 arm
def csr_roofline(spmv_ms, upd_ms, n_local, nnz, peak, peak_src):
    """Assembled path: per iteration 12 nnz + 4 (N + 1) + 88 N algorithmic bytes = 152 B/DOF-it at nnz -> 5 N.
    Dominant kernel: the fused direction update + SpMV + dots (values 8 + columns 4 per non-zero, row_map, z gathered
    once, r, z_old in, z, Az out)."""
    spmv_bytes = 12.0 * nnz + 4.0 * (n_local + 1) + 40.0 * n_local  # r, z_old in; z, Az out; gather counted once
    upd_bytes = 48.0 * n_local                                        # x, z, r, Az in; x, r out
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
    return {"bound": "hbm", "kernel": "csr_spmv_kernel<1> (direction update + SpMV + dots)", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_bytes, "avg_launch_ms": spmv_ms,
            "update_kernel": {"achieved": upd_bytes / (upd_ms * 1e-3) / 1e9 if upd_ms > 0 else None, "avg_launch_ms": upd_ms,
                              "algorithmic_bytes_per_launch": upd_bytes},
            "algorithmic_bytes_per_dof_iter": (spmv_bytes + upd_bytes) / n_local, "model_bytes_per_dof_iter": BYTES_CSR,
            "single_sweep": False, "x_deferral": False}


def csr_leg(capi, device, n, iters, steps, peak, peak_src):
    """BASELINE.json configs[3]: assembled GridSystem CSR SpMV CG on the n x n grid vs the matrix-free iteration."""
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, device=device) as p:
        p.build_rhs()
        nnz = p.assemble_csr()
        out = {"workload": workload_name(n, iters, "lshape", "csr"), "unknowns": p.N, "nnz": nnz}
        for name, kw in (("csr", dict(op=capi.OP_CSR)), ("matrix_free", dict(op=capi.OP_MATRIX_FREE))):
            kw.update(rule=capi.RULE_REL_L2, eps_rel=0.0, max_it=iters, rhs_on_device=True, keep_x_on_device=True)
            for _ in range(2):
                p.solve(**kw)
            ms, its, spmv, upd = 0.0, 0, 0.0, 0.0
            for _ in range(steps):
                _, info = p.solve(**kw)
                ms += info["device_ms"]
                its += info["iterations"]
                spmv += info["dot_kernel_ms"]
                upd += info["upd_kernel_ms"]
            out[name] = {"value": float(p.N) * its / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms / steps}
            if name == "csr":
                out["csr"]["roofline"] = csr_roofline(spmv / steps, upd / steps, p.N, nnz, peak, peak_src)
        out["csr_vs_matrix_free"] = out["csr"]["value"] / out["matrix_free"]["value"]
        return out


def run_b200(args):
    import torch  # plumbing only: process group, barrier, max-over-ranks

    from iterative_solvers_b200 import capi

    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    capi.lib()
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device - libb200cg has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else {"bound": False, "why": "--no-numa-bind"}
    comm_id = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            blob = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(blob, src=0)
        comm_id = bytes(blob.cpu().tolist())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- N > 1: correctness of the sharded solve on THIS build, on the driver's record (tests/run_multigpu.py cases:
    # both iteration schemes, both rule sets, callbacks, 2-row stages) - before and outside the timed region
    parity = None
    if world > 1 and not args.no_parity:
        import importlib.util

        spec = importlib.util.spec_from_file_location("run_multigpu", os.path.join(ROOT, "tests", "run_multigpu.py"))
        mg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mg)
        t_par = time.perf_counter()
        try:
            parity = mg.run_cases(rank, world, local_rank, mg.BENCH_CASES, log=lambda m: print(m, file=sys.stderr, flush=True))
        except Exception as exc:  # a failing check must not cost the timing line; it is reported as not ok
            parity = {"cases": len(mg.BENCH_CASES), "ok": False, "error": repr(exc)}
        parity["seconds"] = time.perf_counter() - t_par

    n = grid_side(args.n, world if args.scaling == "weak" else 1, args.domain)
    domain = capi.DOMAIN_LSHAPE if args.domain == "lshape" else capi.DOMAIN_RECT
    op = capi.OP_MATRIX_FREE if args.op == "mf" else capi.OP_CSR
    plan = capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, domain=domain, device=local_rank, rank=rank, world=world,
                     comm_id=comm_id, tile_rows=args.tile_rows)
    plan.build_rhs()  # synthetic, deterministic: the reference's analytic f and Dirichlet data
    plan_nnz = plan.assemble_csr() if op == capi.OP_CSR else 0
    n_local = plan.n_local
    solve_kw = dict(op=op, rule=capi.RULE_REL_L2, eps_rel=0.0, max_it=args.iters, iters_per_graph=args.iters_per_graph,
                    single_sweep=args.single_sweep)

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        plan.solve(rhs_on_device=True, keep_x_on_device=True, **solve_kw)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, dot_ms, upd_ms, samples, its_done = 0.0, 0, 0.0, 0.0, 0, 0
    upd_even_ms, upd_odd_ms, xdefer, peer_exchange, single_sweep = 0.0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        _, info = plan.solve(rhs_on_device=True, keep_x_on_device=True, **solve_kw)
        dev_ms += info["device_ms"]
        launches += info["kernel_launches"]
        its_done += info["iterations"]
        if info["kernel_samples"]:
            dot_ms += info["dot_kernel_ms"]
            upd_ms += info["upd_kernel_ms"]
            upd_even_ms += info["upd_even_ms"]
            upd_odd_ms += info["upd_odd_ms"]
            xdefer = info["x_deferral"]
            single_sweep = info["single_sweep"]
            peer_exchange = info["peer_exchange"]
            samples += 1
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    total_dofit = float(plan.N) * its_done  # all ranks run the same iteration count
    value = total_dofit / (dev_ms * 1e-3) / 1e9
    launches_all = int(sum_over_ranks(float(launches)))

    # ---- e2e: host buffers through the C ABI. Headline: the steps as ONE queue of right-hand sides (b200cg_solve_batch):
    # every step's b goes H2D from pinned memory and every step's x comes back D2H and is read on the host, all inside
    # the timed region - the copies of neighbouring steps run on copy streams under the current step's iterations.
    # Beside it: the same steps as separate b200cg_solve calls (copies serialised with the solve), on fewer steps.
    e2e = None
    if not args.no_e2e:
        hb = capi.PinnedArray(n_local)
        hx = [capi.PinnedArray(n_local), capi.PinnedArray(n_local)]
        hb.array[:] = plan.get_rhs()
        batch_kw = dict(rule=capi.RULE_REL_L2, eps_rel=0.0, max_it=args.iters, iters_per_graph=args.iters_per_graph,
                        single_sweep=args.single_sweep)
        checks = []

        def serial_leg(steps):
            for _ in range(max(1, min(args.warmup, 2))):
                plan.solve(b=hb.array, x_out=hx[0].array, **solve_kw)
            barrier()
            t0 = time.perf_counter()
            its, dev = 0, 0.0
            for _ in range(steps):
                _, info = plan.solve(b=hb.array, x_out=hx[0].array, **solve_kw)
                its += info["iterations"]
                dev += info["device_ms"]
                checks.append(float(hx[0].array[0] + hx[0].array[-1]))  # the D2H result is read on the host every step
            barrier()
            wall = max_over_ranks((time.perf_counter() - t0) * 1e3)
            return {"value": float(plan.N) * its / (wall * 1e-3) / 1e9, "unit": UNIT, "steps": steps,
                    "ms_per_step": wall / max(steps, 1), "device_ms_per_step": max_over_ranks(dev) / max(steps, 1)}

        def batch_leg(steps):
            outs = [hx[i & 1].array for i in range(steps)]
            plan.solve_batch([hb.array] * 2, [hx[0].array, hx[1].array], **batch_kw)  # warm-up
            barrier()
            t0 = time.perf_counter()
            infos = plan.solve_batch([hb.array] * steps, outs, **batch_kw,
                                     done=lambda i, info: checks.append(float(outs[i][0] + outs[i][-1])))
            barrier()
            wall = max_over_ranks((time.perf_counter() - t0) * 1e3)
            its = sum(info["iterations"] for info in infos)
            return {"value": float(plan.N) * its / (wall * 1e-3) / 1e9, "unit": UNIT, "steps": steps,
                    "ms_per_step": wall / max(steps, 1)}

        serial = serial_leg(args.steps if args.op == "csr" else max(1, min(args.steps, 3)))
        batch, batch_err = None, None
        if args.op == "mf":
            try:
                batch = batch_leg(args.steps)
            except Exception as exc:  # the queue is an addition: its failure must not cost the line
                batch_err = repr(exc)
        head = batch or serial
        e2e = {"value": head["value"], "unit": UNIT,
               "h2d_bytes_per_step": int(sum_over_ranks(float(n_local * 8))),
               "d2h_bytes_per_step": int(sum_over_ranks(float(n_local * 8))),
               "ms_per_step": head["ms_per_step"],
               "mode": ("b200cg_solve_batch: the steps as one queue of right-hand sides, each step's H2D of b and D2H of x "
                        "(pinned host buffers) on copy streams under the neighbouring steps' iterations"
                        if batch else "b200cg_solve per step: H2D of b, solve, D2H of x one after the other"),
               "one_call_per_step": serial, "batch_error": batch_err,
               "timing": "host wall clock around the C-ABI calls, barrier + device synchronize on both sides",
               "checksum": checks[-1] if checks else None, "results_read_on_host": len(checks),
               "host_numa_binding_rank0": numa}
        hb.free()
        for h in hx:
            h.free()

    if rank != 0:
        plan.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    roofline = None
    if samples and op == capi.OP_CSR:
        roofline = csr_roofline(dot_ms / samples, upd_ms / samples, n_local, plan_nnz, peak, peak_src)
    if samples and op == capi.OP_MATRIX_FREE:
        # dominant kernel = the update phase that touches x (48 B/unknown). Under x-deferral even iterations run
        # the lighter variant (32 B/unknown); both and the dot phase (16 B) are listed.
        upd_s = (upd_odd_ms / samples) * 1e-3
        nox_s = (upd_even_ms / samples) * 1e-3
        dot_s = (dot_ms / samples) * 1e-3
        achieved = BYTES_UPD * n_local / upd_s / 1e9
        even_bytes = BYTES_UPD_NOX if xdefer else BYTES_UPD
        iter_s = dot_s + 0.5 * (upd_s + nox_s)
        roofline = {"bound": "hbm",
                    "kernel": ("cg_fused_kernel<F_X2> (single-sweep iteration touching x)" if single_sweep
                               else "cg_stream_kernel<MODE_UPD> (update phase touching x)"), "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(single_sweep) if n == 16384 and world == 1 else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_UPD * n_local,
                    "avg_launch_ms": upd_s * 1e3, "launches_sampled": samples,
                    "x_deferral": bool(xdefer),
                    "update_kernel_even_iterations": {"achieved": even_bytes * n_local / nox_s / 1e9 if nox_s > 0 else None,
                                                      "avg_launch_ms": nox_s * 1e3,
                                                      "algorithmic_bytes_per_launch": even_bytes * n_local},
                    "dot_kernel": None if single_sweep else
                                  {"achieved": BYTES_DOT * n_local / dot_s / 1e9 if dot_s > 0 else None,
                                   "avg_launch_ms": dot_s * 1e3, "algorithmic_bytes_per_launch": BYTES_DOT * n_local},
                    "single_sweep": bool(single_sweep),
                    "algorithmic_bytes_per_dof_iter": (0.0 if single_sweep else BYTES_DOT) + 0.5 * (BYTES_UPD + even_bytes),
                    "kernel_share_of_step": iter_s * 1e3 * (its_done / max(args.steps, 1)) /
                                            (dev_ms / max(args.steps, 1))}
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        N, its, secs, kind = time_reference(args.cpu_sample_n, args.cpu_sample_iters)
        cpu_baseline = {"value": N * its / secs / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                        "host_cores": os.cpu_count(), "seconds": secs,
                        "sample": (f"MatrixFreeSolver::solve, {args.cpu_sample_n}^2 L-shaped grid ({N} unknowns), "
                                   f"{its} iterations, 1 core (the reference code is serial)")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / max(args.steps, 1), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, args.iters, args.domain, args.op), "grid_n": n, "unknowns": plan.N,
                   "unknowns_per_gpu": plan.N / world, "iterations_per_step": args.iters,
                   "iteration": ("single sweep per iteration (default; alpha from the single-reduction CG recurrence)"
                                 if single_sweep else "two sweeps per iteration (alpha = r.r / p.Ap)"),
                   "parallelism": (f"row-slab x{world}, " + ("NVLink peer-memory halo + reductions" if peer_exchange
                                                            else "NCCL halo + all-reduce")) if world > 1 else "single GPU",
                   "l2": "inputs_exceed_l2" if n_local * 8 > 200e6 else "inputs_fit_l2_no_flush",
                   "timing": "CUDA events on the library's solve stream, summed over steps, max over ranks",
                   "wall_ms_per_step": wall_ms / max(args.steps, 1)},
        # real traffic: what the implementation moves (algorithmic bytes of its kernels x value), per GPU
        "hbm_gbs_actual": (value * roofline["algorithmic_bytes_per_dof_iter"] / world) if roofline else None,
        "frac_of_measured_peak_actual": (value * roofline["algorithmic_bytes_per_dof_iter"] / world / peak) if roofline else None,
        "frac_of_8tbs_actual": (value * roofline["algorithmic_bytes_per_dof_iter"] / world / NOMINAL_HBM_GBS) if roofline else None,
        # equivalent work in SURVEY 8d's 80-byte store-Ap MODEL (not traffic: it may exceed the physical bandwidth)
        "model_80B": {"model": "80 B/DOF-it (SURVEY 8d: store-Ap formulation)",
                      "equivalent_gbs_per_gpu": value * BYTES_MODEL / world,
                      "equivalent_frac_of_8tbs": value * BYTES_MODEL / world / NOMINAL_HBM_GBS},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_all, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "multi_gpu_parity": parity, "two_sweep_extra": None, "csr_extra": None,
    }

    # ---- extras (1 GPU, default workload): (1) the same grid with the two-sweep iteration (alpha = r.r / p.Ap, what
    # the max-norm rules and the report callback run); (2) BASELINE.json configs[3]: assembled-CSR CG at 8192^2 beside
    # the matrix-free iteration on that grid. They run last, under a watchdog: whatever happens, the line is printed.
    if world == 1 and op == capi.OP_MATRIX_FREE and args.single_sweep == 0 and not args.no_extras:
        import threading

        finished = threading.Event()

        def watchdog():
            if not finished.wait(240.0):
                line.setdefault("extras_error", "timed out")
                print(json.dumps(line), flush=True)
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()
        try:
            kw = dict(solve_kw, single_sweep=2)
            for _ in range(2):
                plan.solve(rhs_on_device=True, keep_x_on_device=True, **kw)
            s_ms, s_its, s_on = 0.0, 0, 1
            for _ in range(args.steps):
                _, info = plan.solve(rhs_on_device=True, keep_x_on_device=True, **kw)
                s_ms += info["device_ms"]
                s_its += info["iterations"]
                s_on = info["single_sweep"]
            line["two_sweep_extra"] = {
                "value": float(plan.N) * s_its / (s_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": s_ms / max(args.steps, 1),
                "active": not s_on, "algorithmic_bytes_per_dof_iter": 56.0,
                "dot_kernel_ms": info["dot_kernel_ms"], "upd_even_ms": info["upd_even_ms"], "upd_odd_ms": info["upd_odd_ms"],
                "note": "b200cg_params.single_sweep = 2: dot sweep + update sweep per iteration, alpha = r.r / p.Ap"}
        except Exception as exc:  # the extra legs must never cost the headline line
            line["two_sweep_extra"] = {"error": repr(exc)}
        try:
            plan.close()
            line["csr_extra"] = csr_leg(capi, local_rank, 8192, args.iters, max(2, min(args.steps, 3)), peak, peak_src)
        except Exception as exc:
            line["csr_extra"] = {"error": repr(exc)}
        finished.set()
    print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
