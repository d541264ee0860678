"""GPU parity of the single-sweep iteration (the default of the REL_L2 rule; csrc/fused_kernel.cuh).

One kernel per iteration; alpha comes from the single-reduction CG recurrence instead of p.Ap. The bar is the same as
for the default path (BASELINE.json north_star): iteration count within +-1 of the reference, solution within 1e-10
relative. tests/studies/single_reduction_cg.py (numpy) and scripts/model_single_sweep.py (lane-level model of the
kernel's data flow) are the CPU-side evidence; these tests are the product check through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-10
DOMAINS = {0: (0.0, 1.0), 1: (1.0, 2.0)}


@pytest.fixture(scope="module")
def capi():
    from iterative_solvers_b200 import capi as c

    c.lib()
    assert c.device_count() >= 1, "these tests need a CUDA device"
    return c


def relmax(x, ref):
    return np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300)


def fused_solve(p, **kw):
    # small_grid_path=1: small grids would otherwise take the cluster-resident kernel
    x, info = p.solve(single_sweep=1, small_grid_path=1, **kw)
    assert info["single_sweep"] == 1 and info["cluster_path"] == 0
    return x, info


@pytest.mark.parametrize("n,a_tag,iters", [(6, 1, 13), (30, 1, 88), (64, 0, 178), (128, 0, 352), (128, 1, 362)])
def test_reference_fixtures(capi, golden_ref, n, a_tag, iters):
    """The grids the unmodified reference was run on (tests/golden): same iteration counts, same solution."""
    tag = f"mf_n{n}_a{a_tag}"
    a, b = DOMAINS[a_tag]
    with capi.Plan(n, n, a, b, a, b) as p:
        x, info = fused_solve(p, b=golden_ref[tag + "_rhs"], eps_rel=1e-8, max_it=10000)
        assert abs(info["iterations"] - iters) <= 1
        assert info["converged"]
        assert relmax(x, golden_ref[tag + "_x"]) < REL
        assert info["r_l2"] <= 1e-8 * info["r0_l2"]


def check_against_oracle(o, p, eps):
    """Converged solve (eps > 0): the reference's iteration count and solution. eps = None (grids on which a converged
    oracle solve would take minutes of host time): 80 iterations - every one runs the same kernel over the same strips and
    tiles - against the oracle with long-double sums (the reference's sequential fp64 sums are the inaccurate side at these
    sizes, test_gpu_parity.py::test_config2_fixed_iterations_vs_oracle)."""
    if eps is None:
        ref = o.mf_solve(eps=0.0, max_it=80, accurate_dots=True)
        x, info = fused_solve(p, b=o.rhs(), eps_rel=0.0, max_it=80)
        assert info["iterations"] == ref["iterations"] == 80
        assert abs(info["r_l2"] - ref["r_norm"]) <= 1e-9 * ref["r_norm"]
    else:
        ref = o.mf_solve(eps=eps, max_it=20000)
        x, info = fused_solve(p, b=o.rhs(), eps_rel=eps, max_it=20000)
        assert abs(info["iterations"] - ref["iterations"]) <= 1
    assert relmax(x, ref["x"]) < REL


@pytest.mark.parametrize("n,domain,tile_rows,eps", [(256, 0, 0, 1e-8), (600, 0, 0, None), (1030, 0, 7, None),
                                                     (333, 1, 0, 1e-8), (1009, 1, 5, None), (64, 0, 1, 1e-9),
                                                     (64, 0, 3, 1e-9)])
def test_strips_tiles_and_domains_vs_oracle(capi, oracle_mod, n, domain, tile_rows, eps):
    """Several strips (n > 420), ragged and one-row tiles, the full rectangle."""
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, domain)
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, domain=domain, tile_rows=tile_rows) as p:
        check_against_oracle(o, p, eps)


@pytest.mark.parametrize("n,m", [(7, 7), (9, 12), (33, 20), (481, 333)])
def test_general_lshape(capi, oracle_mod, n, m):
    o = oracle_mod.Oracle(m, n, 0.0, 1.0, 0.0, 1.0, oracle_mod.LSHAPE_ANY)
    ref = o.mf_solve(eps=1e-7, max_it=20000)
    with capi.Plan(m, n, 0.0, 1.0, 0.0, 1.0, domain=capi.DOMAIN_LSHAPE_ANY) as p:
        x, info = fused_solve(p, b=o.rhs(), eps_rel=1e-7, max_it=20000)
        assert abs(info["iterations"] - ref["iterations"]) <= 1
        assert relmax(x, ref["x"]) < REL


def test_fixed_iteration_counts_match_the_default_path(capi, oracle_mod):
    """Stopped early at odd and even counts (the pending x update of an even last sweep is flushed): iterate by iterate
    the single-sweep path stays at rounding distance from the two-sweep path and from the oracle."""
    n = 512
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, 0)
    b = o.rhs()
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        for iters in (1, 2, 7, 40):
            ref = o.mf_solve(b=b, eps=0.0, max_it=iters)
            xd, idf = p.solve(b=b, eps_rel=0.0, max_it=iters, single_sweep=2)
            xs, isf = fused_solve(p, b=b, eps_rel=0.0, max_it=iters, iters_per_graph=6)
            assert idf["single_sweep"] == 0
            assert isf["iterations"] == idf["iterations"] == iters
            # the oracle sums its dot products sequentially like the reference (~1e-12 accurate on this rhs); the two GPU
            # paths share the tree sums and sit closer to each other
            assert relmax(xs, ref["x"]) < REL and relmax(xd, ref["x"]) < REL
            assert relmax(xs, xd) < 1e-11
            assert abs(isf["r_l2"] - idf["r_l2"]) <= 1e-10 * idf["r_l2"]


def test_edge_cases_and_fallbacks(capi, oracle_mod):
    o = oracle_mod.Oracle(64, 64, 0.0, 1.0, 0.0, 1.0, 0)
    b, u = o.rhs(), o.true_solution()
    with capi.Plan(64, 64, 0.0, 1.0, 0.0, 1.0) as p:
        x, info = fused_solve(p, b=np.zeros_like(b), eps_rel=1e-8, max_it=100)  # zero rhs: nothing to do
        assert info["iterations"] == 0 and not np.any(x)
        x, info = fused_solve(p, b=b, eps_rel=1e-8, max_it=0)
        assert info["iterations"] == 0 and not info["converged"]
        # the max-norm rules have their own single-sweep flavour (tests below) ...
        ref = o.msg_solve(b=b, u=u, eps_p=1e-8, eps_r=1e-8, max_it=20000)
        x, info = p.solve(b=b, u=u, rule=capi.RULE_MAXNORM, eps_p=1e-8, eps_r=1e-8, max_it=20000, single_sweep=1,
                          small_grid_path=1)
        assert info["single_sweep"] == 1 and info["x_deferral"] == 0 and info["iterations"] == ref["iterations"]
        # ... where the single sweep does not apply (per-iteration report, assembled operator) the request is ignored
        p.assemble_csr()
        x, info = p.solve(b=b, u=u, op=capi.OP_CSR, rule=capi.RULE_MAXNORM, eps_p=1e-8, eps_r=1e-8, max_it=20000,
                          single_sweep=1)
        assert info["single_sweep"] == 0 and info["iterations"] == ref["iterations"]
        got = []
        x, info = p.solve(b=b, u=u, eps_rel=1e-8, max_it=20000, single_sweep=1, small_grid_path=1,
                          callback=lambda it, pr, r, e: got.append(it))
        assert info["single_sweep"] == 0 and len(got) > 0


def test_large_grid_property(capi):
    """4096^2 (config 2 size): 60 single-sweep iterations against 60 two-sweep iterations on the device-built rhs."""
    n = 4096
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        p.build_rhs()
        xd, idf = p.solve(rhs_on_device=True, eps_rel=0.0, max_it=60, single_sweep=2)
        xs, isf = p.solve(rhs_on_device=True, eps_rel=0.0, max_it=60, single_sweep=1)
        assert isf["single_sweep"] == 1 and idf["single_sweep"] == 0
        assert isf["iterations"] == idf["iterations"] == 60
        assert relmax(xs, xd) < 1e-11
        assert abs(isf["r_l2"] - idf["r_l2"]) <= 1e-10 * idf["r_l2"]


@pytest.mark.parametrize("n,domain,tile_rows,eps", [(64, 0, 0, 1e-8), (64, 0, 3, 1e-8), (900, 0, 0, None), (1030, 0, 7, None),
                                                     (333, 1, 0, 1e-8), (1709, 1, 5, None)])
def test_wide_geometry_forced_on_small_grids(capi, oracle_mod, n, domain, tile_rows, eps):
    """Slabs of >= 4 M unknowns run the single sweep as one 15-warp CTA per SM on 840-column strips (test_large_grid_property,
    the 4096^2 and 16384^2 tests); B200CG_FUSED_CW=14 forces that geometry onto grids the oracle handles in seconds."""
    import os

    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, domain)
    os.environ["B200CG_FUSED_CW"] = "14"
    try:
        with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, domain=domain, tile_rows=tile_rows) as p:  # the knob is read here
            check_against_oracle(o, p, eps)
    finally:
        os.environ.pop("B200CG_FUSED_CW", None)


# ---------------------------------------------------------------- MSGSolver's max-norm rules in one sweep (F_MAXN)
def maxnorm_solve(p, capi, ss=1, **kw):
    x, info = p.solve(rule=capi.RULE_MAXNORM, single_sweep=ss, small_grid_path=1, **kw)
    assert info["single_sweep"] == (1 if ss != 2 else 0) and info["cluster_path"] == 0 and info["x_deferral"] == 0
    return x, info


@pytest.mark.parametrize("n,a_tag", [(6, 1), (30, 1), (128, 0)])
def test_maxnorm_reference_fixtures(capi, golden_ref, n, a_tag):
    """MSGSolver::solve (msg_solver.cpp:10-212) as ONE sweep per iteration against the fixtures of the unmodified
    reference: 14 / 79 / 102 / 355 / 482 iterations, stop reasons, solution, final norms, callback cadence and values."""
    tag = f"grid_n{n}_a{a_tag}"
    eps = 1e-6 if n <= 30 else 1e-8
    a, b = DOMAINS[a_tag]
    with capi.Plan(n, n, a, b, a, b) as p:
        for cname, kw in {"pr": dict(eps_p=eps, eps_r=eps), "r": dict(eps_p=-1.0, eps_r=eps)}.items():
            info_ref = golden_ref[f"{tag}_msg_{cname}_info"]
            cb_ref = golden_ref[f"{tag}_msg_{cname}_cb"]
            got = []
            x, info = maxnorm_solve(p, capi, b=golden_ref[tag + "_rhs"], u=golden_ref[tag + "_true"], max_it=10000,
                                    callback=lambda it, pr, rs, er: got.append((it, pr, rs, er)), **kw)
            assert info["iterations"] == int(info_ref[0]) and info["converged"] == bool(info_ref[1])
            assert info["stop_reason"] == capi.STOP_NAMES[int(info_ref[2])]
            assert relmax(x, golden_ref[f"{tag}_msg_{cname}_x"]) < REL
            bscale = np.max(np.abs(golden_ref[tag + "_rhs"]))
            assert abs(info["r_max"] - info_ref[3]) <= REL * bscale and abs(info["r_max"] - info_ref[3]) <= 1e-3 * abs(info_ref[3])
            assert abs(info["dx_max"] - info_ref[4]) <= 1e-6 * abs(info_ref[4])
            assert abs(info["err_max"] - info_ref[5]) <= 1e-9 * abs(info_ref[5])
            got = np.array(got)
            assert np.array_equal(got[:, 0], cb_ref[:, 0])  # it 0, 1, every 100, final (msg_solver.cpp:75,172,193)
            assert got[0, 1] == cb_ref[0, 1] == np.finfo(np.float64).max
            assert np.allclose(got[1:, 1:], cb_ref[1:, 1:], rtol=1e-3, atol=0)


@pytest.mark.parametrize("n,domain,tile_rows,with_u,wide,converged",
                         [(256, 0, 0, True, False, True), (700, 0, 0, True, False, False), (1030, 0, 7, False, False, False),
                          (333, 1, 0, True, False, True), (900, 0, 0, True, True, False), (1009, 1, 5, False, True, False),
                          (64, 0, 3, True, True, True)])
def test_maxnorm_strips_tiles_and_domains_vs_oracle(capi, oracle_mod, n, domain, tile_rows, with_u, wide, converged):
    """Several strips, ragged tiles, the full rectangle, with and without the true solution (four / three streams), both
    strip geometries (B200CG_FUSED_CW=14 forces the wide one of large slabs onto these grids). converged: the oracle's
    MSGSolver (reference arithmetic on the assembled matrix) stops at the same iteration for the same reason with the same
    solution and norms. Otherwise (grids on which that would take minutes of host time): 80 iterations against the oracle
    with long-double sums. Either way the two-sweep iteration agrees."""
    import os

    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, domain)
    b = o.rhs()
    u = o.true_solution() if with_u else None
    rules = [dict(eps_p=1e-7, eps_r=1e-7, max_it=20000), dict(eps_p=-1.0, eps_r=1e-5, max_it=20000)] if converged else \
            [dict(eps_p=-1.0, eps_r=1e-300, max_it=80)]
    if wide:
        os.environ["B200CG_FUSED_CW"] = "14"
    try:
        with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0, domain=domain, tile_rows=tile_rows) as p:  # the knob is read here
            for kw in rules:
                ref = o.msg_solve(b=b, u=u, accurate_dots=not converged, **kw)
                x, info = maxnorm_solve(p, capi, b=b, u=u, **kw)
                x2, info2 = maxnorm_solve(p, capi, ss=2, b=b, u=u, **kw)
                assert info["iterations"] == info2["iterations"] and abs(info["iterations"] - ref["iterations"]) <= 1
                assert info["stop_reason"] == ref["stop_reason"] == info2["stop_reason"]
                assert relmax(x, x2) < 1e-11
                if info["iterations"] == ref["iterations"]:
                    assert relmax(x, ref["x"]) < REL
                    assert abs(info["dx_max"] - ref["dx_max"]) <= 1e-6 * ref["dx_max"]
                    assert abs(info["r_max"] - ref["r_max"]) <= 1e-3 * ref["r_max"] + REL * np.max(np.abs(b))
                    if with_u:  # (a converged |x - u|_inf is the discretisation error: x agrees to REL of |u|, not of that)
                        assert abs(info["err_max"] - ref["err_max"]) <= 1e-9 * ref["err_max"] + REL * np.max(np.abs(u))
                if not with_u:
                    assert info["err_max"] == np.finfo(np.float64).max
    finally:
        os.environ.pop("B200CG_FUSED_CW", None)


def test_maxnorm_iterate_by_iterate(capi, oracle_mod):
    """Stopped after 1, 2, 7, 40 iterations at 512^2 (FULL and generic stages, two blocks of the L): x and the three
    max-norms of every stop agree with the oracle's MSGSolver (long-double sums) and with the two-sweep iteration to rounding - the maxima are
    order-independent, so they pin the masks of the branch-free path (no stale column, no lane counted that does not write).
    Then the exact-error rule and an interrupt."""
    import ctypes

    n = 512
    o = oracle_mod.Oracle(n, n, 0.0, 1.0, 0.0, 1.0, 0)
    b, u = o.rhs(), o.true_solution()
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        for iters in (1, 2, 7, 40):
            ref = o.msg_solve(b=b, u=u, eps_p=-1.0, eps_r=1e-300, max_it=iters, accurate_dots=True)
            x, info = maxnorm_solve(p, capi, b=b, u=u, eps_p=-1.0, eps_r=1e-300, max_it=iters, iters_per_graph=6)
            res, err = p.postprocess()  # x is current after every iteration: no pending update to settle
            assert abs(np.max(np.abs(err)) - info["err_max"]) <= 1e-15 * np.max(np.abs(u))
            x2, info2 = maxnorm_solve(p, capi, ss=2, b=b, u=u, eps_p=-1.0, eps_r=1e-300, max_it=iters)
            assert info["iterations"] == info2["iterations"] == ref["iterations"] == iters
            assert info["stop_reason"] == "ITERATIONS" and not info["converged"]
            assert relmax(x, ref["x"]) < REL and relmax(x, x2) < 1e-11
            for key, tol in (("r_max", 1e-9), ("dx_max", 1e-9), ("err_max", 1e-10)):
                assert abs(info[key] - ref[key]) <= tol * abs(ref[key]), (iters, key, info[key], ref[key])
                assert abs(info[key] - info2[key]) <= tol * abs(info2[key]), (iters, key)
        ref = o.msg_solve(b=b, u=u, eps_p=-1.0, eps_r=-1.0, eps_e=1e-3, max_it=20000)
        x, info = maxnorm_solve(p, capi, b=b, u=u, eps_p=-1.0, eps_r=-1.0, eps_e=1e-3, max_it=20000)
        assert ref["stop_reason"] == info["stop_reason"] == "EXACT_ERROR" and info["iterations"] == ref["iterations"]
        flag = ctypes.c_int(1)
        x, info = maxnorm_solve(p, capi, b=b, u=u, eps_p=-1.0, eps_r=1e-300, max_it=100000, iters_per_graph=100,
                                stop_flag=flag)
        assert info["stop_reason"] == "INTERRUPTED" and info["iterations"] == 16


def test_maxnorm_large_grid_property(capi):
    """4096^2 (the wide geometry by slab size), 60 iterations with the true solution: single sweep against two sweeps."""
    n = 4096
    with capi.Plan(n, n, 0.0, 1.0, 0.0, 1.0) as p:
        p.build_rhs()
        u = p.true_solution()
        kw = dict(rhs_on_device=True, u=u, rule=capi.RULE_MAXNORM, eps_p=-1.0, eps_r=1e-300, max_it=60)
        xd, idf = p.solve(single_sweep=2, **kw)
        xs, isf = p.solve(single_sweep=0, **kw)  # the default
        assert isf["single_sweep"] == 1 and idf["single_sweep"] == 0 and isf["iterations"] == idf["iterations"] == 60
        assert relmax(xs, xd) < 1e-11
        for key in ("r_max", "dx_max", "err_max"):
            assert abs(isf[key] - idf[key]) <= 1e-9 * abs(idf[key]), key
