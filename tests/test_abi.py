"""CPU tests of the drop-in boundary: the C-ABI library builds, loads without a GPU, exports exactly the
symbols include/b200cg.h declares, validates arguments, and fails loudly (no CPU fallback) on compute calls."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from iterative_solvers_b200 import build, capi as c

    build.build_library()
    c.lib()
    return c


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200cg.h"), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200cg_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_what_the_mirror_binds(capi):
    assert header_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(capi):
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, f"declared in include/b200cg.h but not exported: {missing}"
    for s in header_symbols():
        assert hasattr(capi.lib(), s)


def test_library_is_sm100a_native(capi):
    """The fat binary holds sm_100a code for the hot kernels (cuobjdump works without a GPU)."""
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_version_and_struct_sizes(capi):
    import ctypes as C

    assert capi.lib().b200cg_version() == 100
    # layouts the header promises (ints before doubles are padded as a C compiler would)
    assert C.sizeof(capi.PlanDesc) == 96
    assert C.sizeof(capi.Params) == 88
    assert C.sizeof(capi.SolveInfo) % 8 == 0


def test_partition_is_pure_geometry(capi):
    for n, domain, world in [(128, 0, 1), (128, 0, 2), (128, 0, 8), (4096, 0, 8), (46341, 1, 8), (1000, 1, 3)]:
        parts = [capi.partition(n, n, domain, r, world) for r in range(world)]
        N = parts[0][4]
        expect = (n // 2 - 1) * (n // 2) + (n - 1) * (n // 2 - 1) if domain == 0 else (n - 1) * (n - 1)
        assert N == expect
        assert parts[0][0] == 1 and parts[0][2] == 0
        assert parts[-1][1] == n and parts[-1][3] == N
        for a, b in zip(parts, parts[1:]):
            assert a[1] == b[0] and a[3] == b[2]  # contiguous rows and contiguous compact ranges
        sizes = np.array([p[3] - p[2] for p in parts])
        assert sizes.min() > 0
        assert sizes.max() - sizes.min() <= 2 * (n - 1)  # balanced to within two full rows
    assert capi.partition(46342, 46342, 1, 7, 8)[3] == 46341**2 > 2**31 - 1  # 64-bit global indices


def test_invalid_geometry_is_rejected(capi):
    for n, m, domain in [(8, 6, 0), (7, 7, 0), (2, 2, 0), (1, 5, 1)]:
        with pytest.raises(capi.B200CGError) as e:
            capi.partition(m, n, domain, 0, 1)
        assert e.value.status == 1
    with pytest.raises(capi.B200CGError):
        capi.partition(128, 128, 0, 3, 2)  # rank outside world


def test_no_cpu_fallback(capi):
    """Without a CUDA device every plan creation fails with ERR_NO_DEVICE; with one this test is skipped."""
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.B200CGError) as e:
        capi.Plan(6, 6, 1.0, 2.0, 1.0, 2.0)
    assert e.value.status == capi.ERR_NO_DEVICE
    assert "no" in str(e.value).lower()


def test_solve_batch_validates_its_arguments_without_a_device(capi):
    """b200cg_solve_batch rejects a NULL plan / negative count before it touches CUDA."""
    import ctypes as C

    L = capi.lib()
    prm = capi.Params(op=capi.OP_MATRIX_FREE, rule=capi.RULE_REL_L2, eps_rel=1e-8, max_it=10)
    info = (capi.SolveInfo * 1)()
    ptrs = (C.c_void_p * 1)()
    assert L.b200cg_solve_batch(None, C.byref(prm), 1, ptrs, ptrs, info, None, None, None) == 1
    assert b"NULL" in L.b200cg_last_error()
    assert L.b200cg_solve_batch(None, C.byref(prm), -1, ptrs, ptrs, info, None, None, None) == 1


def test_product_does_not_reach_into_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "iterative_solvers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "cg_oracle" not in text and "oracle." not in text and "oracle/" not in text.replace(
                    "oracle/shim/KokkosSparse_spmv.hpp", ""), f"{f} references the oracle"
