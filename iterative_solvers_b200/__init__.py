"""iterative_solvers_b200: B200-native (sm_100a, fp64) conjugate-gradient Dirichlet-Poisson hot path of
Ruslan361/iterative_solvers behind a C ABI (include/b200cg.h).

The product is libb200cg.so (CUDA/C++) plus the C++ drop-in classes in dropin/. This Python package only
holds a ctypes mirror of the C ABI (`capi`) used by the tests and bench.py, and the in-tree build recipe
(`build`). There is no CPU fallback: `capi.lib()` raises if the library is missing, and every compute call
raises without a CUDA device.
"""
from . import build, capi  # noqa: F401
from .capi import B200CGError, Plan, SolveInfo  # noqa: F401

__all__ = ["build", "capi", "Plan", "SolveInfo", "B200CGError"]
