"""Builds libb200cg.so (nvcc, sm_100a) and the C++ drop-in test driver in-tree. No JIT cache: the built
files sit next to the sources so they travel to the GPU box with the repository snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libb200cg.so")
DROPIN_DIR = os.path.join(PKG, "dropin")
DROPIN_LIB = os.path.join(PKG, "libb200_dropin.so")
DROPIN_TEST = os.path.join(PKG, "dropin_test")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(directory: str, exts: tuple[str, ...]) -> list[str]:
    return sorted(os.path.join(directory, f) for f in os.listdir(directory) if f.endswith(exts))


def build_library(force: bool = False, verbose: bool = False) -> str:
    deps = _sources(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "b200cg.h")]
    if force or _newer(LIB, deps):
        # one nvcc per translation unit, in parallel (the sweep-kernel instantiations dominate the build time)
        units = ["plan.cu", "solve.cu", "comm.cu", "mg.cu"]
        objs = [os.path.join(CSRC, u[:-3] + ".o") for u in units]
        flags = [*ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"] + (["-Xptxas=-v"] if verbose else [])
        procs = [subprocess.Popen([NVCC, *flags, "-c", os.path.join(CSRC, u), "-o", o]) for u, o in zip(units, objs)]
        if any(p.wait() != 0 for p in procs):
            raise RuntimeError("nvcc failed")
        subprocess.check_call([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl"])
    return LIB


def build_dropin_test(force: bool = False) -> str | None:
    """C++ host classes mirroring the reference's public surface (libb200_dropin.so) + their test driver."""
    if not os.path.isdir(DROPIN_DIR):
        return None
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    hdrs = _sources(DROPIN_DIR, (".hpp", ".h")) + [os.path.join(ROOT, "include", "b200cg.h")]
    common = ["-std=c++17", "-O2", "-Wall", "-Wextra", "-I", DROPIN_DIR, "-I", os.path.join(ROOT, "include"),
              "-L", PKG, f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN"]
    lib_src = os.path.join(DROPIN_DIR, "b200_dropin.cpp")
    if force or _newer(DROPIN_LIB, [lib_src, LIB] + hdrs):
        subprocess.check_call([cxx, "-fPIC", "-shared", "-o", DROPIN_LIB, lib_src, *common, "-lb200cg"])
    test_src = os.path.join(DROPIN_DIR, "dropin_test.cpp")
    if force or _newer(DROPIN_TEST, [test_src, DROPIN_LIB] + hdrs):
        subprocess.check_call([cxx, "-o", DROPIN_TEST, test_src, *common, "-lb200_dropin", "-lb200cg", "-lpthread"])
    return DROPIN_TEST


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_library(force, verbose)
    build_dropin_test(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
