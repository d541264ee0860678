"""INTEGRATION.md shows a maintainer how to bind the reference's classes to the C ABI. tests/snippets/binding_example.cpp
is that code; it must keep compiling against include/b200cg.h and linking against libb200cg.so (CPU: no device needed to
compile, link and ask the library for its version)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_binding_example_compiles_links_and_loads(tmp_path):
    from iterative_solvers_b200 import build

    lib = build.build_library()
    exe = os.path.join(tmp_path, "binding_example")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "snippets", "binding_example.cpp"), "-o", exe,
                           "-L", os.path.dirname(lib), f"-Wl,-rpath,{os.path.dirname(lib)}", "-lb200cg"])
    assert subprocess.run([exe], timeout=120).returncode == 0
