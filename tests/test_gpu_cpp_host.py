"""Runs the C++ conformance program of the compiled host mirror (binius_b200/host/compute_layer.hpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_conformance():
    exe = os.path.join(ROOT, "tests", "cpp", "conformance")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", ROOT, "tests/cpp/conformance"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp conformance ok" in out.stdout


def test_cpp_host_header_compiles():
    # CPU-only: the header is self-contained C++17 over the C ABI
    src = '#include "binius_b200/host/compute_layer.hpp"\n#include "binius_b200/host/computation_backend.hpp"\nint main() { return 0; }\n'
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", ROOT, "-"], input=src, text=True, capture_output=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
