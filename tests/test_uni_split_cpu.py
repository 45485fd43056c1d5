"""CPU unit test of the composition-range split planner of the univariate-skip fast path
(binius_b200/csrc/uni_split.hpp; C++, compiled with g++ -- no CUDA needed)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_planner(tmp_path):
    exe = str(tmp_path / "uni_split_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "uni_split_test.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and "uni split ok" in out.stdout
