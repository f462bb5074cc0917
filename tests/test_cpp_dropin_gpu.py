"""Builds tests/cpp/dropin_main.cc against include/sparse-matrix.h (the C++ drop-in for the reference's
header) and runs it on the GPU: the reference's own lab3 self-test flow (main6.cc:192-253)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_dropin_replays_reference_selftest(gsb):
    lib_dir = os.path.join(ROOT, "coursecomputationalphotography_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "dropin_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "dropin_main.cc"), "-o", exe, "-L", lib_dir, "-lgsb200",
                    "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dropin ok" in r.stdout
