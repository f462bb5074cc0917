"""Builds tests/cpp/dropin_main.cc against include/sparse-matrix.h (the C++ drop-in for the reference's
header) and runs it on the GPU: the reference's own lab3 self-test flow (main6.cc:192-253)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    lib_dir = os.path.join(ROOT, "coursecomputationalphotography_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "dropin_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "dropin_main.cc"), "-o", exe, "-L", lib_dir, "-lgsb200",
                    "-Wl,-rpath," + lib_dir], check=True)
    return exe


def test_cpp_dropin_compiles_in_both_error_modes():
    """CPU: the header compiles as the reference's callers include it, non-throwing (default) and throwing."""
    for extra in ([], ["-DGSB_THROW_ON_ERROR"]):
        subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include")] + extra +
                       [os.path.join(ROOT, "tests", "cpp", "dropin_main.cc")], check=True)


def test_cpp_host_accessors_against_dense_mirror(gsb):
    """CPU: at / coeff / insert* of the C++ drop-in against a dense mirror (the reference's CheckEqual idea,
    main6.cc:19-33): 30 000 random edits over zero -> slack -> refill -> grow, int and double."""
    lib_dir = os.path.join(ROOT, "coursecomputationalphotography_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "host_accessors")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "host_accessors.cc"), "-o", exe, "-L", lib_dir, "-lgsb200",
                    "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host accessors ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_dropin_replays_reference_selftest(gsb):
    r = subprocess.run([_build()], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dropin ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_cpp_dropin_multi_device(gsb, ndev):
    """SparseMatrix<double>::gaussSeidel[Multi] on an imported masked system, N devices of one process (setDevices)
    == one device, bit for bit, same stop sweep."""
    if gsb._lib.device_count() < ndev:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (ndev, ndev))
    r = subprocess.run([_build(), "mgpu", str(ndev)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dropin multi-device ok" in r.stdout
