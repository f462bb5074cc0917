"""CPU-side checks of the drop-in boundary: libgsb200.so loads, exports every symbol include/gsb200.h
declares (and nothing the header does not), the ctypes table matches the header, and compute entry
points fail loudly without a GPU instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsb200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(gsb):
    names = _declared()
    assert len(names) >= 40
    lib = gsb.load()
    out = subprocess.run(["nm", "-D", "--defined-only", gsb._lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = {ln.split()[-1] for ln in out.stdout.splitlines() if " T " in ln and ln.split()[-1].startswith("gsb_")}
    missing = [n for n in names if n not in exported]
    assert not missing, "declared in gsb200.h but not exported: %s" % missing
    # the ctypes table covers exactly the header
    assert sorted(gsb._lib.SIGNATURES) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_no_torch_types_in_the_abi():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # declarations only, comments stripped
    assert "torch" not in code.lower() and "tensor" not in code.lower()
    # plain pointers and sizes only
    assert not re.search(r"std::|at::|c10::", src)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the package (or the C sources) may reference it."""
    pkg = os.path.join(ROOT, "coursecomputationalphotography_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"\boracle\b|gs_oracle|liboracle|libgsref", txt):
                    bad.append(os.path.join(dp, f))
    for f in ("gsb200.h", "gsb_sparse_matrix.hpp", "sparse-matrix.h"):
        txt = open(os.path.join(ROOT, "include", f)).read()
        if re.search(r"gs_oracle|liboracle|libgsref", txt):
            bad.append(f)
    assert not bad, bad
    out = subprocess.run(["ldd", os.path.join(pkg, "libgsb200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "gsref" not in out


def test_status_text_and_version(gsb):
    lib = gsb.load()
    assert lib.gsb_version() >= 100
    assert isinstance(lib.gsb_last_error(), bytes)
    o = gsb.GsOptions()
    lib.gsb_gs_default_options(C.byref(o))
    assert (o.ordering, o.check_every, o.use_graph, o.kernel) == (0, 1, -1, 0)


def test_fails_loudly_without_a_gpu(gsb):
    if gsb._lib.device_count() > 0:
        pytest.skip("a GPU is visible: the no-device path cannot be exercised here")
    with pytest.raises(gsb.GsbError) as e:
        gsb.SparseMatrix(np.float64)
    assert e.value.status == 6 and "no CPU fallback" in str(e.value)
    with pytest.raises(gsb.GsbError) as e:
        gsb.manhattonDist([1.0, 2.0], [2.0, 1.0])
    assert e.value.status == 6
    with pytest.raises(gsb.GsbError):
        gsb.writeback_u8(np.zeros(4))
    # the multi-device entry points likewise: no device, no group, no device list
    from coursecomputationalphotography_b200 import strips
    with pytest.raises(gsb.GsbError) as e:
        strips.LocalGroup([0, 1])
    assert e.value.status == 6
    with pytest.raises(gsb.GsbError) as e:
        strips.set_devices([0, 1])
    assert e.value.status == 6
    strips.set_devices([])  # clearing the list needs no device
    assert strips.get_devices() == []


def test_cpp_dropin_header_compiles():
    """include/sparse-matrix.h must compile as the reference's translation units use it (C++17,
    optional USE_NAME_SPACE wrapper, both element types)."""
    src = r'''
    #define USE_NAME_SPACE refns
    #include "sparse-matrix.h"
    int main() {
        refns::SparseMatrix<double> a; refns::SparseMatrix<int> b;
        refns::SparseMatrix<double>::Triplet t{0, 0, 1.0}; (void)t;
        return (int)(a.rows() + b.cols());
    }'''
    p = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), "-x", "c++",
                        "-"], input=src, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
