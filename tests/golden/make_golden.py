"""Generates tests/golden/ref_golden.json by running the UNMODIFIED reference headers
(oracle/_ref/libgsref.so, compiled in place from /root/reference by oracle/Makefile).
Run where /root/reference exists:  python tests/golden/make_golden.py
Floats are stored as C99 hex strings so the fixtures are bit-exact."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from coursecomputationalphotography_b200 import workloads as wl  # noqa: E402

pyoracle.build()
hx = lambda a: [float(v).hex() for v in np.asarray(a, np.float64).ravel()]
li = lambda a: [int(v) for v in np.asarray(a).ravel()]


def lay(ref):
    L = ref.layout()
    return {"vals": [float(v) if ref.sfx == "f64" else int(v) for v in L.values], "cols": li(L.cols),
            "row_begin": li(L.row_begin), "row_nnz": li(L.row_nnz), "row_left": li(L.row_left),
            "n_rows": L.n_rows, "n_cols": L.n_cols, "dense": ref.dense().tolist()}


out = {"source": "oracle/_ref/libgsref.so = /root/reference labs/lab3 (v1) and labs/lab8 (v2) sparse-matrix.h, "
                 "g++ -std=c++17 -O2 -include cstring -include cmath -fno-access-control"}

# --- labs/lab3/src/OpenCVHW1/main6.cc:192-231 -------------------------------------------------
for ver in (1, 2):
    ref = pyoracle.Ref(ver, "i32")
    ref.init_from_vector([0, 0, 0, 2, 2], [0, 3, 4, 0, 2], [1, 1, 0, 8, 1])
    steps = {"init": lay(ref)}
    for name, (x, r, c) in (("T1", (0, 1, 0)), ("T2", (0, 0, 0)), ("T3", (1, 2, 2)), ("T4", (8, 0, 0)),
                            ("T5", (9, 1, 1))):
        ref.insert(x, r, c)
        steps[name] = lay(ref)
    out["lab3_fixture" if ver == 1 else "lab3_fixture_v2"] = steps

# --- main6.cc:233-253 ----------------------------------------------------------------------------
out["manhatton"] = pyoracle.ref_manhatton([1.0, 2.0, 3.0, 10.0], [2.0, 1.0, 3.0, 8.0], 1)
A = [10, -1, 2, 0, -1, 11, -1, 3, 2, -1, 10, -1, 0, 3, -1, 8]
b = np.array([6, 25, -11, 15], np.float64)
for ver in (1, 2):
    ref = pyoracle.Ref(ver, "i32").init_dense(4, 4, A)
    out["gs_4x4_v%d" % ver] = hx(ref.gauss_seidel(b))
    out["cg_4x4_v%d" % ver] = hx(ref.cg(b))
    out["layout_4x4_v%d" % ver] = lay(ref)

# --- Poisson import (initializeFromEigenRowMajor, compressed) ------------------------------------
ro, ci, va = pyoracle.poisson_csr(4, 3)
ref = pyoracle.Ref(2, "f64").import_csr(va, ro[:-1], ci, 12)
out["poisson_4x3_import"] = lay(ref)

W, H = 8, 6
ro, ci, va = pyoracle.poisson_csr(W, H)
img = wl.synth_image(W, H, 1, seed=7)
gx, gy = wl.forward_gradients(img)
bp = pyoracle.poisson_rhs(W, H, gx[0], gy[0], float(img[0, 0, 0]))
ref = pyoracle.Ref(2, "f64").import_csr(va, ro[:-1], ci, W * H)
sgx, sgy = wl.seamless_gradients(img)
bs = pyoracle.poisson_rhs(W, H, sgx[0], sgy[0], float(img[0, 0, 0]))
out["poisson_8x6"] = {"b": hx(bp), "gs": {str(k): hx(ref.gauss_seidel(bp, 0.0, k)) for k in (1, 10, 100)},
                      "spmv_of_b": hx(ref.spmv(bp)), "b_seamless": hx(bs),
                      "cg_12": hx(ref.cg(bp, 1e-10, 12)),
                      "cg_8_init": hx(ref.cg(bs, 1e-10, 8, img[0].ravel().astype(float))),
                      "pcg_10": hx(ref.pcg(bs, 1e-10, 10))}

# --- uncompressed import (per-row counts) and trailing empty rows ---------------------------------
rng = np.random.default_rng(4)
cap = rng.integers(2, 7, 12)
used = np.minimum(cap, rng.integers(0, 7, 12))
off = np.r_[0, np.cumsum(cap)].astype(np.int32)
vals = np.round(rng.uniform(1, 2, off[-1]), 3)
cols = np.concatenate([np.sort(rng.choice(20, k, replace=False)) for k in cap]).astype(np.int32)
ref = pyoracle.Ref(2, "f64").import_csr(vals, off[:-1], cols, 20, used)
out["import_counts"] = {"values": vals.tolist(), "row_off": li(off[:-1]), "cols": li(cols), "used": li(used),
                        "n_cols": 20, "layout": {k: v for k, v in lay(ref).items() if k != "dense"}}
lens = np.r_[rng.integers(1, 5, 9), 0, 0, 0]
off = np.r_[0, np.cumsum(lens)].astype(np.int32)
vals = np.round(rng.uniform(1, 2, off[-1]), 3)
cols = np.concatenate([np.sort(rng.choice(12, k, replace=False)) for k in lens if k]).astype(np.int32)
ref = pyoracle.Ref(2, "f64").import_csr(vals, off[:-1], cols, 12)
out["import_trailing_empty"] = {"values": vals.tolist(), "row_off": li(off[:-1]), "cols": li(cols), "n_cols": 12,
                                "layout": {k: v for k, v in lay(ref).items() if k != "dense"}}

# --- config-1 analogue, small: reference defaults ---------------------------------------------------
r, c, v, bb, xs = wl.diag_dominant_system(300, 4, seed=42)
ref = pyoracle.Ref(2, "f64").init_from_vector(r, c, v)
out["diag_dominant_300"] = {"gs_default": hx(ref.gauss_seidel(bb)), "gs_3": hx(ref.gauss_seidel(bb, 0.0, 3))}

# --- initializeFromVector with explicit zeros, both element types -----------------------------------
rows = [0, 0, 0, 0, 2, 2, 5, 5, 5]
colz = [1, 2, 5, 7, 0, 3, 2, 4, 7]
valz = [0, 3, 0, 4, 0, 0, 6, 0, 0]
for sfx in ("i32", "f64"):
    ref = pyoracle.Ref(2, sfx).init_from_vector(rows, colz, valz)
    out["zeros_" + sfx] = {"rows": rows, "cols": colz, "vals": valz, "layout": lay(ref)}

# --- Dirichlet-masked blend (small): the reference's converged Gauss-Seidel solution ---------------
W = H = 40
mask = wl.blob_mask(W, H, 0.30, 12, seed=11)
guide, target = wl.synth_image(W, H, 3, seed=7), wl.synth_image(W, H, 3, seed=9)
mro, mci, mva, mb, pix, colors = wl.masked_poisson_system(mask, guide, target)
ref = pyoracle.Ref(2, "f64").import_csr(mva, mro[:-1], mci, len(pix))
out["masked_blend_40"] = {"W": W, "H": H, "coverage": 0.30, "thickness": 12, "mask_seed": 11, "guide_seed": 7,
                          "target_seed": 9, "n": int(len(pix)), "nnz": int(len(mva)),
                          "gs_default": [hx(ref.gauss_seidel(mb[c])) for c in range(3)],
                          "gs_25": hx(ref.gauss_seidel(mb[0], 0.0, 25))}

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, os.path.getsize(path), "bytes")
