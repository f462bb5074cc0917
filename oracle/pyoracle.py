"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle and the compiled reference.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package never imports this module.

  Oracle(...)  -> oracle/liboracle.so      this repo's C restatement (gs_oracle.c)
  Ref(...)     -> oracle/_ref/libgsref.so  the unmodified reference headers, compiled in
                                           place from /root/reference by oracle/Makefile
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgsref.so")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(quiet=True):
    """Compile liboracle.so and (only where /root/reference exists) _ref/libgsref.so."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def ref_available():
    return os.path.exists(REF_SO)


class _OrcMatrix(C.Structure):
    _fields_ = [("values", C.POINTER(C.c_double)), ("cols", C.POINTER(C.c_int)), ("store", C.c_int64),
                ("cap", C.c_int64), ("row_begin", C.POINTER(C.c_int)), ("row_nnz", C.POINTER(C.c_int)),
                ("row_left", C.POINTER(C.c_int)), ("n_rows", C.c_int), ("n_cols", C.c_int)]


_lib = None


def _oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        P = C.POINTER(_OrcMatrix)
        L.orc_new.restype = P
        L.orc_free.argtypes = [P]
        L.orc_init_from_vector.argtypes = [P, _i32p, _i32p, _f64p, C.c_int64]
        L.orc_init_dense.argtypes = [P, C.c_int, C.c_int, _f64p]
        L.orc_import_csr.argtypes = [P, _f64p, C.c_int, _i32p, C.c_int, _i32p, C.c_int, C.c_void_p]
        L.orc_at.argtypes = [P, C.c_int, C.c_int]
        L.orc_at.restype = C.c_double
        L.orc_insert.argtypes = [P, C.c_double, C.c_int, C.c_int]
        L.orc_gauss_seidel.argtypes = [P, _f64p, C.c_int64, C.c_double, C.c_int, _f64p, C.POINTER(C.c_int),
                                       C.POINTER(C.c_double)]
        L.orc_gauss_seidel_trace.argtypes = [P, _f64p, C.c_int64, _f64p, C.c_int, C.c_int, _f64p, _i32p, _f64p]
        L.orc_spmv.argtypes = [P, _f64p, _f64p]
        L.orc_l1_dist.argtypes = [_f64p, _f64p, C.c_int64]
        L.orc_l1_dist.restype = C.c_double
        L.orc_dot.argtypes = [_f64p, _f64p, C.c_int64]
        L.orc_dot.restype = C.c_double
        L.orc_cg.argtypes = [P, _f64p, C.c_int64, C.c_double, C.c_int, C.c_void_p, _f64p, C.POINTER(C.c_int)]
        L.orc_pcg.argtypes = [P, _f64p, C.c_int64, C.c_double, C.c_int, _f64p, C.POINTER(C.c_int)]
        L.orc_poisson_nnz.argtypes = [C.c_int, C.c_int]
        L.orc_poisson_nnz.restype = C.c_int64
        L.orc_poisson_csr.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p]
        L.orc_poisson_rhs.argtypes = [C.c_int, C.c_int, _f32p, _f32p, C.c_double, _f64p]
        L.orc_writeback_u8.argtypes = [_f64p, C.c_int64, _u8p]
        L.orc_gdf_gradients.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, _f32p, _f32p]
        L.orc_gdf_gradients.restype = C.c_int64
        L.orc_gdf_composite.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, _f64p]
        L.orc_gdf_composite.restype = C.c_int64
        L.orc_gdf_writeback.argtypes = [_f64p, C.c_int64, _u8p]
        L.orc_pano_mask_image.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _u8p]
        L.orc_pano_gradients.argtypes = [_u8p, C.c_int, C.c_int, _f32p, _f32p]
        L.orc_pano_gradients_masked.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _f32p, _f32p]
        L.orc_pano_gradients_masked.restype = None
        L.orc_pano_merge2_f32.argtypes = [_f32p, _f32p, _u8p, _u8p, _u8p, C.c_int, C.c_int]
        L.orc_pano_merge_u8.argtypes = [_u8p, _u8p, _u8p, _u8p, C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_pano_enforce_gradient_bound.argtypes = [_f32p, _f32p, _u8p, _u8p, C.c_int, C.c_int]
        for fn in (L.orc_pano_mask_image, L.orc_pano_gradients, L.orc_pano_merge2_f32, L.orc_pano_merge_u8,
                   L.orc_pano_enforce_gradient_bound):
            fn.restype = None
        _lib = L
    return _lib


def _a(x, dt):
    return np.ascontiguousarray(x, dtype=dt)


class Layout:
    """The five reference arrays, as numpy copies."""

    def __init__(self, values, cols, row_begin, row_nnz, row_left, n_rows, n_cols):
        self.values, self.cols = values, cols
        self.row_begin, self.row_nnz, self.row_left = row_begin, row_nnz, row_left
        self.n_rows, self.n_cols = n_rows, n_cols

    def same(self, o, values_dtype=None):
        v0, v1 = self.values, o.values
        if values_dtype is not None:
            v0, v1 = v0.astype(values_dtype), v1.astype(values_dtype)
        return (self.n_rows == o.n_rows and self.n_cols == o.n_cols and np.array_equal(v0, v1)
                and np.array_equal(self.cols, o.cols) and np.array_equal(self.row_begin, o.row_begin)
                and np.array_equal(self.row_nnz, o.row_nnz) and np.array_equal(self.row_left, o.row_left))


class Oracle:
    """This repo's C restatement (values held as double for T=int and T=double alike)."""

    def __init__(self):
        self.L = _oracle()
        self.m = self.L.orc_new()

    def __del__(self):
        try:
            self.L.orc_free(self.m)
        except Exception:
            pass

    def init_from_vector(self, rows, cols, vals):
        rows, cols, vals = _a(rows, np.int32), _a(cols, np.int32), _a(vals, np.float64)
        rc = self.L.orc_init_from_vector(self.m, rows, cols, vals, len(rows))
        if rc:
            raise ValueError("orc_init_from_vector rc=%d" % rc)
        return self

    def init_dense(self, n_rows, n_cols, dense):
        rc = self.L.orc_init_dense(self.m, n_rows, n_cols, _a(np.ravel(dense), np.float64))
        if rc:
            raise ValueError("orc_init_dense rc=%d" % rc)
        return self

    def import_csr(self, values, row_off, col_idx, n_cols, nnz_per_row=None, n_values=None):
        values, row_off, col_idx = _a(values, np.float64), _a(row_off, np.int32), _a(col_idx, np.int32)
        nv = len(values) if n_values is None else n_values
        nz = None
        if nnz_per_row is not None:
            self._nz = _a(nnz_per_row, np.int32)
            nz = self._nz.ctypes.data
        rc = self.L.orc_import_csr(self.m, values, nv, row_off, len(row_off), col_idx, n_cols, nz)
        if rc:
            raise ValueError("orc_import_csr rc=%d" % rc)
        return self

    def layout(self):
        s = self.m.contents
        n, nr = s.store, s.n_rows
        g = lambda p, k, dt: np.ctypeslib.as_array(p, shape=(max(k, 1),))[:k].astype(dt).copy()
        return Layout(g(s.values, n, np.float64), g(s.cols, n, np.int32), g(s.row_begin, nr, np.int32),
                      g(s.row_nnz, nr, np.int32), g(s.row_left, nr, np.int32), s.n_rows, s.n_cols)

    @property
    def n_rows(self):
        return self.m.contents.n_rows

    @property
    def n_cols(self):
        return self.m.contents.n_cols

    def at(self, r, c):
        return self.L.orc_at(self.m, r, c)

    def insert(self, v, r, c):
        self.L.orc_insert(self.m, float(v), r, c)

    def dense(self):
        return np.array([[self.at(i, j) for j in range(self.n_cols)] for i in range(self.n_rows)])

    def gauss_seidel(self, b, eps=1e-6, max_iter=1000):
        b = _a(b, np.float64)
        x = np.empty_like(b)
        sw, le = C.c_int(0), C.c_double(0)
        self.L.orc_gauss_seidel(self.m, b, len(b), eps, max_iter, x, C.byref(sw), C.byref(le))
        return x, sw.value, le.value

    def gauss_seidel_trace(self, b, eps_list, max_iter):
        """One run, snapshots at every threshold of the descending `eps_list`:
        -> x (len(eps_list), n), sweeps (len(eps_list),), eps_hist (sweeps run,)"""
        b, eps_list = _a(b, np.float64), _a(eps_list, np.float64)
        assert np.all(np.diff(eps_list) <= 0)
        xs = np.empty((len(eps_list), len(b)), np.float64)
        sw = np.zeros(len(eps_list), np.int32)
        hist = np.zeros(max(max_iter, 1), np.float64)
        self.L.orc_gauss_seidel_trace(self.m, b, len(b), eps_list, len(eps_list), max_iter, xs.reshape(-1), sw, hist)
        return xs, sw, hist[:int(sw.max())]

    def spmv(self, v):
        v = _a(v, np.float64)
        out = np.zeros(self.n_rows, np.float64)
        self.L.orc_spmv(self.m, v, out)
        return out

    def cg(self, b, eps=1e-16, max_iter=1000, x0=None):
        b = _a(b, np.float64)
        x = np.empty_like(b)
        it = C.c_int(0)
        p0 = None
        if x0 is not None:
            self._x0 = _a(x0, np.float64)
            p0 = self._x0.ctypes.data
        self.L.orc_cg(self.m, b, len(b), eps, max_iter, p0, x, C.byref(it))
        return x, it.value

    def pcg(self, b, eps=1e-16, max_iter=180):
        b = _a(b, np.float64)
        x = np.empty_like(b)
        it = C.c_int(0)
        self.L.orc_pcg(self.m, b, len(b), eps, max_iter, x, C.byref(it))
        return x, it.value


def l1_dist(a, b):
    a, b = _a(a, np.float64), _a(b, np.float64)
    return _oracle().orc_l1_dist(a, b, len(a))


def poisson_csr(W, H):
    """Closed-form A^T*A of the reference's forward-difference system (compressed CSR)."""
    L = _oracle()
    nnz = max(int(L.orc_poisson_nnz(W, H)), 1)  # the closed form holds for W,H >= 2; 1-pixel-wide grids keep only the pin
    ro = np.empty(W * H + 1, np.int32)
    ci = np.empty(nnz, np.int32)
    va = np.empty(nnz, np.float64)
    L.orc_poisson_csr(W, H, ro, ci, va)
    k = int(ro[-1])
    return ro, ci[:k].copy(), va[:k].copy()


def poisson_rhs(W, H, gx, gy, constraint):
    gx, gy = _a(gx, np.float32), _a(gy, np.float32)
    b = np.empty(W * H, np.float64)
    _oracle().orc_poisson_rhs(W, H, gx, gy, float(constraint), b)
    return b


def writeback_u8(x):
    x = _a(x, np.float64)
    out = np.empty(len(x), np.uint8)
    _oracle().orc_writeback_u8(x, len(x), out)
    return out


def gdf_gradients(images, labels):
    """PhotoMontage.cpp:399-425.  images (n, H, W, 3) uint8, labels (H, W) uint8 -> gx, gy (3, H, W) float32."""
    images, labels = _a(images, np.uint8), _a(labels, np.uint8)
    n, H, W, _ = images.shape
    gx, gy = np.empty((3, H, W), np.float32), np.empty((3, H, W), np.float32)
    bad = _oracle().orc_gdf_gradients(images.reshape(-1), n, labels.reshape(-1), W, H, gx.reshape(-1), gy.reshape(-1))
    if bad:
        raise ValueError("%d labels out of range" % bad)
    return gx, gy


def gdf_composite(images, labels):
    """PhotoMontage.cpp:599-610 -> (3, H*W) float64"""
    images, labels = _a(images, np.uint8), _a(labels, np.uint8)
    n, H, W, _ = images.shape
    x0 = np.empty((3, H * W), np.float64)
    bad = _oracle().orc_gdf_composite(images.reshape(-1), n, labels.reshape(-1), W, H, x0.reshape(-1))
    if bad:
        raise ValueError("%d labels out of range" % bad)
    return x0


def gdf_writeback(x, H, W):
    """PhotoMontage.cpp:617-626: (3, H*W) float64 -> (H, W, 3) uint8"""
    x = _a(x, np.float64).reshape(-1)
    out = np.empty(H * W * 3, np.uint8)
    _oracle().orc_gdf_writeback(x, H * W, out)
    return out.reshape(H, W, 3)


# ---- lab8 panorama right-hand-side producers (hw8_pa.cc:338-498, :604-636); arrays in cv::Mat layouts ----------
def pano_mask_image(src, mask):
    src, mask = _a(src, np.uint8), _a(mask, np.uint8)
    H, W = mask.shape
    out = np.empty_like(src)
    _oracle().orc_pano_mask_image(src.reshape(-1), mask.reshape(-1), W, H, out.reshape(-1))
    return out


def pano_gradients(img):
    img = _a(img, np.uint8)
    H, W, _ = img.shape
    gx, gy = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.float32)
    _oracle().orc_pano_gradients(img.reshape(-1), W, H, gx.reshape(-1), gy.reshape(-1))
    return gx, gy


def pano_gradients_masked(img, mask):
    """struct Gradients, second constructor (hw8_pa.cc:638-676)"""
    img, mask = _a(img, np.uint8), _a(mask, np.uint8)
    H, W, _ = img.shape
    gx, gy = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.float32)
    _oracle().orc_pano_gradients_masked(img.reshape(-1), mask.reshape(-1), W, H, gx.reshape(-1), gy.reshape(-1))
    return gx, gy


def pano_merge2_f32(target, src, target_mask, src_outer_mask, src_inner_mask):
    """MergeImage2<float>: returns the updated copy of `target` (H, W, 3) float32."""
    out = _a(target, np.float32).copy()
    H, W, _ = out.shape
    _oracle().orc_pano_merge2_f32(out.reshape(-1), _a(src, np.float32).reshape(-1), _a(target_mask, np.uint8).reshape(-1),
                                  _a(src_outer_mask, np.uint8).reshape(-1), _a(src_inner_mask, np.uint8).reshape(-1),
                                  W, H)
    return out


def pano_merge_u8(target, src, target_mask, src_mask, skip_how_many):
    """MergeImage<uchar, channel>: channel = 3 for (H, W, 3) images, 1 for (H, W) masks; returns the updated copy."""
    out = _a(target, np.uint8).copy()
    H, W = out.shape[:2]
    ch = 3 if out.ndim == 3 else 1
    _oracle().orc_pano_merge_u8(out.reshape(-1), _a(src, np.uint8).reshape(-1), _a(target_mask, np.uint8).reshape(-1),
                                _a(src_mask, np.uint8).reshape(-1), ch, float(skip_how_many), W, H)
    return out


def pano_merge_step(raw, dx, dy, mask, warped, erode_mask, erode_mask2):
    """One iteration of the stitch loop, hw8_pa.cc:740-768, after its warps -> new (raw, dx, dy, mask)."""
    gx, gy = pano_gradients(pano_mask_image(warped, erode_mask2))
    dx = pano_merge2_f32(dx, gx, mask, erode_mask2, erode_mask)
    dy = pano_merge2_f32(dy, gy, mask, erode_mask2, erode_mask)
    raw = pano_merge_u8(raw, warped, mask, erode_mask, 1)
    mask = pano_merge_u8(mask, erode_mask, mask, erode_mask, 0)
    return raw, dx, dy, mask


def pano_enforce_gradient_bound(dx, dy, src, mask):
    dx, dy = _a(dx, np.float32).copy(), _a(dy, np.float32).copy()
    H, W, _ = dx.shape
    _oracle().orc_pano_enforce_gradient_bound(dx.reshape(-1), dy.reshape(-1), _a(src, np.uint8).reshape(-1),
                                              _a(mask, np.uint8).reshape(-1), W, H)
    return dx, dy


# --------------------------------------------------------------------------------------
# The compiled reference's lab8 producers (oracle/_ref/libpanoref.so: hw8_pa.cc:316-498 and :602-676, unmodified,
# over a stand-in for cv::Mat whose images carry zero guard rows -- see ref_pano_shim.cc)
# --------------------------------------------------------------------------------------
PANO_REF_SO = os.path.join(_HERE, "_ref", "libpanoref.so")
_pref = None


def pano_ref_available():
    return os.path.exists(PANO_REF_SO)


def _panoref():
    global _pref
    if _pref is None:
        if not os.path.exists(PANO_REF_SO):
            raise RuntimeError("oracle/_ref/libpanoref.so missing: run `make -C oracle` where /root/reference exists")
        R = C.CDLL(PANO_REF_SO)
        R.ref_pano_mask_image.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _u8p]
        R.ref_pano_gradients.argtypes = [_u8p, C.c_void_p, C.c_int, C.c_int, _f32p, _f32p]
        R.ref_pano_merge2_f32.argtypes = [_f32p, _f32p, _u8p, _u8p, _u8p, C.c_int, C.c_int]
        R.ref_pano_merge_u8.argtypes = [_u8p, _u8p, _u8p, _u8p, C.c_int, C.c_double, C.c_int, C.c_int]
        R.ref_pano_enforce_gradient_bound.argtypes = [_f32p, _f32p, _u8p, _u8p, C.c_int, C.c_int]
        for fn in (R.ref_pano_mask_image, R.ref_pano_gradients, R.ref_pano_merge2_f32, R.ref_pano_merge_u8,
                   R.ref_pano_enforce_gradient_bound):
            fn.restype = None
        _pref = R
    return _pref


def ref_pano_mask_image(src, mask):
    src, mask = _a(src, np.uint8), _a(mask, np.uint8)
    H, W = mask.shape
    out = np.empty_like(src)
    _panoref().ref_pano_mask_image(src.reshape(-1), mask.reshape(-1), W, H, out.reshape(-1))
    return out


def ref_pano_gradients(img, mask=None):
    """struct Gradients: first constructor, or (mask given) the mask-driven second one (hw8_pa.cc:638-676)."""
    img = _a(img, np.uint8)
    H, W, _ = img.shape
    gx, gy = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.float32)
    mp = None
    if mask is not None:
        mask = _a(mask, np.uint8)
        mp = mask.ctypes.data
    _panoref().ref_pano_gradients(img.reshape(-1), mp, W, H, gx.reshape(-1), gy.reshape(-1))
    return gx, gy


def ref_pano_merge2_f32(target, src, target_mask, src_outer_mask, src_inner_mask):
    out = _a(target, np.float32).copy()
    H, W, _ = out.shape
    _panoref().ref_pano_merge2_f32(out.reshape(-1), _a(src, np.float32).reshape(-1), _a(target_mask, np.uint8).reshape(-1),
                                   _a(src_outer_mask, np.uint8).reshape(-1), _a(src_inner_mask, np.uint8).reshape(-1), W, H)
    return out


def ref_pano_merge_u8(target, src, target_mask, src_mask, skip_how_many):
    out = _a(target, np.uint8).copy()
    H, W = out.shape[:2]
    ch = 3 if out.ndim == 3 else 1
    _panoref().ref_pano_merge_u8(out.reshape(-1), _a(src, np.uint8).reshape(-1), _a(target_mask, np.uint8).reshape(-1),
                                 _a(src_mask, np.uint8).reshape(-1), ch, float(skip_how_many), W, H)
    return out


def ref_pano_enforce_gradient_bound(dx, dy, src, mask):
    dx, dy = _a(dx, np.float32).copy(), _a(dy, np.float32).copy()
    H, W, _ = dx.shape
    _panoref().ref_pano_enforce_gradient_bound(dx.reshape(-1), dy.reshape(-1), _a(src, np.uint8).reshape(-1),
                                               _a(mask, np.uint8).reshape(-1), W, H)
    return dx, dy


# --------------------------------------------------------------------------------------
# The compiled reference (oracle/_ref/libgsref.so)
# --------------------------------------------------------------------------------------
_ref = None


def _reflib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libgsref.so missing: run `make -C oracle` where /root/reference exists")
        _ref = C.CDLL(REF_SO)
    return _ref


class Ref:
    """One reference SparseMatrix<T>.  version: 1 (lab3 header) or 2 (lab8/project header);
    dtype: 'i32' (SparseMatrix<int>) or 'f64' (SparseMatrix<double>)."""

    def __init__(self, version=2, dtype="f64"):
        self.R = _reflib()
        self.v, self.sfx = version, dtype
        self.np_t = np.int32 if dtype == "i32" else np.float64
        self.c_t = C.c_int if dtype == "i32" else C.c_double
        self._tp = _i32p if dtype == "i32" else _f64p
        new = self._f("new")
        new.restype = C.c_void_p
        self.h = C.c_void_p(new())

    def _f(self, name):
        return getattr(self.R, "ref%d_%s_%s" % (self.v, name, self.sfx))

    def __del__(self):
        try:
            f = self._f("free")
            f.argtypes = [C.c_void_p]
            f(self.h)
        except Exception:
            pass

    def init_from_vector(self, rows, cols, vals):
        f = self._f("init_from_vector")
        f.argtypes = [C.c_void_p, _i32p, _i32p, self._tp, C.c_int64]
        rows, cols, vals = _a(rows, np.int32), _a(cols, np.int32), _a(vals, self.np_t)
        f(self.h, rows, cols, vals, len(rows))
        return self

    def init_dense(self, n_rows, n_cols, dense):
        r, c = np.divmod(np.arange(n_rows * n_cols, dtype=np.int32), n_cols)
        return self.init_from_vector(r, c, np.ravel(dense))

    def import_csr(self, values, row_off, col_idx, n_cols, nnz_per_row=None, n_values=None):
        assert self.v == 2
        f = self._f("import_csr")
        f.argtypes = [C.c_void_p, self._tp, C.c_int, _i32p, C.c_int, _i32p, C.c_int, C.c_void_p, C.c_int]
        values, row_off, col_idx = _a(values, self.np_t), _a(row_off, np.int32), _a(col_idx, np.int32)
        nv = len(values) if n_values is None else n_values
        nz, nnz = None, 0
        if nnz_per_row is not None:
            self._nz = _a(nnz_per_row, np.int32)
            nz, nnz = self._nz.ctypes.data, len(self._nz)
        f(self.h, values, nv, row_off, len(row_off), col_idx, n_cols, nz, nnz)
        return self

    def shape(self):
        f = self._f("shape")
        f.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        s, nr, nc = C.c_int64(), C.c_int(), C.c_int()
        f(self.h, C.byref(s), C.byref(nr), C.byref(nc))
        return s.value, nr.value, nc.value

    def layout(self):
        f = self._f("layout_sizes")
        f.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 5
        sz = [C.c_int64() for _ in range(5)]
        f(self.h, *[C.byref(s) for s in sz])
        nv, ncol, nrb, nnz, nleft = [s.value for s in sz]
        vals = np.zeros(max(nv, 1), self.np_t)
        cols = np.zeros(max(ncol, 1), np.int32)
        rb = np.zeros(max(nrb, 1), np.int32)
        rn = np.zeros(max(nnz, 1), np.int32)
        rl = np.zeros(max(nleft, 1), np.int32)
        g = self._f("layout")
        g.argtypes = [C.c_void_p, self._tp, _i32p, _i32p, _i32p, _i32p]
        g(self.h, vals, cols, rb, rn, rl)
        _, nr, nc = self.shape()
        return Layout(vals[:nv], cols[:ncol], rb[:nrb], rn[:nnz], rl[:nleft], nr, nc)

    def at(self, r, c):
        f = self._f("at")
        f.argtypes = [C.c_void_p, C.c_int, C.c_int]
        f.restype = self.c_t
        return f(self.h, r, c)

    def insert(self, v, r, c):
        f = self._f("insert")
        f.argtypes = [C.c_void_p, self.c_t, C.c_int, C.c_int]
        f(self.h, v, r, c)

    def dense(self):
        _, nr, nc = self.shape()
        out = np.zeros(max(nr * nc, 1), self.np_t)
        f = self._f("dense")
        f.argtypes = [C.c_void_p, self._tp]
        f(self.h, out)
        return out[:nr * nc].reshape(nr, nc)

    def gauss_seidel(self, b, eps=1e-6, max_iter=1000):
        f = self._f("gauss_seidel")
        f.argtypes = [C.c_void_p, _f64p, C.c_int64, C.c_double, C.c_int, _f64p]
        b = _a(b, np.float64)
        x = np.empty_like(b)
        f(self.h, b, len(b), eps, max_iter, x)
        return x

    def spmv(self, v, n_out=None):
        f = self._f("apply")
        f.argtypes = [C.c_void_p, _f64p, C.c_int64, _f64p, C.c_int64]
        v = _a(v, np.float64)
        n_out = self.shape()[1] if n_out is None else n_out
        out = np.zeros(n_out, np.float64)
        f(self.h, v, len(v), out, n_out)
        return out

    def cg(self, b, eps=1e-16, max_iter=1000, x0=None):
        b = _a(b, np.float64)
        x = np.empty_like(b)
        if self.v == 2:
            f = self._f("cg_init")
            f.argtypes = [C.c_void_p, _f64p, C.c_int64, C.c_double, C.c_int, C.c_void_p, _f64p]
            p0 = None
            if x0 is not None:
                self._x0 = _a(x0, np.float64)
                p0 = self._x0.ctypes.data
            f(self.h, b, len(b), eps, max_iter, p0, x)
        else:
            assert x0 is None
            f = self._f("cg")
            f.argtypes = [C.c_void_p, _f64p, C.c_int64, C.c_double, C.c_int, _f64p]
            f(self.h, b, len(b), eps, max_iter, x)
        return x

    def pcg(self, b, eps=1e-16, max_iter=180):
        assert self.v == 2 and self.sfx == "f64"
        f = self.R.ref2_pcg_f64
        f.argtypes = [C.c_void_p, _f64p, C.c_int64, C.c_double, C.c_int, _f64p]
        b = _a(b, np.float64)
        x = np.empty_like(b)
        f(self.h, b, len(b), eps, max_iter, x)
        return x


def ref_manhatton(a, b, version=2):
    f = getattr(_reflib(), "ref%d_manhatton_dist" % version)
    f.argtypes = [_f64p, _f64p, C.c_int64]
    f.restype = C.c_double
    a, b = _a(a, np.float64), _a(b, np.float64)
    return f(a, b, len(a))
