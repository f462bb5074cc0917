"""Python mirror of the reference's `SparseMatrix<T, IndexType=int>` (v2 = labs/lab8/src/OpenCVHW1/sparse-matrix.h)
over the C ABI of libgsb200.so.  Same member names, argument meaning and error behaviour as the
reference; bulk work (assembly, import, Gauss-Seidel, SpMV, CG) runs on the B200, single-element
access/modify (`at`, `insert`) runs on the host copy of the five layout arrays exactly as the C++
drop-in header include/sparse-matrix.h does, and marks the device copy stale.

The C++ header is the drop-in for reference users; this module exists so that tests/ and bench.py
can drive the same C entry points from Python.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import GSB_F64, GSB_I32, GsOptions, GsStats, check, load, ptr


def manhattonDist(a, b):
    """v2 :45-49 (spelling is the reference's)."""
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    assert a.size == b.size, "dims of 2 vecs mismatch"
    out = C.c_double(0)
    check(load().gsb_l1_dist(ptr(a), ptr(b), a.size, C.byref(out)), "gsb_l1_dist")
    return out.value


def dotProd(a, b):
    """v2 :58-63"""
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    out = C.c_double(0)
    check(load().gsb_dot(ptr(a), ptr(b), a.size, C.byref(out)), "gsb_dot")
    return out.value


def veclen2(a):
    """v2 :51-55"""
    return dotProd(a, a)


def vecadd(a, b, scale_b):
    """v2 :81-85: a + scale_b * b"""
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    out = np.empty_like(a)
    check(load().gsb_axpy(ptr(a), ptr(b), float(scale_b), a.size, ptr(out)), "gsb_axpy")
    return out


def vecsub(a, b):
    """v2 :75-79"""
    return vecadd(a, b, -1.0)


def vecmul(a, b):
    """v2 :101-105"""
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    out = np.empty_like(a)
    check(load().gsb_vecmul(ptr(a), ptr(b), a.size, ptr(out)), "gsb_vecmul")
    return out


class SparseMatrix:
    """`dtype`: np.int32 (SparseMatrix<int>, lab3) or np.float64 (SparseMatrix<double>, lab8/project)."""

    def __init__(self, dtype=np.float64, device=None):
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.int32), np.dtype(np.float64)):
            raise TypeError("SparseMatrix supports int32 and float64 elements (the reference's two instantiations)")
        self.L = load()
        if device is not None:
            check(self.L.gsb_set_device(int(device)), "gsb_set_device")
        self._h = C.c_void_p()
        check(self.L.gsb_matrix_create(C.byref(self._h), GSB_I32 if self.dtype == np.int32 else GSB_F64),
              "gsb_matrix_create")
        # host copy of the reference's five arrays (values_, col_offset_, row_begin_, row_num_nze_, row_space_left_)
        self.values_ = np.zeros(0, self.dtype)
        self.col_offset_ = np.zeros(0, np.int32)
        self.row_begin_ = np.zeros(0, np.int32)
        self.row_num_nze_ = np.zeros(0, np.int32)
        self.row_space_left_ = np.zeros(0, np.int32)
        self.n_rows_ = 0
        self.n_cols_ = 0
        self._host_valid = False   # host arrays mirror the device
        self._device_stale = False  # host arrays were edited by insert()
        self.last_stats = None

    def __del__(self):
        try:
            if self._h:
                self.L.gsb_matrix_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- shape ------------------------------------------------------------------------------
    def rows(self):
        return self.n_rows_

    def cols(self):
        return self.n_cols_

    def _refresh_shape(self):
        store, nr, nc, nnz = C.c_int64(), C.c_int(), C.c_int(), C.c_int64()
        check(self.L.gsb_matrix_shape(self._h, C.byref(store), C.byref(nr), C.byref(nc), C.byref(nnz)),
              "gsb_matrix_shape")
        self.n_rows_, self.n_cols_, self._store, self._nnz = nr.value, nc.value, store.value, nnz.value
        self._host_valid = False
        self._device_stale = False

    def _pull(self):
        """Bring the five arrays to the host (lazy: bulk users never pay for it)."""
        if self._host_valid:
            return
        self.values_ = np.empty(self._store, self.dtype)
        self.col_offset_ = np.empty(self._store, np.int32)
        self.row_begin_ = np.empty(self.n_rows_, np.int32)
        self.row_num_nze_ = np.empty(self.n_rows_, np.int32)
        self.row_space_left_ = np.empty(self.n_rows_, np.int32)
        check(self.L.gsb_matrix_download(self._h, ptr(self.values_), ptr(self.col_offset_), ptr(self.row_begin_),
                                         ptr(self.row_num_nze_), ptr(self.row_space_left_)), "gsb_matrix_download")
        self._host_valid = True

    def _push(self):
        """Re-upload after host-side insert() edits (the device handle is rebuilt lazily)."""
        if not self._device_stale:
            return
        check(self.L.gsb_matrix_upload(self._h, ptr(self.values_), ptr(self.col_offset_), len(self.values_),
                                       ptr(self.row_begin_), ptr(self.row_num_nze_), ptr(self.row_space_left_),
                                       self.n_rows_, self.n_cols_), "gsb_matrix_upload")
        self._store = len(self.values_)
        self._device_stale = False

    def layout(self):
        """(values_, col_offset_, row_begin_, row_num_nze_, row_space_left_) as numpy arrays."""
        self._pull()
        return self.values_, self.col_offset_, self.row_begin_, self.row_num_nze_, self.row_space_left_

    # ---- assembly ---------------------------------------------------------------------------
    def initializeFromVector(self, rows, cols, vals):
        """v2 :265-319.  Sorted COO (row, col), explicit zeros allowed."""
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, self.dtype)
        if not (rows.size == cols.size == vals.size):
            raise ValueError("rows, cols, vals must have equal length")
        check(self.L.gsb_matrix_assemble_sorted_coo(self._h, ptr(rows), ptr(cols), ptr(vals), rows.size),
              "gsb_matrix_assemble_sorted_coo")
        self._refresh_shape()

    def initialize(self, row, col, x=None):
        """v2 :321-347.  With a dense row-major list: every entry (zeros too) becomes a COO item."""
        if x is None:
            # reference leaves row_num_nze_ unsized here; a usable empty matrix is kept instead
            self.n_rows_, self.n_cols_ = int(row), int(col)
            self.values_ = np.zeros(0, self.dtype)
            self.col_offset_ = np.zeros(0, np.int32)
            self.row_begin_ = np.zeros(self.n_rows_, np.int32)
            self.row_num_nze_ = np.zeros(self.n_rows_, np.int32)
            self.row_space_left_ = np.zeros(self.n_rows_, np.int32)
            self._store = 0
            self._host_valid, self._device_stale = True, True
            return
        x = np.ascontiguousarray(np.ravel(x), self.dtype)
        r, c = np.divmod(np.arange(int(row) * int(col), dtype=np.int32), np.int32(col))
        self.initializeFromVector(r, c, x)

    def initializeFromTriplets(self, rows, cols, vals):
        """v2 :249-263 (a working version: last duplicate wins, zeros dropped). Needs initialize(r, c) first."""
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, self.dtype)
        check(self.L.gsb_matrix_assemble_coo(self._h, ptr(rows), ptr(cols), ptr(vals), rows.size, self.n_rows_,
                                             self.n_cols_), "gsb_matrix_assemble_coo")
        self._refresh_shape()

    def initializeFromEigenRowMajor(self, values, n_values, row_offset, n_row_offset, col_offset, n_col_offset,
                                    non_zeros=None, n_non_zeros=0):
        """v2 :537-620, same argument order."""
        values = np.ascontiguousarray(values, self.dtype)
        row_offset = np.ascontiguousarray(row_offset, np.int32)
        col_offset = np.ascontiguousarray(col_offset, np.int32)
        nz = None if non_zeros is None else np.ascontiguousarray(non_zeros, np.int32)
        check(self.L.gsb_matrix_import_csr(self._h, ptr(values), int(n_values), ptr(row_offset), int(n_row_offset),
                                           ptr(col_offset), int(n_col_offset), ptr(nz), int(n_non_zeros)),
              "gsb_matrix_import_csr")
        self._refresh_shape()

    def poisson(self, W, H):
        """A9: the reference's A^T*A for a W x H image, built on the device (PhotoMontage.cpp:541-597)."""
        check(self.L.gsb_poisson_matrix(self._h, int(W), int(H)), "gsb_poisson_matrix")
        self._refresh_shape()

    # ---- access / modify (host side, as in the C++ header) -----------------------------------
    def _nearest(self, row, col):
        """v2 :627-645"""
        idx = int(self.row_begin_[row])
        end = idx + int(self.row_num_nze_[row]) - 1
        co = self.col_offset_
        if co[idx] == col:
            return idx
        while end > idx:
            mid = (end + idx) // 2
            if co[mid] < col:
                idx = mid + 1
            else:
                end = mid
        return idx

    def at(self, row, col):
        """v2 :162-173"""
        self._pull()
        if not self.row_num_nze_[row]:
            return self.dtype.type(0)
        idx = self._nearest(row, col)
        return self.values_[idx] if self.col_offset_[idx] == col else self.dtype.type(0)

    coeff = at  # v2 :176

    def at_many(self, rows, cols):
        """Batched at() on the device copy (gsb_matrix_at)."""
        self._push()
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        out = np.empty(rows.size, np.float64)
        check(self.L.gsb_matrix_at(self._h, ptr(rows), ptr(cols), rows.size, ptr(out)), "gsb_matrix_at")
        return out

    def insertZero(self, row, col):
        """v2 :183-201 with the column memmove sized by the index type (upstream uses sizeof(T))."""
        self._pull()
        nz = int(self.row_num_nze_[row])
        if nz == 0:
            return
        idx = self._nearest(row, col)
        if self.col_offset_[idx] != col:
            return
        end = int(self.row_begin_[row]) + nz
        self.values_[idx:end - 1] = self.values_[idx + 1:end].copy()
        self.col_offset_[idx:end - 1] = self.col_offset_[idx + 1:end].copy()
        self.row_num_nze_[row] -= 1
        self.row_space_left_[row] += 1
        self._device_stale = True

    def insertNoneZero(self, val, row, col):
        """v2 :203-237, keeping rows sorted (upstream mis-places a column larger than all live ones)."""
        self._pull()
        rb, nz = int(self.row_begin_[row]), int(self.row_num_nze_[row])
        idx = rb
        if nz:
            idx = self._nearest(row, col)
            if self.col_offset_[idx] == col:
                self.values_[idx] = val
                self._device_stale = True
                return
            if self.col_offset_[idx] < col:
                idx += 1
        end = rb + nz
        if self.row_space_left_[row]:
            self.row_space_left_[row] -= 1
            self.values_[idx + 1:end + 1] = self.values_[idx:end].copy()
            self.col_offset_[idx + 1:end + 1] = self.col_offset_[idx:end].copy()
            self.values_[idx] = val
            self.col_offset_[idx] = col
        else:
            self.values_ = np.insert(self.values_, idx, val)
            self.col_offset_ = np.insert(self.col_offset_, idx, col).astype(np.int32)
            self.row_begin_[row + 1:] += 1
        self.row_num_nze_[row] += 1
        self._device_stale = True

    def insert(self, val, row, col):
        """v2 :239-247"""
        if val == 0:
            self.insertZero(row, col)
        else:
            self.insertNoneZero(val, row, col)

    # ---- ordering ---------------------------------------------------------------------------
    def analyze(self, ordering=_lib.ORDER_AUTO, colors=None):
        self._push()
        c = None if colors is None else np.ascontiguousarray(colors, np.int32)
        check(self.L.gsb_matrix_analyze(self._h, int(ordering), ptr(c)), "gsb_matrix_analyze")
        return self.coloring()

    def coloring(self):
        nc, used, w = C.c_int(), C.c_int(), C.c_int()
        check(self.L.gsb_matrix_coloring(self._h, C.byref(nc), C.byref(used), C.byref(w)), "gsb_matrix_coloring")
        return {"n_colors": nc.value, "ordering": used.value, "grid_width": w.value}

    def ordering(self):
        perm = np.empty(self.n_rows_, np.int32)
        colors = np.empty(self.n_rows_, np.int32)
        check(self.L.gsb_matrix_ordering(self._h, ptr(perm), ptr(colors)), "gsb_matrix_ordering")
        return perm, colors

    # ---- solvers ----------------------------------------------------------------------------
    @staticmethod
    def options(**kw):
        o = GsOptions()
        load().gsb_gs_default_options(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def gaussSeidel(self, b, epsilon=1e-6, max_iteration=1000, options=None, x0=None):
        """v2 :350-380.  b: n doubles, or (nrhs, n) for up to 4 right-hand sides sharing the matrix pass.
        x0 is an extension (the reference always starts from 1.0)."""
        self._push()
        b = np.ascontiguousarray(b, np.float64)
        nrhs = 1 if b.ndim == 1 else b.shape[0]
        if b.size != nrhs * self.n_cols_:
            raise ValueError("len(b) must match matrix's column")  # the reference's debug-only assert, v2 :351
        x = np.empty_like(b)
        st = GsStats()
        op = C.byref(options) if options is not None else None
        if x0 is None:
            check(self.L.gsb_gauss_seidel(self._h, ptr(b), nrhs, float(epsilon), int(max_iteration), op, ptr(x),
                                          C.byref(st)), "gsb_gauss_seidel")
        else:
            x0 = np.ascontiguousarray(x0, np.float64)
            check(self.L.gsb_gauss_seidel_x0(self._h, ptr(b), ptr(x0), nrhs, float(epsilon), int(max_iteration), op,
                                             ptr(x), C.byref(st)), "gsb_gauss_seidel_x0")
        self.last_stats = st
        return x

    def gaussSeidel_dev(self, b_ptr, x_ptr, nrhs=1, epsilon=1e-6, max_iteration=1000, options=None):
        """Device-pointer form: b_ptr/x_ptr are integer addresses on this matrix's device."""
        self._push()
        st = GsStats()
        op = C.byref(options) if options is not None else None
        check(self.L.gsb_gauss_seidel_dev(self._h, C.c_void_p(b_ptr), nrhs, float(epsilon), int(max_iteration), op,
                                          C.c_void_p(x_ptr), C.byref(st)), "gsb_gauss_seidel_dev")
        self.last_stats = st
        return st

    def applyToVector(self, vin, out=None):
        """v2 :382-393"""
        self._push()
        vin = np.ascontiguousarray(vin, np.float64)
        if vin.size != self.n_cols_:
            raise ValueError("input length must equal cols()")
        if out is None:
            out = np.empty(self.n_rows_, np.float64)
        check(self.L.gsb_spmv(self._h, ptr(vin), ptr(out)), "gsb_spmv")
        return out

    def residual(self, b, x):
        self._push()
        b, x = np.ascontiguousarray(b, np.float64), np.ascontiguousarray(x, np.float64)
        out = C.c_double(0)
        check(self.L.gsb_residual_l2(self._h, ptr(b), ptr(x), C.byref(out)), "gsb_residual_l2")
        return out.value

    def conjugateGradient(self, b, epsilon=1e-16, max_iteration=1000, initialize=None):
        """v2 :396-434"""
        self._push()
        b = np.ascontiguousarray(b, np.float64)
        x = np.empty_like(b)
        it = C.c_int(0)
        x0 = None
        if initialize is not None and len(initialize):
            x0 = np.ascontiguousarray(initialize, np.float64)
        check(self.L.gsb_conjugate_gradient(self._h, ptr(b), float(epsilon), int(max_iteration), ptr(x0), ptr(x),
                                            C.byref(it)), "gsb_conjugate_gradient")
        self.last_iters = it.value
        return x

    def conjugateGradientMulti(self, b, epsilon=1e-16, max_iteration=1000, initialize=None, out=None):
        """EXTENSION: b (nrhs, n), nrhs <= 4: the right-hand sides share every pass over the matrix; each stops on its
        own.  self.last_iters = list of the nrhs loop counts.  `out`: a C-contiguous float64 array of b's shape to
        receive x (e.g. page-locked memory, so that the download runs at PCIe speed)."""
        self._push()
        b = np.ascontiguousarray(b, np.float64)
        nrhs = 1 if b.ndim == 1 else b.shape[0]
        if out is not None and (out.dtype != np.float64 or out.shape != b.shape or not out.flags.c_contiguous):
            raise ValueError("out must be a C-contiguous float64 array of the shape of b")
        x = np.empty_like(b) if out is None else out
        it = (C.c_int * 4)()
        x0 = None
        if initialize is not None and len(initialize):
            x0 = np.ascontiguousarray(initialize, np.float64)
            if x0.shape != b.shape:
                raise ValueError("initialize must have the shape of b")
        check(self.L.gsb_conjugate_gradient_multi(self._h, ptr(b), nrhs, float(epsilon), int(max_iteration), ptr(x0),
                                                  ptr(x), it), "gsb_conjugate_gradient_multi")
        self.last_iters = [it[i] for i in range(nrhs)]
        return x

    def conjugateGradientEigen(self, b, epsilon=1e-16, max_iteration=180):
        """v2 :494-535 (Jacobi-preconditioned CG)"""
        self._push()
        b = np.ascontiguousarray(b, np.float64)
        x = np.empty_like(b)
        it = C.c_int(0)
        check(self.L.gsb_conjugate_gradient_jacobi(self._h, ptr(b), float(epsilon), int(max_iteration), ptr(x),
                                                   C.byref(it)), "gsb_conjugate_gradient_jacobi")
        self.last_iters = it.value
        return x


def poisson_rhs(W, H, gx, gy, constraint):
    """A^T*b of the reference's SolveChannel (PhotoMontage.cpp:551-592) for 1..4 channels.
    gx, gy: (nch, H, W) or (H, W) float32; constraint: scalar or nch values."""
    gx, gy = np.ascontiguousarray(gx, np.float32), np.ascontiguousarray(gy, np.float32)
    nch = 1 if gx.ndim == 2 else gx.shape[0]
    c = np.ascontiguousarray(np.broadcast_to(np.asarray(constraint, np.float64), (nch,)))
    b = np.empty(nch * W * H, np.float64)
    check(load().gsb_poisson_rhs(W, H, nch, ptr(gx), ptr(gy), ptr(c), ptr(b)), "gsb_poisson_rhs")
    return b if nch == 1 else b.reshape(nch, W * H)


def writeback_u8(x):
    """A10: uchar(clamp(x, 0, 255)), truncating (PhotoMontage.cpp:617-626)"""
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty(x.size, np.uint8)
    check(load().gsb_writeback_u8(ptr(x), x.size, ptr(out)), "gsb_writeback_u8")
    return out.reshape(x.shape)
