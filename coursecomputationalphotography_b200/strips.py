"""Row-strip multi-GPU solve: one rank per GPU over the gsb_dist_* C entry points.

The rendezvous (who is rank r, how the 128-byte NCCL unique id reaches every rank) belongs to the caller:
bench.py and the tests use torch.distributed for it; the data path (halo exchange per colour phase, the
stop-rule all-reduce) lives inside libgsb200.so and talks NCCL directly.
"""
import ctypes as C

import numpy as np

from ._lib import UNIQUE_ID_BYTES, GsStats, check, load, ptr


def make_unique_id():
    """Rank 0 only: a fresh NCCL unique id as bytes."""
    buf = (C.c_ubyte * UNIQUE_ID_BYTES)()
    check(load().gsb_dist_unique_id(buf), "gsb_dist_unique_id")
    return bytes(buf)


def broadcast_unique_id(dist, rank, device=None):
    """Create the id on rank 0 and broadcast it with torch.distributed (any backend)."""
    import torch
    if rank == 0:
        t = torch.tensor(list(make_unique_id()), dtype=torch.uint8)
    else:
        t = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


class StripSolver:
    """One rank of a row-strip Gauss-Seidel solve."""

    def __init__(self, unique_id, rank, world, device):
        self.L = load()
        self.rank, self.world, self.device = rank, world, device
        self._h = C.c_void_p()
        idbuf = (C.c_ubyte * UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        check(self.L.gsb_dist_init(C.byref(self._h), idbuf, rank, world, device), "gsb_dist_init")
        self.last_stats = None

    def close(self):
        if self._h:
            self.L.gsb_dist_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def poisson_strip(self, W, H, y0, y1):
        """Reference-faithful full-grid Poisson rows of image rows [y0, y1), generated on the device."""
        check(self.L.gsb_dist_poisson_strip(self._h, W, H, y0, y1), "gsb_dist_poisson_strip")
        self.n_local = W * (y1 - y0)

    def set_colors(self, colors):
        """The caller's two-colouring (0/1) of ALL global rows, instead of pixel parity (masked / compact systems)."""
        c = np.ascontiguousarray(colors, np.uint8)
        check(self.L.gsb_dist_set_colors(self._h, ptr(c), len(c)), "gsb_dist_set_colors")

    def matrix_rows(self, values, row_off, col_idx, row0, n_global, grid_width):
        values = np.ascontiguousarray(values, np.float64)
        row_off = np.ascontiguousarray(row_off, np.int32)
        col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.n_local = len(row_off) - 1
        check(self.L.gsb_dist_matrix_rows(self._h, ptr(values), ptr(row_off), ptr(col_idx), row0, self.n_local,
                                          n_global, grid_width), "gsb_dist_matrix_rows")

    def gauss_seidel_dev(self, b_ptr, x_ptr, nrhs=1, epsilon=1e-6, max_iteration=1000, options=None):
        st = GsStats()
        op = C.byref(options) if options is not None else None
        check(self.L.gsb_dist_gauss_seidel_dev(self._h, C.c_void_p(b_ptr), nrhs, float(epsilon), int(max_iteration),
                                               op, C.c_void_p(x_ptr), C.byref(st)), "gsb_dist_gauss_seidel_dev")
        self.last_stats = st
        return st

    def residual_dev(self, b_ptr, x_ptr):
        out = C.c_double(0)
        check(self.L.gsb_dist_residual_l2_dev(self._h, C.c_void_p(b_ptr), C.c_void_p(x_ptr), C.byref(out)),
              "gsb_dist_residual_l2_dev")
        return out.value


def set_devices(devices):
    """Device list of the host entry points (SparseMatrix.gaussSeidel): two or more -> row strips, one per device,
    from the one blocking call.  [] or one device: single-device solves."""
    arr = (C.c_int * max(len(devices), 1))(*devices)
    check(load().gsb_set_devices(arr, len(devices)), "gsb_set_devices")


def get_devices():
    arr = (C.c_int * 16)()
    n = load().gsb_get_devices(arr, 16)
    return [arr[i] for i in range(n)]


class LocalGroup:
    """Single-process multi-device solve (gsb_dist_init_local): N devices of this box, one row strip each, driven
    from the calling thread; host vectors in and out."""

    def __init__(self, devices):
        self.L = load()
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        check(self.L.gsb_dist_init_local(C.byref(self._h), arr, len(devices)), "gsb_dist_init_local")
        self.devices = list(devices)
        self.n = 0
        self.last_stats = None

    def close(self):
        if self._h:
            self.L.gsb_dist_group_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def matrix(self, sp):
        """Shard an assembled SparseMatrix by rows (two-colourable matrices)."""
        sp._push()
        check(self.L.gsb_dist_group_matrix(self._h, sp._h), "gsb_dist_group_matrix")
        self.n = sp.rows()

    def poisson(self, W, H):
        check(self.L.gsb_dist_group_poisson(self._h, W, H), "gsb_dist_group_poisson")
        self.n = W * H

    def gauss_seidel(self, b, epsilon=1e-6, max_iteration=1000, options=None, out=None):
        b = np.ascontiguousarray(b, np.float64)
        nrhs = 1 if b.ndim == 1 else b.shape[0]
        if b.size != nrhs * self.n:
            raise ValueError("len(b) must match matrix's column")
        x = np.empty_like(b) if out is None else out
        st = GsStats()
        op = C.byref(options) if options is not None else None
        check(self.L.gsb_dist_group_gauss_seidel(self._h, ptr(b), nrhs, float(epsilon), int(max_iteration), op, ptr(x),
                                                 C.byref(st)), "gsb_dist_group_gauss_seidel")
        self.last_stats = st
        return x

    def residual(self, b, x):
        out = C.c_double(0)
        check(self.L.gsb_dist_group_residual_l2(self._h, ptr(np.ascontiguousarray(b, np.float64)),
                                                ptr(np.ascontiguousarray(x, np.float64)), C.byref(out)),
              "gsb_dist_group_residual_l2")
        return out.value


def strip_rhs(W, H, channels, y0, y1, seed=7):
    """A^T b for the strip's rows (host array, channels x n_local) from the synthetic two-exposure image."""
    from . import workloads as wl
    gx, gy, ya, pin = wl.strip_gradients(W, H, channels, y0, y1, seed=seed)
    gx, gy = np.ascontiguousarray(gx), np.ascontiguousarray(gy)
    b = np.empty(channels * W * (y1 - y0), np.float64)
    check(load().gsb_poisson_rhs_rows(W, H, y0, y1, channels, ptr(gx), ptr(gy), ptr(np.ascontiguousarray(pin)), ptr(b)),
          "gsb_poisson_rhs_rows")
    return b.reshape(channels, W * (y1 - y0))
