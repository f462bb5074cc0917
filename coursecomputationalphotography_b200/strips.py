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


def strip_rhs(W, H, channels, y0, y1, seed=7):
    """A^T b for the strip's rows (host array, channels x n_local) from the synthetic two-exposure image."""
    from . import workloads as wl
    gx, gy, ya, pin = wl.strip_gradients(W, H, channels, y0, y1, seed=seed)
    gx, gy = np.ascontiguousarray(gx), np.ascontiguousarray(gy)
    b = np.empty(channels * W * (y1 - y0), np.float64)
    check(load().gsb_poisson_rhs_rows(W, H, y0, y1, channels, ptr(gx), ptr(gy), ptr(np.ascontiguousarray(pin)), ptr(b)),
          "gsb_poisson_rhs_rows")
    return b.reshape(channels, W * (y1 - y0))
