"""Gradient-domain fusion driver (SURVEY 8f, rows N2 + N4): Python mirror of the gsb_gdf_* entry points.

Names follow the reference (project/src/PhotoMontage/PhotoMontage.cpp): `GradientAt` + label pick (:399-425),
the composite initial guess (:599-610), `SolveChannel` (:535-628; lab8: hw8_pa.cc:902-986) and
`BuildSolveGradientFusion` (:410-433).  Arrays use the reference's cv::Mat layouts: images (n, H, W, 3) uint8,
labels (H, W) uint8, gradients (3, H, W) float32, result (H, W, 3) uint8.  No CPU path: everything runs in
libgsb200.so on the device.
"""
import ctypes as C

import numpy as np

from ._lib import GDF_CG, GDF_GS, GdfOptions, GdfStats, check, load, ptr

__all__ = ["gdf_options", "gdf_gradients", "gdf_composite", "SolveChannels", "BuildSolveGradientFusion", "GDF_GS",
           "GDF_CG", "GdfOptions", "GdfStats"]


def gdf_options(**kw):
    """Defaults of gsb_gdf_default_options (GS, eps 1e-6, 1000 sweeps); keyword overrides, `gs=dict(...)` for the
    nested Gauss-Seidel options."""
    o = GdfOptions()
    load().gsb_gdf_default_options(C.byref(o))
    for k, v in kw.items():
        if k == "gs":
            for k2, v2 in v.items():
                setattr(o.gs, k2, v2)
        else:
            setattr(o, k, v)
    return o


def _images(images, labels):
    images = np.ascontiguousarray(images, np.uint8)
    labels = np.ascontiguousarray(labels, np.uint8)
    if images.ndim != 4 or images.shape[3] != 3 or labels.shape != images.shape[1:3]:
        raise ValueError("images must be (n, H, W, 3) uint8 and labels (H, W) uint8")
    return images, labels


def gdf_gradients(images, labels):
    """GradientAt of the labelled image per pixel (PhotoMontage.cpp:399-425) -> gx, gy as (3, H, W) float32."""
    images, labels = _images(images, labels)
    n, H, W, _ = images.shape
    gx, gy = np.empty((3, H, W), np.float32), np.empty((3, H, W), np.float32)
    check(load().gsb_gdf_gradients(ptr(images), n, ptr(labels), W, H, ptr(gx), ptr(gy)), "gsb_gdf_gradients")
    return gx, gy


def gdf_composite(images, labels):
    """fast_init_value (PhotoMontage.cpp:599-610): (3, H*W) float64 composite of the labelled images."""
    images, labels = _images(images, labels)
    n, H, W, _ = images.shape
    x0 = np.empty((3, H * W), np.float64)
    check(load().gsb_gdf_composite(ptr(images), n, ptr(labels), W, H, ptr(x0)), "gsb_gdf_composite")
    return x0


def SolveChannels(gx, gy, constraint, init=None, options=None):
    """The three SolveChannel calls of hw8_pa.cc:808-810 / PhotoMontage.cpp:428-433 with the gradients given.
    gx, gy: (3, H, W) float32; constraint: 3 values; init: (3, H*W) float64 or None.
    Returns (result (H, W, 3) uint8, GdfStats)."""
    gx, gy = np.ascontiguousarray(gx, np.float32), np.ascontiguousarray(gy, np.float32)
    if gx.ndim != 3 or gx.shape[0] != 3 or gy.shape != gx.shape:
        raise ValueError("gx, gy must be (3, H, W) float32")
    _, H, W = gx.shape
    c = np.ascontiguousarray(constraint, np.float64)
    if c.shape != (3,):
        raise ValueError("constraint must hold 3 values")
    ini = None
    if init is not None:
        ini = np.ascontiguousarray(init, np.float64)
        if ini.size != 3 * W * H:
            raise ValueError("init must hold 3 * W * H values")
    out = np.empty((H, W, 3), np.uint8)
    st = GdfStats()
    op = C.byref(options) if options is not None else None
    check(load().gsb_gdf_solve(W, H, ptr(gx), ptr(gy), ptr(c), ptr(ini), op, ptr(out), C.byref(st)), "gsb_gdf_solve")
    return out, st


def BuildSolveGradientFusion(images, labels, fast_init=False, options=None):
    """PhotoMontage.cpp:410-433: label map + source images -> fused image (H, W, 3) uint8, GdfStats."""
    images, labels = _images(images, labels)
    n, H, W, _ = images.shape
    out = np.empty((H, W, 3), np.uint8)
    st = GdfStats()
    op = C.byref(options) if options is not None else None
    check(load().gsb_gdf_fuse(ptr(images), n, ptr(labels), W, H, 1 if fast_init else 0, op, ptr(out), C.byref(st)),
          "gsb_gdf_fuse")
    return out, st
