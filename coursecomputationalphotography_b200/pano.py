"""lab8 panorama right-hand-side producers: Python mirror of the gsb_pano_* entry points.
Function names follow labs/lab8/src/OpenCVHW1/hw8_pa.cc; arrays use its cv::Mat layouts (images (H, W, 3) uint8,
gradients (H, W, 3) float32, masks (H, W) uint8).  In-place arguments of the reference are returned as new arrays.
No CPU path: everything runs in libgsb200.so on the device."""
import numpy as np

from ._lib import check, load, ptr

__all__ = ["MaskImage", "Gradients", "MergeImage2", "MergeImage", "EnforceGradientBound", "merge_step", "split_planes"]


def _img(a, dt=np.uint8):
    a = np.ascontiguousarray(a, dt)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected an (H, W, 3) array")
    return a


def _mask(a, shape):
    a = np.ascontiguousarray(a, np.uint8)
    if a.shape != tuple(shape):
        raise ValueError("mask must be (H, W) uint8 of the image's size")
    return a


def MaskImage(src, mask):
    """hw8_pa.cc:443-466"""
    src = _img(src)
    H, W, _ = src.shape
    mask = _mask(mask, (H, W))
    out = np.empty_like(src)
    check(load().gsb_pano_mask_image(ptr(src), ptr(mask), W, H, ptr(out)), "gsb_pano_mask_image")
    return out


def Gradients(img, mask=None):
    """struct Gradients, hw8_pa.cc:604-636 (m) and :638-676 (m, mask) -> (x, y), each (H, W, 3) float32"""
    img = _img(img)
    H, W, _ = img.shape
    gx, gy = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.float32)
    if mask is None:
        check(load().gsb_pano_gradients(ptr(img), W, H, ptr(gx), ptr(gy)), "gsb_pano_gradients")
    else:
        mask = _mask(mask, (H, W))
        check(load().gsb_pano_gradients_masked(ptr(img), ptr(mask), W, H, ptr(gx), ptr(gy)), "gsb_pano_gradients_masked")
    return gx, gy


def MergeImage2(target, src, target_mask, src_outer_mask, src_inner_mask):
    """MergeImage2<float>, hw8_pa.cc:338-385; returns the merged copy of `target`."""
    out = _img(target, np.float32).copy()
    src = _img(src, np.float32)
    H, W, _ = out.shape
    tm, so, si = (_mask(m, (H, W)) for m in (target_mask, src_outer_mask, src_inner_mask))
    check(load().gsb_pano_merge2_f32(ptr(out), ptr(src), ptr(tm), ptr(so), ptr(si), W, H), "gsb_pano_merge2_f32")
    return out


def MergeImage(target, src, target_mask, src_mask, SkipHowMany):
    """MergeImage<uchar, channel>, hw8_pa.cc:387-441; (H, W, 3) images or (H, W) masks; returns the merged copy."""
    out = np.ascontiguousarray(target, np.uint8).copy()
    src = np.ascontiguousarray(src, np.uint8)
    if out.shape != src.shape or out.ndim not in (2, 3) or (out.ndim == 3 and out.shape[2] != 3):
        raise ValueError("target and src must both be (H, W, 3) or both (H, W)")
    H, W = out.shape[:2]
    ch = 3 if out.ndim == 3 else 1
    tm, sm = _mask(target_mask, (H, W)), _mask(src_mask, (H, W))
    check(load().gsb_pano_merge_u8(ptr(out), ptr(src), ptr(tm), ptr(sm), ch, float(SkipHowMany), W, H),
          "gsb_pano_merge_u8")
    return out


def EnforceGradientBound(dx, dy, src, mask):
    """hw8_pa.cc:468-498; returns the updated copies of dx, dy."""
    dx, dy = _img(dx, np.float32).copy(), _img(dy, np.float32).copy()
    src = _img(src)
    H, W, _ = src.shape
    mask = _mask(mask, (H, W))
    check(load().gsb_pano_enforce_gradient_bound(ptr(dx), ptr(dy), ptr(src), ptr(mask), W, H),
          "gsb_pano_enforce_gradient_bound")
    return dx, dy


def merge_step(raw, dx, dy, mask, warped, erode_mask, erode_mask2):
    """One iteration of the stitch loop (hw8_pa.cc:740-768) after its warps; returns the new (raw, dx, dy, mask)."""
    raw, warped = _img(raw).copy(), _img(warped)
    dx, dy = _img(dx, np.float32).copy(), _img(dy, np.float32).copy()
    H, W, _ = raw.shape
    mask = _mask(mask, (H, W)).copy()
    e1, e2 = _mask(erode_mask, (H, W)), _mask(erode_mask2, (H, W))
    check(load().gsb_pano_merge_step(ptr(raw), ptr(dx), ptr(dy), ptr(mask), ptr(warped), ptr(e1), ptr(e2), W, H),
          "gsb_pano_merge_step")
    return raw, dx, dy, mask


def split_planes(interleaved):
    """(H, W, 3) float32 (CV_32FC3) -> (3, H, W): the layout gdf.SolveChannels / poisson_rhs take."""
    a = _img(interleaved, np.float32)
    H, W, _ = a.shape
    out = np.empty((3, H, W), np.float32)
    check(load().gsb_pano_split_planes_f32(ptr(a), W, H, ptr(out)), "gsb_pano_split_planes_f32")
    return out
