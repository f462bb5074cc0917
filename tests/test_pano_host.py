"""CPU checks of the lab8 panorama right-hand-side producers (SURVEY 8f row N3, hw8_pa.cc:338-498, :604-636):
the oracle restates the reference's pointer walks literally (flat buffers, reads past a row's end land in the next
row as on a continuous cv::Mat); the device kernels use bounded per-row bodies (csrc/gsb_pano_body.h).  Here the
bodies, compiled for the host, are held to the literal restatement bit for bit on masks built to hit every branch:
empty rows, runs that reach the row end, covered / uncovered targets, inner masks with holes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pano") / "libpano_body_host.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "cpp", "pano_body_host.cc")], check=True)
    L = C.CDLL(so)
    u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.host_pano_mask_image.argtypes = [u8, u8, C.c_int, C.c_int, u8]
    L.host_pano_gradients.argtypes = [u8, C.c_int, C.c_int, f32, f32]
    L.host_pano_gradients_masked.argtypes = [u8, u8, C.c_int, C.c_int, f32, f32]
    L.host_pano_gradients_masked.restype = None
    L.host_pano_merge2_f32.argtypes = [f32, f32, u8, u8, u8, C.c_int, C.c_int]
    L.host_pano_merge_u8.argtypes = [u8, u8, u8, u8, C.c_int, C.c_double, C.c_int, C.c_int]
    L.host_pano_enforce_gradient_bound.argtypes = [f32, f32, u8, u8, C.c_int, C.c_int]
    for f in (L.host_pano_mask_image, L.host_pano_gradients, L.host_pano_merge2_f32, L.host_pano_merge_u8,
              L.host_pano_enforce_gradient_bound):
        f.restype = None
    return L


def row_masks(rng, H, W, style):
    """Masks as the panorama produces them (one run per row, values 0 / 255) plus adversarial variants."""
    m = np.zeros((H, W), np.uint8)
    for i in range(H):
        kind = style if style != "mixed" else rng.choice(["run", "empty", "full", "to_end", "holes", "noise"])
        if kind == "run":
            a = rng.integers(0, W)
            b = rng.integers(a, W + 1)
            m[i, a:b] = 255
        elif kind == "full":
            m[i] = 255
        elif kind == "to_end":
            m[i, rng.integers(0, W):] = 255
        elif kind == "holes":
            m[i] = 255
            m[i, rng.integers(0, W, max(1, W // 5))] = 0
        elif kind == "noise":
            m[i] = rng.integers(0, 2, W) * 255
    return m


SHAPES = [(1, 1), (1, 9), (7, 1), (6, 8), (23, 31), (40, 64)]


@pytest.mark.parametrize("H,W", SHAPES)
@pytest.mark.parametrize("style", ["run", "mixed", "noise", "empty", "full"])
def test_merge_bodies_equal_the_literal_restatement(oracle_mod, host, H, W, style):
    rng = np.random.default_rng(H * 131 + W * 7 + len(style))
    for rep in range(6):
        tmask = row_masks(rng, H, W, "mixed" if rep % 2 else style)
        outer = row_masks(rng, H, W, style)
        inner = (outer & row_masks(rng, H, W, "holes")) if rep % 3 else row_masks(rng, H, W, "mixed")
        tgt = rng.standard_normal((H, W, 3)).astype(np.float32)
        src = rng.standard_normal((H, W, 3)).astype(np.float32)
        want = oracle_mod.pano_merge2_f32(tgt, src, tmask, outer, inner)
        got = tgt.copy()
        host.host_pano_merge2_f32(got.reshape(-1), src.reshape(-1), tmask.reshape(-1), outer.reshape(-1),
                                  inner.reshape(-1), W, H)
        assert np.array_equal(got, want)
        for skip in (0.0, 1.0, 10.0, 2.5):
            ti = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            si = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            want = oracle_mod.pano_merge_u8(ti, si, tmask, outer, skip)
            got = ti.copy()
            host.host_pano_merge_u8(got.reshape(-1), si.reshape(-1), tmask.reshape(-1), outer.reshape(-1), 3, skip, W, H)
            assert np.array_equal(got, want), skip
        # MergeImage<uchar, 1>(mask, erode_mask, mask, erode_mask, 0): target and target mask are the same buffer
        want = oracle_mod.pano_merge_u8(tmask, outer, tmask, outer, 0.0)
        got = tmask.copy()
        host.host_pano_merge_u8(got.reshape(-1), outer.reshape(-1), got.reshape(-1), outer.reshape(-1), 1, 0.0, W, H)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("H,W", SHAPES)
def test_pixel_bodies_equal_the_literal_restatement(oracle_mod, host, H, W):
    rng = np.random.default_rng(H * 17 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    mask = row_masks(rng, H, W, "mixed")
    out = np.empty_like(img)
    host.host_pano_mask_image(img.reshape(-1), mask.reshape(-1), W, H, out.reshape(-1))
    assert np.array_equal(out, oracle_mod.pano_mask_image(img, mask))
    assert np.array_equal(out, np.where(mask[..., None] != 0, img, 0))
    gx, gy = np.full((H, W, 3), np.nan, np.float32), np.full((H, W, 3), np.nan, np.float32)
    host.host_pano_gradients(img.reshape(-1), W, H, gx.reshape(-1), gy.reshape(-1))
    ox, oy = oracle_mod.pano_gradients(img)
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)
    # the interleaved gradients are the planar ones of the fusion driver (same GradientAt), transposed
    px, py = oracle_mod.gdf_gradients(img[None], np.zeros((H, W), np.uint8))
    assert np.array_equal(np.moveaxis(px, 0, 2), ox) and np.array_equal(np.moveaxis(py, 0, 2), oy)
    dx = rng.standard_normal((H, W, 3)).astype(np.float32)
    dy = rng.standard_normal((H, W, 3)).astype(np.float32)
    for style in ("mixed", "full", "noise"):
        bound = row_masks(rng, H, W, style)
        wx, wy = oracle_mod.pano_enforce_gradient_bound(dx, dy, img, bound)
        hx, hy = dx.copy(), dy.copy()
        host.host_pano_enforce_gradient_bound(hx.reshape(-1), hy.reshape(-1), img.reshape(-1), bound.reshape(-1), W, H)
        assert np.array_equal(hx, wx) and np.array_equal(hy, wy)


def _erode_cross(m):
    k = m > 0
    e = k.copy()
    e[1:, :] &= k[:-1, :]
    e[:-1, :] &= k[1:, :]
    e[:, 1:] &= k[:, :-1]
    e[:, :-1] &= k[:, 1:]
    e[0, :] = e[-1, :] = False
    e[:, 0] = e[:, -1] = False
    return (e * 255).astype(np.uint8)


def test_restated_stitch_flow_removes_the_seam(oracle_mod):
    """The stage's purpose, in oracle terms: two exposures of one scene stitched by the restated merge step
    (hw8_pa.cc:722-788 without the OpenCV warps) and fused by the reference's own solver call
    (conjugateGradient, 50 iterations from the raw composite, hw8_pa.cc:808-810, :972) -- the exposure step at the
    seam shrinks to the level of the image's own gradients.  Pins that the restatement does what the reference
    does it for, not only that two of my formulations agree."""
    from coursecomputationalphotography_b200 import workloads as wl
    H, W = 96, 160
    scene = np.moveaxis(wl.synth_image(W, H, 3, seed=4), 0, 2).astype(np.float64)
    left = np.zeros((H, W), np.uint8)
    left[8:H - 6, 5:100] = 255
    right = np.zeros((H, W), np.uint8)
    right[4:H - 10, 70:W - 4] = 255
    im1 = np.where(left[..., None] > 0, np.clip(scene, 0, 255), 0).astype(np.uint8)
    im2 = np.where(right[..., None] > 0, np.clip(scene * 0.8 + 20, 0, 255), 0).astype(np.uint8)
    raw, mask = im1.copy(), left.copy()
    dx, dy = oracle_mod.pano_gradients(im1)
    e2 = _erode_cross(_erode_cross(right))
    e1 = _erode_cross(_erode_cross(e2))
    raw, dx, dy, mask = oracle_mod.pano_merge_step(raw, dx, dy, mask, im2, e1, e2)
    assert (mask > 0).sum() > (left > 0).sum()
    dx, dy = oracle_mod.pano_enforce_gradient_bound(dx, dy, raw, mask - _erode_cross(mask))
    ro, ci, va = oracle_mod.poisson_csr(W, H)
    m = oracle_mod.Oracle().import_csr(va, ro[:-1], ci, W * H)
    sol = np.empty((3, H * W))
    for c in range(3):
        b = oracle_mod.poisson_rhs(W, H, np.ascontiguousarray(dx[..., c]), np.ascontiguousarray(dy[..., c]),
                                   float(im1[0, 0, c]))
        sol[c], _ = m.cg(b, 1e-10, 50, x0=raw[..., c].astype(np.float64).ravel())
    out = oracle_mod.gdf_writeback(sol, H, W).astype(np.float64)
    rows = slice(20, 70)
    r = raw.astype(np.float64)
    jump_raw = np.abs(r[rows, 1:] - r[rows, :-1]).mean(axis=(0, 2))
    jump_out = np.abs(out[rows, 1:] - out[rows, :-1]).mean(axis=(0, 2))
    seam = int(np.argmax(jump_raw[60:110])) + 60
    assert jump_raw[seam] > 2.5 * np.median(jump_raw[10:150])  # the composite has a visible exposure step ...
    assert jump_out[seam] < 0.6 * jump_raw[seam]               # ... which the fusion takes most of the way down
