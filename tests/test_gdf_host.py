"""CPU checks of the gradient-domain-fusion driver pieces (SURVEY 8f rows N2 + N4):
  * the oracle's restatement of PhotoMontage.cpp:399-425 / :599-610 / :617-626 against an independent numpy
    formulation (the reference driver itself needs OpenCV + Eigen and cannot run here: parity unpinned);
  * the per-pixel bodies the CUDA kernels are made of (csrc/gsb_gdf_body.h), compiled for the host, against the
    oracle -- bit for bit, including the edge shapes and the out-of-range label check;
  * the semantics of the whole stage in oracle terms: with one source image the fused result is that image."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(n, H, W, seed):
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)
    labels = rng.integers(0, n, size=(H, W), dtype=np.uint8)
    return images, labels


def _np_gradients(images, labels):
    n, H, W, _ = images.shape
    im = images.astype(np.int32)
    yy, xx = np.mgrid[0:H, 0:W]
    gx, gy = np.zeros((3, H, W), np.float32), np.zeros((3, H, W), np.float32)
    if H > 1 and W > 1:
        l = labels[:-1, :-1]
        y, x = yy[:-1, :-1], xx[:-1, :-1]
        here, right, below = im[l, y, x], im[l, y, x + 1], im[l, y + 1, x]  # (H-1, W-1, 3)
        gx[:, :-1, :-1] = np.moveaxis(right - here, 2, 0)
        gy[:, :-1, :-1] = np.moveaxis(below - here, 2, 0)
    return gx, gy


SHAPES = [(3, 7, 9), (1, 1, 1), (2, 1, 6), (2, 5, 1), (4, 2, 2), (5, 33, 17)]


@pytest.mark.parametrize("n,H,W", SHAPES)
def test_oracle_gdf_matches_numpy(oracle_mod, n, H, W):
    images, labels = _case(n, H, W, seed=n * 100 + H)
    gx, gy = oracle_mod.gdf_gradients(images, labels)
    ex, ey = _np_gradients(images, labels)
    assert np.array_equal(gx, ex) and np.array_equal(gy, ey)
    x0 = oracle_mod.gdf_composite(images, labels)
    yy, xx = np.mgrid[0:H, 0:W]
    comp = images[labels, yy, xx].astype(np.float64)  # (H, W, 3)
    assert np.array_equal(x0, np.moveaxis(comp, 2, 0).reshape(3, H * W))
    x = np.random.default_rng(1).uniform(-40, 300, size=(3, H * W))
    x[0, 0] = 255.999
    out = oracle_mod.gdf_writeback(x, H, W)
    assert np.array_equal(out, np.moveaxis(np.clip(x, 0, 255).astype(np.uint8).reshape(3, H, W), 0, 2))
    # one channel of the interleaved write-back == the reference's per-channel write-back (A10)
    assert np.array_equal(out[..., 1].reshape(-1), oracle_mod.writeback_u8(x[1]))


@pytest.fixture(scope="module")
def host_bodies(tmp_path_factory):
    """gsb_gdf_body.h compiled with g++ (GSB_HD expands to `inline`)."""
    so = str(tmp_path_factory.mktemp("gdf") / "libgdf_body_host.so")
    src = os.path.join(ROOT, "tests", "cpp", "gdf_body_host.cc")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", so, src], check=True)
    L = C.CDLL(so)
    u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    L.host_gdf_gradients.argtypes = [u8, C.c_int, u8, C.c_int, C.c_int, f32, f32]
    L.host_gdf_composite.argtypes = [u8, C.c_int, u8, C.c_int, C.c_int, f64]
    L.host_gdf_writeback.argtypes = [f64, C.c_int64, u8]
    L.host_gdf_writeback.restype = None
    return L


@pytest.mark.parametrize("n,H,W", SHAPES)
def test_kernel_bodies_match_oracle(oracle_mod, host_bodies, n, H, W):
    images, labels = _case(n, H, W, seed=7 * n + W)
    gx, gy = np.full((3, H, W), np.nan, np.float32), np.full((3, H, W), np.nan, np.float32)
    assert host_bodies.host_gdf_gradients(images.reshape(-1), n, labels.reshape(-1), W, H, gx.reshape(-1),
                                          gy.reshape(-1)) == 0
    ox, oy = oracle_mod.gdf_gradients(images, labels)
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)  # every element written, last row / column 0
    x0 = np.full((3, H * W), np.nan)
    assert host_bodies.host_gdf_composite(images.reshape(-1), n, labels.reshape(-1), W, H, x0.reshape(-1)) == 0
    assert np.array_equal(x0, oracle_mod.gdf_composite(images, labels))
    x = np.random.default_rng(3).uniform(-300, 600, size=(3, H * W))
    x[2, -1] = np.nan  # the clamp sends NaN to 0 in both
    out = np.zeros(H * W * 3, np.uint8)
    host_bodies.host_gdf_writeback(x.reshape(-1), H * W, out)
    assert np.array_equal(out.reshape(H, W, 3), oracle_mod.gdf_writeback(x, H, W))
    # a label beyond the image list is reported by both
    bad = labels.copy()
    bad[H - 1, W - 1] = n
    assert host_bodies.host_gdf_gradients(images.reshape(-1), n, bad.reshape(-1), W, H, gx.reshape(-1),
                                          gy.reshape(-1)) == 1
    assert host_bodies.host_gdf_composite(images.reshape(-1), n, bad.reshape(-1), W, H, x0.reshape(-1)) == 1
    with pytest.raises(ValueError):
        oracle_mod.gdf_gradients(images, bad)


def test_single_source_fusion_reproduces_the_image(oracle_mod):
    """BuildSolveGradientFusion with one source image: the gradients are consistent, so the solution of
    A^T A v = A^T b is the image itself (pixel (W-1, H-1) has an empty row and stays wherever it started).
    Oracle terms only -- this pins the stage's semantics the GPU test then has to reproduce."""
    W, H = 12, 9
    images, _ = _case(1, H, W, seed=5)
    labels = np.zeros((H, W), np.uint8)
    gx, gy = oracle_mod.gdf_gradients(images, labels)
    ro, ci, va = oracle_mod.poisson_csr(W, H)
    m = oracle_mod.Oracle().import_csr(va, ro[:-1], ci, W * H)
    sol = np.empty((3, H * W))
    for c in range(3):
        b = oracle_mod.poisson_rhs(W, H, gx[c], gy[c], float(images[0, 0, 0, c]))
        # zero start: with the composite (== the exact solution here) as initial guess the reference's CG
        # divides 0 by 0 in its first step (v2 :419-420)
        x, it = m.cg(b, 1e-10, 500)
        sol[c] = x
    out = oracle_mod.gdf_writeback(sol + 0.5, H, W)  # +0.5: the reference truncates; avoid 41.9999 -> 41
    keep = np.ones((H, W), bool)
    keep[H - 1, W - 1] = False  # empty row: the pixel keeps its start value
    assert np.array_equal(out[keep], images[0][keep])
    assert np.all(out[H - 1, W - 1] == 0)
