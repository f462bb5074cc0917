"""Row N3 pinned to the reference: the lab8 right-hand-side producers (hw8_pa.cc:316-498, struct Gradients :602-676)
compiled UNMODIFIED from /root/reference into oracle/_ref/libpanoref.so (oracle/ref_pano_shim.cc: a stand-in for the
few cv:: names they use, images with zero guard rows) against (a) the oracle's literal restatement orc_pano_* and
(b) the kernels' bounded per-pixel bodies compiled for the host -- bit for bit, on masks built to hit every branch.
The GPU tests (tests/test_zz_pano_gpu.py) hold the device kernels to the same restatement."""
import numpy as np
import pytest

from test_pano_host import SHAPES, host, row_masks  # noqa: F401  (host is a fixture)


@pytest.fixture(scope="module")
def ref(oracle_mod):
    if not oracle_mod.pano_ref_available():
        pytest.skip("oracle/_ref/libpanoref.so absent (built only where /root/reference exists)")
    return oracle_mod


@pytest.mark.parametrize("H,W", SHAPES + [(64, 200)])
@pytest.mark.parametrize("style", ["run", "mixed", "noise", "empty", "full"])
def test_restatement_equals_compiled_reference(ref, host, H, W, style):
    o = ref
    rng = np.random.default_rng(H * 977 + W * 13 + len(style))
    for rep in range(5):
        tmask = row_masks(rng, H, W, "mixed" if rep % 2 else style)
        outer = row_masks(rng, H, W, style)
        inner = (outer & row_masks(rng, H, W, "holes")) if rep % 3 else row_masks(rng, H, W, "mixed")
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        tgt = rng.standard_normal((H, W, 3)).astype(np.float32)
        src = rng.standard_normal((H, W, 3)).astype(np.float32)
        assert np.array_equal(o.pano_mask_image(img, outer), o.ref_pano_mask_image(img, outer))
        for a, b in zip(o.pano_gradients(img), o.ref_pano_gradients(img)):
            assert np.array_equal(a, b)
        assert np.array_equal(o.pano_merge2_f32(tgt, src, tmask, outer, inner),
                              o.ref_pano_merge2_f32(tgt, src, tmask, outer, inner))
        for skip in (0.0, 1.0, 10.0, 2.5):
            ti = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            si = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            assert np.array_equal(o.pano_merge_u8(ti, si, tmask, outer, skip), o.ref_pano_merge_u8(ti, si, tmask, outer, skip))
        m1, s1 = rng.integers(0, 2, (H, W), dtype=np.uint8) * 255, rng.integers(0, 2, (H, W), dtype=np.uint8) * 255
        assert np.array_equal(o.pano_merge_u8(m1, s1, tmask, outer, 0.0), o.ref_pano_merge_u8(m1, s1, tmask, outer, 0.0))
        dx = rng.standard_normal((H, W, 3)).astype(np.float32)
        dy = rng.standard_normal((H, W, 3)).astype(np.float32)
        want = o.ref_pano_enforce_gradient_bound(dx, dy, img, outer)
        for a, b in zip(o.pano_enforce_gradient_bound(dx, dy, img, outer), want):
            assert np.array_equal(a, b)
        gx, gy = dx.copy(), dy.copy()  # the kernels' body
        host.host_pano_enforce_gradient_bound(gx.reshape(-1), gy.reshape(-1), img.reshape(-1), outer.reshape(-1), W, H)
        assert np.array_equal(gx, want[0]) and np.array_equal(gy, want[1])
        # struct Gradients, second (mask-driven) constructor
        want = o.ref_pano_gradients(img, outer)
        for a, b in zip(o.pano_gradients_masked(img, outer), want):
            assert np.array_equal(a, b)
        gx, gy = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.float32)
        host.host_pano_gradients_masked(img.reshape(-1), outer.reshape(-1), W, H, gx.reshape(-1), gy.reshape(-1))
        assert np.array_equal(gx, want[0]) and np.array_equal(gy, want[1])
