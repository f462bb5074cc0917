"""GPU parity of the gradient-domain-fusion driver (SURVEY 8f rows N2 + N4; gsb_gdf_* in include/gsb200.h):
bit-exact against the oracle for the image-space pieces, bit-exact against the composition of the already
verified solver entry points for the whole stage, and the reference-semantics check (one source image in ->
that image out).  Named test_zz_* so that it runs after the hot-path suites."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(n, H, W, seed):
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)
    labels = rng.integers(0, n, size=(H, W), dtype=np.uint8)
    return images, labels


def _smooth_case(n, H, W, seed):
    """n smooth exposures of one scene, label map = vertical bands (what a graph cut typically returns)."""
    from coursecomputationalphotography_b200 import workloads as wl
    base = wl.synth_image(W, H, 3, seed=seed).astype(np.float64)  # (3, H, W)
    images = np.stack([np.clip(np.moveaxis(base, 0, 2) * (0.7 + 0.2 * k) + 10 * k, 0, 255).astype(np.uint8)
                       for k in range(n)])
    labels = ((np.arange(W)[None, :] * n) // W).astype(np.uint8).repeat(H, axis=0)
    return np.ascontiguousarray(images), np.ascontiguousarray(labels)


@pytest.mark.parametrize("n,H,W", [(3, 7, 9), (1, 1, 1), (2, 1, 6), (2, 5, 1), (4, 2, 2), (5, 33, 17), (3, 301, 203)])
def test_gradients_and_composite_bitexact(gsb, oracle_mod, n, H, W):
    from coursecomputationalphotography_b200 import gdf
    images, labels = _case(n, H, W, seed=11 * n + H)
    gx, gy = gdf.gdf_gradients(images, labels)
    ox, oy = oracle_mod.gdf_gradients(images, labels)
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)
    assert np.array_equal(gdf.gdf_composite(images, labels), oracle_mod.gdf_composite(images, labels))


def test_label_out_of_range_is_an_error(gsb):
    from coursecomputationalphotography_b200 import gdf
    images, labels = _case(2, 6, 5, seed=1)
    labels[3, 2] = 2
    for fn in (gdf.gdf_gradients, gdf.gdf_composite, gdf.BuildSolveGradientFusion):
        with pytest.raises(gsb.GsbError) as e:
            fn(images, labels)
        assert e.value.status == 1
    with pytest.raises(ValueError):
        gdf.gdf_gradients(images[:, :, :, :2], labels)


@pytest.mark.parametrize("W,H,sweeps", [(64, 48, 9), (301, 203, 5)])
def test_solve_channels_equals_composition_of_the_solver_api(gsb, oracle_mod, W, H, sweeps):
    """gsb_gdf_solve == poisson_rhs -> poisson matrix -> gaussSeidel (3 right-hand sides fused, x0) -> clamp,
    bit for bit; and the image-space ends of it equal the oracle."""
    from coursecomputationalphotography_b200 import gdf
    images, labels = _smooth_case(3, H, W, seed=3)
    gx, gy = oracle_mod.gdf_gradients(images, labels)
    init = oracle_mod.gdf_composite(images, labels)
    constraint = images[0, 0, 0].astype(np.float64)
    opts = gdf.gdf_options(epsilon=0.0, max_iteration=sweeps)
    for ini in (init, None):
        out, st = gdf.SolveChannels(gx, gy, constraint, ini, opts)
        b = gsb.poisson_rhs(W, H, gx, gy, constraint)
        sp = gsb.SparseMatrix(np.float64)
        sp.poisson(W, H)
        x = sp.gaussSeidel(b, epsilon=0.0, max_iteration=sweeps, x0=ini)
        assert np.array_equal(out, oracle_mod.gdf_writeback(x, H, W))
        assert list(st.iterations)[:3] == [sweeps] * 3
        assert list(st.last_eps)[:3] == list(sp.last_stats.last_eps)[:3]
        for c in range(3):
            r = sp.residual(b[c], x[c])
            assert abs(st.residual_l2[c] - r) <= 1e-9 * max(1.0, r)
    # the whole stage from images + labels
    out2, st2 = gdf.BuildSolveGradientFusion(images, labels, fast_init=True, options=opts)
    out1, _ = gdf.SolveChannels(gx, gy, constraint, init, opts)
    assert np.array_equal(out2, out1) and st2.total_ms > 0.0


def test_single_source_fusion_reproduces_the_image(gsb):
    """Reference semantics (see tests/test_gdf_host.py): one source image in -> that image out, with the
    reference's own solver call (conjugateGradient, eps 1e-10) from a zero start and with Gauss-Seidel from the
    composite (already the solution: the sweeps must leave it in place)."""
    from coursecomputationalphotography_b200 import gdf
    W, H = 12, 9
    images, _ = _case(1, H, W, seed=5)
    labels = np.zeros((H, W), np.uint8)
    keep = np.ones((H, W), bool)
    keep[H - 1, W - 1] = False  # empty row of A^T A: the pixel keeps its start value
    gx, gy = gdf.gdf_gradients(images, labels)
    # SolveChannel writes uchar(solution) (truncation): shift the constraint by 0.5 so that 41.9999.. stays 41+
    c = images[0, 0, 0].astype(np.float64) + 0.5
    out, st = gdf.SolveChannels(gx, gy, c, None, gdf.gdf_options(solver=gdf.GDF_CG, epsilon=1e-10, max_iteration=500))
    assert np.array_equal(out[keep], images[0][keep]) and np.all(out[H - 1, W - 1] == 0)
    assert max(list(st.residual_l2)[:3]) < 1e-6 and max(list(st.iterations)[:3]) < 500
    out, st = gdf.BuildSolveGradientFusion(images, labels, fast_init=True,
                                           options=gdf.gdf_options(epsilon=1e-9, max_iteration=50))
    assert np.array_equal(out, images[0])
    assert st.iterations[0] == 1 and max(list(st.last_eps)[:3]) <= 1e-9  # nothing moves


def test_matrix_is_rebuilt_when_the_size_changes(gsb, oracle_mod):
    from coursecomputationalphotography_b200 import gdf
    opts = gdf.gdf_options(epsilon=0.0, max_iteration=3)
    outs = {}
    for (W, H) in ((40, 30), (30, 40), (40, 30)):
        images, labels = _smooth_case(2, H, W, seed=9)
        outs.setdefault((W, H), []).append(gdf.BuildSolveGradientFusion(images, labels, True, opts)[0])
    assert np.array_equal(outs[(40, 30)][0], outs[(40, 30)][1])
    assert gsb.load().gsb_gdf_release() == 0
    images, labels = _smooth_case(2, 30, 40, seed=9)
    assert np.array_equal(gdf.BuildSolveGradientFusion(images, labels, True, opts)[0], outs[(40, 30)][0])
