"""Single-process multi-device solve (gsb_dist_init_local / gsb_set_devices): N devices of one box behind one
blocking call, no NCCL / torch / IPC.  Needs >= 2 GPUs (gpurun --gpus N); every case compares with the single-device
solve BIT FOR BIT."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _need(gsb, n):
    if gsb._lib.device_count() < n:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (n, n))


def _poisson_b(gsb, W, H, ch):
    from coursecomputationalphotography_b200 import workloads as wl
    img = wl.synth_image(W, H, ch, seed=7)
    gx, gy = wl.seamless_gradients(img)
    return np.asarray(gsb.poisson_rhs(W, H, gx, gy, img[:, 0, 0].astype(np.float64))).reshape(ch, W * H)


@pytest.mark.parametrize("ndev", [2, 3, 4, 8])
def test_local_group_poisson_strips_equal_single_device(gsb, ndev):
    """The full-grid Poisson system generated strip by strip on the devices (config C4's path), odd sizes included."""
    _need(gsb, ndev)
    from coursecomputationalphotography_b200 import strips
    g = strips.LocalGroup(list(range(ndev)))
    for (W, H, ch, sweeps, ce) in ((64, 48, 3, 9, 1), (301, 203, 1, 5, 2), (1024, 1024, 3, 6, 1)):
        b = _poisson_b(gsb, W, H, ch)
        g.poisson(W, H)
        opts = gsb.SparseMatrix.options(check_every=ce)
        xs = [g.gauss_seidel(b if ch > 1 else b[0], 0.0, sweeps, opts).reshape(ch, -1) for _ in range(2)]  # epochs reused
        assert g.last_stats.sweeps == sweeps and g.last_stats.kernel_used >= 30
        sp = gsb.SparseMatrix(np.float64)
        sp.poisson(W, H)
        x1 = sp.gaussSeidel(b if ch > 1 else b[0], epsilon=0.0, max_iteration=sweeps, options=opts).reshape(ch, -1)
        assert np.array_equal(xs[0], x1) and np.array_equal(xs[1], x1), (W, H, ndev)
        r1 = sp.residual(b[0], x1[0])
        assert abs(g.residual(b[0], xs[0][0]) - r1) <= 1e-9 * max(1.0, r1)
    g.close()


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_host_api_shards_imported_csr(gsb, ndev):
    """SparseMatrix.gaussSeidel on an IMPORTED host CSR (initializeFromEigenRowMajor) with a device list set: the
    masked blend (compact unknowns, caller's colours) and the full grid (red-black probe); same bits, same stop
    sweep as one device; matrices that need more colours stay on one device."""
    _need(gsb, ndev)
    from coursecomputationalphotography_b200 import strips, workloads as wl
    W = H = 1024
    mask = wl.blob_mask(W, H, 0.35, 48, seed=11)
    guide = wl.synth_image(W, H, 3, seed=7)
    ro, ci, va, b, pix, colors = wl.masked_poisson_system(mask, guide, np.ascontiguousarray(guide[:, ::-1, ::-1]))
    n = len(pix)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    sp.analyze(gsb._lib.ORDER_USER, colors)
    try:
        strips.set_devices([])
        x1 = sp.gaussSeidel(b, epsilon=0.0, max_iteration=40)
        assert sp.last_stats.kernel_used < 10 and sp.last_stats.sweeps == 40
        # a real epsilon: a right-hand side whose solution is close to the start vector (the reference's loop only
        # runs while the update norm is below its initial eps = 10, v2 :354-356), stop threshold = 1.5 x the update
        # norm of sweep 40 -> the solve must stop at that sweep or a little earlier, on one device and on N alike
        bs = sp.applyToVector(np.ones(n))[None, :] + 1e-7 * b
        sp.gaussSeidel(bs, epsilon=0.0, max_iteration=40)
        eps = 1.5 * max(list(sp.last_stats.last_eps)[:3])
        assert 0.0 < eps < 10.0
        xs1 = sp.gaussSeidel(bs, epsilon=eps, max_iteration=500)
        s1 = sp.last_stats.sweeps
        assert 5 < s1 <= 40
        strips.set_devices(list(range(ndev)))
        assert strips.get_devices() == list(range(ndev))
        for rep in range(2):
            xn = sp.gaussSeidel(b, epsilon=0.0, max_iteration=40)
            assert sp.last_stats.kernel_used >= 30 and sp.last_stats.sweeps == 40
            assert np.array_equal(xn, x1), "masked system: %d devices differ from one (rep %d)" % (ndev, rep)
        xsn = sp.gaussSeidel(bs, epsilon=eps, max_iteration=500)
        assert sp.last_stats.sweeps == s1 and np.array_equal(xsn, xs1)
        # the full grid, imported as Eigen would hand it over
        Wg, Hg = 640, 512
        fg = gsb.SparseMatrix(np.float64)
        fg.poisson(Wg, Hg)
        v2, c2, _, rn, _ = fg.layout()
        ro2 = np.zeros(Wg * Hg, np.int32)
        ro2[1:] = np.cumsum(rn[:-1])
        bg = _poisson_b(gsb, Wg, Hg, 3)
        imp = gsb.SparseMatrix(np.float64)
        imp.initializeFromEigenRowMajor(v2, len(v2), ro2, Wg * Hg, c2, Wg * Hg)
        xn = imp.gaussSeidel(bg, epsilon=0.0, max_iteration=12)
        assert imp.last_stats.kernel_used >= 30
        strips.set_devices([])
        assert np.array_equal(xn, imp.gaussSeidel(bg, epsilon=0.0, max_iteration=12))
        # a matrix that needs more than two colours does not shard: it runs on one device
        strips.set_devices(list(range(ndev)))
        nm = 100_000 * ndev
        rr, cc, vv = wl.random_spd_coo(nm, 6, seed=3)  # ~13 entries per row, random coupling: needs > 2 colours
        mc = gsb.SparseMatrix(np.float64)
        mc.initialize(nm, nm)
        mc.initializeFromTriplets(rr, cc, vv)
        b2 = np.random.default_rng(2).standard_normal(nm)
        xm = mc.gaussSeidel(b2, epsilon=0.0, max_iteration=5)
        assert mc.last_stats.kernel_used < 10 and mc.last_stats.n_colors > 2
        strips.set_devices([])
        assert np.array_equal(xm, mc.gaussSeidel(b2, epsilon=0.0, max_iteration=5))
    finally:
        strips.set_devices([])


def test_local_group_rejects_bad_input(gsb):
    from coursecomputationalphotography_b200 import strips
    with pytest.raises(gsb.GsbError):
        strips.LocalGroup([0, 0])
    with pytest.raises(gsb.GsbError):
        strips.LocalGroup([gsb._lib.device_count()])
    g = strips.LocalGroup([0])  # a group of one is the single-rank strip solver
    g.poisson(96, 64)
    b = _poisson_b(gsb, 96, 64, 1)[0]
    x = g.gauss_seidel(b, 0.0, 7)
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(96, 64)
    assert np.array_equal(x, sp.gaussSeidel(b, epsilon=0.0, max_iteration=7))
    with pytest.raises(ValueError):
        g.gauss_seidel(b[:-1])
    g.close()
