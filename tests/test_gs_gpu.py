"""GPU parity tests of the Gauss-Seidel path (A6-A8) through the C ABI, against the CPU oracle.

Two kinds of statement (BASELINE.json north_star):
  1. per-sweep bit-exactness: multicolour GS on the device == the oracle's lexicographic GS on the
     symmetrically permuted matrix P A P^T, for the same number of sweeps;
  2. converged parity: the fixed point matches the oracle's natural-order solution within
     max-abs <= 1e-4 * 255 (image range), with residual norms reported.
"""
import numpy as np
import pytest

from helpers import oracle_from_csr, permuted_csr

pytestmark = pytest.mark.gpu

TOL_IMAGE = 1e-4 * 255.0  # north_star: max-abs <= 1e-4 relative to the image range


def _poisson_rhs_for(pkg, wl, W, H, C=1, seed=7):
    img = wl.synth_image(W, H, C, seed=seed)
    gx, gy = wl.seamless_gradients(img)
    b = pkg.poisson_rhs(W, H, gx if C > 1 else gx[0], gy if C > 1 else gy[0], img[:, 0, 0].astype(np.float64))
    return img, b


def test_known_answer_4x4(gsb, oracle_mod):
    """labs/lab3/src/OpenCVHW1/main6.cc:238-249: expected print `1 2 -1 1`."""
    A = [10, -1, 2, 0, -1, 11, -1, 3, 2, -1, 10, -1, 0, 3, -1, 8]
    b = np.array([6, 25, -11, 15], np.float64)
    sp = gsb.SparseMatrix(np.int32)
    sp.initialize(4, 4, A)
    x = sp.gaussSeidel(b)
    assert np.allclose(x, [1, 2, -1, 1], atol=1e-6)
    o = oracle_mod.Oracle().init_dense(4, 4, A)
    xo, sw, _ = o.gauss_seidel(b)
    assert np.abs(x - xo).max() < 1e-6
    assert sp.last_stats.sweeps > 0 and sp.last_stats.last_eps[0] <= 1e-6
    # rows 0 and 3 are not coupled (a03 = a30 = 0): 3 colours suffice, 4 is the lexicographic order itself
    info = sp.coloring()
    assert info["n_colors"] in (3, 4)
    perm, _ = sp.ordering()
    if np.array_equal(perm, np.arange(4)):
        assert np.array_equal(x, xo) and sp.last_stats.sweeps == sw


@pytest.mark.parametrize("W,H", [(16, 12), (33, 17), (64, 64), (2, 2), (5, 1), (1, 7)])
@pytest.mark.parametrize("sweeps", [1, 7])
def test_redblack_poisson_bitexact_vs_oracle_on_permuted(gsb, oracle_mod, W, H, sweeps):
    from coursecomputationalphotography_b200 import workloads as wl
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    img, b = _poisson_rhs_for(gsb, wl, W, H)
    x = sp.gaussSeidel(b, epsilon=0.0, max_iteration=sweeps)
    info = sp.coloring()
    assert info["ordering"] == gsb._lib.ORDER_REDBLACK and info["n_colors"] <= 2
    perm, colors = sp.ordering()
    ro, ci, va = oracle_mod.poisson_csr(W, H)
    pro, pci, pva = permuted_csr(ro, ci, va, perm)
    o = oracle_from_csr(oracle_mod, pro, pci, pva)
    xp, sw, _ = o.gauss_seidel(b[perm], 0.0, sweeps)
    xo = np.empty_like(xp)
    xo[perm] = xp
    assert np.array_equal(x, xo), "device sweep differs from the oracle on P A P^T (max %g)" % np.abs(x - xo).max()
    # the empty last row is skipped: x stays at 1.0 (v2 :360-363)
    if W > 1 and H > 1:
        assert x[-1] == 1.0


def test_multicolor_random_bitexact_and_converged(gsb, oracle_mod):
    from coursecomputationalphotography_b200 import workloads as wl
    n = 5000
    r, c, v, b, xstar = wl.diag_dominant_system(n, 4, seed=3)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromVector(r, c, v)
    x5 = sp.gaussSeidel(b, epsilon=0.0, max_iteration=5)
    info = sp.coloring()
    assert info["ordering"] == gsb._lib.ORDER_MULTICOLOR and 2 <= info["n_colors"] <= 64
    perm, colors = sp.ordering()
    ro, ci, va = wl.coo_to_csr(r, c, v, n)
    # proper colouring: no row reads an unknown of its own colour
    rows = np.repeat(np.arange(n), np.diff(ro))
    off = rows != ci
    assert not np.any(colors[rows[off]] == colors[ci[off]])
    pro, pci, pva = permuted_csr(ro, ci, va, perm)
    o = oracle_from_csr(oracle_mod, pro, pci, pva)
    xp, _, _ = o.gauss_seidel(b[perm], 0.0, 5)
    xo = np.empty_like(xp)
    xo[perm] = xp
    assert np.array_equal(x5, xo)
    # converged parity against the natural-order oracle (reference defaults)
    x = sp.gaussSeidel(b)
    on = oracle_from_csr(oracle_mod, ro, ci, va)
    xn, sw, _ = on.gauss_seidel(b)
    assert np.abs(x - xn).max() <= 1e-6
    assert np.abs(x - xstar).max() <= 1e-6
    assert abs(sp.last_stats.sweeps - sw) <= 6  # ordering changes the count slightly, not the fixed point
    assert sp.residual(b, x) <= 1e-5


def test_masked_blend_converged_parity_user_colors(gsb, oracle_mod):
    """Dirichlet-masked 5-point blend (SURVEY 8d C3 at 192^2): converged solution vs the oracle."""
    from coursecomputationalphotography_b200 import workloads as wl
    W = H = 192
    mask = wl.blob_mask(W, H, 0.30, 24, seed=11)
    guide = wl.synth_image(W, H, 3, seed=7)
    target = wl.synth_image(W, H, 3, seed=9)
    ro, ci, va, b, pix, colors = wl.masked_poisson_system(mask, guide, target)
    n = len(pix)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    info = sp.analyze(gsb._lib.ORDER_USER, colors)
    assert info["n_colors"] == 2
    x = sp.gaussSeidel(b, epsilon=1e-7, max_iteration=20000)
    assert sp.last_stats.sweeps < 20000
    o = oracle_from_csr(oracle_mod, ro, ci, va)
    for ch in range(3):
        xo, sw, _ = o.gauss_seidel(b[ch], 1e-7, 20000)
        err = np.abs(x[ch] - xo).max()
        r_gpu, r_cpu = sp.residual(b[ch], x[ch]), np.linalg.norm(b[ch] - o.spmv(xo))
        print("ch%d: max-abs %.3e  ||r|| gpu %.3e cpu %.3e  sweeps gpu %d cpu %d" %
              (ch, err, r_gpu, r_cpu, sp.last_stats.sweeps, sw))
        assert err <= TOL_IMAGE
        assert np.array_equal(gsb.writeback_u8(x[ch]), oracle_mod.writeback_u8(xo)) or err < TOL_IMAGE
    # automatic ordering on the same (irregular) system also converges to the same point
    sp2 = gsb.SparseMatrix(np.float64)
    sp2.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    x2 = sp2.gaussSeidel(b[0], epsilon=1e-7, max_iteration=20000)
    assert np.abs(x2 - x[0]).max() <= TOL_IMAGE


def test_multi_rhs_equals_single_rhs(gsb):
    from coursecomputationalphotography_b200 import workloads as wl
    W, H = 48, 40
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    img, b = _poisson_rhs_for(gsb, wl, W, H, C=3)
    x3 = sp.gaussSeidel(b, epsilon=0.0, max_iteration=9)
    for ch in range(3):
        x1 = sp.gaussSeidel(b[ch], epsilon=0.0, max_iteration=9)
        assert np.array_equal(x1, x3[ch])
    b4 = np.concatenate([b, b[:1]])
    x4 = sp.gaussSeidel(b4, epsilon=0.0, max_iteration=9)
    assert np.array_equal(x4[:3], x3) and np.array_equal(x4[3], x3[0])


def test_stop_rule_semantics(gsb):
    from coursecomputationalphotography_b200 import workloads as wl
    r, c, v, b, _ = wl.diag_dominant_system(2000, 4, seed=5)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromVector(r, c, v)
    # eps starts at 10 (v2 :354): epsilon >= 10 or max_iteration <= 0 means no sweep at all, x stays 1.0
    x = sp.gaussSeidel(b, epsilon=10.0)
    assert sp.last_stats.sweeps == 0 and np.all(x == 1.0)
    x = sp.gaussSeidel(b, max_iteration=0)
    assert sp.last_stats.sweeps == 0 and np.all(x == 1.0)
    # max_iteration caps the count exactly, whatever the batching
    for k in (1, 3, 5):
        for opts in (gsb.SparseMatrix.options(use_graph=0, batch_sweeps=2),
                     gsb.SparseMatrix.options(use_graph=1, batch_sweeps=4)):
            sp.gaussSeidel(b, epsilon=0.0, max_iteration=k, options=opts)
            assert sp.last_stats.sweeps == k
    # the accepted iterate is the one the rule fired on: graph / plain / different batch sizes agree bitwise
    xa = sp.gaussSeidel(b, options=gsb.SparseMatrix.options(use_graph=0, batch_sweeps=1))
    sa = sp.last_stats.sweeps
    xb = sp.gaussSeidel(b, options=gsb.SparseMatrix.options(use_graph=1, batch_sweeps=16))
    sb = sp.last_stats.sweeps
    xc = sp.gaussSeidel(b, options=gsb.SparseMatrix.options(use_graph=0, batch_sweeps=7))
    assert sa == sb == sp.last_stats.sweeps and np.array_equal(xa, xb) and np.array_equal(xa, xc)
    assert sp.last_stats.last_eps[0] <= 1e-6
    # check_every = 4: stops at a multiple of 4 that is >= the every-sweep count
    sp.gaussSeidel(b, options=gsb.SparseMatrix.options(check_every=4))
    assert sp.last_stats.sweeps % 4 == 0 and sa <= sp.last_stats.sweeps < sa + 4


def test_zero_diagonal_rows_are_skipped(gsb, oracle_mod):
    # 3x3: row 1 has no diagonal entry -> x[1] stays 1.0 (v2 :360-363)
    rows = np.array([0, 0, 1, 2, 2], np.int32)
    cols = np.array([0, 1, 0, 1, 2], np.int32)
    vals = np.array([4.0, 1.0, 2.0, 1.0, 5.0])
    b = np.array([1.0, 2.0, 3.0])
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromVector(rows, cols, vals)
    x = sp.gaussSeidel(b)
    o = oracle_mod.Oracle().init_from_vector(rows, cols, vals)
    xo, _, _ = o.gauss_seidel(b)
    assert x[1] == 1.0 and np.allclose(x, xo, atol=1e-12)


def test_x0_extension_and_errors(gsb):
    from coursecomputationalphotography_b200 import workloads as wl
    r, c, v, b, xstar = wl.diag_dominant_system(500, 4, seed=8)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromVector(r, c, v)
    x = sp.gaussSeidel(b, x0=xstar)
    assert sp.last_stats.sweeps <= 2 and np.abs(x - xstar).max() < 1e-9
    with pytest.raises(ValueError):
        sp.gaussSeidel(b[:-1])
    rect = gsb.SparseMatrix(np.float64)
    rect.initializeFromVector([0, 1], [0, 3], [1.0, 2.0])  # 2 x 4
    with pytest.raises(gsb.GsbError) as e:
        rect.gaussSeidel(np.ones(4))
    assert e.value.status == 2  # GSB_ERR_SHAPE
    bad = gsb.SparseMatrix(np.float64)
    bad.initialize(3, 3, [2, 1, 0, 1, 2, 1, 0, 1, 2])
    with pytest.raises(gsb.GsbError) as e:
        bad.analyze(gsb._lib.ORDER_USER, [0, 0, 1])  # rows 0 and 1 are coupled
    assert e.value.status == 8  # GSB_ERR_COLORING


def test_spmv_bitexact_and_vector_helpers(gsb, oracle_mod):
    from coursecomputationalphotography_b200 import workloads as wl
    rng = np.random.default_rng(0)
    r, c, v, b, _ = wl.diag_dominant_system(3000, 6, seed=2)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromVector(r, c, v)
    o = oracle_mod.Oracle().init_from_vector(r, c, v)
    vin = rng.standard_normal(3000)
    assert np.array_equal(sp.applyToVector(vin), o.spmv(vin))
    a, bb = rng.standard_normal(100001), rng.standard_normal(100001)
    assert gsb.manhattonDist([1, 2, 3, 10], [2, 1, 3, 8]) == 4.0  # main6.cc:233-235
    assert abs(gsb.manhattonDist(a, bb) - oracle_mod.l1_dist(a, bb)) <= 1e-9 * np.abs(a - bb).sum()
    assert abs(gsb.dotProd(a, bb) - float(np.dot(a, bb))) <= 1e-9 * np.abs(a * bb).sum()
    assert np.array_equal(gsb.vecadd(a, bb, 0.37), a + 0.37 * bb)
    assert np.array_equal(gsb.vecsub(a, bb), a - bb)
    assert np.array_equal(gsb.vecmul(a, bb), a * bb)


def test_cg_and_pcg_vs_oracle(gsb, oracle_mod):
    from coursecomputationalphotography_b200 import workloads as wl
    A = [10, -1, 2, 0, -1, 11, -1, 3, 2, -1, 10, -1, 0, 3, -1, 8]
    b = np.array([6, 25, -11, 15], np.float64)
    sp = gsb.SparseMatrix(np.float64)
    sp.initialize(4, 4, A)
    assert np.allclose(sp.conjugateGradient(b), [1, 2, -1, 1], atol=1e-12)  # main6.cc:251-253
    assert np.allclose(sp.conjugateGradientEigen(b), [1, 2, -1, 1], atol=1e-12)
    # Poisson 32x24 with the composite as the initial guess (PhotoMontage.cpp:599-613)
    W, H = 32, 24
    img, bp = _poisson_rhs_for(gsb, wl, W, H)
    sp2 = gsb.SparseMatrix(np.float64)
    sp2.poisson(W, H)
    ro, ci, va = oracle_mod.poisson_csr(W, H)
    o = oracle_from_csr(oracle_mod, ro, ci, va)
    init = img[0].astype(np.float64).ravel()
    x = sp2.conjugateGradient(bp, 1e-10, 50, init)
    xo, it = o.cg(bp, 1e-10, 50, init)
    assert np.abs(x - xo).max() < 1e-6 * 255
    xj = sp2.conjugateGradientEigen(bp, 1e-10, 400)
    xjo, _ = o.pcg(bp, 1e-10, 400)
    assert np.abs(xj[:-1] - xjo[:-1]).max() < 1e-5 * 255


def test_multicolour_on_unsymmetric_pattern(gsb, oracle_mod):
    """Jones-Plassmann on the SYMMETRISED pattern: a structurally unsymmetric matrix (row i stores column j but row j
    does not store column i) of 400 000 rows -- the case on which round 1's speculate-and-repair colouring ran into
    its round cap -- gets a proper colouring with few colours in a fraction of a second, and every sweep is bit-exact
    against the oracle on P A P^T."""
    import time
    from coursecomputationalphotography_b200 import workloads as wl
    for n, seed in ((400_000, 3), (7001, 13)):
        r, c, v, b, _ = wl.diag_dominant_system(n, 4, seed=seed)
        sp = gsb.SparseMatrix(np.float64)
        sp.initializeFromVector(r, c, v)
        t0 = time.perf_counter()
        info = sp.analyze()
        dt = time.perf_counter() - t0
        assert 3 <= info["n_colors"] <= 16 and dt < 2.0, (info, dt)
        perm, colors = sp.ordering()
        assert not np.any(colors[r[r != c]] == colors[c[r != c]])  # no row reads an unknown of its own colour
        x = sp.gaussSeidel(b, epsilon=0.0, max_iteration=4)
        ro, ci, va = wl.coo_to_csr(r, c, v, n)
        o = oracle_from_csr(oracle_mod, *permuted_csr(ro, ci, va, perm))
        xp, _, _ = o.gauss_seidel(b[perm], 0.0, 4)
        xo = np.empty_like(xp)
        xo[perm] = xp
        assert np.array_equal(x, xo), n


def test_cg_device_loop_multi_rhs_and_stop_semantics(gsb, oracle_mod):
    """conjugateGradient as a device loop: (a) three right-hand sides in one call == three single calls, bit for bit,
    each with its own loop count; (b) the reference's stop semantics (v2 :417-431): max_iteration caps the count, a
    break on sqrt(r.r) < eps leaves the count un-bumped, max_iteration = 0 returns the initial guess; (c) vs the
    oracle (itself bit-pinned to the reference's CG) with an initial guess, to rounding."""
    from coursecomputationalphotography_b200 import workloads as wl
    W, H = 96, 80
    img, b = _poisson_rhs_for(gsb, wl, W, H, C=3)
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    init = img.reshape(3, -1).astype(np.float64)
    xm = sp.conjugateGradientMulti(b, 1e-8, 3000, init)  # (the oracle needs ~600 iterations on this Neumann + pin system)
    its = list(sp.last_iters)
    assert len(its) == 3 and all(100 < k < 3000 for k in its)
    for c in range(3):
        x1 = sp.conjugateGradient(b[c], 1e-8, 3000, init[c])
        assert sp.last_iters == its[c] and np.array_equal(x1, xm[c]), c
    ro, ci, va = oracle_mod.poisson_csr(W, H)
    o = oracle_from_csr(oracle_mod, ro, ci, va)
    for c in range(3):
        xo, it = o.cg(b[c], 1e-8, 3000, init[c])
        assert abs(it - its[c]) <= 0.2 * it and np.abs(xo - xm[c]).max() < 1e-4 * 255
    # caps and the zero-iteration case
    x5 = sp.conjugateGradient(b[0], 1e-30, 5, init[0])
    assert sp.last_iters == 5
    xo5, it5 = o.cg(b[0], 1e-30, 5, init[0])
    assert it5 == 5 and np.abs(x5 - xo5).max() < 1e-9 * 255
    x0 = sp.conjugateGradient(b[0], 1e-8, 0, init[0])
    assert sp.last_iters == 0 and np.array_equal(x0, init[0])
    # mixed stopping: one channel starts at its solution (stops at once, count 0), the others go on
    start = init.copy()
    start[1] = xm[1]
    xs = sp.conjugateGradientMulti(b, 1e-6, 3000, start)
    assert sp.last_iters[1] <= 1 and sp.last_iters[0] > 3 and np.abs(xs - xm).max() < 1e-3
    # Jacobi variant unchanged in meaning
    xj = sp.conjugateGradientEigen(b[0], 1e-10, 600)
    xjo, _ = o.pcg(b[0], 1e-10, 600)
    assert np.abs(xj[:-1] - xjo[:-1]).max() < 1e-5 * 255


def test_full_size_properties_4096(gsb):
    """BASELINE configs[2] size: 4096^2 x 3 channels.  Size-independent properties instead of an oracle run."""
    import ctypes as C
    from coursecomputationalphotography_b200 import workloads as wl
    W = H = 4096
    n = W * H
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    assert sp._nnz == wl.poisson_nnz(W, H) == 83_853_315
    # (1) a constant image is a fixed point: b = A*const (only the pin row is nonzero) -> x stays const
    rng = np.random.default_rng(1)
    b = np.zeros((3, n))
    b[:, 0] = 1.0  # pin v(0,0) = 1, zero gradients: the solution is the all-ones start vector
    x = sp.gaussSeidel(b, epsilon=0.0, max_iteration=3)
    # the update is exactly zero, so eps = 0 and `eps > epsilon` (v2 :356) already fails after one sweep
    assert sp.last_stats.sweeps == 1 and np.all(x == 1.0)
    assert sp.last_stats.last_eps[0] == 0.0
    # (2) linearity of one sweep chain is not exact in floating point, but the residual must fall
    #     monotonically for an SPD system under GS: compare ||b-Ax|| after 2 and after 12 sweeps
    img = wl.synth_image(W, H, 1, seed=3)
    gx, gy = wl.seamless_gradients(img)
    b1 = gsb.poisson_rhs(W, H, gx[0], gy[0], float(img[0, 0, 0]))
    xa = sp.gaussSeidel(b1, epsilon=0.0, max_iteration=2)
    ra = sp.residual(b1, xa)
    xb = sp.gaussSeidel(b1, epsilon=0.0, max_iteration=12)
    rb = sp.residual(b1, xb)
    assert rb < ra
    # (3) red-black: two colours, both half the grid; empty corner row untouched
    info = sp.coloring()
    assert info["n_colors"] == 2 and info["grid_width"] == W
    assert xb[-1] == 1.0
    # (4) the A-energy error decreases: applyToVector parity with the residual kernel
    Ax = sp.applyToVector(xb)
    assert abs(np.linalg.norm(b1 - Ax) - rb) <= 1e-9 * max(rb, 1.0)


@pytest.mark.parametrize("nrhs", [1, 3])
def test_all_kernels_agree_bitwise(gsb, nrhs):
    """direct (1), staged (2), ring (3), window-staged ring (4) and the fused two-colour sweep (5) implement the
    same arithmetic in the same order."""
    from coursecomputationalphotography_b200 import workloads as wl
    W, H = 300, 217  # odd sizes: tiles and bulk-copy spans start at unaligned offsets
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    img, b = _poisson_rhs_for(gsb, wl, W, H, C=3)
    bb = b[:nrhs] if nrhs > 1 else b[0]
    res = {}
    for k in (1, 2, 3, 4, 5):
        for ce in (1, 3):
            x = sp.gaussSeidel(bb, epsilon=0.0, max_iteration=12, options=gsb.SparseMatrix.options(kernel=k, check_every=ce))
            assert sp.last_stats.kernel_used == k and sp.last_stats.sweeps == 12
            res[(k, ce)] = (x, sp.last_stats.last_eps[0])
    x0, e0 = res[(1, 1)]
    for key, (x, e) in res.items():
        assert np.array_equal(x, x0), key
        # the stop norm is summed in a fixed but kernel-specific order: equal to rounding
        assert abs(e - e0) <= 1e-12 * abs(e0), key
    # ... and run-to-run identical for one kernel
    for k in (3, 4, 5):
        o = gsb.SparseMatrix.options(kernel=k)
        sp.gaussSeidel(bb, epsilon=0.0, max_iteration=5, options=o)
        e1 = sp.last_stats.last_eps[0]
        sp.gaussSeidel(bb, epsilon=0.0, max_iteration=5, options=o)
        assert sp.last_stats.last_eps[0] == e1
    # a general (multicolour) matrix through the staged/ring kernels
    r, c, v, b2, _ = wl.diag_dominant_system(7001, 6, seed=13)
    sg = gsb.SparseMatrix(np.float64)
    sg.initializeFromVector(r, c, v)
    outs = [sg.gaussSeidel(b2, epsilon=0.0, max_iteration=6, options=gsb.SparseMatrix.options(kernel=k)) for k in (1, 2, 3, 6, 0)]
    assert sg.last_stats.kernel_used == 6  # auto: a small system runs in one persistent launch
    assert all(np.array_equal(outs[0], o) for o in outs[1:])
    for k in (4, 5):  # no gather windows, more than two colours
        with pytest.raises(gsb.GsbError):
            sg.gaussSeidel(b2, epsilon=0.0, max_iteration=1, options=gsb.SparseMatrix.options(kernel=k))
    # auto on the grid: measured policy, see gsb_plan_effective_kernel
    sp.gaussSeidel(bb, epsilon=0.0, max_iteration=1)
    assert sp.last_stats.kernel_used == AUTO_GRID_KERNEL(nrhs)


def AUTO_GRID_KERNEL(nrhs):
    """300 x 217 x 2 colours is a small system: the whole solve in one persistent launch (kernel 6)."""
    return 6


def test_small_system_persistent_kernel(gsb, oracle_mod):
    """Kernel 6 (one launch per solve, grid barriers between the colours, stop rule inside the loop) against kernel 1:
    x bit for bit, the same stop sweep, on the lab3 4 x 4 system, BASELINE configs[0] (n = 1e4, multicolour), a small
    Poisson grid with three right-hand sides and a one-row system; and it is what small systems get by default."""
    from coursecomputationalphotography_b200 import workloads as wl
    K = lambda k, **kw: gsb.SparseMatrix.options(kernel=k, **kw)
    sp = gsb.SparseMatrix(np.int32)
    sp.initialize(4, 4, [10, -1, 2, 0, -1, 11, -1, 3, 2, -1, 10, -1, 0, 3, -1, 8])
    b = np.array([6.0, 25.0, -11.0, 15.0])
    x6 = sp.gaussSeidel(b, options=K(6))
    s6 = sp.last_stats.sweeps
    assert sp.last_stats.kernel_used == 6 and np.allclose(x6, [1, 2, -1, 1], atol=1e-6)
    x1 = sp.gaussSeidel(b, options=K(1, use_graph=0))
    assert sp.last_stats.sweeps == s6 and np.array_equal(x1, x6)
    assert sp.gaussSeidel(b) is not None and sp.last_stats.kernel_used == 6  # the default for small systems
    # configs[0]
    r, c, v, bb, xstar = wl.diag_dominant_system(10_000, 4, seed=42)
    sd = gsb.SparseMatrix(np.float64)
    sd.initializeFromVector(r, c, v)
    for eps, cap in ((1e-6, 1000), (0.0, 7), (1e-3, 1000)):
        y6 = sd.gaussSeidel(bb, epsilon=eps, max_iteration=cap, options=K(6))
        n6, e6 = sd.last_stats.sweeps, sd.last_stats.last_eps[0]
        assert sd.last_stats.kernel_used == 6 and sd.last_stats.kernel_launches <= 5
        y1 = sd.gaussSeidel(bb, epsilon=eps, max_iteration=cap, options=K(1, use_graph=0))
        assert sd.last_stats.sweeps == n6 and np.array_equal(y1, y6), (eps, cap)
        assert abs(sd.last_stats.last_eps[0] - e6) <= 1e-12 * max(abs(e6), 1e-300)
    assert np.abs(y6 - xstar).max() < 1e-2
    # small Poisson grid, three channels, odd sizes
    W, H = 97, 61
    img, bp = _poisson_rhs_for(gsb, wl, W, H, C=3)
    pg = gsb.SparseMatrix(np.float64)
    pg.poisson(W, H)
    z6 = pg.gaussSeidel(bp, epsilon=0.0, max_iteration=33, options=K(6))
    assert pg.last_stats.kernel_used == 6 and pg.last_stats.sweeps == 33
    assert np.array_equal(z6, pg.gaussSeidel(bp, epsilon=0.0, max_iteration=33, options=K(3, use_graph=0)))
    # a system that takes several CTAs (grid barrier in global memory), two right-hand sides
    W2, H2 = 300, 217
    img2, bp2 = _poisson_rhs_for(gsb, wl, W2, H2, C=3)
    pm = gsb.SparseMatrix(np.float64)
    pm.poisson(W2, H2)
    for cap in (1, 2, 25):
        w6 = pm.gaussSeidel(bp2[:2], epsilon=0.0, max_iteration=cap, options=K(6))
        assert pm.last_stats.kernel_used == 6 and pm.last_stats.sweeps == cap
        assert np.array_equal(w6, pm.gaussSeidel(bp2[:2], epsilon=0.0, max_iteration=cap, options=K(3, use_graph=0))), cap
    # repeated solves reuse the barrier words; a one-row system
    assert np.array_equal(z6, pg.gaussSeidel(bp, epsilon=0.0, max_iteration=33, options=K(6)))
    one = gsb.SparseMatrix(np.float64)
    one.initializeFromVector([0], [0], [4.0])
    assert one.gaussSeidel(np.array([2.0]), options=K(6))[0] == 0.5
    # check_every > 1 is not what this kernel implements
    with pytest.raises(gsb.GsbError):
        pg.gaussSeidel(bp, epsilon=0.0, max_iteration=4, options=K(6, check_every=2))


@pytest.mark.parametrize("W,H", [(128, 100), (127, 129)])
def test_small_cluster_two_passes_per_colour(gsb, W, H):
    """Kernel 6 on a thread-block cluster when a colour has more rows than the cluster has threads (8 x 512): the
    colour's phase then spans two passes that share one mbarrier phase.  1 to 4 right-hand sides (every instantiation
    of gs_small_cluster), x bit for bit against kernel 1, also over a long run with the stop rule armed.  (The rule's
    semantics on the cluster kernel are held by test_small_system_persistent_kernel.)"""
    from coursecomputationalphotography_b200 import workloads as wl
    K = lambda k, **kw: gsb.SparseMatrix.options(kernel=k, **kw)
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    _, b3 = _poisson_rhs_for(gsb, wl, W, H, C=3)
    b = np.concatenate([b3, b3[:1][:, ::-1]], axis=0)  # a fourth right-hand side
    assert W * H <= 16384 and (W * H + 1) // 2 > 8 * 512
    for k in (1, 2, 3, 4):
        bb = np.ascontiguousarray(b[:k]) if k > 1 else np.ascontiguousarray(b[0])
        x6 = sp.gaussSeidel(bb, epsilon=0.0, max_iteration=7, options=K(6))
        assert sp.last_stats.kernel_used == 6 and sp.last_stats.sweeps == 7
        x1 = sp.gaussSeidel(bb, epsilon=0.0, max_iteration=7, options=K(1, use_graph=0))
        assert np.array_equal(x6, x1), k
    # a longer run with the stop rule armed: the same sweep count (the rule, or the cap) and the same bits
    y6 = sp.gaussSeidel(b[0], epsilon=1e-1, max_iteration=300, options=K(6))
    s6 = sp.last_stats.sweeps
    y1 = sp.gaussSeidel(b[0], epsilon=1e-1, max_iteration=300, options=K(1, use_graph=0))
    assert sp.last_stats.sweeps == s6 and np.array_equal(y6, y1)


@pytest.mark.parametrize("W,H,nrhs", [(300, 217, 3), (1024, 512, 1), (1024, 512, 3), (2048, 1600, 3)])
def test_fused_sweep_wavefront_bitwise(gsb, W, H, nrhs):
    """Kernel 5 (both colours in one launch, colour 1 trailing colour 0 by `lead` tiles and waiting on per-tile
    flags) against the per-phase ring kernel: x AND the stop norm bit for bit, for leads from "just the dependency
    distance" (every colour-1 tile waits on tiles in flight) to "more than there are tiles" (degenerates to two
    phases), with and without graph replay, on the full grid and on a compact masked system whose two colours
    have different sizes."""
    from coursecomputationalphotography_b200 import workloads as wl
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    img, b = _poisson_rhs_for(gsb, wl, W, H, C=3)
    bb = b[:nrhs] if nrhs > 1 else b[0]
    ref = sp.gaussSeidel(bb, epsilon=0.0, max_iteration=9, options=gsb.SparseMatrix.options(kernel=3, use_graph=0))
    if W * H > 3_000_000:
        # leads around half the grid size (592 CTAs on a B200): a tile's dependency is then the CURRENT tile of another
        # CTA in the same wave -- the configuration that deadlocked before a waiting CTA retired its own tile first
        for lead in range(262, 322, 6):
            x = sp.gaussSeidel(bb, epsilon=0.0, max_iteration=9, options=gsb.SparseMatrix.options(kernel=5, fused_lead=lead))
            assert sp.last_stats.kernel_used == 5 and np.array_equal(x, ref), lead
        x = sp.gaussSeidel(bb, epsilon=0.0, max_iteration=9)  # auto: the fused sweep from 8 M rows on (measured policy)
        assert sp.last_stats.kernel_used == (5 if W * H >= (8 << 20) else 3) and np.array_equal(x, ref)
    for lead in (1, 2, 37, 0, 1 << 20):
        for graph in (0, 1):
            o = gsb.SparseMatrix.options(kernel=5, fused_lead=lead, use_graph=graph, batch_sweeps=4)
            x = sp.gaussSeidel(bb, epsilon=0.0, max_iteration=9, options=o)
            assert sp.last_stats.kernel_used == 5 and sp.last_stats.sweeps == 9
            assert np.array_equal(x, ref), (lead, graph)
    # compact masked system (Dirichlet blend): user colouring, unequal colour sizes, ragged tiles
    mask = wl.blob_mask(W // 2, H // 2, 0.35, 24, seed=5)
    guide = wl.synth_image(W // 2, H // 2, 3, seed=7)
    ro, ci, va, bm, pix, colors = wl.masked_poisson_system(mask, guide, guide[:, ::-1, ::-1])
    n = len(pix)
    sm = gsb.SparseMatrix(np.float64)
    sm.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    sm.analyze(gsb._lib.ORDER_USER, colors)
    bq = bm[:nrhs] if nrhs > 1 else bm[0]
    r1 = sm.gaussSeidel(bq, epsilon=0.0, max_iteration=15, options=gsb.SparseMatrix.options(kernel=1))
    e1 = list(sm.last_stats.last_eps)[:nrhs]
    for lead in (1, 5, 0):
        x = sm.gaussSeidel(bq, epsilon=0.0, max_iteration=15, options=gsb.SparseMatrix.options(kernel=5, fused_lead=lead))
        assert sm.last_stats.kernel_used == 5 and np.array_equal(x, r1), lead
        assert all(abs(a - c) <= 1e-12 * abs(c) for a, c in zip(list(sm.last_stats.last_eps)[:nrhs], e1))
    # the stop rule fires on the same sweep with the same accepted iterate
    x3 = sm.gaussSeidel(bq, epsilon=1e-2, max_iteration=20000, options=gsb.SparseMatrix.options(kernel=3))
    s3 = sm.last_stats.sweeps
    x5 = sm.gaussSeidel(bq, epsilon=1e-2, max_iteration=20000, options=gsb.SparseMatrix.options(kernel=5))
    assert 15 < s3 < 20000 and sm.last_stats.sweeps == s3 and np.array_equal(x3, x5)


def test_fused_sweep_needs_banded_two_colour_system(gsb):
    """A two-colour (bipartite) matrix with random coupling has no wavefront: kernel 5 is refused, auto keeps the
    per-phase kernels; a one-colour (diagonal) system likewise."""
    rng = np.random.default_rng(3)
    n = 200_000
    half = n // 2
    rows = np.repeat(np.arange(n), 3)
    cols = np.where(rows < half, rng.integers(half, n, rows.size), rng.integers(0, half, rows.size))
    key = np.unique(rows.astype(np.int64) * n + cols)
    r, c = (key // n).astype(np.int32), (key % n).astype(np.int32)
    v = rng.uniform(-1, 0, r.size)
    r = np.concatenate([r, np.arange(n, dtype=np.int32)])
    c = np.concatenate([c, np.arange(n, dtype=np.int32)])
    v = np.concatenate([v, np.full(n, 5.0)])
    o = np.lexsort((c, r))
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromVector(r[o], c[o], v[o])
    sp.analyze(gsb._lib.ORDER_USER, (np.arange(n) >= half).astype(np.int32))
    b = rng.standard_normal(n)
    x = sp.gaussSeidel(b, epsilon=0.0, max_iteration=4)
    assert sp.last_stats.kernel_used in (3, 4, 6) and sp.last_stats.n_colors == 2
    with pytest.raises(gsb.GsbError):
        sp.gaussSeidel(b, epsilon=0.0, max_iteration=1, options=gsb.SparseMatrix.options(kernel=5))
    assert np.array_equal(x, sp.gaussSeidel(b, epsilon=0.0, max_iteration=4, options=gsb.SparseMatrix.options(kernel=1)))


@pytest.mark.parametrize("W,H,nrhs", [(64, 48, 3), (300, 217, 1), (300, 217, 3), (1024, 512, 3)])
def test_dependent_launch_equals_graph_replay(gsb, W, H, nrhs):
    """Plain launches use programmatic dependent launch (the next colour phase stages its first tiles while the
    previous one drains); the CUDA-graph path does not.  Solution AND stop norm must agree bit for bit, every
    sweep -- a kernel that staged x_old before its writer had finished would show up in the norm only."""
    from coursecomputationalphotography_b200 import workloads as wl
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    img, b = _poisson_rhs_for(gsb, wl, W, H, C=3)
    bb = b[:nrhs] if nrhs > 1 else b[0]
    for k in (3, 4, 5):
        for sweeps in (1, 2, 7):
            out = []
            for graph in (1, 0):
                o = gsb.SparseMatrix.options(kernel=k, use_graph=graph, batch_sweeps=7)
                x = sp.gaussSeidel(bb, epsilon=0.0, max_iteration=sweeps, options=o)
                out.append((x, list(sp.last_stats.last_eps)[:nrhs], sp.last_stats.sweeps))
            assert out[0][2] == out[1][2] == sweeps
            assert np.array_equal(out[0][0], out[1][0]), (k, sweeps)
            assert out[0][1] == out[1][1], (k, sweeps, out[0][1], out[1][1])


def test_masked_blend_vs_reference_golden(gsb):
    """The converged device solution against the golden the UNMODIFIED reference produced (no oracle involved):
    max-abs <= 1e-4 * 255 (north-star tolerance), for the automatic ordering and for the parity colouring, one
    right-hand side at a time and the three channels fused."""
    import json
    import os
    from coursecomputationalphotography_b200 import workloads as wl
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")))["masked_blend_40"]
    mask = wl.blob_mask(G["W"], G["H"], G["coverage"], G["thickness"], seed=G["mask_seed"])
    guide = wl.synth_image(G["W"], G["H"], 3, seed=G["guide_seed"])
    target = wl.synth_image(G["W"], G["H"], 3, seed=G["target_seed"])
    ro, ci, va, b, pix, colors = wl.masked_poisson_system(mask, guide, target)
    n = len(pix)
    assert n == G["n"] and len(va) == G["nnz"]
    gold = np.array([[float.fromhex(v) for v in G["gs_default"][c]] for c in range(3)])
    tol = 1e-4 * 255.0
    for order in ("auto", "user"):
        sp = gsb.SparseMatrix(np.float64)
        sp.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
        if order == "user":
            sp.analyze(gsb._lib.ORDER_USER, colors)
        x3 = sp.gaussSeidel(b)  # reference defaults, three channels fused
        assert sp.last_stats.sweeps < 1000 and max(list(sp.last_stats.last_eps)[:3]) <= 1e-6
        assert np.abs(x3 - gold).max() <= tol, (order, np.abs(x3 - gold).max())
        x1 = sp.gaussSeidel(b[1])
        assert np.abs(x1 - gold[1]).max() <= tol
        print("masked 40x40 (%s): max-abs vs reference %.3e, %d sweeps, ||r|| %.3e" %
              (order, np.abs(x3 - gold).max(), sp.last_stats.sweeps, sp.residual(b[1], x1)))
