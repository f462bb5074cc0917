"""The BASELINE.json configurations that are not the bench workload, as parity cases:
  configs[0]  lab3 Gauss-Seidel, n = 1e4, ~5 nnz/row, reference defaults          (C1)
  configs[1]  1024 x 1024 single-channel Poisson blend                              (C2)
  configs[4]  unstructured random sparse SPD, ~27 nnz/row, multicolour ordering     (C5, at n = 3e5 / 2e6)
configs[2] is bench.py's workload (full-size properties in test_gs_gpu.py); configs[3] is the strip solver
(tests/test_dist.py)."""
import numpy as np
import pytest

from helpers import oracle_from_csr, permuted_csr

pytestmark = pytest.mark.gpu
TOL_IMAGE = 1e-4 * 255.0


def test_config0_lab3_diag_dominant_1e4(gsb, oracle_mod):
    from coursecomputationalphotography_b200 import workloads as wl
    n = 10_000
    r, c, v, b, xstar = wl.diag_dominant_system(n, 4, seed=42)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromVector(r, c, v)
    x = sp.gaussSeidel(b)  # reference defaults: eps 1e-6, 1000 sweeps
    o = oracle_mod.Oracle().init_from_vector(r, c, v)
    xo, sw, eps = o.gauss_seidel(b)
    st = sp.last_stats
    print("C1: gpu %d sweeps (%d colours, eps %.2e, %.3f ms)  cpu %d sweeps" % (st.sweeps, st.n_colors, st.last_eps[0],
                                                                             st.solve_ms, sw))
    assert st.last_eps[0] <= 1e-6 and st.sweeps < 60 and abs(st.sweeps - sw) <= 8
    assert np.abs(x - xo).max() <= 1e-6 and np.abs(x - xstar).max() <= 1e-6
    assert sp.residual(b, x) <= 1e-5 * np.linalg.norm(b)
    if oracle_mod.ref_available():  # the compiled reference itself
        xr = oracle_mod.Ref(2, "f64").init_from_vector(r, c, v).gauss_seidel(b)
        assert np.abs(x - xr).max() <= 1e-6


def test_config1_poisson_1024_masked_blend(gsb, oracle_mod):
    """1024^2 single channel: (a) one sweep of the reference-faithful full-grid system, bit-exact against the
    oracle on P A P^T; (b) the Dirichlet-masked blend converged against the oracle within 1e-4 of the range."""
    from coursecomputationalphotography_b200 import workloads as wl
    W = H = 1024
    img = wl.synth_image(W, H, 1, seed=7)
    gx, gy = wl.seamless_gradients(img)
    b = gsb.poisson_rhs(W, H, gx[0], gy[0], float(img[0, 0, 0]))
    sp = gsb.SparseMatrix(np.float64)
    sp.poisson(W, H)
    assert sp._nnz == 5_234_691
    x = sp.gaussSeidel(b, epsilon=0.0, max_iteration=3)
    perm, _ = sp.ordering()
    ro, ci, va = oracle_mod.poisson_csr(W, H)
    o = oracle_from_csr(oracle_mod, *permuted_csr(ro, ci, va, perm))
    xp, _, _ = o.gauss_seidel(b[perm], 0.0, 3)
    xo = np.empty_like(xp)
    xo[perm] = xp
    assert np.array_equal(x, xo)
    # (b) masked blend: thickness <= 32 px so that plain GS converges in a few thousand sweeps
    mask = wl.blob_mask(W, H, 0.30, 32, seed=11)
    guide, target = wl.synth_image(W, H, 1, seed=7), wl.synth_image(W, H, 1, seed=9)
    mro, mci, mva, mb, pix, colors = wl.masked_poisson_system(mask, guide, target)
    n = len(pix)
    sm = gsb.SparseMatrix(np.float64)
    sm.initializeFromEigenRowMajor(mva, len(mva), mro[:-1], n, mci, n)
    sm.analyze(gsb._lib.ORDER_USER, colors)
    xm = sm.gaussSeidel(mb[0], epsilon=1e-6, max_iteration=50000)
    om = oracle_from_csr(oracle_mod, mro, mci, mva)
    xmo, sw, _ = om.gauss_seidel(mb[0], 1e-6, 50000)
    err = np.abs(xm - xmo).max()
    print("C2 masked: n=%d  gpu %d sweeps / cpu %d sweeps, max-abs %.3e, ||r|| gpu %.3e cpu %.3e" %
          (n, sm.last_stats.sweeps, sw, err, sm.residual(mb[0], xm), np.linalg.norm(mb[0] - om.spmv(xmo))))
    assert sm.last_stats.sweeps < 50000 and err <= TOL_IMAGE
    assert np.array_equal(gsb.writeback_u8(xm), oracle_mod.writeback_u8(xmo)) or err <= TOL_IMAGE


def test_config4_random_spd_multicolor(gsb, oracle_mod):
    from coursecomputationalphotography_b200 import workloads as wl
    n = 300_000
    ro, ci, va, b, xstar = wl.random_spd_system(n, 13, seed=5)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    x4 = sp.gaussSeidel(b, epsilon=0.0, max_iteration=4)
    info = sp.coloring()
    assert info["ordering"] == gsb._lib.ORDER_MULTICOLOR and 8 <= info["n_colors"] <= 64
    perm, colors = sp.ordering()
    rows = np.repeat(np.arange(n), np.diff(ro))
    off = rows != ci
    assert not np.any(colors[rows[off]] == colors[ci[off]])  # proper colouring
    o = oracle_from_csr(oracle_mod, *permuted_csr(ro, ci, va, perm))
    xp, _, _ = o.gauss_seidel(b[perm], 0.0, 4)
    xo = np.empty_like(xp)
    xo[perm] = xp
    assert np.array_equal(x4, xo)  # per-sweep bit-exactness on P A P^T
    x = sp.gaussSeidel(b, epsilon=1e-8, max_iteration=3000)
    on = oracle_from_csr(oracle_mod, ro, ci, va)
    xn, sw, _ = on.gauss_seidel(b, 1e-8, 3000)
    print("C5: %d colours, gpu %d sweeps cpu %d sweeps, max-abs vs cpu %.2e, vs x* %.2e" %
          (info["n_colors"], sp.last_stats.sweeps, sw, np.abs(x - xn).max(), np.abs(x - xstar).max()))
    assert np.abs(x - xn).max() <= 1e-8 and np.abs(x - xstar).max() <= 1e-8
    assert np.array_equal(sp.applyToVector(x), on.spmv(x))  # SpMV in storage order is bit-exact


def test_config4_random_spd_1e6_triplets_vs_oracle(gsb, oracle_mod):
    """"CPU oracle on n = 1e6 of the same family".  The full-size path (unsorted triplets -> device radix
    sort -> slack CSR -> Jones-Plassmann -> sweeps, bench.py's c5 leg) at n = 1e6: the assembled layout equals the
    oracle's import of the host-sorted CSR, three sweeps are bit-exact against the oracle on P A P^T, and the converged
    solution matches the oracle's natural-order solve."""
    from coursecomputationalphotography_b200 import workloads as wl
    n = 1_000_000
    rows, cols, vals = wl.random_spd_coo(n, 13, seed=5)
    sp = gsb.SparseMatrix(np.float64)
    sp.initialize(n, n)
    sp.initializeFromTriplets(rows, cols, vals)
    # host CSR of the same triplets (last duplicate wins, as the insert loop of v2 :256-263 would leave it)
    key = rows.astype(np.int64) * n + cols
    order = np.argsort(key, kind="stable")
    ks, vs = key[order], vals[order]
    last = np.ones(ks.size, bool)
    last[:-1] = ks[1:] != ks[:-1]
    ks, vs = ks[last], vs[last]
    r, c = (ks // n).astype(np.int32), (ks % n).astype(np.int32)
    ro, ci, va = wl.coo_to_csr(r, c, vs, n)
    gv, gc, gb, gn, gl = sp.layout()
    assert sp._nnz == len(va) and np.array_equal(gn, np.diff(ro)) and np.array_equal(gb, ro[:-1])
    assert np.array_equal(gc, ci) and np.array_equal(gv, va) and not gl.any()
    rng = np.random.default_rng(1)
    xstar = rng.uniform(-1.0, 1.0, n)
    on = oracle_from_csr(oracle_mod, ro, ci, va)
    b = on.spmv(xstar)
    assert np.array_equal(sp.applyToVector(xstar), b)
    x3 = sp.gaussSeidel(b, epsilon=0.0, max_iteration=3)
    perm, colors = sp.ordering()
    assert not np.any(colors[r[r != c]] == colors[c[r != c]])
    o = oracle_from_csr(oracle_mod, *permuted_csr(ro, ci, va, perm))
    xp, _, _ = o.gauss_seidel(b[perm], 0.0, 3)
    xo = np.empty_like(xp)
    xo[perm] = xp
    assert np.array_equal(x3, xo)
    x = sp.gaussSeidel(b, epsilon=1e-8, max_iteration=1000)
    xn, sw, _ = on.gauss_seidel(b, 1e-8, 1000)
    print("C5 1e6: %d colours, gpu %d sweeps cpu %d sweeps, max-abs vs cpu %.2e" %
          (sp.last_stats.n_colors, sp.last_stats.sweeps, sw, np.abs(x - xn).max()))
    assert np.abs(x - xn).max() <= 1e-8 and np.abs(x - xstar).max() <= 1e-8


def test_config4_scale_2e6_residual(gsb):
    """Larger instance of the same family (2e6 rows, ~54e6 nnz): no oracle run, size-independent checks."""
    from coursecomputationalphotography_b200 import workloads as wl
    n = 2_000_000
    ro, ci, va, b, xstar = wl.random_spd_system(n, 13, seed=6)
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    x = sp.gaussSeidel(b, epsilon=1e-7, max_iteration=3000)
    st = sp.last_stats
    print("C5 2e6: %d colours, %d sweeps, %.1f ms, %.1f Gnnz/s" % (st.n_colors, st.sweeps, st.solve_ms,
                                                                  len(va) * st.sweeps / st.solve_ms / 1e6))
    assert st.sweeps < 3000 and st.last_eps[0] <= 1e-7
    assert np.abs(x - xstar).max() <= 1e-7
    assert sp.residual(b, x) <= 1e-6 * np.linalg.norm(b)


def test_config2_4096_masked_blend_vs_reference_golden(gsb):
    """BASELINE configs[2] at full size, converged: the 4096^2 x 3-channel Dirichlet-masked blend solved to the
    reference's stop rule, against the golden the UNMODIFIED reference gaussSeidel produced on this very system
    (tests/golden/c3_masked_4096.*, generator tests/golden/make_golden_c3.py, ~14 CPU-minutes per channel):
    max-abs <= 1e-4 * 255 on the 3 x 65 536 sampled unknowns, the written-back 8-bit values equal, and about the
    same number of sweeps as the reference's lexicographic order needs."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from make_golden_c3 import c3_system, sample_index
    meta = json.load(open(os.path.join(here, "golden", "c3_masked_4096.json")))
    gold = np.load(os.path.join(here, "golden", "c3_masked_4096.npz"))
    ro, ci, va, b, pix, colors = c3_system(4096)
    n = len(pix)
    assert n == meta["n"] and len(va) == meta["nnz"]
    idx = sample_index(n)
    assert np.array_equal(idx, gold["index"])
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    sp.analyze(gsb._lib.ORDER_USER, colors)
    for eps in (1e-5, 1e-6):
        x = sp.gaussSeidel(b, epsilon=eps, max_iteration=meta["cap"])
        st = sp.last_stats
        runs = [r for r in meta["runs"] if r["epsilon"] == eps]
        assert st.sweeps < meta["cap"] and max(list(st.last_eps)[:3]) <= eps
        worst = 0.0
        for c in range(3):
            ref = gold["x_eps%g_ch%d" % (eps, c)]
            worst = max(worst, float(np.abs(x[c][idx] - ref).max()))
            assert np.array_equal(gsb.writeback_u8(x[c][idx]), gsb.writeback_u8(ref))
        assert worst <= TOL_IMAGE, worst
        ref_sweeps = max(r["sweeps"] for r in runs)
        assert abs(st.sweeps - ref_sweeps) <= 0.02 * ref_sweeps, (st.sweeps, ref_sweeps)
        print("C3 4096^2 masked, eps %g: %d sweeps (reference %s), max-abs vs reference %.3e, %.0f ms" %
              (eps, st.sweeps, [r["sweeps"] for r in runs], worst, st.solve_ms))
