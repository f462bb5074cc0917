"""GPU parity tests of the container's bulk entry points (A0-A5, A9, A10) through the C ABI.
Bit-exact against the oracle (and through it the compiled reference, see test_oracle_vs_reference.py)."""
import json
import os

import numpy as np
import pytest

from helpers import random_sorted_coo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")


def _same_layout(sp, lay, dtype):
    v, c, rb, rn, rl = sp.layout()
    assert sp.rows() == lay.n_rows and sp.cols() == lay.n_cols
    assert np.array_equal(v, lay.values.astype(dtype))
    assert np.array_equal(c, lay.cols)
    assert np.array_equal(rb, lay.row_begin)
    assert np.array_equal(rn, lay.row_nnz)
    assert np.array_equal(rl, lay.row_left)


def test_lab3_fixture_layout_and_inserts(gsb):
    """main6.cc:192-231: the 3x5 fixture, then T1-T5, each checked against the dense mirror (CheckEqual)
    and against the layouts the reference itself produced (tests/golden/ref_golden.json)."""
    gold = json.load(open(GOLD))["lab3_fixture"]
    mat = np.array([[1, 0, 0, 1, 0], [0, 0, 0, 0, 0], [8, 0, 1, 0, 0]], np.int32)
    sp = gsb.SparseMatrix(np.int32)
    sp.initializeFromVector([0, 0, 0, 2, 2], [0, 3, 4, 0, 2], [1, 1, 0, 8, 1])

    def check(step):
        dense = np.array([[sp.at(i, j) for j in range(sp.cols())] for i in range(sp.rows())])
        assert np.array_equal(dense, mat), step
        g = gold[step]
        v, c, rb, rn, rl = sp.layout()
        assert list(rb) == g["row_begin"] and list(rn) == g["row_nnz"] and list(rl) == g["row_left"], step
        # live entries must agree with the reference; slack slots are compared where the reference's
        # defects do not touch them
        for r in range(sp.rows()):
            a, k = rb[r], rn[r]
            assert list(c[a:a + k]) == g["cols"][a:a + k] and list(v[a:a + k]) == g["vals"][a:a + k], step
        # the device copy answers the same queries
        ii, jj = np.divmod(np.arange(mat.size), mat.shape[1])
        assert np.array_equal(sp.at_many(ii, jj), mat.ravel().astype(np.float64)), step

    check("init")
    for step, (x, r, c) in (("T1", (0, 1, 0)), ("T2", (0, 0, 0)), ("T3", (1, 2, 2)), ("T4", (8, 0, 0)),
                            ("T5", (9, 1, 1))):
        sp.insert(x, r, c)
        mat[r, c] = x
        check(step)


@pytest.mark.parametrize("dtype", [np.int32, np.float64])
@pytest.mark.parametrize("shape,density", [((40, 40), 0.3), ((300, 200), 0.05), ((1, 9), 0.9), ((7, 1), 0.5),
                                           ((2000, 2000), 0.004)])
def test_initialize_from_vector_bitexact(gsb, oracle_mod, dtype, shape, density):
    rng = np.random.default_rng(hash((shape, density)) % 2**32)
    r, c, v = random_sorted_coo(rng, shape[0], shape[1], density, 0.25, dtype)
    sp = gsb.SparseMatrix(dtype)
    sp.initializeFromVector(r, c, v)
    o = oracle_mod.Oracle().init_from_vector(r, c, v)
    _same_layout(sp, o.layout(), dtype)
    if oracle_mod.ref_available():
        ref = oracle_mod.Ref(2, "i32" if dtype == np.int32 else "f64").init_from_vector(r, c, v)
        _same_layout(sp, ref.layout(), dtype)


def test_initialize_from_vector_edge_cases(gsb, oracle_mod):
    cases = [
        ([5], [3], [2.5]),                                  # single entry far from the origin: rows 0..4 empty
        ([0, 0, 0], [0, 1, 2], [0.0, 0.0, 0.0]),            # a row of explicit zeros only
        ([0, 3, 3, 9], [1, 0, 5, 2], [1.0, 0.0, -0.0, 4.0]),  # -0.0 counts as zero; empty rows in between
        (np.zeros(5000, np.int32), np.arange(5000), np.r_[np.zeros(2500), np.ones(2500)]),  # one long row
    ]
    for r, c, v in cases:
        sp = gsb.SparseMatrix(np.float64)
        sp.initializeFromVector(r, c, v)
        o = oracle_mod.Oracle().init_from_vector(r, c, v)
        _same_layout(sp, o.layout(), np.float64)
    sp = gsb.SparseMatrix(np.float64)
    with pytest.raises(gsb.GsbError) as e:
        sp.initializeFromVector([2, 1, 3], [0, 0, 0], [1.0, 1.0, 1.0])
    assert e.value.status == 3  # GSB_ERR_UNSORTED
    with pytest.raises(gsb.GsbError):
        sp.initializeFromVector([], [], [])


def test_large_assembly_matches_oracle(gsb, oracle_mod):
    """config-1 scale and beyond: 2e6 entries, 10 % explicit zeros."""
    rng = np.random.default_rng(12)
    n_rows, per = 200_000, 10
    r = np.repeat(np.arange(n_rows, dtype=np.int32), per)
    c = np.sort(rng.integers(0, 1_000_000, (n_rows, per)), axis=1).ravel().astype(np.int32)
    v = rng.uniform(-1, 1, r.size)
    v[rng.random(r.size) < 0.1] = 0.0
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromVector(r, c, v)
    o = oracle_mod.Oracle().init_from_vector(r, c, v)
    _same_layout(sp, o.layout(), np.float64)


def test_import_csr_bitexact(gsb, oracle_mod):
    rng = np.random.default_rng(4)
    W, H = 4, 3
    ro, ci, va = oracle_mod.poisson_csr(W, H)
    n = W * H
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    gold = json.load(open(GOLD))["poisson_4x3_import"]
    v, c, rb, rn, rl = sp.layout()
    assert list(rb) == gold["row_begin"] and list(rn) == gold["row_nnz"] and list(rl) == gold["row_left"]
    assert list(c) == gold["cols"] and list(v) == gold["vals"]
    # compressed, with several trailing empty rows (v2 :608-614) and with none
    for tail in (0, 1, 4):
        lens = np.r_[rng.integers(0, 6, 50), np.zeros(tail, np.int64)]
        if tail == 0:
            lens[-1] = 3
        off = np.r_[0, np.cumsum(lens)].astype(np.int32)
        cols = np.concatenate([np.sort(rng.choice(60, k, replace=False)) for k in lens] + [np.zeros(0, np.int64)])
        vals = rng.uniform(1, 2, off[-1])
        sp = gsb.SparseMatrix(np.float64)
        sp.initializeFromEigenRowMajor(vals, len(vals), off[:-1], len(lens), cols, 60)
        o = oracle_mod.Oracle().import_csr(vals, off[:-1], cols, 60)
        _same_layout(sp, o.layout(), np.float64)
    # uncompressed (per-row counts given, v2 :560-589): rows with holes
    cap = rng.integers(2, 7, 40)
    used = np.minimum(cap, rng.integers(0, 7, 40))
    off = np.r_[0, np.cumsum(cap)].astype(np.int32)
    vals = rng.uniform(1, 2, off[-1])
    cols = np.concatenate([np.sort(rng.choice(50, k, replace=False)) for k in cap])
    sp = gsb.SparseMatrix(np.float64)
    sp.initializeFromEigenRowMajor(vals, len(vals), off[:-1], 40, cols, 50, used, 40)
    o = oracle_mod.Oracle().import_csr(vals, off[:-1], cols, 50, used)
    _same_layout(sp, o.layout(), np.float64)


def test_triplets_unsorted_with_duplicates(gsb):
    rng = np.random.default_rng(21)
    nr, nc, k = 300, 257, 20000
    r = rng.integers(0, nr, k).astype(np.int32)
    c = rng.integers(0, nc, k).astype(np.int32)
    v = rng.integers(-3, 4, k).astype(np.float64)  # plenty of zeros and duplicates
    dense = np.zeros((nr, nc))
    for i in range(k):  # the reference's insert() loop semantics: the last write wins, zero clears
        dense[r[i], c[i]] = v[i]
    sp = gsb.SparseMatrix(np.float64)
    sp.initialize(nr, nc)
    sp.initializeFromTriplets(r, c, v)
    vals, cols, rb, rn, rl = sp.layout()
    assert sp.rows() == nr and sp.cols() == nc and rl.sum() == 0 and rn.sum() == np.count_nonzero(dense)
    got = np.zeros_like(dense)
    for i in range(nr):
        cc = cols[rb[i]:rb[i] + rn[i]]
        assert np.all(np.diff(cc) > 0)
        got[i, cc] = vals[rb[i]:rb[i] + rn[i]]
    assert np.array_equal(got, dense)
    ii, jj = rng.integers(0, nr, 5000), rng.integers(0, nc, 5000)
    assert np.array_equal(sp.at_many(ii, jj), dense[ii, jj])


def test_poisson_assembly_matches_oracle_import(gsb, oracle_mod):
    from coursecomputationalphotography_b200 import workloads as wl
    for W, H in ((4, 3), (31, 17), (2, 2), (64, 1), (1, 5), (1, 1), (130, 77)):
        sp = gsb.SparseMatrix(np.float64)
        sp.poisson(W, H)
        n = W * H
        if W > 1 and H > 1:
            ro, ci, va = oracle_mod.poisson_csr(W, H)
            o = oracle_mod.Oracle().import_csr(va, ro[:-1], ci, n)
            _same_layout(sp, o.layout(), np.float64)
        else:  # one-pixel-wide grids have no forward-difference rows at all: only the pin on pixel 0 survives
            v, c, rb, rn, rl = sp.layout()
            assert list(rn) == [1] + [0] * (n - 1) and list(v) == [1.0] and list(c) == [0] and rl.sum() == 0
            continue
        img = wl.synth_image(W, H, 3, seed=W * 100 + H)
        gx, gy = wl.forward_gradients(img)
        b = gsb.poisson_rhs(W, H, gx, gy, img[:, 0, 0].astype(np.float64))
        for ch in range(3):
            assert np.array_equal(b[ch], oracle_mod.poisson_rhs(W, H, gx[ch], gy[ch], float(img[ch, 0, 0])))


def test_writeback_u8(gsb, oracle_mod):
    x = np.array([-5.0, -0.0, 0.0, 0.999, 1.0, 127.5, 254.999, 255.0, 255.5, 1e9, np.nan, -np.inf, np.inf])
    assert np.array_equal(gsb.writeback_u8(x), oracle_mod.writeback_u8(x))
    rng = np.random.default_rng(3)
    y = rng.uniform(-50, 300, 100003)
    assert np.array_equal(gsb.writeback_u8(y), oracle_mod.writeback_u8(y))
