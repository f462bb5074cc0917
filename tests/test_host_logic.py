"""CPU tests of host-side logic: input generators, the Python mirror's host-resident access/modify
path (the part of the reference API that stays on the host), strip partitioning."""
import numpy as np
import pytest


def test_generators_are_deterministic_and_well_formed():
    from coursecomputationalphotography_b200 import workloads as wl
    r, c, v, b, xs = wl.diag_dominant_system(2000, 4, seed=42)
    r2, c2, v2, b2, _ = wl.diag_dominant_system(2000, 4, seed=42)
    assert np.array_equal(r, r2) and np.array_equal(c, c2) and np.array_equal(v, v2) and np.array_equal(b, b2)
    key = r.astype(np.int64) * 2000 + c
    assert np.all(np.diff(key) > 0)  # sorted by (row, col), no duplicates: the initializeFromVector contract
    diag = v[r == c]
    off = np.zeros(2000)
    np.add.at(off, r[r != c], np.abs(v[r != c]))
    assert len(diag) == 2000 and np.all(diag > off)  # strictly diagonally dominant
    img = wl.synth_image(40, 30, 3, seed=1)
    assert img.dtype == np.uint8 and img.shape == (3, 30, 40)
    gx, gy = wl.forward_gradients(img)
    assert gx.dtype == np.float32 and np.all(gx[:, :, -1] == 0) and np.all(gy[:, -1, :] == 0)
    assert np.array_equal(gx[0, 3, 4], np.float32(int(img[0, 3, 5]) - int(img[0, 3, 4])))
    assert wl.poisson_nnz(4096, 4096) == 83_853_315 and wl.poisson_nnz(1024, 1024) == 5_234_691


def test_masked_system_is_a_dirichlet_5_point_operator():
    from coursecomputationalphotography_b200 import workloads as wl
    mask = wl.blob_mask(96, 80, 0.3, 16, seed=2)
    assert not mask[0].any() and not mask[-1].any() and not mask[:, 0].any() and not mask[:, -1].any()
    g, t = wl.synth_image(96, 80, 2, seed=3), wl.synth_image(96, 80, 2, seed=4)
    ro, ci, va, b, pix, colors = wl.masked_poisson_system(mask, g, t)
    n = len(pix)
    assert n == mask.sum() and ro[-1] == len(ci) == len(va)
    rows = np.repeat(np.arange(n), np.diff(ro))
    assert np.all(va[rows == ci] == 4.0) and np.all(va[rows != ci] == -1.0)
    assert np.all(colors[rows[rows != ci]] != colors[ci[rows != ci]])  # parity is a proper 2-colouring
    # if guide == target the blend reproduces the target exactly: A t = b
    ro, ci, va, b, pix, _ = wl.masked_poisson_system(mask, t, t)
    tx = t.reshape(2, -1)[:, pix].astype(np.float64)
    Ax = np.zeros_like(tx)
    for ch in range(2):
        np.add.at(Ax[ch], rows, va * tx[ch][ci])
    assert np.array_equal(Ax, b)


def test_strip_bounds_cover_the_image():
    from coursecomputationalphotography_b200 import workloads as wl
    for H, world in ((4096, 8), (16384, 4), (10, 3), (7, 7), (5, 2)):
        s = wl.strip_bounds(H, world)
        assert s[0][0] == 0 and s[-1][1] == H and all(a[1] == b[0] for a, b in zip(s, s[1:]))
        sizes = [b - a for a, b in s]
        assert max(sizes) - min(sizes) <= 1


def test_python_mirror_host_side_insert_matches_oracle(gsb, oracle_mod):
    """at()/insert() live on the host copy of the five arrays (as in the C++ header); starting from
    initialize(r, c) no device call is needed, so this runs without a GPU."""
    if gsb._lib.device_count() > 0:
        pytest.skip("covered by the GPU tests")
    rng = np.random.default_rng(0)

    class HostOnly(gsb.SparseMatrix):
        def __init__(self, dtype):  # no device handle: only the host-resident part of the mirror is used
            self.dtype = np.dtype(dtype)
            self._h = None

        def __del__(self):
            pass

    sp = HostOnly(np.float64)
    gsb.SparseMatrix.initialize(sp, 12, 9)
    dense = np.zeros((12, 9))
    for _ in range(400):
        i, j = int(rng.integers(0, 12)), int(rng.integers(0, 9))
        v = float(rng.integers(-2, 3))
        sp.insert(v, i, j)
        dense[i, j] = v
    got = np.array([[sp.at(i, j) for j in range(9)] for i in range(12)])
    assert np.array_equal(got, dense)
    vals, cols, rb, rn, rl = sp.values_, sp.col_offset_, sp.row_begin_, sp.row_num_nze_, sp.row_space_left_
    for i in range(12):
        cc = cols[rb[i]:rb[i] + rn[i]]
        assert np.all(np.diff(cc) > 0)  # rows stay sorted (the reference's insert does not guarantee this)
    assert rb[-1] + rn[-1] + rl[-1] == len(vals)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours): one JSON line with the contract's
    keys, produced by the compiled reference when it is present, else by the restatement."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--size", "256",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Gnnz/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    # the workload-defining keys are the SAME dict the GPU arm and the strip arm emit (what each arm did with the
    # workload is in "run"), and the warm-up clamp is the GPU arm's
    sys.path.insert(0, root)
    import bench
    from coursecomputationalphotography_b200 import workloads as wl
    assert d["config"] == bench.workload_config(256, 256, 3, 256 * 256, wl.poisson_nnz(256, 256))
    assert d["warmup"] == 3 and "sweeps_per_step" in d["run"] and "sweeps_per_step" not in d["config"]


def test_bench_gpu_arm_fails_loudly_without_a_gpu(gsb):
    """No CPU fallback behind the GPU arm; the time-to-tolerance child reports its failure instead of raising."""
    if gsb._lib.device_count() > 0:
        pytest.skip("needs a machine without a GPU")
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=root)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
    sys.path.insert(0, root)
    import bench

    class A:
        size, channels, check_every, kernel = 64, 3, 1, 0
    out = bench.time_to_tol_child(A)
    assert "error" in out and "NO_DEVICE" in out["error"]


def test_e2e_warmup_rule():
    """bench.py's e2e leg: at least three untimed steps, then timed as soon as two consecutive steps agree to 10 %, at
    most twelve warm-up steps."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    done = bench.e2e_warm_done
    assert not done([]) and not done([80.0, 80.0])
    assert done([1700.0, 80.0, 81.0])              # settled by the third step
    assert not done([1700.0, 80.0, 120.0])         # the last two differ by more than 10 %
    assert done([1700.0, 80.0, 120.0, 118.0])
    assert done([100.0 * 1.5 ** i for i in range(12)])   # never settles: stop warming up after twelve
    assert not done([100.0 * 1.5 ** i for i in range(11)])
    assert done([5.0, 5.0, 5.0, 5.0], warm_min=4) and not done([5.0, 5.0, 5.0], warm_min=4)
