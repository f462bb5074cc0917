"""Generates tests/golden/c3_masked_<size>.npz (+ .json): the REFERENCE's own gaussSeidel
(labs/lab8/src/OpenCVHW1/sparse-matrix.h:350-380, compiled unmodified into oracle/_ref/libgsref.so) run to its
stop rule on BASELINE configs[2]'s converged-parity system: the Dirichlet-masked 5-point blend,
three channels, x0 = 1, natural (lexicographic) order.

    python tests/golden/make_golden_c3.py [--size 4096] [--eps 1e-5 1e-6] [--cap 40000]

Costs tens of CPU-minutes at 4096^2; run where /root/reference exists.  What is kept (the full x is 120 MB):
  * per epsilon and channel: sweeps to stop, ||b - A x||_2, SHA-256 of x (float64 bytes) and of the u8
    write-back, a strided sample of x (float64, bit-exact) with its indices, wall seconds of the reference call;
  * the sweep count comes from the oracle port's traced run of the same loop (the reference returns only x);
    `oracle_bitwise` records that the port's iterate at that sweep has the reference's SHA-256, which pins the count.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from coursecomputationalphotography_b200 import workloads as wl  # noqa: E402

N_SAMPLE = 65536


def c3_system(size, channels=3):
    """The system bench.py's time-to-tolerance leg and tests/ solve (same generator, same seeds)."""
    return wl.c3_masked_system(size, channels)


def sample_index(n):
    return np.unique(np.linspace(0, n - 1, N_SAMPLE).astype(np.int64))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--eps", type=float, nargs="+", default=[1e-5, 1e-6])
    ap.add_argument("--cap", type=int, default=40000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    eps_list = sorted(a.eps, reverse=True)
    pyoracle.build()
    assert pyoracle.ref_available(), "needs oracle/_ref/libgsref.so (make -C oracle where /root/reference exists)"
    t0 = time.time()
    ro, ci, va, b, pix, colors = c3_system(a.size)
    n, nnz, ch = len(pix), len(va), b.shape[0]
    print("system: n=%d nnz=%d (%.1f s)" % (n, nnz, time.time() - t0), flush=True)
    ref = pyoracle.Ref(2, "f64").import_csr(va, ro[:-1], ci, n)
    orc = pyoracle.Oracle().import_csr(va, ro[:-1], ci, n)
    idx = sample_index(n)
    res = {}
    lock = threading.Lock()

    def ref_run(e, c):
        t = time.time()
        x = ref.gauss_seidel(b[c], e, a.cap)
        dt = time.time() - t
        r = b[c] - ref.spmv(x)
        out = {"x_sha256": sha(x), "u8_sha256": sha(pyoracle.writeback_u8(x)), "residual_l2": float(np.sqrt(r @ r)),
               "sample": x[idx].copy(), "cpu_s": dt}
        with lock:
            res[("ref", e, c)] = out
            print("ref eps=%g ch=%d: %.0f s  resid %.3e" % (e, c, dt, out["residual_l2"]), flush=True)

    def orc_run(c):
        t = time.time()
        xs, sw, hist = orc.gauss_seidel_trace(b[c], eps_list, a.cap)
        with lock:
            for k, e in enumerate(eps_list):
                res[("orc", e, c)] = {"sweeps": int(sw[k]), "x_sha256": sha(xs[k]),
                                      "last_eps": float(hist[sw[k] - 1])}
            print("oracle trace ch=%d: sweeps %s (%.0f s)" % (c, sw.tolist(), time.time() - t), flush=True)

    def run_all(ths):
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    # phase 1: the first epsilon on `ch` threads, one channel each (the bench's CPU arm layout), next to the traced port
    run_all([threading.Thread(target=ref_run, args=(eps_list[0], c)) for c in range(ch)] +
            [threading.Thread(target=orc_run, args=(c,)) for c in range(ch)])
    for e in eps_list[1:]:
        run_all([threading.Thread(target=ref_run, args=(e, c)) for c in range(ch)])

    meta = {"source": "oracle/_ref/libgsref.so gaussSeidel (reference v2 :350-380), g++ -O2, x0 = 1, natural order",
            "generator": "tests/golden/make_golden_c3.py --size %d" % a.size, "size": a.size, "n": n, "nnz": nnz,
            "channels": ch, "cap": a.cap, "host_threads": "%d concurrent reference solves + %d traced port solves "
            "on %d cores" % (ch, ch, os.cpu_count()), "runs": []}
    arrays = {"index": idx}
    for e in eps_list:
        for c in range(ch):
            r, o = res[("ref", e, c)], res[("orc", e, c)]
            meta["runs"].append({"epsilon": e, "channel": c, "sweeps": o["sweeps"], "last_eps": o["last_eps"],
                                 "stopped": bool(o["sweeps"] < a.cap), "residual_l2": r["residual_l2"],
                                 "x_sha256": r["x_sha256"], "u8_sha256": r["u8_sha256"], "cpu_s": r["cpu_s"],
                                 "oracle_bitwise": bool(o["x_sha256"] == r["x_sha256"])})
            arrays["x_eps%g_ch%d" % (e, c)] = r["sample"]
    out = a.out or os.path.join(ROOT, "tests", "golden", "c3_masked_%d" % a.size)
    np.savez_compressed(out + ".npz", **arrays)
    json.dump(meta, open(out + ".json", "w"), indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
