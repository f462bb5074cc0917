"""Prints a hash of the solution after a fixed number of sweeps (used to compare launch modes bit for bit).
usage: python tools/x_hash.py SIZE CHANNELS SWEEPS [KERNEL]"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coursecomputationalphotography_b200 as pkg
from coursecomputationalphotography_b200 import workloads as wl

W = H = int(sys.argv[1])
ch, sweeps = int(sys.argv[2]), int(sys.argv[3])
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else 0
img = wl.synth_image(W, H, ch, seed=7)
gx, gy = wl.seamless_gradients(img)
b = pkg.poisson_rhs(W, H, gx, gy, img[:, 0, 0].astype(np.float64))
sp = pkg.SparseMatrix(np.float64)
sp.poisson(W, H)
b = np.asarray(b).reshape(ch, -1)
x = sp.gaussSeidel(b if ch > 1 else b[0], epsilon=0.0, max_iteration=sweeps,
                   options=pkg.SparseMatrix.options(kernel=kernel, use_graph=0))
print(hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest(), sp.last_stats.kernel_used, sp.last_stats.sweeps,
      "eps", list(sp.last_stats.last_eps)[:ch])
