"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def permuted_csr(ro, ci, va, perm):
    """P A P^T in compressed CSR with ascending columns; perm[new] = old."""
    n = len(perm)
    iperm = np.empty(n, np.int64)
    iperm[perm] = np.arange(n)
    lens = (ro[1:] - ro[:-1])[perm]
    new_ro = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=new_ro[1:])
    rows_old = np.repeat(np.arange(n), ro[1:] - ro[:-1])
    r_new = iperm[rows_old]
    c_new = iperm[ci]
    order = np.lexsort((c_new, r_new))
    return new_ro.astype(np.int32), c_new[order].astype(np.int32), np.asarray(va, np.float64)[order]


def compact_csr_from_layout(values, cols, row_begin, row_nnz):
    """Live entries of a slack-CSR layout as compressed CSR."""
    n = len(row_begin)
    ro = np.zeros(n + 1, np.int64)
    np.cumsum(row_nnz, out=ro[1:])
    idx = np.concatenate([np.arange(b, b + k) for b, k in zip(row_begin, row_nnz)]) if n else np.zeros(0, np.int64)
    idx = idx.astype(np.int64)
    return ro.astype(np.int32), np.asarray(cols)[idx].astype(np.int32), np.asarray(values, np.float64)[idx]


def oracle_from_csr(pyoracle, ro, ci, va, n_cols=None):
    o = pyoracle.Oracle()
    n = len(ro) - 1
    o.import_csr(va, np.ascontiguousarray(ro[:-1]), ci, n if n_cols is None else n_cols)
    return o


def random_sorted_coo(rng, n_rows, n_cols, density, zero_frac=0.2, dtype=np.float64, force_last=True):
    """Sorted COO with explicit zeros, as the lab3 benchmark builds it (main6.cc:106-123)."""
    m = rng.random((n_rows, n_cols)) < density
    r, c = np.nonzero(m)
    if dtype == np.int32:
        v = rng.integers(1, 10000, r.size).astype(np.int32)
    else:
        v = rng.uniform(-5, 5, r.size)
    z = rng.random(r.size) < zero_frac
    v = np.where(z, 0, v).astype(dtype)
    if force_last and (r.size == 0 or r[-1] != n_rows - 1 or c[-1] != n_cols - 1):
        r = np.append(r, n_rows - 1)
        c = np.append(c, n_cols - 1)
        v = np.append(v, dtype(0) if dtype == np.int32 else 0.0).astype(dtype)
    return r.astype(np.int32), c.astype(np.int32), v
