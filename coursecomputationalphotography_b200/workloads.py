"""Synthetic inputs for the five BASELINE.json configurations, numpy only.

Everything here is input generation: images, gradients as the reference computes them
(GradientAt, project/src/PhotoMontage/PhotoMontage.cpp:399-408: forward differences of 8-bit
pixels stored as float32), masks, and CSR systems.  Fixed seeds; the same arrays feed the GPU
path and the CPU checkers used by tests/ and bench.py.
"""
import numpy as np


# ---- C1: lab3 Gauss-Seidel on a diagonally dominant sparse system -------------------------------
def diag_dominant_system(n=10_000, off_per_row=4, seed=42):
    """Sorted COO (rows, cols, vals) with `off_per_row` distinct random off-diagonals per row,
    values U(-1,1), diagonal = sum|off| + 1; x* ~ U(-100,100); b = A x*."""
    rng = np.random.default_rng(seed)
    cols = np.empty((n, off_per_row + 1), np.int64)
    for k in range(off_per_row):
        cols[:, k] = rng.integers(0, n, n)
    cols[:, off_per_row] = np.arange(n)
    vals = rng.uniform(-1.0, 1.0, (n, off_per_row + 1))
    # drop duplicate columns inside a row (keep the first), and clashes with the diagonal
    order = np.argsort(cols, axis=1, kind="stable")
    cols = np.take_along_axis(cols, order, 1)
    vals = np.take_along_axis(vals, order, 1)
    is_diag = cols == np.arange(n)[:, None]
    dup = np.zeros_like(is_diag)
    dup[:, 1:] = cols[:, 1:] == cols[:, :-1]
    # among equal columns keep exactly one; if the diagonal is among them keep a single entry as diagonal
    keep = ~dup
    vals = np.where(keep, vals, 0.0)
    offsum = np.where(keep & ~is_diag, np.abs(vals), 0.0).sum(1)
    first_diag = is_diag & keep
    vals = np.where(first_diag, (offsum + 1.0)[:, None], vals)
    rows = np.repeat(np.arange(n), off_per_row + 1).reshape(n, -1)
    m = keep.ravel()
    r, c, v = rows.ravel()[m], cols.ravel()[m], vals.ravel()[m]
    xstar = rng.uniform(-100.0, 100.0, n)
    b = np.zeros(n)
    np.add.at(b, r, v * xstar[c])
    return r.astype(np.int32), c.astype(np.int32), v.astype(np.float64), b, xstar


def coo_to_csr(rows, cols, vals, n):
    ro = np.zeros(n + 1, np.int64)
    np.add.at(ro, rows.astype(np.int64) + 1, 1)
    return np.cumsum(ro).astype(np.int32), cols.astype(np.int32), vals.astype(np.float64)


# ---- images and gradients -------------------------------------------------------------------------
def _pixel_noise(x, y, c, seed):
    """Deterministic per-pixel noise in -3..3 (integer hash), so any row range can be generated on its own."""
    h = (x.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + y.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
         + np.uint64((c * 0x165667B19E3779F9 + seed * 0x27D4EB2F165667C5) & 0xFFFFFFFFFFFFFFFF))
    h ^= h >> np.uint64(29)
    h *= np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    return (h % np.uint64(7)).astype(np.int64) - 3


def synth_image(W, H, channels=1, seed=7, y0=0, y1=None):
    """uint8 image rows [y0, y1) of a W x H picture: smooth low-frequency field plus a wavy vertical seam
    step (two 'exposures') plus per-pixel noise.  Every pixel depends only on (x, y, channel, seed)."""
    y1 = H if y1 is None else y1
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[y0:y1, 0:W].astype(np.float64)
    yi, xi = np.mgrid[y0:y1, 0:W]
    out = np.empty((channels, y1 - y0, W), np.uint8)
    for c in range(channels):
        ph = rng.uniform(0, 2 * np.pi, 4)
        f = (110 + 60 * np.sin(2 * np.pi * x / max(W, 2) * 1.5 + ph[0]) * np.cos(2 * np.pi * y / max(H, 2) + ph[1])
             + 25 * np.sin(2 * np.pi * (x + 2 * y) / max(W + H, 2) * 3 + ph[2]))
        seam = W // 2 + (8 * np.sin(2 * np.pi * y / max(H, 2) * 2 + ph[3])).astype(np.int64)
        f = f + np.where(x >= seam, 28.0 + 4 * c, 0.0)
        f += _pixel_noise(xi, yi, c, seed)
        out[c] = np.clip(f, 0, 255).astype(np.uint8)
    return out


def forward_gradients(img):
    """GradientAt over y < H-1, x < W-1 (PhotoMontage.cpp:399-425): float32 arrays (C, H, W); the last
    row/column stay 0 (never read by the system)."""
    img = img.astype(np.int32)
    gx = np.zeros(img.shape, np.float32)
    gy = np.zeros(img.shape, np.float32)
    gx[:, :-1, :-1] = (img[:, :-1, 1:] - img[:, :-1, :-1]).astype(np.float32)
    gy[:, :-1, :-1] = (img[:, 1:, :-1] - img[:, :-1, :-1]).astype(np.float32)
    return gx, gy


def seamless_gradients(img, seed=7):
    """Gradients of the two half-images with the seam step removed: what gradient-domain fusion feeds the
    solver (the blended result should not show the exposure step)."""
    gx, gy = forward_gradients(img)
    big = np.abs(gx) > 20
    gx = np.where(big, 0.0, gx).astype(np.float32)
    return gx, gy


def strip_gradients(W, H, channels, y0, y1, seed=7):
    """Seam-free forward gradients for image rows [ya, y1) with ya = max(y0-1, 0): what the strip solver's
    right-hand side generator (gsb_poisson_rhs_rows) needs for the strip [y0, y1).  Returns gx, gy, ya and
    the image's pixel (0,0) per channel (the pin constraint)."""
    ya = max(y0 - 1, 0)
    yb = min(y1 + 1, H)
    img = synth_image(W, H, channels, seed=seed, y0=ya, y1=yb).astype(np.int32)
    rows = y1 - ya
    gx = np.zeros((channels, rows, W), np.float32)
    gy = np.zeros((channels, rows, W), np.float32)
    gx[:, :, :-1] = (img[:, :rows, 1:] - img[:, :rows, :-1]).astype(np.float32)
    ny = min(rows, img.shape[1] - 1)
    gy[:, :ny, :] = (img[:, 1:ny + 1, :] - img[:, :ny, :]).astype(np.float32)
    # rows/columns the reference never reads (x == W-1 or y == H-1) stay zero, as in forward_gradients
    gx[:, :, -1] = 0
    gy[:, :, -1] = 0
    if y1 == H:
        gx[:, -1, :] = 0
        gy[:, -1, :] = 0
    gx = np.where(np.abs(gx) > 20, 0.0, gx).astype(np.float32)
    pin = synth_image(W, H, channels, seed=seed, y0=0, y1=1)[:, 0, 0].astype(np.float64)
    return gx, gy, ya, pin


# ---- Dirichlet-masked 5-point blend (C2/C3 converged-parity systems) ---------------------------------
def blob_mask(W, H, coverage=0.30, max_thickness=48, seed=11):
    """Union of random axis-aligned ellipses with semi-axes <= max_thickness/2, never touching the frame."""
    rng = np.random.default_rng(seed)
    mask = np.zeros((H, W), bool)
    target = coverage * W * H
    r_max = max(2, max_thickness // 2)
    guard = 0
    covered = 0  # == mask.sum(), kept incrementally (a full count per ellipse is minutes at 4096^2)
    while covered < target and guard < 200000:
        guard += 1
        ry, rx = rng.integers(max(2, r_max // 3), r_max + 1, 2)
        if H - 2 * ry - 2 <= 1 or W - 2 * rx - 2 <= 1:
            break
        cy = rng.integers(ry + 1, H - ry - 1)
        cx = rng.integers(rx + 1, W - rx - 1)
        yy, xx = np.ogrid[-ry:ry + 1, -rx:rx + 1]
        e = (yy / ry) ** 2 + (xx / rx) ** 2 <= 1.0
        win = mask[cy - ry:cy + ry + 1, cx - rx:cx + rx + 1]
        covered += int((e & ~win).sum())
        win |= e
    mask[0, :] = mask[-1, :] = False
    mask[:, 0] = mask[:, -1] = False
    return mask


def masked_poisson_system(mask, guide, target):
    """5-point Dirichlet system over the pixels with mask==1:
       4 v_p - sum_{q in N(p) & mask} v_q = sum_{q in N(p)} (g_p - g_q) + sum_{q in N(p) \\ mask} t_q
    guide/target: (C, H, W) arrays (uint8 or float).  Returns CSR (row_off, col_idx, values), b (C, n),
    pixel index of every unknown, and the parity colouring (x+y)&1 of the unknowns."""
    H, W = mask.shape
    g = guide.astype(np.float64)
    t = target.astype(np.float64)
    C = g.shape[0]
    idx = -np.ones((H, W), np.int64)
    ys, xs = np.nonzero(mask)
    n = ys.size
    idx[ys, xs] = np.arange(n)
    b = np.zeros((C, n))
    ent_r, ent_c, ent_v = [], [], []
    # neighbour order up, left, (diag), right, down == ascending compact index
    for dy, dx in ((-1, 0), (0, -1), (0, 1), (1, 0)):
        qy, qx = ys + dy, xs + dx
        inside = mask[qy, qx]
        b += g[:, ys, xs] - g[:, qy, qx]
        b += np.where(inside, 0.0, t[:, qy, qx])
        ent_r.append(np.arange(n)[inside])
        ent_c.append(idx[qy, qx][inside])
        ent_v.append(np.full(inside.sum(), -1.0))
    ent_r.append(np.arange(n))
    ent_c.append(np.arange(n))
    ent_v.append(np.full(n, 4.0))
    r = np.concatenate(ent_r)
    c = np.concatenate(ent_c)
    v = np.concatenate(ent_v)
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    ro, ci, va = coo_to_csr(r, c, v, n)
    colors = ((ys + xs) & 1).astype(np.int32)
    return ro, ci, va, b, (ys * W + xs), colors


def c3_masked_system(size=4096, channels=3):
    """BASELINE configs[2]'s converged-parity system: the Dirichlet-masked blend on a size x size frame,
    blob mask 30 %, thickness <= 48 px, guide = the synthetic two-exposure image, target = the same scene rotated.
    One definition for the golden generator (tests/golden/make_golden_c3.py), the tests and both bench arms."""
    mask = blob_mask(size, size, 0.30, 48, seed=11)
    guide = synth_image(size, size, channels, seed=7)
    target = np.ascontiguousarray(guide[:, ::-1, ::-1])
    return masked_poisson_system(mask, guide, target)


# ---- C5: random sparse SPD -----------------------------------------------------------------------
def random_spd_system(n=1_000_000, pairs_per_row=13, seed=5):
    """~2*pairs_per_row+1 nnz/row: symmetric random pattern, off-diagonals U(-1,0), diagonal = sum|off| + 1.
    Returns CSR and b = A x* with x* ~ U(-1,1)."""
    rng = np.random.default_rng(seed)
    i = np.repeat(np.arange(n, dtype=np.int64), pairs_per_row)
    j = rng.integers(0, n, i.size)
    keep = i != j
    i, j = i[keep], j[keep]
    v = rng.uniform(-1.0, 0.0, i.size)
    r = np.concatenate([i, j])
    c = np.concatenate([j, i])
    v = np.concatenate([v, v])
    key = r * n + c
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    first = np.ones(key.size, bool)
    first[1:] = key[1:] != key[:-1]
    # merge duplicates by summing
    seg = np.cumsum(first) - 1
    vs = np.zeros(seg[-1] + 1)
    np.add.at(vs, seg, v)
    key = key[first]
    r, c = key // n, key % n
    diag = np.zeros(n)
    np.add.at(diag, r, np.abs(vs))
    diag += 1.0
    r = np.concatenate([r, np.arange(n)])
    c = np.concatenate([c, np.arange(n)])
    vs = np.concatenate([vs, diag])
    order = np.lexsort((c, r))
    r, c, vs = r[order], c[order], vs[order]
    ro, ci, va = coo_to_csr(r, c, vs, n)
    xstar = rng.uniform(-1.0, 1.0, n)
    b = np.zeros(n)
    np.add.at(b, r, vs * xstar[c])
    return ro, ci, va, b, xstar


def random_spd_coo(n=10_000_000, pairs_per_row=13, seed=5):
    """C5 at full size, cheap to generate: UNSORTED COO triplets (the input of initializeFromTriplets, whose device
    radix sort does the ordering).  Each row draws `pairs_per_row` random partners; pairs are stored once per
    direction with the same value U(-1, 0) (orientation canonicalised, so that the last-duplicate-wins rule of the
    triplet build keeps the matrix symmetric); diagonal = (entries in the row) + 1 >= sum|off| + 1 (strictly
    dominant => SPD).  ~2 * pairs_per_row + 1 stored entries per row.  Returns rows, cols, vals (int32, int32, f64)."""
    rng = np.random.default_rng(seed)
    i = np.repeat(np.arange(n, dtype=np.int32), pairs_per_row)
    j = rng.integers(0, n, i.size, dtype=np.int32)
    keep = i != j
    i, j = i[keep], j[keep]
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    v = rng.uniform(-1.0, 0.0, lo.size)
    cnt = np.bincount(lo, minlength=n) + np.bincount(hi, minlength=n)
    d = np.arange(n, dtype=np.int32)
    rows = np.concatenate([lo, hi, d])
    cols = np.concatenate([hi, lo, d])
    vals = np.concatenate([v, v, cnt.astype(np.float64) + 1.0])
    return rows, cols, vals


# ---- numpy restatement of the reference Poisson stencil (host-side reference for generators) ----------
def poisson_nnz(W, H):
    return (W * H - 1) + 4 * (W - 1) * (H - 1)


def strip_bounds(H, world):
    """Row-strip partition of an H-row image over `world` ranks: [y0, y1) per rank, as even as possible."""
    base, rem = divmod(H, world)
    out, y = [], 0
    for r in range(world):
        h = base + (1 if r < rem else 0)
        out.append((y, y + h))
        y += h
    return out
