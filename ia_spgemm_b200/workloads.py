"""Synthetic operands for the BASELINE.json configs (host side, NumPy).

All generators are deterministic functions of (shape, seed) built on a counter-based hash, so the
CPU oracle, the GPU engine and the device-side generators (csrc/generate.cu, same hash) see
bit-identical arrays.  Matrices are canonical CSR: int32 row pointers / sorted, duplicate-free
column indices / float64 values.  Values are U[0.5, 1.5) of hash(seed, i, j) -- zero-free and
cancellation-free, so the structure of A*A is value independent -- except
Poisson, which carries the usual 4 / -1 stencil.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def _key(seed, i, j):
    with np.errstate(over="ignore"):
        s = mix64(np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019))
        return (np.asarray(i, dtype=np.uint64) << np.uint64(32) | np.asarray(j, dtype=np.uint64)) ^ s


def hash_values(seed, rows_of_entries, cols):
    """U[0.5,1.5) value for each (i, j)."""
    h = mix64(_key(seed, rows_of_entries, cols))
    return 0.5 + (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def row_index(row_ptr):
    n = len(row_ptr) - 1
    return np.repeat(np.arange(n, dtype=np.int64), np.diff(row_ptr).astype(np.int64))


def poisson2d(nx, ny=None):
    """2-D 5-point Poisson operator on an nx x ny grid (ny defaults to nx), row-major node order."""
    ny = nx if ny is None else ny
    n = nx * ny
    r = np.arange(n, dtype=np.int64)
    x = r % nx
    y = r // nx
    cand = np.stack([r - nx, r - 1, r, r + 1, r + nx], axis=1)
    mask = np.stack([y > 0, x > 0, np.ones(n, bool), x < nx - 1, y < ny - 1], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), (n, 5))
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=row_ptr[1:])
    return n, n, row_ptr.astype(np.int32), cand[mask].astype(np.int32), np.ascontiguousarray(vals[mask])


def uniform_rows(n, per_row, seed=1, chunk=1 << 20):
    """Every row has `per_row` distinct, sorted columns drawn uniformly from [0, n)."""
    col = np.empty((n, per_row), dtype=np.int32)
    t = np.arange(per_row, dtype=np.uint64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        i = np.arange(s, e, dtype=np.uint64)[:, None]
        draw = 0
        c = (mix64(_key(seed + 7919, i, t[None, :])) % np.uint64(n)).astype(np.int64)
        c.sort(axis=1)
        bad = np.nonzero((np.diff(c, axis=1) == 0).any(axis=1))[0]
        while len(bad):          # re-draw the few rows that collided, with a fresh counter block
            draw += 1
            ib = i[bad]
            cb = (mix64(_key(seed + 7919, ib, t[None, :] + np.uint64(draw * per_row))) % np.uint64(n)).astype(np.int64)
            cb.sort(axis=1)
            c[bad] = cb
            bad = bad[(np.diff(cb, axis=1) == 0).any(axis=1)]
        col[s:e] = c
    row_ptr = (np.arange(n + 1, dtype=np.int64) * per_row).astype(np.int32)
    ci = col.reshape(-1)
    vals = hash_values(seed, row_index(row_ptr), ci)
    return n, n, row_ptr, ci, vals


def rmat(scale, edge_factor=16, seed=1, a=0.57, b=0.19, c=0.19, chunk=1 << 24):
    """Graph500-style R-MAT: edge_factor * 2^scale directed draws, duplicates removed,
    self loops kept, no vertex permutation."""
    n = 1 << scale
    m = edge_factor * n
    ab, abc = a + b, a + b + c
    keys = []
    for s in range(0, m, chunk):
        e = np.arange(s, min(m, s + chunk), dtype=np.uint64)
        i = np.zeros(len(e), dtype=np.uint64)
        j = np.zeros(len(e), dtype=np.uint64)
        for lvl in range(scale):
            u = (mix64(_key(seed + 104729, e, lvl)) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
            ibit = (u >= ab).astype(np.uint64)
            jbit = (((u >= a) & (u < ab)) | (u >= abc)).astype(np.uint64)
            i = (i << np.uint64(1)) | ibit
            j = (j << np.uint64(1)) | jbit
        keys.append((i << np.uint64(32)) | j)
    k = np.unique(np.concatenate(keys))
    ri = (k >> np.uint64(32)).astype(np.int64)
    ci = (k & np.uint64(0xFFFFFFFF)).astype(np.int32)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(ri, minlength=n), out=row_ptr[1:])
    vals = hash_values(seed, ri, ci)
    return n, n, row_ptr.astype(np.int32), ci, vals


def banded(n, offsets, seed=1):
    """All diagonals in `offsets` fully populated (DIA-friendly test operand)."""
    offsets = np.array(sorted(offsets), dtype=np.int64)
    r = np.arange(n, dtype=np.int64)[:, None]
    cand = r + offsets[None, :]
    mask = (cand >= 0) & (cand < n)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=row_ptr[1:])
    ci = cand[mask].astype(np.int32)
    vals = hash_values(seed, row_index(row_ptr), ci)
    return n, n, row_ptr.astype(np.int32), ci, vals


def random_sparse(rows, cols, density, seed=1, sort_columns=True):
    """Small irregular test operand with empty rows allowed; optionally unsorted columns."""
    rng = np.random.default_rng(seed)
    mask = rng.random((rows, cols)) < density
    row_ptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=row_ptr[1:])
    ri, ci = np.nonzero(mask)
    if not sort_columns:
        # shuffle within rows, keeping the row grouping
        perm = np.lexsort((rng.random(len(ci)), ri))
        ci = ci[perm]
    vals = hash_values(seed, ri, ci)
    return rows, cols, row_ptr.astype(np.int32), ci.astype(np.int32), vals


def poisson_counts(n_grid):
    """Closed forms for the Poisson config: nnz(A), products, nnz(C)."""
    N = n_grid
    return 5 * N * N - 4 * N, 25 * (N - 2) ** 2 + 64 * (N - 2) + 36, 13 * N * N - 20 * N + 4
