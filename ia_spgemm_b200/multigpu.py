"""Row-block multi-GPU plumbing (one process per GPU, torch.distributed).

The reference is single-device; this is the new functionality BASELINE.json's
north_star asks for: A is partitioned into contiguous row blocks balanced by intermediate products,
B is broadcast once (NCCL over NVLink on GPUs, gloo in the CPU tests), every rank produces its row
block of C, and only a handful of scalars are reduced afterwards.  No collective on the data path.
"""
import numpy as np


def per_row_products(rp_a, ci_a, rp_b):
    """ub[i] = sum of B row lengths over A(i,:)  (the row's share of GetFlop, csr:290-304)."""
    rp_a = np.asarray(rp_a, dtype=np.int64)
    lens_b = np.diff(np.asarray(rp_b, dtype=np.int64))
    if len(ci_a) == 0:
        return np.zeros(len(rp_a) - 1, dtype=np.int64)
    csum = np.concatenate(([0], np.cumsum(lens_b[np.asarray(ci_a)], dtype=np.int64)))
    return csum[rp_a[1:]] - csum[rp_a[:-1]]


def balanced_row_blocks(products_per_row, parts):
    """Contiguous row blocks with (nearly) equal shares of products: bounds[0..parts].
    Same rule as ias_partition_rows (csrc/csr_api.cu k_split): bounds[p] is the first row whose
    inclusive prefix exceeds total*p/parts."""
    incl = np.cumsum(np.asarray(products_per_row, dtype=np.int64))
    n = len(incl)
    total = int(incl[-1]) if n else 0
    bounds = [0]
    for p in range(1, parts):
        target = int(float(total) * p / parts)
        bounds.append(int(np.searchsorted(incl, target, side="right")))
    bounds.append(n)
    return bounds


def snake_row_share(products_per_row, parts, part):
    """Rows of one rank for strong scaling: the rows sorted by decreasing products (stable: ties keep index order) are
    dealt to the ranks in snake order -- round k gives position k*parts + part (k even) or k*parts + parts-1-part
    (k odd) -- so every rank gets its share of hub rows and of tail rows.  Same rule as ias_row_share
    (csrc/csr_api.cu: radix sort by work, k_take_share); returns row indices in the order they are dealt."""
    work = np.asarray(products_per_row, dtype=np.int64)
    n = len(work)
    order = np.argsort(-work, kind="stable")
    k = np.arange((n + parts - 1) // parts, dtype=np.int64)
    pos = np.where(k % 2 == 0, k * parts + part, k * parts + (parts - 1 - part))
    return order[pos[pos < n]].astype(np.int32)


def broadcast_csr(dist, rows, cols, rp, ci, v, src=0, device="cpu"):
    """Broadcast a CSR operand from `src` to every rank.  On `src` pass the torch tensors; elsewhere pass
    None.  Returns (rows, cols, rp, ci, v) as torch tensors on `device`."""
    import torch
    rank = dist.get_rank()
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == src:
        meta = torch.tensor([rows, cols, int(ci.numel())], dtype=torch.int64, device=device)
    dist.broadcast(meta, src)
    rows, cols, nnz = (int(x) for x in meta.tolist())
    if rank != src:
        rp = torch.empty(rows + 1, dtype=torch.int32, device=device)
        ci = torch.empty(nnz, dtype=torch.int32, device=device)
        v = torch.empty(nnz, dtype=torch.float64, device=device)
    for t in (rp, ci, v):
        dist.broadcast(t, src)
    return rows, cols, rp, ci, v


def reduce_scalars(dist, ints=(), floats=(), max_floats=(), device="cpu"):
    """Sum integer / float scalars and max float scalars over ranks (the only post-multiply traffic)."""
    import torch
    out_i, out_f, out_m = list(ints), list(floats), list(max_floats)
    if ints:
        t = torch.tensor(list(ints), dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out_i = [int(x) for x in t.tolist()]
    if floats:
        t = torch.tensor(list(floats), dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out_f = [float(x) for x in t.tolist()]
    if max_floats:
        t = torch.tensor(list(max_floats), dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out_m = [float(x) for x in t.tolist()]
    return out_i, out_f, out_m
