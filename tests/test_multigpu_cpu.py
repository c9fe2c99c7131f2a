"""World-size-2 test of the row-block sharding logic on CPU (gloo): rank 0 owns the operand, broadcasts
it, each rank multiplies its products-balanced row block (with the CPU oracle standing in for the
engine, which needs a GPU), and the concatenated blocks must equal the full product."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ia_spgemm_b200 import multigpu as M
    from ia_spgemm_b200 import workloads as W
    from oracle.binding import Oracle
    ora = Oracle()
    if rank == 0:
        rows, cols, rp, ci, v = W.rmat(9, 8, seed=4)
        args = (rows, cols, torch.from_numpy(rp.copy()), torch.from_numpy(ci.copy()), torch.from_numpy(v.copy()))
    else:
        args = (0, 0, None, None, None)
    rows, cols, rp, ci, v = M.broadcast_csr(dist, *args, src=0)
    rp, ci, v = rp.numpy(), ci.numpy(), v.numpy()
    ub = M.per_row_products(rp, ci, rp)
    bounds = M.balanced_row_blocks(ub, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    s, e = int(rp[r0]), int(rp[r1])
    blk_rp = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
    c_rp, c_ci, c_v = ora.csr_mul_csr(r1 - r0, cols, blk_rp, ci[s:e], v[s:e], rp, ci, v)
    (nnz, products), (checksum,), (tmax,) = M.reduce_scalars(dist, ints=(int(c_rp[-1]), int(ub[r0:r1].sum())),
                                                              floats=(float(c_v.sum()),), max_floats=(float(rank + 1),))
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), rp=c_rp, ci=c_ci, v=c_v, bounds=np.array(bounds), nnz=nnz, products=products,
             checksum=checksum, tmax=tmax, ub_block=int(ub[r0:r1].sum()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_row_blocks_match_full_product(tmp_path, oracle):
    from ia_spgemm_b200 import workloads as W
    from util import sort_rows
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows, cols, rp, ci, v = W.rmat(9, 8, seed=4)
    f_rp, f_ci, f_v = sort_rows(*oracle.csr_mul_csr(rows, cols, rp, ci, v, rp, ci, v))
    parts = [np.load(os.path.join(tmp_path, "r%d.npz" % r)) for r in range(world)]
    bounds = parts[0]["bounds"].tolist()
    assert bounds == parts[1]["bounds"].tolist() and bounds[0] == 0 and bounds[-1] == rows
    off = 0
    for r, p in enumerate(parts):
        b_rp, b_ci, b_v = sort_rows(p["rp"], p["ci"], p["v"])
        r0, r1 = bounds[r], bounds[r + 1]
        assert np.array_equal(b_rp, f_rp[r0:r1 + 1] - f_rp[r0])
        n = int(b_rp[-1])
        assert np.array_equal(b_ci, f_ci[off:off + n])
        assert np.allclose(b_v, f_v[off:off + n], rtol=1e-13, atol=0)
        off += n
    assert off == int(f_rp[-1]) == int(parts[0]["nnz"]) == int(parts[1]["nnz"])          # all-reduced totals agree
    assert int(parts[0]["products"]) == oracle.getflop(rp, ci, rp)
    assert np.isclose(float(parts[0]["checksum"]), f_v.sum(), rtol=1e-12)
    assert float(parts[0]["tmax"]) == 2.0
    # products-balanced: the un-permuted R-MAT puts the heavy rows first, so an equal-rows split would be lopsided
    shares = [int(p["ub_block"]) for p in parts]
    assert max(shares) < 0.6 * sum(shares)
    assert bounds[1] < rows // 2


def _strong_worker(rank, world, port, out_dir):
    """Strong scaling as bench.py's rmat22_strong leg does it: B broadcast once, the rows dealt in snake order by
    decreasing products, every rank multiplies its (non-contiguous) row list; a few scalars are reduced."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ia_spgemm_b200 import multigpu as M
    from ia_spgemm_b200 import workloads as W
    from oracle.binding import Oracle
    ora = Oracle()
    if rank == 0:
        rows, cols, rp, ci, v = W.rmat(9, 8, seed=5)
        args = (rows, cols, torch.from_numpy(rp.copy()), torch.from_numpy(ci.copy()), torch.from_numpy(v.copy()))
    else:
        args = (0, 0, None, None, None)
    rows, cols, rp, ci, v = M.broadcast_csr(dist, *args, src=0)
    rp, ci, v = rp.numpy(), ci.numpy(), v.numpy()
    ub = M.per_row_products(rp, ci, rp)
    mine = M.snake_row_share(ub, world, rank)
    lens = (rp[mine + 1] - rp[mine]).astype(np.int64)
    g_rp = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)                  # the gathered rows as their own CSR block
    idx = np.concatenate([np.arange(rp[r], rp[r + 1]) for r in mine]) if len(mine) else np.zeros(0, dtype=np.int64)
    c_rp, c_ci, c_v = ora.csr_mul_csr(len(mine), cols, g_rp, ci[idx], v[idx], rp, ci, v)
    (nnz, products), (checksum,), _ = M.reduce_scalars(dist, ints=(int(c_rp[-1]), int(ub[mine].sum())), floats=(float(c_v.sum()),))
    np.savez(os.path.join(out_dir, "s%d.npz" % rank), rows=mine, rp=c_rp, ci=c_ci, v=c_v, nnz=nnz, products=products, checksum=checksum,
             share=int(ub[mine].sum()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_snake_shares_match_full_product(tmp_path, oracle):
    from ia_spgemm_b200 import workloads as W
    from util import sort_rows
    world, port = 2, 31000 + os.getpid() % 2000
    mp.spawn(_strong_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows, cols, rp, ci, v = W.rmat(9, 8, seed=5)
    f_rp, f_ci, f_v = sort_rows(*oracle.csr_mul_csr(rows, cols, rp, ci, v, rp, ci, v))
    parts = [np.load(os.path.join(tmp_path, "s%d.npz" % r)) for r in range(world)]
    seen = np.concatenate([p["rows"] for p in parts])
    assert len(seen) == rows and np.array_equal(np.sort(seen), np.arange(rows))      # every row exactly once
    for p in parts:
        b_rp, b_ci, b_v = sort_rows(p["rp"], p["ci"], p["v"])
        for k, r in enumerate(p["rows"]):
            s, e = int(b_rp[k]), int(b_rp[k + 1])
            fs, fe = int(f_rp[r]), int(f_rp[r + 1])
            assert e - s == fe - fs
            assert np.array_equal(b_ci[s:e], f_ci[fs:fe])
            assert np.allclose(b_v[s:e], f_v[fs:fe], rtol=1e-13, atol=0)
    assert int(parts[0]["nnz"]) == int(parts[1]["nnz"]) == int(f_rp[-1])
    assert int(parts[0]["products"]) == oracle.getflop(rp, ci, rp)
    assert np.isclose(float(parts[0]["checksum"]), f_v.sum(), rtol=1e-12)
    shares = [int(p["share"]) for p in parts]
    assert max(shares) <= 0.51 * sum(shares)                                          # snake dealing balances the hubs


def test_snake_row_share_properties():
    from ia_spgemm_b200 import multigpu as M
    rng = np.random.default_rng(1)
    work = rng.integers(0, 1000, size=1003)
    work[:7] = 10 ** 6                            # hubs, with ties among them
    for parts in (1, 2, 3, 8):
        shares = [M.snake_row_share(work, parts, p) for p in range(parts)]
        allrows = np.concatenate(shares)
        assert np.array_equal(np.sort(allrows), np.arange(len(work)))
        assert max(len(s) for s in shares) - min(len(s) for s in shares) <= 1
        for s in shares:
            assert np.all(np.diff(work[s]) <= 0)                                     # dealt in order of decreasing work
        tot = [int(work[s].sum()) for s in shares]
        assert max(tot) - min(tot) <= 10 ** 6                                         # within one hub of each other
    # ties keep index order (what a stable radix sort does): equal work -> rank 0 gets row 0, rank 1 row 1, then snake back
    eq = np.full(6, 5)
    assert M.snake_row_share(eq, 2, 0).tolist() == [0, 3, 4] and M.snake_row_share(eq, 2, 1).tolist() == [1, 2, 5]
    assert M.snake_row_share(np.zeros(0, dtype=np.int64), 4, 2).tolist() == []


def test_balanced_row_blocks_properties():
    from ia_spgemm_b200 import multigpu as M
    rng = np.random.default_rng(0)
    ub = rng.integers(0, 1000, size=5000)
    ub[:10] = 100000                             # heavy head, like R-MAT
    for parts in (1, 2, 3, 8):
        b = M.balanced_row_blocks(ub, parts)
        assert len(b) == parts + 1 and b[0] == 0 and b[-1] == len(ub) and all(x <= y for x, y in zip(b, b[1:]))
        shares = [int(ub[x:y].sum()) for x, y in zip(b, b[1:])]
        assert sum(shares) == int(ub.sum())
        assert max(shares) <= ub.sum() / parts + ub.max()
    assert M.balanced_row_blocks(np.zeros(0, dtype=np.int64), 4) == [0, 0, 0, 0, 0]
    assert M.balanced_row_blocks(np.zeros(7, dtype=np.int64), 2)[-1] == 7


def test_per_row_products_matches_oracle(oracle):
    from ia_spgemm_b200 import multigpu as M
    from ia_spgemm_b200 import workloads as W
    rows, cols, rp, ci, v = W.random_sparse(60, 60, 0.1, seed=3)
    ub = M.per_row_products(rp, ci, rp)
    assert int(ub.sum()) == oracle.getflop(rp, ci, rp)
    assert ub[np.diff(rp) == 0].sum() == 0
