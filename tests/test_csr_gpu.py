"""Parity of the CUDA CSR hot path (through the C ABI) with the CPU oracle.

Bar (BASELINE.json north_star): row_ptr and column indices bit-exact after sorting, values within
1e-12 relative per entry.  Oracle rows come out unsorted (reverse first-touch), the engine's are
column sorted, so the oracle side is sorted before comparison (tests/util.py).
"""
import os

import numpy as np
import pytest

from ia_spgemm_b200 import workloads as W
from util import RTOL, SQUARE, abs_product, assert_csr_parity, sort_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ia_spgemm_b200.engine import get_engine
    return get_engine()


def _mul(eng, A, B=None):
    dA = eng.upload(*A)
    dB = dA if B is None else eng.upload(*B)
    got, st = eng.CSR_MUL_CSR_DEV(dA, dB)
    dA.close()
    if B is not None:
        dB.close()
    return got, st


def _oracle(oracle, A, B=None):
    B = A if B is None else B
    return oracle.csr_mul_csr(A[0], B[1], A[2], A[3], A[4], B[2], B[3], B[4])


def _check(eng, oracle, A, B=None, mag=True):
    got, st = _mul(eng, A, B)
    want = _oracle(oracle, A, B)
    m = abs_product(oracle, A, A if B is None else B) if mag else None
    assert_csr_parity(got, want, mag=m)
    assert st["nnz"] == int(want[0][-1])
    assert st["products"] == oracle.getflop(A[2], A[3], (A if B is None else B)[2])
    return got, st


@pytest.mark.parametrize("name", SQUARE)
def test_bundled_square_inputs(eng, oracle, golden, mtx_dir, name):
    """A^2 of the reference's own Inputs/*.mtx, against the oracle and the committed reference output."""
    A = oracle.mtx_load(os.path.join(mtx_dir, name + ".mtx"))
    got, st = _check(eng, oracle, A)
    g = golden["inputs"][name]
    if "csr_row_ptr" not in g:
        return
    assert st["nnz"] == g["csr_row_ptr"][-1]
    assert st["products"] == g["flop"]
    ref = sort_rows(np.array(g["csr_row_ptr"]), np.array(g["csr_col_ind"]), np.array(g["csr_values"], dtype=np.float64))
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    scale = np.maximum(np.abs(ref[2]), abs_product(oracle, A, A))
    assert np.all(np.abs(got[2] - ref[2]) <= RTOL * scale)
    assert eng.lib.ias_sizeof_csr(A[0], st["nnz"]) == g["sizeof_csr_c"]


def test_dia_mtx_known_answer(eng, oracle, mtx_dir):
    """dia.mtx A^2 = rows {0:1,1:2,2:1} {1:1,2:2,3:1} {2:1,3:2} {3:1}."""
    A = oracle.mtx_load(os.path.join(mtx_dir, "dia.mtx"))
    (rp, ci, v), st = _mul(eng, A)
    assert rp.tolist() == [0, 3, 6, 8, 9]
    assert ci.tolist() == [0, 1, 2, 1, 2, 3, 2, 3, 3]
    assert v.tolist() == [1, 2, 1, 1, 2, 1, 1, 2, 1]
    assert st["products"] == 12


def test_keeps_numerical_zeros(eng, oracle, mtx_dir):
    """b1_ss has three entries that cancel to exactly 0.0; the structure must keep them."""
    A = oracle.mtx_load(os.path.join(mtx_dir, "b1_ss.mtx"))
    (rp, ci, v), _ = _mul(eng, A)
    assert int(rp[-1]) == 30
    assert np.count_nonzero(np.abs(v) < 1e-15) == 3


@pytest.mark.parametrize("rows,inner,cols,density,seed", [
    (1, 1, 1, 1.0, 1), (7, 5, 9, 0.5, 2), (64, 64, 64, 0.1, 3), (300, 200, 500, 0.05, 4),
    (1000, 1000, 1000, 0.02, 5), (513, 2000, 3000, 0.3, 6), (50, 4000, 20000, 0.6, 7),
])
def test_random_rectangular_unsorted(eng, oracle, rows, inner, cols, density, seed):
    """A*B with empty rows, unsorted column order inside rows, every smem bin."""
    A = W.random_sparse(rows, inner, density, seed=seed, sort_columns=False)
    B = W.random_sparse(inner, cols, density, seed=seed + 100, sort_columns=False)
    _check(eng, oracle, A, B)


def test_empty_operands(eng, oracle):
    z = (5, 5, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0))
    (rp, ci, v), st = _mul(eng, z)
    assert rp.tolist() == [0] * 6 and len(ci) == 0 and st["products"] == 0
    A = W.random_sparse(5, 5, 0.5, seed=1)
    (rp, ci, v), st = _mul(eng, A, z)
    assert int(rp[-1]) == 0
    z0 = (0, 4, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    B = W.random_sparse(4, 4, 0.5, seed=1)
    (rp, ci, v), st = _mul(eng, z0, B)
    assert rp.tolist() == [0]


def test_duplicate_entries_are_merged_like_csr_mul_csr(eng, oracle):
    """The reference loader does not merge duplicate (i,j); CSR_MUL_CSR accumulates them."""
    rp = np.array([0, 3, 5], np.int32)
    ci = np.array([1, 1, 0, 0, 0], np.int32)
    v = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    _check(eng, oracle, (2, 2, rp, ci, v))


def test_shape_mismatch_is_an_error(eng):
    from ia_spgemm_b200.engine import EngineError
    A = W.random_sparse(4, 6, 0.5, seed=1)
    B = W.random_sparse(5, 4, 0.5, seed=2)
    dA, dB = eng.upload(*A), eng.upload(*B)
    with pytest.raises(EngineError) as e:
        eng.CSR_MUL_CSR_DEV(dA, dB)
    assert e.value.code == 2


def test_poisson_tiny_bin(eng, oracle):
    A = W.poisson2d(192)
    got, st = _check(eng, oracle, A, mag=True)
    nnz_a, prod, nnz_c = W.poisson_counts(192)
    assert (st["products"], st["nnz"]) == (prod, nnz_c)
    assert st["sym_bin_rows"][1] == 192 * 192           # every row in the tiny bin


def test_uniform_warp_bin(eng, oracle):
    A = W.uniform_rows(30000, 16, seed=1)
    got, st = _check(eng, oracle, A, mag=False)
    assert st["products"] == 30000 * 256
    assert st["sym_bin_rows"][2] == 30000


@pytest.mark.parametrize("scale,takes_b2", [(10, 1), (13, 1), (13, 0)])
def test_rmat_all_bins(eng, oracle, scale, takes_b2):
    A = W.rmat(scale, 16, seed=1)
    saved = eng.get_option("gwin_takes_b2")
    eng.set_option("gwin_takes_b2", takes_b2)
    try:
        got, st = _check(eng, oracle, A, mag=False)
    finally:
        eng.set_option("gwin_takes_b2", saved)
    if scale == 13:
        assert st["sym_bin_rows"][5] > 0
        if takes_b2:      # one super-window: rows beyond the small CTA hash go to the windowed kernel
            assert st["num_bin_rows"][4] == 0 and st["num_bin_rows"][5] > 0
        else:             # global + large CTA bins exercised
            assert st["num_bin_rows"][4] > 0


def test_rmat_forced_global_bin(eng, oracle):
    """A dense-ish operand whose rows exceed the shared-memory hash: bitmap + rank path."""
    A = W.random_sparse(40, 300, 0.9, seed=3)
    B = W.random_sparse(300, 40000, 0.4, seed=4, sort_columns=False)
    got, st = _check(eng, oracle, A, B, mag=False)
    assert st["num_bin_rows"][5] > 0


def test_row_blocks_concatenate_to_full_result(eng, oracle):
    """Multi-GPU row blocks: C[r0:r1,:] computed independently equals the slice of the full product."""
    A = W.rmat(11, 16, seed=2)
    dA = eng.upload(*A)
    (rp, ci, v), st = eng.CSR_MUL_CSR_DEV(dA, dA)
    bounds = eng.partition_rows(dA, dA, 4)
    assert bounds[0] == 0 and bounds[-1] == A[0] and all(b0 <= b1 for b0, b1 in zip(bounds, bounds[1:]))
    tot = 0
    for b0, b1 in zip(bounds, bounds[1:]):
        (rpb, cib, vb), sb = eng.CSR_MUL_CSR_DEV(dA, dA, rows=(b0, b1))
        assert np.array_equal(rpb, rp[b0:b1 + 1] - rp[b0])
        assert np.array_equal(cib, ci[rp[b0]:rp[b1]])
        assert np.allclose(vb, v[rp[b0]:rp[b1]], rtol=1e-13, atol=0)
        tot += sb["products"]
    assert tot == st["products"] == eng.GetFlop(dA, dA)
    # balance: no block above 2x the mean share (contiguous split of a skewed matrix)
    shares = [eng.CSR_MUL_CSR_DEV(dA, dA, rows=(b0, b1), download=False)[1]["products"] for b0, b1 in zip(bounds, bounds[1:])]
    assert max(shares) <= 2.0 * st["products"] / 4 + max(A[2][1:] - A[2][:-1]) * 2000
    dA.close()


def test_streaming_matches_materialised(eng, oracle):
    """Streaming mode (row batches under a byte budget) reproduces nnz, checksum, structure hash, row nnz."""
    A = W.rmat(12, 16, seed=3)
    dA = eng.upload(*A)
    c64, st = eng.CSR_MUL_CSR_DEV(dA, dA, keep=True)
    h = eng.structure_hash(c64)
    s = eng.checksum_ptr(c64.values_dev, c64.nnz)
    (rp, ci, v) = eng._take_csr64(c64)
    for budget in (0, 12 * 40000, 12 * 5000):
        d = eng.csr_mul_csr_stream(dA, dA, budget_bytes=budget, want_row_nnz=True)
        assert d["nnz"] == st["nnz"] and d["products"] == st["products"]
        assert d["structure_hash"] == h
        assert np.isclose(d["checksum"], s, rtol=1e-11)
        assert np.array_equal(d["row_nnz"], np.diff(rp))
        if budget:
            assert d["batches"] > 1
    # host restatement of the hash: sum of mix64(row<<32|col)
    rows = np.repeat(np.arange(A[0], dtype=np.uint64), np.diff(rp))
    with np.errstate(over="ignore"):
        want = int(W.mix64((rows << np.uint64(32)) | ci.astype(np.uint64)).sum(dtype=np.uint64))
    assert h == want
    dA.close()


def test_stream_consumer_reassembles_c(eng, oracle):
    """The streaming entry with a consumer callback: the batches, concatenated on the host, are C."""
    A = W.rmat(11, 16, seed=5)
    dA = eng.upload(*A)
    (rp, ci, v), st = eng.CSR_MUL_CSR_DEV(dA, dA)
    parts = []

    def consumer(b):
        assert b["nnz_total"] == st["nnz"] and b["row_ptr"][0] == 0
        parts.append(b)

    d = eng.csr_mul_csr_stream(dA, dA, budget_bytes=12 * 30000, consumer=consumer)
    assert d["batches"] == len(parts) > 1 and [b["batch_index"] for b in parts] == list(range(len(parts)))
    assert parts[0]["row_begin"] == 0 and parts[-1]["row_end"] == A[0]
    assert all(x["row_end"] == y["row_begin"] for x, y in zip(parts, parts[1:]))
    assert np.array_equal(np.concatenate([b["col_ind"] for b in parts]), ci)
    # (the order of the shared-memory fp64 additions differs from run to run: values agree to rounding, not to the bit)
    assert np.allclose(np.concatenate([b["values"] for b in parts]), v, rtol=1e-13, atol=0)
    assert np.array_equal(np.concatenate([b["row_ptr"][:-1] + b["entry_base"] for b in parts] + [[st["nnz"]]]), rp)
    # a consumer that fails stops the multiply with its status
    from ia_spgemm_b200.engine import EngineError
    with pytest.raises(EngineError) as ei:
        eng.csr_mul_csr_stream(dA, dA, budget_bytes=12 * 30000, consumer=lambda b: 7)
    assert ei.value.code == 7
    dA.close()


def test_row_list_streaming_equals_rows_of_the_full_product(eng, oracle):
    """ias_csr_mul_csr_rowlist_stream (the multi-GPU entry for skewed operands): an arbitrary list of rows of A -- the cyclic
    share of a rank, a reversed range, duplicates -- gives exactly those rows of A*B, through every bin of the pipeline."""
    import torch
    A = W.rmat(12, 16, seed=9)
    dA = eng.upload(*A)
    (rp, ci, v), st = eng.CSR_MUL_CSR_DEV(dA, dA)
    n = A[0]
    for rows in (np.arange(1, n, 3), np.arange(n - 1, -1, -7), np.array([5, 5, 0, n - 1, 5]), np.arange(0, 0)):
        t = torch.from_numpy(rows.astype(np.int32)).cuda()
        parts = []
        d = eng.csr_mul_csr_rowlist_stream(dA, dA, t.data_ptr(), len(rows), budget_bytes=12 * 20000, want_row_nnz=True,
                                           consumer=lambda b: parts.append(b))
        want_nnz = np.diff(rp)[rows] if len(rows) else np.zeros(0, np.int64)
        assert np.array_equal(d["row_nnz"], want_nnz) and d["nnz"] == int(want_nnz.sum())
        if len(rows) == 0:
            continue
        got_ci = np.concatenate([b["col_ind"] for b in parts])
        got_v = np.concatenate([b["values"] for b in parts])
        want_ci = np.concatenate([ci[rp[r]:rp[r + 1]] for r in rows])
        want_v = np.concatenate([v[rp[r]:rp[r + 1]] for r in rows])
        assert np.array_equal(got_ci, want_ci) and np.allclose(got_v, want_v, rtol=1e-13, atol=0)
        assert parts[0]["row_begin"] == 0 and parts[-1]["row_end"] == len(rows)
    # work-balanced shares: a partition of the rows, products within a few percent of each other
    lens = np.diff(A[2]).astype(np.int64)
    cs = np.concatenate(([0], np.cumsum(lens[A[3]])))
    per_row = cs[A[2][1:]] - cs[A[2][:-1]]
    shares = [eng.row_share(dA, dA, 3, p).cpu().numpy() for p in range(3)]
    assert sorted(np.concatenate(shares).tolist()) == list(range(n))
    from ia_spgemm_b200.multigpu import snake_row_share            # the dealing rule restated in numpy (CPU multi-rank tests use it)
    for p in range(3):
        assert np.array_equal(shares[p], snake_row_share(per_row, 3, p))
    tot = [int(per_row[sh].sum()) for sh in shares]
    assert max(tot) - min(tot) <= 0.05 * sum(tot) / 3 + per_row.max()
    from ia_spgemm_b200.engine import EngineError
    bad = torch.tensor([0, n], dtype=torch.int32).cuda()
    with pytest.raises(EngineError):
        eng.csr_mul_csr_rowlist_stream(dA, dA, bad.data_ptr(), 2)
    dA.close()


def test_int32_layout_and_host_path(eng, oracle):
    A = W.random_sparse(200, 200, 0.05, seed=9)
    dA = eng.upload(*A)
    rp32, ci32, v32, ms = eng.csr_mul_csr_dev32(dA, dA)
    want = _oracle(oracle, A)
    assert_csr_parity((rp32, ci32, v32), want, mag=abs_product(oracle, A, A))
    assert rp32.dtype == np.int32 and ms > 0
    (rp, ci, v), st, h2d, d2h = eng.CSR_MUL_CSR(A, A)
    assert_csr_parity((rp.copy(), ci.copy(), v.copy()), want, mag=abs_product(oracle, A, A))
    assert eng.is_canonical(dA)
    dA.close()


def test_launch_counter_moves(eng):
    A = W.poisson2d(32)
    dA = eng.upload(*A)
    before = eng.kernel_launches()
    _, st = eng.CSR_MUL_CSR_DEV(dA, dA, download=False)
    assert eng.kernel_launches() - before == st["kernel_launches"] > 0
    dA.close()


def test_canonical_flag_cache_is_invalidated(eng, oracle):
    """Opt-in cache ("trust_operand_cache"): row-block multiplies remember whether B is canonical; freeing or
    forgetting the operand drops that."""
    eng.set_option("trust_operand_cache", 1)
    try:
        for sort_columns in (True, False, True, False):        # the pool hands the same addresses out again
            A = W.random_sparse(300, 300, 0.03, seed=11, sort_columns=sort_columns)
            dA = eng.upload(*A)
            want = sort_rows(*_oracle(oracle, A))
            for _ in range(2):                                  # second call hits the cache
                (rp, ci, v), st = eng.CSR_MUL_CSR_DEV(dA, dA, rows=(100, 250))
                assert np.array_equal(rp, want[0][100:251] - want[0][100])
                assert np.array_equal(ci, want[1][want[0][100]:want[0][250]])
                assert np.allclose(v, want[2][want[0][100]:want[0][250]], rtol=1e-12, atol=1e-15)
            eng.forget_operand(dA)
            dA.close()
    finally:
        eng.set_option("trust_operand_cache", 0)


def test_caller_owned_operand_reallocated_at_the_same_address(eng, oracle):
    """a torch tensor freed and re-allocated at the same address with the same shape and nnz must not
    inherit the previous operand's canonical flag.  Default (cache off) and wrap_device's forget both cover it."""
    import torch
    A1 = W.random_sparse(300, 300, 0.03, seed=11, sort_columns=True)
    perm_rows = np.repeat(np.arange(300), np.diff(A1[2]))
    rng = np.random.default_rng(5)
    order = np.lexsort((rng.random(len(A1[3])), perm_rows))        # same shape, same nnz, columns shuffled inside rows
    A2 = (A1[0], A1[1], A1[2], A1[3][order], A1[4][order])
    for trust in (0, 1):
        eng.set_option("trust_operand_cache", trust)
        ptrs = []
        for A in (A1, A2):
            t = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in A[2:]]
            ptrs.append(t[1].data_ptr())
            d = eng.wrap_device(A[0], A[1], len(A[3]), t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr())
            want = sort_rows(*_oracle(oracle, A))
            (rp, ci, v), st = eng.CSR_MUL_CSR_DEV(d, d, rows=(50, 250))
            assert np.array_equal(rp, want[0][50:251] - want[0][50])
            assert np.array_equal(ci, want[1][want[0][50]:want[0][250]])
            d.close()
            del t, d
            torch.cuda.synchronize()
        # (the caching allocator normally returns the same block: the test is meaningful when it does, harmless otherwise)
    eng.set_option("trust_operand_cache", 0)


def test_long_row_inside_a_small_row_block(eng, oracle):
    """a row block with few rows that contains an A row of more than 64 entries, in a matrix whose
    average row is short.  The long-row analysis kernel must run for it (it used to be gated on avg * block rows)."""
    n = 4000
    rng = np.random.default_rng(3)
    rows_ci = [np.sort(rng.choice(n, size=2, replace=False)) for _ in range(n)]
    for hub, ln in ((7, 65), (1234, 200), (3999, 900)):
        rows_ci[hub] = np.sort(rng.choice(n, size=ln, replace=False))
    rp = np.zeros(n + 1, np.int32)
    rp[1:] = np.cumsum([len(r) for r in rows_ci])
    ci = np.concatenate(rows_ci).astype(np.int32)
    v = W.hash_values(1, W.row_index(rp), ci)
    A = (n, n, rp, ci, v)
    dA = eng.upload(*A)
    for r0, r1 in ((7, 8), (1234, 1235), (1230, 1240), (3999, 4000), (0, 16)):
        s, e = int(rp[r0]), int(rp[r1])
        blk_rp = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
        want = oracle.csr_mul_csr(r1 - r0, n, blk_rp, ci[s:e], v[s:e], rp, ci, v)
        got, st = eng.CSR_MUL_CSR_DEV(dA, dA, rows=(r0, r1))
        assert_csr_parity(got, want)
        assert st["products"] == oracle.getflop(blk_rp, ci[s:e], rp)
        d = eng.csr_mul_csr_stream(dA, dA, rows=(r0, r1), want_row_nnz=True)
        assert d["nnz"] == st["nnz"] and np.array_equal(d["row_nnz"], np.diff(want[0]))
    bounds = eng.partition_rows(dA, dA, 64)                      # many parts: blocks of a handful of rows
    total = sum(eng.CSR_MUL_CSR_DEV(dA, dA, rows=(x, y), download=False)[1]["nnz"] for x, y in zip(bounds, bounds[1:]) if y > x)
    assert total == eng.CSR_MUL_CSR_DEV(dA, dA, download=False)[1]["nnz"]
    dA.close()


def test_touched_b_bytes(eng):
    """Algorithmic-bytes helper: 4*|T| + 12*sum len(B_j) over the columns referenced by an A row block."""
    A = W.rmat(10, 8, seed=7)
    dA = eng.upload(*A)
    rows, cols, rp, ci, v = A
    lens = np.diff(rp).astype(np.int64)
    for r0, r1 in ((0, rows), (100, 300), (5, 5)):
        T = np.unique(ci[rp[r0]:rp[r1]])
        assert eng.touched_b_bytes(dA, dA, rows=(r0, r1)) == int(4 * len(T) + 12 * lens[T].sum())
    dA.close()


def test_wide_column_space_uses_64bit_sort_keys(eng, oracle):
    """More than 2^23 columns: the register sort of the warp bin packs (column, index) into 64-bit keys."""
    rng = np.random.default_rng(5)
    ncols = (1 << 24) + 5
    A = W.random_sparse(200, 300, 0.03, seed=21)
    per = 40
    cols = np.sort(np.stack([rng.choice(ncols, size=per, replace=False) for _ in range(300)]), axis=1)
    cols[0, -1] = ncols - 1                                  # the very last column is used
    B = (300, ncols, (np.arange(301) * per).astype(np.int32), cols.reshape(-1).astype(np.int32),
         rng.uniform(0.5, 1.5, size=300 * per))
    got, st = _check(eng, oracle, A, B, mag=False)
    assert st["num_bin_rows"][2] > 0                         # warp bin populated
    assert int(got[1].max()) == ncols - 1


# ---------------------------------------------------------------- global rows: windowed shared-memory kernels
_GWIN_OPTS = ("global_rows_smem", "gwin_swords", "gwin_win", "gwin_sym_swords", "gwin_smem_kb", "gwin_max_sw", "g_win", "g_coop", "gwin_takes_b2",
              "g_v2", "g_tbl", "g_lpt", "g_block", "g_scr", "g_split", "g_split_ub", "g_split_parts", "g2_takes_b2")


@pytest.fixture
def options(eng):
    saved = {k: eng.get_option(k) for k in _GWIN_OPTS}
    yield eng.set_option
    for k, v in saved.items():
        eng.set_option(k, v)


def _global_operands(kind):
    if kind == "few_a_entries":        # <= 128 A entries per row: warp-wide lower bounds, tile reused by the accumulate pass
        return W.random_sparse(12, 100, 0.8, seed=11), W.random_sparse(100, 70001, 0.02, seed=12)
    if kind == "mid_a_entries":        # 129..1024 A entries: one lower bound per thread, single tile
        return W.random_sparse(10, 400, 0.7, seed=13), W.random_sparse(400, 50000, 0.005, seed=14)
    if kind == "long_a_rows":          # more than 1024 A entries: several A tiles per window
        return W.random_sparse(6, 3000, 0.9, seed=15), W.random_sparse(3000, 30011, 0.0007, seed=16)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["few_a_entries", "mid_a_entries", "long_a_rows"])
@pytest.mark.parametrize("swords,win,sym_swords", [(0, 0, 0), (32, 32, 32), (64, 96, 64), (8192, 64, 0), (32, 0, 0), (512, 4096, 1024)])
def test_global_rows_windowed_kernels(eng, oracle, options, kind, swords, win, sym_swords):
    """k_sym_gwin / k_num_gwin with every window geometry: one or many super-windows (32*swords columns),
    one or many rank windows (win entries), last super-window partial (column counts are not multiples of 32)."""
    A, B = _global_operands(kind)
    options("global_rows_smem", 1)
    options("gwin_max_sw", 0)
    options("gwin_swords", swords)
    options("gwin_win", win)
    options("gwin_sym_swords", sym_swords)
    got, st = _check(eng, oracle, A, B, mag=False)
    assert st["num_bin_rows"][5] > 0 and st["sym_bin_rows"][5] > 0


@pytest.mark.parametrize("kind", ["few_a_entries", "mid_a_entries", "long_a_rows"])
@pytest.mark.parametrize("g_win,g_coop", [(20480, 1), (4096, 1), (64, 1), (4096, 0)])
def test_global_rows_l2_kernel_with_shared_memory_mark(eng, oracle, options, kind, g_win, g_coop):
    """First generation of the L2 bitmap/rank kernel (g_v2 = 0): mark pass in shared memory one super-window
    (64 * g_win columns) at a time, flushed to the row's cells, separate rank + emit pass over L2."""
    A, B = _global_operands(kind)
    options("global_rows_smem", 1)
    options("gwin_swords", 32)
    options("gwin_max_sw", 1)              # 32 * 32 columns per super-window: the windowed numeric kernel is refused
    options("g_v2", 0)
    options("g_win", g_win)
    options("g_coop", g_coop)
    got, st = _check(eng, oracle, A, B, mag=False)
    assert st["num_bin_rows"][5] > 0


@pytest.mark.parametrize("kind", ["few_a_entries", "mid_a_entries", "long_a_rows"])
@pytest.mark.parametrize("g_win,g_tbl,g_lpt", [(16384, 8192, 1), (4096, 8192, 1), (64, 8192, 0), (16, 16384, 1), (256, 0, 1), (128, 300, 0), (2048, 8192, 1)])
def test_global_rows_second_generation(eng, oracle, options, kind, g_win, g_tbl, g_lpt):
    """k_num_global2 (default for wide column spaces): rank + emit from the shared-memory bitmap, split tables for the
    super-windows of the mark pass and for the rank windows of the accumulate pass.  Window geometries: one or many
    super-windows (64 * g_win columns each), one or hundreds of rank windows, tables that fit / do not fit / are off,
    rows with more A entries than threads, rows handed out in order of work or of index."""
    A, B = _global_operands(kind)
    options("global_rows_smem", 1)
    options("gwin_swords", 32)
    options("gwin_max_sw", 1)
    options("g_v2", 1)
    options("g_win", g_win)
    options("g_tbl", g_tbl)
    options("g_lpt", g_lpt)
    got, st = _check(eng, oracle, A, B, mag=False)
    assert st["num_bin_rows"][5] > 0


@pytest.mark.parametrize("g_win,g_scr", [(64, 8 << 20), (256, 8 << 20), (64, 0), (64, 40000), (16384, 8 << 20)])
def test_global_rows_long_rows_use_global_split_tables(eng, oracle, options, g_win, g_scr):
    """Rows with more than 1024 entries in A: their split tables live in a per-CTA global scratch (g_scr ints); too
    small a scratch, or none, falls back to per-window searches."""
    A, B = _global_operands("long_a_rows")
    options("global_rows_smem", 1)
    options("gwin_swords", 32)
    options("gwin_max_sw", 1)
    options("g_v2", 1)
    options("g_win", g_win)
    options("g_scr", g_scr)
    got, st = _check(eng, oracle, A, B, mag=False)
    assert st["num_bin_rows"][5] > 0


@pytest.mark.parametrize("kind", ["few_a_entries", "mid_a_entries", "long_a_rows"])
@pytest.mark.parametrize("g_win,parts,g_scr,g_tbl", [(16384, 128, 8 << 20, 8192), (64, 128, 8 << 20, 8192), (256, 3, 8 << 20, 0), (2048, 2, 0, 8192),
                                                       (64, 7, 40000, 300), (1024, 128, 8 << 20, 8192)])
def test_global_rows_split_over_ctas(eng, oracle, options, kind, g_win, parts, g_scr, g_tbl):
    """k_num_global2 with every global row cut into column-range parts (g_split_ub = 1: any row qualifies), one work
    item per part: parts publish their counts and wait for the counts before them.  Geometries: parts of one bitmap
    chunk (4096 columns) up to half the column space, one or many rank windows per part, empty parts, long A rows with
    and without the global scratch (restricted cached B rows vs. per-window searches)."""
    A, B = _global_operands(kind)
    options("global_rows_smem", 1)
    options("gwin_swords", 32)
    options("gwin_max_sw", 1)
    options("g_v2", 1)
    options("g_lpt", 1)
    options("g_split", 1)
    options("g_split_ub", 1)
    options("g_split_parts", parts)
    options("g_win", g_win)
    options("g_scr", g_scr)
    options("g_tbl", g_tbl)
    got, st = _check(eng, oracle, A, B, mag=False)
    assert st["num_bin_rows"][5] > 0
    options("g_split", 0)                      # and the same C with one CTA per row
    got1, st1 = _mul(eng, A, B)
    assert np.array_equal(got[0], got1[0]) and np.array_equal(got[1], got1[1])
    assert np.all(np.abs(got[2] - got1[2]) <= RTOL * np.maximum(np.abs(got[2]), np.abs(got1[2])))


@pytest.mark.parametrize("parts,g_split_ub", [(4, 1), (2, 20000), (128, 0)])
def test_global_rows_split_many_rows(eng, oracle, options, parts, g_split_ub):
    """R-MAT scale 14 with the global rows on k_num_global2: hundreds of rows cut into parts at once (more work items
    than CTAs, so parts wait for counts published by CTAs that drew the earlier items); a threshold in the middle of
    the work-ordered list; the automatic threshold."""
    A = W.rmat(14, 16, seed=3)
    options("global_rows_smem", 1)
    options("gwin_swords", 32)
    options("gwin_max_sw", 1)
    options("g_v2", 1)
    options("g_win", 1024)
    options("g2_takes_b2", 1)                  # rows of 4097..12288 entries too: enough rows for every CTA
    options("g_split_ub", g_split_ub)
    options("g_split_parts", parts)
    got, st = _check(eng, oracle, A, mag=False)
    assert st["num_bin_rows"][5] > 20


def test_global_rows_second_generation_many_boundaries(eng, oracle, options):
    """More than 512 rank windows in a row: the window boundaries no longer fit shared memory and are read back
    from the emitted column list."""
    A = W.random_sparse(3, 60, 0.9, seed=31)
    B = W.random_sparse(60, 40000, 0.25, seed=32)
    options("global_rows_smem", 1)
    options("gwin_swords", 32)
    options("gwin_max_sw", 1)
    options("g_v2", 1)
    options("g_win", 16)
    got, st = _check(eng, oracle, A, B, mag=False)
    assert st["num_bin_rows"][5] > 0 and st["nnz"] > 3 * 16 * 512


def test_global_rows_both_kernel_families_agree(eng, oracle, options):
    """The L2 bitmap kernels (kept for non-canonical B) and the windowed kernels give the same C."""
    A, B = _global_operands("mid_a_entries")
    dA, dB = eng.upload(*A), eng.upload(*B)
    options("global_rows_smem", 0)
    (rp0, ci0, v0), st0 = eng.CSR_MUL_CSR_DEV(dA, dB)
    options("global_rows_smem", 1)
    options("gwin_max_sw", 0)
    options("gwin_swords", 64)
    options("gwin_win", 1024)
    (rp1, ci1, v1), st1 = eng.CSR_MUL_CSR_DEV(dA, dB)
    dA.close()
    dB.close()
    assert st0["num_bin_rows"][5] > 0 and st1["num_bin_rows"][5] > 0
    assert np.array_equal(rp0, rp1) and np.array_equal(ci0, ci1)
    assert np.all(np.abs(v0 - v1) <= RTOL * np.maximum(np.abs(v0), np.abs(v1)))


def test_unknown_option_is_an_error(eng):
    from ia_spgemm_b200.engine import EngineError
    with pytest.raises(EngineError):
        eng.set_option("no_such_knob", 1)
