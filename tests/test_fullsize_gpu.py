"""Parity at BASELINE.json's full sizes through size-independent properties, and against the oracle on
sampled row blocks where the whole product is too large for the CPU."""
import ctypes as C

import numpy as np
import pytest

from ia_spgemm_b200 import workloads as W
from util import abs_product, assert_csr_parity, sort_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ia_spgemm_b200.engine import get_engine
    return get_engine()


def test_poisson_4096_closed_forms(eng):
    """configs[1]: nnz(A), products and nnz(C) have closed forms; sum(C) = |A*1|^2 = 16 + 4(N-2) because the
    row sums of the 5-point operator are 0 inside, 1 on edges and 2 in corners; rows come out strictly sorted."""
    from ia_spgemm_b200.engine import CsrMatrixDev, DeviceCsr
    N = 4096
    dA = eng.gen_poisson2d(N)
    nnz_a, products, nnz_c = W.poisson_counts(N)
    assert dA.nnz == nnz_a
    assert eng.GetFlop(dA, dA) == products
    cdev, ms = CsrMatrixDev(), C.c_double()
    eng._ck(eng.lib.ias_csr_mul_csr_dev(C.byref(dA.dev), C.byref(dA.dev), C.byref(cdev), C.byref(ms)))
    Cm = DeviceCsr(eng, cdev)
    assert Cm.nnz == nnz_c
    assert eng.is_canonical(Cm)                                   # every row strictly increasing in column
    total = eng.checksum_ptr(cdev.values_dev, cdev.nnz)
    assert total == pytest.approx(16 + 4 * (N - 2), rel=1e-12)
    f = eng.GetInfo1(Cm)
    assert (f[4], f[5]) == (13, 6)                                # interior rows 13 entries, corner rows 6
    Cm.close()
    # the DIA path stores the same matrix (plus explicit zeros on 13 diagonals)
    d = eng.CSRtoDIA(dA, gate=20.0)
    assert d.choice and d.num_diagonals == 5
    c, _ = eng.DIA_MUL_DIA_DEV(d, d)
    assert c.num_diagonals == 13
    assert eng.checksum_ptr(c.values_dev, c.row * c.num_diagonals) == pytest.approx(total, rel=1e-12)
    eng.free_dia(c); eng.free_dia(d)
    # streaming and row blocks agree with the materialised run
    st = eng.csr_mul_csr_stream(dA, dA, budget_bytes=600 * 10**6)
    assert st["nnz"] == nnz_c and st["products"] == products and st["batches"] > 1
    assert st["checksum"] == pytest.approx(total, rel=1e-11)
    b = eng.partition_rows(dA, dA, 8)
    assert sum(eng.CSR_MUL_CSR_DEV(dA, dA, rows=(x, y), download=False)[1]["nnz"] for x, y in zip(b, b[1:])) == nnz_c
    dA.close()


def test_rmat18_sampled_row_blocks_against_oracle(eng, oracle):
    """R-MAT scale 18: the head rows (global bin, several windows per row), a middle block and the tail,
    each compared entry by entry with CSR_MUL_CSR(A[r0:r1,:], A) on the CPU."""
    A = W.rmat(18, 16, seed=1)
    rows, cols, rp, ci, v = A
    dA = eng.gen_rmat(18, 16, seed=1)
    assert dA.nnz == int(rp[-1])
    for r0, r1 in ((0, 48), (3000, 3400), (rows // 2, rows // 2 + 4000), (rows - 20000, rows)):
        s, e = int(rp[r0]), int(rp[r1])
        blk_rp = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
        want = oracle.csr_mul_csr(r1 - r0, cols, blk_rp, ci[s:e], v[s:e], rp, ci, v)
        got, st = eng.CSR_MUL_CSR_DEV(dA, dA, rows=(r0, r1))
        assert_csr_parity(got, want)                    # values are positive: the entry itself is the scale
        assert st["products"] == oracle.getflop(blk_rp, ci[s:e], rp)
    dA.close()


def test_rmat16_streaming_equals_materialised(eng):
    dA = eng.gen_rmat(16, 16, seed=1)
    c64, st = eng.CSR_MUL_CSR_DEV(dA, dA, keep=True)
    h = eng.structure_hash(c64)
    s = eng.checksum_ptr(c64.values_dev, c64.nnz)
    eng.free_csr64(c64)
    d = eng.csr_mul_csr_stream(dA, dA, budget_bytes=400 * 10**6, want_row_nnz=True)
    assert d["batches"] > 2 and d["nnz"] == st["nnz"] and d["structure_hash"] == h
    assert d["checksum"] == pytest.approx(s, rel=1e-11)
    assert int(d["row_nnz"].sum()) == st["nnz"]
    dA.close()


def test_uniform_full_size_properties(eng):
    """configs[2]: 8M x 8M, 16 distinct columns per row: products = 2^11 * 10^6 exactly, every row of C has at
    most 256 entries, nnz(C) <= products, ELL path agrees on nnz and checksum."""
    n = 8_000_000
    dA = eng.gen_uniform(n, 16, seed=1)
    assert eng.is_canonical(dA)
    st = eng.csr_mul_csr_stream(dA, dA, want_row_nnz=True)
    assert st["products"] == n * 256
    assert st["nnz"] <= st["products"] and st["nnz"] > 0.9999 * st["products"]
    assert int(st["row_nnz"].max()) <= 256 and int(st["row_nnz"].min()) >= 200
    e = eng.CSRtoELL(dA, gate=20.0)
    assert e.choice and e.max_nnz_per_row == 16
    c, _ = eng.ELL_MUL_ELL_DEV(e, e)
    assert c.nnz == st["nnz"] and c.max_nnz_per_row == int(st["row_nnz"].max())
    assert eng.checksum_ptr(c.values_dev, c.row * c.max_nnz_per_row) == pytest.approx(st["checksum"], rel=1e-11)
    eng.free_ell(c); eng.free_ell(e)
    dA.close()


# ---------------------------------------------------------------- entry-wise parity at BASELINE.json's full sizes
# C does not fit the host oracle as a whole, so deterministic row blocks are compared entry by
# entry with CSR_MUL_CSR(A[r0:r1,:], A) (the reference function takes a rectangular A, csr:88-89): row_ptr and sorted
# columns bit-exact, values within 1e-12 of the entry.  The device operand is downloaded (the generators are pinned
# bit-identical to workloads.py in test_formats_gpu.py) so that the host never regenerates 10^8 entries.
def _host_memory_available():
    """Bytes this process may still use: the smaller of the machine's available memory and the cgroup's headroom."""
    import psutil
    avail = psutil.virtual_memory().available
    for lim, cur in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                     ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            limit = open(lim).read().strip()
            if limit != "max" and int(limit) < 2**60:
                avail = min(avail, int(limit) - int(open(cur).read().strip()))
        except Exception:
            pass
    return avail


def _block_parity(eng, oracle, dA, host, blocks, mag=False):
    rows, cols, rp, ci, v = host
    # the oracle keeps a dense accumulator of `cols` entries per thread (16 bytes each): bound the host memory
    oracle.set_threads(min(oracle.threads(), 8 if cols <= (1 << 23) else 2))
    for r0, r1 in blocks:
        s, e = int(rp[r0]), int(rp[r1])
        blk_rp = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
        want = oracle.csr_mul_csr(r1 - r0, cols, blk_rp, ci[s:e], v[s:e], rp, ci, v)
        got, st = eng.CSR_MUL_CSR_DEV(dA, dA, rows=(r0, r1))
        m = None
        if mag:
            a = oracle.csr_mul_csr(r1 - r0, cols, blk_rp, ci[s:e], np.abs(v[s:e]), rp, ci, np.abs(v))
            m = sort_rows(*a)[2]
        assert_csr_parity(got, want, mag=m)
        assert st["products"] == oracle.getflop(blk_rp, ci[s:e], rp)
        yield (r0, r1), got, st


def test_poisson_4096_sampled_blocks_entrywise(eng, oracle):
    """configs[1] at full size: the first grid line (boundary rows), an interior band, the last grid line."""
    N = 4096
    dA = eng.gen_poisson2d(N)
    host = dA.download()
    n = N * N
    blocks = ((0, N + 5), (N * 2000 - 3, N * 2000 + 3000), (n - N - 7, n))
    for (r0, r1), got, st in _block_parity(eng, oracle, dA, host, blocks, mag=True):      # 4 / -1 stencil: terms cancel
        inner = np.diff(got[0])
        # 13 entries inside, 12 on the second / second-to-last grid line, down to 6 in the corners
        assert inner.max() == (13 if r0 > 2 * N and r1 < n - 2 * N else 12) and inner.min() >= 6
    dA.close()


def test_uniform_8m_sampled_blocks_entrywise(eng, oracle):
    """configs[2] at full size: three row blocks of the warp bin (16 x 16 products per row, nnz(C_i) <= 256)."""
    n = 8_000_000
    dA = eng.gen_uniform(n, 16, seed=1)
    host = dA.download()
    for (r0, r1), got, st in _block_parity(eng, oracle, dA, host, ((0, 3000), (n // 2 + 11, n // 2 + 3011), (n - 3000, n))):
        assert st["products"] == 256 * (r1 - r0)
    dA.close()


def test_rmat22_sampled_blocks_entrywise_and_streamed_totals(eng, oracle):
    """configs[3] at full size.  Rows 0-32 are the hub rows (tens of millions of products each, through the
    multi-window global-row kernel), then a middle block and the tail.  The streamed run of the whole matrix must
    report the same per-row nnz on those rows, and its totals must be self-consistent."""
    dA = eng.gen_rmat(22, 16, seed=1)
    host = dA.download()
    rows = host[0]
    blocks = ((0, 32), (40_000, 40_400), (rows // 2, rows // 2 + 3000), (rows - 30_000, rows))
    seen = {}
    for (r0, r1), got, st in _block_parity(eng, oracle, dA, host, blocks):
        seen[(r0, r1)] = (np.diff(got[0]), float(got[2].sum()))
        if r0 == 0:
            assert st["num_bin_rows"][5] > 0 and np.diff(got[0]).max() > 500_000       # really the global-row path
    st = eng.csr_mul_csr_stream(dA, dA, want_row_nnz=True)
    assert st["batches"] > 1 and st["products"] == eng.GetFlop(dA, dA) and st["products"] > 2 ** 37
    assert int(st["row_nnz"].astype(np.int64).sum()) == st["nnz"] and st["nnz"] > 2 ** 35      # beyond int32 by far
    for (r0, r1), (cnt, _) in seen.items():
        assert np.array_equal(st["row_nnz"][r0:r1], cnt)
    dA.close()


def test_rmat25_blocks_on_one_gpu(eng, oracle):
    """configs[4]: the scale-25 operand (5.3e8 entries, 6.5 GB) on one GPU, two row blocks against the oracle --
    33.5 M columns (1 M bitmap words, 8 MB workspace slots), products and offsets far beyond int32."""
    import os
    if os.environ.get("IAS_TEST_RMAT25") != "1":
        # Opt-in: on the shared GPU hosts of this pool the test process was killed twice (SIGKILL, no Python error) while
        # it held the 6.5 GB host copy of the operand, although the box reported 195 GB free and no cgroup limit -- the
        # rest of the suite must not depend on that.  IAS_TEST_RMAT25=1 python -m pytest tests/test_fullsize_gpu.py -m gpu -k rmat25
        pytest.skip("opt-in (IAS_TEST_RMAT25=1): holds a 6.5 GB host copy of the scale-25 operand")
    info = eng.device_info()
    if info["free_bytes"] < 60 * 10**9:
        pytest.skip("needs ~60 GB of free HBM for the scale-25 generator")
    if _host_memory_available() < 24 * 2**30:
        pytest.skip("needs ~10 GB of host memory (operand copy + the oracle's per-thread accumulators) with headroom")
    oracle.set_threads(2)
    dA = eng.gen_rmat(25, 16, seed=1)
    assert dA.nnz > 500_000_000
    host = dA.download()
    rows = host[0]
    for (r0, r1), got, st in _block_parity(eng, oracle, dA, host, ((64, 72), (rows // 2 + 5, rows // 2 + 2005))):
        if r0 == 64:
            assert st["num_bin_rows"][5] > 0
    dA.close()
