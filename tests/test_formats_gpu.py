"""Parity of the DIA / ELL / COO kernels, the converters, the density image, the features and the
device generators with the CPU oracle and the committed reference fixtures (through the C ABI)."""
import os

import numpy as np
import pytest

from ia_spgemm_b200 import workloads as W
from util import RECT, RTOL, SQUARE, abs_product, assert_csr_parity, decode_img, sort_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ia_spgemm_b200.engine import get_engine
    return get_engine()


def _load(oracle, mtx_dir, name):
    return oracle.mtx_load(os.path.join(mtx_dir, name + ".mtx"))


CASES = [
    ("poisson40", lambda: W.poisson2d(40)),
    ("banded", lambda: W.banded(700, [-9, -1, 0, 2, 17], seed=4)),
    ("uniform", lambda: W.uniform_rows(3000, 8, seed=3)),
    ("irregular", lambda: W.random_sparse(257, 257, 0.03, seed=5)),
    ("unsorted", lambda: W.random_sparse(120, 120, 0.08, seed=6, sort_columns=False)),
    ("rmat9", lambda: W.rmat(9, 8, seed=2)),
]


def _close(got, want, scale=None, rtol=RTOL):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    s = np.abs(want) if scale is None else np.maximum(np.abs(scale), np.abs(want))
    return bool(np.all(np.abs(got - want) <= rtol * s))


# ---------------------------------------------------------------- DIA
@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_csr_to_dia_matches_oracle(eng, oracle, name, make):
    A = make()
    dA = eng.upload(*A)
    for gate in (20.0, 50.0):
        want = oracle.csr_to_dia(*A, gate=gate)
        d = eng.CSRtoDIA(dA, gate=gate)
        assert bool(d.choice) == want["choice"] and d.num_diagonals == want["num_diagonals"]
        if want["choice"]:
            got = eng.download_dia(d)
            assert np.array_equal(got["diagonal_offsets"], want["diagonal_offsets"])
            assert np.array_equal(got["diagonal_ind"], want["diagonal_ind"])
            assert np.array_equal(got["values"], want["values"])
        eng.free_dia(d)
    dA.close()


@pytest.mark.parametrize("name,make", CASES[:2] + CASES[4:5], ids=[c[0] for c in CASES[:2] + CASES[4:5]])
def test_dia_mul_dia_matches_oracle(eng, oracle, name, make):
    A = make()
    want_a = oracle.csr_to_dia(*A, gate=1e9)
    want = oracle.dia_mul_dia(want_a, want_a)
    dA = eng.upload(*A)
    d = eng.CSRtoDIA(dA, gate=1e9)
    c, ms = eng.DIA_MUL_DIA_DEV(d, d)
    got = eng.download_dia(c)
    assert np.array_equal(got["diagonal_offsets"], want["diagonal_offsets"])     # offsets exact
    assert np.array_equal(got["diagonal_ind"], want["diagonal_ind"])
    absa = dict(want_a, values=np.abs(want_a["values"]))
    mag = oracle.dia_mul_dia(absa, absa)["values"]
    assert _close(got["values"], want["values"], scale=mag)                      # full padded array, 1e-12
    # (ii) SURVEY appendix A: the CSR result sits inside the DIA result, every other position is 0
    rp, ci, v = sort_rows(*oracle.csr_mul_csr(A[0], A[1], A[2], A[3], A[4], A[2], A[3], A[4]))
    rows = np.repeat(np.arange(A[0]), np.diff(rp))
    slot = np.searchsorted(got["diagonal_offsets"], ci.astype(np.int64) - rows)
    dense = got["values"].copy()
    magc = abs_product(oracle, A, A)
    assert _close(dense[rows, slot], v, scale=magc)
    dense[rows, slot] = 0.0
    assert np.all(np.abs(dense) <= RTOL * mag)
    assert ms > 0
    eng.free_dia(c); eng.free_dia(d); dA.close()


@pytest.mark.parametrize("name", SQUARE)
def test_dia_golden(eng, oracle, golden, mtx_dir, name):
    g = golden["inputs"][name]
    if "dia_offsets_c" not in g:
        pytest.skip("reference refused DIA for this input")
    A = _load(oracle, mtx_dir, name)
    dA = eng.upload(*A)
    d = eng.CSRtoDIA(dA, gate=50.0)
    assert eng.download_dia(d)["diagonal_offsets"].tolist() == g["dia_offsets_a"]
    c, _ = eng.DIA_MUL_DIA_DEV(d, d)
    got = eng.download_dia(c)
    assert got["diagonal_offsets"].tolist() == g["dia_offsets_c"]
    assert got["diagonal_ind"].tolist() == g["dia_diag_ind_c"]
    want = np.array(g["dia_values_c"], dtype=np.float64)
    absd = oracle.csr_to_dia(A[0], A[1], A[2], A[3], np.abs(A[4]), gate=50.0)
    mag = oracle.dia_mul_dia(absd, absd)["values"].ravel()
    assert _close(got["values"].ravel(), want, scale=mag)
    assert eng.lib.ias_sizeof_dia(c.row, c.col, c.num_diagonals) == oracle.sizeof_dia(c.row, c.col, c.num_diagonals)
    eng.free_dia(c); eng.free_dia(d); dA.close()


def test_dia_gate_refuses_scattered(eng, oracle):
    from ia_spgemm_b200.engine import EngineError
    S = W.uniform_rows(4000, 2, seed=1)
    dS = eng.upload(*S)
    d = eng.CSRtoDIA(dS, gate=20.0)
    assert not d.choice and d.num_diagonals == oracle.csr_to_dia(*S, gate=20.0)["num_diagonals"]
    with pytest.raises(EngineError) as e:
        eng.DIA_MUL_DIA_DEV(d, d)
    assert e.value.code == 5
    dS.close()


def test_dia_rectangular(eng, oracle):
    A = W.random_sparse(30, 45, 0.2, seed=1)
    B = W.random_sparse(45, 25, 0.2, seed=2)
    oa, ob = oracle.csr_to_dia(*A, gate=1e9), oracle.csr_to_dia(*B, gate=1e9)
    want = oracle.dia_mul_dia(oa, ob)
    dA, dB = eng.upload(*A), eng.upload(*B)
    da, db = eng.CSRtoDIA(dA, gate=1e9), eng.CSRtoDIA(dB, gate=1e9)
    c, _ = eng.DIA_MUL_DIA_DEV(da, db)
    got = eng.download_dia(c)
    assert np.array_equal(got["diagonal_offsets"], want["diagonal_offsets"])
    assert np.array_equal(got["diagonal_ind"], want["diagonal_ind"])
    assert np.allclose(got["values"], want["values"], rtol=1e-12, atol=1e-14)
    for x in (c, da, db):
        eng.free_dia(x)
    dA.close(); dB.close()


# ---------------------------------------------------------------- ELL
@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_ell_matches_oracle(eng, oracle, name, make):
    A = make()
    dA = eng.upload(*A)
    want_a = oracle.csr_to_ell(*A, gate=20.0)
    e = eng.CSRtoELL(dA, gate=20.0)
    assert bool(e.choice) == want_a["choice"] and e.max_nnz_per_row == want_a["width"]
    if want_a["choice"]:
        got_a = eng.download_ell(e)
        for key in ("nnz_row", "col_ind", "values"):
            assert np.array_equal(got_a[key], want_a[key]), key
        want = oracle.ell_mul_ell(want_a, want_a)
        c, ms = eng.ELL_MUL_ELL_DEV(e, e)
        got = eng.download_ell(c)
        assert got["width"] == want["width"] and got["nnz"] == want["nnz"]
        assert np.array_equal(got["nnz_row"], want["nnz_row"])
        mag = abs_product(oracle, A, A)
        p = 0
        for i in range(A[0]):                     # engine rows are sorted, the reference's are reverse first-touch
            n = int(want["nnz_row"][i])
            o = np.argsort(want["col_ind"][i, :n], kind="stable")
            assert np.array_equal(got["col_ind"][i, :n], want["col_ind"][i, :n][o])
            assert _close(got["values"][i, :n], want["values"][i, :n][o], scale=mag[p:p + n])
            assert not got["col_ind"][i, n:].any() and not got["values"][i, n:].any()      # 0 / 0.0 padding
            p += n
        eng.free_ell(c)
    eng.free_ell(e); dA.close()


@pytest.mark.parametrize("name", SQUARE)
def test_ell_golden(eng, oracle, golden, mtx_dir, name):
    g = golden["inputs"][name]
    if "ell_width_c" not in g:
        pytest.skip("reference refused ELL for this input")
    A = _load(oracle, mtx_dir, name)
    dA = eng.upload(*A)
    e = eng.CSRtoELL(dA, gate=50.0)
    c, _ = eng.ELL_MUL_ELL_DEV(e, e)
    got = eng.download_ell(c)
    assert got["width"] == g["ell_width_c"] and got["nnz_row"].tolist() == g["ell_nnz_row_c"]
    w = got["width"]
    wc = np.array(g["ell_col_ind_c"]).reshape(A[0], w)
    wv = np.array(g["ell_values_c"], dtype=np.float64).reshape(A[0], w)
    for i in range(A[0]):
        n = got["nnz_row"][i]
        o = np.argsort(wc[i, :n], kind="stable")
        assert np.array_equal(got["col_ind"][i, :n], wc[i, :n][o])
        assert np.allclose(got["values"][i, :n], wv[i, :n][o], rtol=1e-12, atol=1e-9 if name in ("LFAT5", "b1_ss") else 0)
    eng.free_ell(c); eng.free_ell(e); dA.close()


# ---------------------------------------------------------------- COO
@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_coo_matches_oracle(eng, oracle, name, make):
    A = make()
    dA = eng.upload(*A)
    k = eng.CSRtoCOO(dA)
    got_a = eng.download_coo(k)
    want_a = oracle.csr_to_coo(A[0], A[2], A[3], A[4])
    for key in ("row_offset", "row_ind", "col_ind", "values"):
        assert np.array_equal(got_a[key], want_a[key]), key
    c, ms = eng.COO_MUL_COO_DEV(k, k)
    got = eng.download_coo(c)
    want = oracle.csr_mul_csr(A[0], A[1], A[2], A[3], A[4], A[2], A[3], A[4])     # same symbolic pass as COO_MUL_COO
    rp, ci, v = sort_rows(*want)
    assert np.array_equal(got["row_offset"], rp) and np.array_equal(got["col_ind"], ci)
    assert np.array_equal(got["row_ind"], np.repeat(np.arange(A[0]), np.diff(rp)))
    assert _close(got["values"], v, scale=abs_product(oracle, A, A))
    if A[0] <= 600:                               # the reference's own COO kernel (first-touch order), small cases only
        wc = oracle.coo_mul_coo(A[0], A[1], want_a, want_a)
        o = np.lexsort((wc["col_ind"], wc["row_ind"]))
        assert np.array_equal(got["col_ind"], wc["col_ind"][o]) and np.array_equal(got["row_offset"], wc["row_offset"])
        assert _close(got["values"], wc["values"][o], scale=abs_product(oracle, A, A))
    assert eng.lib.ias_sizeof_coo(A[0], c.nnz) == oracle.sizeof_coo(A[0], c.nnz)
    c32, _ = eng.COO_MUL_COO_DEV(k, k, int32=True)               # the reference's CooMatrixDev layout (int32 nnz / row_offset)
    got32 = eng.download_coo(c32)
    for key in ("row_offset", "row_ind", "col_ind"):
        assert np.array_equal(got32[key], got[key]), key
    assert np.allclose(got32["values"], got["values"], rtol=1e-13, atol=0)      # two runs: shared-memory additions in another order
    eng.free_coo(c32)
    eng.free_coo(c); eng.free_coo(k); dA.close()


# ---------------------------------------------------------------- density + features
@pytest.mark.parametrize("name", SQUARE + RECT)
def test_density_golden(eng, oracle, golden, mtx_dir, name):
    A = _load(oracle, mtx_dir, name)
    dA = eng.upload(*A)
    assert np.array_equal(eng.density_image(dA), decode_img(golden["inputs"][name]["density"]))
    dA.close()


def test_density_shipped_images(eng, oracle, golden, mtx_dir):
    ship = golden["shipped_imgs"]
    for name, key in (("dia", "gpu_img1_dia"), ("small", "cpu_img1_small")):
        A = _load(oracle, mtx_dir, name)
        dA = eng.upload(*A)
        assert np.array_equal(eng.density_image(dA), decode_img(ship[key]))
        dA.close()


@pytest.mark.parametrize("name,make", CASES + [("exact128", lambda: W.random_sparse(128, 128, 0.1, seed=8)),
                                                ("wide", lambda: W.random_sparse(50, 5000, 0.01, seed=9))],
                         ids=[c[0] for c in CASES] + ["exact128", "wide"])
def test_density_and_features_match_oracle(eng, oracle, name, make):
    A = make()
    dA = eng.upload(*A)
    assert np.array_equal(eng.density_image(dA), oracle.density(A[0], A[1], A[2], A[3]))
    want = oracle.features26(A, A, gate=20.0)
    got = eng.features26(dA, dA)
    assert np.allclose(got, want, rtol=1e-12, atol=0), (got, want)
    assert eng.count_diagonals(dA) == oracle.csr_to_dia(*A, gate=1e9)["num_diagonals"]
    assert eng.max_row_nnz(dA) == int(np.diff(A[2]).max())
    assert eng.GetFlop(dA, dA) == oracle.getflop(A[2], A[3], A[2]) if A[0] == A[1] else True
    dA.close()


def test_features_screenshot(eng, oracle, golden, mtx_dir):
    A = _load(oracle, mtx_dir, "dia")
    dA = eng.upload(*A)
    assert eng.features26(dA, dA).tolist() == golden["screenshot_features_dia"]
    dA.close()


# ---------------------------------------------------------------- loader + generators
@pytest.mark.parametrize("name", SQUARE + RECT)
def test_engine_loader_golden(eng, golden, mtx_dir, name):
    g = golden["inputs"][name]
    rows, cols, rp, ci, v = eng.mtx_load(os.path.join(mtx_dir, name + ".mtx"))
    assert (rows, cols) == (g["rows"], g["cols"])
    assert rp.tolist() == g["loader_row_ptr"] and ci.tolist() == g["loader_col_ind"] and v.tolist() == g["loader_values"]


def test_device_generators_are_bit_identical_to_numpy(eng):
    for dev, host in ((eng.gen_poisson2d(37), W.poisson2d(37)), (eng.gen_poisson2d(16, 48), W.poisson2d(16, 48)),
                      (eng.gen_uniform(5000, 16, seed=1), W.uniform_rows(5000, 16, seed=1)),
                      (eng.gen_uniform(40, 16, seed=3), W.uniform_rows(40, 16, seed=3)),     # collisions -> redraws
                      (eng.gen_rmat(12, 16, seed=1), W.rmat(12, 16, seed=1)),
                      (eng.gen_rmat(9, 4, seed=5), W.rmat(9, 4, seed=5))):
        rows, cols, rp, ci, v = dev.download()
        assert (rows, cols) == (host[0], host[1])
        assert np.array_equal(rp, host[2]) and np.array_equal(ci, host[3]) and np.array_equal(v, host[4])
        assert eng.is_canonical(dev)
        dev.close()


# ---------------------------------------------------------------- transpose / A * A^T (GPU/main.cu:261-269) / writer
@pytest.mark.parametrize("name", SQUARE + RECT)
def test_a_times_a_transpose_bundled(eng, oracle, mtx_dir, name):
    """What the reference's GPU driver actually multiplies: B := A^T.  Works for the rectangular inputs too."""
    import scipy.sparse as sp
    A = _load(oracle, mtx_dir, name)
    rows, cols, rp, ci, v = A
    dA = eng.upload(*A)
    dT = eng.transpose(dA)
    t_rows, t_cols, t_rp, t_ci, t_v = dT.download()
    T = sp.csr_matrix((v, ci, rp), shape=(rows, cols)).T.tocsr()
    T.sort_indices()
    assert (t_rows, t_cols) == (cols, rows)
    assert np.array_equal(t_rp, T.indptr) and np.array_equal(t_ci, T.indices) and np.array_equal(t_v, T.data)
    assert eng.is_canonical(dT)
    got, st = eng.CSR_MUL_CSR_DEV(dA, dT)
    B = (cols, rows, t_rp, t_ci, t_v)
    want = oracle.csr_mul_csr(rows, rows, rp, ci, v, t_rp, t_ci, t_v)
    assert_csr = __import__("util").assert_csr_parity
    assert_csr(got, want, mag=abs_product(oracle, A, B))
    assert st["products"] == oracle.getflop(rp, ci, t_rp)
    dT.close(); dA.close()


def test_dia_times_dia_transpose_screenshot_values(eng, oracle, golden, mtx_dir):
    """GPU/2.jpg: dia.mtx * dia.mtx^T has 13 products, 10 stored entries, memory_size 152."""
    g = golden["dia_times_diaT"]
    A = _load(oracle, mtx_dir, "dia")
    dA = eng.upload(*A)
    dT = eng.transpose(dA)
    assert dT.download()[2].tolist() == g["b_row_ptr"] and dT.download()[3].tolist() == g["b_col_ind"]
    (rp, ci, v), st = eng.CSR_MUL_CSR_DEV(dA, dT)
    assert st["nnz"] == g["nnz"] == 10 and st["products"] == g["flop"] == 13
    assert eng.lib.ias_sizeof_csr(4, st["nnz"]) == g["sizeof_csr"] == 152.0
    dT.close(); dA.close()


def test_transpose_large_random_and_roundtrip(eng):
    import scipy.sparse as sp
    A = W.rmat(12, 8, seed=9)
    dA = eng.upload(*A)
    dT = eng.transpose(dA)
    dTT = eng.transpose(dT)
    r = dTT.download()
    assert np.array_equal(r[2], A[2]) and np.array_equal(r[3], A[3]) and np.array_equal(r[4], A[4])
    T = sp.csr_matrix((A[4], A[3], A[2]), shape=(A[0], A[1])).T.tocsr(); T.sort_indices()
    t = dT.download()
    assert np.array_equal(t[2], T.indptr) and np.array_equal(t[3], T.indices) and np.array_equal(t[4], T.data)
    for d in (dTT, dT, dA):
        d.close()


def test_matrix_market_writer_roundtrip(eng, oracle, tmp_path):
    A = W.random_sparse(40, 40, 0.1, seed=2)
    dA = eng.upload(*A)
    c64, st = eng.CSR_MUL_CSR_DEV(dA, dA, keep=True)
    path = str(tmp_path / "c.mtx")
    eng.mtx_write(path, c64)
    want = eng._take_csr64(c64)
    rows, cols, rp, ci, v = eng.mtx_load(path)
    assert (rows, cols) == (40, 40)
    assert np.array_equal(rp, want[0]) and np.array_equal(ci, want[1]) and np.array_equal(v, want[2])   # %.17g round-trips fp64
    dA.close()


def test_ell_and_coo_views_through_every_bin(eng, oracle):
    """The CTA / global kernels are templated on the operand view: run them on ELL (fixed width) and COO
    (64-bit row offsets) operands that populate the CTA bins (R-MAT scale 13) and the global bin (a dense-ish
    rectangular product)."""
    R13 = W.rmat(13, 16, seed=5)
    cases = [(R13, None, (3, 4), 0),        # large CTA hash kept for its rows
             (R13, None, (3, 5), 1),        # default: those rows go to the windowed shared-memory kernel
             (W.random_sparse(40, 300, 0.9, seed=3), W.random_sparse(300, 40000, 0.4, seed=4), (5,), 1)]
    for A, B, bins, takes_b2 in cases:
        eng.set_option("gwin_takes_b2", takes_b2)
        B = A if B is None else B
        dA = eng.upload(*A)
        dB = dA if B is A else eng.upload(*B)
        want = sort_rows(*oracle.csr_mul_csr(A[0], B[1], A[2], A[3], A[4], B[2], B[3], B[4]))
        _, st = eng.CSR_MUL_CSR_DEV(dA, dB, download=False)
        assert all(st["num_bin_rows"][b] > 0 for b in bins), st["num_bin_rows"]
        ka = eng.CSRtoCOO(dA)
        kb = ka if B is A else eng.CSRtoCOO(dB)
        c, _ = eng.COO_MUL_COO_DEV(ka, kb)
        got = eng.download_coo(c)
        assert np.array_equal(got["row_offset"], want[0]) and np.array_equal(got["col_ind"], want[1])
        assert np.allclose(got["values"], want[2], rtol=1e-12, atol=0)
        eng.free_coo(c); eng.free_coo(ka)
        if B is not A:
            eng.free_coo(kb)
        ea = eng.CSRtoELL(dA, gate=1e9)                  # the 20x gate would refuse a power-law operand
        eb = ea if B is A else eng.CSRtoELL(dB, gate=1e9)
        assert ea.choice and eb.choice
        c, _ = eng.ELL_MUL_ELL_DEV(ea, eb)
        got = eng.download_ell(c)
        assert np.array_equal(got["nnz_row"], np.diff(want[0]))
        for i in range(0, A[0], 37):
            n = int(got["nnz_row"][i]); s0 = int(want[0][i])
            assert np.array_equal(got["col_ind"][i, :n], want[1][s0:s0 + n])
            assert np.allclose(got["values"][i, :n], want[2][s0:s0 + n], rtol=1e-12, atol=0)
        eng.free_ell(c); eng.free_ell(ea)
        if B is not A:
            eng.free_ell(eb); dB.close()
        dA.close()
    eng.set_option("gwin_takes_b2", 1)


# ---------------------------------------------------------------- the front end's path in one call (ias_spgemm_auto_host)
def test_auto_host_picks_the_format_and_matches_the_oracle(eng, oracle):
    """features -> selection -> conversion -> multiply -> host result.  Banded operands come back as DIA, fixed-width
    operands as ELL, everything else as CSR; each equals the reference kernel of that format on the same input."""
    # DIA
    A = W.poisson2d(48)
    r = eng.spgemm_auto(A, A)
    assert r["format"] == "dia" and r["num_diagonals"] == 13
    wa = oracle.csr_to_dia(*A, gate=20.0)
    want = oracle.dia_mul_dia(wa, wa)
    assert np.array_equal(r["diagonal_offsets"], want["diagonal_offsets"]) and np.array_equal(r["diagonal_ind"], want["diagonal_ind"])
    absa = dict(wa, values=np.abs(wa["values"]))
    assert _close(r["values"], want["values"], scale=oracle.dia_mul_dia(absa, absa)["values"])
    assert np.allclose(r["features"], oracle.features26(A, A, gate=20.0), rtol=1e-12)
    assert r["h2d_bytes"] == 4 * (A[0] + 1) + 12 * len(A[3]) and r["d2h_bytes"] >= 8 * A[0] * 13
    # ELL
    U = W.uniform_rows(5000, 8, seed=3)
    r = eng.spgemm_auto(U, U)
    assert r["format"] == "ell"
    we = oracle.csr_to_ell(*U, gate=20.0)
    want = oracle.ell_mul_ell(we, we)
    assert r["width"] == want["width"] and np.array_equal(r["nnz_row"], want["nnz_row"])
    for i in range(0, U[0], 97):
        n = int(want["nnz_row"][i])
        o = np.argsort(want["col_ind"][i, :n], kind="stable")
        assert np.array_equal(r["col_ind"][i, :n], want["col_ind"][i, :n][o])
        assert np.allclose(r["values"][i, :n], want["values"][i, :n][o], rtol=1e-12, atol=0)
        assert not r["values"][i, n:].any() and not r["col_ind"][i, n:].any()
    # CSR (two different operands)
    X, Y = W.random_sparse(300, 200, 0.05, seed=21), W.random_sparse(200, 500, 0.04, seed=22, sort_columns=False)
    r = eng.spgemm_auto(X, Y)
    assert r["format"] == "csr"
    want = oracle.csr_mul_csr(X[0], Y[1], X[2], X[3], X[4], Y[2], Y[3], Y[4])
    assert_csr_parity((r["row_ptr"].copy(), r["col_ind"].copy(), r["values"].copy()), want, mag=abs_product(oracle, X, Y))
    eng.lib.ias_release_host()


def test_select_format_rule(eng):
    f = np.zeros(26)
    f[0], f[2], f[18], f[24], f[8] = 1000, 4900, 5, 0.98, 0.01
    assert eng.select_format(f, True, True) == 2 and eng.select_format(f, False, True) == 3 and eng.select_format(f, False, False) == 1
    f[18] = 400
    assert eng.select_format(f, True, True) == 3
    f[8] = 0.5
    assert eng.select_format(f, True, True) == 1


def test_dia_row_blocks_concatenate(eng):
    """ias_dia_mul_dia_rows_dev (multi-GPU row blocks of the DIA path): the blocks stack up to the full product."""
    A = W.banded(1500, [-40, -1, 0, 3, 41], seed=8)
    dA = eng.upload(*A)
    d = eng.CSRtoDIA(dA, gate=20.0)
    c, _ = eng.DIA_MUL_DIA_DEV(d, d)
    full = eng.download_dia(c)
    parts = []
    for r0, r1 in ((0, 1), (1, 700), (700, 700), (700, 1499), (1499, 1500)):
        cb, ms = eng.DIA_MUL_DIA_DEV(d, d, rows=(r0, r1))
        assert cb.row == r1 - r0 and cb.num_diagonals == c.num_diagonals
        got = eng.download_dia(cb)
        assert np.array_equal(got["diagonal_offsets"], full["diagonal_offsets"])
        parts.append(got["values"].reshape(r1 - r0, c.num_diagonals))
        eng.free_dia(cb)
    assert np.array_equal(np.concatenate(parts, axis=0), full["values"])
    eng.free_dia(c); eng.free_dia(d); dA.close()


@pytest.mark.parametrize("make", [lambda: W.uniform_rows(6000, 16, seed=2), lambda: W.poisson2d(50), lambda: W.banded(900, [-7, 0, 1, 2, 30, 31], seed=3),
                                  lambda: W.random_sparse(400, 400, 0.01, seed=8, sort_columns=False),
                                  lambda: W.uniform_rows(70000, 3, seed=5)],
                         ids=["uniform16", "poisson", "banded6", "unsorted", "wide_cols"])
def test_ell_onepass_kernel_equals_pipeline_and_oracle(eng, oracle, make):
    """The one-pass ELL x ELL kernel (register sort started from the sorted B rows, no symbolic pass) and the Gustavson
    pipeline on fixed-width rows give the same matrix, and both match ELL_MUL_ELL of the oracle."""
    A = make()
    dA = eng.upload(*A)
    e = eng.CSRtoELL(dA, gate=1e9)
    res = {}
    for onepass in (1, 0):
        eng.set_option("ell_onepass", onepass)
        c, ms = eng.ELL_MUL_ELL_DEV(e, e)
        res[onepass] = (eng.download_ell(c), c.nnz)
        eng.free_ell(c)
    eng.set_option("ell_onepass", 1)
    a, b = res[1][0], res[0][0]
    assert res[1][1] == res[0][1] and a["width"] == b["width"]
    assert np.array_equal(a["nnz_row"], b["nnz_row"]) and np.array_equal(a["col_ind"], b["col_ind"])
    assert np.allclose(a["values"], b["values"], rtol=1e-12, atol=0)
    we = oracle.csr_to_ell(*A, gate=1e9)
    want = oracle.ell_mul_ell(we, we)
    assert a["width"] == want["width"] and np.array_equal(a["nnz_row"], want["nnz_row"]) and res[1][1] == want["nnz"]
    mag = oracle.ell_mul_ell(dict(we, values=np.abs(we["values"])), dict(we, values=np.abs(we["values"])))["values"]
    for i in range(0, A[0], max(1, A[0] // 400)):
        n = int(want["nnz_row"][i])
        o = np.argsort(want["col_ind"][i, :n], kind="stable")
        assert np.array_equal(a["col_ind"][i, :n], want["col_ind"][i, :n][o])
        assert np.all(np.abs(a["values"][i, :n] - want["values"][i, :n][o]) <= 1e-12 * np.abs(mag[i, :n][o]))
        assert not a["values"][i, n:].any() and not a["col_ind"][i, n:].any()
    eng.free_ell(e); dA.close()


@pytest.mark.parametrize("make,rows", [(lambda: W.poisson2d(64), None), (lambda: W.banded(2000, [-301, -2, -1, 0, 1, 7, 300], seed=9), (100, 1700)),
                                        (lambda: W.banded(1001, [-3, 0, 4], seed=2), None)], ids=["poisson", "banded_block", "odd_rows"])
def test_dia_vectorised_kernel_equals_scalar(eng, make, rows):
    """k_dia_mul_dia_v2 (two adjacent rows per thread, 128-bit loads where the diagonal offset keeps them aligned) adds the
    pairs of a row in the same order as the scalar kernel: bit-identical results.  Odd row counts take the scalar kernel."""
    A = make()
    dA = eng.upload(*A)
    d = eng.CSRtoDIA(dA, gate=1e9)
    out = {}
    saved = eng.get_option("dia_vec")
    for vec in (1, 0):
        eng.set_option("dia_vec", vec)
        c, ms = eng.DIA_MUL_DIA_DEV(d, d, rows=rows)
        out[vec] = eng.download_dia(c)
        eng.free_dia(c)
    eng.set_option("dia_vec", saved)
    assert np.array_equal(out[1]["diagonal_offsets"], out[0]["diagonal_offsets"])
    assert np.array_equal(out[1]["values"], out[0]["values"])
    eng.free_dia(d); dA.close()


def test_auto_host_pipelined_banded_square(eng, oracle):
    """ias_spgemm_auto_host on a banded A^2 overlaps upload / DIA multiply / download chunk by chunk, taking the diagonal
    set from the first chunk and verifying it afterwards.  Same result as the plain sequence, bit for bit; an operand
    that breaks the speculation (a stray entry in a later chunk, a band wider than a chunk) falls back and is still right."""
    A = W.poisson2d(300)                                   # 90 000 rows: above the pipelining threshold
    r1 = eng.spgemm_auto(A, A)
    assert r1["format"] == "dia" and r1["pipelined"] and r1["num_diagonals"] == 13
    v1, o1, i1 = r1["values"].copy(), r1["diagonal_offsets"].copy(), r1["diagonal_ind"].copy()
    eng.set_option("e2e_pipeline", 0)
    r0 = eng.spgemm_auto(A, A)
    eng.set_option("e2e_pipeline", 1)
    assert r0["format"] == "dia" and not r0["pipelined"]
    assert np.array_equal(v1, r0["values"]) and np.array_equal(o1, r0["diagonal_offsets"]) and np.array_equal(i1, r0["diagonal_ind"])
    wa = oracle.csr_to_dia(*A, gate=20.0)
    want = oracle.dia_mul_dia(wa, wa)
    assert np.array_equal(o1, want["diagonal_offsets"]) and np.array_equal(i1, want["diagonal_ind"])
    absa = dict(wa, values=np.abs(wa["values"]))
    assert _close(v1, want["values"], scale=oracle.dia_mul_dia(absa, absa)["values"])
    assert np.allclose(r1["features"], oracle.features26(A, A, gate=20.0), rtol=1e-12)
    # a stray entry far from the band in the last rows: not in the first chunk's diagonal set -> verified, refused, redone
    rows, cols, rp, ci, v = A
    rp2 = rp.copy(); rp2[-1] += 1
    ci2 = np.concatenate((ci[:rp[-2]], [7], ci[rp[-2]:])).astype(np.int32)       # row rows-1 gets column 7 in front (sorted)
    v2 = np.concatenate((v[:rp[-2]], [0.25], v[rp[-2]:]))
    A2 = (rows, cols, rp2, ci2, v2)
    r2 = eng.spgemm_auto(A2, A2)
    assert not r2["pipelined"]
    want2 = oracle.csr_mul_csr(rows, cols, rp2, ci2, v2, rp2, ci2, v2)
    if r2["format"] == "csr":
        assert_csr_parity((r2["row_ptr"].copy(), r2["col_ind"].copy(), r2["values"].copy()), want2, mag=abs_product(oracle, A2, A2))
    else:
        w2 = oracle.csr_to_dia(*A2, gate=20.0)
        d2 = oracle.dia_mul_dia(w2, w2)
        assert r2["format"] == "dia" and np.array_equal(r2["diagonal_offsets"], d2["diagonal_offsets"])
        ab2 = dict(w2, values=np.abs(w2["values"]))
        assert _close(r2["values"], d2["values"], scale=oracle.dia_mul_dia(ab2, ab2)["values"])
    # a band wider than a chunk: plausible first chunk, but block k would need rows that are not converted yet
    Wd = W.banded(70000, [-40000, 0, 3], seed=4)
    r3 = eng.spgemm_auto(Wd, Wd)
    assert not r3["pipelined"]
    want3 = oracle.csr_mul_csr(Wd[0], Wd[1], Wd[2], Wd[3], Wd[4], Wd[2], Wd[3], Wd[4])
    if r3["format"] == "csr":
        assert_csr_parity((r3["row_ptr"].copy(), r3["col_ind"].copy(), r3["values"].copy()), want3)
    else:
        assert r3["format"] == "dia" and float(r3["values"].sum()) == pytest.approx(float(want3[2].sum()), rel=1e-11)
    eng.lib.ias_release_host()
