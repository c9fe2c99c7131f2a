"""Pin the CPU oracle (oracle/ia_oracle.c) before anything trusts it.

Three anchors:
  1. committed golden fixtures produced by the reference itself (tests/golden/make_golden.py):
     loader dump + density image from the reference's unmodified main.cpp, kernel outputs from
     the reference's unmodified headers;
  2. the known-answer files the reference ships (imgs/img{1,2}.txt) and the feature vector in
     its screenshots;
  3. live comparison with oracle/_ref/libiaref.so on seeded random inputs (skipped when the
     reference build is not present).
"""
import os

import numpy as np
import pytest

from ia_spgemm_b200 import workloads as W
from util import RECT, SQUARE, decode_img, sort_rows


def _load(oracle, mtx_dir, name):
    return oracle.mtx_load(os.path.join(mtx_dir, name + ".mtx"))


@pytest.mark.parametrize("name", SQUARE + RECT)
def test_loader_matches_reference_dump(oracle, golden, mtx_dir, name):
    g = golden["inputs"][name]
    rows, cols, rp, ci, v = _load(oracle, mtx_dir, name)
    assert (rows, cols) == (g["rows"], g["cols"])
    assert rp.tolist() == g["loader_row_ptr"]
    assert ci.tolist() == g["loader_col_ind"]          # file order inside rows, mirrored entries interleaved
    assert v.tolist() == g["loader_values"]


def test_loader_known_counts(oracle, mtx_dir):
    # nnz after load (LFAT5 is symmetric: 30 stored -> 46)
    want = {"dia": 7, "small": 8, "b1_ss": 15, "LFAT5": 46, "Ragusa18": 64}
    for name, nnz in want.items():
        assert int(_load(oracle, mtx_dir, name)[2][-1]) == nnz


def test_loader_errors(oracle, tmp_path):
    with pytest.raises(IOError):
        oracle.mtx_load(str(tmp_path / "missing.mtx"))
    p = tmp_path / "cplx.mtx"
    p.write_text("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0.0\n")
    with pytest.raises(IOError, match="-3"):
        oracle.mtx_load(str(p))
    p = tmp_path / "nobanner.mtx"
    p.write_text("hello world\n")
    with pytest.raises(IOError, match="-2"):
        oracle.mtx_load(str(p))


@pytest.mark.parametrize("name", SQUARE + RECT)
def test_density_matches_reference_main(oracle, golden, mtx_dir, name):
    rows, cols, rp, ci, v = _load(oracle, mtx_dir, name)
    assert np.array_equal(oracle.density(rows, cols, rp, ci), decode_img(golden["inputs"][name]["density"]))


def test_density_matches_shipped_images(oracle, golden, mtx_dir):
    import scipy.sparse as sp
    ship = golden["shipped_imgs"]
    for name, k1, k2 in (("dia", "gpu_img1_dia", "gpu_img2_diaT"), ("small", "cpu_img1_small", "cpu_img2_smallT")):
        rows, cols, rp, ci, v = _load(oracle, mtx_dir, name)
        assert np.array_equal(oracle.density(rows, cols, rp, ci), decode_img(ship[k1]))
        T = sp.csr_matrix((v, ci, rp), shape=(rows, cols)).T.tocsr()
        assert np.array_equal(oracle.density(cols, rows, T.indptr, T.indices), decode_img(ship[k2]))


def test_density_large_dims(oracle):
    # rows > 128: single cell per entry; 64-bit index arithmetic
    n, _, rp, ci, v = W.banded(1000, [-3, 0, 5])
    img = oracle.density(n, n, rp, ci)
    assert img.sum() == rp[-1]
    ri = W.row_index(rp)
    want = np.zeros((128, 128), dtype=np.int64)
    np.add.at(want, (ri * 128 // n, ci.astype(np.int64) * 128 // n), 1)
    assert np.array_equal(img, want)


def test_features_match_screenshot(oracle, golden, mtx_dir):
    A = _load(oracle, mtx_dir, "dia")
    f = oracle.features26(A, A)
    assert f.tolist() == golden["screenshot_features_dia"]


@pytest.mark.parametrize("name", SQUARE)
def test_features_match_reference(oracle, golden, mtx_dir, name):
    A = _load(oracle, mtx_dir, name)
    assert oracle.features26(A, A).tolist() == golden["inputs"][name]["features26"]


@pytest.mark.parametrize("name", SQUARE)
def test_csr_mul_csr_golden(oracle, golden, mtx_dir, name):
    g = golden["inputs"][name]
    rows, cols, rp, ci, v = _load(oracle, mtx_dir, name)
    assert oracle.getflop(rp, ci, rp) == g["flop"]
    c_rp, c_ci, c_v = oracle.csr_mul_csr(rows, cols, rp, ci, v, rp, ci, v)
    assert c_rp.tolist() == g["csr_row_ptr"]
    assert c_ci.tolist() == g["csr_col_ind"]            # same (reverse first-touch) order as the reference
    assert c_v.tolist() == g["csr_values"]              # same accumulation order -> bit identical
    assert oracle.sizeof_csr(rows, int(c_rp[-1])) == g["sizeof_csr_c"]
    # MKL (Algorithm 1) agrees on structure, and on values to rounding
    m = sort_rows(g["mkl_row_ptr"], np.array(g["mkl_col_ind"]), np.array(g["mkl_values"]))
    o = sort_rows(c_rp, c_ci, c_v)
    assert np.array_equal(m[0], o[0]) and np.array_equal(m[1], o[1])
    np.testing.assert_allclose(m[2], o[2], rtol=1e-12, atol=1e-9 if name == "LFAT5" else 1e-15)


def test_known_answers_appendix_b(oracle, mtx_dir):
    # SURVEY.md appendix B
    want = {"dia": (12, 9, 12.0), "small": (15, 9, 15.0), "b1_ss": (33, 30, 3.733128554),
            "LFAT5": (166, 72, 7.895731823e13), "Ragusa18": (251, 172, 422.0)}
    for name, (flop, nnz, total) in want.items():
        rows, cols, rp, ci, v = _load(oracle, mtx_dir, name)
        c_rp, c_ci, c_v = oracle.csr_mul_csr(rows, cols, rp, ci, v, rp, ci, v)
        assert oracle.getflop(rp, ci, rp) == flop
        assert int(c_rp[-1]) == nnz
        assert c_v.sum() == pytest.approx(total, rel=1e-9)
    # b1_ss keeps three numerically-zero entries
    rows, cols, rp, ci, v = _load(oracle, mtx_dir, "b1_ss")
    assert int((oracle.csr_mul_csr(rows, cols, rp, ci, v, rp, ci, v)[2] == 0.0).sum()) == 3


def test_dia_times_dia_transpose_screenshot(oracle, golden, mtx_dir):
    g = golden["dia_times_diaT"]
    rows, cols, rp, ci, v = _load(oracle, mtx_dir, "dia")
    b_rp, b_ci = np.array(g["b_row_ptr"], np.int32), np.array(g["b_col_ind"], np.int32)
    c_rp, _, _ = oracle.csr_mul_csr(rows, rows, rp, ci, v, b_rp, b_ci, np.ones(len(b_ci)))
    assert int(c_rp[-1]) == g["nnz"] == 10
    assert oracle.getflop(rp, ci, b_rp) == g["flop"] == 13
    assert oracle.sizeof_csr(rows, 10) == g["sizeof_csr"] == 152.0


@pytest.mark.parametrize("name", SQUARE)
def test_dia_ell_coo_golden(oracle, golden, mtx_dir, name):
    g = golden["inputs"][name]
    A = _load(oracle, mtx_dir, name)
    rows, cols, rp, ci, v = A
    if "dia_offsets_c" in g:
        d = oracle.csr_to_dia(*A)
        assert d["choice"] and d["diagonal_offsets"].tolist() == g["dia_offsets_a"]
        c = oracle.dia_mul_dia(d, d)
        assert c["diagonal_offsets"].tolist() == g["dia_offsets_c"]
        assert c["diagonal_ind"].tolist() == g["dia_diag_ind_c"]
        assert c["values"].ravel().tolist() == g["dia_values_c"]
    if "ell_width_c" in g:
        e = oracle.csr_to_ell(*A)
        c = oracle.ell_mul_ell(e, e)
        assert c["width"] == g["ell_width_c"]
        assert c["nnz_row"].tolist() == g["ell_nnz_row_c"]
        assert c["col_ind"].ravel().tolist() == g["ell_col_ind_c"]
        assert c["values"].ravel().tolist() == g["ell_values_c"]
    k = oracle.csr_to_coo(rows, rp, ci, v)
    c = oracle.coo_mul_coo(rows, cols, k, k)
    assert c["row_offset"].tolist() == g["coo_row_offset_c"]
    assert c["row_ind"].tolist() == g["coo_row_ind_c"]
    assert c["col_ind"].tolist() == g["coo_col_ind_c"]
    assert c["values"].tolist() == g["coo_values_c"]


# ---- live comparison against the compiled reference on seeded inputs ----------------------------
CASES = [
    ("poisson32", lambda: W.poisson2d(32)),
    ("uniform", lambda: W.uniform_rows(3000, 8, seed=3)),
    ("rmat10", lambda: W.rmat(10, 8, seed=2)),
    ("irregular", lambda: W.random_sparse(257, 257, 0.03, seed=5)),
    ("unsorted", lambda: W.random_sparse(120, 120, 0.08, seed=6, sort_columns=False)),
    ("banded", lambda: W.banded(500, [-7, -1, 0, 2, 9], seed=4)),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_reference_build(oracle, ref, name, make):
    A = make()
    rows, cols, rp, ci, v = A
    r_rp, r_ci, r_v, _ = ref.csr_mul_csr(A, A)
    o_rp, o_ci, o_v = oracle.csr_mul_csr(rows, cols, rp, ci, v, rp, ci, v)
    assert np.array_equal(r_rp, o_rp) and np.array_equal(r_ci, o_ci) and np.array_equal(r_v, o_v)
    assert ref.getflop(A, A) == oracle.getflop(rp, ci, rp)
    assert np.array_equal(ref.features26(A, A), oracle.features26(A, A))
    # MKL agrees after sorting
    m_rp, m_ci, m_v, _, _ = ref.mkl_mul_mkl(A, A)
    m, o = sort_rows(m_rp, m_ci, m_v), sort_rows(o_rp, o_ci, o_v)
    assert np.array_equal(m[0], o[0]) and np.array_equal(m[1], o[1])
    np.testing.assert_allclose(m[2], o[2], rtol=1e-12)
    # DIA / ELL / COO
    rd = ref.dia_mul_dia(A, A)
    od = oracle.csr_to_dia(*A)
    assert (rd is not None) == od["choice"]
    if rd is not None:
        oc = oracle.dia_mul_dia(od, od)
        assert np.array_equal(rd["diagonal_offsets"], oc["diagonal_offsets"])
        assert np.array_equal(rd["diagonal_ind"], oc["diagonal_ind"])
        assert np.array_equal(rd["values"], oc["values"])
    re_ = ref.ell_mul_ell(A, A)
    oe = oracle.csr_to_ell(*A)
    assert (re_ is not None) == oe["choice"]
    if re_ is not None:
        oc = oracle.ell_mul_ell(oe, oe)
        assert re_["width"] == oc["width"] and re_["nnz"] == oc["nnz"]
        assert np.array_equal(re_["nnz_row"], oc["nnz_row"])
        assert np.array_equal(re_["col_ind"], oc["col_ind"]) and np.array_equal(re_["values"], oc["values"])
    if rows <= 600:
        rc = ref.coo_mul_coo(A, A)
        ok = oracle.csr_to_coo(rows, rp, ci, v)
        oc = oracle.coo_mul_coo(rows, cols, ok, ok)
        for key in ("row_offset", "row_ind", "col_ind", "values"):
            assert np.array_equal(rc[key], oc[key]), key


def test_reference_converters(oracle, ref):
    A = W.banded(300, [-2, 0, 1], seed=9)
    rd, od = ref.csr_to_dia(A), oracle.csr_to_dia(*A)
    for key in ("diagonal_ind", "diagonal_offsets", "values"):
        assert np.array_equal(rd[key], od[key]), key
    re_, oe = ref.csr_to_ell(A), oracle.csr_to_ell(*A)
    for key in ("nnz_row", "col_ind", "values"):
        assert np.array_equal(re_[key], oe[key]), key
    # the 50x gate: a scattered matrix is refused as DIA by both
    S = W.uniform_rows(4000, 2, seed=1)
    assert ref.csr_to_dia(S)["choice"] is False and oracle.csr_to_dia(*S)["choice"] is False


def test_poisson_closed_forms(oracle):
    N = 48
    rows, cols, rp, ci, v = W.poisson2d(N)
    nnz, products, nnz_c = W.poisson_counts(N)
    assert int(rp[-1]) == nnz
    assert oracle.getflop(rp, ci, rp) == products
    assert int(oracle.csr_mul_csr(rows, cols, rp, ci, v, rp, ci, v)[0][-1]) == nnz_c
