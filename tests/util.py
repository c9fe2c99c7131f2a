"""Helpers shared by the parity tests."""
import base64
import zlib

import numpy as np

SQUARE = ["dia", "small", "b1_ss", "LFAT5", "Ragusa18"]
RECT = ["Trec5", "ch3-3-b2", "relat3", "sample"]
RTOL = 1e-12     # BASELINE.json north_star: values within 1e-12 relative per entry


def decode_img(s):
    return np.frombuffer(zlib.decompress(base64.b64decode(s)), dtype="<i4").astype(np.int64).reshape(128, 128)


def sort_rows(rp, ci, v):
    """Column-sort every row of a CSR triple (the reference emits unsorted rows)."""
    rp = np.asarray(rp, dtype=np.int64)
    ci = np.asarray(ci)
    v = np.asarray(v)
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    order = np.lexsort((ci, rows))
    return rp, ci[order], v[order]


def assert_csr_parity(got, want, mag=None, rtol=RTOL):
    """got = engine (rp, ci, v) -- must already be column sorted; want = oracle triple (any order).

    Structure bit-exact; values within rtol of the entry's magnitude.  `mag` (same layout as
    want after sorting) bounds |sum of |a*b||, which is the right scale when terms cancel; if
    omitted the entry itself is the scale.
    """
    g_rp, g_ci, g_v = got
    w_rp, w_ci, w_v = sort_rows(*want)
    g_rp = np.asarray(g_rp, dtype=np.int64)
    assert np.array_equal(g_rp, w_rp), "row_ptr differs"
    assert np.array_equal(np.asarray(g_ci), w_ci), "column indices differ"
    # engine rows must be strictly increasing in column
    if len(g_ci) > 1:
        d = np.diff(np.asarray(g_ci, dtype=np.int64))
        starts = g_rp[1:-1]
        inner = np.ones(len(d), bool)
        inner[starts[(starts > 0) & (starts < len(g_ci))] - 1] = False
        assert (d[inner] > 0).all(), "engine rows are not strictly column sorted"
    scale = np.abs(w_v) if mag is None else np.abs(mag)
    err = np.abs(np.asarray(g_v) - w_v)
    bad = err > rtol * scale
    assert not bad.any(), "max rel err %.3e at %d entries" % (float((err[bad] / np.maximum(scale[bad], 1e-300)).max()), int(bad.sum()))


def abs_product(oracle, A, B):
    """|A|*|B| through the oracle: per-entry magnitude bound for cancellation-prone inputs."""
    rp, ci, v = oracle.csr_mul_csr(A[0], B[1], A[2], A[3], np.abs(A[4]), B[2], B[3], np.abs(B[4]))
    return sort_rows(rp, ci, v)[2]
