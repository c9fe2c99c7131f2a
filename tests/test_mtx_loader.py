"""ias_mtx_load parses the entries from memory (several threads on regular files, a sequential tokenizer otherwise)
instead of the reference's fscanf loop (CPU/main.cpp:143-458).  Same result as that loop -- kept in the library as
IAS_MTX_LOADER=fscanf -- and as the oracle's restatement of the reference loader, on regular files, on files whose
entries wander over line ends, and on files that stop converting half way."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ia_spgemm_b200 import engine as E      # noqa: E402


@pytest.fixture(scope="module")
def lib():
    return E.load_library()


def _load(lib, path, mode=None):
    old = os.environ.pop("IAS_MTX_LOADER", None)
    if mode:
        os.environ["IAS_MTX_LOADER"] = mode
    try:
        h = E.CsrMatrix()
        rc = lib.ias_mtx_load(str(path).encode(), C.byref(h))
        if rc != 0:
            return rc
        out = (h.row, h.col, np.ctypeslib.as_array(h.row_ind, shape=(h.row + 1,)).copy(),
               np.ctypeslib.as_array(h.col_ind, shape=(max(h.nnz, 1),))[: h.nnz].copy(),
               np.ctypeslib.as_array(h.values, shape=(max(h.nnz, 1),))[: h.nnz].copy())
        lib.ias_free_host_csr(C.byref(h))
        return out
    finally:
        os.environ.pop("IAS_MTX_LOADER", None)
        if old is not None:
            os.environ["IAS_MTX_LOADER"] = old


def _same(a, b):
    if isinstance(a, int) or isinstance(b, int):
        return a == b
    return a[:2] == b[:2] and all(np.array_equal(x, y) for x, y in zip(a[2:4], b[2:4])) and \
        np.array_equal(a[4].view(np.uint64), b[4].view(np.uint64))          # values bit for bit (NaNs included)


HEAD = "%%MatrixMarket matrix coordinate {field} {symm}\n% a comment\n{m} {n} {nz}\n"

CASES = {
    # regular
    "plain": ("real", "general", 4, 4, 3, "1 1 1.5\n2 3 -2.25\n4 4 1e-3\n"),
    "no_final_newline": ("real", "general", 4, 4, 2, "1 1 1.5\n2 3 -2.25"),
    "crlf_tabs_blank_lines": ("real", "general", 4, 4, 3, "1\t1\t1.5\r\n\r\n  2 3   -2.25  \r\n\n4 4 7\n"),
    "more_entries_than_declared": ("real", "general", 4, 4, 2, "1 1 1\n2 2 2\n3 3 3\n4 4 4\n"),
    "fewer_entries_than_declared": ("real", "general", 4, 4, 9, "1 1 1\n2 2 2\n"),
    "out_of_range_entries_are_skipped_but_counted": ("real", "general", 3, 3, 3, "0 1 5\n1 4 6\n2 2 7\n3 3 8\n"),
    "integer_field": ("integer", "general", 4, 4, 3, "1 2 -7\n2 2 +3\n4 1 0\n"),
    "pattern_symmetric": ("pattern", "symmetric", 4, 4, 3, "2 1\n3 3\n4 2\n"),
    "real_hermitian": ("real", "hermitian", 3, 3, 2, "2 1 0.5\n3 3 9\n"),
    "rectangular_symmetric_mirror_is_bounded": ("real", "symmetric", 2, 5, 2, "1 4 1.0\n2 1 2.0\n"),
    "spellings_of_reals": ("real", "general", 9, 9, 9, "1 1 +1.5\n2 2 .5\n3 3 5.\n4 4 1E+2\n5 5 0x1p3\n6 6 inf\n7 7 -nan\n8 8 1e-400\n9 9 1e400\n"),
    "seventeen_digits": ("real", "general", 3, 3, 3, "1 1 0.10000000000000001\n2 2 1.7976931348623157e308\n3 3 4.9406564584124654e-324\n"),
    "signed_indices": ("real", "general", 3, 3, 2, "+1 +2 1\n-1 2 2\n"),
    # irregular: the tokenizer takes over
    "entry_over_two_lines": ("real", "general", 4, 4, 3, "1 1\n1.5\n2\n3 -2.25 4 4\n7\n"),
    "two_entries_on_a_line": ("real", "general", 4, 4, 3, "1 1 1.5 2 3 -2.25\n4 4 7\n"),
    "value_missing_takes_the_next_token": ("real", "general", 4, 4, 2, "1 1\n2 3 4\n5 6 7\n"),
    "comment_inside_the_data_stops_the_read": ("real", "general", 4, 4, 3, "1 1 1.5\n% not allowed here\n2 3 -2.25\n"),
    "garbage_stops_the_read": ("real", "general", 4, 4, 3, "1 1 1.5\n2 x 3\n4 4 7\n"),
    "float_where_an_index_belongs": ("real", "general", 4, 4, 3, "1 1 1.5\n2.5 3 1\n4 4 7\n"),
    "trailing_junk_after_a_value": ("real", "general", 4, 4, 3, "1 1 1.5abc\n2 3 1\n"),
    "form_feed_is_white_space": ("real", "general", 4, 4, 2, "1 1 1.5\f2 3 1\n"),
    "pattern_with_values_present": ("pattern", "general", 4, 4, 2, "1 1 9\n2 2 9\n"),
    "integer_field_with_reals": ("integer", "general", 4, 4, 2, "1 1 2.5\n2 2 3\n"),
    "huge_index": ("real", "general", 4, 4, 2, "99999999999 1 1\n2 2 2\n"),
    "empty_body": ("real", "general", 4, 4, 0, ""),
    "declared_but_empty": ("real", "general", 4, 4, 5, "\n\n"),
}


# files with exactly the declared number of entries, all inside the declared shape (the reference loop neither checks
# what fscanf returns nor the indices: CPU/main.cpp:417-433)
WELL_FORMED = {"plain", "no_final_newline", "crlf_tabs_blank_lines", "integer_field", "pattern_symmetric", "real_hermitian",
               "spellings_of_reals", "seventeen_digits", "entry_over_two_lines", "two_entries_on_a_line", "form_feed_is_white_space",
               "empty_body"}


@pytest.mark.parametrize("name", sorted(CASES))
def test_same_as_the_fscanf_loop_and_the_oracle(lib, oracle, tmp_path, name):
    field, symm, m, n, nz, body = CASES[name]
    p = tmp_path / (name + ".mtx")
    p.write_bytes((HEAD.format(field=field, symm=symm, m=m, n=n, nz=nz) + body).encode())
    new, old = _load(lib, p), _load(lib, p, "fscanf")
    assert _same(new, old), (new, old)
    if name in WELL_FORMED:               # the oracle restates the reference loop, which is only defined on such files
        want = oracle.mtx_load(str(p))
        assert new[:2] == want[:2] and np.array_equal(new[2], want[2]) and np.array_equal(new[3], want[3])
        assert np.array_equal(new[4].view(np.uint64), np.asarray(want[4], dtype=np.float64).view(np.uint64))


def test_known_answers(lib, tmp_path):
    p = tmp_path / "a.mtx"
    p.write_text(HEAD.format(field="real", symm="general", m=9, n=9, nz=9) +
                 "1 1 +1.5\n2 2 .5\n3 3 5.\n4 4 1E+2\n5 5 0x1p3\n6 6 inf\n7 7 0.10000000000000001\n8 8 1e-400\n9 9 1e400\n")
    rows, cols, rp, ci, v = _load(lib, p)
    assert rp.tolist() == list(range(10)) and ci.tolist() == list(range(9))
    assert v.tolist() == [1.5, 0.5, 5.0, 100.0, 8.0, float("inf"), 0.1, 0.0, float("inf")]
    p.write_text(HEAD.format(field="real", symm="general", m=4, n=4, nz=3) + "1 1\n1.5\n2\n3 -2.25 4 4\n7\n")
    rows, cols, rp, ci, v = _load(lib, p)
    assert rp.tolist() == [0, 1, 2, 2, 3] and ci.tolist() == [0, 2, 3] and v.tolist() == [1.5, -2.25, 7.0]


@pytest.mark.parametrize("field,symm", [("real", "general"), ("integer", "symmetric"), ("pattern", "general")])
def test_large_regular_file_uses_every_thread_and_keeps_file_order(lib, tmp_path, field, symm):
    """6 MB and more of text: the file is cut into pieces at line ends and parsed by several threads; rows are filled in
    file order (columns unsorted, duplicates kept), exactly as the sequential loop fills them."""
    rng = np.random.default_rng(7)
    n, nz = 5000, 600000
    i, j = rng.integers(1, n + 1, nz), rng.integers(1, n + 1, nz)
    if symm == "symmetric":
        i, j = np.maximum(i, j), np.minimum(i, j)
    if field == "real":
        vals = ["%.17g" % x for x in rng.normal(size=nz)]
    elif field == "integer":
        vals = ["%d" % x for x in rng.integers(-99, 99, nz)]
    else:
        vals = [""] * nz
    p = tmp_path / "big.mtx"
    with open(p, "w") as f:
        f.write(HEAD.format(field=field, symm=symm, m=n, n=n, nz=nz - 7))          # the last 7 entries are beyond the declared count
        f.write("\n".join(("%d %d %s" % t).rstrip() for t in zip(i, j, vals)) + "\n")
    assert os.path.getsize(p) > 4 << 20
    new, old = _load(lib, p), _load(lib, p, "fscanf")
    assert _same(new, old)
    assert new[2][-1] == (nz - 7) * (1 if symm == "general" else 2) - (0 if symm == "general" else int(np.sum(i[:nz - 7] == j[:nz - 7])))
    # an irregular line at the end of the LAST piece: every piece is thrown away and the tokenizer reads the whole file
    with open(p, "w") as f:
        f.write(HEAD.format(field=field, symm=symm, m=n, n=n, nz=nz + 2))
        f.write("\n".join(("%d %d %s" % t).rstrip() for t in zip(i, j, vals)) + "\n")
        f.write("7 7 1 8 8 1 9\n" if field != "pattern" else "7 7 8 8 9\n")
    new, old = _load(lib, p), _load(lib, p, "fscanf")
    assert _same(new, old)
    assert new[2][-1] == old[2][-1] > 0


_FUZZ = r'''
import ctypes as C, os, random, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
from ia_spgemm_b200 import engine as E
lib = E.load_library()
path, rng = sys.argv[2], random.Random(int(sys.argv[3]))
TOKENS = ["1", "2", "3", "4", "5", "0", "-1", "+2", "99999999999", "1.5", "-2e3", "nan", "inf", "0x1p3", "abc", "%", "1e400", ".", "+", "-",
          "1e", "\n", "\n\n", "\t", " ", "\r\n", "\f", "\v", "\0"]


def load(mode):
    if mode:
        os.environ["IAS_MTX_LOADER"] = mode
    else:
        os.environ.pop("IAS_MTX_LOADER", None)
    h = E.CsrMatrix()
    rc = lib.ias_mtx_load(path.encode(), C.byref(h))
    if rc:
        return rc
    out = (h.row, h.col, h.nnz, np.ctypeslib.as_array(h.row_ind, shape=(h.row + 1,)).tobytes(),
           np.ctypeslib.as_array(h.col_ind, shape=(max(h.nnz, 1),))[:h.nnz].tobytes(),
           np.ctypeslib.as_array(h.values, shape=(max(h.nnz, 1),))[:h.nnz].tobytes())
    lib.ias_free_host_csr(C.byref(h))
    return out


for t in range(int(sys.argv[4])):
    field = rng.choice(["real", "integer", "pattern"])
    symm = rng.choice(["general", "symmetric", "hermitian", "skew-symmetric"])
    m, n, nz = rng.randrange(0, 7), rng.randrange(0, 7), rng.randrange(0, 12)
    body = []
    for e in range(rng.randrange(0, 14)):
        if rng.random() < 0.8:                       # an entry line, indices sometimes out of range
            toks = [str(rng.randrange(0, 8)), str(rng.randrange(0, 8))]
            if field == "real":
                toks.append(rng.choice(["1.5", "-2", "3e1", "7"]))
            elif field == "integer":
                toks.append(str(rng.randrange(-9, 9)))
            body.append(rng.choice([" ", "  ", "\t"]).join(toks) + rng.choice(["\n", "\n", "\r\n", " \n"]))
        else:                                        # anything
            body.append("".join(rng.choice(TOKENS) + rng.choice([" ", "", "\n"]) for _ in range(rng.randrange(1, 6))))
    text = "%%%%MatrixMarket matrix coordinate %s %s\n%% c\n%d %d %d\n" % (field, symm, m, n, nz) + "".join(body)
    open(path, "wb").write(text.encode("latin1"))
    a, b = load(None), load("fscanf")
    if a != b:
        print("MISMATCH", repr(text))
        sys.exit(3)
print("agreed")
'''


@pytest.mark.timeout(300)
def test_random_files_parse_like_the_fscanf_loop(tmp_path):
    """Differential fuzz in a child process (a crash would be seen as one): regular entry lines mixed with arbitrary token
    soup -- signs, hex floats, inf/nan, NUL bytes, form feeds, comments, numbers that overflow -- 600 files."""
    import subprocess
    for seed in (21, 22):
        r = subprocess.run([sys.executable, "-c", _FUZZ, ROOT, str(tmp_path / "fuzz.mtx"), str(seed), "300"],
                           capture_output=True, text=True, timeout=280)
        assert r.returncode == 0, (r.stdout[-800:], r.stderr[-800:])
        assert r.stdout.strip().endswith("agreed")
