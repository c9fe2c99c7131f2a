"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/iaspgemm.h declares, has reference-compatible struct layouts, refuses to run without a GPU
(no CPU fallback), and its host-only pieces (Matrix-Market loader, sizeof formulas) match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ia_spgemm_b200 import engine as E
from util import RECT, SQUARE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(E.LIB_PATH):
        g.build()
    return E.load_library()


def _declared():
    text = open(os.path.join(ROOT, "include", "iaspgemm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ias_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_mirror_agree():
    assert _declared() == sorted(E.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in _declared() if not hasattr(lib, s)]
    assert not missing, missing


def test_every_entry_point_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "iaspgemm.h")).read()
    for needle in ("csr_dev/common_csr_dev.h:134", "dia_dev/common_dia_dev.h:138", "ell_dev/common_ell_dev.h:310",
                   "coo_dev/common_coo_dev.h:279", "common_csr.h:290", "common_csr.h:257", "main.cpp:516", "main.cpp:143"):
        assert needle in text, needle


def test_struct_layouts_match_reference_format_h():
    # bool choice; int row, col, nnz; three pointers  ->  4 + 12 (+0 pad) + 24 on LP64
    for S in (E.CsrMatrix, E.CsrMatrixDev):
        assert C.sizeof(S) == 40
        assert (S.choice.offset, S.row.offset, S.col.offset, S.nnz.offset) == (0, 4, 8, 12)
    assert E.CsrMatrix.row_ind.offset == 16 and E.CsrMatrix.values.offset == 32
    assert E.DiaDev.num_diagonals.offset == 12 and E.DiaDev.values_dev.offset == 32
    # CooMatrixDev: bool, 3 int, 4 pointers; EllMatrixDev: bool, 4 int, (pad), 3 pointers (GPU/detail/format.h:29-40,108-119)
    assert C.sizeof(E.CooDev) == 48 and E.CooDev.nnz.offset == 12 and E.CooDev.row_offset_dev.offset == 16 and E.CooDev.values_dev.offset == 40
    assert C.sizeof(E.EllDev) == 48 and E.EllDev.max_nnz_per_row.offset == 16 and E.EllDev.nnz_row_dev.offset == 24 and E.EllDev.values_dev.offset == 40
    assert C.sizeof(E.SpgemmStats) == 8 * 8 + 16 * 8 + 16 * 8 + 2 * 4 + 8 + 8
    assert C.sizeof(E.StreamBatch) == 4 * 4 + 3 * 8 + 4 * 8


REF = "/root/reference"
CXX = os.path.join(ROOT, "tests", "abi_cxx")


def test_layouts_against_the_reference_headers(tmp_path):
    """Compiles tests/abi_cxx/layout_check.cpp: the reference's own GPU/detail/format.h and CPU/detail/format.h
    (included from where they lie) against include/iaspgemm.h, sizeof/offsetof static_asserts for CSR, COO, DIA, ELL."""
    import subprocess
    gpu_h = os.path.join(REF, "IA-SPGEMM-GPU_release", "detail", "format.h")
    cpu_h = os.path.join(REF, "IA-SPGEMM-CPU_release", "detail", "format.h")
    if not (os.path.exists(gpu_h) and os.path.exists(cpu_h)):
        pytest.skip("reference checkout absent (GPU box): the layouts were proven where it is present")
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-DVALUE_TYPE=double",
                        "-I", os.path.join(CXX, "cusp_stub"), "-I", os.path.join(ROOT, "oracle", "ref_shim"),
                        '-DREF_GPU_FORMAT_H="%s"' % gpu_h, '-DREF_CPU_FORMAT_H="%s"' % cpu_h,
                        os.path.join(CXX, "layout_check.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # and the check bites: a struct with a widened field must be refused
    bad = tmp_path / "bad.cpp"
    src = open(os.path.join(CXX, "layout_check.cpp")).read().replace('#include "../../include/iaspgemm.h"',
                                                                     '#include "%s"' % (tmp_path / "iaspgemm_bad.h"))
    hdr = open(os.path.join(ROOT, "include", "iaspgemm.h")).read()
    marker = "    int row, col, nnz;\n    int *row_offset_dev;"
    assert marker in hdr
    (tmp_path / "iaspgemm_bad.h").write_text(hdr.replace(marker, "    int row, col; long long nnz;\n    int *row_offset_dev;"))
    bad.write_text(src)
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-DVALUE_TYPE=double",
                        "-I", os.path.join(CXX, "cusp_stub"), "-I", os.path.join(ROOT, "oracle", "ref_shim"),
                        '-DREF_GPU_FORMAT_H="%s"' % gpu_h, '-DREF_CPU_FORMAT_H="%s"' % cpu_h, str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "CooMatrixDev" in r.stderr


def test_cxx_caller_compiles_and_links(lib, tmp_path):
    """tests/abi_cxx/abi_driver.cpp (a compiled C++ caller, no ctypes) builds against the header and the library;
    it runs on the GPU box (tests/test_abi_cxx_gpu.py)."""
    import subprocess
    out = tmp_path / "abi_driver"
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-o", str(out), os.path.join(CXX, "abi_driver.cpp"),
                        "-L", os.path.dirname(E.LIB_PATH), "-liaspgemm", "-Wl,-rpath," + os.path.dirname(E.LIB_PATH)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(out)], capture_output=True, text=True)
    import torch
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "no CPU fallback" in r.stdout       # loud, not silent


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.ias_init(0) == 1          # IAS_E_CUDA
    assert b"no CPU fallback" in lib.ias_last_error()
    with pytest.raises(E.EngineError):
        E.Engine()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "ia_spgemm_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libiaoracle" not in text and "oracle." not in text.replace("the oracle.", ""), f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_sizeof_formulas(lib, oracle):
    for rows, nnz, cols, nd, w in ((4, 9, 4, 3, 3), (16777216, 218021892, 16777216, 13, 13), (8000000, 2048000000, 8000000, 5, 256)):
        assert lib.ias_sizeof_csr(rows, nnz) == oracle.sizeof_csr(rows, nnz)
        assert lib.ias_sizeof_coo(rows, nnz) == oracle.sizeof_coo(rows, nnz)
        assert lib.ias_sizeof_dia(rows, cols, nd) == oracle.sizeof_dia(rows, cols, nd)
        assert lib.ias_sizeof_ell(rows, w) == oracle.sizeof_ell(rows, w)
    assert lib.ias_sizeof_csr(4, 9) == 140.0 and lib.ias_sizeof_csr(4, 10) == 152.0      # screenshot values


def _mtx_load(lib, path):
    h = E.CsrMatrix()
    rc = lib.ias_mtx_load(path.encode(), C.byref(h))
    if rc != 0:
        return rc
    out = (h.row, h.col, np.ctypeslib.as_array(h.row_ind, shape=(h.row + 1,)).copy(),
           np.ctypeslib.as_array(h.col_ind, shape=(max(h.nnz, 1),))[: h.nnz].copy(),
           np.ctypeslib.as_array(h.values, shape=(max(h.nnz, 1),))[: h.nnz].copy())
    lib.ias_free_host_csr(C.byref(h))
    return out


@pytest.mark.parametrize("name", SQUARE + RECT)
def test_loader_matches_reference_dump(lib, golden, mtx_dir, name):
    g = golden["inputs"][name]
    rows, cols, rp, ci, v = _mtx_load(lib, os.path.join(mtx_dir, name + ".mtx"))
    assert (rows, cols) == (g["rows"], g["cols"])
    assert rp.tolist() == g["loader_row_ptr"]
    assert ci.tolist() == g["loader_col_ind"]
    assert v.tolist() == g["loader_values"]


def test_loader_error_codes(lib, tmp_path):
    assert _mtx_load(lib, str(tmp_path / "missing.mtx")) == -1
    p = tmp_path / "nobanner.mtx"; p.write_text("hello world\n")
    assert _mtx_load(lib, str(p)) == -2
    p = tmp_path / "cplx.mtx"; p.write_text("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0.0\n")
    assert _mtx_load(lib, str(p)) == -3
    p = tmp_path / "nosize.mtx"; p.write_text("%%MatrixMarket matrix coordinate real general\n% only comments\n")
    assert _mtx_load(lib, str(p)) == -4
    p = tmp_path / "skew.mtx"; p.write_text("%%MatrixMarket matrix coordinate real skew-symmetric\n3 3 2\n2 1 5.0\n3 2 -1.5\n")
    rows, cols, rp, ci, v = _mtx_load(lib, str(p))
    assert rp.tolist() == [0, 0, 1, 2] and ci.tolist() == [0, 1] and v.tolist() == [5.0, -1.5]     # not mirrored


def test_loader_agrees_with_oracle_on_random_files(lib, oracle, tmp_path):
    rng = np.random.default_rng(0)
    for t, (field, symm) in enumerate([("real", "general"), ("integer", "symmetric"), ("pattern", "hermitian"), ("real", "symmetric")]):
        n = 17
        ent = {(int(i), int(j)) for i, j in rng.integers(1, n + 1, size=(60, 2)) if symm == "general" or i >= j}
        lines = ["%%MatrixMarket matrix coordinate " + field + " " + symm, "% c", "%d %d %d" % (n, n, len(ent))]
        for i, j in ent:
            val = "" if field == "pattern" else (" %d" % rng.integers(-9, 9) if field == "integer" else " %.17g" % rng.normal())
            lines.append("%d %d%s" % (i, j, val))
        p = tmp_path / ("m%d.mtx" % t)
        p.write_text("\n".join(lines) + "\n")
        got, want = _mtx_load(lib, str(p)), oracle.mtx_load(str(p))
        assert got[:2] == want[:2]
        for a, b in zip(got[2:], want[2:]):
            assert np.array_equal(a, b)


def test_options_round_trip_without_a_device(lib):
    """ias_set_option / ias_get_option are host-only bookkeeping: every documented knob exists, unknown names and
    negative values are refused (status 2 = IAS_E_ARG), and nothing needs a GPU."""
    names = ["global_rows_smem", "gwin_swords", "gwin_win", "gwin_sym_swords", "gwin_smem_kb", "gwin_max_sw", "g_win",
             "g_coop", "gwin_takes_b2", "trust_operand_cache", "ell_onepass", "g_block", "g_ldca", "g_v2", "g_tbl", "g_lpt", "bulk_store", "dia_vec", "g_scr", "g2_takes_b2", "e2e_pipeline",
             "block_cache", "g_split", "g_split_ub", "g_split_parts"]
    header = open(os.path.join(ROOT, "include", "iaspgemm.h")).read()
    for n in names:
        assert '"%s"' % n in header, n            # documented where the entry point is declared
        old = C.c_longlong(-1)
        assert lib.ias_get_option(n.encode(), C.byref(old)) == 0 and old.value >= 0
        assert lib.ias_set_option(n.encode(), 7) == 0
        v = C.c_longlong(-1)
        assert lib.ias_get_option(n.encode(), C.byref(v)) == 0 and v.value == 7
        assert lib.ias_set_option(n.encode(), old.value) == 0
    assert lib.ias_set_option(b"no_such_knob", 1) == 2
    assert b"no_such_knob" in lib.ias_last_error()
    assert lib.ias_set_option(b"gwin_win", -5) == 2
    assert b"negative" in lib.ias_last_error()
