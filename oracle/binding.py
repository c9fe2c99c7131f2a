"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Nothing under ia_spgemm_b200/ does.

  Oracle()  -> oracle/libiaoracle.so      (plain-C restatement, oracle/ia_oracle.c)
  Ref()     -> oracle/_ref/libiaref.so    (the reference's own headers, oracle/ref_harness.cpp);
               Ref.available() is False when it has not been built.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_I = C.POINTER(C.c_int)
_L = C.POINTER(C.c_int64)
_D = C.POINTER(C.c_double)


def build(ref=True):
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=False)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


def _take(ptr, n, dtype, free):
    """Copy n elements out of a malloc'd C array and free it."""
    if n <= 0 or not ptr:
        if ptr:
            free(ptr)
        return np.zeros(0, dtype=dtype)
    out = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
    free(ptr)
    return out


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "libiaoracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        L = self.lib
        L.ora_getflop.restype = C.c_longlong
        for n in ("ora_sizeof_csr", "ora_sizeof_dia", "ora_sizeof_ell", "ora_sizeof_coo"):
            getattr(L, n).restype = C.c_double
        L.ora_sizeof_csr.argtypes = [C.c_int, C.c_int64]
        L.ora_sizeof_coo.argtypes = [C.c_int, C.c_int64]
        L.ora_sizeof_dia.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ora_sizeof_ell.argtypes = [C.c_int, C.c_int]
        L.ora_free.argtypes = [C.c_void_p]
        self._free = lambda p: L.ora_free(C.cast(p, C.c_void_p))

    def threads(self):
        return int(self.lib.ora_num_threads())

    def set_threads(self, n):
        self.lib.ora_set_threads(C.c_int(int(n)))

    # -- loader ---------------------------------------------------------------
    def mtx_load(self, path):
        r, c, n = C.c_int(), C.c_int(), C.c_int()
        rp, ci, v = _I(), _I(), _D()
        rc = self.lib.ora_mtx_load(path.encode(), C.byref(r), C.byref(c), C.byref(n),
                                   C.byref(rp), C.byref(ci), C.byref(v))
        if rc != 0:
            raise IOError("ora_mtx_load(%s) -> %d" % (path, rc))
        return (r.value, c.value,
                _take(rp, r.value + 1, np.int32, self._free),
                _take(ci, n.value, np.int32, self._free),
                _take(v, n.value, np.float64, self._free))

    # -- features -------------------------------------------------------------
    def density(self, rows, cols, rp, ci):
        rp, ci = _i(rp), _i(ci)
        img = np.zeros(128 * 128, dtype=np.int64)
        self.lib.ora_density(C.c_int(rows), C.c_int(cols), _p(rp, _I), _p(ci, _I),
                             img.ctypes.data_as(C.POINTER(C.c_longlong)))
        return img.reshape(128, 128)

    def getinfo1(self, rows, cols, rp):
        rp = _i(rp)
        f = np.zeros(9)
        self.lib.ora_getinfo1(C.c_int(rows), C.c_int(cols), C.c_int(int(rp[-1])), _p(rp, _I), _p(f, _D))
        return f

    def getinfo2(self, rows, cols, nd):
        f = np.zeros(3)
        self.lib.ora_getinfo2(C.c_int(rows), C.c_int(cols), C.c_int(nd), _p(f, _D))
        return f

    def getinfo3(self, rows, nnz, width):
        f = np.zeros(1)
        self.lib.ora_getinfo3(C.c_int(rows), C.c_int(nnz), C.c_int(width), _p(f, _D))
        return f

    def getflop(self, a_rp, a_ci, b_rp):
        a_rp, a_ci, b_rp = _i(a_rp), _i(a_ci), _i(b_rp)
        return int(self.lib.ora_getflop(C.c_int(len(a_rp) - 1), _p(a_rp, _I), _p(a_ci, _I), _p(b_rp, _I)))

    def features26(self, A, B, gate=50.0):
        """A, B = (rows, cols, rp, ci, v).  Order as CPU/main.cpp:655-679."""
        f = np.zeros(26)
        f[0:9] = self.getinfo1(A[0], A[1], A[2])
        f[9:18] = self.getinfo1(B[0], B[1], B[2])
        for M, o in ((A, 18), (B, 21)):
            d = self.csr_to_dia(*M, gate=gate)
            f[o:o + 3] = self.getinfo2(M[0], M[1], d["num_diagonals"])
        for M, o in ((A, 24), (B, 25)):
            e = self.csr_to_ell(*M, gate=gate)
            f[o] = self.getinfo3(M[0], int(M[2][-1]), e["width"])[0]
        return f

    # -- Algorithm 2 ------------------------------------------------------------
    def csr_mul_csr(self, a_rows, b_cols, a_rp, a_ci, a_v, b_rp, b_ci, b_v):
        a_rp, a_ci, a_v = _i(a_rp), _i(a_ci), _d(a_v)
        b_rp, b_ci, b_v = _i(b_rp), _i(b_ci), _d(b_v)
        rp, ci, v = _L(), _I(), _D()
        rc = self.lib.ora_csr_mul_csr(C.c_int(a_rows), C.c_int(b_cols),
                                      _p(a_rp, _I), _p(a_ci, _I), _p(a_v, _D),
                                      _p(b_rp, _I), _p(b_ci, _I), _p(b_v, _D),
                                      C.byref(rp), C.byref(ci), C.byref(v))
        if rc != 0:
            raise MemoryError("ora_csr_mul_csr")
        c_rp = _take(rp, a_rows + 1, np.int64, self._free)
        nnz = int(c_rp[-1])
        return c_rp, _take(ci, nnz, np.int32, self._free), _take(v, nnz, np.float64, self._free)

    # -- DIA --------------------------------------------------------------------
    def csr_to_dia(self, rows, cols, rp, ci, v, gate=50.0):
        rp, ci, v = _i(rp), _i(ci), _d(v)
        ch, nd = C.c_int(), C.c_int()
        di, off, val = _I(), _I(), _D()
        self.lib.ora_csr_to_dia(C.c_int(rows), C.c_int(cols), C.c_int(int(rp[-1])), _p(rp, _I), _p(ci, _I), _p(v, _D),
                                C.c_double(gate), C.byref(ch), C.byref(nd), C.byref(di), C.byref(off), C.byref(val))
        out = {"choice": bool(ch.value), "num_diagonals": nd.value, "row": rows, "col": cols}
        if ch.value:
            out["diagonal_ind"] = _take(di, rows + cols - 1, np.int32, self._free)
            out["diagonal_offsets"] = _take(off, nd.value, np.int32, self._free)
            out["values"] = _take(val, rows * nd.value, np.float64, self._free).reshape(rows, nd.value)
        return out

    def dia_mul_dia(self, A, B):
        a_off, a_val = _i(A["diagonal_offsets"]), _d(A["values"])
        b_off, b_val = _i(B["diagonal_offsets"]), _d(B["values"])
        nd = C.c_int()
        di, off, val = _I(), _I(), _D()
        self.lib.ora_dia_mul_dia(C.c_int(A["row"]), C.c_int(A["col"]), C.c_int(len(a_off)), _p(a_off, _I), _p(a_val, _D),
                                 C.c_int(B["col"]), C.c_int(len(b_off)), _p(b_off, _I), _p(b_val, _D),
                                 C.byref(nd), C.byref(di), C.byref(off), C.byref(val))
        rows = A["row"]
        return {"row": rows, "col": B["col"], "num_diagonals": nd.value, "choice": True,
                "diagonal_ind": _take(di, rows + B["col"] - 1, np.int32, self._free),
                "diagonal_offsets": _take(off, nd.value, np.int32, self._free),
                "values": _take(val, rows * nd.value, np.float64, self._free).reshape(rows, nd.value)}

    # -- ELL --------------------------------------------------------------------
    def csr_to_ell(self, rows, cols, rp, ci, v, gate=50.0):
        rp, ci, v = _i(rp), _i(ci), _d(v)
        ch, w = C.c_int(), C.c_int()
        nr, ec, ev = _I(), _I(), _D()
        self.lib.ora_csr_to_ell(C.c_int(rows), C.c_int(cols), C.c_int(int(rp[-1])), _p(rp, _I), _p(ci, _I), _p(v, _D),
                                C.c_double(gate), C.byref(ch), C.byref(w), C.byref(nr), C.byref(ec), C.byref(ev))
        out = {"choice": bool(ch.value), "width": w.value, "row": rows, "col": cols, "nnz": int(rp[-1])}
        if ch.value:
            out["nnz_row"] = _take(nr, rows, np.int32, self._free)
            out["col_ind"] = _take(ec, rows * w.value, np.int32, self._free).reshape(rows, w.value)
            out["values"] = _take(ev, rows * w.value, np.float64, self._free).reshape(rows, w.value)
        return out

    def ell_mul_ell(self, A, B):
        a_nr, a_ci, a_v = _i(A["nnz_row"]), _i(A["col_ind"]), _d(A["values"])
        b_nr, b_ci, b_v = _i(B["nnz_row"]), _i(B["col_ind"]), _d(B["values"])
        w, nnz = C.c_int(), C.c_int64()
        nr, ec, ev = _I(), _I(), _D()
        self.lib.ora_ell_mul_ell(C.c_int(A["row"]), C.c_int(A["width"]), _p(a_nr, _I), _p(a_ci, _I), _p(a_v, _D),
                                 C.c_int(B["col"]), C.c_int(B["width"]), _p(b_nr, _I), _p(b_ci, _I), _p(b_v, _D),
                                 C.byref(w), C.byref(nnz), C.byref(nr), C.byref(ec), C.byref(ev))
        rows = A["row"]
        return {"row": rows, "col": B["col"], "width": w.value, "nnz": nnz.value, "choice": True,
                "nnz_row": _take(nr, rows, np.int32, self._free),
                "col_ind": _take(ec, rows * w.value, np.int32, self._free).reshape(rows, w.value),
                "values": _take(ev, rows * w.value, np.float64, self._free).reshape(rows, w.value)}

    # -- COO --------------------------------------------------------------------
    def csr_to_coo(self, rows, rp, ci, v):
        rp, ci, v = _i(rp), _i(ci), _d(v)
        nnz = int(rp[-1])
        ro, ri, cc, vv = _I(), _I(), _I(), _D()
        self.lib.ora_csr_to_coo(C.c_int(rows), C.c_int(nnz), _p(rp, _I), _p(ci, _I), _p(v, _D),
                                C.byref(ro), C.byref(ri), C.byref(cc), C.byref(vv))
        return {"row": rows, "nnz": nnz,
                "row_offset": _take(ro, rows + 1, np.int32, self._free),
                "row_ind": _take(ri, nnz, np.int32, self._free),
                "col_ind": _take(cc, nnz, np.int32, self._free),
                "values": _take(vv, nnz, np.float64, self._free)}

    def coo_mul_coo(self, a_rows, b_cols, A, B):
        a_ro, a_ci, a_v = _i(A["row_offset"]), _i(A["col_ind"]), _d(A["values"])
        b_ro, b_ci, b_v = _i(B["row_offset"]), _i(B["col_ind"]), _d(B["values"])
        ro, ri, cc, vv = _L(), _I(), _I(), _D()
        self.lib.ora_coo_mul_coo(C.c_int(a_rows), C.c_int(b_cols), _p(a_ro, _I), _p(a_ci, _I), _p(a_v, _D),
                                 _p(b_ro, _I), _p(b_ci, _I), _p(b_v, _D),
                                 C.byref(ro), C.byref(ri), C.byref(cc), C.byref(vv))
        c_ro = _take(ro, a_rows + 1, np.int64, self._free)
        nnz = int(c_ro[-1])
        return {"row": a_rows, "col": b_cols, "nnz": nnz, "row_offset": c_ro,
                "row_ind": _take(ri, nnz, np.int32, self._free),
                "col_ind": _take(cc, nnz, np.int32, self._free),
                "values": _take(vv, nnz, np.float64, self._free)}

    def sizeof_csr(self, rows, nnz):
        return float(self.lib.ora_sizeof_csr(rows, nnz))

    def sizeof_dia(self, rows, cols, nd):
        return float(self.lib.ora_sizeof_dia(rows, cols, nd))

    def sizeof_ell(self, rows, w):
        return float(self.lib.ora_sizeof_ell(rows, w))

    def sizeof_coo(self, rows, nnz):
        return float(self.lib.ora_sizeof_coo(rows, nnz))


class Ref:
    """The reference's own CPU kernels (oracle/_ref/libiaref.so)."""
    PATH = os.path.join(HERE, "_ref", "libiaref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        if not self.available():
            raise FileNotFoundError(self.PATH)
        self.lib = C.CDLL(self.PATH)
        L = self.lib
        for n in ("ref_csr_mul_csr", "ref_mkl_mul_mkl", "ref_sizeof_csr", "ref_dia_mul_dia_from_csr",
                  "ref_ell_mul_ell_from_csr", "ref_coo_mul_coo_from_csr"):
            getattr(L, n).restype = C.c_double
        L.ref_getflop.restype = C.c_longlong
        L.ref_free.argtypes = [C.c_void_p]
        self._free = lambda p: L.ref_free(C.cast(p, C.c_void_p))

    def threads(self):
        return int(self.lib.ref_omp_threads()), int(self.lib.ref_mkl_threads())

    def set_threads(self, n):
        """OpenMP and MKL thread counts for the timed baseline (torchrun exports OMP_NUM_THREADS=1)."""
        self.lib.ref_set_threads(C.c_int(int(n)))

    def mkl_version(self):
        buf = C.create_string_buffer(256)
        self.lib.ref_mkl_version(buf, 256)
        return buf.value.decode(errors="replace").strip()

    @staticmethod
    def _args(M):
        rows, cols, rp, ci, v = M
        rp, ci, v = _i(rp), _i(ci), _d(v)
        keep = (rp, ci, v)
        return keep, [C.c_int(rows), C.c_int(cols), C.c_int(int(rp[-1])), _p(rp, _I), _p(ci, _I), _p(v, _D)]

    def csr_mul_csr(self, A, B):
        ka, aa = self._args(A)
        kb, bb = self._args(B)
        n = C.c_int()
        rp, ci, v = _I(), _I(), _D()
        ms = self.lib.ref_csr_mul_csr(*aa, *bb, C.byref(n), C.byref(rp), C.byref(ci), C.byref(v))
        return (_take(rp, A[0] + 1, np.int64, self._free), _take(ci, n.value, np.int32, self._free),
                _take(v, n.value, np.float64, self._free), ms)

    def mkl_mul_mkl(self, A, B, keep=True):
        ka, aa = self._args(A)
        kb, bb = self._args(B)
        n = C.c_int()
        rp, ci, v = _I(), _I(), _D()
        ms = self.lib.ref_mkl_mul_mkl(*aa, *bb, C.c_int(1 if keep else 0), C.byref(n), C.byref(rp), C.byref(ci), C.byref(v))
        if not keep:
            return None, None, None, ms, n.value
        return (_take(rp, A[0] + 1, np.int64, self._free), _take(ci, n.value, np.int32, self._free),
                _take(v, n.value, np.float64, self._free), ms, n.value)

    def getflop(self, A, B):
        ka, aa = self._args(A)
        kb, bb = self._args(B)
        return int(self.lib.ref_getflop(*aa, *bb))

    def sizeof_csr(self, rows, cols, nnz):
        return float(self.lib.ref_sizeof_csr(rows, cols, nnz))

    def features26(self, A, B):
        ka, aa = self._args(A)
        kb, bb = self._args(B)
        f = np.zeros(26)
        self.lib.ref_features26(*aa, *bb, _p(f, _D))
        return f

    def csr_to_dia(self, A):
        ka, aa = self._args(A)
        ch, nd = C.c_int(), C.c_int()
        di, off, val = _I(), _I(), _D()
        self.lib.ref_csr_to_dia(*aa, C.byref(ch), C.byref(nd), C.byref(di), C.byref(off), C.byref(val))
        out = {"choice": bool(ch.value), "num_diagonals": nd.value, "row": A[0], "col": A[1]}
        if ch.value:
            out["diagonal_ind"] = _take(di, A[0] + A[1] - 1, np.int32, self._free)
            out["diagonal_offsets"] = _take(off, nd.value, np.int32, self._free)
            out["values"] = _take(val, A[0] * nd.value, np.float64, self._free).reshape(A[0], nd.value)
        return out

    def dia_mul_dia(self, A, B):
        ka, aa = self._args(A)
        kb, bb = self._args(B)
        ok, nd = C.c_int(), C.c_int()
        di, off, val = _I(), _I(), _D()
        ms = self.lib.ref_dia_mul_dia_from_csr(*aa, *bb, C.byref(ok), C.byref(nd), C.byref(di), C.byref(off), C.byref(val))
        if not ok.value:
            return None
        return {"row": A[0], "col": B[1], "num_diagonals": nd.value, "ms": ms,
                "diagonal_ind": _take(di, A[0] + B[1] - 1, np.int32, self._free),
                "diagonal_offsets": _take(off, nd.value, np.int32, self._free),
                "values": _take(val, A[0] * nd.value, np.float64, self._free).reshape(A[0], nd.value)}

    def csr_to_ell(self, A):
        ka, aa = self._args(A)
        ch, w = C.c_int(), C.c_int()
        nr, ec, ev = _I(), _I(), _D()
        self.lib.ref_csr_to_ell(*aa, C.byref(ch), C.byref(w), C.byref(nr), C.byref(ec), C.byref(ev))
        out = {"choice": bool(ch.value), "width": w.value, "row": A[0], "col": A[1]}
        if ch.value:
            out["nnz_row"] = _take(nr, A[0], np.int32, self._free)
            out["col_ind"] = _take(ec, A[0] * w.value, np.int32, self._free).reshape(A[0], w.value)
            out["values"] = _take(ev, A[0] * w.value, np.float64, self._free).reshape(A[0], w.value)
        return out

    def ell_mul_ell(self, A, B):
        ka, aa = self._args(A)
        kb, bb = self._args(B)
        ok, w, n = C.c_int(), C.c_int(), C.c_int()
        nr, ec, ev = _I(), _I(), _D()
        ms = self.lib.ref_ell_mul_ell_from_csr(*aa, *bb, C.byref(ok), C.byref(w), C.byref(n), C.byref(nr), C.byref(ec), C.byref(ev))
        if not ok.value:
            return None
        return {"row": A[0], "col": B[1], "width": w.value, "nnz": n.value, "ms": ms,
                "nnz_row": _take(nr, A[0], np.int32, self._free),
                "col_ind": _take(ec, A[0] * w.value, np.int32, self._free).reshape(A[0], w.value),
                "values": _take(ev, A[0] * w.value, np.float64, self._free).reshape(A[0], w.value)}

    def coo_mul_coo(self, A, B):
        ka, aa = self._args(A)
        kb, bb = self._args(B)
        n = C.c_int()
        ro, ri, cc, vv = _I(), _I(), _I(), _D()
        ms = self.lib.ref_coo_mul_coo_from_csr(*aa, *bb, C.byref(n), C.byref(ro), C.byref(ri), C.byref(cc), C.byref(vv))
        return {"row": A[0], "col": B[1], "nnz": n.value, "ms": ms,
                "row_offset": _take(ro, A[0] + 1, np.int64, self._free),
                "row_ind": _take(ri, n.value, np.int32, self._free),
                "col_ind": _take(cc, n.value, np.int32, self._free),
                "values": _take(vv, n.value, np.float64, self._free)}
