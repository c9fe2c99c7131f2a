"""Host-side mirror of the reference's operator interface, over the C ABI (include/iaspgemm.h).

The reference exposes free functions on POD structs (`CSR_MUL_CSR(A, B, C)`, `CSRtoDIA`,
`GetInfo1`, `GetFlop`, ... -- SURVEY.md section 8b); this module offers the same names on NumPy
operands for the test-suite and bench harness.  Everything here calls libiaspgemm.so through
ctypes: there is no CPU fallback, and import fails loudly if the library is missing.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IAS_LIB") or os.path.join(HERE, "libiaspgemm.so")     # IAS_LIB: an instrumented build (make prof)

_I = C.POINTER(C.c_int)
_L = C.POINTER(C.c_longlong)
_D = C.POINTER(C.c_double)


class CsrMatrix(C.Structure):          # IasCsrMatrix == CsrMatrix (GPU/detail/format.h:47-57)
    _fields_ = [("choice", C.c_bool), ("row", C.c_int), ("col", C.c_int), ("nnz", C.c_int),
                ("row_ind", _I), ("col_ind", _I), ("values", _D)]


class CsrMatrixDev(C.Structure):       # IasCsrMatrixDev == CsrMatrixDev (format.h:59-69)
    _fields_ = [("choice", C.c_bool), ("row", C.c_int), ("col", C.c_int), ("nnz", C.c_int),
                ("row_ind_dev", C.c_void_p), ("col_ind_dev", C.c_void_p), ("values_dev", C.c_void_p)]


class Csr64Dev(C.Structure):
    _fields_ = [("row", C.c_int), ("col", C.c_int), ("nnz", C.c_longlong),
                ("row_ptr_dev", C.c_void_p), ("col_ind_dev", C.c_void_p), ("values_dev", C.c_void_p)]


class CooDev(C.Structure):             # IasCooDev == CooMatrixDev (format.h:29-40)
    _fields_ = [("choice", C.c_bool), ("row", C.c_int), ("col", C.c_int), ("nnz", C.c_int),
                ("row_offset_dev", C.c_void_p), ("row_ind_dev", C.c_void_p), ("col_ind_dev", C.c_void_p),
                ("values_dev", C.c_void_p)]


class Coo64Dev(C.Structure):           # results beyond int32
    _fields_ = [("choice", C.c_bool), ("row", C.c_int), ("col", C.c_int), ("nnz", C.c_longlong),
                ("row_offset_dev", C.c_void_p), ("row_ind_dev", C.c_void_p), ("col_ind_dev", C.c_void_p),
                ("values_dev", C.c_void_p)]


class DiaDev(C.Structure):
    _fields_ = [("choice", C.c_bool), ("row", C.c_int), ("col", C.c_int), ("num_diagonals", C.c_int),
                ("diagonal_ind_dev", C.c_void_p), ("diagonal_offsets_dev", C.c_void_p), ("values_dev", C.c_void_p)]


class EllDev(C.Structure):             # IasEllDev == EllMatrixDev (format.h:108-119)
    _fields_ = [("choice", C.c_bool), ("row", C.c_int), ("col", C.c_int), ("nnz", C.c_int),
                ("max_nnz_per_row", C.c_int),
                ("nnz_row_dev", C.c_void_p), ("col_ind_dev", C.c_void_p), ("values_dev", C.c_void_p)]


class Ell64Dev(C.Structure):           # results beyond int32
    _fields_ = [("choice", C.c_bool), ("row", C.c_int), ("col", C.c_int), ("nnz", C.c_longlong),
                ("max_nnz_per_row", C.c_int),
                ("nnz_row_dev", C.c_void_p), ("col_ind_dev", C.c_void_p), ("values_dev", C.c_void_p)]


class StreamBatch(C.Structure):        # IasStreamBatch
    _fields_ = [("row_begin", C.c_int), ("row_end", C.c_int), ("batch_index", C.c_int), ("batch_count", C.c_int),
                ("nnz_total", C.c_longlong), ("entry_base", C.c_longlong), ("batch_nnz", C.c_longlong),
                ("row_ptr_dev", C.c_void_p), ("col_ind_dev", C.c_void_p), ("values_dev", C.c_void_p),
                ("cuda_stream", C.c_void_p)]


class AutoResult(C.Structure):         # IasAutoResult
    _fields_ = [("format", C.c_int), ("row", C.c_int), ("col", C.c_int), ("nnz", C.c_longlong),
                ("row_ptr", _L), ("col_ind", _I), ("values", _D),
                ("num_diagonals", C.c_int), ("diagonal_ind", _I), ("diagonal_offsets", _I),
                ("max_nnz_per_row", C.c_int), ("nnz_row", _I),
                ("features", C.c_double * 26),
                ("ms_h2d", C.c_double), ("ms_select", C.c_double), ("ms_convert", C.c_double), ("ms_multiply", C.c_double),
                ("ms_d2h", C.c_double), ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong), ("ms_wall", C.c_double), ("ms_host", C.c_double * 6), ("pipelined", C.c_int)]


STREAM_CONSUMER = C.CFUNCTYPE(C.c_int, C.POINTER(StreamBatch), C.c_void_p)


class SpgemmStats(C.Structure):
    _fields_ = [("products", C.c_longlong), ("nnz", C.c_longlong), ("ms_total", C.c_double),
                ("ms_analyze", C.c_double), ("ms_symbolic", C.c_double), ("ms_scan", C.c_double),
                ("ms_numeric", C.c_double), ("ms_consume", C.c_double),
                ("ms_bin_sym", C.c_double * 8), ("ms_bin_num", C.c_double * 8),
                ("sym_bin_rows", C.c_longlong * 8), ("num_bin_rows", C.c_longlong * 8),
                ("batches", C.c_int), ("kernel_launches", C.c_int), ("checksum", C.c_double),
                ("structure_hash", C.c_ulonglong)]

    def as_dict(self):
        return {"products": self.products, "nnz": self.nnz, "ms_total": self.ms_total,
                "ms_analyze": self.ms_analyze, "ms_symbolic": self.ms_symbolic, "ms_scan": self.ms_scan,
                "ms_numeric": self.ms_numeric, "ms_bin_sym": list(self.ms_bin_sym)[:6],
                "ms_bin_num": list(self.ms_bin_num)[:6], "sym_bin_rows": list(self.sym_bin_rows)[:6],
                "num_bin_rows": list(self.num_bin_rows)[:6], "batches": self.batches,
                "kernel_launches": self.kernel_launches, "checksum": self.checksum,
                "structure_hash": self.structure_hash}


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libiaspgemm status %d: %s" % (code, msg))
        self.code = code


BIN_NAMES = ["empty", "tiny", "warp", "cta_s", "cta_l", "global"]

# every symbol include/iaspgemm.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "ias_init", "ias_set_stream", "ias_use_own_stream", "ias_sync", "ias_trim_pool", "ias_last_error", "ias_version", "ias_device_info",
    "ias_kernel_launches", "ias_set_option", "ias_get_option",
    "ias_upload_csr", "ias_free_csr_dev", "ias_free_csr64_dev", "ias_download_csr64", "ias_download_csr",
    "ias_csr_is_canonical", "ias_copy", "ias_forget_operand",
    "ias_csr_mul_csr_dev64", "ias_csr_mul_csr_dev", "ias_csr_mul_csr_rows_dev64", "ias_csr_mul_csr_stream",
    "ias_csr_mul_csr_stream_cb", "ias_csr_mul_csr_rowlist_stream",
    "ias_csr_mul_csr_host", "ias_release_host", "ias_getflop", "ias_touched_b_bytes", "ias_partition_rows", "ias_row_share", "ias_checksum", "ias_device_alloc", "ias_device_free",
    "ias_structure_hash",
    "ias_csr_to_dia", "ias_dia_mul_dia_dev", "ias_dia_mul_dia_rows_dev", "ias_download_dia", "ias_free_dia_dev", "ias_dia_relayout",
    "ias_csr_to_ell", "ias_ell_mul_ell_dev", "ias_ell_mul_ell_dev64", "ias_download_ell", "ias_download_ell64", "ias_free_ell_dev", "ias_free_ell64_dev",
    "ias_csr_to_coo", "ias_coo_mul_coo_dev", "ias_coo_mul_coo_dev64", "ias_download_coo", "ias_download_coo64",
    "ias_free_coo_dev", "ias_free_coo64_dev",
    "ias_density_image", "ias_getinfo1", "ias_getinfo2", "ias_getinfo3", "ias_count_diagonals",
    "ias_max_row_nnz", "ias_features26",
    "ias_sizeof_csr", "ias_sizeof_dia", "ias_sizeof_ell", "ias_sizeof_coo",
    "ias_mtx_load", "ias_free_host_csr", "ias_mtx_write_csr64", "ias_csr_transpose",
    "ias_gen_poisson2d", "ias_gen_uniform", "ias_gen_rmat",
    "ias_matnet_load", "ias_matnet_create", "ias_matnet_set_tensor", "ias_matnet_get_tensor", "ias_matnet_shape",
    "ias_matnet_predict", "ias_matnet_free",
    "ias_select_format", "ias_spgemm_auto_host",
]


class MatNet:
    """Native MatNet selector (MatNet.Pred, CPU/MatNet.py:24-96).  Host only: usable without a GPU."""
    TENSORS = ["conv2d_%d/%s" % (i, k) for i in range(1, 7) for k in ("kernel", "bias")] + \
              ["dense_%d/%s" % (i, k) for i in range(1, 5) for k in ("kernel", "bias")]

    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle

    @classmethod
    def load(cls, path, lib=None):
        lib = lib or load_library()
        h = C.c_void_p()
        rc = lib.ias_matnet_load(path.encode(), C.byref(h))
        if rc != 0:
            raise EngineError(rc, (lib.ias_last_error() or b"").decode(errors="replace"))
        return cls(lib, h)

    @classmethod
    def from_arrays(cls, tensors, lib=None):
        lib = lib or load_library()
        h = C.c_void_p()
        lib.ias_matnet_create(C.byref(h))
        for name, a in tensors.items():
            a = np.ascontiguousarray(a, dtype=np.float32)
            dims = (C.c_longlong * a.ndim)(*a.shape)
            rc = lib.ias_matnet_set_tensor(h, name.encode(), a.ctypes.data_as(C.POINTER(C.c_float)), dims, a.ndim)
            if rc != 0:
                raise EngineError(rc, (lib.ias_last_error() or b"").decode(errors="replace"))
        return cls(lib, h)

    def shape(self):
        f, c, p = C.c_int(), C.c_int(), C.c_longlong()
        self.lib.ias_matnet_shape(self.h, C.byref(f), C.byref(c), C.byref(p))
        return f.value, c.value, p.value

    def tensor(self, name):
        dims, rank = (C.c_longlong * 4)(), C.c_int()
        if self.lib.ias_matnet_get_tensor(self.h, name.encode(), None, dims, C.byref(rank)) != 0:
            raise KeyError(name)
        shape = tuple(dims[i] for i in range(rank.value))
        out = np.empty(shape, dtype=np.float32)
        self.lib.ias_matnet_get_tensor(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), None, None)
        return out

    def predict(self, img1, img2, features):
        img1 = np.ascontiguousarray(img1, dtype=np.int64).reshape(-1)
        img2 = np.ascontiguousarray(img2, dtype=np.int64).reshape(-1)
        f = np.ascontiguousarray(features, dtype=np.float64)
        nf, nc, _ = self.shape()
        assert img1.size == 16384 and img2.size == 16384 and f.size >= nf
        cls_, probs = C.c_int(), np.zeros(nc)
        rc = self.lib.ias_matnet_predict(self.h, img1.ctypes.data_as(_L), img2.ctypes.data_as(_L), f.ctypes.data_as(_D),
                                         C.byref(cls_), probs.ctypes.data_as(_D))
        if rc != 0:
            raise EngineError(rc, (self.lib.ias_last_error() or b"").decode(errors="replace"))
        return cls_.value, probs

    def close(self):
        if self.h:
            self.lib.ias_matnet_free(self.h)
            self.h = None


def load_library():
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C ia_spgemm_b200/csrc).  The engine has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.ias_last_error.restype = C.c_char_p
    lib.ias_version.restype = C.c_char_p
    lib.ias_kernel_launches.restype = C.c_longlong
    for n in ("ias_sizeof_csr", "ias_sizeof_dia", "ias_sizeof_ell", "ias_sizeof_coo"):
        getattr(lib, n).restype = C.c_double
    sigs = {
        "ias_sizeof_csr": [C.c_int, C.c_longlong], "ias_sizeof_coo": [C.c_int, C.c_longlong],
        "ias_sizeof_dia": [C.c_int, C.c_int, C.c_int], "ias_sizeof_ell": [C.c_int, C.c_int],
        "ias_copy": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int],
        "ias_set_option": [C.c_char_p, C.c_longlong], "ias_get_option": [C.c_char_p, C.c_void_p],
        "ias_set_stream": [C.c_void_p], "ias_checksum": [C.c_void_p, C.c_longlong, _D],
        "ias_csr_mul_csr_stream": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p],
        "ias_csr_mul_csr_stream_cb": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, STREAM_CONSUMER, C.c_void_p,
                                      C.c_void_p],
        "ias_dia_relayout": [C.c_void_p, C.c_int, C.c_void_p],
        "ias_csr_mul_csr_rowlist_stream": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p],
        "ias_select_format": [C.c_void_p, C.c_int, C.c_int],
        "ias_spgemm_auto_host": [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p],
        "ias_getinfo3": [C.c_int, C.c_longlong, C.c_int, _D],
        "ias_gen_rmat": [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p],
        "ias_csr_to_dia": [C.c_void_p, C.c_double, C.c_void_p], "ias_csr_to_ell": [C.c_void_p, C.c_double, C.c_void_p],
        "ias_matnet_load": [C.c_char_p, C.c_void_p], "ias_matnet_create": [C.c_void_p], "ias_matnet_free": [C.c_void_p],
        "ias_matnet_set_tensor": [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int],
        "ias_matnet_get_tensor": [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p],
        "ias_matnet_shape": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        "ias_matnet_predict": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    }
    for name, args in sigs.items():
        if hasattr(lib, name):          # a missing export is reported by tests/test_abi.py, not here
            getattr(lib, name).argtypes = args
    return lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class DeviceCsr:
    """A CSR operand resident on the GPU (CsrMatrixDev).  Freed on close()/GC."""

    def __init__(self, eng, dev, owned=True):
        self.eng, self.dev, self.owned = eng, dev, owned

    @property
    def shape(self):
        return self.dev.row, self.dev.col

    @property
    def nnz(self):
        return self.dev.nnz

    def download(self):
        d = self.dev
        rp = np.empty(d.row + 1, np.int32)
        ci = np.empty(d.nnz, np.int32)
        v = np.empty(d.nnz, np.float64)
        self.eng._ck(self.eng.lib.ias_download_csr(C.byref(d), rp.ctypes.data_as(_I), ci.ctypes.data_as(_I), v.ctypes.data_as(_D)))
        return d.row, d.col, rp, ci, v

    def close(self):
        if self.owned and self.dev.row_ind_dev:
            self.eng.lib.ias_free_csr_dev(C.byref(self.dev))
        elif self.dev.col_ind_dev:
            self.eng.lib.ias_forget_operand(C.byref(self.dev))      # caller-owned memory may be re-used at the same address
        self.owned = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, device=0):
        self.lib = load_library()
        self._ck(self.lib.ias_init(C.c_int(device)))

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(rc, (self.lib.ias_last_error() or b"").decode(errors="replace"))

    # -- context -----------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        """Run on the caller's cudaStream_t; 0 / None is the legacy default stream (torch's default stream handle)."""
        self._ck(self.lib.ias_set_stream(C.c_void_p(cuda_stream_handle or 0)))

    def use_own_stream(self):
        self._ck(self.lib.ias_use_own_stream())

    def sync(self):
        self._ck(self.lib.ias_sync())

    def trim_pool(self):
        self._ck(self.lib.ias_trim_pool())

    def kernel_launches(self):
        return int(self.lib.ias_kernel_launches())

    def set_option(self, name, value):
        """Kernel-selection knob (include/iaspgemm.h: ias_set_option); results never depend on it."""
        self._ck(self.lib.ias_set_option(name.encode(), int(value)))

    def global_numeric_kernel(self, cols, b_canonical=True):
        """Name of the kernel that takes the rows beyond the CTA hash (mirrors gwin_numeric_pays in spgemm_host.cuh)."""
        if not b_canonical or not self.get_option("global_rows_smem"):
            return "k_num_global"
        words = ((cols + 31) // 32 + 31) // 32 * 32
        sw = self.get_option("gwin_swords") or 8192
        sw = max(32, (sw + 31) // 32 * 32)
        nsw = (words + sw - 1) // sw
        mx = self.get_option("gwin_max_sw")
        if mx == 0 or nsw <= mx:
            return "k_num_gwin"
        second_generation = self.get_option("g_v2") != 0 and self.get_option("g_block") != 512
        return "k_num_global2" if second_generation else "k_num_global"

    def get_option(self, name):
        v = C.c_longlong()
        self._ck(self.lib.ias_get_option(name.encode(), C.byref(v)))
        return int(v.value)

    def device_info(self):
        sm, smem, fr, tot = C.c_int(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._ck(self.lib.ias_device_info(C.byref(sm), C.byref(smem), C.byref(fr), C.byref(tot)))
        return {"sm_count": sm.value, "smem_optin": smem.value, "free_bytes": fr.value, "total_bytes": tot.value}

    # -- transfers -----------------------------------------------------------------------
    @staticmethod
    def host_csr(rows, cols, rp, ci, v):
        rp, ci, v = _i32(rp), _i32(ci), _f64(v)
        h = CsrMatrix(True, rows, cols, int(rp[-1]), rp.ctypes.data_as(_I), ci.ctypes.data_as(_I), v.ctypes.data_as(_D))
        h._keep = (rp, ci, v)
        return h

    def upload(self, rows, cols, rp, ci, v):
        """UploadCsrMatrix (csr_dev:111)."""
        h = self.host_csr(rows, cols, rp, ci, v)
        d = CsrMatrixDev()
        self._ck(self.lib.ias_upload_csr(C.byref(h), C.byref(d)))
        return DeviceCsr(self, d)

    def wrap_device(self, rows, cols, nnz, rp_ptr, ci_ptr, v_ptr):
        """Operand living in caller-owned device memory (e.g. torch tensors)."""
        d = DeviceCsr(self, CsrMatrixDev(True, rows, cols, nnz, rp_ptr, ci_ptr, v_ptr), owned=False)
        self.lib.ias_forget_operand(C.byref(d.dev))                 # whatever lived at these addresses before is gone
        return d

    def copy(self, dst_ptr, src_ptr, nbytes, kind):
        """kind: 0 host->device, 1 device->host, 2 device->device."""
        self._ck(self.lib.ias_copy(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), nbytes, kind))

    def transpose(self, A):
        """B := A^T on device (GPU/main.cu:261-269 does it with mkl_dcsrcsc on the host)."""
        d = CsrMatrixDev()
        self._ck(self.lib.ias_csr_transpose(C.byref(A.dev), C.byref(d)))
        return DeviceCsr(self, d)

    def mtx_write(self, path, c64, row_base=0):
        self._ck(self.lib.ias_mtx_write_csr64(path.encode(), C.byref(c64), C.c_int(row_base)))

    def forget_operand(self, A=None):
        self.lib.ias_forget_operand(C.byref(A.dev) if A is not None else None)

    def is_canonical(self, A):
        r = C.c_int()
        self._ck(self.lib.ias_csr_is_canonical(C.byref(A.dev), C.byref(r)))
        return bool(r.value)

    def _take_csr64(self, c64, download=True):
        out = None
        if download:
            rp = np.empty(c64.row + 1, np.int64)
            ci = np.empty(c64.nnz, np.int32)
            v = np.empty(c64.nnz, np.float64)
            self._ck(self.lib.ias_download_csr64(C.byref(c64), rp.ctypes.data_as(_L), ci.ctypes.data_as(_I), v.ctypes.data_as(_D)))
            out = (rp, ci, v)
        self.lib.ias_free_csr64_dev(C.byref(c64))
        return out

    # -- Algorithm 2 ---------------------------------------------------------------------
    def CSR_MUL_CSR_DEV(self, A, B, rows=None, download=True, keep=False):
        """CSR_MUL_CSR_DEV (csr_dev:134): C = A*B on device operands.
        Returns ((row_ptr int64, col_ind, values) or None, stats dict[, Csr64Dev if keep])."""
        c64, st = Csr64Dev(), SpgemmStats()
        if rows is None:
            self._ck(self.lib.ias_csr_mul_csr_dev64(C.byref(A.dev), C.byref(B.dev), C.byref(c64), C.byref(st)))
        else:
            self._ck(self.lib.ias_csr_mul_csr_rows_dev64(C.byref(A.dev), C.byref(B.dev), C.c_int(rows[0]), C.c_int(rows[1]),
                                                         C.byref(c64), C.byref(st)))
        if keep:
            return c64, st.as_dict()
        return self._take_csr64(c64, download), st.as_dict()

    def free_csr64(self, c64):
        self.lib.ias_free_csr64_dev(C.byref(c64))

    def csr_mul_csr_dev32(self, A, B):
        """int32 reference layout (CsrMatrixDev result)."""
        cdev, ms = CsrMatrixDev(), C.c_double()
        self._ck(self.lib.ias_csr_mul_csr_dev(C.byref(A.dev), C.byref(B.dev), C.byref(cdev), C.byref(ms)))
        out = DeviceCsr(self, cdev)
        res = out.download()
        out.close()
        return res[2], res[3], res[4], ms.value

    def CSR_MUL_CSR(self, A, B):
        """CSR_MUL_CSR(A, B, C) on host operands (CPU/detail/csr/common_csr.h:85) -- the e2e path.
        A, B = (rows, cols, rp, ci, v).  Returns (rp, ci, v) views of engine-owned pinned memory
        (valid until the next host call), stats, ms_h2d, ms_d2h."""
        hA = self.host_csr(*A)
        hB = hA if B is A else self.host_csr(*B)
        rp, ci, v = _L(), _I(), _D()
        nnz, st, h2d, d2h = C.c_longlong(), SpgemmStats(), C.c_double(), C.c_double()
        self._ck(self.lib.ias_csr_mul_csr_host(C.byref(hA), C.byref(hB), C.byref(rp), C.byref(ci), C.byref(v), C.byref(nnz),
                                               C.byref(st), C.byref(h2d), C.byref(d2h)))
        n = nnz.value
        out = (np.ctypeslib.as_array(rp, shape=(A[0] + 1,)),
               np.ctypeslib.as_array(ci, shape=(n,)) if n else np.zeros(0, np.int32),
               np.ctypeslib.as_array(v, shape=(n,)) if n else np.zeros(0, np.float64))
        return out, st.as_dict(), h2d.value, d2h.value

    def spgemm_auto(self, A, B, gate=20.0, matnet=None):
        """The front end's path on host operands (ias_spgemm_auto_host): features -> selection -> conversion ->
        multiply -> host result in the selected format.  Returns a dict with views of engine-owned pinned memory."""
        hA = self.host_csr(*A)
        hB = hA if B is A else self.host_csr(*B)
        r = AutoResult()
        self._ck(self.lib.ias_spgemm_auto_host(C.byref(hA), C.byref(hB), gate, matnet.h if matnet is not None else None, C.byref(r)))
        out = {"format": {1: "csr", 2: "dia", 3: "ell"}[r.format], "row": r.row, "col": r.col, "nnz": r.nnz,
               "features": np.array(r.features), "h2d_bytes": r.h2d_bytes, "d2h_bytes": r.d2h_bytes,
               "ms": {k: getattr(r, "ms_" + k) for k in ("h2d", "select", "convert", "multiply", "d2h", "wall")},
               "host_ms_at": [round(x, 3) for x in r.ms_host], "pipelined": bool(r.pipelined)}
        arr = np.ctypeslib.as_array
        if r.format == 1:
            out["row_ptr"] = arr(r.row_ptr, shape=(r.row + 1,))
            out["col_ind"] = arr(r.col_ind, shape=(r.nnz,)) if r.nnz else np.zeros(0, np.int32)
            out["values"] = arr(r.values, shape=(r.nnz,)) if r.nnz else np.zeros(0)
        elif r.format == 2:
            nd = r.num_diagonals
            out["num_diagonals"] = nd
            out["diagonal_offsets"] = arr(r.diagonal_offsets, shape=(nd,)) if nd else np.zeros(0, np.int32)
            out["diagonal_ind"] = arr(r.diagonal_ind, shape=(max(r.row + r.col - 1, 1),))
            out["values"] = arr(r.values, shape=(r.row, nd)) if r.row * nd else np.zeros((r.row, nd))
        else:
            w = r.max_nnz_per_row
            out["width"] = w
            out["nnz_row"] = arr(r.nnz_row, shape=(r.row,)) if r.row else np.zeros(0, np.int32)
            out["col_ind"] = arr(r.col_ind, shape=(r.row, w)) if r.row * w else np.zeros((r.row, w), np.int32)
            out["values"] = arr(r.values, shape=(r.row, w)) if r.row * w else np.zeros((r.row, w))
        return out

    def select_format(self, features26, dia_ok=True, ell_ok=True):
        f = np.ascontiguousarray(features26, dtype=np.float64)
        return int(self.lib.ias_select_format(f.ctypes.data_as(_D), int(dia_ok), int(ell_ok)))

    def csr_mul_csr_stream(self, A, B, rows=None, budget_bytes=0, want_row_nnz=False, consumer=None):
        """consumer(batch_dict) is called once per row batch with host copies of the batch (row_begin, row_end,
        row_ptr relative to the batch, col_ind, values, nnz_total): the §8(b) streaming consumer."""
        r0, r1 = rows if rows is not None else (0, A.dev.row)
        st = SpgemmStats()
        row_nnz = None
        ptr = None
        if want_row_nnz:
            import torch
            row_nnz = torch.empty(max(r1 - r0, 1), dtype=torch.int32, device="cuda")
            ptr = row_nnz.data_ptr()
        if consumer is not None:
            def _cb(bp, _user):
                b = bp.contents
                n = b.row_end - b.row_begin
                rp = np.empty(n + 1, np.int64)
                ci = np.empty(b.batch_nnz, np.int32)
                v = np.empty(b.batch_nnz, np.float64)
                self.copy(rp.ctypes.data, b.row_ptr_dev, 8 * (n + 1), 1)
                if b.batch_nnz:
                    self.copy(ci.ctypes.data, b.col_ind_dev, 4 * b.batch_nnz, 1)
                    self.copy(v.ctypes.data, b.values_dev, 8 * b.batch_nnz, 1)
                try:
                    rc = consumer({"row_begin": b.row_begin, "row_end": b.row_end, "batch_index": b.batch_index,
                                   "batch_count": b.batch_count, "nnz_total": b.nnz_total, "entry_base": b.entry_base,
                                   "row_ptr": rp - b.entry_base, "col_ind": ci, "values": v})
                except Exception:          # an exception cannot cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 99
                return int(rc or 0)
            cb = STREAM_CONSUMER(_cb)
            self._ck(self.lib.ias_csr_mul_csr_stream_cb(C.byref(A.dev), C.byref(B.dev), r0, r1, budget_bytes, ptr, cb, None, C.byref(st)))
        else:
            self._ck(self.lib.ias_csr_mul_csr_stream(C.byref(A.dev), C.byref(B.dev), r0, r1, budget_bytes, ptr, C.byref(st)))
        d = st.as_dict()
        if want_row_nnz:
            d["row_nnz"] = row_nnz[: r1 - r0].cpu().numpy()
        return d

    def csr_mul_csr_rowlist_stream(self, A, B, rows_dev_ptr, nrows, budget_bytes=0, want_row_nnz=False, consumer=None):
        """Rows rows_dev[0..nrows) of A*B, streamed (ias_csr_mul_csr_rowlist_stream).  rows_dev_ptr: device pointer to int32
        row indices (e.g. a torch tensor's data_ptr()).  consumer(batch_dict) as in csr_mul_csr_stream; rows are in list order."""
        st = SpgemmStats()
        row_nnz, ptr = None, None
        if want_row_nnz:
            import torch
            row_nnz = torch.empty(max(nrows, 1), dtype=torch.int32, device="cuda")
            ptr = row_nnz.data_ptr()
        cb = None
        if consumer is not None:
            def _cb(bp, _user):
                b = bp.contents
                n = b.row_end - b.row_begin
                rp = np.empty(n + 1, np.int64)
                ci = np.empty(b.batch_nnz, np.int32)
                v = np.empty(b.batch_nnz, np.float64)
                self.copy(rp.ctypes.data, b.row_ptr_dev, 8 * (n + 1), 1)
                if b.batch_nnz:
                    self.copy(ci.ctypes.data, b.col_ind_dev, 4 * b.batch_nnz, 1)
                    self.copy(v.ctypes.data, b.values_dev, 8 * b.batch_nnz, 1)
                try:
                    rc = consumer({"row_begin": b.row_begin, "row_end": b.row_end, "nnz_total": b.nnz_total,
                                   "row_ptr": rp - b.entry_base, "col_ind": ci, "values": v})
                except Exception:
                    import traceback
                    traceback.print_exc()
                    return 99
                return int(rc or 0)
            cb = STREAM_CONSUMER(_cb)
        self._ck(self.lib.ias_csr_mul_csr_rowlist_stream(C.byref(A.dev), C.byref(B.dev), C.c_void_p(rows_dev_ptr), nrows, budget_bytes, ptr,
                                                         cb if cb is not None else C.cast(None, STREAM_CONSUMER), None, C.byref(st)))
        d = st.as_dict()
        if want_row_nnz:
            d["row_nnz"] = row_nnz[:nrows].cpu().numpy()
        return d

    def GetFlop(self, A, B):
        p = C.c_longlong()
        self._ck(self.lib.ias_getflop(C.byref(A.dev), C.byref(B.dev), C.byref(p)))
        return p.value

    def touched_b_bytes(self, A, B, rows=None):
        r0, r1 = rows if rows is not None else (0, A.dev.row)
        b = C.c_longlong()
        self._ck(self.lib.ias_touched_b_bytes(C.byref(A.dev), C.byref(B.dev), C.c_int(r0), C.c_int(r1), C.byref(b)))
        return b.value

    def partition_rows(self, A, B, parts):
        b = (C.c_int * (parts + 1))()
        self._ck(self.lib.ias_partition_rows(C.byref(A.dev), C.byref(B.dev), C.c_int(parts), b))
        return list(b)

    def row_share(self, A, B, parts, part):
        """Device tensor (torch int32) with this part's rows: rows sorted by decreasing products, dealt in snake order."""
        import torch
        cap = (A.dev.row + parts - 1) // parts + 1
        t = torch.empty(cap, dtype=torch.int32, device="cuda")
        n = C.c_int()
        self._ck(self.lib.ias_row_share(C.byref(A.dev), C.byref(B.dev), C.c_int(parts), C.c_int(part), C.c_void_p(t.data_ptr()), C.byref(n)))
        return t[: n.value]

    def checksum_ptr(self, ptr, n):
        s = C.c_double()
        self._ck(self.lib.ias_checksum(C.c_void_p(ptr), n, C.byref(s)))
        return s.value

    def structure_hash(self, c64, row_base=0):
        h = C.c_ulonglong()
        self._ck(self.lib.ias_structure_hash(C.byref(c64), C.c_int(row_base), C.byref(h)))
        return h.value

    # -- DIA -----------------------------------------------------------------------------
    def CSRtoDIA(self, A, gate=20.0):
        d = DiaDev()
        self._ck(self.lib.ias_csr_to_dia(C.byref(A.dev), gate, C.byref(d)))
        return d

    def download_dia(self, d):
        di = np.zeros(max(d.row + d.col - 1, 1), np.int32)
        off = np.zeros(max(d.num_diagonals, 1), np.int32)
        val = np.zeros(max(d.row * d.num_diagonals, 1), np.float64)
        self._ck(self.lib.ias_download_dia(C.byref(d), di.ctypes.data_as(_I), off.ctypes.data_as(_I), val.ctypes.data_as(_D)))
        return {"row": d.row, "col": d.col, "num_diagonals": d.num_diagonals, "choice": bool(d.choice),
                "diagonal_ind": di[: d.row + d.col - 1], "diagonal_offsets": off[: d.num_diagonals],
                "values": val[: d.row * d.num_diagonals].reshape(d.row, d.num_diagonals)}

    def DIA_MUL_DIA_DEV(self, A, B, rows=None):
        c, ms = DiaDev(), C.c_double()
        if rows is None:
            self._ck(self.lib.ias_dia_mul_dia_dev(C.byref(A), C.byref(B), C.byref(c), C.byref(ms)))
        else:
            self._ck(self.lib.ias_dia_mul_dia_rows_dev(C.byref(A), C.byref(B), C.c_int(rows[0]), C.c_int(rows[1]), C.byref(c), C.byref(ms)))
        return c, ms.value

    def free_dia(self, d):
        self.lib.ias_free_dia_dev(C.byref(d))

    def dia_relayout(self, d, to_row_major):
        out = DiaDev()
        self._ck(self.lib.ias_dia_relayout(C.byref(d), C.c_int(1 if to_row_major else 0), C.byref(out)))
        return out

    # -- ELL -----------------------------------------------------------------------------
    def CSRtoELL(self, A, gate=20.0):
        e = EllDev()
        self._ck(self.lib.ias_csr_to_ell(C.byref(A.dev), gate, C.byref(e)))
        return e

    def download_ell(self, e):
        w = e.max_nnz_per_row
        nr = np.zeros(max(e.row, 1), np.int32)
        ci = np.zeros(max(e.row * w, 1), np.int32)
        v = np.zeros(max(e.row * w, 1), np.float64)
        f = self.lib.ias_download_ell64 if isinstance(e, Ell64Dev) else self.lib.ias_download_ell
        self._ck(f(C.byref(e), nr.ctypes.data_as(_I), ci.ctypes.data_as(_I), v.ctypes.data_as(_D)))
        return {"row": e.row, "col": e.col, "width": w, "nnz": e.nnz, "choice": bool(e.choice), "nnz_row": nr[: e.row],
                "col_ind": ci[: e.row * w].reshape(e.row, w), "values": v[: e.row * w].reshape(e.row, w)}

    def ELL_MUL_ELL_DEV(self, A, B, int32=False):
        """ELL_MUL_ELL_DEV (ell_dev:310).  Default: the 64-bit-nnz result; int32=True: the reference's EllMatrixDev."""
        c, ms = (EllDev() if int32 else Ell64Dev()), C.c_double()
        f = self.lib.ias_ell_mul_ell_dev if int32 else self.lib.ias_ell_mul_ell_dev64
        self._ck(f(C.byref(A), C.byref(B), C.byref(c), C.byref(ms)))
        return c, ms.value

    def free_ell(self, e):
        if isinstance(e, Ell64Dev):
            self.lib.ias_free_ell64_dev(C.byref(e))
        else:
            self.lib.ias_free_ell_dev(C.byref(e))

    # -- COO -----------------------------------------------------------------------------
    def CSRtoCOO(self, A):
        c = CooDev()
        self._ck(self.lib.ias_csr_to_coo(C.byref(A.dev), C.byref(c)))
        return c

    def download_coo(self, c):
        wide = isinstance(c, Coo64Dev)
        ro = np.zeros(c.row + 1, np.int64 if wide else np.int32)
        ri = np.zeros(max(c.nnz, 1), np.int32)
        ci = np.zeros(max(c.nnz, 1), np.int32)
        v = np.zeros(max(c.nnz, 1), np.float64)
        if wide:
            self._ck(self.lib.ias_download_coo64(C.byref(c), ro.ctypes.data_as(_L), ri.ctypes.data_as(_I), ci.ctypes.data_as(_I), v.ctypes.data_as(_D)))
        else:
            self._ck(self.lib.ias_download_coo(C.byref(c), ro.ctypes.data_as(_I), ri.ctypes.data_as(_I), ci.ctypes.data_as(_I), v.ctypes.data_as(_D)))
        return {"row": c.row, "col": c.col, "nnz": c.nnz, "row_offset": ro.astype(np.int64), "row_ind": ri[: c.nnz], "col_ind": ci[: c.nnz],
                "values": v[: c.nnz]}

    def COO_MUL_COO_DEV(self, A, B, int32=False):
        """COO_MUL_COO_DEV (coo_dev:279).  Default: the 64-bit result; int32=True: the reference's CooMatrixDev."""
        c, ms = (CooDev() if int32 else Coo64Dev()), C.c_double()
        f = self.lib.ias_coo_mul_coo_dev if int32 else self.lib.ias_coo_mul_coo_dev64
        self._ck(f(C.byref(A), C.byref(B), C.byref(c), C.byref(ms)))
        return c, ms.value

    def free_coo(self, c):
        if isinstance(c, Coo64Dev):
            self.lib.ias_free_coo64_dev(C.byref(c))
        else:
            self.lib.ias_free_coo_dev(C.byref(c))

    # -- features ------------------------------------------------------------------------
    def density_image(self, A):
        img = np.zeros(128 * 128, np.int64)
        self._ck(self.lib.ias_density_image(C.byref(A.dev), img.ctypes.data_as(_L)))
        return img.reshape(128, 128)

    def GetInfo1(self, A):
        f = np.zeros(9)
        self._ck(self.lib.ias_getinfo1(C.byref(A.dev), f.ctypes.data_as(_D)))
        return f

    def GetInfo2(self, rows, cols, nd):
        f = np.zeros(3)
        self._ck(self.lib.ias_getinfo2(rows, cols, nd, f.ctypes.data_as(_D)))
        return f

    def GetInfo3(self, rows, nnz, width):
        f = np.zeros(1)
        self._ck(self.lib.ias_getinfo3(rows, nnz, width, f.ctypes.data_as(_D)))
        return f

    def count_diagonals(self, A):
        n = C.c_int()
        self._ck(self.lib.ias_count_diagonals(C.byref(A.dev), C.byref(n)))
        return n.value

    def max_row_nnz(self, A):
        n = C.c_int()
        self._ck(self.lib.ias_max_row_nnz(C.byref(A.dev), C.byref(n)))
        return n.value

    def features26(self, A, B):
        f = np.zeros(26)
        self._ck(self.lib.ias_features26(C.byref(A.dev), C.byref(B.dev), f.ctypes.data_as(_D)))
        return f

    # -- front end -----------------------------------------------------------------------
    def mtx_load(self, path):
        h = CsrMatrix()
        rc = self.lib.ias_mtx_load(path.encode(), C.byref(h))
        if rc != 0:
            raise IOError("ias_mtx_load(%s) -> %d" % (path, rc))
        rp = np.ctypeslib.as_array(h.row_ind, shape=(h.row + 1,)).copy()
        ci = np.ctypeslib.as_array(h.col_ind, shape=(h.nnz,)).copy() if h.nnz else np.zeros(0, np.int32)
        v = np.ctypeslib.as_array(h.values, shape=(h.nnz,)).copy() if h.nnz else np.zeros(0, np.float64)
        self.lib.ias_free_host_csr(C.byref(h))
        return h.row, h.col, rp, ci, v

    # -- synthetic operands ----------------------------------------------------------------
    def gen_poisson2d(self, nx, ny=None):
        d = CsrMatrixDev()
        self._ck(self.lib.ias_gen_poisson2d(C.c_int(nx), C.c_int(nx if ny is None else ny), C.byref(d)))
        return DeviceCsr(self, d)

    def gen_uniform(self, n, per_row, seed=1):
        d = CsrMatrixDev()
        self._ck(self.lib.ias_gen_uniform(C.c_int(n), C.c_int(per_row), C.c_int(seed), C.byref(d)))
        return DeviceCsr(self, d)

    def gen_rmat(self, scale, edge_factor=16, seed=1, a=0.57, b=0.19, c=0.19):
        d = CsrMatrixDev()
        self._ck(self.lib.ias_gen_rmat(scale, edge_factor, seed, a, b, c, C.byref(d)))
        return DeviceCsr(self, d)


_ENGINE = None


def get_engine(device=0):
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(device)
    return _ENGINE
