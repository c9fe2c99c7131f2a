#!/usr/bin/env python
"""Generate tests/golden/*.json from the REFERENCE itself (run in the build container only).

Sources of truth, all under /root/reference (never copied as source):
  * the reference's unmodified CPU/main.cpp, compiled to /tmp and run in testing mode:
      - its A_csr dump pins the Matrix-Market loader (row pointers + column order),
      - ./imgs/img1.txt pins the density representation for every bundled input;
  * oracle/_ref/libiaref.so (the reference's unmodified headers behind oracle/ref_harness.cpp):
      - CSR_MUL_CSR / MKL_MUL_MKL / DIA_mul_DIA / ELL_MUL_ELL / COO_MUL_COO results, GetFlop,
        the 26 features, sizeofcsr;
  * the density images the reference ships (CPU/imgs, GPU/imgs) -- known-answer files;
  * the bundled Inputs/*.mtx, re-encoded as JSON (entries as text lines) so the GPU box,
    where /root/reference does not exist, can re-materialise them.

Usage:  python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.binding import Oracle, Ref, build  # noqa: E402

REF = "/root/reference"
CPU = os.path.join(REF, "IA-SPGEMM-CPU_release")
GPU = os.path.join(REF, "IA-SPGEMM-GPU_release")
OUT = os.path.dirname(os.path.abspath(__file__))
SQUARE = ["dia", "small", "b1_ss", "LFAT5", "Ragusa18"]
RECT = ["Trec5", "ch3-3-b2", "relat3", "sample"]


def encode_inputs():
    out = {}
    for name in SQUARE + RECT:
        with open(os.path.join(CPU, "Inputs", name + ".mtx")) as f:
            lines = [ln.rstrip("\n") for ln in f]
        banner = lines[0]
        body = [ln for ln in lines[1:] if ln.strip() and not ln.startswith("%")]
        out[name] = {"banner": banner, "size": body[0], "entries": body[1:]}
    return out


def build_refmain(tmp):
    import torch
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    os.makedirs(os.path.join(tmp, "imgs"), exist_ok=True)
    with open(os.path.join(tmp, "MatNet.py"), "w") as f:
        f.write("def Pred(*a):\n    return 1\n")   # keras/tensorflow are absent; the selector is not on this path
    exe = os.path.join(tmp, "refmain")
    cmd = ["g++", "-std=c++11", "-O1", "-fopenmp", "-fpermissive", "-w", "-DVALUE_TYPE=double",
           "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I/usr/include/python3.12", "-I" + CPU,
           os.path.join(CPU, "main.cpp"), "-o", exe, "-L" + lib, "-ltorch_cpu", "-lc10",
           "-Wl,-rpath," + lib, "-lpython3.12", "-lpthread"]
    subprocess.run(cmd, check=True)
    return exe


def run_refmain(exe, tmp, a, b):
    """Returns (row_ptr, col_ind of A as printed, density image of A)."""
    env = dict(os.environ, PYTHONPATH=tmp)
    p = subprocess.run([exe, a, b, "1"], cwd=tmp, env=env, capture_output=True, text=True, timeout=120)
    lines = p.stdout.splitlines()   # the run dies after "DONE MKL" (missing returns in thread_fun_*): ignore rc
    k = lines.index("A_csr:")
    rp = [int(x) for x in lines[k + 2].split(",") if x != ""]
    ci = [int(x) for x in lines[k + 3].split(",") if x != ""]
    img = np.loadtxt(os.path.join(tmp, "imgs", "img1.txt"), dtype=np.int64)
    return rp, ci, img


def sparse_img(img):
    """128x128 int64 image -> base64(zlib(little-endian int32 row-major)); decoded by tests/util.py."""
    import base64
    import zlib
    img = np.asarray(img).reshape(128, 128).astype("<i4")
    return base64.b64encode(zlib.compress(img.tobytes(), 9)).decode()


def main():
    build(ref=True)
    ora, ref = Oracle(), Ref()
    inputs = encode_inputs()
    with open(os.path.join(OUT, "bundled_inputs.json"), "w") as f:
        json.dump(inputs, f, indent=0)

    tmp = "/tmp/ias_refmain"
    exe = build_refmain(tmp)
    gold = {"_generated_by": "tests/golden/make_golden.py", "mkl": ref.mkl_version(), "inputs": {}}
    dia_path = os.path.join(CPU, "Inputs", "dia.mtx")
    for name in SQUARE + RECT:
        path = os.path.join(CPU, "Inputs", name + ".mtx")
        rows, cols, rp, ci, v = ora.mtx_load(path)
        g = {"rows": rows, "cols": cols}
        ref_rp, ref_ci, img = run_refmain(exe, tmp, path, path if name in SQUARE else dia_path)
        g["loader_row_ptr"] = ref_rp
        g["loader_col_ind"] = ref_ci
        g["loader_values"] = [float(x) for x in v]      # values: ours (reference prints 2 decimals only)
        g["density"] = sparse_img(img)
        A = (rows, cols, np.array(ref_rp, np.int32), np.array(ref_ci, np.int32), v)
        if name in SQUARE:
            c_rp, c_ci, c_v, _ = ref.csr_mul_csr(A, A)
            g["flop"] = ref.getflop(A, A)
            g["csr_row_ptr"] = [int(x) for x in c_rp]
            g["csr_col_ind"] = [int(x) for x in c_ci]          # reference order (reverse first touch)
            g["csr_values"] = [float(x) for x in c_v]
            g["sizeof_csr_c"] = ref.sizeof_csr(rows, cols, int(c_rp[-1]))
            m_rp, m_ci, m_v, _, _ = ref.mkl_mul_mkl(A, A)
            g["mkl_row_ptr"] = [int(x) for x in m_rp]
            g["mkl_col_ind"] = [int(x) for x in m_ci]
            g["mkl_values"] = [float(x) for x in m_v]
            g["features26"] = [float(x) for x in ref.features26(A, A)]
            d = ref.dia_mul_dia(A, A)
            if d is not None:
                g["dia_offsets_a"] = [int(x) for x in ref.csr_to_dia(A)["diagonal_offsets"]]
                g["dia_offsets_c"] = [int(x) for x in d["diagonal_offsets"]]
                g["dia_diag_ind_c"] = [int(x) for x in d["diagonal_ind"]]
                g["dia_values_c"] = [float(x) for x in d["values"].ravel()]
            e = ref.ell_mul_ell(A, A)
            if e is not None:
                g["ell_width_c"] = e["width"]
                g["ell_nnz_row_c"] = [int(x) for x in e["nnz_row"]]
                g["ell_col_ind_c"] = [int(x) for x in e["col_ind"].ravel()]
                g["ell_values_c"] = [float(x) for x in e["values"].ravel()]
            c = ref.coo_mul_coo(A, A)
            g["coo_row_offset_c"] = [int(x) for x in c["row_offset"]]
            g["coo_row_ind_c"] = [int(x) for x in c["row_ind"]]
            g["coo_col_ind_c"] = [int(x) for x in c["col_ind"]]
            g["coo_values_c"] = [float(x) for x in c["values"]]
        gold["inputs"][name] = g

    # density images the reference ships: GPU/imgs <-> dia.mtx / dia^T, CPU/imgs <-> small.mtx / small^T
    gold["shipped_imgs"] = {
        "gpu_img1_dia": sparse_img(np.loadtxt(os.path.join(GPU, "imgs", "img1.txt"), dtype=np.int64)),
        "gpu_img2_diaT": sparse_img(np.loadtxt(os.path.join(GPU, "imgs", "img2.txt"), dtype=np.int64)),
        "cpu_img1_small": sparse_img(np.loadtxt(os.path.join(CPU, "imgs", "img1.txt"), dtype=np.int64)),
        "cpu_img2_smallT": sparse_img(np.loadtxt(os.path.join(CPU, "imgs", "img2.txt"), dtype=np.int64)),
    }
    # the feature vector printed in CPU/1.jpg and GPU/2.jpg for dia.mtx (transcribed in SURVEY section 4)
    gold["screenshot_features_dia"] = ([4, 4, 7, 0.4375, 2, 1, 1.75, 0.25, 0.2857142857142857] * 2
                                       + [2, 0.2857142857142857, 0.5] * 2 + [0.875, 0.875])
    # GPU/2.jpg: the GPU program multiplied dia.mtx by its transpose: nnz(C)=10, memory_size 152, 13 products
    rows, cols, rp, ci, v = ora.mtx_load(dia_path)
    import scipy.sparse as sp
    At = sp.csr_matrix((v, ci, rp), shape=(rows, cols)).T.tocsr()
    A = (rows, cols, rp, ci, v)
    B = (cols, rows, At.indptr.astype(np.int32), At.indices.astype(np.int32), At.data)
    c_rp, c_ci, c_v, _ = ref.csr_mul_csr(A, B)
    gold["dia_times_diaT"] = {"nnz": int(c_rp[-1]), "flop": ref.getflop(A, B),
                              "sizeof_csr": ref.sizeof_csr(rows, rows, int(c_rp[-1])),
                              "b_row_ptr": [int(x) for x in B[2]], "b_col_ind": [int(x) for x in B[3]]}
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(gold, f)
    print("wrote", os.path.join(OUT, "golden.json"), os.path.getsize(os.path.join(OUT, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
