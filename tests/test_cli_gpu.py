"""The spgemm-gpu front end on the reference's bundled inputs: report block, density files, exit codes."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from util import SQUARE, decode_img

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ia_spgemm_b200", "spgemm-gpu")


def _run(args, cwd):
    return subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=300)


@pytest.mark.parametrize("name", SQUARE)
def test_cli_report_matches_reference_numbers(golden, mtx_dir, tmp_path, name):
    g = golden["inputs"][name]
    r = _run([os.path.join(mtx_dir, name + ".mtx"), "--all", "--json"], str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "The Chosen One = Algorithm" in out and "MAX SPEED IS" in out
    blocks = re.findall(r"Algorithm (\d):\nrun_time: (\S+)\ntrans_time: (\S+)\nmemory_size: (\S+)\nverified_sum: (\S+)\nGflops: (\S+)\nSpeedup: (\S+)", out)
    assert [b[0] for b in blocks] == ["1", "2", "3", "4", "5"]
    j = json.loads(out.strip().splitlines()[-1])
    want_sum = float(np.sum(np.array(g["csr_values"], dtype=np.float64)))
    assert j["products"] == g["flop"]
    assert j["features"] == pytest.approx(g["features26"], rel=1e-12)      # the gate (20x vs 50x) changes `choice`, never a feature
    assert j["memory_size"][1] == g["sizeof_csr_c"]                                 # Algorithm 2 = CSR
    assert j["verified_sum"][1] == pytest.approx(want_sum, rel=1e-11, abs=1e-9)
    assert j["verified_sum"][4] == pytest.approx(want_sum, rel=1e-11, abs=1e-9)     # COO stores the same entries
    if "dia_values_c" in g and j["run_ms"][2] > 0:
        assert j["verified_sum"][2] == pytest.approx(float(np.sum(g["dia_values_c"])), rel=1e-11, abs=1e-9)
    if "ell_values_c" in g and j["run_ms"][3] > 0:
        assert j["verified_sum"][3] == pytest.approx(float(np.sum(g["ell_values_c"])), rel=1e-11, abs=1e-9)
    img1 = np.loadtxt(os.path.join(tmp_path, "imgs", "img1.txt"), dtype=np.int64).reshape(128, 128)
    assert np.array_equal(img1, decode_img(g["density"]))
    assert np.array_equal(np.loadtxt(os.path.join(tmp_path, "imgs", "img2.txt"), dtype=np.int64).reshape(128, 128), img1)


@pytest.mark.parametrize("name", ["LFAT5", "Ragusa18"])
def test_cli_gpus_flag_deals_the_rows_to_one_process_per_gpu(mtx_dir, tmp_path, name):
    """--gpus 3: two helper processes plus the main one, each multiplying its snake-dealt share of the rows (on a box
    with fewer devices the helpers share device 0): nnz (memory_size), products and verified_sum equal the one-GPU run."""
    path = os.path.join(mtx_dir, name + ".mtx")
    one = _run([path, "--json", "--no-matnet", "--all"], str(tmp_path))
    three = _run([path, "--json", "--no-matnet", "--gpus", "3"], str(tmp_path))
    assert one.returncode == 0 and three.returncode == 0, three.stdout + three.stderr
    j1, j3 = json.loads(one.stdout.strip().splitlines()[-1]), json.loads(three.stdout.strip().splitlines()[-1])
    assert j3["gpus"] == 3 and j1["gpus"] == 1
    shares = re.findall(r"share (\d) on device (\d): (\d+) rows, (\d+) products, nnz (\d+)", three.stdout)
    assert [s[0] for s in shares] == ["0", "1", "2"]
    assert sum(int(s[2]) for s in shares) == j3["rows"] and sum(int(s[3]) for s in shares) == j3["products"] == j1["products"]
    assert "DONE CSR (rows dealt to 3 processes" in three.stdout
    assert j3["memory_size"][1] == j1["memory_size"][1]
    assert j3["verified_sum"][1] == pytest.approx(j1["verified_sum"][1], rel=1e-12, abs=1e-9)


def test_cli_dia_known_answer(mtx_dir, tmp_path):
    r = _run([os.path.join(mtx_dir, "dia.mtx"), "--all"], str(tmp_path))
    assert r.returncode == 0
    assert "memory_size: 140.000000" in r.stdout and "verified_sum: 12.000000" in r.stdout


def test_cli_errors(tmp_path):
    assert _run([], str(tmp_path)).returncode != 0
    r = _run([str(tmp_path / "missing.mtx")], str(tmp_path))
    assert r.returncode == 255 and "could not load" in r.stdout          # main() returns -1
    p = tmp_path / "cplx.mtx"
    p.write_text("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0.0\n")
    assert _run([str(p)], str(tmp_path)).returncode == 253                # -3


def test_cli_transpose_b_and_write_c(golden, mtx_dir, tmp_path):
    """--transpose-b reproduces the GPU release's operand (B := A^T): dia.mtx gives the screenshot's 152 bytes."""
    out = str(tmp_path / "c.mtx")
    r = _run([os.path.join(mtx_dir, "dia.mtx"), "--transpose-b", "--write-c", out, "--json"], str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["products"] == 13 and j["memory_size"][1] == 152.0
    lines = open(out).read().strip().splitlines()
    assert lines[0].startswith("%%MatrixMarket matrix coordinate real general") and lines[1].split() == ["4", "4", "10"]
    assert len(lines) == 12
    r = _run([os.path.join(mtx_dir, "Trec5.mtx"), "--transpose-b", "--all"], str(tmp_path))     # rectangular 3x7: A*A^T is 3x3
    assert r.returncode == 0 and "DONE CSR" in r.stdout


def test_cli_matnet_selection(mtx_dir, tmp_path):
    """--matnet runs the native MatNet on the density images + features (weights: any Keras-2.1 MatNet file).
    The reference's weight files are not in this repository, so the test writes nothing and only runs when the
    reference checkout is present (build container); the forward pass itself is pinned in tests/test_matnet.py."""
    w = "/root/reference/IA-SPGEMM-CPU_release/NetWeights/Intel_weights.h5"
    if not os.path.exists(w):
        pytest.skip("reference checkout not present on this box")
    r = _run([os.path.join(mtx_dir, "dia.mtx"), "--matnet", w], str(tmp_path))
    assert r.returncode == 0 and "MatNet class" in r.stdout and "The Chosen One = Algorithm" in r.stdout


def test_cli_loads_netweights_by_default(mtx_dir, tmp_path):
    """The reference always loads ./NetWeights/<machine>_weights.h5 from the working directory (CPU/MatNet.py:81,
    GPU/MatNet.py); so does spgemm-gpu when the files are there -- no flag needed.  With the shipped Intel weights the
    pick for dia.mtx is the screenshot's "Algorithm 3" (CPU/1.jpg)."""
    import shutil
    src = "/root/reference/IA-SPGEMM-CPU_release/NetWeights"
    if not os.path.exists(os.path.join(src, "Intel_weights.h5")):
        pytest.skip("reference checkout not present on this box")
    os.makedirs(tmp_path / "NetWeights")
    for n in ("Intel_weights.h5", "P100_weights.h5"):
        if os.path.exists(os.path.join(src, n)):
            shutil.copy(os.path.join(src, n), tmp_path / "NetWeights" / n)
    r = _run([os.path.join(mtx_dir, "dia.mtx")], str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MatNet class 2 of 5" in r.stdout and "The Chosen One = Algorithm 3" in r.stdout and "DONE DIA" in r.stdout
    assert "MatNet predicts Algorithm" in r.stdout                     # the 3-class GPU net's line (GPU/main.cu:543-544)
    r = _run([os.path.join(mtx_dir, "dia.mtx"), "--no-matnet"], str(tmp_path))
    assert r.returncode == 0 and "MatNet class" not in r.stdout


def test_cli_streams_when_asked_or_when_c_does_not_fit(mtx_dir, tmp_path):
    """--stream produces the CSR result in row batches through the consumer callback: same report numbers and the
    same .mtx output as the materialised run."""
    a = os.path.join(mtx_dir, "Ragusa18.mtx")
    o1, o2 = str(tmp_path / "c1.mtx"), str(tmp_path / "c2.mtx")
    r1 = _run([a, "--no-matnet", "--write-c", o1, "--json"], str(tmp_path))
    r2 = _run([a, "--no-matnet", "--write-c", o2, "--json", "--stream", "0.000001"], str(tmp_path))
    assert r1.returncode == 0 and r2.returncode == 0, r2.stdout + r2.stderr
    assert "DONE CSR (streamed in" in r2.stdout and "DONE CSR\n" in r1.stdout
    j1, j2 = (json.loads(r.stdout.strip().splitlines()[-1]) for r in (r1, r2))
    assert j1["memory_size"][1] == j2["memory_size"][1] and j1["verified_sum"][1] == pytest.approx(j2["verified_sum"][1], rel=1e-12)
    assert open(o1).read() == open(o2).read()
