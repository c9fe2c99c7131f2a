"""The spgemm-gpu front end without a device: argument handling, loader exit codes (the reference's -1 / -3,
CPU/main.cpp:156-180) and the loud failure where the reference's GPU program would call exit(1) -- never a CPU fallback."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ia_spgemm_b200", "spgemm-gpu")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(CLI):
        import __graft_entry__ as g
        g.build()


def _run(args, cwd):
    return subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=120)


def test_help_lists_every_flag(tmp_path):
    r = _run(["--help"], str(tmp_path))
    assert r.returncode == 0
    for flag in ("--all", "--json", "--gate", "--repeat", "--transpose-b", "--write-c", "--matnet", "--opt", "--gpus", "--stream"):
        assert flag in r.stdout


def test_usage_and_loader_errors_need_no_device(tmp_path):
    r = _run([], str(tmp_path))
    assert r.returncode == 255 and "please use command like this" in r.stdout          # the reference's message, main() returns -1
    r = _run([str(tmp_path / "missing.mtx")], str(tmp_path))
    assert r.returncode == 255 and "could not load" in r.stdout
    p = tmp_path / "cplx.mtx"
    p.write_text("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0.0\n")
    assert _run([str(p)], str(tmp_path)).returncode == 253                               # -3: complex is rejected


def test_opt_flag_is_validated_before_any_work(tmp_path):
    r = _run(["--opt", "no_such_knob=1", "x.mtx"], str(tmp_path))
    assert r.returncode == 250 and "unknown option 'no_such_knob'" in r.stdout          # -6
    r = _run(["--opt", "gwin_win", "x.mtx"], str(tmp_path))
    assert r.returncode == 250 and "NAME=VALUE" in r.stdout
    r = _run(["--opt", "gwin_win=-3", "x.mtx"], str(tmp_path))
    assert r.returncode == 250 and "negative" in r.stdout


def test_gpus_flag_is_validated_before_any_work(tmp_path):
    r = _run(["--gpus", "0", "x.mtx"], str(tmp_path))
    assert r.returncode == 250 and "--gpus expects" in r.stdout                          # -6
    r = _run(["--gpus", "2", "--write-c", "c.mtx", "x.mtx"], str(tmp_path))
    assert r.returncode == 250 and "--write-c" in r.stdout
    r = _run(["--gpus", "2", str(tmp_path / "missing.mtx")], str(tmp_path))              # the loader runs before any process is started
    assert r.returncode == 255 and "could not load" in r.stdout


def test_shape_mismatch_is_reported_before_the_device_is_touched(mtx_dir, tmp_path):
    r = _run([os.path.join(mtx_dir, "sample.mtx"), os.path.join(mtx_dir, "Trec5.mtx")], str(tmp_path))    # 8x5 times 3x7: A has more columns than B has rows
    assert r.returncode == 251 and "shape mismatch" in r.stdout                          # -5


def test_no_cpu_fallback(mtx_dir, tmp_path):
    """Without a CUDA device the program says so and fails; with one this test has nothing to check."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run([os.path.join(mtx_dir, "dia.mtx")], str(tmp_path))
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and "Algorithm" not in r.stdout
    r = _run([os.path.join(mtx_dir, "dia.mtx"), "--gpus", "2"], str(tmp_path))          # nor with one process per GPU
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and "Algorithm" not in r.stdout
