"""bench.py's contract where it can be checked without a GPU: the reference arm (the reference's CPU path on the host
cores) prints exactly one JSON line with the agreed keys, alone and under a 2-rank launch; the engine arm refuses to run
without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
METRIC_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
               "dtype", "data", "config", "cpu_baseline", "e2e")


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


def _check_reference_line(d, n_gpus, steps, warmup):
    for k in METRIC_KEYS + ("impl",):
        assert k in d, k
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["steps"] == steps and d["warmup"] == warmup
    assert d["unit"] == "GFLOP/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                                     # BASELINE.md holds no published number for this metric
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("poisson2d_5pt_")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "2", "--warmup", "1", "--grid", "192"],
                       capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1 and len([l for l in r.stdout.splitlines() if l.strip()]) == 1      # stdout carries the line and nothing else
    _check_reference_line(lines[0], 1, 2, 1)
    assert lines[0]["config"]["workload"] == "poisson2d_5pt_192x192_A2_fp64"


@pytest.mark.timeout(300)
def test_reference_arm_under_a_two_rank_launch():
    """Launched as the driver launches N > 1: rank 0 alone runs and prints; the other rank exits 0 without work.  The
    workload is the N x grid the engine arm would run (a bounded sample of it when it is large)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29600 + os.getpid() % 300), BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--grid", "128"], capture_output=True, text=True, timeout=280, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 2, 1, 0)
    assert lines[0]["config"]["workload"] == "poisson2d_5pt_128x256_A2_fp64"
    assert lines[0]["cpu_baseline"]["cores"] >= 1               # the launcher's OMP_NUM_THREADS=1 is overridden explicitly


def test_engine_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1", "--grid", "64"], capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not _json_lines(r.stdout)
