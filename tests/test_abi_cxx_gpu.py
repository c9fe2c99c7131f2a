"""The C ABI from a compiled C++ caller (tests/abi_cxx/abi_driver.cpp): no ctypes, no Python in the data path.
Known answers: dia.mtx A^2 (CSR / DIA / ELL / COO), the 26 features, Poisson 64^2 through the streaming entry with a
consumer callback."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cxx_caller_runs(tmp_path):
    from ia_spgemm_b200 import engine as E
    out = tmp_path / "abi_driver"
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-o", str(out), os.path.join(ROOT, "tests", "abi_cxx", "abi_driver.cpp"),
                        "-L", os.path.dirname(E.LIB_PATH), "-liaspgemm", "-Wl,-rpath," + os.path.dirname(E.LIB_PATH)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ABI_CXX_OK" in r.stdout, r.stdout + r.stderr
