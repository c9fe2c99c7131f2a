import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def mtx_dir(tmp_path_factory):
    """Re-materialise the bundled Matrix-Market inputs (tests/golden/bundled_inputs.json)."""
    d = tmp_path_factory.mktemp("inputs")
    with open(os.path.join(GOLDEN, "bundled_inputs.json")) as f:
        inputs = json.load(f)
    for name, m in inputs.items():
        with open(os.path.join(d, name + ".mtx"), "w") as f:
            f.write(m["banner"] + "\n% re-materialised from tests/golden/bundled_inputs.json\n")
            f.write(m["size"] + "\n")
            f.write("\n".join(m["entries"]) + "\n")
    return str(d)


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle, build
    build(ref=False)
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.binding import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libiaref.so not built (reference sources absent)")
    return Ref()


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through its C ABI.  Fails loudly when the library is missing."""
    import ia_spgemm_b200 as ias
    return ias
