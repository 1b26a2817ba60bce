import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "golden_vectors.json")))


@pytest.fixture(scope="session")
def built():
    """Make sure every native artefact exists (oracle, reference lib when sources are present, product .so)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def orc(built):
    import refshim
    return refshim.OracleLib()


@pytest.fixture(scope="session")
def ref(built):
    import refshim
    if not refshim.ref_available():
        pytest.skip("oracle/_ref/libmzref.so not built (reference sources absent)")
    return refshim.RefLib()
