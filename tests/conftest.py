import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def set8():
    from libmultirobotplanning_b200 import instances
    return instances.load_set(os.path.join(GOLDEN, "bench_8x8.npz"))


@pytest.fixture(scope="session")
def set32():
    from libmultirobotplanning_b200 import instances
    return instances.load_set(os.path.join(GOLDEN, "bench_32x32.npz"))


@pytest.fixture(scope="session")
def ref_fixtures():
    with open(os.path.join(GOLDEN, "ref_fixtures.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_golden():
    with open(os.path.join(GOLDEN, "oracle_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def capi():
    """The CUDA library, initialised on cuda:0 (GPU tests only)."""
    from libmultirobotplanning_b200 import capi as _capi
    _capi.init(0)
    return _capi
