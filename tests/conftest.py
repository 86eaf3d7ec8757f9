import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_built():
    import oracle_api
    oracle_api.build(ref=True)
    return True


@pytest.fixture(scope="session")
def lib_built():
    import tps_b200
    if not os.path.exists(tps_b200.library_path()):
        tps_b200.build_library()
    return tps_b200.lib()
