import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU check")


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_lib import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def gpu_lib():
    import sad_monte_carlo_b200 as pkg
    return pkg.load_library()


def pytest_collection_modifyitems(config, items):
    # every GPU test is bounded: a hung kernel must fail the test, not eat the GPU budget
    for item in items:
        if "gpu" in item.keywords and not any(m.name == "timeout" for m in item.iter_markers()):
            item.add_marker(pytest.mark.timeout(300))
