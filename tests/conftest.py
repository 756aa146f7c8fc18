import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import ko

    ko.build()
    return ko


@pytest.fixture(scope="session")
def kt():
    """The product binding.  GPU tests fail loudly if the CUDA library is missing -- no fallback."""
    import kube_throttler_b200 as kt

    if not os.path.exists(kt.LIB_PATH):
        kt.build()
    kt.lib()
    return kt
