import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native():
    """Builds (if stale) and loads the native library; never falls back to anything else."""
    from nbodyhpc_b200 import _build, capi

    _build.build_all()
    return capi


@pytest.fixture(scope="session")
def gpu(native):
    if native.lib().nbk_device_count() <= 0:
        pytest.fail("test marked gpu but no CUDA device is visible: " + native.lib().nbk_last_error().decode())
    return native
