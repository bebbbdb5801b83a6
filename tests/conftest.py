import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from nrays_b200 import _lib

        return _lib.load().nrb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a device must fail loudly, not skip: the product has no CPU path.
    # A GPU test that hangs (a kernel that never ends) must not hold the box: pytest-timeout's thread method ends the
    # process, which tears the CUDA context down.  No GPU test needs more than a minute; the limit is generous.
    if config.pluginmanager.hasplugin("timeout"):
        for item in items:
            if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(600, method="thread"))


@pytest.fixture(scope="session")
def gpu():
    from nrays_b200 import _lib

    lib = _lib.load()
    assert lib.nrb_device_count() > 0, "no CUDA device: -m gpu tests must run on the B200 box"
    return lib
