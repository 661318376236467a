import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # pytest-timeout is in the image; registering the marker keeps `--strict-markers` runs and plugin-less runs quiet
    config.addinivalue_line("markers", "timeout(seconds): per-test limit (pytest-timeout; ignored when the plugin is absent)")


@pytest.fixture(scope="session")
def rtp():
    import rtp_b200

    return rtp_b200


@pytest.fixture(scope="session")
def gpu(rtp):
    """Initialise device 0; a GPU test that reaches this without a usable device must FAIL, not skip."""
    rtp.api.init(0)
    return rtp
