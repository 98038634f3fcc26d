import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """One rfb_ctx for the whole GPU session.  Fails loudly (no skip) when the device is missing."""
    import rfb200
    c = rfb200.Context(int(os.environ.get("LOCAL_RANK", "0")))
    yield c
    c.close()
