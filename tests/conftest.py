import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session.  No skip-on-failure: if the CUDA library or the
    GPU is missing the GPU tests must FAIL (there is no fallback to hide behind)."""
    from pylda_b200 import native
    c = native.EStepContext(int(os.environ.get("PYLDA_DEVICE", "0")))
    yield c
    c.close()
