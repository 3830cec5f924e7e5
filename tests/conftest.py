import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("DISABLE_TQDM", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product library and the C oracle exist (nvcc / gcc cross-compile without a GPU)."""
    from bm25s_b200 import _lib

    if not os.path.exists(_lib.SO_PATH):
        _lib.build_extension()
    from oracle import c_oracle

    c_oracle.lib()
    yield


def has_gpu():
    try:
        from bm25s_b200 import _lib

        return _lib.device_count() > 0
    except Exception:
        return False
