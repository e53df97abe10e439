import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """CPU checkers + the CUDA library (cross-compiled; loads without a GPU)."""
    import oracle
    oracle.build()
    from mapf_gpt_b200 import _lib
    if not _lib.LIB_PATH.exists():
        _lib.build()
    return _lib.lib()
