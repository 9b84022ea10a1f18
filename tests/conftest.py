import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built():
    """the CUDA library and the C oracle are built in-tree (nvcc / gcc cross-compile without a GPU)"""
    import __graft_entry__ as g
    from ctpa_clip_b200 import _lib
    if not _lib.LIB_PATH.exists():
        g.build()
    from oracle import build_oracle
    build_oracle.build()
