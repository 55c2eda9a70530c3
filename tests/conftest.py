import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for both tiers (CPU tier checks loading/symbols)."""
    import __graft_entry__ as g
    g.build_native()
    g.build_oracle()
    yield


@pytest.fixture(scope="session")
def golden():
    return GOLDEN
