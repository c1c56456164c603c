import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build every native artefact once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
    return True
