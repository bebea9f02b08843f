import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The C oracle (test infrastructure), built on demand with gcc."""
    from oracle import canonical
    canonical.build()
    return canonical


@pytest.fixture(scope="session")
def ctx():
    """One native context on cuda:0 for the whole GPU session."""
    from speaker_diarization_toolkit_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()
