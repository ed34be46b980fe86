import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) every native library that can be built on this machine."""
    from stencilstream_b200 import _build
    return _build.build_all()


@pytest.fixture(scope="session")
def oracle_best(built):
    import oracle
    return oracle.best()


@pytest.fixture(scope="session")
def oracle_port(built):
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def oracle_ref(built):
    import oracle
    ref = oracle.reference()
    if ref is None:
        pytest.skip("reference-built oracle (oracle/_ref) not available on this machine")
    return ref
