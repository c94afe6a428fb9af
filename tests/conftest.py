import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libdmx.so once per session when sources are newer (no-op on the GPU box: the .so travels)."""
    from vqe_errormitigation_b200 import build
    try:
        build.build()
    except Exception:
        if not os.path.exists(build.OUT):
            raise
    yield
