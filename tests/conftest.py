import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One b200rt context on cuda:0 for the whole GPU session; fails loudly if it cannot be made."""
    import ensem3a_openclraytracer_b200 as rt
    ctx = rt.Context(0)
    yield ctx
    ctx.close()
