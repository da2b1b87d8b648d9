import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): our restatement, fp32 + fp64 twin."""
    from oracle import oracle as orc
    return orc.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference itself compiled headless (oracle/_ref); skipped where it was never built."""
    from oracle import oracle as orc
    if os.path.isdir("/root/reference"):
        orc.build(ref=True)
    r = orc.Ref()
    if not r.available:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return r


@pytest.fixture(scope="session")
def heightmap():
    from tests.common import reference_heightmap
    return reference_heightmap()
