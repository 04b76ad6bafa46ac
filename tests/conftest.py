import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    so = os.path.join(ROOT, "uzliti_slam_b200", "libuzliti_edge.so")
    if not os.path.exists(so) or not os.path.exists(os.path.join(ROOT, "oracle", "libuz_oracle.so")) or \
            not os.path.exists(os.path.join(ROOT, "tests", "host_shim", "libhost_shim.so")):
        g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def est(built):
    """The CUDA path.  Fails loudly (no skip, no fallback) when the library or the GPU is missing."""
    from uzliti_slam_b200 import EdgeEstimator
    e = EdgeEstimator(0)
    yield e
    e.close()
