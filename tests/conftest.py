import os
import sys

import pytest

# Every device buffer of the library gets canaries while the tests run (astre_gpu_debug_guard_check; GpuContext.close
# raises when a kernel stored outside its buffer).  Must be set before the library reads it.
os.environ.setdefault("ASTRE_GUARD", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library must be the thing that runs: no fallback, loud failure."""
    import astre_b200
    return astre_b200.load_library()
