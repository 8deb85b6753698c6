import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """librt_b200.so, built in-tree if necessary (nvcc cross-compiles without a GPU)."""
    from hslu_i.ba_raytracing.f2501_raytracer_b200 import _lib, build

    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle

    oracle.load()
    return oracle


@pytest.fixture(scope="session")
def pkg(lib):
    import hslu_i.ba_raytracing.f2501_raytracer_b200 as p

    return p
