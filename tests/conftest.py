import os
import sys

os.environ.setdefault("RUNMAT_B200_OZAKI_CHECK", "1")  # surface pipeline-protocol errors of the tcgen05 kernel in tests
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Builds (or reuses) the CUDA library and the oracle. nvcc cross-compiles without a GPU."""
    from runmat_b200 import build as b

    b.build_library()
    b.build_oracle()
    return True


@pytest.fixture(scope="session")
def orc(built):
    import oracle_binding

    return oracle_binding.Oracle()


@pytest.fixture(scope="module")
def prov(built):
    from runmat_b200 import B200Provider

    p = B200Provider(0, device_id=7, precision="f64")
    yield p
    p.close()


@pytest.fixture(scope="module")
def prov32(built):
    from runmat_b200 import B200Provider

    p = B200Provider(0, device_id=8, precision="f32")
    yield p
    p.close()
