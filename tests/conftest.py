import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle / emulator (test infrastructure) and the product library if stale."""
    from rocoder_b200 import build

    build.build_oracle()
    build.build_emulator()
    if not os.path.exists(build.LIB) or os.environ.get("RCG_REBUILD"):
        build.build_product()
    yield
