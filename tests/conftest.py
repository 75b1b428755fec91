import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


def _no_device() -> bool:
    """True only when the C ABI itself reports RCG_ERR_NO_DEVICE (no CUDA device / driver, or not an sm_100
    part).  Any other failure -- a missing library, a CUDA error on a B200 -- must stay loud: the tests then run
    and fail instead of being skipped."""
    try:
        import ctypes as C
        from rocoder_b200 import _abi
        cfg = _abi.make_config(1, 48000, 1024)
        h = C.c_void_p()
        rc = _abi.lib().rcg_create(C.byref(cfg), C.byref(h))
        if rc == _abi.RCG_OK:
            _abi.lib().rcg_destroy(h)
        return rc == _abi.RCG_ERR_NO_DEVICE
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a B200 skips the gpu-marked tests instead of failing in the first one"""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or not _no_device():
        return
    skip = pytest.mark.skip(reason="rcg_create: RCG_ERR_NO_DEVICE -- GPU tests need a B200")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle / emulator (test infrastructure) and the product library if stale."""
    import shutil
    from rocoder_b200 import build

    build.build_oracle()
    build.build_emulator()
    if (not os.path.exists(build.LIB) or os.environ.get("RCG_REBUILD")) and (
            shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        build.build_product()
    yield
