import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once per session if they are missing (the CUDA library cross-compiles without a GPU)."""
    from swiftshader_b200 import capi
    if not os.path.exists(capi.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "libswref.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def device():
    import torch  # noqa: F401  (only for the availability probe; the draw path itself does not use torch)
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swiftshader_b200.scene import Device
    dev = Device(0)
    yield dev
    dev.close()
