import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref/libswegl_ref.so); prebuilt in the build container."""
    from oracle.binding import Ref, REF_LIB
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libswegl_ref.so not built (needs /root/reference; run `make -C oracle ref`)")
    return Ref()


@pytest.fixture(scope="session")
def renderer():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swegl_b200 import Renderer
    r = Renderer(0)
    yield r
    r.close()
