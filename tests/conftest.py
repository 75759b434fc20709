import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def port():
    from oracle.bindings import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def reference():
    from oracle import bindings
    if not bindings.available("reference") and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libshs_ref.so not built and /root/reference absent")
    return bindings.Oracle("reference")


@pytest.fixture(scope="session")
def gpu():
    from leisure_software_renderer_b200 import build
    build.build()
    from leisure_software_renderer_b200.renderer import Context
    ctx = Context(0)  # raises if there is no CUDA device: there is no CPU fallback
    yield ctx
    ctx.close()
