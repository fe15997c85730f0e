import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native_lib():
    """Builds (if needed) and loads libnesvor_b200.so; never falls back."""
    from nesvor_b200.csrc import build as nsv_build
    from nesvor_b200 import _lib

    nsv_build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def oracle():
    from oracle import native

    return native.Oracle()


@pytest.fixture(scope="session")
def reference_cpu():
    from oracle import native

    ref = native.Reference()
    if ref is None:
        pytest.skip("oracle/_ref not available (needs /root/reference or a prebuilt copy)")
    return ref
