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
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def ref():
    """The reference's own CUDA ops (oracle/_ref/libvolrend_ref.so), GPU only."""
    from tests import refops

    if not refops.available():
        pytest.skip("oracle/_ref/libvolrend_ref.so not built (needs /root/reference at build time)")
    return refops


@pytest.fixture(scope="session")
def built_lib():
    """libngp_b200.so, cross-compiled on first use (nvcc needs no GPU); CPU tests only look at its symbols."""
    from jaxngp_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()
