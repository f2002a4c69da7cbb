"""pytest configuration: the `gpu` marker, path setup and the shared builders of the test-only native shims."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.join(ROOT, "tests")
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def rbd_lib():
    """librbd_b200.so built in-tree (nvcc cross-compiles without a GPU)."""
    from sofa.rigidbodydynamics_b200 import build, _capi
    build.build()
    return _capi.load()


@pytest.fixture(scope="session")
def emul():
    import helpers
    return helpers.load_emul()
