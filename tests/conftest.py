import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

if os.path.join(ROOT, "examples") not in sys.path:  # the demo drivers (known-answer problems) live with the examples
    sys.path.insert(0, os.path.join(ROOT, "examples"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        from dolfinx_external_operator_b200 import _lib

        return _lib.load().eo_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ctx():
    import dolfinx_external_operator_b200 as eo

    c = eo.Context(0)
    yield c
    c.sync()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
