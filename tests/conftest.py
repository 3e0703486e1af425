import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_cpu():
    """The CPU oracle library, built on demand (test infrastructure, never the product path)."""
    import subprocess

    from oracle import binding as B

    if not os.path.exists(B.CPU_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "cpu"])
    return B


@pytest.fixture(scope="session")
def renderer():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sunvolumerender_b200.render import Renderer

    r = Renderer(0)
    yield r
    r.close()
