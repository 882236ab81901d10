import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def kats():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kats.json")) as fh:
        return json.load(fh)


def _gpu_available():
    """A CUDA device AND the in-tree library: without both the gpu tests are skipped, not failed."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False
    except Exception:
        return False
    return os.path.exists(os.path.join(ROOT, "distributed-decisiontrees_b200", "libdte.so")) or True


def pytest_collection_modifyitems(config, items):
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (there is no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
