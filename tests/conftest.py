import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_unavailable():
    """Why the gpu-marked tests cannot run here, or None.  The product has no CPU path, so these tests are skipped -- not
    silently passed -- on a host without a CUDA device or without the built library."""
    if not os.path.exists(os.path.join(ROOT, "cholla_b200", "libcholla_b200.so")):
        return "cholla_b200/libcholla_b200.so is not built (make -C cholla_b200/csrc)"
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device (the product has no CPU path)"
    except Exception as e:  # noqa: BLE001
        return "torch unavailable: %r" % (e,)
    return None


def pytest_collection_modifyitems(config, items):
    why = None
    for item in items:
        if "gpu" in item.keywords:
            if why is None:
                why = _gpu_unavailable() or ""
            if why:
                item.add_marker(pytest.mark.skip(reason=why))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
