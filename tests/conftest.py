import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "transformer-inertial-poser_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")
CKPT_DIR = os.path.join(ROOT, "baseline", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # -m gpu tests must never silently pass on a CPU-only box.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_checkpoint(name):
    """Checkpoint state-dict as numpy arrays, or None when baseline/_ref is not staged."""
    import torch
    path = os.path.join(CKPT_DIR, name + ".pt")
    if not os.path.exists(path):
        return None
    return {k: v.numpy() for k, v in torch.load(path, map_location="cpu").items()}
