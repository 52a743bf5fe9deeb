import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
# run-time specialised kernels (csrc/ox_jit.cpp) compiled by the tests are cached in-tree: build/ is git-ignored but
# travels to the GPU box with the snapshot, so the GPU tests load the cubins the CPU tests compiled here
os.environ.setdefault("OX_B200_CACHE_DIR", os.path.join(ROOT, "build", "jit_cache"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ox():
    import oxide_control_b200
    return oxide_control_b200
