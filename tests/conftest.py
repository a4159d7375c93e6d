import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "nuclei-feature-extraction_b200"), os.path.join(ROOT, "oracle"), ROOT,
          os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run through gpurun); everything else is CPU-only")


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
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def libnfx():
    """Build (if stale) and load libnfx.so. The product has no CPU fallback: a missing library is an error."""
    sys.path.insert(0, os.path.join(ROOT, "nuclei-feature-extraction_b200"))
    import build as nfx_build
    nfx_build.build_lib()
    from nfx._lib import lib
    return lib()
