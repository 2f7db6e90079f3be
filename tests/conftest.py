import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The built libraries are git-ignored: on a fresh checkout build them once (nvcc cross-compiles without a GPU).
    # This only builds; if nvcc is missing the ABI / GPU tests fail loudly, there is no fallback.
    try:
        from ska_sdp_accelerate_gridding_b200 import _lib, build as _build
        if not os.path.exists(_lib.LIB_PATH):
            _build.build()
    except Exception as e:  # pragma: no cover
        print(f"[conftest] libskagrid.so could not be built: {e}", file=sys.stderr)


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is present, so a bare `pytest tests/`
    works in the CPU-only dev container.  On a GPU box a missing libskagrid.so is a hard failure."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
