"""skagrid-b200: B200-native (sm_100a) AW-projection gridding hot path behind the API of
sakehl/SKA-SDP-Accelerate-gridding's src/Gridding.hs and src/ImageDataset.hs.

Importing this package does not load the CUDA library; the first call does, and fails loudly if
libskagrid.so is missing or no CUDA device is present (there is no CPU fallback).
"""
from . import _lib  # noqa: F401
from .context import Context, get_context  # noqa: F401

__all__ = ["Context", "get_context", "gridding", "image_dataset", "device", "distributed"]
