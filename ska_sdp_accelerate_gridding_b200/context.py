"""Context handling: one `skagrid_ctx` per process/GPU, created on first use."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_CTX = {}


class Context:
    """Owns a skagrid_ctx* (device memory pool, streams, cuFFT plans)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.skagrid_create(int(device), C.byref(h))
        if rc != 0:
            raise _lib.SkagridError(rc, self.lib.skagrid_last_error(None).decode())
        self.h = h
        self.device = int(device)

    def close(self):
        if self.h:
            self.lib.skagrid_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            raise _lib.SkagridError(rc, self.lib.skagrid_last_error(self.h).decode())

    @property
    def last_device_ms(self) -> float:
        return float(self.lib.skagrid_last_device_ms(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.skagrid_launch_count(self.h))

    def fp64_tflops(self) -> float:
        out = C.c_double()
        self.check(self.lib.skagrid_measure_fp64_tflops(self.h, C.byref(out)))
        return out.value



    def l2_read_tbs(self, nbytes: int = 8 << 20, pattern: int = 0) -> float:
        """Measured L2 -> SM read bandwidth in TB/s: pattern 0 coalesced stream, 1 random 15x15-tap slices (useful bytes)."""
        out = C.c_double()
        self.check(self.lib.skagrid_measure_l2_pattern_tbs(self.h, int(nbytes), int(pattern), C.byref(out)))
        return out.value


def get_context(device: int | None = None) -> Context:
    if device is None:
        device = 0
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def int64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def ptr(a):
    return None if a is None else a.ctypes.data
