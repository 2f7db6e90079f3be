"""Single-process multi-GPU gridding through the C ABI (`skagrid_*_mgpu_{vis,tile}`, csrc/mgpu.cu): one host thread
drives one context per device.  This is the route a Haskell caller of the reference takes (it has no
torch.distributed); `distributed.py` is the one-process-per-GPU equivalent used by bench.py.

Both modes are BASELINE.json's: "vis" = visibility-sharded + reduce (config 4), "tile" = uv-tile-sharded + routing
(config 5).  Same argument conventions as gridding.convgrid2 / convdegrid2 (src/Gridding.hs:199-244)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .context import Context, c128, int64, ptr
from .gridding import _uv


class MultiDevice:
    """`devices`: CUDA device ordinals, one context each (repeats are allowed: several contexts on one device)."""

    def __init__(self, devices):
        self.lib = _lib.load()
        self.ctxs = [Context(int(d)) for d in devices]
        self.handles = (C.c_void_p * len(self.ctxs))(*[c.h for c in self.ctxs])
        self.bounds = None
        self._resident_shape = None

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def _check(self, rc):
        self.ctxs[0].check(rc)

    @property
    def last_device_ms(self):
        return self.ctxs[0].last_device_ms

    def _table(self, gcf):
        gcf = c128(gcf)
        if gcf.ndim == 4:
            gcf = gcf[None]
        if gcf.ndim != 5 or gcf.shape[1] != gcf.shape[2]:
            raise ValueError("gcf must be [nw,qpx,qpx,gh,gw] (or [qpx,qpx,gh,gw])")
        return gcf

    def convgrid2(self, gcf, a, p, wbin, v, mode="vis"):
        """a + sum_k v_k * gcf[wbin_k, yf_k, xf_k] over all devices; returns the new grid."""
        gcf = self._table(gcf)
        a = c128(a).copy()
        u, vv = _uv(p)
        vis = c128(v)
        wb = None if wbin is None else int64(wbin)
        nw, qpx, _, gh, gw = gcf.shape
        args = (self.handles, len(self.ctxs), nw, qpx, gh, gw, ptr(gcf), a.shape[0], a.shape[1], ptr(a), u.size, ptr(u), ptr(vv), ptr(wb),
                ptr(vis))
        if mode == "vis":
            self._check(self.lib.skagrid_convgrid2_mgpu_vis(*args))
            self._resident_shape = a.shape
        elif mode == "tile":
            b = np.zeros(len(self.ctxs) + 1, np.int64)
            self._check(self.lib.skagrid_convgrid2_mgpu_tile(*args, ptr(b)))
            self.bounds = b.tolist()
        else:
            raise ValueError("mode must be 'vis' or 'tile'")
        return a

    def convdegrid2(self, gcf, a, p, wbin, mode="vis", count=None):
        """Adjoint of convgrid2.  Mode 'vis' only: a=None = the grid the previous 'vis' call left on the devices;
        p=None, wbin=None, count=n = at the coordinates that call uploaded (nothing but the results crosses PCIe)."""
        gcf = self._table(gcf)
        if p is None:
            if mode != "vis" or count is None:
                raise ValueError("p=None needs mode 'vis' and count")
            u = vv = wb = None
            n = int(count)
        else:
            u, vv = _uv(p)
            wb = None if wbin is None else int64(wbin)
            n = u.size
        nw, qpx, _, gh, gw = gcf.shape
        if a is None:
            if mode != "vis":
                raise ValueError("a=None needs mode 'vis'")
            shape, pa = self._resident_shape, None
        else:
            a = c128(a)
            shape, pa = a.shape, ptr(a)
        out = np.empty(n, np.complex128)
        args = (self.handles, len(self.ctxs), nw, qpx, gh, gw, ptr(gcf), shape[0], shape[1], pa, n, ptr(u), ptr(vv), ptr(wb), ptr(out))
        if mode == "vis":
            self._check(self.lib.skagrid_convdegrid2_mgpu_vis(*args))
            self._resident_shape = tuple(shape)
        elif mode == "tile":
            b = np.zeros(len(self.ctxs) + 1, np.int64)
            self._check(self.lib.skagrid_convdegrid2_mgpu_tile(*args, ptr(b)))
            self.bounds = b.tolist()
        else:
            raise ValueError("mode must be 'vis' or 'tile'")
        return out

    def conv_grid_resident(self, gcf, shape, p, wbin, v):
        """Visibility-sharded gridding from a zero grid that stays on the devices (no grid over PCIe); follow with
        convdegrid2(a=None) or gridding.grid_to_image(None, ctx=self.ctxs[0], n=shape[0])."""
        gcf = self._table(gcf)
        u, vv = _uv(p)
        vis = c128(v)
        wb = None if wbin is None else int64(wbin)
        nw, qpx, _, gh, gw = gcf.shape
        self._check(self.lib.skagrid_convgrid2_mgpu_vis(self.handles, len(self.ctxs), nw, qpx, gh, gw, ptr(gcf), shape[0], shape[1], None, u.size,
                                                        ptr(u), ptr(vv), ptr(wb), ptr(vis)))
        self._resident_shape = tuple(shape)
