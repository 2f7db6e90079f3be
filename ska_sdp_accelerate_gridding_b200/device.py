"""Device-resident API: torch tensors as device memory, torch's current CUDA stream as the launch stream,
libskagrid.so's `skagrid_dev_*` entry points for the arithmetic.  torch is plumbing only (allocation,
streams, NCCL); no torch op touches the gridding arithmetic."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .context import Context, get_context


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA tensor of {dtype}")


def context_for_current_device() -> Context:
    return get_context(torch.cuda.current_device())


def synth_vis(seed, first, count, n, support, nw, uniform=False, with_vis=True, ctx=None):
    """SURVEY.md 8d synthetic SKA1-Low-shaped visibilities [first, first+count) -> (u, v, wbin, vis) on the device."""
    ctx = ctx or context_for_current_device()
    dev = torch.device("cuda", ctx.device)
    u = torch.empty(count, dtype=torch.float64, device=dev)
    v = torch.empty(count, dtype=torch.float64, device=dev)
    wbin = torch.empty(count, dtype=torch.int64, device=dev)
    vis = torch.empty(count, dtype=torch.complex128, device=dev) if with_vis else None
    ctx.check(ctx.lib.skagrid_dev_synth_vis(ctx.h, seed, first, count, n, support, nw, int(uniform), _p(u), _p(v), _p(wbin), _p(vis), _stream()))
    return u, v, wbin, vis


def frac_coord(n, qpx, p, normalise=True, ctx=None):
    """Bit-exact frac_coord (src/Gridding.hs:126-140) of a CUDA float64 tensor -> (fl, frac) int64 CUDA tensors."""
    ctx = ctx or context_for_current_device()
    _chk(p, torch.float64, "p")
    fl = torch.empty(p.shape, dtype=torch.int64, device=p.device)
    fr = torch.empty(p.shape, dtype=torch.int64, device=p.device)
    ctx.check(ctx.lib.skagrid_dev_frac_coord(ctx.h, n, qpx, p.numel(), _p(p), _p(fl), _p(fr), int(normalise), _stream()))
    return fl, fr


def uvw_lambda_(freq, u, v, w, ctx=None):
    """In place (u,v,w) *= freq/299792458.0 (src/ImageDataset.hs:181-187)."""
    ctx = ctx or context_for_current_device()
    for t in (u, v, w):
        _chk(t, torch.float64, "uvw")
    ctx.check(ctx.lib.skagrid_dev_uvw_scale(ctx.h, u.numel(), _p(u), _p(v), _p(w), float(freq) / 299792458.0, 0, _stream()))


def div3_(lam, u, v, w, ctx=None):
    """In place (u,v,w) /= lam (src/Gridding.hs:838-839)."""
    ctx = ctx or context_for_current_device()
    for t in (u, v, w):
        _chk(t, torch.float64, "uvw")
    ctx.check(ctx.lib.skagrid_dev_uvw_scale(ctx.h, u.numel(), _p(u), _p(v), _p(w), float(lam), 1, _stream()))


def mirror_uvw_(u, v, w, vis=None, ctx=None):
    """In place mirror_uvw (src/Gridding.hs:551-562)."""
    ctx = ctx or context_for_current_device()
    for t in (u, v, w):
        _chk(t, torch.float64, "uvw")
    _chk(vis, torch.complex128, "vis")
    ctx.check(ctx.lib.skagrid_dev_mirror_uvw(ctx.h, u.numel(), _p(u), _p(v), _p(w), _p(vis), _stream()))


def find_closest(wbins, w, ctx=None):
    """findClosest (src/Gridding.hs:895-907) of a CUDA vector of w against the sorted CUDA vector wbins."""
    ctx = ctx or context_for_current_device()
    _chk(wbins, torch.float64, "wbins"); _chk(w, torch.float64, "w")
    out = torch.empty(w.shape, dtype=torch.int64, device=w.device)
    ctx.check(ctx.lib.skagrid_dev_find_closest(ctx.h, wbins.numel(), _p(wbins), w.numel(), _p(w), _p(out), _stream()))
    return out


def doweight_(theta, lam, u, v, vis, ctx=None):
    """In place doweight (src/Gridding.hs:564-583): vis /= number of visibilities sharing its cell."""
    ctx = ctx or context_for_current_device()
    _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(vis, torch.complex128, "vis")
    ctx.check(ctx.lib.skagrid_dev_doweight(ctx.h, float(theta), int(lam), u.numel(), _p(u), _p(v), _p(vis), _stream()))
    flags = C.c_int()
    ctx.check(ctx.lib.skagrid_dev_take_error(ctx.h, _stream(), C.byref(flags)))
    if flags.value & 2:
        raise _lib.SkagridError(-5, "doweight: a visibility falls outside the weight grid")


def slab_fft_rows_(n, row0, slab, ctx=None):
    """Stage 1 of the slab-distributed grid -> image (include/skagrid.h), in place on rows [row0, row0 + slab.shape[0])."""
    ctx = ctx or context_for_current_device()
    _chk(slab, torch.complex128, "slab")
    if slab.dim() != 2 or slab.shape[1] != n:
        raise ValueError("slab must be [rows, n]")
    ctx.check(ctx.lib.skagrid_dev_slab_fft_rows(ctx.h, int(n), int(row0), slab.shape[0], _p(slab), _stream()))


def slab_fft_cols_(n, col0, cols, want_image=True, ctx=None):
    """Stage 2, in place on the [n, ncols] column slab; returns (image [n, ncols] float64 or None, max as a 1-element tensor)."""
    ctx = ctx or context_for_current_device()
    _chk(cols, torch.complex128, "cols")
    if cols.dim() != 2 or cols.shape[0] != n:
        raise ValueError("cols must be [n, ncols]")
    img = torch.empty(cols.shape, dtype=torch.float64, device=cols.device) if want_image else None
    mx = torch.empty(1, dtype=torch.float64, device=cols.device)
    ctx.check(ctx.lib.skagrid_dev_slab_fft_cols(ctx.h, int(n), int(col0), cols.shape[1], _p(cols), _p(img), _p(mx), _stream()))
    return img, mx


def weight_count_(theta, lam, u, v, hist, ctx=None):
    """First phase of doweight for sharded visibilities: hist (n x n int32, n = round(theta*lam)) += cell counts of (u, v)."""
    ctx = ctx or context_for_current_device()
    _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(hist, torch.int32, "hist")
    ctx.check(ctx.lib.skagrid_dev_weight_count(ctx.h, float(theta), int(lam), u.numel(), _p(u), _p(v), _p(hist), _stream()))


def weight_apply_(theta, lam, u, v, hist, vis, ctx=None):
    """Second phase: vis /= hist[cell of (u, v)], in place.  Raises if a visibility fell outside the weight grid."""
    ctx = ctx or context_for_current_device()
    _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(hist, torch.int32, "hist"); _chk(vis, torch.complex128, "vis")
    ctx.check(ctx.lib.skagrid_dev_weight_apply(ctx.h, float(theta), int(lam), u.numel(), _p(u), _p(v), _p(hist), _p(vis), _stream()))
    flags = C.c_int()
    ctx.check(ctx.lib.skagrid_dev_take_error(ctx.h, _stream(), C.byref(flags)))
    if flags.value & 2:
        raise _lib.SkagridError(-5, "doweight: a visibility falls outside the weight grid")


def w_kernel_table(theta, ws, npixff, npixkern, qpx, conjugate=True, ctx=None):
    """w_kernel (src/Gridding.hs:610-728) for every w in `ws`, built in device memory -> [nw,qpx,qpx,s,s]."""
    ctx = ctx or context_for_current_device()
    ws = np.ascontiguousarray(ws, dtype=np.float64)
    out = torch.empty((ws.size, qpx, qpx, npixkern, npixkern), dtype=torch.complex128, device=torch.device("cuda", ctx.device))
    ctx.check(ctx.lib.skagrid_dev_w_kernels(ctx.h, float(theta), ws.size, ws.ctypes.data, npixff, npixkern, qpx, int(conjugate), _p(out), _stream()))
    return out


class Plan:
    """Bit-exact binning + uv-tile bucketing of one batch of visibilities (skagrid_dev_plan_*)."""

    def __init__(self, height, width, table_shape, u, v, wbin=None, vis=None, rows=None, slice_override=False, ctx=None):
        self.ctx = ctx or context_for_current_device()
        if len(table_shape) == 4:
            table_shape = (1,) + tuple(table_shape)
        nw, qpx, qpx2, gh, gw = table_shape
        if qpx != qpx2:
            raise ValueError("kernel table must be [nw,qpx,qpx,gh,gw]")
        row0, row1 = rows if rows is not None else (0, height)
        self.geom = _lib.Geom(height, width, row0, row1, nw, qpx, gh, gw)
        self.rows = (row0, row1)
        self.width = width
        _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(wbin, torch.int64, "wbin"); _chk(vis, torch.complex128, "vis")
        self.count = int(u.numel())
        self.capacity = max(self.count, 1)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.skagrid_dev_plan_create(self.ctx.h, C.byref(self.geom), self.count, _p(u), _p(v), _p(wbin), _p(vis),
                                                            int(slice_override), _stream(), C.byref(h)))
        self.h = h

    def update(self, u, v, wbin=None, vis=None):
        _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(wbin, torch.int64, "wbin"); _chk(vis, torch.complex128, "vis")
        self.ctx.check(self.ctx.lib.skagrid_dev_plan_update(self.ctx.h, self.h, int(u.numel()), _p(u), _p(v), _p(wbin), _p(vis), _stream()))
        self.count = int(u.numel())

    def stats(self):
        out = (C.c_int64 * 5)()
        self.ctx.check(self.ctx.lib.skagrid_dev_plan_stats(self.ctx.h, self.h, _stream(), C.byref(out)))
        return dict(kept=out[0], dropped=out[1], work_items=out[2], tiles=out[3], nonempty_tiles=out[4])

    def grid(self, table, grid, variant=0):
        """grid[row0:row1] += sum vis_k * table[slice_k].  variant 0 tiled, 1 atomic scatter (literal permute (+)), 2 tiled with a shallower tap pipeline."""
        _chk(table, torch.complex128, "table"); _chk(grid, torch.complex128, "grid")
        if grid.numel() != (self.rows[1] - self.rows[0]) * self.width:
            raise ValueError("grid tensor does not match the plan's owned rows")
        self.ctx.check(self.ctx.lib.skagrid_dev_grid(self.ctx.h, self.h, _p(table), _p(grid), int(variant), _stream()))

    def degrid(self, table, grid, out=None):
        _chk(table, torch.complex128, "table"); _chk(grid, torch.complex128, "grid")
        if out is None:
            out = torch.empty(self.count, dtype=torch.complex128, device=grid.device)
        _chk(out, torch.complex128, "out")
        self.ctx.check(self.ctx.lib.skagrid_dev_degrid(self.ctx.h, self.h, _p(table), _p(grid), _p(out), _stream()))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.skagrid_dev_plan_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def grid_to_image(grid, want_image=True, ctx=None):
    """In place on `grid` (n x n complex128 CUDA tensor): hermitian -> centred IFFT; returns (image or None, max tensor)."""
    ctx = ctx or context_for_current_device()
    _chk(grid, torch.complex128, "grid")
    n = grid.shape[0]
    img = torch.empty((n, n), dtype=torch.float64, device=grid.device) if want_image else None
    mx = torch.empty(1, dtype=torch.float64, device=grid.device)
    ctx.check(ctx.lib.skagrid_dev_grid_to_image(ctx.h, n, _p(grid), _p(img), _p(mx), _stream()))
    return img, mx
