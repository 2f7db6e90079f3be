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


def check_errors(ctx=None):
    """Reads and clears the context's device error word (synchronises the current stream) and raises on ANY pending bit:
    bit 0 = a w-plane / sub-cell / antenna index out of range in a plan (those visibilities were dropped), bit 1 = a
    visibility outside the weight grid.  The asynchronous entry points (Plan.update, weight_count_) only set the bits."""
    ctx = ctx or context_for_current_device()
    flags = C.c_int()
    ctx.check(ctx.lib.skagrid_dev_take_error(ctx.h, _stream(), C.byref(flags)))
    if flags.value & 1:
        raise _lib.SkagridError(-5, "plan: a w-plane / oversampling / antenna index is out of range (the visibility was dropped)")
    if flags.value & 2:
        raise _lib.SkagridError(-5, "doweight: a visibility falls outside the weight grid")
    if flags.value & 4:
        raise _lib.SkagridError(-2, "peer barrier: a rank did not arrive (timed out)")
    if flags.value:
        raise _lib.SkagridError(-5, f"device error word {flags.value:#x}")


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
    check_errors(ctx)


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


def grid_side(theta, lam, ctx=None):
    """N = P.round (theta * lam), half to even (src/Gridding.hs:571): the library's own definition."""
    return int(_lib.load().skagrid_grid_side(float(theta), int(lam)))


def _chk_hist(ctx, theta, lam, hist):
    n = grid_side(theta, lam)
    if tuple(hist.shape) != (n, n):
        raise ValueError(f"hist must be [{n}, {n}] = round(theta*lam) squared, got {tuple(hist.shape)}")


def weight_count_(theta, lam, u, v, hist, ctx=None):
    """First phase of doweight for sharded visibilities: hist (n x n int32, n = round(theta*lam)) += cell counts of (u, v)."""
    ctx = ctx or context_for_current_device()
    _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(hist, torch.int32, "hist")
    _chk_hist(ctx, theta, lam, hist)
    ctx.check(ctx.lib.skagrid_dev_weight_count(ctx.h, float(theta), int(lam), u.numel(), _p(u), _p(v), _p(hist), _stream()))


def weight_apply_(theta, lam, u, v, hist, vis, ctx=None):
    """Second phase: vis /= hist[cell of (u, v)], in place.  Raises if a visibility fell outside the weight grid."""
    ctx = ctx or context_for_current_device()
    _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(hist, torch.int32, "hist"); _chk(vis, torch.complex128, "vis")
    _chk_hist(ctx, theta, lam, hist)
    ctx.check(ctx.lib.skagrid_dev_weight_apply(ctx.h, float(theta), int(lam), u.numel(), _p(u), _p(v), _p(hist), _p(vis), _stream()))
    check_errors(ctx)


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

    @classmethod
    def empty(cls, height, width, table_shape, capacity, rows=None, slice_override=False, ctx=None):
        """A plan with room for `capacity` visibilities and no batch yet (skagrid_dev_plan_alloc)."""
        self = cls.__new__(cls)
        self.ctx = ctx or context_for_current_device()
        if len(table_shape) == 4:
            table_shape = (1,) + tuple(table_shape)
        nw, qpx, qpx2, gh, gw = table_shape
        if qpx != qpx2:
            raise ValueError("kernel table must be [nw,qpx,qpx,gh,gw]")
        row0, row1 = rows if rows is not None else (0, height)
        self.geom = _lib.Geom(height, width, row0, row1, nw, qpx, gh, gw)
        self.rows, self.width = (row0, row1), width
        self.count, self.capacity = 0, max(int(capacity), 1)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.skagrid_dev_plan_alloc(self.ctx.h, C.byref(self.geom), self.capacity, int(slice_override), C.byref(h)))
        self.h = h
        return self

    def update_packed(self, rec, check=True):
        """Re-bins records [count, W] of W doubles {u, v, wbin (int64 bits), [re, im]} (W = 5, or 3 for a degrid-only
        batch): the layout the uv-tile-sharded routing delivers (skagrid_dev_route_pack)."""
        _chk(rec, torch.float64, "rec")
        if rec.dim() != 2 or rec.shape[1] not in (3, 5):
            raise ValueError("rec must be [count, 3] or [count, 5] float64")
        if rec.shape[0] > self.capacity:
            raise ValueError("batch exceeds the plan's capacity")
        self.ctx.check(self.ctx.lib.skagrid_dev_plan_update_packed(self.ctx.h, self.h, int(rec.shape[0]), _p(rec), int(rec.shape[1]), _stream()))
        self.count = int(rec.shape[0])
        if check:
            self.check()

    def update(self, u, v, wbin=None, vis=None, check=True):
        """Re-bins a new batch into the existing buffers.  check=True (default) reads the device error word afterwards
        (one stream synchronisation) and raises SKAGRID_ERANGE for an out-of-range w-plane index, exactly as the
        constructor does; timed loops pass check=False and call `check()` once at the end."""
        _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(wbin, torch.int64, "wbin"); _chk(vis, torch.complex128, "vis")
        self.ctx.check(self.ctx.lib.skagrid_dev_plan_update(self.ctx.h, self.h, int(u.numel()), _p(u), _p(v), _p(wbin), _p(vis), _stream()))
        self.count = int(u.numel())
        if check:
            self.check()

    def check(self):
        check_errors(self.ctx)

    def set_vis(self, vis, in_plan_order=False):
        """New visibility values at the coordinates the plan was built from: no re-binning, no re-sorting -- what a major cycle
        over the same uvw needs (skagrid_dev_plan_set_vis).  vis: caller's order, one value per visibility of the batch; or,
        in_plan_order=True, one value per KEPT record in the plan's own order (vis[order()]): a sequential refresh."""
        _chk(vis, torch.complex128, "vis")
        if not in_plan_order and vis.numel() != self.count:
            raise ValueError("vis must have one value per visibility of the plan's batch")
        if in_plan_order and vis.numel() < self.stats()["kept"]:
            raise ValueError("vis must have one value per kept record")
        self.ctx.check(self.ctx.lib.skagrid_dev_plan_set_vis(self.ctx.h, self.h, _p(vis), int(bool(in_plan_order)), _stream()))

    def order(self):
        """index[r] = position in the caller's arrays of the visibility behind record r (int32 CUDA tensor, one entry per kept
        record): permute the data once with it and refresh / read back in plan order from then on."""
        if self.count >= 2 ** 31:
            raise ValueError("order(): batches of 2^31 visibilities or more do not fit the int32 index tensor")
        kept = self.stats()["kept"]
        idx = torch.empty(kept, dtype=torch.int32, device=torch.device("cuda", self.ctx.device))
        self.ctx.check(self.ctx.lib.skagrid_dev_plan_order(self.ctx.h, self.h, _p(idx), _stream()))
        return idx

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

    def degrid(self, table, grid, out=None, plan_order=False):
        """vis_out[k] for the plan's visibilities in the caller's order (others 0) -- or, plan_order=True, vis_out[r] for record r in
        the plan's own order (`order()`): sequential full-sector writes."""
        _chk(table, torch.complex128, "table"); _chk(grid, torch.complex128, "grid")
        if out is None:
            out = torch.empty(self.count, dtype=torch.complex128, device=grid.device)
        _chk(out, torch.complex128, "out")
        fn = self.ctx.lib.skagrid_dev_degrid_plan_order if plan_order else self.ctx.lib.skagrid_dev_degrid
        self.ctx.check(fn(self.ctx.h, self.h, _p(table), _p(grid), _p(out), _stream()))
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


class _DevArray:
    """A device buffer owned by the library, exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def resident_grid(ctx=None):
    """The grid the last host-pointer call left resident in the context, as a complex128 CUDA tensor VIEW (no copy): lets a
    one-process-per-GPU caller all-reduce the per-process grids between skagrid_conv_imaging2 and the NULL-grid calls."""
    ctx = ctx or context_for_current_device()
    p, h, w = C.c_void_p(), C.c_int64(), C.c_int64()
    ctx.check(ctx.lib.skagrid_resident_grid(ctx.h, C.byref(p), C.byref(h), C.byref(w)))
    real = torch.as_tensor(_DevArray(p.value, (h.value, w.value, 2), "<f8"), device=torch.device("cuda", ctx.device))
    return torch.view_as_complex(real)


def _bounds_arg(bounds):
    b = np.ascontiguousarray(bounds, dtype=np.int64)
    return b, C.c_void_p(b.ctypes.data)


def row_hist_(height, qpx, gh, v, hist, ctx=None):
    """hist[row of the footprint centre of every visibility] += 1 (int32 [height]): input of the slab balance."""
    ctx = ctx or context_for_current_device()
    _chk(v, torch.float64, "v"); _chk(hist, torch.int32, "hist")
    if hist.numel() != height:
        raise ValueError("hist must have one entry per grid row")
    ctx.check(ctx.lib.skagrid_dev_row_hist(ctx.h, int(height), int(qpx), int(gh), v.numel(), _p(v), _p(hist), _stream()))


def route_count(height, qpx, gh, bounds, v, ctx=None):
    """Records this device sends to every rank under the row-slab `bounds` (int32 CUDA tensor [nranks]; no host sync)."""
    ctx = ctx or context_for_current_device()
    _chk(v, torch.float64, "v")
    b, bp = _bounds_arg(bounds)
    counts = torch.empty(len(b) - 1, dtype=torch.int32, device=v.device)
    ctx.check(ctx.lib.skagrid_dev_route_count(ctx.h, int(height), int(qpx), int(gh), len(b) - 1, bp, v.numel(), _p(v), _p(counts), _stream()))
    return counts


def route_pack(height, qpx, gh, bounds, u, v, wbin, vis, send_counts, keep_index=False, ctx=None, out=None):
    """Destination-major send buffer [sum(send_counts), W] float64 (W = 5 with vis, 3 without) and, with keep_index, the
    source index of every record (int32).  send_counts: host list from route_count.  out: a float64 CUDA tensor with room
    for the records (e.g. a view of peer-visible memory) to pack into instead of a fresh tensor."""
    ctx = ctx or context_for_current_device()
    _chk(u, torch.float64, "u"); _chk(v, torch.float64, "v"); _chk(wbin, torch.int64, "wbin"); _chk(vis, torch.complex128, "vis")
    b, bp = _bounds_arg(bounds)
    seg = np.zeros(len(b) - 1, dtype=np.int64)
    seg[1:] = np.cumsum(np.asarray(send_counts, dtype=np.int64))[:-1]
    total = int(np.sum(send_counts))
    w = 3 if vis is None else 5
    if out is None:
        send = torch.empty((total, w), dtype=torch.float64, device=u.device)
    else:
        _chk(out, torch.float64, "out")
        if out.numel() < total * w:
            raise ValueError(f"send buffer too small: {out.numel()} doubles for {total} records of {w}")
        send = out.reshape(-1)[:total * w].view(total, w)
    sidx = torch.empty(total, dtype=torch.int32, device=u.device) if keep_index else None
    ctx.check(ctx.lib.skagrid_dev_route_pack(ctx.h, int(height), int(qpx), int(gh), len(b) - 1, bp, u.numel(), _p(u), _p(v), _p(wbin), _p(vis),
                                             C.c_void_p(seg.ctypes.data), _p(send), _p(sidx), _stream()))
    return send, sidx


def scatter_add_(out, sidx, back, ctx=None):
    """out[sidx[i]] += back[i] (complex128): the degridding partial sums returned by the slab owners."""
    ctx = ctx or context_for_current_device()
    _chk(out, torch.complex128, "out"); _chk(sidx, torch.int32, "sidx"); _chk(back, torch.complex128, "back")
    if sidx.numel() != back.numel():
        raise ValueError("one index per returned value")
    ctx.check(ctx.lib.skagrid_dev_scatter_add(ctx.h, sidx.numel(), _p(sidx), _p(back), _p(out), _stream()))
    return out


def grid_to_image(grid, want_image=True, ctx=None):
    """hermitian -> centred IFFT -> real of an n x n complex128 CUDA grid; returns (image or None, max tensor).  `grid` is
    left untouched for even n (complex-to-real transform of the hermitian half), transformed in place for odd n."""
    ctx = ctx or context_for_current_device()
    _chk(grid, torch.complex128, "grid")
    n = grid.shape[0]
    img = torch.empty((n, n), dtype=torch.float64, device=grid.device) if want_image else None
    mx = torch.empty(1, dtype=torch.float64, device=grid.device)
    ctx.check(ctx.lib.skagrid_dev_grid_to_image(ctx.h, n, _p(grid), _p(img), _p(mx), _stream()))
    return img, mx
