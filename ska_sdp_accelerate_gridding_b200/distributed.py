"""Multi-GPU partitioning of the gridding hot path: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch on the B200 box; gloo in the CPU tests of the host logic).

Two modes, both from BASELINE.json:
  * visibility-sharded (config 4): every rank grids a contiguous slice of the visibilities into a full local
    grid; the grids are summed with one NCCL (all-)reduce.  Gridding is linear in the visibilities
    (`permute (+)`, src/Gridding.hs:377), so the result differs from the 1-GPU grid only by summation order.
    Degridding replicates the grid and shards the visibilities: no collective afterwards.
  * uv-tile-sharded (config 5): rank g owns grid rows [bounds[g], bounds[g+1]); every visibility is routed
    (all-to-all) to each owner its footprint rows intersect (at most two when a slab is taller than the
    kernel) and each owner clips taps to its slab -- the same rule as fixoutofbounds (src/Gridding.hs:883-891).
    No grid reduction.

The routing/partition logic here is integer-only and backend-agnostic (it runs on CPU tensors under gloo in
tests/test_distributed.py); the gridding itself is done by `device.Plan` on CUDA tensors.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def shard_range(count: int, rank: int, world: int):
    """Contiguous, balanced slice [first, first+n) of `count` items for `rank` (first ranks get the remainder)."""
    base, rem = divmod(int(count), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def slab_bounds(height: int, world: int, align: int = 32):
    """Row slabs of the uv-tile-sharded mode: world+1 increasing bounds, interior ones aligned to the uv tile."""
    bounds = [0]
    for g in range(1, world):
        b = (height * g // world) // align * align
        bounds.append(max(b, bounds[-1]))
    bounds.append(height)
    return bounds


def balanced_slab_bounds(row_hist: torch.Tensor, world: int):
    """Row slabs holding (nearly) equal numbers of visibilities: bounds at the k/world quantiles of the per-row
    histogram of footprint-centre rows (already summed over ranks).  SKA1-Low uv coverage is core-dominated and
    mirrored to v >= 0, so equal-height slabs would leave half of the GPUs idle.  Integer only; every slab has at
    least one row."""
    height = int(row_hist.numel())
    cum = torch.cumsum(row_hist.to(torch.int64), 0)
    total = int(cum[-1].item())
    bounds = [0]
    for g in range(1, world):
        target = (total * g) // world
        b = int(torch.searchsorted(cum, torch.tensor([target], dtype=torch.int64, device=cum.device), right=True).item())
        b = min(max(b, bounds[-1] + 1), height - (world - g))
        bounds.append(b)
    bounds.append(height)
    return bounds


def owners_of_rows(y0: torch.Tensor, gh: int, bounds: Sequence[int]):
    """For footprints covering rows [y0, y0+gh): first and last owning rank (clamped to the grid).  Integer only.
    Returns (lo, hi, on_grid)."""
    b = torch.as_tensor(list(bounds[1:-1]), dtype=torch.int64, device=y0.device)
    height = bounds[-1]
    first = torch.clamp(y0, 0, height - 1)
    last = torch.clamp(y0 + gh - 1, 0, height - 1)
    on_grid = (y0 + gh > 0) & (y0 < height)
    lo = torch.bucketize(first, b, right=True)
    hi = torch.bucketize(last, b, right=True)
    return lo, hi, on_grid


def route_by_rows(y0: torch.Tensor, gh: int, bounds: Sequence[int], payload: Sequence[torch.Tensor], group=None, return_route=False):
    """All-to-all of per-visibility payload tensors (each [count, ...]) to the ranks owning their footprint rows.
    Returns the received payload tensors (concatenated over source ranks, in source-rank order) and the
    per-source receive counts; with return_route=True the second value is the full route
    {"order", "in_splits", "out_splits"} needed to send per-record results back (`return_to_source`)."""
    world = dist.get_world_size(group)
    lo, hi, on_grid = owners_of_rows(y0, gh, bounds)
    send_idx = []
    for g in range(world):
        sel = on_grid & (lo <= g) & (hi >= g)
        send_idx.append(torch.nonzero(sel, as_tuple=False).reshape(-1))
    send_counts = torch.tensor([int(i.numel()) for i in send_idx], dtype=torch.int64, device=y0.device)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    order = torch.cat(send_idx)
    in_splits = [int(x) for x in send_counts.tolist()]
    out_splits = [int(x) for x in recv_counts.tolist()]
    received = []
    for t in payload:
        src = t.index_select(0, order).contiguous()
        as_real = torch.view_as_real(src) if src.is_complex() else src
        out = torch.empty((sum(out_splits),) + tuple(as_real.shape[1:]), dtype=as_real.dtype, device=as_real.device)
        dist.all_to_all_single(out, as_real, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
        received.append(torch.view_as_complex(out) if src.is_complex() else out)
    if return_route:
        return received, {"order": order, "in_splits": in_splits, "out_splits": out_splits}
    return received, out_splits


def return_to_source(values: torch.Tensor, route, count: int, group=None):
    """Inverse of route_by_rows for per-record results: `values[i]` belongs to the i-th received record.  Sends them
    back to the ranks the records came from and sums the contributions of every source visibility (a visibility
    whose footprint straddles slabs was sent to several owners; each returns the partial sum over its own rows)."""
    as_real = torch.view_as_real(values) if values.is_complex() else values
    back = torch.empty((sum(route["in_splits"]),) + tuple(as_real.shape[1:]), dtype=as_real.dtype, device=as_real.device)
    dist.all_to_all_single(back, as_real.contiguous(), output_split_sizes=route["in_splits"], input_split_sizes=route["out_splits"], group=group)
    out = torch.zeros((count,) + tuple(as_real.shape[1:]), dtype=as_real.dtype, device=as_real.device)
    out.index_add_(0, route["order"], back)
    return torch.view_as_complex(out) if values.is_complex() else out


def allreduce_grid(grid: torch.Tensor, group=None, dst: int | None = None):
    """Sum the per-rank grids in place (complex128 viewed as 2 x float64).  dst=None: all-reduce, else reduce to dst."""
    real = torch.view_as_real(grid) if grid.is_complex() else grid
    if dst is None:
        dist.all_reduce(real, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(real, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return grid


def rows_to_columns(block: torch.Tensor, rows: Sequence[int], n: int, group=None):
    """Transpose of the ownership of an n x n array across ranks: in, this rank holds rows [rows[0], rows[1]) as
    [rows[1]-rows[0], n] (the row intervals of the ranks are disjoint and increasing with the rank; rows held by nobody
    are zero); out, rank h holds columns [n*h/P, n*(h+1)/P) of every row as [n, cols].  One all-to-all: the block
    (my rows) x (columns of rank h) goes to rank h, which places the blocks at their rows.  Backend-agnostic (complex
    tensors travel as pairs of reals).  Returns (columns, (c0, c1))."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    a, b = int(rows[0]), int(rows[1])
    if tuple(block.shape) != (b - a, n):
        raise ValueError("block must be [rows[1]-rows[0], n]")
    cb = [n * h // world for h in range(world + 1)]
    c0, c1 = cb[rank], cb[rank + 1]
    if world == 1:
        if (a, b) == (0, n):
            return block, (c0, c1)
        cols = torch.zeros((n, n), dtype=block.dtype, device=block.device)
        cols[a:b] = block
        return cols, (c0, c1)
    mine = torch.tensor([a, b], dtype=torch.int64, device=block.device)
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine, group=group)
    spans = [(int(t[0].item()), int(t[1].item())) for t in every]
    send = torch.cat([block[:, cb[h]:cb[h + 1]].reshape(-1) for h in range(world)])
    in_splits = [(b - a) * (cb[h + 1] - cb[h]) for h in range(world)]
    out_splits = [(hi - lo) * (c1 - c0) for lo, hi in spans]
    recv = torch.empty(sum(out_splits), dtype=block.dtype, device=block.device)
    x, y = (torch.view_as_real(recv), torch.view_as_real(send)) if block.is_complex() else (recv, send)
    dist.all_to_all_single(x, y, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    if sum(hi - lo for lo, hi in spans) == n:
        return recv.reshape(n, c1 - c0), (c0, c1)  # the intervals tile the rows: the stacked blocks are the column slab
    cols = torch.zeros((n, c1 - c0), dtype=block.dtype, device=block.device)
    off = 0
    for (lo, hi), cnt in zip(spans, out_splits):
        cols[lo:hi] = recv[off:off + cnt].reshape(hi - lo, c1 - c0)
        off += cnt
    return cols, (c0, c1)


def slab_grid_to_image(slab: torch.Tensor, bounds: Sequence[int], group=None, want_image=True, nonzero=None):
    """Grid -> image (make_grid_hermitian, centred ifft, real, maximum: src/ImageDataset.hs:74-77) for an n x n grid held as
    row slabs, rank g owning rows [bounds[g], bounds[g+1]) -- the layout uv-tile-sharded gridding produces -- WITHOUT
    gathering the grid: row transforms on the owners, one all-to-all transpose, column transforms on the column owners.
    `slab` ([rows, n] complex128) is transformed in place.  nonzero = (lo, hi): only rows [lo, hi) of this rank's slab
    can be non-zero (e.g. TileShardedGridder.nonzero_rows(): mirrored uv coverage leaves half of the grid empty), the
    others are neither transformed nor sent.  Returns (image columns [n, c1-c0] float64 or None, (c0, c1), maximum over
    the whole image as a float)."""
    from . import device as dv
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(bounds[-1])
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    if tuple(slab.shape) != (r1 - r0, n):
        raise ValueError("slab must be [bounds[rank+1]-bounds[rank], n]")
    lo, hi = (r0, r1) if nonzero is None else (max(r0, int(nonzero[0])), min(r1, int(nonzero[1])))
    hi = max(hi, lo)
    block = slab[lo - r0:hi - r0]
    dv.slab_fft_rows_(n, lo, block)
    cols, (c0, c1) = rows_to_columns(block, (lo, hi), n, group)
    img, mx = dv.slab_fft_cols_(n, c0, cols, want_image=want_image)
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return img, (c0, c1), float(mx.item())


def doweight_sharded_(theta, lam, u, v, vis, group=None):
    """doweight (src/Gridding.hs:564-583) for visibilities sharded over ranks, in place on this rank's `vis`: the weight
    of a visibility is the number of visibilities of ALL ranks in its cell, so the per-rank cell counts are summed with
    one all-reduce of the n x n count grid (n = round(theta*lam)) between counting and dividing."""
    from . import device as dv
    n = int(round(theta * lam))
    hist = torch.zeros((n, n), dtype=torch.int32, device=u.device)
    dv.weight_count_(theta, lam, u, v, hist)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    dv.weight_apply_(theta, lam, u, v, hist, vis)
    return vis


class VisShardedGridder:
    """Visibility-sharded gridding / degridding on CUDA tensors (config 4)."""

    def __init__(self, height, width, table, group=None):
        self.h, self.w, self.table, self.group = height, width, table, group
        self.plan = None

    def _plan(self, u, v, wbin, vis):
        from . import device as dv
        if self.plan is None or self.plan.count < u.numel():
            self.plan = dv.Plan(self.h, self.w, self.table.shape, u, v, wbin, vis)
        else:
            self.plan.update(u, v, wbin, vis)
        return self.plan

    def grid(self, u, v, wbin, vis, out=None, dst=None):
        """Grids this rank's visibilities and sums over ranks.  Returns the full grid (on every rank, or on dst)."""
        if out is None:
            out = torch.zeros((self.h, self.w), dtype=torch.complex128, device=u.device)
        self._plan(u, v, wbin, vis).grid(self.table, out)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            allreduce_grid(out, self.group, dst)
        return out

    def degrid(self, grid, u, v, wbin):
        """Every rank holds the full (model) grid; each degrids its own visibilities."""
        return self._plan(u, v, wbin, None).degrid(self.table, grid)


class TileShardedGridder:
    """uv-tile-sharded gridding on CUDA tensors (config 5): this rank owns rows [bounds[rank], bounds[rank+1])."""

    def __init__(self, height, width, table, group=None):
        self.h, self.w, self.table, self.group = height, width, table, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.set_bounds(slab_bounds(height, self.world))

    def set_bounds(self, bounds):
        self.bounds = [int(b) for b in bounds]
        self.rows = (self.bounds[self.rank], self.bounds[self.rank + 1])

    def balance(self, v):
        """Choose the slab bounds from this data set's uv coverage (collective): equal visibility counts per rank."""
        from . import device as dv
        y, _ = dv.frac_coord(self.h, self.table.shape[-3], v)
        hist = torch.bincount(torch.clamp(y, 0, self.h - 1), minlength=self.h)
        if self.world > 1:
            dist.all_reduce(hist, group=self.group)
        self.set_bounds(balanced_slab_bounds(hist, self.world))
        return self.bounds

    def route(self, u, v, wbin, vis=None, return_route=False):
        from . import device as dv
        gh = self.table.shape[-2]
        qpx = self.table.shape[-3]
        payload = (u, v, wbin) if vis is None else (u, v, wbin, vis)
        if self.world == 1:
            return payload, (None if return_route else [u.numel()])
        y, _ = dv.frac_coord(self.h, qpx, v)
        return route_by_rows(y - gh // 2, gh, self.bounds, payload, self.group, return_route=return_route)

    def grid(self, u, v, wbin, vis, out=None):
        """Routes, then grids into this rank's slab [rows, width] (no reduction)."""
        from . import device as dv
        (ru, rv, rwb, rvis), _ = self.route(u, v, wbin, vis)
        self.last_routed = int(ru.numel())
        self._last_rv = rv
        if out is None:
            out = torch.zeros((self.rows[1] - self.rows[0], self.w), dtype=torch.complex128, device=u.device)
        if ru.numel() > 0:
            plan = getattr(self, "_plan", None)
            if plan is None or plan.capacity < ru.numel() or plan.rows != self.rows:
                if plan is not None:
                    plan.close()
                self._plan = plan = dv.Plan(self.h, self.w, self.table.shape, ru, rv, rwb, rvis, rows=self.rows)
            else:
                plan.update(ru, rv, rwb, rvis)
            plan.grid(self.table, out)
        return out

    def nonzero_rows(self):
        """Rows of this rank's slab the last `grid` call (into a zeroed slab) can have touched: [lo, hi) in grid rows."""
        from . import device as dv
        r0, r1 = self.rows
        rv = getattr(self, "_last_rv", None)
        if rv is None or rv.numel() == 0:
            return (r0, r0)
        gh = self.table.shape[-2]
        y, _ = dv.frac_coord(self.h, self.table.shape[-3], rv)
        lo, hi = int(y.min().item()) - gh // 2, int(y.max().item()) - gh // 2 + gh
        return (min(max(lo, r0), r1), min(max(hi, r0), r1))

    def degrid(self, slab, u, v, wbin):
        """Adjoint of `grid`: `slab` holds this rank's rows of the (model) grid.  Coordinates are routed to the owners,
        every owner degrids the taps that fall on its rows, the partial sums travel back and are added per visibility."""
        from . import device as dv
        (ru, rv, rwb), route = self.route(u, v, wbin, return_route=True)
        partial = torch.zeros(ru.numel(), dtype=torch.complex128, device=u.device)
        if ru.numel() > 0:
            plan = dv.Plan(self.h, self.w, self.table.shape, ru, rv, rwb, None, rows=self.rows)
            plan.degrid(self.table, slab, partial)
            plan.close()
        if route is None:
            return partial
        return return_to_source(partial, route, u.numel(), self.group)
