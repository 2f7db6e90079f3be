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


TRACE = None   # a list: the exchange functions below append (label, CUDA event) at their sub-stage boundaries (bench.py)


def _mark(label):
    if TRACE is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        TRACE.append((label, ev))


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


def weighted_slab_bounds(row_cost: torch.Tensor, world: int):
    """Row slabs of (nearly) equal total cost: bounds at the k/world quantiles of the cumulative per-row cost (any
    non-negative float tensor, e.g. visibility counts scaled by a measured cost per visibility).  Every slab keeps at least
    one row."""
    height = int(row_cost.numel())
    cum = torch.cumsum(row_cost.to(torch.float64).cpu(), 0)
    total = float(cum[-1].item())
    bounds = [0]
    for g in range(1, world):
        b = int(torch.searchsorted(cum, torch.tensor([total * g / world], dtype=torch.float64), right=True).item())
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


def rows_to_columns(block: torch.Tensor, rows: Sequence[int], n: int, group=None, spans=None):
    """Transpose of the ownership of an n x n array across ranks: in, this rank holds rows [rows[0], rows[1]) as
    [rows[1]-rows[0], n] (the row intervals of the ranks are disjoint and increasing with the rank; rows held by nobody
    are zero); out, rank h holds columns [n*h/P, n*(h+1)/P) of every row as [n, cols].  One all-to-all: the block
    (my rows) x (columns of rank h) goes to rank h, which places the blocks at their rows.  Backend-agnostic (complex
    tensors travel as pairs of reals).  spans: the (a, b) of every rank when the caller knows them.  Returns (columns, (c0, c1))."""
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
    if spans is None:  # every rank's (a, b); callers that know them (regular slabs) pass them and save the exchange + host sync
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


def slab_grid_to_image(slab: torch.Tensor, bounds: Sequence[int], group=None, want_image=True, nonzero=None, spans=None, sync_max=True):
    """Grid -> image (make_grid_hermitian, centred ifft, real, maximum: src/ImageDataset.hs:74-77) for an n x n grid held as
    row slabs, rank g owning rows [bounds[g], bounds[g+1]) -- the layout uv-tile-sharded gridding produces -- WITHOUT
    gathering the grid: row transforms on the owners, one all-to-all transpose, column transforms on the column owners.
    `slab` ([rows, n] complex128) is transformed in place.  nonzero = (lo, hi): only rows [lo, hi) of this rank's slab
    can be non-zero (e.g. TileShardedGridder.nonzero_rows(): mirrored uv coverage leaves half of the grid empty), the
    others are neither transformed nor sent.  spans: instead of `bounds` + `nonzero`, the rows [a, b) every rank holds
    (list of world pairs; rows held by nobody are zero; `slab` is then [b-a, n] of this rank's pair and bounds = n) -- the
    layout the visibility-sharded reduce-scatter produces.  Returns (image columns [n, c1-c0] float64 or None, (c0, c1), maximum over
    the whole image as a float -- or, with sync_max=False, as a 1-element CUDA tensor without synchronising the host)."""
    from . import device as dv
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if spans is not None:
        n = int(bounds)
        lo, hi = (int(x) for x in spans[rank])
        if tuple(slab.shape) != (hi - lo, n):
            raise ValueError("slab must be [spans[rank][1]-spans[rank][0], n]")
        block = slab
    else:
        n = int(bounds[-1])
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        if tuple(slab.shape) != (r1 - r0, n):
            raise ValueError("slab must be [bounds[rank+1]-bounds[rank], n]")
        lo, hi = (r0, r1) if nonzero is None else (max(r0, int(nonzero[0])), min(r1, int(nonzero[1])))
        hi = max(hi, lo)
        block = slab[lo - r0:hi - r0]
    dv.slab_fft_rows_(n, lo, block)
    cols, (c0, c1) = rows_to_columns(block, (lo, hi), n, group, spans=spans)
    img, mx = dv.slab_fft_cols_(n, c0, cols, want_image=want_image)
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return img, (c0, c1), (float(mx.item()) if sync_max else mx)   # sync_max=False: the 1-element device tensor, no host sync


def doweight_sharded_(theta, lam, u, v, vis, group=None):
    """doweight (src/Gridding.hs:564-583) for visibilities sharded over ranks, in place on this rank's `vis`: the weight
    of a visibility is the number of visibilities of ALL ranks in its cell, so the per-rank cell counts are summed with
    one all-reduce of the n x n count grid (n = round(theta*lam)) between counting and dividing."""
    from . import device as dv
    n = dv.grid_side(theta, lam)
    hist = torch.zeros((n, n), dtype=torch.int32, device=u.device)
    dv.weight_count_(theta, lam, u, v, hist)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    dv.weight_apply_(theta, lam, u, v, hist, vis)
    return vis


def route_cuda(height, qpx, gh, bounds, u, v, wbin, vis=None, keep_index=False, group=None):
    """The routing of the uv-tile-sharded mode on CUDA tensors, hand-written kernels + ONE packed all-to-all
    (skagrid_dev_route_count / _route_pack, csrc/route.cu): per-destination counts -> exchange of the counts -> records of
    W doubles {u, v, wbin, [re, im]} packed destination-major -> all-to-all.  One host synchronisation (NCCL needs the split
    sizes on the host).  Returns (received records [n, W] float64, route); route = {"sidx", "in_splits", "out_splits"}."""
    from . import device as dv
    world = dist.get_world_size(group)
    counts = dv.route_count(height, qpx, gh, bounds, v)
    both = torch.empty(2 * world, dtype=torch.int32, device=v.device)
    both[:world] = counts
    dist.all_to_all_single(both[world:], counts, group=group)
    host = both.tolist()
    in_splits, out_splits = host[:world], host[world:]
    send, sidx = dv.route_pack(height, qpx, gh, bounds, u, v, wbin, vis, in_splits, keep_index=keep_index)
    recv = torch.empty((sum(out_splits), send.shape[1]), dtype=torch.float64, device=v.device)
    dist.all_to_all_single(recv, send, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    return recv, {"sidx": sidx, "in_splits": in_splits, "out_splits": out_splits}


def return_cuda(partial: torch.Tensor, route, count: int, group=None):
    """Inverse of route_cuda for per-record complex results: all-to-all back, then out[sidx[i]] += back[i] (hand-written
    scatter-add; a footprint straddling slabs has one partial sum per owner)."""
    from . import device as dv
    back = torch.empty(sum(route["in_splits"]), dtype=torch.complex128, device=partial.device)
    dist.all_to_all_single(torch.view_as_real(back), torch.view_as_real(partial), output_split_sizes=route["in_splits"],
                           input_split_sizes=route["out_splits"], group=group)
    out = torch.zeros(count, dtype=torch.complex128, device=partial.device)
    return dv.scatter_add_(out, route["sidx"], back)


def peer_slab_grid_to_image(pg, pbuf, row_base, spans, n, group=None, want_image=True, sync_max=True):
    """slab_grid_to_image with the transpose pulled over NVLink peer memory instead of an all-to-all: rank s holds grid rows
    [row_base[s], ...) as [rows, n] complex128 at the start of its peer-visible buffer `pbuf`; spans[s] = (a, b) are the rows
    of rank s that can be non-zero (host list, all ranks).  Row transforms in place, device barrier, then every rank pulls its
    column block of every peer's rows straight into place (no pack, no unpack) -- one SM kernel reading all peers at once, or
    one strided copy-engine copy per peer between two or three ranks -- and does the column transforms.
    Returns (image columns or None, (c0, c1), max)."""
    from . import device as dv
    P, me = pg.world, pg.rank
    a, b = spans[me]
    b = max(b, a)
    _mark("image: start")
    if b > a:
        rows = pbuf.tensor(torch.complex128, (b - a, n), offset_bytes=(a - row_base[me]) * n * 16)
        dv.slab_fft_rows_(n, a, rows)
    _mark("image: row transforms")
    cb = [n * h // P for h in range(P + 1)]
    c0, c1 = cb[me], cb[me + 1]
    cw = c1 - c0
    pg.barrier()                                                  # all row transforms are done
    dev = torch.device("cuda", pg.ctx.device)
    cols = torch.zeros((n, cw), dtype=torch.complex128, device=dev)
    copies = []
    for s_ in range(P):
        sa, sb = spans[s_]
        if sb <= sa:
            continue
        src = pbuf.ptrs[s_] + ((sa - row_base[s_]) * n + c0) * 16
        copies.append((cols.data_ptr() + sa * cw * 16, cw * 16, src, n * 16, cw * 16, sb - sa))
    _mark("image: barrier + zeroed columns")
    if P >= 4 or cw * 16 <= (32 << 10):
        pg.gather2d(copies)     # one SM kernel reading all peers at once: from four ranks on it outruns the copy engines (0.13 vs
                                # 0.27 ms for 8 x 8 MB on 8 GPUs), and a strided copy-engine copy pays per row when the rows are short
    else:
        pg.pull(copies)
    _mark("image: transpose pulled")
    img, mx = dv.slab_fft_cols_(n, c0, cols, want_image=want_image)
    if P > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    _mark("image: column transforms + max")
    return img, (c0, c1), (float(mx.item()) if sync_max else mx)


class VisShardedGridder:
    """Visibility-sharded gridding / degridding on CUDA tensors (config 4).

    Two forms of the reduction:
      * `grid`: full local grid + one all-reduce (every rank ends with the sum);
      * `grid_slabs` / `gather_slabs`: reduce-scatter into equal row slabs of the ACTIVE rows -- the rows any footprint of the
        data set can touch, fixed by `set_active_rows` (mirrored uv coverage, v >= 0, leaves the lower half of the grid
        empty) -- which feeds the slab-distributed grid -> image (`slab_grid_to_image(spans=...)`); the all-gather of the
        reduced slabs (for degridding, every rank needs the grid) can then overlap the image stage.  Half the bytes of the
        all-reduce on each leg, and no replicated FFT."""

    def __init__(self, height, width, table, group=None, check=True):
        self.h, self.w, self.table, self.group = height, width, table, group
        self.plan = None
        self.check = check
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.active = None

    def _plan(self, u, v, wbin, vis):
        from . import device as dv
        if self.plan is None or self.plan.capacity < u.numel():
            if self.plan is not None:
                self.plan.close()
            self.plan = dv.Plan(self.h, self.w, self.table.shape, u, v, wbin, vis)
        else:
            self.plan.update(u, v, wbin, vis, check=self.check)
        return self.plan

    def grid(self, u, v, wbin, vis, out=None, dst=None):
        """Grids this rank's visibilities and sums over ranks.  Returns the full grid (on every rank, or on dst)."""
        if out is None:
            out = torch.zeros((self.h, self.w), dtype=torch.complex128, device=u.device)
        self._plan(u, v, wbin, vis).grid(self.table, out)
        if dist.is_initialized() and self.world > 1:
            allreduce_grid(out, self.group, dst)
        return out

    def set_active_rows(self, v):
        """Collective, once per data set (the uv coverage of an observation is known up front): the grid rows any footprint
        can touch, widened to a multiple of the world size.  Returns (lo, rows per rank)."""
        from . import device as dv
        gh, qpx = self.table.shape[-2], self.table.shape[-3]
        if v.numel() > 0:
            y, _ = dv.frac_coord(self.h, qpx, v)
            ext = torch.stack([y.min() - gh // 2, -(y.max() - gh // 2 + gh)])
        else:
            ext = torch.tensor([self.h, 0], dtype=torch.int64, device=v.device)
        if dist.is_initialized() and self.world > 1:
            dist.all_reduce(ext, op=dist.ReduceOp.MIN, group=self.group)
        lo, hi = int(ext[0].item()), -int(ext[1].item())
        self.active = active_row_slabs(self.h, lo, hi, self.world)
        return self.active

    def spans(self):
        lo, m = self.active
        return [(lo + r * m, lo + (r + 1) * m) for r in range(self.world)]

    def grid_slabs(self, u, v, wbin, vis, work, slab):
        """Grids into `work` (full [h, w] grid, zeroed here where needed) and reduce-scatters the active rows: on return
        `slab` ([rows per rank, w]) holds this rank's rows spans()[rank] of the SUM over ranks."""
        lo, m = self.active
        work[lo:lo + m * self.world].zero_()
        self._plan(u, v, wbin, vis).grid(self.table, work)
        act = work[lo:lo + m * self.world]
        if self.world > 1:
            dist.reduce_scatter_tensor(torch.view_as_real(slab), torch.view_as_real(act), group=self.group)
        else:
            slab.copy_(act)
        return slab

    def gather_slabs(self, slab, work, async_op=False):
        """All-gather of the reduced slabs into the active rows of `work`; the other rows of a gridded sum are zero, and are
        zeroed here.  Returns the work handle when async_op (wait() before reading `work`)."""
        lo, m = self.active
        work[:lo].zero_()
        work[lo + m * self.world:].zero_()
        act = work[lo:lo + m * self.world]
        if self.world > 1:
            return dist.all_gather_into_tensor(torch.view_as_real(act), torch.view_as_real(slab), group=self.group, async_op=async_op)
        act.copy_(slab)
        return None

    # ---- the same over peer memory (csrc/ipc.cu): no collective library in the data path
    def enable_peer(self, pg):
        """Collective.  The local grid moves into a peer-visible buffer (`self.work`, [h, w] complex128)."""
        from .peer import PeerBuffer
        self.pg = pg
        self.pgrid = PeerBuffer(pg, self.h * self.w * 16)
        self.work = self.pgrid.tensor(torch.complex128, (self.h, self.w))
        return self.work

    def _slab_off(self, r):
        lo, m = self.active
        return (lo + r * m) * self.w * 16

    def grid_slabs_peer(self, u=None, v=None, wbin=None, vis=None, variant=0, broadcast=False):
        """Reduce-scatter over NVLink peer memory: grid into the local peer-visible grid, barrier, then ONE kernel on every
        rank sums its row slab of all peers' grids in place (all peers in flight), barrier.  Returns this rank's reduced
        slab (a view of rows spans()[rank] of `self.work`).  u = None: the plan was already updated by the caller.
        broadcast: the summing kernel also stores its slab into every peer's grid -- afterwards `self.work` holds the whole
        reduced grid on every rank and no gather is needed."""
        lo, m = self.active
        plan = self.plan if u is None else self._plan(u, v, wbin, vis)
        self.pg.barrier()                      # every peer has finished pulling the previous pass's slabs from this grid
        self.work[lo:lo + m * self.world].zero_()
        plan.grid(self.table, self.work, variant=variant)
        self.pg.barrier()                      # all local grids are complete
        self.pg.peer_sum_(self.pgrid, self._slab_off(self.rank), m * self.w, broadcast=broadcast)
        self.pg.barrier()                      # all slabs are reduced (and, with broadcast, delivered)
        return self.work[lo + self.rank * m:lo + (self.rank + 1) * m]

    def gather_slabs_peer(self, join=True, sm=None):
        """All-gather of the reduced slabs: every rank pulls the other ranks' slabs out of their grids into the same rows of its
        own -- by the copy engines (one copy per peer) or, sm=True, by ONE SM kernel that reads all peers at once on a side
        stream (default from four ranks on: 637 against 270 GB/s per rank on 8 GPUs, the other way round on 2;
        profiles/r02_peer_primitives_*.json).  join=False returns a handle; wait() before reading `self.work`."""
        lo, m = self.active
        nbytes = m * self.w * 16
        copies = [(self.pgrid.local + self._slab_off(p), self.pgrid.ptrs[p] + self._slab_off(p), nbytes) for p in range(self.world) if p != self.rank]
        if sm is None:
            sm = self.world >= 4
        if not sm:
            return self.pg.pull(copies, join=join, pool=1)
        # one block per SM: 148 x 256 threads x 7 peers x 16 bytes in flight cover the NVLink bandwidth-delay product, and the small
        # kernels of the image stage run beside the exchange instead of behind it
        h = self.pg.gather_async(copies, max_blocks=0 if join else self.pg.ctx_sm_count())
        if join:
            h.wait()
            return None
        return h

    def degrid(self, grid, u=None, v=None, wbin=None, out=None):
        """Every rank holds the full (model) grid; each degrids its own visibilities.  u = None: at the coordinates of the
        last grid / degrid call (the plan is reused as it is)."""
        plan = self.plan if u is None else self._plan(u, v, wbin, None)
        return plan.degrid(self.table, grid, out)


def active_row_slabs(height: int, lo: int, hi: int, world: int):
    """Rows [lo, hi) clamped to the grid and widened to `world` equal slabs inside it: (first row, rows per rank)."""
    lo, hi = max(0, min(lo, height)), max(0, min(hi, height))
    if hi <= lo:
        lo, hi = 0, min(height, world)
    m = -(-(hi - lo) // world)
    if m * world > height:
        raise ValueError("grid has fewer rows than ranks")
    lo = min(lo, height - m * world)
    return lo, m


class TileShardedGridder:
    """uv-tile-sharded gridding on CUDA tensors (config 5): this rank owns rows [bounds[rank], bounds[rank+1])."""

    def __init__(self, height, width, table, group=None, check=True):
        self.h, self.w, self.table, self.group = height, width, table, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.check = check
        self._plan = None
        self._route = None
        self._count = 0
        self.set_bounds(slab_bounds(height, self.world))

    def set_bounds(self, bounds):
        self.bounds = [int(b) for b in bounds]
        self.rows = (self.bounds[self.rank], self.bounds[self.rank + 1])

    def balance(self, v):
        """Choose the slab bounds from this data set's uv coverage (collective): equal visibility counts per rank."""
        from . import device as dv
        hist = torch.zeros(self.h, dtype=torch.int32, device=v.device)
        dv.row_hist_(self.h, self.table.shape[-3], self.table.shape[-2], v, hist)
        if self.world > 1:
            dist.all_reduce(hist, group=self.group)
        self._row_hist = hist
        self.set_bounds(balanced_slab_bounds(hist, self.world))
        return self.bounds

    def rebalance(self, seconds):
        """Collective, after `balance` and a measured pass: `seconds` is what this rank's work on its current slab took
        (binning, gridding, the row transforms of the image stage, degridding).  Equal visibility counts do not mean equal
        time: a slab costs  alpha * visibilities + beta * non-empty rows  (the dense core rows are cheap per visibility, the
        sparse outer ones pay per row: tiles to zero and flush, rows to transform).  alpha and beta are fitted (least squares)
        to every measurement so far -- one equation per rank and round -- and the bounds move to the equal-cost quantiles of
        alpha * hist[row] + beta * [hist[row] > 0].  Falls back to scaling each slab's rows by its own measured cost per
        visibility when the fit is degenerate.  Returns the new bounds."""
        import numpy as np
        t = torch.tensor([float(seconds)], dtype=torch.float64, device=self._row_hist.device)
        every = torch.empty(self.world, dtype=torch.float64, device=t.device)
        if self.world > 1:
            dist.all_gather_into_tensor(every, t, group=self.group)
        else:
            every.copy_(t)
        hist = self._row_hist.to(torch.float64).cpu()
        nonempty = (hist > 0).to(torch.float64)
        if not hasattr(self, "_fit_rows"):
            self._fit_rows = []
        times = every.tolist()
        for g, tg in enumerate(times):
            a, b = self.bounds[g], self.bounds[g + 1]
            self._fit_rows.append((float(hist[a:b].sum()), float(nonempty[a:b].sum()), tg))
        A = np.array([[n, r] for n, r, _ in self._fit_rows])
        y = np.array([tg for _, _, tg in self._fit_rows])
        cost = None
        self.last_times = times
        if len(self._fit_rows) >= 2:
            (alpha, beta), *_ = np.linalg.lstsq(A, y, rcond=None)
            resid = np.abs(A @ np.array([alpha, beta]) - y) / np.maximum(y, 1e-12)
            if alpha > 0 and beta >= 0 and float(resid.max()) < 0.03:   # the two-term model explains every measurement: use it
                cost = alpha * hist + beta * nonempty
        if cost is None:
            cost = torch.zeros_like(hist)
            for g, tg in enumerate(times):
                a, b = self.bounds[g], self.bounds[g + 1]
                n = float(hist[a:b].sum())
                if n > 0:
                    cost[a:b] = hist[a:b] * (tg / n)
        self.set_bounds(weighted_slab_bounds(cost, self.world))
        return self.bounds

    def route(self, u, v, wbin, vis=None, keep_index=False):
        """Records [n, W] (W = 5 with vis, else 3) of every visibility whose footprint rows intersect this rank's slab, and
        the route for sending per-record results back."""
        gh, qpx = self.table.shape[-2], self.table.shape[-3]
        if self.world == 1:
            cols = [u, v, wbin.view(torch.float64)] + ([vis.real, vis.imag] if vis is not None else [])
            return torch.stack(cols, dim=1).contiguous(), None
        return route_cuda(self.h, qpx, gh, self.bounds, u, v, wbin, vis, keep_index, self.group)

    def _fill(self, rec):
        from . import device as dv
        plan = self._plan
        if plan is None or plan.capacity < rec.shape[0] or plan.rows != self.rows:
            if plan is not None:
                plan.close()
            self._plan = plan = dv.Plan.empty(self.h, self.w, self.table.shape, int(rec.shape[0] * 1.02) + 1024, rows=self.rows)
        plan.update_packed(rec, check=self.check)
        return plan

    def grid(self, u, v, wbin, vis, out=None, keep_route=False):
        """Routes, then grids into this rank's slab [rows, width] (no reduction).  keep_route: remember where the records came
        from so that `degrid_routed` can send results back without routing again."""
        rec, route = self.route(u, v, wbin, vis, keep_index=keep_route)
        self.last_routed = int(rec.shape[0])
        self._last_rec, self._route, self._count = rec, (route if keep_route else None), int(u.numel())
        if out is None:
            out = torch.zeros((self.rows[1] - self.rows[0], self.w), dtype=torch.complex128, device=u.device)
        if rec.shape[0] > 0:
            self._fill(rec).grid(self.table, out)
        return out

    def nonzero_rows(self):
        """Rows of this rank's slab the last `grid` call (into a zeroed slab) can have touched: [lo, hi) in grid rows."""
        from . import device as dv
        r0, r1 = self.rows
        rec = getattr(self, "_last_rec", None)
        if rec is None or rec.shape[0] == 0:
            return (r0, r0)
        gh = self.table.shape[-2]
        y, _ = dv.frac_coord(self.h, self.table.shape[-3], rec[:, 1].contiguous())
        lo, hi = int(y.min().item()) - gh // 2, int(y.max().item()) - gh // 2 + gh
        return (min(max(lo, r0), r1), min(max(hi, r0), r1))

    def degrid_routed(self, slab):
        """Adjoint of the last `grid(..., keep_route=True)` at the same coordinates: the plan of the routed records is reused,
        every owner degrids the taps on its rows of `slab`, the partial sums travel back (one all-to-all) and are added per
        source visibility."""
        if self._plan is None or (self.world > 1 and self._route is None):
            raise RuntimeError("degrid_routed needs a previous grid(..., keep_route=True)")
        partial = torch.zeros(self.last_routed, dtype=torch.complex128, device=slab.device)
        if self.last_routed > 0:
            self._plan.degrid(self.table, slab, partial)
        if self.world == 1:
            return partial
        return return_cuda(partial, self._route, self._count, self.group)

    def degrid(self, slab, u, v, wbin):
        """Adjoint of `grid`: `slab` holds this rank's rows of the (model) grid.  Coordinates are routed to the owners,
        every owner degrids the taps that fall on its rows, the partial sums travel back and are added per visibility."""
        rec, route = self.route(u, v, wbin, None, keep_index=True)
        partial = torch.zeros(rec.shape[0], dtype=torch.complex128, device=u.device)
        if rec.shape[0] > 0:
            self._fill(rec).degrid(self.table, slab, partial)
        if route is None:
            return partial
        return return_cuda(partial, route, u.numel(), self.group)

    # ---- the same over peer memory (csrc/ipc.cu): NCCL only carries the P x P table of record counts
    def enable_peer(self, pg, send_capacity, recv_capacity=None):
        """Collective, after the bounds are final (`balance`).  send_capacity / recv_capacity: most records this rank packs /
        receives per batch.  Allocates the peer-visible send buffer, partial-sum buffer and slab; returns the slab
        ([rows, w] complex128 view)."""
        from .peer import PeerBuffer
        self.pg = pg
        self.send_cap = int(send_capacity)
        self.recv_cap = int(recv_capacity if recv_capacity is not None else send_capacity)
        self.psend = PeerBuffer(pg, self.send_cap * 40)
        self.ppart = PeerBuffer(pg, self.recv_cap * 16)
        rows = self.rows[1] - self.rows[0]
        self.pslab = PeerBuffer(pg, rows * self.w * 16)
        self.slab = self.pslab.tensor(torch.complex128, (rows, self.w))
        return self.slab

    def route_peer(self, u, v, wbin, vis=None, keep_index=False):
        """route() with the records pulled over NVLink by the destination's copy engines: counts -> all-gather of the count
        table (the one host synchronisation) -> pack into the peer-visible send buffer -> barrier -> every rank pulls its
        segment out of every peer's send buffer."""
        from . import device as dv
        gh, qpx = self.table.shape[-2], self.table.shape[-3]
        P, me = self.world, self.rank
        _mark("route: start")
        counts = dv.route_count(self.h, qpx, gh, self.bounds, v)
        table = torch.empty((P, P), dtype=torch.int32, device=v.device)
        dist.all_gather_into_tensor(table.view(-1), counts, group=self.group)
        cnt = table.tolist()                                      # cnt[s][g]: records rank s sends to rank g
        W = 3 if vis is None else 5
        sent = sum(cnt[me])
        recv = sum(cnt[s][me] for s in range(P))
        if sent > self.send_cap or recv > self.recv_cap:
            raise RuntimeError(f"route_peer: {sent} records to send / {recv} to receive exceed the capacities {self.send_cap} / {self.recv_cap}")
        _mark("route: counts on host")
        self.pg.barrier()                                         # every peer has finished pulling the previous batch
        send_view = self.psend.tensor(torch.float64, (self.send_cap * 5,))
        _, sidx = dv.route_pack(self.h, qpx, gh, self.bounds, u, v, wbin, vis, cnt[me], keep_index=keep_index, out=send_view)
        _mark("route: packed")
        self.pg.barrier()                                         # all send buffers are packed
        rec = torch.empty((recv, W), dtype=torch.float64, device=v.device)
        copies, off = [], 0
        for s in range(P):
            seg = sum(cnt[s][:me])                                # start of my segment in rank s's send buffer
            n = cnt[s][me]
            if n:
                copies.append((rec.data_ptr() + off * W * 8, self.psend.ptrs[s] + seg * W * 8, n * W * 8))
            off += n
        self.pg.bulk(copies)
        _mark("route: pulled")
        return rec, {"sidx": sidx, "cnt": cnt}

    def return_peer(self, count, route):
        """The owners' partial sums (in the peer-visible `ppart`, one per received record) travel back: every source pulls the
        values of its records from every owner and adds them per visibility."""
        from . import device as dv
        P, me = self.world, self.rank
        cnt = route["cnt"]
        sent = sum(cnt[me])
        back = torch.empty(sent, dtype=torch.complex128, device=self.slab.device)
        _mark("return: start")
        self.pg.barrier()                                         # all owners have written their partial sums
        copies, seg = [], 0
        for g in range(P):
            n = cnt[me][g]
            off = sum(cnt[s][g] for s in range(me))               # where my records start in owner g's receive order
            if n:
                copies.append((back.data_ptr() + seg * 16, self.ppart.ptrs[g] + off * 16, n * 16))
            seg += n
        self.pg.bulk(copies)
        _mark("return: pulled")
        out = torch.zeros(count, dtype=torch.complex128, device=self.slab.device)
        dv.scatter_add_(out, route["sidx"], back)
        _mark("return: added")
        return out

    def partial_view(self, n):
        return self.ppart.tensor(torch.complex128, (n,))

    def all_nonzero_rows(self):
        """nonzero_rows() of every rank (collective; host list of (lo, hi))."""
        nz = torch.tensor(self.nonzero_rows(), dtype=torch.int64, device=self.table.device)
        if self.world == 1:
            return [tuple(int(x) for x in nz.tolist())]
        out = torch.empty((self.world, 2), dtype=torch.int64, device=self.table.device)
        dist.all_gather_into_tensor(out.view(-1), nz, group=self.group)
        return [(int(a), int(b)) for a, b in out.tolist()]

    def grid_peer(self, u, v, wbin, vis):
        """grid(..., keep_route=True) over peer memory, into the peer-visible slab of enable_peer (zeroed here)."""
        rec, route = self.route_peer(u, v, wbin, vis, keep_index=True)
        self.last_routed, self._last_rec, self._route, self._count = int(rec.shape[0]), rec, route, int(u.numel())
        self.slab.zero_()
        if rec.shape[0] > 0:
            self._fill(rec).grid(self.table, self.slab)
        return self.slab

    def degrid_routed_peer(self):
        """Adjoint of the last grid_peer at the same coordinates, on whatever the slab holds now: the routed plan is reused,
        the partial sums return through peer memory."""
        partial = self.partial_view(max(self.last_routed, 1))[:self.last_routed]
        if self.last_routed > 0:
            self._plan.degrid(self.table, self.slab, partial)
        return self.return_peer(self._count, self._route)

    def image_peer(self, all_nonzero, want_image=False, sync_max=True):
        """slab_grid_to_image on the peer-visible slab with the transpose pulled over NVLink (peer_slab_grid_to_image).
        all_nonzero: the (lo, hi) non-zero row interval of EVERY rank (host list; static per data set)."""
        return peer_slab_grid_to_image(self.pg, self.pslab, self.bounds[:-1], all_nonzero, self.h, group=self.group, want_image=want_image,
                                       sync_max=sync_max)
