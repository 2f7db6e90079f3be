"""Peer-memory plumbing of the one-process-per-GPU modes (csrc/ipc.cu): exchange buffers every rank can dereference over
NVLink, a stream-ordered device barrier, copy-engine pulls.  torch.distributed is used ONCE per buffer, to swap the 64-byte
CUDA IPC handles; the data path afterwards is the library's own kernels and cudaMemcpyAsync."""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from .context import Context
from .device import _DevArray, _stream, context_for_current_device

_TYPESTR = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}


class PeerBuffer:
    """`nbytes` of device memory on every rank of `pg` (sizes may differ per rank), opened on all the others.
    ptrs[k] = device address of rank k's buffer as seen from this process."""

    def __init__(self, pg: "PeerGroup", nbytes: int):
        self.pg, self.ctx, self.nbytes = pg, pg.ctx, max(int(nbytes), 256)
        lib, h = self.ctx.lib, self.ctx.h
        handle = (C.c_ubyte * 64)()
        p = C.c_void_p()
        problem = None
        if lib.skagrid_ipc_alloc(h, self.nbytes, C.byref(p), handle) != 0:
            if pg.world == 1:
                self.ctx.check(-3)
            problem = f"rank {pg.rank} cannot allocate {self.nbytes} bytes: " + lib.skagrid_last_error(h).decode()   # reported collectively below
        self.local = int(p.value or 0)
        dev = torch.device("cuda", self.ctx.device)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        every = torch.empty(64 * pg.world, dtype=torch.uint8, device=dev)
        if pg.world > 1:
            dist.all_gather_into_tensor(every, mine, group=pg.group)
        else:
            every.copy_(mine)
        host = every.cpu().numpy()
        self.ptrs = []
        for k in range(pg.world):
            if k == pg.rank:
                self.ptrs.append(self.local)
                continue
            if problem is not None or not host[64 * k:64 * (k + 1)].any():   # this rank, or rank k, has no buffer to share
                self.ptrs.append(0)
                continue
            hk = (C.c_ubyte * 64)(*host[64 * k:64 * (k + 1)].tolist())
            q = C.c_void_p()
            rc = lib.skagrid_ipc_open(h, hk, C.byref(q))
            if rc != 0 and problem is None:
                problem = f"rank {pg.rank} cannot open the buffer of rank {k}: " + lib.skagrid_last_error(h).decode()
            self.ptrs.append(int(q.value) if rc == 0 else 0)
        if pg.world > 1:
            # agree on the outcome: a rank that could not map a peer must not leave the others waiting at a barrier
            bad = torch.tensor([1 if problem else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=pg.group)   # also: nobody frees before everybody has opened
            if int(bad.item()):
                for k, q in enumerate(self.ptrs):
                    if k != pg.rank and q:
                        lib.skagrid_ipc_close(h, C.c_void_p(q))
                dist.barrier(group=pg.group)
                if self.local:
                    lib.skagrid_ipc_free(h, C.c_void_p(self.local))
                self.ptrs = None
                raise RuntimeError("peer memory is not available between these devices" + (": " + problem if problem else " (another rank failed)"))

    def tensor(self, dtype, shape, offset_bytes=0):
        """A torch view (no copy) of the LOCAL buffer."""
        itemsize = torch.empty(0, dtype=torch.float64 if dtype == torch.complex128 else dtype).element_size()
        if dtype == torch.complex128:
            n = 1
            for s in shape:
                n *= int(s)
            if offset_bytes < 0 or offset_bytes + n * 16 > self.nbytes:
                raise ValueError("view exceeds the peer buffer")
            t = torch.as_tensor(_DevArray(self.local + offset_bytes, tuple(shape) + (2,), "<f8"), device=torch.device("cuda", self.ctx.device))
            return torch.view_as_complex(t)
        n = 1
        for s in shape:
            n *= int(s)
        if offset_bytes + n * itemsize > self.nbytes:
            raise ValueError("view exceeds the peer buffer")
        return torch.as_tensor(_DevArray(self.local + offset_bytes, shape, _TYPESTR[dtype]), device=torch.device("cuda", self.ctx.device))

    def close(self):
        if getattr(self, "ptrs", None) is None:
            return
        torch.cuda.synchronize()
        if self.pg.world > 1:
            dist.barrier(group=self.pg.group)   # nobody is still reading
        for k, q in enumerate(self.ptrs):
            if k != self.pg.rank:
                self.ctx.lib.skagrid_ipc_close(self.ctx.h, C.c_void_p(q))
        if self.pg.world > 1:
            dist.barrier(group=self.pg.group)   # everybody has unmapped before the owner frees
        self.ctx.lib.skagrid_ipc_free(self.ctx.h, C.c_void_p(self.local))
        self.ptrs = None


class _Pending:
    def __init__(self, streams):
        self.streams = streams

    def wait(self):
        main = torch.cuda.current_stream()
        for s in self.streams:
            main.wait_stream(s)
        self.streams = []


class PeerGroup:
    """The ranks of a torch.distributed group as NVLink peers: flag buffer for the device barrier, side streams for pulls."""

    def __init__(self, group=None, ctx: Context | None = None):
        self.group = group
        self.ctx = ctx or context_for_current_device()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.flags = PeerBuffer(self, 4096)
        self._flag_ptrs = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in self.flags.ptrs])
        self.epoch = 0
        ns = max(4, min(self.world - 1, 8))
        self.pools = [[torch.cuda.Stream() for _ in range(ns)] for _ in range(2)]   # two pools: pulls of different pools never queue behind each other
        self.streams = self.pools[0]
        self._fork = torch.cuda.Event()

    def ctx_sm_count(self):
        return int(torch.cuda.get_device_properties(self.ctx.device).multi_processor_count)

    def barrier(self):
        """Stream-ordered barrier on the current stream: the kernels enqueued after it on ANY rank start only when every
        rank's stream has reached it.  No host synchronisation."""
        if self.world == 1:
            return
        self.epoch += 1
        self.ctx.check(self.ctx.lib.skagrid_dev_peer_barrier(self.ctx.h, self.world, self.rank, self._flag_ptrs, self.epoch & 0xFFFFFFFF, _stream()))

    def peer_sum_(self, buf: PeerBuffer, offset_bytes: int, ncomplex: int, broadcast: bool = False):
        """local[offset ...] += sum over the other ranks of theirs[offset ...] (complex128 values), one kernel, all peers in flight.
        broadcast: the same kernel stores the sum into every peer's buffer too (reduce-scatter + all-gather fused)."""
        others = [buf.ptrs[k] + offset_bytes for k in range(self.world) if k != self.rank]
        arr = (C.c_void_p * max(len(others), 1))(*[C.c_void_p(p) for p in others])
        self.ctx.check(self.ctx.lib.skagrid_dev_peer_sum(self.ctx.h, len(others), arr, C.c_void_p(buf.local + offset_bytes), int(ncomplex), int(bool(broadcast)), _stream()))

    def bulk(self, copies):
        """Bulk pull nothing overlaps with: the SM gather kernel, or copy engines (joined).  SKAGRID_PEER_PULL = ce | sm overrides
        the default for A/B measurements."""
        mode = os.environ.get("SKAGRID_PEER_PULL", "")
        if mode == "sm" or (mode != "ce" and self.world >= 4):   # measured: copy engines win between two GPUs, the SM kernel from four on
            self.gather(copies)
        else:
            self.pull(copies)

    def gather_async(self, copies, max_blocks=0):
        """`gather` on a side stream, after everything already on the current stream; returns a handle whose wait() joins it.
        max_blocks caps the kernel's grid so that what the current stream launches next runs beside it."""
        main = torch.cuda.current_stream()
        side = self.pools[1][0]
        self._fork.record(main)
        side.wait_event(self._fork)
        with torch.cuda.stream(side):
            self.gather(copies, max_blocks=max_blocks)
        return _Pending([side])

    def gather(self, copies, max_blocks=0):
        """copies: iterable of (dst address, src address, bytes), 8-byte aligned: ONE kernel on the current stream in which the
        SMs pull (or, with peer addresses as destinations, push) every segment through peer memory, all peers at once
        (skagrid_dev_peer_gather)."""
        copies = [c for c in copies if c[2] > 0]
        if not copies:
            return
        n = len(copies)
        dst = (C.c_void_p * n)(*[C.c_void_p(c[0]) for c in copies])
        src = (C.c_void_p * n)(*[C.c_void_p(c[1]) for c in copies])
        nb = (C.c_int64 * n)(*[int(c[2]) for c in copies])
        if all(c[2] % 16 == 0 and c[0] % 16 == 0 and c[1] % 16 == 0 for c in copies) and len({c[2] for c in copies}) == 1:
            # equal, 16-byte aligned segments (slabs): the strided kernel with one "row" per segment moves 16 bytes per load
            one = (C.c_int64 * n)(*[1] * n)
            sp = (C.c_int64 * n)(*[int(c[2]) for c in copies])
            self.ctx.check(self.ctx.lib.skagrid_dev_peer_gather2d(self.ctx.h, n, dst, int(copies[0][2]), src, sp, int(copies[0][2]), one, int(max_blocks), _stream()))
            return
        self.ctx.check(self.ctx.lib.skagrid_dev_peer_gather(self.ctx.h, n, dst, src, nb, int(max_blocks), _stream()))

    def gather2d(self, copies):
        """copies: iterable of (dst, dpitch, src, spitch, width_bytes, rows) with one common width and destination pitch, 16-byte
        aligned: ONE kernel on the current stream pulls every block out of peer memory (skagrid_dev_peer_gather2d)."""
        copies = [c for c in copies if c[5] > 0 and c[4] > 0]
        if not copies:
            return
        n = len(copies)
        dst = (C.c_void_p * n)(*[C.c_void_p(c[0]) for c in copies])
        src = (C.c_void_p * n)(*[C.c_void_p(c[2]) for c in copies])
        sp = (C.c_int64 * n)(*[int(c[3]) for c in copies])
        rows = (C.c_int64 * n)(*[int(c[5]) for c in copies])
        self.ctx.check(self.ctx.lib.skagrid_dev_peer_gather2d(self.ctx.h, n, dst, int(copies[0][1]), src, sp, int(copies[0][4]), rows, 0, _stream()))

    def pull(self, copies, join=True, pool=0):
        """copies: iterable of (dst address, src address, bytes) or (dst, dpitch, src, spitch, width_bytes, rows): enqueued
        round-robin on the side streams after everything already on the current stream.  join=True: the current stream
        waits for them; join=False: returns an object whose wait() does that later (the pulls overlap what is enqueued on
        the current stream in between).  pool: which of the two stream pools carries the copies."""
        main = torch.cuda.current_stream()
        self._fork.record(main)
        lib, h = self.ctx.lib, self.ctx.h
        streams = self.pools[pool]
        used = set()
        # large contiguous copies are cut into pieces so that several copy engines work on them when there are few peers
        pieces = []
        for c in copies:
            if len(c) == 3 and c[2] > (256 << 20) and len(copies) < len(streams):
                k = min(4, len(streams))
                step = (-(-c[2] // k) + 255) // 256 * 256
                for off in range(0, c[2], step):
                    pieces.append((c[0] + off, c[1] + off, min(step, c[2] - off)))
            else:
                pieces.append(c)
        for i, c in enumerate(pieces):
            s = streams[i % len(streams)]
            if i < len(streams):
                s.wait_event(self._fork)
            used.add(i % len(streams))
            sp = C.c_void_p(s.cuda_stream)
            if len(c) == 3:
                self.ctx.check(lib.skagrid_dev_peer_copy(h, C.c_void_p(c[0]), C.c_void_p(c[1]), int(c[2]), sp))
            else:
                self.ctx.check(lib.skagrid_dev_peer_copy2d(h, C.c_void_p(c[0]), int(c[1]), C.c_void_p(c[2]), int(c[3]), int(c[4]), int(c[5]), sp))
        pending = _Pending([streams[i] for i in sorted(used)])
        if join:
            pending.wait()
            return None
        return pending

    def close(self):
        self.flags.close()
