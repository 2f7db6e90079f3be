#!/usr/bin/env python
"""Micro-benchmark of the exchange primitives between one-process-per-GPU ranks (torchrun): how fast does every rank get
`--mb` MiB from (or to) EACH peer at the same time -- copy engines or SM kernels, pull or push -- next to NCCL's all-gather of
the same volume and the peer-sum kernel.  Decides which primitive carries which exchange step (profiles/r02_peer_primitives_*.json)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ska_sdp_accelerate_gridding_b200.peer import PeerBuffer, PeerGroup
    pg = PeerGroup()
    seg = a.mb << 20
    buf = PeerBuffer(pg, seg * world)          # slot k of rank r's buffer: data for / from rank k
    dstbuf = PeerBuffer(pg, seg * world)
    local_t = buf.tensor(torch.float64, (seg * world // 8,))
    local_t.copy_(torch.arange(local_t.numel(), dtype=torch.float64, device="cuda") + rank)
    others = [p for p in range(world) if p != rank]
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn):
        best = None
        for _ in range(a.reps + 1):
            torch.cuda.synchronize(); dist.barrier()
            pg.barrier()
            e0, e1 = ev(), ev()
            e0.record(); fn(); pg.barrier(); e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t) if best is None else min(best, float(t))
        return best

    res = {}
    vol = seg * len(others) / 1e9      # GB moved into (pull) or out of (push) every rank
    pulls = [(dstbuf.local + p * seg, buf.ptrs[p] + rank * seg, seg) for p in others]          # my slot of every peer -> local
    pushes = [(dstbuf.ptrs[p] + rank * seg, buf.local + p * seg, seg) for p in others]         # local slot p -> peer p's slot `rank`
    res["ce_pull_ms"] = timed(lambda: pg.pull(pulls))
    res["ce_push_ms"] = timed(lambda: pg.pull(pushes))
    res["sm_pull_ms"] = timed(lambda: pg.gather(pulls))
    res["sm_push_ms"] = timed(lambda: pg.gather(pushes))
    res["peer_sum_ms"] = timed(lambda: pg.peer_sum_(buf, rank * seg, seg // 16))
    res["peer_sum_bcast_ms"] = timed(lambda: pg.peer_sum_(buf, rank * seg, seg // 16, broadcast=True))
    w, cw = 16384, 16384 // world       # bytes per row of the source, bytes per row pulled (a column block)
    rows = seg // w
    c2d = [(dstbuf.local + p * rows * cw, cw, buf.ptrs[p] + rank * cw, w, cw, rows) for p in others]
    res["ce_pull2d_ms"] = timed(lambda: pg.pull(c2d))
    res["sm_pull2d_ms"] = timed(lambda: pg.gather2d(c2d))
    res["pull2d_mb_per_peer"] = rows * cw / 2**20
    src = torch.empty(seg // 8, dtype=torch.float64, device="cuda")
    out = torch.empty(seg * world // 8, dtype=torch.float64, device="cuda")
    res["nccl_all_gather_ms"] = timed(lambda: dist.all_gather_into_tensor(out, src))
    full = torch.empty(seg * world // 8, dtype=torch.float64, device="cuda")
    part = torch.empty(seg // 8, dtype=torch.float64, device="cuda")
    res["nccl_reduce_scatter_ms"] = timed(lambda: dist.reduce_scatter_tensor(part, full))
    a2a_in = torch.empty(seg * world // 8, dtype=torch.float64, device="cuda")
    a2a_out = torch.empty_like(a2a_in)
    res["nccl_all_to_all_ms"] = timed(lambda: dist.all_to_all_single(a2a_out, a2a_in))
    res["barrier_ms"] = timed(lambda: None)
    if rank == 0:
        out = {"world": world, "mib_per_peer": a.mb, "gb_per_rank": vol, "ms": res,
               "gb_per_s_per_rank": {k[:-3]: vol / ((v - res["barrier_ms"]) * 1e-3) for k, v in res.items() if k.endswith("_ms") and k not in ("barrier_ms", "ce_pull2d_ms", "sm_pull2d_ms")},
               "note": "every rank moves mib_per_peer to/from each of its world-1 peers at once; max over ranks, best of reps; the closing device barrier is "
                       "inside the timed region (barrier_ms, subtracted in gb_per_s)"}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
