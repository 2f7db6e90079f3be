#!/usr/bin/env python
"""End-to-end rate of the single-process multi-GPU C ABI (skagrid_*_mgpu_{vis,tile}, csrc/mgpu.cu) on bench.py's
workload (config 4: 8192^2 grid, S = 15, Q = 8, 32 w-planes): ONE host thread, N contexts, every visibility starts and
ends in pinned host memory.  Usage: bench_mgpu_abi.py [--gpus N] [--vis V_total] [--mode vis|tile] [--steps K].

  vis : skagrid_convgrid2_mgpu_vis(grid = NULL: zero start, sum left resident on every device)
        + skagrid_convdegrid2_mgpu_vis(grid = NULL) + skagrid_grid_to_image(ctx 0, NULL)
  tile: skagrid_convgrid2_mgpu_tile + skagrid_convdegrid2_mgpu_tile with the grid in pinned host memory (each device
        only ever holds its row slab)

Prints one JSON line (wall clock around the calls; they return when all devices are done)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402  (constants and the workload generator of the headline bench)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--vis", type=float, default=1e8)
    ap.add_argument("--mode", default="vis")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    from ska_sdp_accelerate_gridding_b200.multi_device import MultiDevice
    ndev = torch.cuda.device_count()
    P = a.gpus or ndev
    V = int(a.vis)
    torch.cuda.set_device(0)
    table = dv.w_kernel_table(B.THETA, np.linspace(-B.WMAX, B.WMAX, B.NW), B.NPIXFF, B.SUPPORT, B.QPX)
    pin = lambda t: t.cpu().pin_memory()
    hu = torch.empty(V, dtype=torch.float64).pin_memory()
    hv = torch.empty(V, dtype=torch.float64).pin_memory()
    hwb = torch.empty(V, dtype=torch.int64).pin_memory()
    hvis = torch.empty(V, dtype=torch.complex128).pin_memory()
    step = 1 << 24
    for f in range(0, V, step):  # generated on device 0 in pieces, parked in pinned host memory
        n = min(step, V - f)
        u, v, wb, vis = dv.synth_vis(B.SEED, f, n, B.N_GRID, B.SUPPORT, B.NW)
        hu[f:f + n].copy_(u); hv[f:f + n].copy_(v); hwb[f:f + n].copy_(wb); hvis[f:f + n].copy_(vis)
    del u, v, wb, vis
    htab = pin(table)
    del table
    torch.cuda.empty_cache()
    hout = torch.empty(V, dtype=torch.complex128).pin_memory()
    hgrid = torch.zeros((B.N_GRID, B.N_GRID), dtype=torch.complex128).pin_memory() if a.mode == "tile" else None
    hmax = np.zeros(1)
    p = lambda t: t.numpy().ctypes.data
    md = MultiDevice([i % ndev for i in range(P)])
    lib, hs = md.lib, md.handles
    N, S, Q, NW = B.N_GRID, B.SUPPORT, B.QPX, B.NW
    bounds = np.zeros(P + 1, np.int64)
    parts = {"grid": [], "image": [], "degrid": []}

    def one():
        t0 = time.perf_counter()
        if a.mode == "vis":
            md._check(lib.skagrid_convgrid2_mgpu_vis(hs, P, NW, Q, S, S, p(htab), N, N, None, V, p(hu), p(hv), p(hwb), p(hvis)))
            t1 = time.perf_counter()
            # grid and coordinates: what the gridding call left on the devices
            md._check(lib.skagrid_convdegrid2_mgpu_vis(hs, P, NW, Q, S, S, p(htab), N, N, None, V, None, None, None, p(hout)))
            t2 = time.perf_counter()
            # last, because grid_to_image transforms context 0's resident copy in place
            md._check(lib.skagrid_grid_to_image(md.ctxs[0].h, N, None, None, hmax.ctypes.data))
            t3 = time.perf_counter()
            return t3 - t0, (t1 - t0, t3 - t2, t2 - t1)
        else:
            hgrid.zero_()
            t0 = time.perf_counter()
            md._check(lib.skagrid_convgrid2_mgpu_tile(hs, P, NW, Q, S, S, p(htab), N, N, p(hgrid), V, p(hu), p(hv), p(hwb), p(hvis),
                                                      bounds.ctypes.data))
            t1 = t2 = time.perf_counter()
            md._check(lib.skagrid_convdegrid2_mgpu_tile(hs, P, NW, Q, S, S, p(htab), N, N, p(hgrid), V, p(hu), p(hv), p(hwb), p(hout),
                                                        bounds.ctypes.data))
        t3 = time.perf_counter()
        return t3 - t0, (t1 - t0, t2 - t1, t3 - t2)

    for _ in range(a.warmup):
        one()
    ts = []
    for _ in range(a.steps):
        t, (tg, ti, td) = one()
        ts.append(t); parts["grid"].append(tg); parts["image"].append(ti); parts["degrid"].append(td)
    t = float(np.mean(ts))
    print(json.dumps({
        "metric": "visibilities gridded+degridded per second, host to host, one process driving all devices through the C ABI",
        "value": V / t, "unit": "vis/s", "n_gpus": P, "distinct_devices": min(P, ndev), "mode": a.mode, "vis_total": V, "ms_per_step": t * 1e3,
        "stages_ms": {k: float(np.mean(x)) * 1e3 for k, x in parts.items()},
        "grid_vis_per_s": V / float(np.mean(parts["grid"])), "degrid_vis_per_s": V / float(np.mean(parts["degrid"])),
        "bounds": bounds.tolist() if a.mode == "tile" else None, "max_pixel": float(hmax[0]), "out_checksum": float(hout.real.sum()),
        "steps": a.steps, "warmup": a.warmup,
        "h2d_bytes_per_step": int(V * 40 + (V * 24 + 2 * N * N * 16 if a.mode == "tile" else 0)),
        "d2h_bytes_per_step": int(V * 16 + (N * N * 16 if a.mode == "tile" else 8)),
    }))
    md.close()


if __name__ == "__main__":
    main()
