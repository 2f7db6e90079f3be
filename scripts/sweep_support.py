#!/usr/bin/env python
"""Gridder / degridder throughput versus kernel support (device-resident, config-4 geometry otherwise: 8192^2 grid,
Q=8, 32 w-planes, 2e7 synthetic core-dominated visibilities).  One JSON line per support."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ska_sdp_accelerate_gridding_b200 import device as dv  # noqa: E402


def main():
    n, q, nw, cnt = 8192, 8, 32, 20_000_000
    for s in (7, 9, 13, 15, 21, 31, 33, 63):
        npixff = 128 if s <= 31 else 256
        table = dv.w_kernel_table(0.01, np.linspace(-300.0, 300.0, nw), npixff, s, q)
        u, v, wb, vis = dv.synth_vis(20261018, 0, cnt, n, s, nw)
        plan = dv.Plan(n, n, table.shape, u, v, wb, vis)
        grid = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
        out = torch.empty(cnt, dtype=torch.complex128, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for _ in range(2):
            plan.update(u, v, wb, vis, check=False); plan.grid(table, grid); plan.degrid(table, grid, out)
        torch.cuda.synchronize()
        ev[0].record(); plan.update(u, v, wb, vis, check=False); ev[1].record(); plan.grid(table, grid); ev[2].record(); plan.degrid(table, grid, out); ev[3].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        print(json.dumps({"support": s, "taps": s * s, "plan_ms": t[0], "grid_ms": t[1], "degrid_ms": t[2], "grid_vis_per_s": cnt / (t[1] * 1e-3),
                          "degrid_vis_per_s": cnt / (t[2] * 1e-3), "grid_tflops_fp64": 8 * s * s * cnt / (t[1] * 1e-3) / 1e12,
                          "grid_taps_per_s": s * s * cnt / (t[1] * 1e-3)}))
        plan.close()
        del table, u, v, wb, vis, grid, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
