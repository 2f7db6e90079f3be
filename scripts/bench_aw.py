#!/usr/bin/env python
"""Times the AW path (configs 1-3 of BASELINE.json) on the real-shaped stand-in R' of SURVEY.md 8d: N = 2400
(theta 0.008 x lam 300000 as src/ImageDataset.hs:32-33), S = 15, Q = 8, nw = 64, 64 antennas, through the
host-pointer C ABI (skagrid_aw_gridding: uvw_lambda, doweight, mirror, aw_imaging, hermitian, ifft, real, max).
The real SKA1_Low_*.h5 files are git-LFS stubs in the reference, hence the stand-in.  `measure` is what bench.py's `aw`
sub-record calls; as a script it prints one JSON line:  bench_aw.py [V] [--oracle]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

WORKLOAD = "R' stand-in for configs 1-3: skagrid_aw_gridding (pre-steps + AW gridding + image), N=2400, S=15, Q=8, nw=64, nant=64"
THETA, LAM, NW, Q, S, NANT = 0.008, 300000, 64, 8, 15, 64


def standin(V):
    from ska_sdp_accelerate_gridding_b200 import gridding as G
    rng = np.random.default_rng(20261018)
    ko = G.KernelOptions(qpx=Q, npixFF=128, npixKern=S)
    wbins = np.linspace(-3000.0, 3000.0, NW)
    wk = G.w_kernel(THETA, wbins, ko)
    yy, xx = np.mgrid[-7:8, -7:8]
    ak = np.stack([np.exp(-(xx ** 2 + yy ** 2) / (8.0 + 0.1 * a)) * np.exp(1j * 0.05 * (a % 7) * xx + 1j * 0.03 * (a % 5) * yy) for a in range(NANT)])
    ak /= np.abs(ak).sum(axis=(1, 2), keepdims=True)
    freq = 1.0e8
    sc = 299792458.0 / freq
    r = np.abs(rng.normal(0, 0.12, V)) * 0.45 * LAM
    ang = rng.uniform(0, 2 * np.pi, V)
    uvw_m = np.stack([np.clip(r * np.cos(ang), -0.49 * LAM, 0.49 * LAM) * sc, np.clip(r * np.sin(ang), -0.49 * LAM, 0.49 * LAM) * sc,
                      rng.uniform(-2900, 2900, V) * sc], axis=1)
    a1, a2 = rng.integers(0, NANT, V), rng.integers(0, NANT, V)
    vis = rng.standard_normal(V) + 1j * rng.standard_normal(V)
    return wk, wbins, ak, uvw_m, a1, a2, freq, vis


def measure(V, oracle=False, oracle_n=2000):
    import torch
    from ska_sdp_accelerate_gridding_b200 import image_dataset as D
    from ska_sdp_accelerate_gridding_b200 import gridding as G
    from ska_sdp_accelerate_gridding_b200.context import get_context
    ctx = get_context(torch.cuda.current_device())
    wk, wbins, ak, uvw_m, a1, a2, freq, vis = standin(V)
    times = []
    for i in range(4):
        t0 = time.perf_counter()
        mx, img, _ = D.aw_gridding_arrays(THETA, LAM, wk, wbins, ak, uvw_m, a1, a2, freq, vis, want_image=True, ctx=ctx)
        times.append((time.perf_counter() - t0, ctx.last_device_ms))
    wall, devms = min(t[0] for t in times[1:]), min(t[1] for t in times[1:])
    out = {"vis": V, "wall_ms": wall * 1e3, "device_ms": devms, "vis_per_s_e2e": V / wall, "vis_per_s_device": V / (devms * 1e-3), "image_max": mx,
           # formation = one S x S AW kernel per visibility by two direct same-convolutions: 2 * 8 * S^4 flop, pairs de-duplicated
           "formation_flop_per_vis_algorithmic": 2 * 8 * S ** 4,
           "note": "wall/device: pageable numpy buffers (what a plain caller passes); *_pinned: w-kernels and image in page-locked memory"}
    # the same call with the two large buffers (14.7 MB of w-kernels in, 46 MB of image out) page-locked
    wk_pin = torch.from_numpy(wk).pin_memory().numpy()
    side = G._grid_side(THETA, LAM)
    img_pin = torch.empty((side, side), dtype=torch.float64).pin_memory().numpy()
    times = []
    for i in range(4):
        t0 = time.perf_counter()
        D.aw_gridding_arrays(THETA, LAM, wk_pin, wbins, ak, uvw_m, a1, a2, freq, vis, out_image=img_pin, ctx=ctx)
        times.append((time.perf_counter() - t0, ctx.last_device_ms))
    wallp, devp = min(t[0] for t in times[1:]), min(t[1] for t in times[1:])
    out.update({"wall_ms_pinned": wallp * 1e3, "device_ms_pinned": devp, "vis_per_s_e2e_pinned": V / wallp,
                "pinned_image_rel_diff": float(np.abs(img_pin - img).max() / np.abs(img).max())})
    if oracle:
        from oracle import oracle as orc
        n = min(V, oracle_n)
        t0 = time.perf_counter()
        oimg, omx, _ = orc.aw_gridding(THETA, LAM, wk, wbins, ak, uvw_m[:n, 0], uvw_m[:n, 1], uvw_m[:n, 2], a1[:n], a2[:n], freq, vis[:n])
        out["cpu_port_vis_per_s_1core"] = n / (time.perf_counter() - t0)
        mx2, img2, _ = D.aw_gridding_arrays(THETA, LAM, wk, wbins, ak, uvw_m, a1, a2, freq, vis, n=n, want_image=True, ctx=ctx)
        out["parity_image_max_abs_err_over_peak_first_%d" % n] = float(np.abs(img2 - oimg).max() / np.abs(oimg).max())
        out["parity_max_rel_err"] = float(abs(mx2 - omx) / abs(omx))
    return out


if __name__ == "__main__":
    V = int(float(sys.argv[1])) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 100000
    r = measure(V, oracle="--oracle" in sys.argv)
    r["workload"] = WORKLOAD
    print(json.dumps(r))
