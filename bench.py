#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config 4:

    synthetic SKA1-Low-shaped visibilities, 8192^2 complex128 grid, support 15, oversampling 8, 32 w-planes,
    visibility-sharded over N GPUs with an NCCL grid reduction.

One step = one full imaging major-cycle pass over one batch of V visibilities per GPU:
    bin + uv-tile bucket sort  ->  tiled gridder  ->  (N>1: NCCL all-reduce of the grid)
    ->  hermitian + centred inverse FFT + real/max (grid -> image)  ->  degridder (adjoint) of the same batch.
`value` = visibilities gridded+degridded per second over all GPUs, inputs resident in HBM.
`e2e`   = the same pass through the host-pointer C ABI (skagrid_convgrid2 / skagrid_grid_to_image /
          skagrid_convdegrid2, the functions the reference's Haskell layer would bind) from pinned host
          buffers, every host<->device copy inside the timed region.
`--impl reference` times the CPU restatement of the reference semantics (oracle/, OpenMP) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
N_GRID, SUPPORT, QPX, NW = 8192, 15, 8, 32
THETA, NPIXFF = 0.01, 128
WMAX = 300.0
FLOP_PER_VIS = 8 * SUPPORT * SUPPORT
BYTES_PER_VIS = 64                                   # SURVEY.md 8d compulsory record bytes
UPD_BYTES_PER_VIS = 64 + 16 * SUPPORT * SUPPORT      # grid-update-equivalent accounting


def workload_name():
    return (("config 4: " if (N_GRID, SUPPORT, NW) == (8192, 15, 32) else "") +
            f"synthetic SKA1-Low-shaped visibilities, {N_GRID}^2 c128 grid, support {SUPPORT}, oversampling {QPX}, {NW} w-planes, "
            "visibility-sharded (one batch per GPU) with NCCL all-reduce of the grid")


def parse():
    global N_GRID, SUPPORT, NW, FLOP_PER_VIS, UPD_BYTES_PER_VIS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--vis", type=float, default=1e8, help="visibilities per GPU per step")
    ap.add_argument("--uniform", action="store_true", help="uniform uv coverage instead of the core-dominated mixture")
    ap.add_argument("--variant", type=int, default=0, help="gridder variant (0 tiled, 1 atomic scatter, 2 tiled with two tap loads in flight)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-vis", type=float, default=None, help="visibilities per e2e step (default: --vis)")
    ap.add_argument("--cpu-sample", type=float, default=4e6)
    ap.add_argument("--grid", type=int, default=N_GRID, help="grid side (config 4: 8192; config 5: 32768)")
    ap.add_argument("--support", type=int, default=SUPPORT, help="kernel support (config 4: 15; config 5: 31)")
    ap.add_argument("--nw", type=int, default=NW, help="w-planes in the kernel table")
    ap.add_argument("--mode", default="vis", choices=["vis", "tile"],
                    help="vis: visibility-sharded + all-reduce (config 4, the headline); tile: uv-tile-sharded + routing (config 5, gridding only)")
    args = ap.parse_args()
    N_GRID, SUPPORT, NW = args.grid, args.support, args.nw
    FLOP_PER_VIS = 8 * SUPPORT * SUPPORT
    UPD_BYTES_PER_VIS = 64 + 16 * SUPPORT * SUPPORT
    return args


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_rate(sample, u, v, wb, vis, table):
    """The reference semantics on the host cores (oracle/oracle.c, OpenMP): grid + degrid of `sample`
    visibilities on the same 8192^2 grid and kernel table.  Returns (vis/s, threads, seconds)."""
    from oracle import oracle as orc
    orc.use_all_cores()
    grid0 = np.zeros((N_GRID, N_GRID), np.complex128)
    t0 = time.perf_counter()
    g = orc.convgrid(table, grid0, u, v, vis, wbin=wb, parallel=True)
    t1 = time.perf_counter()
    orc.convdegrid(table, g, u, v, wbin=wb, parallel=True)
    t2 = time.perf_counter()
    return sample / (t2 - t0), orc.num_threads(), (t1 - t0, t2 - t1)


def synth_host(count, first=0):
    """Host copy of the device-generated synthetic visibilities (the generator lives in libskagrid.so)."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    u, v, wb, vis = dv.synth_vis(SEED, first, count, N_GRID, SUPPORT, NW)
    table = dv.w_kernel_table(THETA, np.linspace(-WMAX, WMAX, NW), NPIXFF, SUPPORT, QPX)
    torch.cuda.synchronize()
    return u.cpu().numpy(), v.cpu().numpy(), wb.cpu().numpy(), vis.cpu().numpy(), table.cpu().numpy()


def run_reference(args, rank):
    """--impl reference: rank 0 only; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    sample = int(args.cpu_sample)
    u, v, wb, vis, table = synth_host(sample)
    rates, t_g, t_d, threads = [], [], [], 1
    for i in range(args.warmup + args.steps):
        r, threads, (tg, td) = cpu_reference_rate(sample, u, v, wb, vis, table)
        if i >= args.warmup:
            rates.append(r); t_g.append(tg); t_d.append(td)
    ms = 1e3 * sample / float(np.mean(rates))
    val = sample / (ms * 1e-3)
    desc = f"{sample} of the workload's visibilities per step (grid + degrid on the same {N_GRID}^2 grid and kernel table), CPU restatement of the reference semantics"
    print(json.dumps({
        "impl": "reference", "metric": "visibilities/sec gridded+degridded", "value": val, "unit": "vis/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(), "sample": desc},
        "cpu_baseline": {"value": val, "unit": "vis/s", "cores": threads, "kind": "port", "sample": desc,
                         "grid_s": float(np.mean(t_g)), "degrid_s": float(np.mean(t_d))},
        "e2e": {"value": val, "unit": "vis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------ GPU arm
def run_tile_mode(args, rank, world, dev, ctx, table, u, v, wb, vis):
    """Config 5 shape: uv-tile-sharded gridding.  step = owner computation + all-to-all routing + bin/bucket +
    tiled gridder into the owned row slab; no grid reduction.  value = visibilities gridded per second."""
    import torch
    import torch.distributed as dist
    from ska_sdp_accelerate_gridding_b200 import distributed as D
    V = int(args.vis)
    ts = D.TileShardedGridder(N_GRID, N_GRID, table)
    bounds = ts.balance(v)   # once per data set: slabs with equal visibility counts (the uv coverage is known up front)
    slab = torch.zeros((ts.rows[1] - ts.rows[0], N_GRID), dtype=torch.complex128, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    kept = [0]

    def step():
        slab.zero_()
        ts.grid(u, v, wb, vis, out=slab)   # owners -> all-to-all -> bin+bucket -> tiled gridder (plan reused across steps)
        kept[0] = ts.last_routed

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    l0 = ctx.launch_count
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = t0.elapsed_time(t1)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        k = torch.tensor([kept[0]], dtype=torch.int64, device=dev)
        dist.all_reduce(k)
        kept[0] = int(k.item())
    ms_per_step = ms / args.steps
    # grid -> image straight from the row slabs (no gather): row transforms, all-to-all transpose, column transforms
    nz = ts.nonzero_rows()  # rows the gridder cannot have touched (v >= 0 after mirroring: half of the grid) are skipped
    img_ms = []
    for i in range(3):
        work = slab.clone()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = ev(), ev()
        a.record()
        _, _, mx = D.slab_grid_to_image(work, bounds, want_image=True, nonzero=nz)
        b.record()
        torch.cuda.synchronize()
        img_ms.append(a.elapsed_time(b))
        del work
    image_ms = min(img_ms[1:])
    if world > 1:
        t = torch.tensor([image_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        image_ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "slab_grid_to_image_ms": image_ms, "image_max": mx, "nonzero_rows_rank0": list(nz),
            "metric": "visibilities/sec gridded (uv-tile-sharded)", "value": world * V / (ms_per_step * 1e-3), "unit": "vis/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config 5 shape: {N_GRID}^2 c128 grid, support {SUPPORT}, oversampling {QPX}, {NW} w-planes, uv-tile-sharded "
                                   "(row slabs, all-to-all routing, no grid reduce)", "vis_per_gpu_per_step": V, "routed_records": kept[0], "slab_bounds": bounds,
                       "step": "owners (bit-exact y cell) -> all-to-all -> bin+bucket -> tiled gridder into the owned slab"},
            "gpu_launches": int(ctx.launch_count - l0), "clocks": clocks, "e2e": None}))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    torch.cuda.set_device(local_rank)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from ska_sdp_accelerate_gridding_b200 import device as dv
    from ska_sdp_accelerate_gridding_b200 import gridding as G
    from ska_sdp_accelerate_gridding_b200.context import get_context

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = get_context(local_rank)
    dev = torch.device("cuda", local_rank)
    V = int(args.vis)

    table = dv.w_kernel_table(THETA, np.linspace(-WMAX, WMAX, NW), NPIXFF, SUPPORT, QPX)
    u, v, wb, vis = dv.synth_vis(SEED, rank * V, V, N_GRID, SUPPORT, NW, uniform=args.uniform)
    if args.mode == "tile":
        run_tile_mode(args, rank, world, dev, ctx, table, u, v, wb, vis)
        return
    grid = torch.zeros((N_GRID, N_GRID), dtype=torch.complex128, device=dev)
    vis_out = torch.empty(V, dtype=torch.complex128, device=dev)
    plan = dv.Plan(N_GRID, N_GRID, table.shape, u, v, wb, vis)
    stats = plan.stats()
    fp64_peak = ctx.fp64_tflops()
    # measured L2->SM read bandwidth over a table-sized buffer: a coalesced stream, and (S=15) the kernels' own access
    # pattern -- 15 taps of 15 consecutive 256-byte rows of a random slice per half-warp, useful bytes only
    l2_stream = ctx.l2_read_tbs(int(table.numel()) * 16, 0)
    l2_peak = ctx.l2_read_tbs(int(table.numel()) * 16, 1) if SUPPORT == 15 else l2_stream

    ev = lambda: torch.cuda.Event(enable_timing=True)
    stage_ms = {k: [] for k in ("plan", "grid", "reduce", "image", "degrid")}

    def step(record):
        e = [ev() for _ in range(6)]
        e[0].record()
        plan.update(u, v, wb, vis)          # bit-exact binning + bucket sort (part of gridding)
        e[1].record()
        grid.zero_()
        plan.grid(table, grid, variant=args.variant)
        e[2].record()
        if world > 1:
            dist.all_reduce(torch.view_as_real(grid))
        e[3].record()
        _, mx = dv.grid_to_image(grid, want_image=False)   # in place: the buffer now holds the transformed plane
        e[4].record()
        plan.degrid(table, grid, vis_out)   # adjoint pass over the same batch; the transformed plane stands in for the model grid
        e[5].record()
        if record:
            torch.cuda.synchronize()
            for k, a, b in (("plan", 0, 1), ("grid", 1, 2), ("reduce", 2, 3), ("image", 3, 4), ("degrid", 4, 5)):
                stage_ms[k].append(e[a].elapsed_time(e[b]))

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    torch.cuda.synchronize()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        step(False)
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = t0.elapsed_time(t1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * V / (ms_per_step * 1e-3)

    # per-stage breakdown and the dominant kernel's launch duration (CUDA events on the launching stream)
    for _ in range(max(2, min(3, args.steps))):
        step(True)
    sm = {k: float(np.mean(val)) for k, val in stage_ms.items()}
    grid_ms = sm["grid"]
    # isolate the gridder kernel from the grid.zero_() memset that shares its bracket
    ez = [ev() for _ in range(3)]
    ez[0].record(); grid.zero_(); ez[1].record(); plan.grid(table, grid, variant=args.variant); ez[2].record()
    torch.cuda.synchronize()
    kern_ms = ez[1].elapsed_time(ez[2])
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    alg_bytes = BYTES_PER_VIS * V + 16 * N_GRID * N_GRID + table.numel() * 16
    traffic = None   # dram__bytes_read+write of the gridder per launch, from the committed ncu capture of this workload
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_v21_traffic.json")
    if os.path.exists(tpath) and N_GRID == 8192 and SUPPORT == 15 and not args.uniform and args.variant == 0:
        tj = json.load(open(tpath))
        if int(tj.get("vis_per_launch", 0)) == V:
            k = tj["grid_tiled_kernel"]
            traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "grid_tiled_kernel<R=16,MT=2,DEPTH=3,TY=8>" if SUPPORT <= 15 else "grid_tiled_kernel<R=32,MT=2,DEPTH=3,TY=32>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "traffic": traffic, "algorithmic_bytes": alg_bytes, "peak_source": peak_src, "kernel_ms": kern_ms,
        "note": "compulsory-byte accounting (64 B/vis + 16 N^2 + table); the tiled gridder is bound by L2->SM kernel-tap traffic (3.7 KB/vis), see DESIGN.md 4.2 and the l2_taps entry",
        "fp64": {"achieved_tflops": FLOP_PER_VIS * V / (kern_ms * 1e-3) / 1e12, "peak_tflops_measured": fp64_peak,
                 "frac": FLOP_PER_VIS * V / (kern_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None},
        "l2_taps": {"achieved_tbs": 16.0 * SUPPORT * SUPPORT * V / (kern_ms * 1e-3) / 1e12, "peak_tbs_measured": l2_peak,
                    "frac": 16.0 * SUPPORT * SUPPORT * V / (kern_ms * 1e-3) / 1e12 / l2_peak if l2_peak else None,
                    "peak_tbs_coalesced_stream": l2_stream,
                    "note": "kernel taps streamed L2->SM (16 B x S^2 per visibility) vs the L2 read bandwidth measured in this run with the same "
                            "access pattern and nothing else (skagrid_measure_l2_pattern_tbs: random 15x15-tap slices of a table-sized buffer, "
                            "useful bytes); this is the ceiling of any gridder that fetches one kernel slice per visibility from L2"},
        "hbm_update_equiv": {"achieved": UPD_BYTES_PER_VIS * V / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak,
                             "frac": UPD_BYTES_PER_VIS * V / (kern_ms * 1e-3) / 1e9 / hbm_peak},
    }

    out = {
        "metric": "visibilities/sec gridded+degridded", "value": value, "unit": "vis/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_name(),
            "vis_per_gpu_per_step": V, "uv": "uniform" if args.uniform else "core-dominated mixture (SURVEY 8d)", "seed": SEED,
            "step": "bin+bucket -> tiled gridder -> all-reduce (N>1) -> hermitian+IFFT+real/max -> degridder",
            "l2": f"inputs ({V * 40 / 1e9:.1f} GB) and grid ({N_GRID * N_GRID * 16 / 1e9:.2f} GB) exceed the 126 MB L2; no explicit flush",
            "gridder_variant": args.variant, "plan": stats,
        },
        "stages_ms": sm,
        "rates": {"grid_vis_per_s_per_gpu": V / ((sm["plan"] + sm["grid"]) * 1e-3), "gridder_kernel_vis_per_s": V / (kern_ms * 1e-3),
                  "degrid_vis_per_s_per_gpu": V / (sm["degrid"] * 1e-3), "grid_to_image_ms": sm["image"]},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }

    # ---------------------------------------------------------------- e2e through the host-pointer C ABI
    if not args.no_e2e:
        Ve = int(args.e2e_vis) if args.e2e_vis else V
        pin = lambda t: t.cpu().pin_memory()
        hu, hv, hwb, hvis = pin(u[:Ve]), pin(v[:Ve]), pin(wb[:Ve]), pin(vis[:Ve])
        htab = pin(table)
        hgrid = torch.zeros((N_GRID, N_GRID), dtype=torch.complex128).pin_memory()
        hout = torch.empty(Ve, dtype=torch.complex128).pin_memory()
        hmax = np.zeros(1)
        nu, nv, nwb, nvis, ntab, ngrid, nout = (x.numpy() for x in (hu, hv, hwb, hvis, htab, hgrid, hout))
        lib, h = ctx.lib, ctx.h
        p = lambda a: a.ctypes.data
        # free the device-resident working set of the first measurement so both fit comfortably
        plan.close()
        del u, v, wb, vis, vis_out, grid
        torch.cuda.empty_cache()

        # conv_imaging2 takes uvw in wavelengths and divides by lam itself (src/Gridding.hs:115-124): theta*lam = N_GRID
        E2E_THETA, E2E_LAM = 0.01, N_GRID * 100
        hu_wl, hv_wl = (hu * float(E2E_LAM)).pin_memory(), (hv * float(E2E_LAM)).pin_memory()
        nu_wl, nv_wl = hu_wl.numpy(), hv_wl.numpy()

        e2e_parts = []

        def e2e_step():
            tp = [time.perf_counter()]
            # grid pointers are NULL: the grid stays resident in the context between the three calls (include/skagrid.h),
            # so only visibilities, w-plane indices and the kernel table cross PCIe -- what the reference's single fused
            # Accelerate program does (`use` the inputs, return the result)
            ctx.check(lib.skagrid_conv_imaging2(h, NW, QPX, SUPPORT, SUPPORT, p(ntab), E2E_THETA, E2E_LAM, Ve, p(nu_wl), p(nv_wl), p(nu_wl),
                                                p(nwb), p(nvis), None))
            tp.append(time.perf_counter())
            ctx.check(lib.skagrid_grid_to_image(h, N_GRID, None, None, p(hmax)))
            tp.append(time.perf_counter())
            # ... and the coordinates too (u = v = wbin = NULL: those conv_imaging2 uploaded, already divided by lam)
            ctx.check(lib.skagrid_convdegrid2(h, NW, QPX, SUPPORT, SUPPORT, p(ntab), N_GRID, N_GRID, None, Ve, None, None, None, p(nout)))
            tp.append(time.perf_counter())
            e2e_parts.append([1e3 * (b - a) for a, b in zip(tp, tp[1:])])

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        if world > 1:
            dist.barrier()
        ts = time.perf_counter()
        ksteps = max(1, min(args.steps, 3))
        for _ in range(ksteps):
            e2e_step()
        torch.cuda.synchronize()
        te = (time.perf_counter() - ts) / ksteps
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        grid_b = N_GRID * N_GRID * 16
        parts = np.mean(np.array(e2e_parts[-ksteps:]), axis=0)
        out["e2e"] = {"value": world * Ve / te, "unit": "vis/s", "ms_per_step": te * 1e3, "vis_per_gpu_per_step": Ve,
                      "calls_ms": {"conv_imaging2": float(parts[0]), "grid_to_image": float(parts[1]), "convdegrid2": float(parts[2])},
                      "h2d_bytes_per_step": int(Ve * 40 + 2 * ntab.nbytes), "d2h_bytes_per_step": int(Ve * 16 + 8),
                      "api": "skagrid_conv_imaging2 (vis -> grid) + skagrid_grid_to_image (grid -> max) + skagrid_convdegrid2 (grid -> vis): host pointers "
                             "(pinned) for visibilities / indices / table / results; the grid and the uploaded coordinates stay resident in the "
                             "context between the calls (NULL pointers), so every input crosses PCIe once per step"}
    else:
        out["e2e"] = None

    # ---------------------------------------------------------------- CPU baseline beside it (rank 0, N=1 only)
    if not args.no_cpu and world == 1 and rank == 0:
        sample = int(min(args.cpu_sample, V))
        cu, cv, cwb, cvis, ctab = synth_host(sample)
        r, threads, (tg, td) = cpu_reference_rate(sample, cu, cv, cwb, cvis, ctab)
        out["cpu_baseline"] = {"value": r, "unit": "vis/s", "cores": threads, "kind": "port",
                               "sample": f"first {sample} visibilities of the same workload, grid + degrid on the host (oracle/oracle.c, OpenMP)",
                               "grid_s": tg, "degrid_s": td}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
