"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Both sharding modes against the CPU oracle on a seeded problem.  Exit code 0 = parity within 1e-10 of the peak."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import oracle as orc
    from ska_sdp_accelerate_gridding_b200 import distributed as D

    rng = np.random.default_rng(2026)
    n, s, q, nw, cnt = 512, 15, 8, 4, 60000
    gcf = rng.standard_normal((nw, q, q, s, s)) + 1j * rng.standard_normal((nw, q, q, s, s))
    u, v = rng.uniform(-0.52, 0.52, cnt), rng.uniform(-0.52, 0.52, cnt)
    wb = rng.integers(0, nw, cnt)
    vis = rng.standard_normal(cnt) + 1j * rng.standard_normal(cnt)
    full = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vis, wbin=wb, parallel=True)
    peak = np.abs(full).max()
    first, m = D.shard_range(cnt, rank, world)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    lu, lv, lwb, lvis = t(u[first:first + m]), t(v[first:first + m]), t(wb[first:first + m]), t(vis[first:first + m])
    table = t(gcf)

    vs = D.VisShardedGridder(n, n, table)
    g = vs.grid(lu, lv, lwb, lvis).cpu().numpy()
    err_v = np.abs(g - full).max() / peak
    d = vs.degrid(t(full), lu, lv, lwb).cpu().numpy()
    od = orc.convdegrid(gcf, full, u[first:first + m], v[first:first + m], wbin=wb[first:first + m])
    err_d = np.abs(d - od).max() / np.abs(od).max()

    # reduce-scatter form: active rows only, slab-distributed grid -> image from the reduced slabs, all-gather for degridding
    lo, m_rows = vs.set_active_rows(lv)
    work = torch.full((n, n), 7.0 + 0j, dtype=torch.complex128, device="cuda")   # garbage: grid_slabs / gather_slabs must zero what they use
    rslab = torch.empty((m_rows, n), dtype=torch.complex128, device="cuda")
    vs.grid_slabs(lu, lv, lwb, lvis, work, rslab)
    a, b = vs.spans()[rank]
    err_v = max(err_v, np.abs(rslab.cpu().numpy() - full[a:b]).max() / peak)
    img, (c0, c1), mx = D.slab_grid_to_image(rslab.clone(), n, spans=vs.spans())
    oimg = np.real(orc.ifft(orc.make_grid_hermitian(full)))
    err_v = max(err_v, np.abs(img.cpu().numpy() - oimg[:, c0:c1]).max() / np.abs(oimg).max(), abs(mx - oimg.max()) / abs(oimg.max()))
    h = vs.gather_slabs(rslab, work, async_op=True)
    if h is not None:
        h.wait()
    err_v = max(err_v, np.abs(work.cpu().numpy() - full).max() / peak)
    d2 = vs.degrid(work).cpu().numpy()   # at the coordinates of the plan grid_slabs filled
    err_d = max(err_d, np.abs(d2 - od).max() / np.abs(od).max())

    # the same reduce-scatter / gather / transpose over NVLink peer memory (csrc/ipc.cu), no NCCL in the data path
    from ska_sdp_accelerate_gridding_b200.peer import PeerBuffer, PeerGroup
    pg = PeerGroup()
    vp = D.VisShardedGridder(n, n, table)
    vp.set_active_rows(lv)
    vp.enable_peer(pg)
    vp.work.fill_(5.0 - 2j)   # garbage in the active rows must not survive
    vp.work[:vp.active[0]].zero_(); vp.work[vp.active[0] + vp.active[1] * world:].zero_()
    for it in range(3):       # repeated passes exercise the barrier that protects the gathered grid; the last one is the fused form
        fused = it == 2       # reduce-scatter + all-gather in one kernel: no gather afterwards
        ps = vp.grid_slabs_peer(lu, lv, lwb, lvis, broadcast=fused)
        err_p = np.abs(ps.cpu().numpy() - full[a:b]).max() / peak
        pis = PeerBuffer(pg, ps.numel() * 16)
        pis.tensor(torch.complex128, tuple(ps.shape)).copy_(ps)
        hnd = None if fused else vp.gather_slabs_peer(join=False, sm=(it == 1))   # copy engines, then the SM kernel, then fused
        pimg, (pc0, pc1), pmx = D.peer_slab_grid_to_image(pg, pis, [x[0] for x in vp.spans()], vp.spans(), n)
        if hnd is not None:
            hnd.wait()
        err_p = max(err_p, np.abs(pimg.cpu().numpy() - oimg[:, pc0:pc1]).max() / np.abs(oimg).max(), abs(pmx - oimg.max()) / abs(oimg.max()))
        err_p = max(err_p, np.abs(vp.work.cpu().numpy() - full).max() / peak)
        dp = vp.degrid(vp.work).cpu().numpy()
        err_p = max(err_p, np.abs(dp - od).max() / np.abs(od).max())
        pis.close()

    ts = D.TileShardedGridder(n, n, table)
    slab_t = ts.grid(lu, lv, lwb, lvis, keep_route=True)
    slab = slab_t.cpu().numpy()
    r0, r1 = ts.rows
    err_t = np.abs(slab - full[r0:r1]).max() / peak
    # adjoint at the same coordinates, reusing the routed plan: partial sums return to the source ranks
    dr = ts.degrid_routed(t(full[r0:r1].copy())).cpu().numpy()
    err_t = max(err_t, np.abs(dr - od).max() / np.abs(od).max())
    ts.balance(lv)   # data-balanced slabs, then the adjoint with partial sums returned to the source ranks
    r0, r1 = ts.rows
    dt = ts.degrid(t(full[r0:r1]), lu, lv, lwb).cpu().numpy()
    err_td = np.abs(dt - od).max() / np.abs(od).max()
    slab2 = ts.grid(lu, lv, lwb, lvis).cpu().numpy()
    err_t = max(err_t, np.abs(slab2 - full[r0:r1]).max() / peak, err_td)
    # grid -> image straight from the row slabs of the tile-sharded gridder (no gather): columns of the oracle's image
    img, (c0, c1), mx = D.slab_grid_to_image(t(full[r0:r1].copy()), ts.bounds)
    err_t = max(err_t, np.abs(img.cpu().numpy() - oimg[:, c0:c1]).max() / np.abs(oimg).max(), abs(mx - oimg.max()) / abs(oimg.max()))
    # the same from the slab the gridder just filled, skipping the rows it cannot have touched (v >= 0 only here: half the grid is empty)
    hu, hv = np.abs(u[first:first + m]) * 0.9, np.abs(v[first:first + m]) * 0.9
    hfull = orc.convgrid(gcf, np.zeros((n, n), complex), np.abs(u) * 0.9, np.abs(v) * 0.9, vis, wbin=wb, parallel=True)
    hslab = ts.grid(t(hu), t(hv), lwb, lvis)
    nz = ts.nonzero_rows()
    img2, (c0, c1), mx2 = D.slab_grid_to_image(hslab, ts.bounds, nonzero=nz)
    himg = np.real(orc.ifft(orc.make_grid_hermitian(hfull)))
    err_t = max(err_t, np.abs(img2.cpu().numpy() - himg[:, c0:c1]).max() / np.abs(himg).max(), abs(mx2 - himg.max()) / abs(himg.max()))
    if nz[1] > nz[0] and nz[0] < n // 2 - s // 2 - 1:
        err_t = 1.0  # v >= 0: no row below the centre minus half a kernel can be non-zero
    # doweight over sharded visibilities: counts all-reduced between the two phases (bit-exact: integer counts, one division)
    theta, lam = 0.01, n * 100
    uw, vw = u * lam * 0.9, v * lam * 0.9
    wvis = t(vis[first:first + m].copy())
    D.doweight_sharded_(theta, lam, t(uw[first:first + m]), t(vw[first:first + m]), wvis)
    ow = orc.doweight(theta, lam, uw, vw, vis)[first:first + m]
    err_t = max(err_t, 0.0 if np.array_equal(wvis.cpu().numpy(), ow) else 1.0)
    # uv-tile-sharded over peer memory: routed records pulled by the owners, partial sums pulled back, transpose pulled
    tp = D.TileShardedGridder(n, n, table)
    tp.balance(t(hv))
    tp.enable_peer(pg, send_capacity=2 * m + 64, recv_capacity=cnt + 64)
    r0, r1 = tp.rows
    err_tp = 0.0
    for _ in range(2):
        pslab = tp.grid_peer(t(hu), t(hv), lwb, lvis)
        err_tp = max(err_tp, np.abs(pslab.cpu().numpy() - hfull[r0:r1]).max() / np.abs(hfull).max())
        hod = orc.convdegrid(gcf, hfull, hu, hv, wbin=wb[first:first + m])
        pslab.copy_(t(hfull[r0:r1].copy()))
        dpp = tp.degrid_routed_peer().cpu().numpy()
        err_tp = max(err_tp, np.abs(dpp - hod).max() / np.abs(hod).max())
        anz = tp.all_nonzero_rows()
        img3, (c0, c1), mx3 = tp.image_peer(anz, want_image=True)
        err_tp = max(err_tp, np.abs(img3.cpu().numpy() - himg[:, c0:c1]).max() / np.abs(himg).max(), abs(mx3 - himg.max()) / abs(himg.max()))
    from ska_sdp_accelerate_gridding_b200 import device as dv
    dv.check_errors()
    res = torch.tensor([err_v, err_d, err_t, err_p, err_tp], dtype=torch.float64, device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} vis-sharded grid err {res[0]:.2e}, degrid err {res[1]:.2e}, tile-sharded (grid, balanced grid, degrid) + slab grid->image + "
              f"sharded doweight err {res[2]:.2e}, vis-sharded over peer memory {res[3]:.2e}, tile-sharded over peer memory {res[4]:.2e}")
    dist.destroy_process_group()
    sys.exit(0 if float(res.max()) < 1e-10 else 1)


if __name__ == "__main__":
    main()
