"""Host-side logic of the two multi-GPU modes on CPU: world_size-2 gloo process groups.  The gridding inside
each rank is done by the CPU oracle here (the CUDA path needs a GPU); what is tested is the partitioning,
the owner computation, the all-to-all routing and the reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ska_sdp_accelerate_gridding_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.default_rng(77)
    n, s, q, nw, cnt = 160, 15, 4, 2, 3000
    gcf = rng.standard_normal((nw, q, q, s, s)) + 1j * rng.standard_normal((nw, q, q, s, s))
    u, v = rng.uniform(-0.55, 0.55, cnt), rng.uniform(-0.55, 0.55, cnt)
    wb = rng.integers(0, nw, cnt)
    vis = rng.standard_normal(cnt) + 1j * rng.standard_normal(cnt)
    return n, s, q, gcf, u, v, wb, vis


def _worker(rank, world, port, mode, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        n, s, q, gcf, u, v, wb, vis = _problem()
        full = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vis, wbin=wb)
        first, cnt = D.shard_range(len(u), rank, world)
        sl = slice(first, first + cnt)
        if mode == "vis":
            local = orc.convgrid(gcf, np.zeros((n, n), complex), u[sl], v[sl], vis[sl], wbin=wb[sl])
            g = torch.from_numpy(local)
            D.allreduce_grid(g)
            err = np.abs(g.numpy() - full).max() / np.abs(full).max()
            ret[rank] = ("ok", float(err), cnt)
        else:
            bounds = D.slab_bounds(n, world)
            y, _ = orc.frac_coord(n, q, v[sl])
            y0 = torch.from_numpy(y - s // 2)
            payload = (torch.from_numpy(u[sl].copy()), torch.from_numpy(v[sl].copy()), torch.from_numpy(wb[sl].copy()),
                       torch.from_numpy(vis[sl].copy()))
            (ru, rv, rwb, rvis), route = D.route_by_rows(y0, s, bounds, payload, return_route=True)
            r0, r1 = bounds[rank], bounds[rank + 1]
            # every received footprint intersects the owned slab, and nothing that does was left behind
            ry, _ = orc.frac_coord(n, q, rv.numpy())
            assert ((ry - s // 2 + s > r0) & (ry - s // 2 < r1)).all()
            ay, _ = orc.frac_coord(n, q, v)
            expected = int(((ay - s // 2 + s > r0) & (ay - s // 2 < r1)).sum())
            assert ru.numel() == expected, (ru.numel(), expected)
            # grid the routed visibilities, clip to the slab, compare with the same rows of the full grid
            g = orc.convgrid(gcf, np.zeros((n, n), complex), ru.numpy(), rv.numpy(), rvis.numpy(), wbin=rwb.numpy())
            err = np.abs(g[r0:r1] - full[r0:r1]).max() / np.abs(full).max()
            # adjoint: every owner degrids only its rows of a model grid; the partial sums go back and add up
            rng = np.random.default_rng(5)
            model = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
            clipped = np.zeros_like(model)
            clipped[r0:r1] = model[r0:r1]
            part = orc.convdegrid(gcf, clipped, ru.numpy(), rv.numpy(), wbin=rwb.numpy())
            back = D.return_to_source(torch.from_numpy(part), route, cnt).numpy()
            ref = orc.convdegrid(gcf, model, u[sl], v[sl], wbin=wb[sl])
            err = max(err, np.abs(back - ref).max() / np.abs(ref).max())
            ret[rank] = ("ok", float(err), int(ru.numel()))
    except Exception as e:  # pragma: no cover
        ret[rank] = ("fail", repr(e), 0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["vis", "tile"])
def test_two_rank_gloo(mode):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, mode, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        status, err, cnt = ret[r]
        assert status == "ok", err
        assert err < 1e-12
    if mode == "vis":
        assert sum(ret[r][2] for r in range(world)) == 3000
    else:
        assert sum(ret[r][2] for r in range(world)) >= 2600  # some footprints straddle the slab boundary, some fall off the grid


def test_shard_range_and_bounds():
    for count in (0, 1, 7, 100, 12345):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == count
            for (f0, n0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + n0 == f1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    b = D.slab_bounds(32768, 8)
    assert b[0] == 0 and b[-1] == 32768 and all(x % 32 == 0 for x in b) and all(b[i] < b[i + 1] for i in range(8))
    assert D.slab_bounds(100, 3) == [0, 32, 64, 100]


def test_owners_of_rows():
    bounds = [0, 32, 64, 100]
    y0 = torch.tensor([-20, -15, -14, 0, 17, 18, 31, 50, 63, 85, 99, 100, 120])
    lo, hi, on = D.owners_of_rows(y0, 15, bounds)
    assert on.tolist() == [False, False, True, True, True, True, True, True, True, True, True, False, False]
    assert lo.tolist()[2:11] == [0, 0, 0, 0, 0, 1, 1, 2, 2]
    assert hi.tolist()[2:11] == [0, 0, 0, 1, 1, 2, 2, 2, 2]


def test_balanced_slab_bounds():
    hist = torch.zeros(1000, dtype=torch.int64)
    hist[500:520] = 100          # dense core
    hist[520:1000] = 1
    b = D.balanced_slab_bounds(hist, 4)
    assert b[0] == 0 and b[-1] == 1000 and all(b[i] < b[i + 1] for i in range(4))
    counts = [int(hist[b[i]:b[i + 1]].sum()) for i in range(4)]
    assert max(counts) - min(counts) <= 200          # within two rows of the densest region
    # degenerate: everything in one row still yields valid, strictly increasing bounds
    h2 = torch.zeros(64, dtype=torch.int64); h2[10] = 5
    b2 = D.balanced_slab_bounds(h2, 8)
    assert b2[0] == 0 and b2[-1] == 64 and all(b2[i] < b2[i + 1] for i in range(8))


def test_weighted_slab_bounds_and_active_rows():
    cost = torch.zeros(1000, dtype=torch.float64)
    cost[500:520] = 50.0         # cheap dense core rows ...
    cost[520:1000] = 4.0         # ... expensive sparse ones
    b = D.weighted_slab_bounds(cost, 4)
    assert b[0] == 0 and b[-1] == 1000 and all(b[i] < b[i + 1] for i in range(4))
    parts = [float(cost[b[i]:b[i + 1]].sum()) for i in range(4)]
    assert max(parts) - min(parts) <= 2 * 50.0
    assert D.weighted_slab_bounds(torch.zeros(16, dtype=torch.float64), 4)[-1] == 16   # degenerate: still valid bounds
    # active rows: clamped to the grid and widened to equal slabs inside it
    assert D.active_row_slabs(8192, 4088, 8192, 8) == (4088, 513)
    lo, m = D.active_row_slabs(100, 90, 100, 4)
    assert lo + 4 * m <= 100 and lo <= 90 and lo + 4 * m >= 100
    assert D.active_row_slabs(64, 10, 5, 2) == (0, 1)                  # empty interval
    with pytest.raises(ValueError):
        D.active_row_slabs(4, 0, 4, 8)


def _transpose_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        rng = np.random.default_rng(3)
        n = 48
        g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        bounds = [0, 7, 29, n][: world] + [n] if world == 3 else [0, 19, n]   # uneven row slabs
        r0, r1 = bounds[rank], bounds[rank + 1]
        # the slab-distributed grid -> image of distributed.slab_grid_to_image, its two device stages restated in numpy:
        # weight 2 except on row 0 / column 0 (replaces make_grid_hermitian when only the real part is kept), centring, ifft along x
        yy, xx = np.mgrid[r0:r1, 0:n]
        slab = g[r0:r1] * np.where((yy == 0) | (xx == 0), 1.0, 2.0) * np.where((yy + xx) % 2 == 1, -1.0, 1.0)
        slab = np.fft.ifft(slab, axis=1) * n   # cuFFT's inverse transform is unnormalised
        cols, (c0, c1) = D.rows_to_columns(torch.from_numpy(slab), (r0, r1), n)
        assert tuple(cols.shape) == (n, c1 - c0)
        # rows nobody sends are zero: rank 0 keeps back its first three rows, rank 1 its last two
        lo, hi = (r0 + 3, r1) if rank == 0 else (r0, r1 - 2)
        part, _ = D.rows_to_columns(torch.from_numpy(slab[lo - r0:hi - r0].copy()), (lo, hi), n)
        expect = cols.clone()
        expect[:3] = 0
        expect[n - 2:] = 0
        assert torch.equal(part, expect)
        colsn = np.fft.ifft(cols.numpy(), axis=0) * n
        yy, xx = np.mgrid[0:n, c0:c1]
        img = colsn.real * np.where((yy + xx) % 2 == 1, -1.0, 1.0) / (n * n)
        ref = np.real(orc.ifft(orc.make_grid_hermitian(g)))[:, c0:c1]
        ret[rank] = ("ok", float(np.abs(img - ref).max() / np.abs(ref).max()), c1 - c0)
    except Exception as e:  # pragma: no cover
        ret[rank] = ("fail", repr(e), 0)
    finally:
        dist.destroy_process_group()


def test_slab_grid_to_image_algorithm_two_rank_gloo():
    """The all-to-all transpose of slab_grid_to_image on uneven row slabs, and the identity it rests on:
    real(ifft(make_grid_hermitian g)) == real(ifft(g weighted 2 off row/column 0)), against the oracle."""
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_transpose_worker, args=(world, port, ret), nprocs=world, join=True)
    assert sum(ret[r][2] for r in range(world)) == 48
    for r in range(world):
        status, err, _ = ret[r]
        assert status == "ok", err
        assert err < 1e-12
