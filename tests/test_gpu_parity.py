"""GPU parity tests: the CUDA path, called through the C ABI (host-pointer functions via the ctypes mirror of
src/Gridding.hs, and the device-resident entry points), against the CPU oracle on the same seeded inputs.
Integer / index results must be bit-exact; grids, images and visibilities within max-abs error <= 1e-10 of
the peak (BASELINE.json north_star)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


def rel_err(a, b):
    peak = np.abs(b).max()
    return np.abs(a - b).max() / (peak if peak > 0 else 1.0)


@pytest.fixture(scope="module")
def G():
    from ska_sdp_accelerate_gridding_b200 import gridding
    return gridding


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _rand_c(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------ binning
@pytest.mark.parametrize("n,qpx", [(2400, 8), (8192, 8), (32768, 8), (10, 1), (33, 4), (2400, 1)])
def test_frac_coord_bit_exact(G, orc, n, qpx):
    rng = np.random.default_rng(n + qpx)
    p = np.concatenate([rng.uniform(-0.5, 0.5, 200000), [0.0, 0.25, -0.25, 0.4999999, -0.5, 1.0 / 3.0],
                        (np.arange(-40, 40) + 0.5) / (n * qpx), np.arange(-40, 40) / (n * qpx * 2.0)])
    for flags, norm in ((1, True), (0, False)):
        fl, fr = G.frac_coord(n, qpx, p, flags=flags)
        ofl, ofr = orc.frac_coord(n, qpx, p, normalise=norm)
        assert np.array_equal(fl, ofl) and np.array_equal(fr, ofr)
    q = p[::-1].copy()
    x, xf, y, yf = G.frac_coords((n, n + 6), qpx, (p, q))
    ox, oxf, oy, oyf = orc.frac_coords(n, n + 6, qpx, p, q)
    assert np.array_equal(x, ox) and np.array_equal(xf, oxf) and np.array_equal(y, oy) and np.array_equal(yf, oyf)


def test_find_closest_bit_exact(G, orc):
    rng = np.random.default_rng(7)
    for nw in (1, 2, 3, 37, 541):
        ws = np.sort(rng.uniform(-3000, 3000, nw))
        w = np.concatenate([rng.uniform(-3500, 3500, 50000), ws, (ws[:-1] + ws[1:]) / 2, [ws[-1] + 1.0, ws[0] - 1.0]])
        assert np.array_equal(G.findClosest(ws, w), orc.find_closest(ws, w))


def test_empty_inputs(G):
    e = np.zeros(0)
    fl, fr = G.frac_coord(16, 4, e)
    assert fl.size == 0 and fr.size == 0
    g = G.convgrid(np.ones((2, 2, 3, 3), complex), np.zeros((8, 8), complex), (e, e), np.zeros(0, complex))
    assert not g.any()
    assert G.convdegrid(np.ones((2, 2, 3, 3), complex), np.ones((8, 8), complex), (e, e)).size == 0


# ------------------------------------------------------------------------------------------------ pre-steps
def test_prestep_parity(G, orc):
    from ska_sdp_accelerate_gridding_b200 import _lib
    from ska_sdp_accelerate_gridding_b200 import image_dataset as D
    rng = np.random.default_rng(11)
    cnt, theta, lam = 50000, 0.01, 30000
    u, v, w = (rng.uniform(-14000, 14000, cnt) for _ in range(3))
    vis = _rand_c(rng, cnt)
    ul = D.uvw_lambda(1.0e8, (u, v, w))
    ol = orc.uvw_lambda(1.0e8, u, v, w)
    assert all(np.array_equal(a, b) for a, b in zip(ul, ol))
    (mu, mv, mw), mvis = G.mirror_uvw((u, v, w), vis)
    ou, ov, ow, ovis = orc.mirror_uvw(u, v, w, vis)
    assert np.array_equal(mu, ou) and np.array_equal(mv, ov) and np.array_equal(mw, ow) and np.array_equal(mvis, ovis)
    wt = G.doweight(theta, lam, (u, v, w), vis)
    assert np.array_equal(wt, orc.doweight(theta, lam, u, v, vis))
    with pytest.raises(_lib.SkagridError):
        G.doweight(theta, lam, (u * 100, v, w), vis)  # outside the weight grid


def test_grid_simple_and_broken_numbers_golden(G, orc):
    j = json.load(open(os.path.join(GOLD, "broken_numbers.json")))
    # the permute (+) golden through the nearest-cell gridder: cell = n/2 + floor(0.5 + n p)  ->  p = (cell - 2) / 5
    k = np.arange(10)
    x, y = (2 * k) % 5, (3 * k + 1) % 5
    val = (k + 5) + 1j
    pu, pv = (x - 2) / 5.0, (y - 2) / 5.0
    g = np.zeros((5, 5), complex)
    for _ in range(2):
        g = G.grid(g, (pu, pv), val)
    assert np.array_equal(g, np.array(j["expected_re"]) + 1j * np.array(j["expected_im"]))
    rng = np.random.default_rng(12)
    u, v = rng.uniform(-0.5, 0.499, 30000), rng.uniform(-0.5, 0.499, 30000)
    vis = _rand_c(rng, 30000)
    a = G.grid(np.zeros((64, 64), complex), (u, v), vis)
    assert rel_err(a, orc.grid_simple(np.zeros((64, 64), complex), u, v, vis)) < TOL
    s = G.simple_imaging(0.01, 6400, (u * 6400, v * 6400, u), None, vis)
    assert rel_err(s, orc.simple_imaging(0.01, 6400, u * 6400, v * 6400, u, vis)) < TOL


# ------------------------------------------------------------------------------------------------ table gridders
CASES = [
    # n, s(h,w), qpx, nw, count, span   (span > 0.5: part of the footprints / visibilities fall off the grid)
    (96, (7, 7), 4, 3, 500, 0.55),
    (256, (15, 15), 8, 4, 20000, 0.5),
    (200, (15, 15), 8, 1, 5000, 0.6),
    (300, (31, 31), 8, 2, 3000, 0.52),
    (128, (9, 9), 2, 2, 4000, 0.5),
    (128, (13, 13), 4, 2, 4000, 0.5),
    (160, (5, 11), 3, 2, 3000, 0.5),
    (160, (33, 33), 2, 1, 600, 0.5),
    (192, (47, 47), 2, 2, 500, 0.5),
    (200, (49, 49), 1, 1, 300, 0.5),
    (224, (63, 63), 1, 1, 200, 0.5),
    (64, (1, 1), 1, 1, 1000, 0.5),
    (260, (65, 65), 1, 1, 100, 0.45),
]


@pytest.mark.parametrize("n,s,qpx,nw,count,span", CASES)
def test_convgrid2_and_degrid_parity(G, orc, n, s, qpx, nw, count, span):
    rng = np.random.default_rng(n * 31 + s[0])
    gcf = _rand_c(rng, (nw, qpx, qpx, s[0], s[1]))
    u, v = rng.uniform(-span, span, count), rng.uniform(-span, span, count)
    wb = rng.integers(0, nw, count)
    vis = _rand_c(rng, count)
    x, xf, y, yf = G.frac_coords((n, n), qpx, (u, v))
    ox = orc.frac_coords(n, n, qpx, u, v)
    assert all(np.array_equal(a, b) for a, b in zip((x, xf, y, yf), ox))
    start = _rand_c(rng, (n, n))
    g = G.convgrid2(gcf, start, (u, v), wb, vis)
    og = orc.convgrid(gcf, start, u, v, vis, wbin=wb)
    assert rel_err(g, og) < TOL
    d = G.convdegrid2(gcf, og, (u, v), wb)
    od = orc.convdegrid(gcf, og, u, v, wbin=wb)
    assert rel_err(d, od) < TOL
    if nw == 1:
        assert rel_err(G.convgrid(gcf[0], start, (u, v), vis), og) < TOL
        assert rel_err(G.convdegrid(gcf[0], og, (u, v)), od) < TOL


def test_golden_random_case(G):
    z = np.load(os.path.join(GOLD, "random_cases.npz"))
    n = int(z["n"])
    g = G.convgrid2(z["gcf"], np.zeros((n, n), complex), (z["u"], z["v"]), z["wbin"], z["vis"])
    assert rel_err(g, z["grid"]) < TOL
    d = G.convdegrid2(z["gcf"], z["grid"], (z["u"], z["v"]), z["wbin"])
    assert rel_err(d, z["degrid"]) < TOL
    x, xf, y, yf = G.frac_coords((n, n), 4, (z["u"], z["v"]))
    assert np.array_equal(x, z["x"]) and np.array_equal(xf, z["xf"]) and np.array_equal(y, z["y"]) and np.array_equal(yf, z["yf"])


def test_wbin_out_of_range_is_an_error(G):
    from ska_sdp_accelerate_gridding_b200 import _lib
    gcf = np.ones((2, 2, 2, 3, 3), complex)
    with pytest.raises(_lib.SkagridError):
        G.convgrid2(gcf, np.zeros((16, 16), complex), ([0.1], [0.1]), [2], [1 + 0j])


def test_dense_tile_many_chunks_and_variants(orc):
    """All visibilities in one uv tile (-> many work items flushing into the same cells) and the three gridder
    variants (tiled with / without L1 prefetch, atomic scatter) agree with the oracle."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    rng = np.random.default_rng(99)
    n, s, q, nw, cnt = 128, 15, 8, 2, 40000
    gcf = _rand_c(rng, (nw, q, q, s, s))
    u, v = rng.uniform(0.0, 20.0 / n, cnt), rng.uniform(0.0, 20.0 / n, cnt)
    wb = rng.integers(0, nw, cnt)
    vis = _rand_c(rng, cnt)
    og = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vis, wbin=wb, parallel=True)
    plan = dv.Plan(n, n, gcf.shape, _t(u), _t(v), _t(wb), _t(vis))
    st = plan.stats()
    assert st["kept"] == cnt and st["work_items"] >= cnt // 4096
    for variant in (0, 1, 2):
        grid = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
        plan.grid(_t(gcf), grid, variant=variant)
        assert rel_err(grid.cpu().numpy(), og) < TOL, variant
    d = plan.degrid(_t(gcf), _t(og)).cpu().numpy()
    assert rel_err(d, orc.convdegrid(gcf, og, u, v, wbin=wb, parallel=True)) < TOL


def test_row_slab_ownership(orc):
    """uv-tile-sharded mode: two plans owning complementary row slabs reproduce the full grid exactly once."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    rng = np.random.default_rng(5)
    n, s, q, nw, cnt = 192, 15, 4, 2, 20000
    gcf = _rand_c(rng, (nw, q, q, s, s))
    u, v = rng.uniform(-0.5, 0.5, cnt), rng.uniform(-0.5, 0.5, cnt)
    wb = rng.integers(0, nw, cnt)
    vis = _rand_c(rng, cnt)
    og = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vis, wbin=wb)
    g2 = _rand_c(rng, (n, n))
    od = orc.convdegrid(gcf, g2, u, v, wbin=wb)
    full = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
    dsum = torch.zeros(cnt, dtype=torch.complex128, device="cuda")
    kept = 0
    for rows in ((0, 70), (70, 192)):
        plan = dv.Plan(n, n, gcf.shape, _t(u), _t(v), _t(wb), _t(vis), rows=rows)
        kept += plan.stats()["kept"]
        slab = torch.zeros((rows[1] - rows[0], n), dtype=torch.complex128, device="cuda")
        plan.grid(_t(gcf), slab)
        full[rows[0]:rows[1]] += slab
        dsum += plan.degrid(_t(gcf), _t(g2[rows[0]:rows[1]]))  # partial sums over the owned rows
    assert rel_err(full.cpu().numpy(), og) < TOL
    assert rel_err(dsum.cpu().numpy(), od) < TOL
    assert cnt <= kept <= cnt * 1.3  # only the footprints straddling row 70 are processed twice


# ------------------------------------------------------------------------------------------------ AW path
def test_convolve2d_and_aw_kernel(G, orc):
    rng = np.random.default_rng(21)
    for n in (1, 3, 5, 7, 8, 9, 11, 13, 15, 16, 17, 19, 31):  # register-tiled kernel for odd n in 5..17, generic one otherwise
        a, b = _rand_c(rng, (n, n)), _rand_c(rng, (n, n))
        assert rel_err(G.convolve2d(a, b), orc.convolve2d(a, b)) < 1e-12
    # aw_kernel_fn2 in batches: distinct antenna pairs are convolved once (more visibilities than pairs, and fewer)
    # (batch sizes that are not multiples of the output kernels per block of the row-pair kernel, every tiled support)
    for nw, q, s, nant, cnt in ((3, 4, 15, 5, 203), (2, 2, 15, 40, 61), (2, 2, 9, 3, 50), (1, 1, 19, 4, 30), (2, 2, 5, 3, 27), (2, 2, 7, 3, 19),
                                (1, 2, 11, 3, 23), (2, 2, 13, 7, 33), (1, 2, 17, 3, 10)):
        wk, ak = _rand_c(rng, (nw, q, q, s, s)), _rand_c(rng, (nant, s, s))
        wb, yf, xf = rng.integers(0, nw, cnt), rng.integers(0, q, cnt), rng.integers(0, q, cnt)
        a1, a2 = rng.integers(0, nant, cnt), rng.integers(0, nant, cnt)
        out = G.aw_kernel_fn2(yf, xf, wk, ak, wb, a1, a2)
        for k in range(cnt):
            assert rel_err(out[k], orc.aw_kernel(wk[wb[k]], yf[k], xf[k], ak[a1[k]], ak[a2[k]])) < 1e-12
    from ska_sdp_accelerate_gridding_b200 import _lib
    bad = a1.copy()
    bad[3] = nant
    with pytest.raises(_lib.SkagridError) as e:
        G.aw_kernel_fn2(yf, xf, wk, ak, wb, bad, a2)
    assert e.value.code == -5


def test_smalltest_fixture(G):
    from tests.golden.make_golden import smalltest_inputs
    wk, ak, uvw, idx, vis = smalltest_inputs()
    stored = np.load(os.path.join(GOLD, "smalltest_oracle.npz"))
    for fn in (G.convgrid3, G.convgrid4):
        g = fn(wk, ak, np.zeros((10, 10), complex), (uvw[:, 0], uvw[:, 1], uvw[:, 2]), (idx[:, 0], idx[:, 1], idx[:, 2]), vis)
        assert rel_err(g, stored["grid"]) < TOL
        assert abs(abs(g[9, 6]) - 46.27454822566) < 1e-9


def test_aw_gridding_end_to_end(G, orc):
    from ska_sdp_accelerate_gridding_b200 import image_dataset as D
    rng = np.random.default_rng(33)
    theta, lam = 0.008, 30000   # N = 240
    nw, q, s, nant, cnt = 5, 4, 15, 6, 3000
    wk, ak = _rand_c(rng, (nw, q, q, s, s)) * 0.1, _rand_c(rng, (nant, s, s)) * 0.1
    wbins = np.linspace(-400.0, 400.0, nw)
    freq = 1.0e8
    sc = 299792458.0 / freq
    uvw_m = np.stack([rng.uniform(-12000, 12000, cnt) * sc, rng.uniform(-12000, 12000, cnt) * sc, rng.uniform(-380, 380, cnt) * sc], axis=1)
    a1, a2 = rng.integers(0, nant, cnt), rng.integers(0, nant, cnt)
    vis = _rand_c(rng, cnt)
    mx, img, grd = D.aw_gridding_arrays(theta, lam, wk, wbins, ak, uvw_m, a1, a2, freq, vis, want_grid=True)
    oimg, omx, ogrid = orc.aw_gridding(theta, lam, wk, wbins, ak, uvw_m[:, 0], uvw_m[:, 1], uvw_m[:, 2], a1, a2, freq, vis)
    assert rel_err(grd, ogrid) < TOL
    assert rel_err(img, oimg) < TOL
    assert abs(mx - omx) <= TOL * abs(omx)
    # the same through several chunks of the per-visibility kernel pipeline (pair ids are per chunk)
    os.environ["SKAGRID_AW_CHUNK"] = "700"
    try:
        mx2, img2, grd2 = D.aw_gridding_arrays(theta, lam, wk, wbins, ak, uvw_m, a1, a2, freq, vis, want_grid=True)
    finally:
        del os.environ["SKAGRID_AW_CHUNK"]
    assert rel_err(grd2, ogrid) < TOL and rel_err(img2, oimg) < TOL
    # aw_imaging alone, and its adjoint against the dot-product identity
    u, v, w = orc.uvw_lambda(freq, uvw_m[:, 0], uvw_m[:, 1], uvw_m[:, 2])
    g = G.aw_imaging(G.noArgs, G.noOtherArgs, theta, lam, wk, wbins, ak, (u, v, w), (a1, a2, None, None), vis)
    assert rel_err(g, orc.aw_imaging(theta, lam, wk, wbins, ak, u, v, w, a1, a2, vis)) < TOL
    wb = orc.find_closest(wbins, w)
    g2 = _rand_c(rng, g.shape)
    d = G.convdegrid3(wk, ak, g2, (u / lam, v / lam), (wb, a1, a2))
    assert rel_err(d, orc.convdegrid_aw(wk, ak, g2, u / lam, v / lam, wb, a1, a2)) < TOL
    lhs, rhs = np.vdot(g2, g), np.vdot(d, vis)
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)


def test_w_kernel_pattern_shift_and_transformation(G, orc):
    """KernelOptions.patHorShift / patVerShift / patTransMat (kernel_coordinates, src/Gridding.hs:620-635): the far field is no
    longer symmetric under transposition, so this also pins the transposing padder (:875) -- and its absence when qpx = 1 (:688)."""
    theta = 0.05
    t = np.array([[0.9, 0.2], [-0.1, 1.1]])
    for qpx, npixff, s, dl, dm, tm in ((4, 32, 9, 0, 0, t), (4, 32, 9, 0.01, -0.02, None), (1, 32, 9, 0.01, -0.02, t), (2, 64, 15, -0.015, 0.005, t)):
        ko = G.KernelOptions(qpx=qpx, npixFF=npixff, npixKern=s, patHorShift=dl, patVerShift=dm, patTransMat=tm)
        for w in (0.0, 37.5, -120.0):
            k = G.w_kernel(theta, w, ko)
            ok = orc.w_kernel(theta, w, npixff, s, qpx, dl=dl, dm=dm, transmat=tm)
            assert k.shape == ok.shape == (qpx, qpx, s, s)
            assert rel_err(k, ok) < 1e-11
    plain = G.w_kernel(theta, 37.5, G.KernelOptions(qpx=4, npixFF=32, npixKern=9))
    assert rel_err(plain, G.w_kernel(theta, 37.5, G.KernelOptions(qpx=4, npixFF=32, npixKern=9, patTransMat=np.eye(2)))) < 1e-15
    assert rel_err(plain, orc.w_kernel(theta, 37.5, 32, 9, 4)) < 1e-11


def test_aw_gridding_config1_shape(G, orc):
    """BASELINE.json configs 1-3 at THEIR shape (src/ImageDataset.hs:32-33: theta 0.008 x lam 300000 = a 2400^2 grid; S = 15, Q = 8,
    64 w-planes, 64 antennas: the R' stand-in of SURVEY 8d, generated by scripts/bench_aw.py) against the oracle, image and grid;
    2400 = 2^5 * 3 * 5^2 exercises the non-power-of-two transform and the 16-cell tile edge that does not divide the grid."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import bench_aw
    from ska_sdp_accelerate_gridding_b200 import image_dataset as D
    cnt = 2500
    wk, wbins, ak, uvw_m, a1, a2, freq, vis = bench_aw.standin(cnt)
    mx, img, grd = D.aw_gridding_arrays(bench_aw.THETA, bench_aw.LAM, wk, wbins, ak, uvw_m, a1, a2, freq, vis, want_grid=True)
    assert img.shape == (2400, 2400)
    oimg, omx, ogrid = orc.aw_gridding(bench_aw.THETA, bench_aw.LAM, wk, wbins, ak, uvw_m[:, 0], uvw_m[:, 1], uvw_m[:, 2], a1, a2, freq, vis)
    assert rel_err(grd, ogrid) < TOL
    assert rel_err(img, oimg) < TOL
    assert abs(mx - omx) <= TOL * abs(omx)
    # config 3: degridding of the model grid at the same coordinates (exact adjoint of the AW gridder)
    u, v, w = orc.uvw_lambda(freq, uvw_m[:, 0], uvw_m[:, 1], uvw_m[:, 2])
    wb = orc.find_closest(wbins, w)
    lam = float(bench_aw.LAM)
    d = G.convdegrid3(wk, ak, ogrid, (u / lam, v / lam), (wb, a1, a2))
    assert rel_err(d, orc.convdegrid_aw(wk, ak, ogrid, u / lam, v / lam, wb, a1, a2)) < TOL


# ------------------------------------------------------------------------------------------------ grid -> image
@pytest.mark.parametrize("n", [8, 9, 240, 255, 2400])
def test_hermitian_fft_image(G, orc, n):
    rng = np.random.default_rng(n)
    g = _rand_c(rng, (n, n))
    h = G.make_grid_hermitian(g)
    assert rel_err(h, orc.make_grid_hermitian(g)) < 1e-15
    assert rel_err(G.ifft(g), orc.ifft(g)) < 1e-12
    if n <= 255:
        assert rel_err(G.fft(g), orc.fft(g)) < 1e-12
    img, mx = G.grid_to_image(g)
    oimg = np.real(orc.ifft(orc.make_grid_hermitian(g)))
    assert rel_err(img, oimg) < TOL and abs(mx - oimg.max()) <= TOL * abs(oimg.max())


def test_w_kernel_and_w_cache_imaging(G, orc):
    ko = G.KernelOptions(qpx=4, npixFF=64, npixKern=15, wstep=200)
    ws = np.array([-400.0, 0.0, 250.0])
    k = G.w_kernel(0.05, ws, ko)
    for i, w in enumerate(ws):
        assert rel_err(k[i], orc.w_kernel(0.05, w, 64, 15, 4)) < 1e-11
    rng = np.random.default_rng(8)
    cnt, theta, lam = 2000, 0.05, 4000  # N = 200
    u, v, w = rng.uniform(-1800, 1800, cnt), rng.uniform(-1800, 1800, cnt), rng.uniform(-500, 500, cnt)
    vis = _rand_c(rng, cnt)
    g = G.w_cache_imaging(ko, G.noOtherArgs, theta, lam, (u, v, w), None, vis)
    rw = (200 * (np.sign(w / 200) * np.floor(np.abs(w / 200) + 0.5))).astype(np.int64)
    wmin = rw.min()
    steps = (rw.max() - wmin) // 200 + 1
    kern = np.stack([np.conj(orc.w_kernel(theta, float(i * 200 + wmin), 64, 15, 4)) for i in range(steps)])
    og = orc.convgrid(kern, np.zeros((200, 200), complex), u / lam, v / lam, vis, wbin=(rw - wmin) // 200)
    assert rel_err(g, og) < TOL
    # OtherImagingArgs.kernelFunction / kernelCache (src/Gridding.hs:401-412): a caller-supplied KernelF replaces w_kernel, called
    # with the four Maybe arguments as Nothing; kernelCache wins over kernelFunction when both are given
    calls = []

    def shifted(theta_, w_, a1, a2, t, f, kernops):
        assert (a1, a2, t, f) == (None, None, None, None)
        calls.append(w_)
        return orc.w_kernel(theta_, w_, 64, 15, 4, dl=0.01, dm=-0.005)

    g2 = G.w_cache_imaging(ko, G.OtherImagingArgs(kernelFunction=shifted), theta, lam, (u, v, w), None, vis)
    assert calls == [float(i * 200 + wmin) for i in range(steps)]
    kern2 = np.stack([np.conj(orc.w_kernel(theta, float(i * 200 + wmin), 64, 15, 4, dl=0.01, dm=-0.005)) for i in range(steps)])
    assert rel_err(g2, orc.convgrid(kern2, np.zeros((200, 200), complex), u / lam, v / lam, vis, wbin=(rw - wmin) // 200)) < TOL
    plain = lambda theta_, w_, a1, a2, t, f, kernops: orc.w_kernel(theta_, w_, 64, 15, 4)
    g3 = G.w_cache_imaging(ko, G.OtherImagingArgs(kernelFunction=shifted, kernelCache=plain), theta, lam, (u, v, w), None, vis)
    assert rel_err(g3, og) < TOL


def test_do_imaging(G, orc):
    rng = np.random.default_rng(9)
    cnt, theta, lam = 3000, 0.02, 6000  # N = 120
    uvw = np.stack([rng.uniform(-2500, 2500, cnt), rng.uniform(-2500, 2500, cnt), rng.uniform(-100, 100, cnt)], axis=1)
    vis = _rand_c(rng, cnt)
    z = np.zeros(cnt)
    drt, psf, pmax = G.do_imaging(theta, lam, uvw, z, z, z, 1e8, vis, G.simple_imaging)
    odrt, opsf, opmax = orc.do_imaging(theta, lam, uvw[:, 0], uvw[:, 1], uvw[:, 2], vis, orc.simple_imaging)
    assert rel_err(drt, odrt) < TOL and rel_err(psf, opsf) < TOL and abs(pmax - opmax) < TOL * abs(opmax)


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_linearity_and_adjoint(G):
    """BASELINE.json config 4 at FULL size (8192^2 grid, S=15, Q=8, 32 w-planes, 1e8 synthetic visibilities):
    size-independent properties -- tiled == atomic scatter, sum(grid) == sum_k vis_k * sum(kernel_k) for
    fully-inside footprints, <grid(v), g> == <v, degrid(g)>."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    n, s, q, nw, cnt = 8192, 15, 8, 32, 100_000_000
    table = dv.w_kernel_table(0.01, np.linspace(-300.0, 300.0, nw), 128, s, q)
    u, v, wb, vis = dv.synth_vis(20261018, 0, cnt, n, s, nw)
    plan = dv.Plan(n, n, table.shape, u, v, wb, vis)
    st = plan.stats()
    assert st["kept"] == cnt and st["dropped"] == 0
    g0 = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
    plan.grid(table, g0, variant=0)
    g1 = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
    plan.grid(table, g1, variant=1)
    peak = g1.abs().max().item()
    assert (g0 - g1).abs().max().item() <= TOL * peak
    # checksum: every footprint is inside the grid, so sum(grid) = sum_k vis_k * sum(table[slice_k])
    _, xf = dv.frac_coord(n, q, u)
    _, yf = dv.frac_coord(n, q, v)
    ksum = table.sum(dim=(-1, -2)).reshape(-1)
    sl = (wb * q + yf) * q + xf
    expect = (vis * ksum[sl]).sum().item()
    got = g0.sum().item()
    assert abs(got - expect) <= 1e-9 * max(abs(expect), peak)
    del g1
    g2 = torch.randn((n, n), dtype=torch.float64, device="cuda").to(torch.complex128)
    d = plan.degrid(table, g2)
    lhs = (g2.conj() * g0).sum().item()
    rhs = (d.conj() * vis).sum().item()
    assert abs(lhs - rhs) <= 1e-9 * abs(lhs)


# ------------------------------------------------------------------------------------------------ sharding modes
def test_tile_sharded_single_rank_and_device_frac_coord(orc):
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    from ska_sdp_accelerate_gridding_b200 import distributed as D
    rng = np.random.default_rng(41)
    n, s, q, nw, cnt = 256, 15, 8, 2, 10000
    gcf = _rand_c(rng, (nw, q, q, s, s))
    u, v = rng.uniform(-0.5, 0.5, cnt), rng.uniform(-0.5, 0.5, cnt)
    wb = rng.integers(0, nw, cnt)
    vis = _rand_c(rng, cnt)
    fl, fr = dv.frac_coord(n, q, _t(v))
    ofl, ofr = orc.frac_coord(n, q, v)
    assert np.array_equal(fl.cpu().numpy(), ofl) and np.array_equal(fr.cpu().numpy(), ofr)
    ts = D.TileShardedGridder(n, n, _t(gcf))
    g = ts.grid(_t(u), _t(v), _t(wb), _t(vis)).cpu().numpy()
    assert rel_err(g, orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vis, wbin=wb)) < TOL
    vs = D.VisShardedGridder(n, n, _t(gcf))
    assert rel_err(vs.grid(_t(u), _t(v), _t(wb), _t(vis)).cpu().numpy(), g) < TOL
    # the single-rank forms of the round-2 paths: routed records (packed plan update), degrid at the routed plan, active-row
    # slabs + slab grid -> image + gather
    model = _rand_c(rng, (n, n))
    ts.grid(_t(u), _t(v), _t(wb), _t(vis), keep_route=True)
    od = orc.convdegrid(gcf, model, u, v, wbin=wb)
    assert rel_err(ts.degrid_routed(_t(model)).cpu().numpy(), od) < TOL
    assert rel_err(ts.degrid(_t(model), _t(u), _t(v), _t(wb)).cpu().numpy(), od) < TOL
    hv = np.abs(v) * 0.8                      # mirrored coverage: the lower half of the grid stays empty
    hg = orc.convgrid(gcf, np.zeros((n, n), complex), u, hv, vis, wbin=wb)
    lo, m = vs.set_active_rows(_t(hv))
    assert lo >= n // 2 - s // 2 - 1 and lo + m <= n
    work = torch.full((n, n), 3.0 - 1j, dtype=torch.complex128, device="cuda")
    slab = torch.empty((m, n), dtype=torch.complex128, device="cuda")
    vs.grid_slabs(_t(u), _t(hv), _t(wb), _t(vis), work, slab)
    assert rel_err(slab.cpu().numpy(), hg[lo:lo + m]) < TOL
    img, (c0, c1), mx = D.slab_grid_to_image(slab.clone(), n, spans=vs.spans())
    oimg = np.real(orc.ifft(orc.make_grid_hermitian(hg)))
    assert (c0, c1) == (0, n) and rel_err(img.cpu().numpy(), oimg) < TOL and abs(mx - oimg.max()) < TOL * abs(oimg.max())
    vs.gather_slabs(slab, work)
    assert rel_err(work.cpu().numpy(), hg) < TOL
    assert rel_err(vs.degrid(work).cpu().numpy(), orc.convdegrid(gcf, hg, u, hv, wbin=wb)) < TOL


def test_route_kernels_against_torch_reference(orc):
    """skagrid_dev_route_count / _route_pack / _row_hist / _scatter_add against the integer torch restatement
    (distributed.owners_of_rows) on slabs that split footprints, and Plan.update must raise on a bad w-plane index."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import _lib
    from ska_sdp_accelerate_gridding_b200 import device as dv
    from ska_sdp_accelerate_gridding_b200 import distributed as D
    rng = np.random.default_rng(43)
    n, s, q, cnt = 512, 15, 8, 50000
    u, v = rng.uniform(-0.55, 0.55, cnt), rng.uniform(-0.55, 0.55, cnt)
    v[:5] = [np.nan, np.inf, -np.inf, 3e9, -0.5]
    wb = rng.integers(0, 4, cnt)
    vis = _rand_c(rng, cnt)
    bounds = [0, 100, 107, 300, 512]            # a slab shorter than the kernel: footprints reach three owners
    y, _ = orc.frac_coord(n, q, np.where(np.isfinite(v) & (np.abs(v) < 1e9), v, 10.0))
    lo, hi, on = D.owners_of_rows(torch.from_numpy(y - s // 2), s, bounds)
    want = [int((on & (lo <= g) & (hi >= g)).sum()) for g in range(4)]
    counts = dv.route_count(n, q, s, bounds, _t(v)).tolist()
    assert counts == want
    send, sidx = dv.route_pack(n, q, s, bounds, _t(u), _t(v), _t(wb), _t(vis), counts, keep_index=True)
    send, sidx = send.cpu().numpy(), sidx.cpu().numpy()
    off = 0
    for g in range(4):
        sel = np.nonzero((on & (lo <= g) & (hi >= g)).numpy())[0]
        seg_idx = sidx[off:off + counts[g]]
        assert np.array_equal(np.sort(seg_idx), sel)              # the right visibilities, each once
        seg = send[off:off + counts[g]]
        assert np.array_equal(seg[:, 0], u[seg_idx]) and np.array_equal(seg[:, 1], v[seg_idx])
        assert np.array_equal(seg[:, 2].view(np.int64), wb[seg_idx])
        assert np.array_equal(seg[:, 3] + 1j * seg[:, 4], vis[seg_idx])
        off += counts[g]
    send3, _ = dv.route_pack(n, q, s, bounds, _t(u), _t(v), _t(wb), None, counts)
    assert tuple(send3.shape) == (sum(counts), 3)
    hist = torch.zeros(n, dtype=torch.int32, device="cuda")
    dv.row_hist_(n, q, s, _t(v), hist)
    oh = np.bincount(np.clip(y[(on).numpy()], 0, n - 1), minlength=n)
    assert np.array_equal(hist.cpu().numpy(), oh)
    out = torch.zeros(cnt, dtype=torch.complex128, device="cuda")
    back = _rand_c(rng, sum(counts))
    dv.scatter_add_(out, _t(sidx), _t(back))
    ref = np.zeros(cnt, complex)
    np.add.at(ref, sidx, back)
    assert rel_err(out.cpu().numpy(), ref) < 1e-15
    # a plan filled from the packed records grids what the SoA plan grids
    gcf = _rand_c(rng, (4, q, q, s, s))
    ok = np.isfinite(v) & (np.abs(v) < 1e9)
    rec = np.stack([u[ok], v[ok], wb[ok].view(np.float64), vis[ok].real, vis[ok].imag], axis=1)
    plan = dv.Plan.empty(n, n, gcf.shape, rec.shape[0] + 10)
    plan.update_packed(_t(rec))
    g = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
    plan.grid(_t(gcf), g)
    assert rel_err(g.cpu().numpy(), orc.convgrid(gcf, np.zeros((n, n), complex), u[ok], v[ok], vis[ok], wbin=wb[ok])) < TOL
    # update() with an out-of-range w-plane index must raise like the constructor does
    bad = wb[ok].copy()
    bad[7] = 4
    p2 = dv.Plan(n, n, gcf.shape, _t(u[ok]), _t(v[ok]), _t(wb[ok]), _t(vis[ok]))
    with pytest.raises(_lib.SkagridError):
        p2.update(_t(u[ok]), _t(v[ok]), _t(bad), _t(vis[ok]))
    p2.update(_t(u[ok]), _t(v[ok]), _t(bad), _t(vis[ok]), check=False)   # deferred: the word stays set until somebody looks
    with pytest.raises(_lib.SkagridError):
        p2.check()
    p2.update(_t(u[ok]), _t(v[ok]), _t(wb[ok]), _t(vis[ok]))


def test_plan_set_vis_reuses_the_sort(orc):
    """Plan.set_vis: new visibility values at the same coordinates grid exactly like a freshly built plan -- also for a plan that
    was built without visibilities (degrid-only) and for a slab plan that drops part of the batch."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    rng = np.random.default_rng(12)
    n, s, q, nw, cnt = 256, 15, 8, 3, 30000
    gcf = _rand_c(rng, (nw, q, q, s, s))
    u, v = rng.uniform(-0.52, 0.52, cnt), rng.uniform(-0.52, 0.52, cnt)
    wb = rng.integers(0, nw, cnt)
    va, vb = _rand_c(rng, cnt), _rand_c(rng, cnt)
    for rows, first in (((0, n), va), ((64, 200), None)):
        plan = dv.Plan(n, n, gcf.shape, _t(u), _t(v), _t(wb), None if first is None else _t(first), rows=rows)
        plan.set_vis(_t(vb))
        g = torch.zeros((rows[1] - rows[0], n), dtype=torch.complex128, device="cuda")
        plan.grid(_t(gcf), g)
        og = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vb, wbin=wb)
        assert rel_err(g.cpu().numpy(), og[rows[0]:rows[1]]) < TOL
        with pytest.raises(ValueError):
            plan.set_vis(_t(vb[:-1]))
        # the plan's own order: permute once, refresh sequentially
        order = plan.order()
        assert order.numel() == plan.stats()["kept"] and len(set(order.tolist())) == order.numel()
        plan.set_vis(_t(va)[order.long()], in_plan_order=True)
        g.zero_()
        plan.grid(_t(gcf), g)
        oga = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, va, wbin=wb)
        assert rel_err(g.cpu().numpy(), oga[rows[0]:rows[1]]) < TOL
        # degridding in plan order: result r belongs to record r
        model = _rand_c(rng, (n, n))
        clipped = np.zeros_like(model)
        clipped[rows[0]:rows[1]] = model[rows[0]:rows[1]]
        od = orc.convdegrid(gcf, clipped, u, v, wbin=wb)
        d_caller = plan.degrid(_t(gcf), _t(model[rows[0]:rows[1]].copy())).cpu().numpy()
        assert rel_err(d_caller, od) < TOL
        d_plan = plan.degrid(_t(gcf), _t(model[rows[0]:rows[1]].copy()), plan_order=True).cpu().numpy()
        oi = order.cpu().numpy()
        assert rel_err(d_plan[:oi.size], od[oi]) < TOL
        plan.close()


def test_peer_memory_kernels_single_device():
    """csrc/ipc.cu on ONE device: the exchange kernels do not care whether a pointer is local or an opened peer buffer, so local
    buffers stand in for the peers (tests/mgpu_check.py runs the real thing on 2-8 GPUs).  The barrier protocol is exercised by
    two streams playing rank 0 and rank 1 against each other."""
    import ctypes as C
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    from ska_sdp_accelerate_gridding_b200.peer import PeerBuffer, PeerGroup
    pg = PeerGroup()                       # world 1: allocation, export and views work without torch.distributed
    ctx, lib, h = pg.ctx, pg.ctx.lib, pg.ctx.h
    assert pg.world == 1 and pg.flags.ptrs == [pg.flags.local]
    rng = np.random.default_rng(9)
    n = 50000
    bufs = [PeerBuffer(pg, n * 16) for _ in range(4)]
    vals = [_rand_c(rng, n) for _ in range(4)]
    for b, v in zip(bufs, vals):
        b.tensor(torch.complex128, (n,)).copy_(_t(v))
    with pytest.raises(ValueError):
        bufs[0].tensor(torch.complex128, (n + 1,))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    peers = (C.c_void_p * 3)(*[C.c_void_p(b.local) for b in bufs[1:]])
    # own += sum of the three "peers"; then the fused form: the sum is stored back into every one of them
    ctx.check(lib.skagrid_dev_peer_sum(h, 3, peers, C.c_void_p(bufs[0].local), n, 0, st))
    want = vals[0] + vals[1] + vals[2] + vals[3]
    assert rel_err(bufs[0].tensor(torch.complex128, (n,)).cpu().numpy(), want) < 1e-15
    assert np.array_equal(bufs[1].tensor(torch.complex128, (n,)).cpu().numpy(), vals[1])
    bufs[0].tensor(torch.complex128, (n,)).copy_(_t(vals[0]))
    ctx.check(lib.skagrid_dev_peer_sum(h, 3, peers, C.c_void_p(bufs[0].local), n, 1, st))
    for b in bufs:
        assert rel_err(b.tensor(torch.complex128, (n,)).cpu().numpy(), want) < 1e-15
    # gather of unequal, 8-byte aligned segments (routed records are 24 or 40 bytes long) and the strided form
    src = _t(rng.standard_normal(4096 * 5))
    dst = torch.zeros(4096 * 5, dtype=torch.float64, device="cuda")
    segs = [(0, 0, 40 * 100), (40 * 100, 40 * 300, 40 * 7), (40 * 107 + 8, 40 * 500 + 8, 8 * 999)]
    pg.gather([(dst.data_ptr() + d, src.data_ptr() + s_, nb) for d, s_, nb in segs])
    ref = np.zeros(4096 * 5)
    hs = src.cpu().numpy()
    for d, s_, nb in segs:
        ref[d // 8:(d + nb) // 8] = hs[s_ // 8:(s_ + nb) // 8]
    assert np.array_equal(dst.cpu().numpy(), ref)
    with pytest.raises(Exception):
        pg.gather([(dst.data_ptr() + 4, src.data_ptr(), 64)])       # misaligned
    # equal 16-byte aligned segments take the 16-byte kernel, also with a capped grid and from a side stream
    dst2 = torch.zeros(3 * 4096, dtype=torch.float64, device="cuda")
    src2 = _t(rng.standard_normal(3 * 4096))
    hnd = pg.gather_async([(dst2.data_ptr() + k * 4096 * 8, src2.data_ptr() + (2 - k) * 4096 * 8, 4096 * 8) for k in range(3)], max_blocks=2)
    hnd.wait()
    hs2 = src2.cpu().numpy()
    assert np.array_equal(dst2.cpu().numpy(), np.concatenate([hs2[2 * 4096:], hs2[4096:2 * 4096], hs2[:4096]]))
    rows, wid, cw = 37, 64, 16
    a = _t(_rand_c(rng, (2, rows, wid)))
    cols = torch.zeros((2 * rows, cw), dtype=torch.complex128, device="cuda")
    c0 = 32
    pg.gather2d([(cols.data_ptr() + k * rows * cw * 16, cw * 16, a[k].data_ptr() + c0 * 16, wid * 16, cw * 16, rows - k) for k in range(2)])
    hc = cols.cpu().numpy()
    ha = a.cpu().numpy()
    assert np.array_equal(hc[:rows], ha[0][:, c0:c0 + cw]) and np.array_equal(hc[rows:2 * rows - 1], ha[1][:rows - 1, c0:c0 + cw])
    assert np.all(hc[2 * rows - 1] == 0)
    cols.zero_()
    pg.pull([(cols.data_ptr(), cw * 16, a[0].data_ptr() + c0 * 16, wid * 16, cw * 16, rows)])      # copy-engine form of the same
    assert np.array_equal(cols.cpu().numpy()[:rows], ha[0][:, c0:c0 + cw])
    # the barrier: two streams are rank 0 and rank 1 of a 2-rank barrier over two flag buffers; three epochs
    f = [PeerBuffer(pg, 4096) for _ in range(2)]
    fl = (C.c_void_p * 2)(C.c_void_p(f[0].local), C.c_void_p(f[1].local))
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    mark = torch.zeros(2, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for epoch in (1, 2, 3):
        with torch.cuda.stream(s0):
            torch.cuda._sleep(20_000_000)                     # rank 0 arrives late ...
            mark[0] = epoch
            ctx.check(lib.skagrid_dev_peer_barrier(h, 2, 0, fl, epoch, C.c_void_p(s0.cuda_stream)))
        with torch.cuda.stream(s1):
            ctx.check(lib.skagrid_dev_peer_barrier(h, 2, 1, fl, epoch, C.c_void_p(s1.cuda_stream)))
            seen = mark[0].clone()                            # ... and rank 1 must not get past the barrier before it has
        s1.synchronize()
        assert float(seen.item()) == epoch
    torch.cuda.synchronize()
    dv.check_errors(ctx)                                      # no barrier timed out
    for b in bufs + f:
        b.close()
    pg.close()


def test_two_gpu_torchrun_if_available():
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29511", os.path.join(root, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_context_resident_grid_chain(G, orc):
    """grid == NULL chains conv_imaging2 -> grid_to_image / convdegrid2 on the grid left on the device."""
    import ctypes as C
    from ska_sdp_accelerate_gridding_b200 import _lib
    from ska_sdp_accelerate_gridding_b200.context import get_context
    ctx = get_context()
    rng = np.random.default_rng(61)
    theta, lam, nw, q, s, cnt = 0.02, 6400, 3, 4, 15, 5000   # N = 128
    n = 128
    gcf = _rand_c(rng, (nw, q, q, s, s))
    u, v = rng.uniform(-0.45, 0.45, cnt) * lam, rng.uniform(-0.45, 0.45, cnt) * lam
    wb = rng.integers(0, nw, cnt)
    vis = _rand_c(rng, cnt)
    p = lambda a: a.ctypes.data
    ctx.check(ctx.lib.skagrid_conv_imaging2(ctx.h, nw, q, s, s, p(gcf), theta, lam, cnt, p(u), p(v), p(u), p(wb), p(vis), None))
    og = orc.convgrid(gcf, np.zeros((n, n), complex), u / lam, v / lam, vis, wbin=wb)
    pu, pv = u / lam, v / lam
    out = np.empty(cnt, complex)
    ctx.check(ctx.lib.skagrid_convdegrid2(ctx.h, nw, q, s, s, p(gcf), n, n, None, cnt, p(pu), p(pv), p(wb), p(out)))
    assert rel_err(out, orc.convdegrid(gcf, og, pu, pv, wbin=wb)) < TOL
    mx = np.zeros(1)
    img = np.empty((n, n))
    ctx.check(ctx.lib.skagrid_grid_to_image(ctx.h, n, None, p(img), p(mx)))
    oimg = np.real(orc.ifft(orc.make_grid_hermitian(og)))
    assert rel_err(img, oimg) < TOL
    # a call that overwrites the scratch invalidates the resident grid; a wrong shape is refused
    G.make_grid_hermitian(np.zeros((8, 8), complex))
    assert ctx.lib.skagrid_grid_to_image(ctx.h, n, None, None, p(mx)) == -1
    with pytest.raises(_lib.SkagridError):
        ctx.check(ctx.lib.skagrid_convdegrid2(ctx.h, nw, q, s, s, p(gcf), 64, 64, None, cnt, p(pu), p(pv), p(wb), p(out)))


def test_c_abi_from_plain_c(tmp_path):
    """The boundary is a C ABI: compile tests/c_abi_smoke.c with gcc against include/skagrid.h and run it."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ska_sdp_accelerate_gridding_b200")
    exe = str(tmp_path / "c_abi_smoke")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "c_abi_smoke.c"),
                        "-L" + pkg, "-lskagrid", "-lm", "-Wl,-rpath," + pkg, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke ok" in r.stdout


def test_device_resident_presteps(orc):
    """uvw_lambda / doweight / mirror / findClosest / div3 on device arrays are bit-identical to the oracle."""
    from ska_sdp_accelerate_gridding_b200 import _lib
    from ska_sdp_accelerate_gridding_b200 import device as dv
    rng = np.random.default_rng(71)
    cnt, theta, lam, freq = 40000, 0.01, 30000, 1.1e8
    sc = 299792458.0 / freq
    u, v, w = (rng.uniform(-14000, 14000, cnt) * sc for _ in range(3))
    vis = _rand_c(rng, cnt)
    wbins = np.sort(rng.uniform(-15000, 15000, 33))
    du, dv_, dw, dvis = _t(u), _t(v), _t(w), _t(vis)
    dv.uvw_lambda_(freq, du, dv_, dw)
    ou, ov, ow = orc.uvw_lambda(freq, u, v, w)
    assert np.array_equal(du.cpu().numpy(), ou) and np.array_equal(dw.cpu().numpy(), ow)
    wt = _t(np.ones(cnt, complex))
    dv.doweight_(theta, lam, du, dv_, wt)
    assert np.array_equal(wt.cpu().numpy(), orc.doweight(theta, lam, ou, ov, np.ones(cnt, complex)))
    dv.mirror_uvw_(du, dv_, dw, dvis)
    mu, mv, mw, mvis = orc.mirror_uvw(ou, ov, ow, vis)
    assert np.array_equal(dv_.cpu().numpy(), mv) and np.array_equal(dvis.cpu().numpy(), mvis)
    assert np.array_equal(dv.find_closest(_t(wbins), dw).cpu().numpy(), orc.find_closest(wbins, mw))
    dv.div3_(lam, du, dv_, dw)
    pu, pv, pw = orc.div3(float(lam), mu, mv, mw)
    assert np.array_equal(du.cpu().numpy(), pu) and np.array_equal(dv_.cpu().numpy(), pv)
    with pytest.raises(_lib.SkagridError):
        dv.doweight_(theta, lam, _t(u * 1000), _t(v), _t(vis))


@pytest.mark.parametrize("tile,cellsort", [(16, 0), (16, 1), (32, 0), (32, 1)])
def test_plan_layout_choices(orc, monkeypatch, tile, cellsort):
    """The per-plan layout choices (uv tile 16/32, cell- or micro-tile-granular buckets) are performance knobs only:
    every combination reproduces the oracle (they select different gridder residency and degridder kernels)."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    monkeypatch.setenv("SKAGRID_TILE", str(tile))
    monkeypatch.setenv("SKAGRID_CELLSORT", str(cellsort))
    rng = np.random.default_rng(100 + tile + cellsort)
    n, s, q, nw, cnt = 160, 15, 8, 3, 30000
    gcf = _rand_c(rng, (nw, q, q, s, s))
    u = np.clip(rng.normal(0, 0.08, cnt), -0.52, 0.52)
    v = np.abs(np.clip(rng.normal(0, 0.08, cnt), -0.52, 0.52))
    wb = rng.integers(0, nw, cnt)
    vis = _rand_c(rng, cnt)
    og = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vis, wbin=wb, parallel=True)
    plan = dv.Plan(n, n, gcf.shape, _t(u), _t(v), _t(wb), _t(vis))
    grid = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
    plan.grid(_t(gcf), grid)
    assert rel_err(grid.cpu().numpy(), og) < TOL
    d = plan.degrid(_t(gcf), _t(og)).cpu().numpy()
    assert rel_err(d, orc.convdegrid(gcf, og, u, v, wbin=wb, parallel=True)) < TOL
    plan.close()


def test_non_square_grid(G, orc):
    """height != width: x from u with the width, y from v with the height (src/Gridding.hs:142-151)."""
    rng = np.random.default_rng(5150)
    h, w, s, q, nw, cnt = 96, 130, 15, 4, 2, 8000
    gcf = _rand_c(rng, (nw, q, q, s, s))
    u, v = rng.uniform(-0.55, 0.55, cnt), rng.uniform(-0.55, 0.55, cnt)
    wb = rng.integers(0, nw, cnt)
    vis = _rand_c(rng, cnt)
    start = _rand_c(rng, (h, w))
    g = G.convgrid2(gcf, start, (u, v), wb, vis)
    og = orc.convgrid(gcf, start, u, v, vis, wbin=wb)
    assert g.shape == (h, w) and rel_err(g, og) < TOL
    assert rel_err(G.convdegrid2(gcf, og, (u, v), wb), orc.convdegrid(gcf, og, u, v, wbin=wb)) < TOL


def test_nan_inf_and_far_off_grid_visibilities_are_dropped(G, orc):
    """Non-finite or absurd coordinates never touch the grid (the reference would index out of bounds); finite
    visibilities in the same batch are unaffected, and their degridded values are exact while the dropped ones read 0."""
    rng = np.random.default_rng(404)
    n, s, q, nw, cnt = 96, 9, 4, 2, 2000
    gcf = _rand_c(rng, (nw, q, q, s, s))
    u, v = rng.uniform(-0.5, 0.5, cnt), rng.uniform(-0.5, 0.5, cnt)
    wb = rng.integers(0, nw, cnt)
    vis = _rand_c(rng, cnt)
    bad = np.array([3, 77, 500, 1999])
    u2, v2 = u.copy(), v.copy()
    u2[bad[0]] = np.nan; v2[bad[1]] = np.inf; u2[bad[2]] = -np.inf; v2[bad[3]] = 1e30
    keep = np.ones(cnt, bool); keep[bad] = False
    g = G.convgrid2(gcf, np.zeros((n, n), complex), (u2, v2), wb, vis)
    og = orc.convgrid(gcf, np.zeros((n, n), complex), u[keep], v[keep], vis[keep], wbin=wb[keep])
    assert rel_err(g, og) < TOL
    d = G.convdegrid2(gcf, og, (u2, v2), wb)
    od = orc.convdegrid(gcf, og, u[keep], v[keep], wbin=wb[keep])
    assert rel_err(d[keep], od) < TOL and not d[bad].any()


def test_cpp_host_mirror_replays_the_reference_tests(tmp_path):
    """include/skagrid.hpp (C++ mirror of Gridding.hs / ImageDataset.hs): test/SmallTest.hs and old/BrokenNumbers.hs."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ska_sdp_accelerate_gridding_b200")
    exe = str(tmp_path / "cpp_smalltest")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp_smalltest.cpp"),
                        "-L" + pkg, "-lskagrid", "-Wl,-rpath," + pkg, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp_smalltest ok" in r.stdout


# ------------------------------------------------------------------------------------------------ single-process multi-GPU ABI
def _mg_devices(k):
    import torch
    nd = torch.cuda.device_count()
    return [i % nd for i in range(k)]  # distinct devices when the box has them, else several contexts on one device


@pytest.mark.parametrize("nctx", [1, 2, 3])
@pytest.mark.parametrize("mode", ["vis", "tile"])
def test_mgpu_abi_parity(orc, nctx, mode):
    """skagrid_conv(de)grid2_mgpu_{vis,tile}: one host thread, nctx contexts; same grid / visibilities as the oracle."""
    from ska_sdp_accelerate_gridding_b200.multi_device import MultiDevice
    rng = np.random.default_rng(77 + nctx)
    n, s, qpx, nw, count = 320, 15, 4, 3, 30011
    gcf = _rand_c(rng, (nw, qpx, qpx, s, s))
    # core-dominated rows (so the balanced slabs differ from equal ones) plus footprints hanging over every grid edge
    u = rng.uniform(-0.52, 0.52, count)
    v = np.where(rng.random(count) < 0.7, rng.normal(0.17, 0.05, count), rng.uniform(-0.52, 0.52, count))
    wb = rng.integers(0, nw, count)
    vis = _rand_c(rng, count)
    start = _rand_c(rng, (n, n))
    md = MultiDevice(_mg_devices(nctx))
    try:
        g = md.convgrid2(gcf, start, (u, v), wb, vis, mode=mode)
        og = orc.convgrid(gcf, start, u, v, vis, wbin=wb)
        assert rel_err(g, og) < TOL
        d = md.convdegrid2(gcf, og, (u, v), wb, mode=mode)
        od = orc.convdegrid(gcf, og, u, v, wbin=wb)
        assert rel_err(d, od) < TOL
        if mode == "tile":
            b = md.bounds
            assert b[0] == 0 and b[-1] == n and all(b[i] < b[i + 1] for i in range(nctx))
            if nctx > 1:
                assert b != [n * i // nctx for i in range(nctx + 1)]  # quantiles of the row histogram, not equal heights
        # fewer visibilities than contexts, none at all, and the 4-D table form (nw = 1, no wbin)
        g2 = md.convgrid2(gcf, start, (u[:1], v[:1]), wb[:1], vis[:1], mode=mode)
        assert rel_err(g2, orc.convgrid(gcf, start, u[:1], v[:1], vis[:1], wbin=wb[:1])) < TOL
        g0 = md.convgrid2(gcf, start, (u[:0], v[:0]), wb[:0], vis[:0], mode=mode)
        assert np.array_equal(g0, start)
        assert md.convdegrid2(gcf, og, (u[:0], v[:0]), wb[:0], mode=mode).size == 0
        g4 = md.convgrid2(gcf[0], start, (u, v), None, vis, mode=mode)
        assert rel_err(g4, orc.convgrid(gcf[:1], start, u, v, vis, wbin=np.zeros(count, np.int64))) < TOL
    finally:
        md.close()


def test_mgpu_abi_resident_chain_and_errors(G, orc):
    """Visibility-sharded gridding leaves the summed grid on every context: degridding and grid_to_image then take
    grid == NULL.  Range errors of any context surface through the first one."""
    import ctypes as C
    from ska_sdp_accelerate_gridding_b200 import _lib
    from ska_sdp_accelerate_gridding_b200.multi_device import MultiDevice
    rng = np.random.default_rng(5)
    n, s, qpx, nw, count = 256, 9, 2, 2, 9000
    gcf = _rand_c(rng, (nw, qpx, qpx, s, s))
    u, v = rng.uniform(-0.45, 0.45, count), rng.uniform(-0.45, 0.45, count)
    wb = rng.integers(0, nw, count)
    vis = _rand_c(rng, count)
    md = MultiDevice(_mg_devices(2))
    try:
        with pytest.raises(_lib.SkagridError) as e:  # nothing resident yet
            md._resident_shape = (n, n)
            md.convdegrid2(gcf, None, (u, v), wb)
        assert e.value.code == -1
        og = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vis, wbin=wb)
        oimg = np.real(orc.ifft(orc.make_grid_hermitian(og)))
        md.conv_grid_resident(gcf, (n, n), (u, v), wb, vis)
        d = md.convdegrid2(gcf, None, (u, v), wb)
        assert rel_err(d, orc.convdegrid(gcf, og, u, v, wbin=wb)) < TOL
        d = md.convdegrid2(gcf, None, None, None, count=count)  # ... and at the coordinates each context still holds
        assert rel_err(d, orc.convdegrid(gcf, og, u, v, wbin=wb)) < TOL
        for c in md.ctxs:  # every context holds the full sum
            img, mx = G.grid_to_image(None, ctx=c, n=n)
            assert rel_err(img, oimg) < TOL and abs(mx - oimg.max()) <= TOL * abs(oimg.max())
        assert md.last_device_ms > 0.0
        bad = wb.copy()
        bad[-1] = nw  # lands in the last context's share
        for mode in ("vis", "tile"):
            with pytest.raises(_lib.SkagridError) as e:
                md.convgrid2(gcf, np.zeros((n, n), complex), (u, v), bad, vis, mode=mode)
            assert e.value.code == -5
        # argument errors: the same context twice, no contexts
        twice = (C.c_void_p * 2)(md.ctxs[0].h, md.ctxs[0].h)
        g = np.zeros((n, n), complex)
        rc = md.lib.skagrid_convgrid2_mgpu_vis(twice, 2, nw, qpx, s, s, gcf.ctypes.data, n, n, g.ctypes.data, 0, None, None, None, None)
        assert rc == -1 and b"twice" in md.lib.skagrid_last_error(md.ctxs[0].h)
        assert md.lib.skagrid_convgrid2_mgpu_vis(twice, 0, nw, qpx, s, s, gcf.ctypes.data, n, n, g.ctypes.data, 0, None, None, None, None) == -1
        # and the contexts stay usable
        assert rel_err(md.convgrid2(gcf, np.zeros((n, n), complex), (u, v), wb, vis), og) < TOL
    finally:
        md.close()


def test_resident_coordinates(G, orc):
    """u == v == wbin == NULL: the next table call works at the coordinates the previous one uploaded."""
    import ctypes as C
    from ska_sdp_accelerate_gridding_b200 import _lib
    from ska_sdp_accelerate_gridding_b200.context import Context
    rng = np.random.default_rng(123)
    n, s, qpx, nw, count = 192, 7, 4, 3, 20000
    gcf = _rand_c(rng, (nw, qpx, qpx, s, s))
    u, v = rng.uniform(-0.5, 0.5, count), rng.uniform(-0.5, 0.5, count)
    wb = rng.integers(0, nw, count)
    vis, vis2 = _rand_c(rng, count), _rand_c(rng, count)
    ctx = Context(0)
    try:
        with pytest.raises(_lib.SkagridError) as e:  # nothing resident on a fresh context
            G.convdegrid2(gcf, np.zeros((n, n), complex), None, None, ctx=ctx, count=count)
        assert e.value.code == -1
        g = G.convgrid2(gcf, np.zeros((n, n), complex), (u, v), wb, vis, ctx=ctx)
        og = orc.convgrid(gcf, np.zeros((n, n), complex), u, v, vis, wbin=wb)
        assert rel_err(g, og) < TOL
        d = G.convdegrid2(gcf, og, None, None, ctx=ctx, count=count)          # same coordinates, nothing re-uploaded
        assert rel_err(d, orc.convdegrid(gcf, og, u, v, wbin=wb)) < TOL
        g2 = G.convgrid2(gcf, og, None, None, vis2, ctx=ctx)                   # new values at the resident coordinates
        assert rel_err(g2, orc.convgrid(gcf, og, u, v, vis2, wbin=wb)) < TOL
        d0 = G.convdegrid(gcf[0], og, None, ctx=ctx, count=count)              # 4-D table: the resident w-plane indices are not used
        assert rel_err(d0, orc.convdegrid(gcf[:1], og, u, v, wbin=np.zeros(count, np.int64))) < TOL
        with pytest.raises(_lib.SkagridError) as e:                            # another count is not what is resident
            G.convdegrid2(gcf, og, None, None, ctx=ctx, count=count - 1)
        assert e.value.code == -1
        # conv_imaging2 keeps the coordinates after the division by lam
        theta, lam = 0.01, n * 100
        G.conv_imaging2(gcf, theta, lam, (u * lam, v * lam, np.zeros(count)), wb, vis, ctx=ctx)
        d = G.convdegrid2(gcf, og, None, None, ctx=ctx, count=count)
        pu, pv = (u * lam) / lam, (v * lam) / lam
        assert rel_err(d, orc.convdegrid(gcf, og, pu, pv, wbin=wb)) < TOL
        # a 4-D call leaves no w-plane indices behind: the 5-D form must refuse them
        G.convgrid(gcf[0], np.zeros((n, n), complex), (u, v), vis, ctx=ctx)
        with pytest.raises(_lib.SkagridError):
            G.convdegrid2(gcf, og, None, None, ctx=ctx, count=count)
    finally:
        ctx.close()


def test_command_line_on_a_standin_dataset(G, orc, tmp_path, capsys):
    """app/Main.hs through `python -m ska_sdp_accelerate_gridding_b200`: loaders + aw_gridding on files."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from make_standin_dataset import make
    from ska_sdp_accelerate_gridding_b200 import __main__ as cli
    from ska_sdp_accelerate_gridding_b200 import image_dataset as D
    d = make(str(tmp_path / "data"), vis=3000, nw=6, nant=5)
    out = str(tmp_path / "result.npz")
    assert cli.main(["-gpu", "-n", "2000", "-i", d, "-o", out]) == 0
    printed = float(capsys.readouterr().out.strip().splitlines()[-1])
    dat = np.load(os.path.join(d, "SKA1_Low_quick.npz"))
    wk, wbins = D.getWKernels(np.load(os.path.join(d, "SKA1_Low_wkern2.npz")), 0.008)
    ak = D.getAKernels(np.load(os.path.join(d, "SKA1_Low_akern3.npz")), 0.008, float(dat["/vis/time"][0]), float(dat["/vis/frequency"][0]))
    assert wk.shape == (6, 8, 8, 15, 15) and ak.shape == (5, 15, 15) and abs(np.abs(ak[0]).sum() - 1.0) < 1e-12  # closest time/frequency picked
    uvw, vis = dat["/vis/uvw"], dat["/vis/vis"].reshape(-1)
    n = 2000
    oimg, omx, _ = orc.aw_gridding(0.008, 300000, wk, wbins, ak, uvw[:n, 0], uvw[:n, 1], uvw[:n, 2], dat["/vis/antenna1"][:n],
                                   dat["/vis/antenna2"][:n], float(dat["/vis/frequency"][0]), vis[:n])
    assert abs(printed - omx) <= TOL * abs(omx)
    assert rel_err(np.load(out)["/img"], oimg) < TOL
    assert cli.main(["-old", "-i", d]) == 0  # default: the first visibility only (app/Main.hs:26)
    one = float(capsys.readouterr().out.strip().splitlines()[-1])
    _, omx1, _ = orc.aw_gridding(0.008, 300000, wk, wbins, ak, uvw[:1, 0], uvw[:1, 1], uvw[:1, 2], dat["/vis/antenna1"][:1],
                                 dat["/vis/antenna2"][:1], float(dat["/vis/frequency"][0]), vis[:1])
    assert abs(one - omx1) <= TOL * abs(omx1)


def test_sharded_doweight_two_phase(orc):
    """doweight split into count / (all-reduce) / apply equals the one-call form; with the visibilities cut into two
    shares whose counts are added by hand it still equals the oracle on the whole set (bit-exact)."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    from ska_sdp_accelerate_gridding_b200 import distributed as D
    rng = np.random.default_rng(9)
    theta, lam, cnt = 0.01, 20000, 50000   # n = 200: many visibilities per cell
    u, v = rng.uniform(-0.45, 0.45, cnt) * lam, rng.uniform(-0.45, 0.45, cnt) * lam
    vis = _rand_c(rng, cnt)
    ow = orc.doweight(theta, lam, u, v, vis)
    one = _t(vis.copy())
    D.doweight_sharded_(theta, lam, _t(u), _t(v), one)  # no process group: a single share
    assert np.array_equal(one.cpu().numpy(), ow)
    h = cnt // 3
    hist = torch.zeros((200, 200), dtype=torch.int32, device="cuda")
    shares = [(_t(u[:h]), _t(v[:h]), _t(vis[:h].copy())), (_t(u[h:]), _t(v[h:]), _t(vis[h:].copy()))]
    for su, sv, _ in shares:
        dv.weight_count_(theta, lam, su, sv, hist)     # what the all-reduce produces: the sum of the shares' counts
    assert int(hist.sum().item()) == cnt
    for su, sv, svis in shares:
        dv.weight_apply_(theta, lam, su, sv, hist, svis)
    assert np.array_equal(torch.cat([s[2] for s in shares]).cpu().numpy(), ow)


@pytest.mark.parametrize("n", [64, 250, 1024])
def test_slab_distributed_grid_to_image_stages(G, orc, n):
    """The two device stages of the slab-distributed grid -> image (row transforms, transpose, column transforms): one slab
    through distributed.slab_grid_to_image, and three uneven row slabs / two column slabs transposed by hand."""
    import torch
    from ska_sdp_accelerate_gridding_b200 import device as dv
    from ska_sdp_accelerate_gridding_b200 import distributed as D
    rng = np.random.default_rng(n)
    g = _rand_c(rng, (n, n))
    oimg = np.real(orc.ifft(orc.make_grid_hermitian(g)))
    img, (c0, c1), mx = D.slab_grid_to_image(_t(g.copy()), [0, n])
    assert (c0, c1) == (0, n)
    assert rel_err(img.cpu().numpy(), oimg) < TOL and abs(mx - oimg.max()) <= TOL * abs(oimg.max())
    bounds = [0, n // 5, n // 5 + 1, n]   # a one-row slab in the middle
    slabs = []
    for r0, r1 in zip(bounds, bounds[1:]):
        s = _t(g[r0:r1].copy())
        dv.slab_fft_rows_(n, r0, s)
        slabs.append(s)
    rows = torch.cat(slabs)               # what the all-to-all assembles, per column block
    mxs = []
    for c0, c1 in ((0, n // 3), (n // 3, n)):
        cols = rows[:, c0:c1].contiguous()
        im, m = dv.slab_fft_cols_(n, c0, cols)
        assert rel_err(im.cpu().numpy(), oimg[:, c0:c1]) < TOL
        mxs.append(float(m.item()))
    assert abs(max(mxs) - oimg.max()) <= TOL * abs(oimg.max())
    with pytest.raises(Exception):
        dv.slab_fft_rows_(n + 1, 0, _t(_rand_c(rng, (2, n + 1))))  # odd sides are not supported by the slab path
