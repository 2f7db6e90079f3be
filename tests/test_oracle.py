"""CPU tests of the oracle (oracle/): pinned against the reference's only stored numbers
(old/BrokenNumbers.hs:85-91), its in-source test inputs (test/SmallTest.hs:51-73) and internal
consistency (direct vs FFT convolve2d, adjointness, C vs pure-Python transliterations)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_broken_numbers_golden():
    j = json.load(open(os.path.join(GOLD, "broken_numbers.json")))
    k = np.arange(10)
    x, y = (2 * k) % 5, (3 * k + 1) % 5
    val = (k + 5) + 1j
    g = np.zeros((5, 5), complex)
    for _ in range(j["applications"]):
        g = orc.scatter_add(g, x, y, val)
    exp = np.array(j["expected_re"]) + 1j * np.array(j["expected_im"])
    assert np.array_equal(g, exp)


def _smalltest():
    from tests.golden.make_golden import smalltest_inputs
    return smalltest_inputs()


def test_smalltest_known_answers():
    """Derived known answers of SURVEY.md 8c for test/SmallTest.hs:51-76 (convgrid3 on a 10x10 grid)."""
    wk, ak, uvw, idx, vis = _smalltest()
    x, xf, y, yf = orc.frac_coords(10, 10, 1, uvw[:, 0], uvw[:, 1])
    assert (x.tolist(), xf.tolist(), y.tolist(), yf.tolist()) == ([6, 4], [0, 0], [7, 9], [0, 0])
    g = orc.convgrid_aw(wk, ak, np.zeros((10, 10), complex), uvw[:, 0], uvw[:, 1], idx[:, 0], idx[:, 1], idx[:, 2], vis)
    assert abs(g.sum() - (2258.959 + 1709.566j)) < 2e-3
    iy, ix = np.unravel_index(np.argmax(np.abs(g)), g.shape)
    assert (iy, ix) == (9, 6) and abs(abs(g[9, 6]) - 46.27454822566) < 1e-9
    assert abs(g[7, 6] - (35.417885 + 25.898745j)) < 1e-5
    assert abs(g[9, 4] - (38.86883 + 23.95676j)) < 1e-4
    assert abs(g[0, 0] - (4.558 + 5.9148j)) < 1e-3
    stored = np.load(os.path.join(GOLD, "smalltest_oracle.npz"))
    assert np.allclose(g, stored["grid"], rtol=0, atol=1e-12)


def test_smalltest_three_formulations_agree():
    """convgrid3 (per-visibility scatter), convgrid4 (materialise vis*conj(AW), one scatter) and the table
    gridder fed with the explicit per-visibility kernels must give the same grid (test/SmallTest.hs:75-158)."""
    wk, ak, uvw, idx, vis = _smalltest()
    u, v = uvw[:, 0], uvw[:, 1]
    g3 = orc.convgrid_aw(wk, ak, np.zeros((10, 10), complex), u, v, idx[:, 0], idx[:, 1], idx[:, 2], vis)
    x, xf, y, yf = orc.frac_coords(10, 10, 1, u, v)
    g4 = np.zeros((10, 10), complex)
    xs, ys, vals = [], [], []
    for k in range(2):
        aw = orc.aw_kernel(wk[idx[k, 0]], yf[k], xf[k], ak[idx[k, 1]], ak[idx[k, 2]])
        kv = vis[k] * np.conj(aw)
        for j in range(15):
            for i in range(15):
                gx, gy = x[k] + i - 7, y[k] + j - 7
                if 0 <= gx < 10 and 0 <= gy < 10:
                    xs.append(gx); ys.append(gy); vals.append(kv[j, i])
    g4 = orc.scatter_add(g4, np.array(xs), np.array(ys), np.array(vals))
    assert np.allclose(g3, g4, rtol=0, atol=1e-12)
    # literal FFT route of convolve2d gives the same AW kernel
    aw_fft = orc.convolve2d_fft(orc.convolve2d_fft(ak[0], ak[1]), wk[0, 0, 0])
    aw_dir = orc.aw_kernel(wk[0], 0, 0, ak[0], ak[1])
    assert np.allclose(aw_fft, aw_dir, rtol=0, atol=1e-11)


@pytest.mark.parametrize("n", [3, 6, 15, 16])
def test_convolve2d_direct_equals_fft_route(n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    b = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    d = orc.convolve2d(a, b)
    f = orc.convolve2d_fft(a, b)
    assert np.abs(d - f).max() < 1e-11 * max(1.0, np.abs(d).max())
    # Q1: equals the transpose of the plain centre-"same" convolution
    if n % 2 == 1:
        from scipy.signal import convolve2d as sc
        assert np.abs(d - sc(a, b, mode="same").T).max() < 1e-11 * np.abs(d).max()


@pytest.mark.parametrize("n", [5, 7, 15, 17])
def test_row_pair_schedule_of_the_formation_kernel(n):
    """The schedule of conv_pair_kernel (csrc/awkern.cu) restated in numpy: column groups of 4 outputs that drop the taps
    whose operand column lies outside the kernel, two output rows per worker sharing one operand window (the lower row's
    window of step ky is the upper row's of step ky+1), operand rows outside the kernel read as zero.  Must equal
    convolve2d: this pins the index arithmetic the CUDA kernel was written from."""
    rng = np.random.default_rng(100 + n)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    b = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    c, tx_per = n // 2, 4
    aT, bT = a.T.copy(), b.T.copy()       # aT[ky][kx] = a[kx,ky], bT[qy][qx] = b[qx,qy]
    zero = np.zeros(n, complex)
    row = lambda q: bT[q] if 0 <= q < n else zero
    out = np.zeros((n, n), complex)
    macs = 0
    for g in range((n + tx_per - 1) // tx_per):
        tx0 = g * tx_per
        txn = min(tx_per, n - tx0)
        for rp in range((n + 1) // 2):
            acc = np.zeros((2, txn), complex)
            q00 = 2 * rp + c
            hi = row(q00 + 1)
            for ky in range(n):
                lo = row(q00 - ky)
                for kx in range(n):
                    for t in range(txn):
                        qx = tx0 + t + c - kx
                        if 0 <= qx < n:  # resolved at compile time in the kernel
                            acc[0, t] += aT[ky, kx] * lo[qx]
                            acc[1, t] += aT[ky, kx] * hi[qx]
                            macs += 1
                hi = lo
            for r in range(2):
                if 2 * rp + r < n:
                    out[2 * rp + r, tx0:tx0 + txn] = acc[r]
    ref = orc.convolve2d(a, b)
    assert np.abs(out - ref).max() < 1e-12 * np.abs(ref).max()
    if n == 15:  # 169 of 240 taps per (output row, step): the x half of the zero padding is skipped
        assert macs == 169 * 15 * 8

def test_frac_coord_properties():
    rng = np.random.default_rng(1)
    p = rng.uniform(-0.5, 0.5, 20000)
    for n, q in [(2400, 8), (8192, 8), (10, 1), (33, 4)]:
        fl, fr = orc.frac_coord(n, q, p)
        assert fr.min() >= 0 and fr.max() < q
        x = n // 2 + p * n
        # (fl, fr) is the nearest 1/q sub-cell: |x - (fl + fr/q)| <= 1/(2q) (+ rounding)
        assert np.abs(x - (fl + fr / q)).max() <= 0.5 / q + 1e-9
    # raw mode reproduces the tie quirk Q3: x + 0.5/qpx integral -> frac = -1
    fl, fr = orc.frac_coord(16, 4, np.array([(3 - 0.125) / 16 - 0.5 + 0.5]), normalise=False)
    fl2, fr2 = orc.frac_coord(16, 4, np.array([(3 - 0.125) / 16 - 0.5 + 0.5]), normalise=True)
    assert 0 <= fr2[0] < 4 and fl[0] * 4 + fr[0] == fl2[0] * 4 + fr2[0]


def test_find_closest_c_equals_python_and_clamps():
    rng = np.random.default_rng(2)
    ws = np.sort(rng.uniform(-100, 100, 37))
    w = np.concatenate([rng.uniform(-120, 120, 500), ws, [ws[-1] + 5.0, ws[0] - 5.0]])
    c = orc.find_closest(ws, w)
    py = np.array([orc.find_closest_py(ws, x) for x in w])
    assert np.array_equal(c, py)
    assert c[-2] == len(ws) - 1 and c[-1] == 0
    inside = (w >= ws[0]) & (w <= ws[-1])
    brute = np.abs(w[inside, None] - ws[None, :]).argmin(axis=1)
    assert np.abs(np.abs(w[inside] - ws[c[inside]]) - np.abs(w[inside] - ws[brute])).max() < 1e-12


def test_golden_random_case_and_adjoint():
    z = np.load(os.path.join(GOLD, "random_cases.npz"))
    n = int(z["n"])
    g = orc.convgrid(z["gcf"], np.zeros((n, n), complex), z["u"], z["v"], z["vis"], wbin=z["wbin"])
    assert np.array_equal(g, z["grid"])
    gp = orc.convgrid(z["gcf"], np.zeros((n, n), complex), z["u"], z["v"], z["vis"], wbin=z["wbin"], parallel=True)
    assert np.allclose(g, gp, rtol=0, atol=1e-12)
    d = orc.convdegrid(z["gcf"], g, z["u"], z["v"], wbin=z["wbin"])
    assert np.array_equal(d, z["degrid"])
    # <grid(v), g'> == <v, degrid(g')>
    rng = np.random.default_rng(3)
    g2 = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    lhs = np.vdot(g2, g)
    rhs = np.vdot(orc.convdegrid(z["gcf"], g2, z["u"], z["v"], wbin=z["wbin"]), z["vis"])
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)


def test_hermitian_and_ifft_conventions():
    rng = np.random.default_rng(4)
    for n in (8, 9):
        g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        h = orc.make_grid_hermitian(g)
        if n % 2 == 0:
            assert np.array_equal(h[0, :], g[0, :]) and np.array_equal(h[:, 0], g[:, 0])
            assert np.allclose(h[1:, 1:], g[1:, 1:] + np.conj(g[1:, 1:][::-1, ::-1]))
        else:
            assert np.allclose(h, g + np.conj(g[::-1, ::-1]))
    n = 8
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    m = (-1.0) ** (np.add.outer(np.arange(n), np.arange(n)))
    assert np.allclose(orc.ifft(g), m * np.fft.ifft2(m * g))  # Q5
    assert np.allclose(orc.fft(orc.ifft(g)), g)


def test_w_kernel_shape_and_symmetry():
    k = orc.w_kernel(0.05, 300.0, 32, 9, 4)
    assert k.shape == (4, 4, 9, 9)
    assert np.allclose(k[0, 0], k[0, 0].T)        # far field symmetric in l,m
    k0 = orc.w_kernel(0.05, 0.0, 32, 9, 4)
    assert abs(k0[0, 0, 4, 4] - 1.0) < 1e-12        # w = 0: delta function at the centre tap
    assert np.abs(k0[0, 0]).sum() - 1.0 < 1e-9


def test_mirror_doweight_pipeline():
    rng = np.random.default_rng(5)
    cnt, theta, lam = 300, 0.01, 3000
    u, v, w = (rng.uniform(-1000, 1000, cnt) for _ in range(3))
    vis = rng.standard_normal(cnt) + 1j * rng.standard_normal(cnt)
    u1, v1, w1, vis1 = orc.mirror_uvw(u, v, w, vis)
    assert (v1 >= 0).all() and np.array_equal(vis1[v < 0], np.conj(vis[v < 0]))
    wt = orc.doweight(theta, lam, u, v, np.ones(cnt, complex))
    x, _, y, _ = orc.frac_coords(30, 30, 1, u / lam, v / lam, normalise=False)
    cells = y * 30 + x
    counts = np.bincount(cells, minlength=900)[cells]
    assert np.array_equal(wt, 1.0 / counts + 0j)


def test_shift_convention_is_the_one_written_out_in_ShiftExample():
    """old/ShiftExample.hs:98-106 spells accelerate-fft's shift1D out as a backpermute: out[i] = in[(i + n `quot` 2 + (1 if odd n)) `rem` n].
    The oracle's shift2d / ishift2d (numpy fftshift / ifftshift) must be that index map per axis -- for odd sizes too, where the
    two differ -- and ifft = shift2D . fft2D Inverse . ishift2D (src/Gridding.hs:828-829) must reduce to the (-1)^(x+y) modulation
    the CUDA path uses for the even sizes of the hot path."""
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    for n in (4, 5, 8, 9, 30):
        a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        sh = n // 2 + (n % 2)
        idx = (np.arange(n) + sh) % n
        assert np.array_equal(orc.shift2d(a), a[np.ix_(idx, idx)])
        assert np.array_equal(orc.ishift2d(orc.shift2d(a)), a)
    n = 8
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    yy, xx = np.mgrid[0:n, 0:n]
    m = np.where((yy + xx) % 2 == 1, -1.0, 1.0)
    assert np.abs(orc.ifft(a) - m * np.fft.ifft2(m * a)).max() < 1e-15


def test_kernel_coordinates_options():
    """src/Gridding.hs:620-635 in the oracle: identity options reproduce the plain w-kernel, the matrix acts as
    (x, y) -> (t00 x + t10 y, t01 x + t11 y), the shifts are added last."""
    from oracle import oracle as orc
    l0, m0 = orc.kernel_coordinates(8, 0.1)
    sc, sr = orc.coordinates2(8)
    assert np.array_equal(l0, sc * 0.1) and np.array_equal(m0, sr * 0.1)
    t = np.array([[2.0, 3.0], [5.0, 7.0]])
    l, m = orc.kernel_coordinates(8, 0.1, dl=0.25, dm=-0.5, transmat=t)
    assert np.allclose(l, 2.0 * l0 + 5.0 * m0 + 0.25) and np.allclose(m, 3.0 * l0 + 7.0 * m0 - 0.5)
    assert np.array_equal(orc.w_kernel(0.05, 10.0, 16, 5, 2), orc.w_kernel(0.05, 10.0, 16, 5, 2, dl=0, dm=0, transmat=np.eye(2)))
