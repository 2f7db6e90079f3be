"""Property tests of the CPU oracle (hypothesis): the size-independent identities the GPU suite relies on at full
scale must hold for the oracle itself on arbitrary small problems."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as orc


def _problem(seed, n, s, q, nw, cnt, span=0.5):
    rng = np.random.default_rng(seed)
    gcf = rng.standard_normal((nw, q, q, s, s)) + 1j * rng.standard_normal((nw, q, q, s, s))
    u, v = rng.uniform(-span, span, cnt), rng.uniform(-span, span, cnt)
    wb = rng.integers(0, nw, cnt)
    vis = rng.standard_normal(cnt) + 1j * rng.standard_normal(cnt)
    return gcf, u, v, wb, vis


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10**6), n=st.integers(8, 40), s=st.integers(1, 9), q=st.integers(1, 5), nw=st.integers(1, 3),
       cnt=st.integers(0, 60))
def test_gridding_is_linear_and_adjoint_to_degridding(seed, n, s, q, nw, cnt):
    gcf, u, v, wb, vis = _problem(seed, n, s, q, nw, cnt, span=0.6)
    z = np.zeros((n, n), complex)
    g = orc.convgrid(gcf, z, u, v, vis, wbin=wb)
    # linearity in the visibilities and accumulation onto an existing grid
    g2 = orc.convgrid(gcf, g, u, v, 2j * vis, wbin=wb)
    assert np.allclose(g2, (1 + 2j) * g, rtol=1e-12, atol=1e-12)
    # split invariance (what visibility-sharding relies on)
    h = cnt // 2
    ga = orc.convgrid(gcf, z, u[:h], v[:h], vis[:h], wbin=wb[:h])
    gb = orc.convgrid(gcf, ga, u[h:], v[h:], vis[h:], wbin=wb[h:])
    assert np.allclose(gb, g, rtol=1e-12, atol=1e-12)
    # <grid(v), m> == <v, degrid(m)>
    rng = np.random.default_rng(seed + 1)
    m = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    d = orc.convdegrid(gcf, m, u, v, wbin=wb)
    lhs, rhs = np.vdot(m, g), np.vdot(d, vis)
    assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs))
    # row-slab clipping (what uv-tile sharding relies on): gridding clipped to complementary slabs adds up
    r = n // 3
    lo, hi = np.zeros_like(m), np.zeros_like(m)
    lo[:r], hi[r:] = m[:r], m[r:]
    assert np.allclose(orc.convdegrid(gcf, lo, u, v, wbin=wb) + orc.convdegrid(gcf, hi, u, v, wbin=wb), d, rtol=1e-12, atol=1e-12)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10**6), n=st.sampled_from([16, 32, 64]), q=st.integers(1, 8), k=st.integers(-3, 3))
def test_frac_coord_integer_shift_equivariance(seed, n, q, k):
    """Moving p by k/n moves the cell by k and leaves the sub-cell index alone (n a power of two: p*n is exact)."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(-0.3, 0.3, 200)
    fl, fr = orc.frac_coord(n, q, p)
    fl2, fr2 = orc.frac_coord(n, q, p + k / n)
    x, x2 = n // 2 + p * n, n // 2 + (p + k / n) * n
    same = np.abs((x2 - x) - k) < 1e-12   # the shifted coordinate is exactly k cells away (no rounding in p + k/n)
    assert np.array_equal((fl2 - fl)[same] * q + (fr2 - fr)[same], np.full(same.sum(), k * q))
    assert fr.min() >= 0 and fr.max() < q


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10**6), n=st.sampled_from([3, 5, 8, 15]))
def test_convolve2d_is_bilinear_and_matches_fft_route(seed, n):
    rng = np.random.default_rng(seed)
    a, b, c = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) for _ in range(3))
    assert np.allclose(orc.convolve2d(a + 2 * c, b), orc.convolve2d(a, b) + 2 * orc.convolve2d(c, b), atol=1e-10)
    assert np.allclose(orc.convolve2d(a, b), orc.convolve2d_fft(a, b), atol=1e-10 * n * n)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10**6), n=st.sampled_from([4, 6, 9, 16]))
def test_hermitian_image_is_real_up_to_the_unsymmetrised_row_and_column(seed, n):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    h = orc.make_grid_hermitian(g)
    if n % 2 == 0:
        # zeroing row/col 0 (which the reference leaves unsymmetrised, Q5) makes the image exactly real
        h0 = h.copy(); h0[0, :] = 0; h0[:, 0] = 0
        assert np.abs(np.imag(orc.ifft(h0))).max() < 1e-12 * max(1.0, np.abs(h0).max())
    else:
        assert np.abs(np.imag(orc.ifft(h))).max() < 1e-12 * max(1.0, np.abs(h).max())
