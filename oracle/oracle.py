"""CPU oracle for the AW-gridding hot path (ctypes over oracle/liboracle.so + numpy).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs may import it.  Nothing under
``ska_sdp_accelerate_gridding_b200/`` imports, calls or falls back to this module.

Parity status ("parity partially pinned"): the reference cannot be built here (no GHC / LLVM 7 /
libhdf5) and its data files are git-LFS stubs, so the oracle is pinned by the only numbers the
reference stores -- old/BrokenNumbers.hs:85-91 (scatter-add golden) -- plus the in-source inputs of
test/SmallTest.hs:51-76 (all three formulations must agree).  FFT normalisation, shift convention
and rounding ties are unpinned; SURVEY.md section 8 Q1-Q6 is the working spec.

Every function cites the reference file:line (relative to /root/reference) it restates.
numpy is used where the reference delegates to accelerate-fft (FFTW): fft2D / shift2D / ishift2D.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64 = C.c_int64
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (gcc; see oracle/Makefile)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_find_closest1.restype = _i64
        _LIB.orc_grid_simple.restype = _i64
        _LIB.orc_num_threads.restype = C.c_int
        _LIB.orc_doweight.restype = C.c_int
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _int64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_cores() -> int:
    """Let OpenMP use every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(C.c_int(n))
    return num_threads()


# ----------------------------------------------------------------------------- binning
def frac_coord(n, qpx, p, normalise=True):
    """src/Gridding.hs:126-140.  Returns (fl, frac) int64 arrays."""
    p = _f64(p)
    fl = np.empty(p.shape, np.int64)
    fr = np.empty(p.shape, np.int64)
    lib().orc_frac_coord(_i64(n), _i64(qpx), _i64(p.size), _d(p), _i(fl), _i(fr), C.c_int(int(normalise)))
    return fl, fr


def frac_coords(h, w, qpx, u, v, normalise=True):
    """src/Gridding.hs:142-151.  Returns (x, xf, y, yf)."""
    x, xf = frac_coord(w, qpx, u, normalise)
    y, yf = frac_coord(h, qpx, v, normalise)
    return x, xf, y, yf


def find_closest(ws, w):
    """src/Gridding.hs:895-907 (+Q4 clamp)."""
    ws = _f64(ws)
    w = _f64(np.atleast_1d(w))
    out = np.empty(w.shape, np.int64)
    lib().orc_find_closest(_i64(ws.size), _d(ws), _i64(w.size), _d(w), _i(out))
    return out


def find_closest_py(ws, w):
    """Pure-Python transliteration of findClosest (src/Gridding.hs:895-907) used to
    cross-check the C version on small cases."""
    mn, mx = 0, len(ws)
    while (mx - mn) // 2 >= 1:
        i = (mx + mn) // 2
        if w > ws[i]:
            mn = i
        else:
            mx = i
    if mx >= len(ws):
        return mn
    return mn if abs(w - ws[mn]) < abs(w - ws[mx]) else mx


# ----------------------------------------------------------------------------- pre-steps
def uvw_lambda(f, u, v, w):
    """src/ImageDataset.hs:181-187."""
    u, v, w = _f64(u).copy(), _f64(v).copy(), _f64(w).copy()
    lib().orc_uvw_lambda(C.c_double(f), _i64(u.size), _d(u), _d(v), _d(w))
    return u, v, w


def div3(lam, u, v, w):
    """src/Gridding.hs:838-839."""
    u, v, w = _f64(u).copy(), _f64(v).copy(), _f64(w).copy()
    lib().orc_div3(C.c_double(lam), _i64(u.size), _d(u), _d(v), _d(w))
    return u, v, w


def mirror_uvw(u, v, w, vis):
    """src/Gridding.hs:551-562."""
    u, v, w, vis = _f64(u).copy(), _f64(v).copy(), _f64(w).copy(), _c128(vis).copy()
    lib().orc_mirror_uvw(_i64(u.size), _d(u), _d(v), _d(w), _d(vis))
    return u, v, w, vis


def doweight(theta, lam, u, v, vis):
    """src/Gridding.hs:564-583; u,v in wavelengths (divided by lam inside)."""
    u, v, vis = _f64(u), _f64(v), _c128(vis).copy()
    rc = lib().orc_doweight(C.c_double(theta), _i64(lam), _i64(u.size), _d(u), _d(v), _d(vis))
    if rc == -2:
        raise MemoryError
    return vis


# ----------------------------------------------------------------------------- gridders
def scatter_add(grid, x, y, val):
    """permute (+) (old/BrokenNumbers.hs:31-44)."""
    grid = _c128(grid).copy()
    x, y, val = _int64(x), _int64(y), _c128(val)
    lib().orc_scatter_add(_i64(grid.shape[0]), _i64(grid.shape[1]), _d(grid), _i64(x.size), _i(x), _i(y), _d(val))
    return grid


def grid_simple(grid, u, v, vis):
    """`grid` src/Gridding.hs:95-112."""
    grid = _c128(grid).copy()
    u, v, vis = _f64(u), _f64(v), _c128(vis)
    lib().orc_grid_simple(_i64(grid.shape[0]), _i64(grid.shape[1]), _d(grid), _i64(u.size), _d(u), _d(v), _d(vis))
    return grid


def convgrid(gcf, grid, u, v, vis, wbin=None, normalise=True, parallel=False):
    """convgrid (4-D gcf, src/Gridding.hs:153-197) / convgrid2 (5-D gcf + wbin, :199-244)."""
    gcf = _c128(gcf)
    grid = _c128(grid).copy()
    if gcf.ndim == 4:
        nw, (qpx, _, gh, gw) = 1, gcf.shape
        wb = None
    else:
        nw, qpx, _, gh, gw = gcf.shape
        wb = _int64(wbin)
    u, v, vis = _f64(u), _f64(v), _c128(vis)
    fn = lib().orc_convgrid2_omp if parallel else lib().orc_convgrid2
    fn(_i64(nw), _i64(qpx), _i64(gh), _i64(gw), _d(gcf), _i64(grid.shape[0]), _i64(grid.shape[1]), _d(grid),
       _i64(u.size), _d(u), _d(v), _i(wb), _d(vis), C.c_int(int(normalise)))
    return grid


def convdegrid(gcf, grid, u, v, wbin=None, normalise=True, parallel=False):
    """Adjoint of convgrid/convgrid2 (absent from the reference; SURVEY 8c definition)."""
    gcf = _c128(gcf)
    grid = _c128(grid)
    if gcf.ndim == 4:
        nw, (qpx, _, gh, gw) = 1, gcf.shape
        wb = None
    else:
        nw, qpx, _, gh, gw = gcf.shape
        wb = _int64(wbin)
    u, v = _f64(u), _f64(v)
    out = np.empty(u.size, np.complex128)
    fn = lib().orc_convdegrid2_omp if parallel else lib().orc_convdegrid2
    fn(_i64(nw), _i64(qpx), _i64(gh), _i64(gw), _d(gcf), _i64(grid.shape[0]), _i64(grid.shape[1]), _d(grid),
       _i64(u.size), _d(u), _d(v), _i(wb), _d(out), C.c_int(int(normalise)))
    return out


def convolve2d(a1, a2):
    """Direct form of convolve2d (src/Gridding.hs:795-811), C implementation."""
    a1, a2 = _c128(a1), _c128(a2)
    out = np.empty_like(a1)
    lib().orc_convolve2d(_i64(a1.shape[0]), _d(a1), _d(a2), _d(out))
    return out


def aw_kernel(wkern_plane, yf, xf, a1, a2):
    """aw_kernel_fn2 (src/Gridding.hs:761-775): convolve2d (convolve2d a1 a2) (w[yf,xf])."""
    return convolve2d(convolve2d(a1, a2), _c128(wkern_plane)[yf, xf])


def convgrid_aw(wkerns, akerns, grid, u, v, wbin, a1, a2, vis, normalise=True):
    """convgrid3 / convgrid4 (src/Gridding.hs:246-317 / :318-396)."""
    wkerns, akerns = _c128(wkerns), _c128(akerns)
    grid = _c128(grid).copy()
    nw, qpx, _, s, _ = wkerns.shape
    u, v, vis = _f64(u), _f64(v), _c128(vis).copy()
    wbin, a1, a2 = _int64(wbin), _int64(a1), _int64(a2)
    lib().orc_convgrid_aw(_i64(nw), _i64(qpx), _i64(s), _d(wkerns), _i64(akerns.shape[0]), _d(akerns),
                          _i64(grid.shape[0]), _i64(grid.shape[1]), _d(grid), _i64(u.size), _d(u), _d(v),
                          _i(wbin), _i(a1), _i(a2), _d(vis), C.c_int(int(normalise)), C.c_int(0))
    return grid


def convdegrid_aw(wkerns, akerns, grid, u, v, wbin, a1, a2, normalise=True):
    """Adjoint of convgrid3/4 (absent from the reference)."""
    wkerns, akerns = _c128(wkerns), _c128(akerns)
    grid = _c128(grid).copy()
    nw, qpx, _, s, _ = wkerns.shape
    u, v = _f64(u), _f64(v)
    vis = np.zeros(u.size, np.complex128)
    wbin, a1, a2 = _int64(wbin), _int64(a1), _int64(a2)
    lib().orc_convgrid_aw(_i64(nw), _i64(qpx), _i64(s), _d(wkerns), _i64(akerns.shape[0]), _d(akerns),
                          _i64(grid.shape[0]), _i64(grid.shape[1]), _d(grid), _i64(u.size), _d(u), _d(v),
                          _i(wbin), _i(a1), _i(a2), _d(vis), C.c_int(int(normalise)), C.c_int(1))
    return vis


def make_grid_hermitian(g):
    """src/Gridding.hs:585-605."""
    g = _c128(g)
    out = np.empty_like(g)
    lib().orc_make_grid_hermitian(_i64(g.shape[0]), _d(g), _d(out))
    return out


# ----------------------------------------------------------------------------- FFT side (numpy)
def shift2d(a):
    """accelerate-fft shift2D: out[i] = in[(i + ceil(n/2)) mod n] per axis (old/ShiftExample.hs:98-106)
    == numpy.fft.fftshift."""
    return np.fft.fftshift(a, axes=(-2, -1))


def ishift2d(a):
    """accelerate-fft ishift2D (inverse of shift2D) == numpy.fft.ifftshift.  Identical to shift2D
    for even sizes, which is every size on the hot path (Q5)."""
    return np.fft.ifftshift(a, axes=(-2, -1))


def ifft(g):
    """src/Gridding.hs:828-829: shift2D . fft2D Inverse . ishift2D (1/N^2-normalised inverse)."""
    return shift2d(np.fft.ifft2(ishift2d(_c128(g))))


def fft(g):
    """src/Gridding.hs:821-826 with the pow2 pad being a no-op for pow2 n; general n pads to the next
    power of two, transforms, and extracts the middle."""
    g = _c128(g)
    n_ = g.shape[0]
    n = 1 << int(np.ceil(np.log2(n_)))
    big = pad_mid(g, n)
    return extract_mid(shift2d(np.fft.fft2(ishift2d(big))), n_)


def padder(a, pad_x, pad_y, cval=0):
    """src/Gridding.hs:863-877.  NB the reference reads array ! index2 oldx oldy, i.e. it
    TRANSPOSES the input while padding (Q1)."""
    x0, x1 = pad_x
    y0, y1 = pad_y
    m, n = a.shape
    out = np.full((m + y0 + y1, n + x0 + x1), cval, dtype=a.dtype)
    for y in range(out.shape[0]):
        for x in range(out.shape[1]):
            ox, oy = x - x0, y - y0
            if 0 <= ox < n and 0 <= oy < m:
                out[y, x] = a[ox, oy]
    return out


def pad_mid(ff, n):
    """src/Gridding.hs:682-691 (returns the input untouched when n == n0)."""
    n0 = ff.shape[0]
    if n == n0:
        return ff
    pw = (n // 2 - n0 // 2, (n + 1) // 2 - (n0 + 1) // 2)
    return padder(ff, pw, pw, 0)


def extract_mid(a, n):
    """src/Gridding.hs:694-707."""
    cx, cy = a.shape[0] // 2, a.shape[1] // 2
    s = n // 2
    return a[cx - s:cx - s + n, cy - s:cy - s + n]


def convolve2d_fft(a1, a2):
    """Literal restatement of convolve2d's FFT route (src/Gridding.hs:795-811)."""
    a1, a2 = _c128(a1), _c128(a2)
    n = a1.shape[0]
    m = 1 << int(np.ceil(np.log2(2 * n - 1)))
    f1 = np.fft.ifft2(ishift2d(pad_mid(a1, m)))
    f2 = np.fft.ifft2(ishift2d(pad_mid(a2, m)))
    conv = shift2d(np.fft.fft2(f1 * f2))
    return extract_mid(conv, n) * float(m * m)


# ----------------------------------------------------------------------------- w-kernel generation
def coordinates2(n):
    """src/Gridding.hs:637-648: (samecolumns, samerows); base = -(n//2)/n + k/n."""
    n2 = n // 2
    step = 1.0 / n
    base = (-n2) * step + np.arange(n) * step
    samecolumns = np.tile(base[None, :], (n, 1))
    samerows = np.tile(base[:, None], (1, n))
    return samecolumns, samerows


def w_kernel_function(l, m, w):
    """src/Gridding.hs:651-667: exp(2 pi i w (1 - sqrt(1 - l^2 - m^2)))."""
    r2 = l * l + m * m
    ph = 1 - np.sqrt(1 - r2)
    return np.exp(1j * (2 * np.pi * w * ph))


def extract_oversampled(a, qpx, n):
    """src/Gridding.hs:709-728: out[yf,xf,y,x] = qpx^2 * a[na/2 - qpx*(n/2) - yf + qpx*y, ... - xf + qpx*x]."""
    na = a.shape[1]
    cons = na // 2 - qpx * (n // 2)
    out = np.empty((qpx, qpx, n, n), np.complex128)
    for yf in range(qpx):
        for xf in range(qpx):
            ys = cons - yf + qpx * np.arange(n)
            xs = cons - xf + qpx * np.arange(n)
            out[yf, xf] = a[np.ix_(ys, xs)]
    return out * float(qpx * qpx)


def kernel_coordinates(n, theta, dl=0, dm=0, transmat=None):
    """src/Gridding.hs:620-635: theta * coordinates2, through patTransMat t as (x, y) -> (t[0,0] x + t[1,0] y, t[0,1] x + t[1,1] y),
    then shifted by (patHorShift, patVerShift)."""
    l, m = coordinates2(n)
    l, m = l * theta, m * theta
    if transmat is not None:
        t = np.asarray(transmat, dtype=np.float64).reshape(2, 2)
        l, m = t[0, 0] * l + t[1, 0] * m, t[0, 1] * l + t[1, 1] * m
    return l + float(dl), m + float(dm)


def w_kernel(theta, w, npixff, npixkern, qpx, dl=0, dm=0, transmat=None):
    """w_kernel (src/Gridding.hs:610-619) = kernel_coordinates (:620-635) -> w_kernel_function ->
    kernel_oversample (:669-680).  Returns [qpx,qpx,npixkern,npixkern]."""
    l, m = kernel_coordinates(npixff, theta, dl, dm, transmat)
    ff = w_kernel_function(l, m, w)
    padff = pad_mid(ff, npixff * qpx)
    af = ifft(padff)
    return extract_oversampled(af, qpx, npixkern)


# ----------------------------------------------------------------------------- drivers
def simple_imaging(theta, lam, u, v, w, vis):
    """src/Gridding.hs:84-93."""
    n = int(round(theta * lam))
    pu, pv, _ = div3(float(lam), u, v, w)
    return grid_simple(np.zeros((n, n), np.complex128), pu, pv, vis)


def conv_imaging(kv, theta, lam, u, v, w, vis):
    """src/Gridding.hs:115-124."""
    n = int(round(theta * lam))
    pu, pv, _ = div3(float(lam), u, v, w)
    return convgrid(kv, np.zeros((n, n), np.complex128), pu, pv, vis)


def aw_imaging(theta, lam, wkernels, wbins, akernels, u, v, w, a1, a2, vis):
    """src/Gridding.hs:452-478: p = uvw/lam; wbin = findClosest wbins w (w in wavelengths, NOT
    divided by lam); convgrid4."""
    n = int(round(theta * lam))
    pu, pv, _ = div3(float(lam), u, v, w)
    closest = find_closest(wbins, w)
    return convgrid_aw(wkernels, akernels, np.zeros((n, n), np.complex128), pu, pv, closest, a1, a2, vis)


def aw_gridding(theta, lam, wkernels, wbins, akernels, u_m, v_m, w_m, a1, a2, freq, vis, count=None):
    """src/ImageDataset.hs:29-83 from the point the arrays are loaded: uvw_lambda, doweight on the
    UN-mirrored uvw (:59), mirror_uvw (:60), aw_imaging on vis*wt (:72-73), make_grid_hermitian,
    real(ifft), max.  Returns (image, max, uvgrid)."""
    cnt = len(vis) if count is None else count
    u, v, w = uvw_lambda(freq, u_m[:cnt], v_m[:cnt], w_m[:cnt])
    vis0 = _c128(vis[:cnt])
    wt = doweight(theta, lam, u, v, np.ones(cnt, np.complex128))
    u1, v1, w1, vis1 = mirror_uvw(u, v, w, vis0)
    uvgrid = aw_imaging(theta, lam, wkernels, wbins, akernels, u1, v1, w1, a1[:cnt], a2[:cnt], vis1 * wt)
    img = np.real(ifft(make_grid_hermitian(uvgrid)))
    return img, float(img.max()), uvgrid


def do_imaging(theta, lam, u, v, w, vis, imgfn):
    """src/Gridding.hs:509-549: mirror, weight (on MIRRORED uvw), dirty image, PSF, normalise by max(psf).
    imgfn(theta, lam, u, v, w, vis) -> uvgrid."""
    u1, v1, w1, vis1 = mirror_uvw(u, v, w, vis)
    wt = doweight(theta, lam, u1, v1, np.ones(len(vis1), np.complex128))
    drt = np.real(ifft(make_grid_hermitian(imgfn(theta, lam, u1, v1, w1, wt * vis1))))
    psf = np.real(ifft(make_grid_hermitian(imgfn(theta, lam, u1, v1, w1, wt))))
    pmax = psf.max()
    return drt / pmax, psf / pmax, pmax
