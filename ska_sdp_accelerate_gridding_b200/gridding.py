"""Host-side mirror of the reference's `src/Gridding.hs` (module Gridding), same names, arity and argument
order, implemented as calls through the C ABI of libskagrid.so (include/skagrid.h) into sm_100a CUDA kernels.

Where the reference takes an `Acc (Array sh e)` this module takes a numpy array of the same shape:
    Vector BaseLines          -> tuple (u, v, w) of float64 vectors        (src/Types.hs:11, SoA as in Accelerate)
    Vector (Antenna, Antenna, Time, Frequency) -> tuple (a1, a2, t, f)
    Vector Visibility         -> complex128 vector
    Kernel / WKernels / AKernels -> complex128 [qpx,qpx,gh,gw] / [nw,qpx,qpx,s,s] / [nant,s,s]
    Matrix Visibility         -> complex128 [height, width], grid[y, x]
There is no CPU implementation here: every function fails if the CUDA library or a device is missing.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _lib
from .context import c128, f64, get_context, int64, ptr

FRAC_RAW = 0
FRAC_NORMALISE = 1


@dataclass
class KernelOptions:
    """src/Gridding.hs:30-38"""
    patHorShift: Optional[int] = None
    patVerShift: Optional[int] = None
    patTransMat: Optional[np.ndarray] = None
    wstep: Optional[int] = None
    qpx: Optional[int] = None
    npixFF: Optional[int] = None
    npixKern: Optional[int] = None


@dataclass
class OtherImagingArgs:
    """src/Gridding.hs:40-46"""
    convolutionKernel: Optional[np.ndarray] = None
    akernels: Optional[np.ndarray] = None
    wkernels: Optional[tuple] = None
    kernelCache: Optional[Callable] = None
    kernelFunction: Optional[Callable] = None


noArgs = KernelOptions()
noOtherArgs = OtherImagingArgs()


def _uv(p):
    u, v = f64(p[0]), f64(p[1])
    if u.shape != v.shape or u.ndim != 1:
        raise ValueError("uvw: u and v must be vectors of equal length")
    return u, v


def _uvw(p):
    u, v = _uv(p)
    w = f64(p[2])
    if w.shape != u.shape:
        raise ValueError("uvw: w must have the length of u")
    return u, v, w


def _grid_side(theta, lam):
    """N = P.round (theta * lam) (src/Gridding.hs:466): half to even, the library's one definition."""
    return int(_lib.load().skagrid_grid_side(float(theta), int(lam)))


# ----------------------------------------------------------------------------------------------- binning
def frac_coord(n, qpx, p, flags=FRAC_NORMALISE, ctx=None):
    """src/Gridding.hs:126-140 -> (flx, fracx) int64."""
    ctx = ctx or get_context()
    p = f64(p)
    fl = np.empty(p.shape, np.int64)
    fr = np.empty(p.shape, np.int64)
    ctx.check(ctx.lib.skagrid_frac_coord(ctx.h, n, qpx, p.size, ptr(p), ptr(fl), ptr(fr), flags))
    return fl, fr


def frac_coords(shape, qpx, p, flags=FRAC_NORMALISE, ctx=None):
    """src/Gridding.hs:142-151: shape = (height, width); -> (x, xf, y, yf)."""
    ctx = ctx or get_context()
    h, w = shape
    u, v = _uv(p)
    out = [np.empty(u.shape, np.int64) for _ in range(4)]
    ctx.check(ctx.lib.skagrid_frac_coords(ctx.h, h, w, qpx, u.size, ptr(u), ptr(v), *[ptr(o) for o in out], flags))
    return tuple(out)


def findClosest(ws, w, ctx=None):
    """src/Gridding.hs:895-907 (vectorised over w)."""
    ctx = ctx or get_context()
    ws = f64(ws)
    w = f64(np.atleast_1d(w))
    out = np.empty(w.shape, np.int64)
    ctx.check(ctx.lib.skagrid_find_closest(ctx.h, ws.size, ptr(ws), w.size, ptr(w), ptr(out)))
    return out


def div3(uvw, lam):
    """src/Gridding.hs:838-839.  Pure element-wise IEEE division; done by the imaging drivers on the device,
    offered here on the host arrays for callers that compose their own pipeline."""
    u, v, w = _uvw(uvw)
    lam = float(lam)
    return u / lam, v / lam, w / lam


# ----------------------------------------------------------------------------------------------- pre-steps
def mirror_uvw(uvw, vis, ctx=None):
    """src/Gridding.hs:551-562 -> ((u,v,w), vis)."""
    ctx = ctx or get_context()
    u, v, w = (a.copy() for a in _uvw(uvw))
    vis = c128(vis).copy()
    ctx.check(ctx.lib.skagrid_mirror_uvw(ctx.h, u.size, ptr(u), ptr(v), ptr(w), ptr(vis)))
    return (u, v, w), vis


def doweight(theta, lam, p, v, ctx=None):
    """src/Gridding.hs:564-583: p = uvw in wavelengths, v = the vector to be divided by the cell counts."""
    ctx = ctx or get_context()
    u, vv = _uv(p)
    out = c128(v).copy()
    ctx.check(ctx.lib.skagrid_doweight(ctx.h, float(theta), int(lam), u.size, ptr(u), ptr(vv), ptr(out)))
    return out


# ----------------------------------------------------------------------------------------------- gridders
def grid(a, p, v, ctx=None):
    """src/Gridding.hs:95-112."""
    ctx = ctx or get_context()
    a = c128(a).copy()
    u, vv = _uv(p)
    vis = c128(v)
    ctx.check(ctx.lib.skagrid_grid(ctx.h, a.shape[0], a.shape[1], ptr(a), u.size, ptr(u), ptr(vv), ptr(vis)))
    return a


def convgrid(gcf, a, p, v, ctx=None):
    """src/Gridding.hs:153-197."""
    ctx = ctx or get_context()
    gcf = c128(gcf)
    if gcf.ndim != 4 or gcf.shape[0] != gcf.shape[1]:
        raise ValueError("convgrid: gcf must be [qpx,qpx,gh,gw]")
    a = c128(a).copy()
    u, vv = _uv(p)
    vis = c128(v)
    qpx, _, gh, gw = gcf.shape
    ctx.check(ctx.lib.skagrid_convgrid(ctx.h, qpx, gh, gw, ptr(gcf), a.shape[0], a.shape[1], ptr(a), u.size, ptr(u), ptr(vv), ptr(vis)))
    return a


def convgrid2(gcf, a, p, wbin, v, ctx=None):
    """src/Gridding.hs:199-244."""
    ctx = ctx or get_context()
    gcf = c128(gcf)
    if gcf.ndim != 5 or gcf.shape[1] != gcf.shape[2]:
        raise ValueError("convgrid2: gcf must be [nw,qpx,qpx,gh,gw]")
    a = c128(a).copy()
    vis = c128(v)
    u, vv, wbin, n = _coords(p, wbin, vis.size)  # p=None: grid new values (weights, residuals) at the previous call's coordinates
    nw, qpx, _, gh, gw = gcf.shape
    ctx.check(ctx.lib.skagrid_convgrid2(ctx.h, nw, qpx, gh, gw, ptr(gcf), a.shape[0], a.shape[1], ptr(a), n, ptr(u), ptr(vv),
                                        ptr(wbin), ptr(vis)))
    return a


def _aw_args(wkerns, akerns, index):
    wkerns, akerns = c128(wkerns), c128(akerns)
    if wkerns.ndim != 5 or akerns.ndim != 3 or wkerns.shape[3] != wkerns.shape[4] or akerns.shape[1:] != wkerns.shape[3:]:
        raise ValueError("AW gridding: wkerns must be [nw,qpx,qpx,s,s] and akerns [nant,s,s]")
    wbin, a1, a2 = (int64(x) for x in index)
    return wkerns, akerns, wbin, a1, a2


def convgrid3(wkerns, akerns, a, p, index, v, ctx=None):
    """src/Gridding.hs:246-317 (index = (wbin, a1, a2)); same result as convgrid4."""
    ctx = ctx or get_context()
    wkerns, akerns, wbin, a1, a2 = _aw_args(wkerns, akerns, index)
    a = c128(a).copy()
    u, vv = _uv(p)
    vis = c128(v)
    nw, qpx, _, s, _ = wkerns.shape
    ctx.check(ctx.lib.skagrid_convgrid_aw(ctx.h, nw, qpx, s, ptr(wkerns), akerns.shape[0], ptr(akerns), a.shape[0], a.shape[1], ptr(a),
                                          u.size, ptr(u), ptr(vv), ptr(wbin), ptr(a1), ptr(a2), ptr(vis)))
    return a


convgrid4 = convgrid3  # src/Gridding.hs:318-377: same semantics, different (batched) evaluation order


# degridding: not in the reference; exact adjoints of the gridders above
def _coords(p, wbin, count):
    """(u, v, wbin, n); p=None: the coordinates the previous table gridder / degridder call left on the device
    (`count` of them; include/skagrid.h "resident coordinates")."""
    if p is None:
        if count is None:
            raise ValueError("p=None (resident coordinates) needs count")
        return None, None, None, int(count)
    u, vv = _uv(p)
    return u, vv, (None if wbin is None else int64(wbin)), u.size


def convdegrid(gcf, a, p, ctx=None, count=None):
    ctx = ctx or get_context()
    gcf, a = c128(gcf), c128(a)
    u, vv, _, n = _coords(p, None, count)
    out = np.empty(n, np.complex128)
    qpx, _, gh, gw = gcf.shape
    ctx.check(ctx.lib.skagrid_convdegrid(ctx.h, qpx, gh, gw, ptr(gcf), a.shape[0], a.shape[1], ptr(a), n, ptr(u), ptr(vv), ptr(out)))
    return out


def convdegrid2(gcf, a, p, wbin, ctx=None, count=None):
    """Adjoint of convgrid2.  p=None, wbin=None, count=n: degrid at the coordinates of the previous call (no re-upload)."""
    ctx = ctx or get_context()
    gcf, a = c128(gcf), c128(a)
    u, vv, wbin, n = _coords(p, wbin, count)
    out = np.empty(n, np.complex128)
    nw, qpx, _, gh, gw = gcf.shape
    ctx.check(ctx.lib.skagrid_convdegrid2(ctx.h, nw, qpx, gh, gw, ptr(gcf), a.shape[0], a.shape[1], ptr(a), n, ptr(u), ptr(vv),
                                          ptr(wbin), ptr(out)))
    return out


def convdegrid3(wkerns, akerns, a, p, index, ctx=None):
    ctx = ctx or get_context()
    wkerns, akerns, wbin, a1, a2 = _aw_args(wkerns, akerns, index)
    a = c128(a)
    u, vv = _uv(p)
    out = np.empty(u.size, np.complex128)
    nw, qpx, _, s, _ = wkerns.shape
    ctx.check(ctx.lib.skagrid_convdegrid_aw(ctx.h, nw, qpx, s, ptr(wkerns), akerns.shape[0], ptr(akerns), a.shape[0], a.shape[1], ptr(a),
                                            u.size, ptr(u), ptr(vv), ptr(wbin), ptr(a1), ptr(a2), ptr(out)))
    return out


convdegrid4 = convdegrid3


# ----------------------------------------------------------------------------------------------- AW kernels
def convolve2d(a1, a2, ctx=None):
    """src/Gridding.hs:795-811."""
    ctx = ctx or get_context()
    a1, a2 = c128(a1), c128(a2)
    if a1.shape != a2.shape or a1.ndim != 2 or a1.shape[0] != a1.shape[1]:
        raise ValueError("convolve2d: two square matrices of equal size")
    out = np.empty_like(a1)
    ctx.check(ctx.lib.skagrid_convolve2d(ctx.h, a1.shape[0], ptr(a1), ptr(a2), ptr(out)))
    return out


def aw_kernel_fn2(yf, xf, wkerns, akerns, wbin, a1, a2, ctx=None):
    """src/Gridding.hs:761-775, batched: kernel k = convolve2d (convolve2d akerns[a1[k]] akerns[a2[k]])
    wkerns[wbin[k], yf[k], xf[k]] -> [count, s, s] (not conjugated)."""
    ctx = ctx or get_context()
    wkerns, akerns, wbin, a1, a2 = _aw_args(wkerns, akerns, (np.atleast_1d(wbin), np.atleast_1d(a1), np.atleast_1d(a2)))
    yf, xf = int64(np.atleast_1d(yf)), int64(np.atleast_1d(xf))
    nw, qpx, _, s, _ = wkerns.shape
    out = np.empty((wbin.size, s, s), np.complex128)
    ctx.check(ctx.lib.skagrid_aw_kernel(ctx.h, nw, qpx, s, ptr(wkerns), akerns.shape[0], ptr(akerns), wbin.size, ptr(wbin), ptr(yf),
                                        ptr(xf), ptr(a1), ptr(a2), ptr(out)))
    return out


def w_kernel(theta, w, kernops: KernelOptions, conjugate=False, ctx=None):
    """src/Gridding.hs:610-619 for one or many w -> [qpx,qpx,s,s] (scalar w) or [nw,qpx,qpx,s,s]."""
    ctx = ctx or get_context()
    ws = f64(np.atleast_1d(w))
    npixff, npixkern, qpx = int(kernops.npixFF), int(kernops.npixKern), int(kernops.qpx)
    out = np.empty((ws.size, qpx, qpx, npixkern, npixkern), np.complex128)
    if kernops.patHorShift or kernops.patVerShift or kernops.patTransMat is not None:   # kernel_coordinates, src/Gridding.hs:620-635
        t = None if kernops.patTransMat is None else f64(kernops.patTransMat).reshape(2, 2)
        ctx.check(ctx.lib.skagrid_w_kernels_ex(ctx.h, float(theta), ws.size, ptr(ws), npixff, npixkern, qpx, int(bool(conjugate)), ptr(t),
                                               float(kernops.patHorShift or 0), float(kernops.patVerShift or 0), ptr(out)))
    else:
        ctx.check(ctx.lib.skagrid_w_kernels(ctx.h, float(theta), ws.size, ptr(ws), npixff, npixkern, qpx, int(bool(conjugate)), ptr(out)))
    return out[0] if np.ndim(w) == 0 else out


# ----------------------------------------------------------------------------------------------- grid -> image
def make_grid_hermitian(g, ctx=None):
    """src/Gridding.hs:585-605."""
    ctx = ctx or get_context()
    g = c128(g)
    out = np.empty_like(g)
    ctx.check(ctx.lib.skagrid_make_grid_hermitian(ctx.h, g.shape[0], ptr(g), ptr(out)))
    return out


def ifft(g, ctx=None):
    """src/Gridding.hs:828-829."""
    ctx = ctx or get_context()
    g = c128(g)
    out = np.empty_like(g)
    ctx.check(ctx.lib.skagrid_ifft(ctx.h, g.shape[0], ptr(g), ptr(out)))
    return out


def fft(g, ctx=None):
    """src/Gridding.hs:821-826."""
    ctx = ctx or get_context()
    g = c128(g)
    out = np.empty_like(g)
    ctx.check(ctx.lib.skagrid_fft(ctx.h, g.shape[0], ptr(g), ptr(out)))
    return out


def grid_to_image(g, want_image=True, ctx=None, n=None):
    """make_grid_hermitian -> ifft -> map real -> maximum, fused (src/ImageDataset.hs:74-77) -> (image, max).
    g=None with n: the n x n grid the previous call left on the context's device."""
    ctx = ctx or get_context()
    if g is None:
        if n is None:
            raise ValueError("grid_to_image: g=None needs n")
        shape = (int(n), int(n))
    else:
        g = c128(g)
        shape = g.shape
    img = np.empty(shape, np.float64) if want_image else None
    mx = np.empty(1, np.float64)
    ctx.check(ctx.lib.skagrid_grid_to_image(ctx.h, shape[0], ptr(g), ptr(img), ptr(mx)))
    return img, float(mx[0])


# ----------------------------------------------------------------------------------------------- imaging drivers
def simple_imaging(theta, lam, uvw, src, vis, ctx=None):
    """src/Gridding.hs:84-93."""
    ctx = ctx or get_context()
    u, v, w = _uvw(uvw)
    vis = c128(vis)
    n = _grid_side(theta, lam)
    out = np.empty((n, n), np.complex128)
    ctx.check(ctx.lib.skagrid_simple_imaging(ctx.h, float(theta), int(lam), u.size, ptr(u), ptr(v), ptr(w), ptr(vis), ptr(out)))
    return out


def conv_imaging(kv, theta, lam, uvw, src, vis, ctx=None):
    """src/Gridding.hs:115-124."""
    ctx = ctx or get_context()
    kv = c128(kv)
    u, v, w = _uvw(uvw)
    vis = c128(vis)
    n = _grid_side(theta, lam)
    qpx, _, gh, gw = kv.shape
    out = np.empty((n, n), np.complex128)
    ctx.check(ctx.lib.skagrid_conv_imaging(ctx.h, qpx, gh, gw, ptr(kv), float(theta), int(lam), u.size, ptr(u), ptr(v), ptr(w), ptr(vis),
                                           ptr(out)))
    return out


def conv_imaging2(kv, theta, lam, uvw, wbin, vis, ctx=None):
    """5-D analogue of conv_imaging (src/Gridding.hs:115-124) = the last step of w_cache_imaging (:421-449):
    zero grid, p = uvw/lam, convgrid2 with the w-plane index per visibility."""
    ctx = ctx or get_context()
    kv = c128(kv)
    u, v, w = _uvw(uvw)
    vis, wbin = c128(vis), int64(wbin)
    n = _grid_side(theta, lam)
    nw, qpx, _, gh, gw = kv.shape
    out = np.empty((n, n), np.complex128)
    ctx.check(ctx.lib.skagrid_conv_imaging2(ctx.h, nw, qpx, gh, gw, ptr(kv), float(theta), int(lam), u.size, ptr(u), ptr(v), ptr(w), ptr(wbin),
                                            ptr(vis), ptr(out)))
    return out


def _round_half_away(x):
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def w_cache_imaging(kernops: KernelOptions, otargs: OtherImagingArgs, theta, lam, uvw, src, vis, ctx=None):
    """src/Gridding.hs:399-449: w rounded to multiples of wstep, one conjugated w-kernel per step, convgrid2.
    otargs.kernelFunction / otargs.kernelCache (KernelF, :48-53: theta w a1 a2 t f kernops -> [qpx,qpx,s,s]) replace the default
    w_kernel exactly as in the reference: kernel_cache = fromMaybe kernel_fn kernel_cache', called with the four Maybe arguments
    set to Nothing (:410-412); the result is conjugated (:441)."""
    ctx = ctx or get_context()
    u, v, w = _uvw(uvw)
    wstep = int(kernops.wstep if kernops.wstep is not None else 2000)
    rounded = (wstep * _round_half_away(w / float(wstep))).astype(np.int64)
    wmin, wmax = int(rounded.min()), int(rounded.max())
    steps = (wmax - wmin) // wstep + 1
    wbins = (rounded - wmin) // wstep
    ws = np.array([float(i * wstep + wmin) for i in range(steps)])
    kernel_fn = otargs.kernelCache if otargs.kernelCache is not None else otargs.kernelFunction
    if kernel_fn is None:
        kernels = w_kernel(theta, ws, kernops, conjugate=True, ctx=ctx)
    else:   # a caller-supplied kernel generator runs on the host, once per w step; the gridding stays on the GPU
        kernels = np.conj(np.stack([c128(kernel_fn(theta, float(wi), None, None, None, None, kernops)) for wi in ws]))
        if kernels.ndim != 5 or kernels.shape[1] != kernels.shape[2]:
            raise ValueError("kernelFunction must return [qpx, qpx, s, s] kernels")
    return conv_imaging2(kernels, theta, lam, (u, v, w), wbins, vis, ctx=ctx)


def aw_imaging(kernops, otargs, theta, lam, wkernels, wbins, akernels, uvw, src, vis, ctx=None):
    """src/Gridding.hs:452-478 (convgrid4 path)."""
    ctx = ctx or get_context()
    wkernels, akernels = c128(wkernels), c128(akernels)
    wbins = f64(wbins)
    u, v, w = _uvw(uvw)
    a1, a2 = int64(src[0]), int64(src[1])
    vis = c128(vis)
    n = _grid_side(theta, lam)
    nw, qpx, _, s, _ = wkernels.shape
    out = np.empty((n, n), np.complex128)
    ctx.check(ctx.lib.skagrid_aw_imaging(ctx.h, float(theta), int(lam), nw, qpx, s, ptr(wkernels), ptr(wbins), akernels.shape[0],
                                         ptr(akernels), u.size, ptr(u), ptr(v), ptr(w), ptr(a1), ptr(a2), ptr(vis), ptr(out)))
    return out


aw_imagingOld = aw_imaging  # src/Gridding.hs:480-506: convgrid3 instead of convgrid4, same grid


def do_imaging(theta, lam, uvw, a1, a2, t, f, vis, imgfn, ctx=None):
    """src/Gridding.hs:509-549: uvw is the [rows,3] matrix; returns (dirty/pmax, psf/pmax, pmax)."""
    ctx = ctx or get_context()
    uvw = f64(uvw)
    uvw0 = (uvw[:, 0].copy(), uvw[:, 1].copy(), uvw[:, 2].copy())
    src0 = (int64(a1), int64(a2), f64(t), np.full(len(vis), float(f)))
    uvw1, vis1 = mirror_uvw(uvw0, vis, ctx=ctx)
    wt = doweight(theta, lam, uvw1, np.ones(len(vis1), np.complex128), ctx=ctx)
    drt, _ = grid_to_image(imgfn(theta, lam, uvw1, src0, wt * vis1), ctx=ctx)
    psf, pmax = grid_to_image(imgfn(theta, lam, uvw1, src0, wt), ctx=ctx)
    return drt / pmax, psf / pmax, pmax
