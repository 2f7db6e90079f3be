"""Host-side mirror of the reference's `src/ImageDataset.hs`: the end-to-end AW-gridding driver and the
dataset / kernel loaders' selection logic.

The reference reads HDF5 through its own C shim (hdf5/hdf5.cc); libhdf5 is not available in this build
environment, so the loaders here take the same *logical* layout from `.npz` archives (one array per HDF5
dataset path, '/' kept in the key) -- see INTEGRATION.md for the mapping.  The arithmetic all happens on
the GPU through libskagrid.so.
"""
from __future__ import annotations

import numpy as np

from . import gridding as G
from .context import c128, f64, get_context, int64, ptr


def uvw_lambda(f, uvw):
    """src/ImageDataset.hs:181-187: (u,v,w) * (f / 299792458.0).  Returns new arrays (device computed)."""
    ctx = get_context()
    u, v, w = (f64(a).copy() for a in uvw)
    ctx.check(ctx.lib.skagrid_uvw_lambda(ctx.h, float(f), u.size, ptr(u), ptr(v), ptr(w)))
    return u, v, w


def findClosestList(ws, w0):
    """src/ImageDataset.hs:151-167: index of the value of the sorted list `ws` closest to w0 (host-side,
    used only to pick the time / frequency of the A-kernels)."""
    ws = list(ws)
    mn, mx = 0, len(ws) - 1
    while (mx - mn) // 2 >= 1:
        i = (mx + mn) // 2
        if w0 > ws[i]:
            mn = i
        else:
            mx = i
    return mn if abs(w0 - ws[mn]) < abs(w0 - ws[mx]) else mx


def convertAndSort(names):
    """src/ImageDataset.hs:170-178: numeric sort of group names, returns (sorted values, sorted names)."""
    vals = [float(s) for s in names]
    order = np.argsort(vals, kind="stable")
    return [vals[i] for i in order], [names[i] for i in order]


def _fmt(x):
    """Haskell `printf "%f"` on a Double prints the shortest digits that round-trip (0.008 -> "0.008")."""
    return np.format_float_positional(float(x), trim="-")


def _members(store, prefix):
    pre = prefix.rstrip("/") + "/"
    out = []
    for k in store.keys():
        if k.startswith(pre):
            m = k[len(pre):].split("/")[0]
            if m not in out:
                out.append(m)
    return out


def getWKernels(store, theta):
    """src/ImageDataset.hs:136-148: stack /wkern/<theta>/<w>/kern in numeric w order -> (wkernels, wbins)."""
    base = "/wkern/%s" % _fmt(theta)
    ws, names = convertAndSort(_members(store, base))
    kerns = np.stack([c128(store["%s/%s/kern" % (base, n)]) for n in names])
    return kerns, np.array(ws, np.float64)


def getAKernels(store, theta, t0, f0):
    """src/ImageDataset.hs:108-133: per antenna (numeric order) the kernel of the closest time and frequency."""
    base = "/akern/%s" % _fmt(theta)
    _, ants = convertAndSort(_members(store, base))
    ts, tnames = convertAndSort(_members(store, "%s/%s" % (base, ants[0])))
    tname = tnames[findClosestList(ts, t0)]
    fs, fnames = convertAndSort(_members(store, "%s/%s/%s" % (base, ants[0], tname)))
    fname = fnames[findClosestList(fs, f0)]
    return np.stack([c128(store["%s/%s/%s/%s/kern" % (base, a, tname, fname)]) for a in ants])


def aw_gridding_arrays(theta, lam, wkernels, wbins, akernels, uvw_m, a1, a2, freq, vis, n=None, want_image=True, want_grid=False,
                       ctx=None, out_image=None):
    """src/ImageDataset.hs:47-77 from the loaded arrays on: uvw in metres ([rows,3] or (u,v,w)), `n` = take
    the first n visibilities (Maybe Int of the reference).  Returns (max, image or None, uvgrid or None).
    out_image: a caller-owned [side, side] float64 array to receive the image (e.g. page-locked memory)."""
    ctx = ctx or get_context()
    if isinstance(uvw_m, (tuple, list)):
        u, v, w = (f64(a) for a in uvw_m)
    else:
        m = f64(uvw_m)
        u, v, w = m[:, 0].copy(), m[:, 1].copy(), m[:, 2].copy()
    vis = c128(vis).reshape(-1)
    cnt = vis.size if n is None else int(n)
    u, v, w, vis = u[:cnt].copy(), v[:cnt].copy(), w[:cnt].copy(), vis[:cnt].copy()
    a1, a2 = int64(a1)[:cnt].copy(), int64(a2)[:cnt].copy()
    wkernels, akernels, wbins = c128(wkernels), c128(akernels), f64(wbins)
    nw, qpx, _, s, _ = wkernels.shape
    side = G._grid_side(theta, lam)
    if out_image is not None:
        if out_image.shape != (side, side) or out_image.dtype != np.float64 or not out_image.flags.c_contiguous:
            raise ValueError("out_image must be a C-contiguous float64 array of shape (%d, %d)" % (side, side))
        img = out_image
    else:
        img = np.empty((side, side), np.float64) if want_image else None
    grd = np.empty((side, side), np.complex128) if want_grid else None
    mx = np.empty(1, np.float64)
    ctx.check(ctx.lib.skagrid_aw_gridding(ctx.h, float(theta), int(lam), nw, qpx, s, ptr(wkernels), ptr(wbins), akernels.shape[0],
                                          ptr(akernels), cnt, ptr(u), ptr(v), ptr(w), ptr(a1), ptr(a2), float(freq), ptr(vis),
                                          ptr(img), ptr(mx), ptr(grd)))
    return float(mx[0]), img, grd


def aw_gridding(wfile, afile, datfile, n=None, outfile=None, old=False):
    """src/ImageDataset.hs:29-83 with `.npz` stand-ins for the three HDF5 files (same dataset paths as keys).
    theta = 0.008 and lam = 300000 are hard-coded as in the reference (:32-33).  Returns the image maximum."""
    theta, lam = 0.008, 300000
    dat = np.load(datfile)
    vis = c128(dat["/vis/vis"]).reshape(-1)
    uvw = f64(dat["/vis/uvw"])
    a1, a2 = int64(dat["/vis/antenna1"]), int64(dat["/vis/antenna2"])
    ts, fs = f64(dat["/vis/time"]), f64(dat["/vis/frequency"])
    akernels = getAKernels(np.load(afile), theta, float(ts[0]), float(fs[0]))
    wkernels, wbins = getWKernels(np.load(wfile), theta)
    mx, img, _ = aw_gridding_arrays(theta, lam, wkernels, wbins, akernels, uvw, a1, a2, float(fs[0]), vis, n=n, want_image=outfile is not None)
    if outfile is not None:
        np.savez(outfile, **{"/img": img})
    return mx
