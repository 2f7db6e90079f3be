"""ctypes binding of libskagrid.so (include/skagrid.h) -- the stub a maintainer of the reference would
write for `foreign import ccall`, in Python.  Loading fails loudly when the library is missing: there
is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libskagrid.so")

i64 = C.c_int64
dbl = C.c_double
vp = C.c_void_p
ip = C.c_int


class Geom(C.Structure):
    """skagrid_geom"""
    _fields_ = [(n, i64) for n in ("height", "width", "row0", "row1", "nw", "qpx", "gh", "gw")]


# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGS = {
    "skagrid_create": [ip, C.POINTER(vp)],
    "skagrid_destroy": [vp],
    "skagrid_last_error": [vp],
    "skagrid_version": [],
    "skagrid_grid_side": [dbl, i64],
    "skagrid_last_device_ms": [vp],
    "skagrid_launch_count": [vp],
    "skagrid_resident_grid": [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)],
    "skagrid_measure_fp64_tflops": [vp, C.POINTER(dbl)],
    "skagrid_measure_l2_read_tbs": [vp, i64, C.POINTER(dbl)],
    "skagrid_measure_l2_pattern_tbs": [vp, i64, ip, C.POINTER(dbl)],
    "skagrid_frac_coord": [vp, i64, i64, i64, vp, vp, vp, ip],
    "skagrid_frac_coords": [vp, i64, i64, i64, i64, vp, vp, vp, vp, vp, vp, ip],
    "skagrid_find_closest": [vp, i64, vp, i64, vp, vp],
    "skagrid_uvw_lambda": [vp, dbl, i64, vp, vp, vp],
    "skagrid_mirror_uvw": [vp, i64, vp, vp, vp, vp],
    "skagrid_doweight": [vp, dbl, i64, i64, vp, vp, vp],
    "skagrid_grid": [vp, i64, i64, vp, i64, vp, vp, vp],
    "skagrid_convgrid": [vp, i64, i64, i64, vp, i64, i64, vp, i64, vp, vp, vp],
    "skagrid_convgrid2": [vp, i64, i64, i64, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp],
    "skagrid_convgrid_aw": [vp, i64, i64, i64, vp, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp, vp, vp],
    "skagrid_convdegrid": [vp, i64, i64, i64, vp, i64, i64, vp, i64, vp, vp, vp],
    "skagrid_convdegrid2": [vp, i64, i64, i64, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp],
    "skagrid_convdegrid_aw": [vp, i64, i64, i64, vp, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp, vp, vp],
    "skagrid_convolve2d": [vp, i64, vp, vp, vp],
    "skagrid_aw_kernel": [vp, i64, i64, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp],
    "skagrid_make_grid_hermitian": [vp, i64, vp, vp],
    "skagrid_ifft": [vp, i64, vp, vp],
    "skagrid_fft": [vp, i64, vp, vp],
    "skagrid_grid_to_image": [vp, i64, vp, vp, vp],
    "skagrid_simple_imaging": [vp, dbl, i64, i64, vp, vp, vp, vp, vp],
    "skagrid_conv_imaging": [vp, i64, i64, i64, vp, dbl, i64, i64, vp, vp, vp, vp, vp],
    "skagrid_conv_imaging2": [vp, i64, i64, i64, i64, vp, dbl, i64, i64, vp, vp, vp, vp, vp, vp],
    "skagrid_aw_imaging": [vp, dbl, i64, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp],
    "skagrid_aw_gridding": [vp, dbl, i64, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp, dbl, vp, vp, vp, vp],
    "skagrid_w_kernels": [vp, dbl, i64, vp, i64, i64, i64, ip, vp],
    "skagrid_w_kernels_ex": [vp, dbl, i64, vp, i64, i64, i64, ip, vp, dbl, dbl, vp],
    "skagrid_convgrid2_mgpu_vis": [C.POINTER(vp), ip, i64, i64, i64, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp],
    "skagrid_convdegrid2_mgpu_vis": [C.POINTER(vp), ip, i64, i64, i64, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp],
    "skagrid_convgrid2_mgpu_tile": [C.POINTER(vp), ip, i64, i64, i64, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp, vp],
    "skagrid_convdegrid2_mgpu_tile": [C.POINTER(vp), ip, i64, i64, i64, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp, vp],
    "skagrid_dev_plan_create": [vp, C.POINTER(Geom), i64, vp, vp, vp, vp, ip, vp, C.POINTER(vp)],
    "skagrid_dev_plan_destroy": [vp, vp],
    "skagrid_dev_plan_update": [vp, vp, i64, vp, vp, vp, vp, vp],
    "skagrid_dev_plan_alloc": [vp, C.POINTER(Geom), i64, ip, C.POINTER(vp)],
    "skagrid_dev_plan_update_packed": [vp, vp, i64, vp, ip, vp],
    "skagrid_dev_row_hist": [vp, i64, i64, i64, i64, vp, vp, vp],
    "skagrid_dev_route_count": [vp, i64, i64, i64, ip, vp, i64, vp, vp, vp],
    "skagrid_dev_route_pack": [vp, i64, i64, i64, ip, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp],
    "skagrid_dev_scatter_add": [vp, i64, vp, vp, vp, vp],
    "skagrid_ipc_alloc": [vp, i64, C.POINTER(vp), vp],
    "skagrid_ipc_free": [vp, vp],
    "skagrid_ipc_open": [vp, vp, C.POINTER(vp)],
    "skagrid_ipc_close": [vp, vp],
    "skagrid_dev_peer_sum": [vp, ip, vp, vp, i64, ip, vp],
    "skagrid_dev_peer_barrier": [vp, ip, ip, vp, C.c_uint32, vp],
    "skagrid_dev_peer_copy": [vp, vp, vp, i64, vp],
    "skagrid_dev_peer_gather": [vp, ip, vp, vp, vp, ip, vp],
    "skagrid_dev_peer_gather2d": [vp, ip, vp, i64, vp, vp, i64, vp, ip, vp],
    "skagrid_dev_peer_copy2d": [vp, vp, i64, vp, i64, i64, i64, vp],
    "skagrid_dev_plan_set_vis": [vp, vp, vp, ip, vp],
    "skagrid_dev_plan_order": [vp, vp, vp, vp],
    "skagrid_dev_plan_stats": [vp, vp, vp, C.POINTER(i64 * 5)],
    "skagrid_dev_grid": [vp, vp, vp, vp, ip, vp],
    "skagrid_dev_degrid": [vp, vp, vp, vp, vp, vp],
    "skagrid_dev_degrid_plan_order": [vp, vp, vp, vp, vp, vp],
    "skagrid_dev_grid_to_image": [vp, i64, vp, vp, vp, vp],
    "skagrid_dev_synth_vis": [vp, C.c_uint64, i64, i64, i64, i64, i64, ip, vp, vp, vp, vp, vp],
    "skagrid_dev_w_kernels": [vp, dbl, i64, vp, i64, i64, i64, ip, vp, vp],
    "skagrid_dev_frac_coord": [vp, i64, i64, i64, vp, vp, vp, ip, vp],
    "skagrid_dev_uvw_scale": [vp, i64, vp, vp, vp, dbl, ip, vp],
    "skagrid_dev_mirror_uvw": [vp, i64, vp, vp, vp, vp, vp],
    "skagrid_dev_find_closest": [vp, i64, vp, i64, vp, vp, vp],
    "skagrid_dev_doweight": [vp, dbl, i64, i64, vp, vp, vp, vp],
    "skagrid_dev_slab_fft_rows": [vp, i64, i64, i64, vp, vp],
    "skagrid_dev_slab_fft_cols": [vp, i64, i64, i64, vp, vp, vp, vp],
    "skagrid_dev_weight_count": [vp, dbl, i64, i64, vp, vp, vp, vp],
    "skagrid_dev_weight_apply": [vp, dbl, i64, i64, vp, vp, vp, vp, vp],
    "skagrid_dev_take_error": [vp, vp, C.POINTER(C.c_int)],
}
_RESTYPES = {
    "skagrid_destroy": None,
    "skagrid_last_error": C.c_char_p,
    "skagrid_version": C.c_char_p,
    "skagrid_last_device_ms": dbl,
    "skagrid_launch_count": i64,
    "skagrid_grid_side": i64,
    "skagrid_dev_plan_destroy": None,
}

EXPORTS = tuple(_SIGS)
_lib = None


class SkagridError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libskagrid error {code}: {msg}")
        self.code = code


def load():
    """dlopen libskagrid.so and declare every prototype.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ska_sdp_accelerate_gridding_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = lib
    return _lib
