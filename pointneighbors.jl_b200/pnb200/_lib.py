"""ctypes binding of libpnb200.so (include/pnb200.h).  One Python function per C entry point.

The library is the product: if it cannot be loaded, or there is no CUDA device, every compute
call raises -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

PNB_OK, PNB_ERR_DOMAIN, PNB_ERR_ARG, PNB_ERR_LIST_FULL, PNB_ERR_BOUNDS, PNB_ERR_CUDA, \
    PNB_ERR_STATE = range(7)


class ArgumentError(ValueError):
    """Julia's ArgumentError (constructor validation)."""


class PointNeighborsError(RuntimeError):
    """Julia's ErrorException raised by `error(...)` in the reference."""


class BoundsError(IndexError):
    """Julia's BoundsError (safe sweep leaving the grid)."""


class CudaError(RuntimeError):
    pass


class WcsphParams(C.Structure):
    _fields_ = [("smoothing_length", C.c_float), ("sound_speed", C.c_float),
                ("alpha", C.c_float), ("beta", C.c_float), ("epsilon", C.c_float),
                ("delta", C.c_float), ("kernel_norm", C.c_float)]


class WcsphParams64(C.Structure):
    """pnb_wcsph_params_f64 (include/pnb200.h)."""
    _fields_ = [("smoothing_length", C.c_double), ("sound_speed", C.c_double), ("alpha", C.c_double),
                ("beta", C.c_double), ("epsilon", C.c_double), ("delta", C.c_double),
                ("kernel_norm", C.c_double)]


class TlsphParams(C.Structure):
    """pnb_tlsph_params (include/pnb200.h)."""
    _fields_ = [("smoothing_length", C.c_float), ("kernel_norm", C.c_float),
                ("young_modulus", C.c_float), ("penalty_alpha", C.c_float)]


class SlabArrays(C.Structure):
    """pnb_slab_arrays (include/pnb200.h)."""
    _fields_ = [("ptr", C.c_void_p * 8), ("width", C.c_int32 * 8), ("n_arrays", C.c_int32)]


_lib = None

_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
_pf = C.POINTER(C.c_float)
_pd = C.POINTER(C.c_double)
_f64 = C.c_double
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol declared in include/pnb200.h
SIGNATURES = {
    "pnb_version": (C.c_int, []),
    "pnb_last_error": (C.c_char_p, []),
    "pnb_device_count": (C.c_int, []),
    "pnb_grid_params_f32": (C.c_int, [C.c_int, _f32, _pf, _pf, _pf, _pf, _pf, _pf, _pi64, _pi64, _pf]),
    "pnb_grid_create_f32": (C.c_int, [C.c_int, _f32, _pf, _pf, _pf, _pf, C.POINTER(_vp)]),
    "pnb_grid_create_window_f32": (C.c_int, [C.c_int, _f32, _pf, _pf, _pf, _pf, _pi64, _pi64,
                                             C.POINTER(_vp)]),
    "pnb_slab_classify_f32": (C.c_int, [_vp, _i64, C.c_int, _f32, _f32, _i64, _i64, C.c_int, C.c_int,
                                        _vp, _vp, _vp, _i64, _vp, _pi64, _vp]),
    "pnb_slab_pack_f32": (C.c_int, [C.POINTER(SlabArrays), _i64, C.c_int, _f32, _f32, _i64, _i64,
                                    C.c_int, C.c_int, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _pi64, _vp]),
    "pnb_slab_unpack_f32": (C.c_int, [C.POINTER(SlabArrays), _i64, C.c_int, _f32, _f32, _i64, _i64,
                                      C.c_int, C.c_int, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64,
                                      _vp, _i64, _vp, _pi64, _vp]),
    "pnb_grid_params_f64": (C.c_int, [C.c_int, _f64, _pd, _pd, _pd, _pd, _pd, _pd, _pi64, _pi64, _pd]),
    "pnb_grid_create_f64": (C.c_int, [C.c_int, _f64, _pd, _pd, _pd, _pd, C.POINTER(_vp)]),
    "pnb_grid_create_padded_f64": (C.c_int, [C.c_int, _f64, _pd, _pd, _pd, _pd, C.POINTER(_vp)]),
    "pnb_grid_create_padded_f32": (C.c_int, [C.c_int, _f32, _pf, _pf, _pf, _pf, C.POINTER(_vp)]),
    "pnb_grid_params_mixed": (C.c_int, [C.c_int, _f32, _pd, _pd, _pf, _pf, _pd, _pd, _pi64, _pi64, _pf]),
    "pnb_grid_create_mixed": (C.c_int, [C.c_int, _f32, _pd, _pd, _pf, _pf, C.POINTER(_vp)]),
    "pnb_grid_create_padded_mixed": (C.c_int, [C.c_int, _f32, _pd, _pd, _pf, _pf, C.POINTER(_vp)]),
    "pnb_grid_build_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, _vp]),
    "pnb_point_cells_f64": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "pnb_count_neighbors_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, C.c_int, _vp, _vp]),
    "pnb_nbody_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, C.c_int, _vp, _f64, _vp, _vp]),
    "pnb_wcsph_interact_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, C.c_int, _vp, _vp, _vp,
                                         _vp, _vp, _vp, C.POINTER(WcsphParams64), _vp, _vp]),
    "pnb_nbody_mixed": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, C.c_int, _vp, _f32, _vp, _vp]),
    "pnb_wcsph_interact_mixed": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, C.c_int, _vp, _vp, _vp,
                                           _vp, _vp, _vp, C.POINTER(WcsphParams), _vp, _vp]),
    "pnb_nlist_build_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, C.POINTER(_vp), _vp]),
    "pnb_nlist_pairs_f64": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pnb_grid_create_hashed_f32": (C.c_int, [C.c_int, _f32, _i64, _pf, _pf, C.POINTER(_vp)]),
    "pnb_grid_export_hash_table": (C.c_int, [_vp, _vp, _vp, _vp]),
    "pnb_spatial_hash": (_i64, [C.c_int, _pi64, _i64]),
    "pnb_grid_destroy": (None, [_vp]),
    "pnb_grid_total_cells": (_i64, [_vp]),
    "pnb_grid_n_points": (_i64, [_vp]),
    "pnb_grid_layout": (C.c_int, [_vp]),
    "pnb_grid_build_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, _vp]),
    "pnb_grid_build_async_f32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "pnb_grid_check": (C.c_int, [_vp, _vp]),
    "pnb_grid_check_settled": (C.c_int, [_vp]),
    "pnb_wcsph_interact_async_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp,
                                               C.POINTER(WcsphParams), _vp, _vp]),
    "pnb_hoststep_create": (C.c_int, [_vp, _i64, C.POINTER(_vp)]),
    "pnb_hoststep_wcsph_submit": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.POINTER(WcsphParams), _vp]),
    "pnb_hoststep_wait": (C.c_int, [_vp]),
    "pnb_hoststep_set_state_equation": (C.c_int, [_vp, _f32, _f32, _f32, _f32]),
    "pnb_hoststep_host_times": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pnb_hoststep_destroy": (None, [_vp]),
    "pnb_grid_append_f32": (C.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "pnb_wcsph_interact_layers_async_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp,
                                                      C.POINTER(WcsphParams), _vp, C.c_int, C.c_int,
                                                      C.c_int, C.c_int, C.c_int, _vp]),
    "pnb_slab_pack_rows_f32": (C.c_int, [C.POINTER(SlabArrays), _vp, _i64, _vp, _vp]),
    "pnb_slab_append_f32": (C.c_int, [C.POINTER(SlabArrays), _i64, C.c_int, _f32, _f32, _i64, _i64,
                                      _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "pnb_slab_append_strided_f32": (C.c_int, [C.POINTER(SlabArrays), _i64, C.c_int, _f32, _f32, _i64, _i64,
                                              _i64, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp]),
    "pnb_slab_link_row_stride": (C.c_int, [_vp]),
    "pnb_slab_link_set_timeout": (C.c_int, [_vp, C.c_double]),
    "pnb_slab_compact_f32": (C.c_int, [C.POINTER(SlabArrays), _i64, _i64, _vp, _i64, _i64, _vp, _vp,
                                       _vp, _vp]),
    "pnb_slab_link_create": (C.c_int, [_i64, C.c_int, C.POINTER(_vp)]),
    "pnb_slab_link_export": (C.c_int, [_vp, _vp]),
    "pnb_slab_link_connect": (C.c_int, [_vp, _vp, _vp]),
    "pnb_slab_link_connect_local": (C.c_int, [_vp, _vp, _vp]),
    "pnb_slab_link_send": (C.c_int, [_vp, C.POINTER(SlabArrays), _i64, C.c_int, _f32, _f32, _i64, _i64,
                                     _vp, C.c_uint64, _vp]),
    "pnb_slab_link_recv": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp), C.POINTER(_vp),
                                     C.POINTER(C.c_int64), _vp]),
    "pnb_slab_link_destroy": (None, [_vp]),
    "pnb_point_cells_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "pnb_grid_export_csr": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "pnb_grid_export_dvov": (C.c_int, [_vp, _vp, _vp, _i32, C.c_int, _vp]),
    "pnb_count_neighbors_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, C.c_int, _vp, _vp]),
    "pnb_nbody_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, C.c_int, _vp, _f32, _vp, _vp]),
    "pnb_wcsph_interact_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, C.c_int, _vp, _vp,
                                         _vp, _vp, _vp, _vp, C.POINTER(WcsphParams), _vp, _vp]),
    "pnb_nlist_build_f32": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, C.POINTER(_vp), _vp]),
    "pnb_nlist_destroy": (None, [_vp]),
    "pnb_nlist_n_points": (_i64, [_vp]),
    "pnb_nlist_n_pairs": (_i64, [_vp]),
    "pnb_nlist_max_length": (_i64, [_vp]),
    "pnb_nlist_export_csr": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "pnb_nlist_export_dvov": (C.c_int, [_vp, _vp, _vp, _i32, C.c_int, C.c_int, _vp]),
    "pnb_nlist_pairs_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pnb_tlsph_deformation_grad_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32,
                                                 _vp, _vp]),
    "pnb_tlsph_interact_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                         C.POINTER(TlsphParams), _vp, _vp]),
    "pnb_tlsph_pk1_corrected_f32": (C.c_int, [C.c_int, _i64, _vp, _vp, _f32, _f32, _vp, _vp]),
    "pnb_wcsph_compute_pressure_f32": (C.c_int, [C.c_int, _i64, _vp, _f32, _f32, _f32, _f32, _vp,
                                                 _vp]),
    "pnb_malloc": (C.c_int, [C.POINTER(_vp), _i64]),
    "pnb_free": (C.c_int, [_vp]),
    "pnb_malloc_host": (C.c_int, [C.POINTER(_vp), _i64]),
    "pnb_free_host": (C.c_int, [_vp]),
    "pnb_memcpy_h2d": (C.c_int, [_vp, _vp, _i64, _vp]),
    "pnb_memcpy_d2h": (C.c_int, [_vp, _vp, _i64, _vp]),
    "pnb_memset": (C.c_int, [_vp, C.c_int, _i64, _vp]),
    "pnb_stream_synchronize": (C.c_int, [_vp]),
    "pnb_launch_count": (_i64, []),
    "pnb_set_exact_arithmetic": (None, [C.c_int]),
    "pnb_get_exact_arithmetic": (C.c_int, []),
    "pnb_set_tuning": (None, [C.c_int, C.c_int]),
    "pnb_set_build_tuning": (None, [C.c_int]),
    "pnb_set_build_layout": (None, [C.c_int]),
    "pnb_set_bucket_order": (None, [C.c_int]),
    "pnb_set_twoset_tiles": (None, [C.c_int]),
    "pnb_set_sweep_left": (None, [C.c_int]),
    "pnb_set_sweep_kernel": (None, [C.c_int]),
    "pnb_set_sweep_reserve": (None, [C.c_int]),
    "pnb_profile_enable": (None, [C.c_int]),
    "pnb_profile_reset": (None, []),
    "pnb_profile_phases": (C.c_int, []),
    "pnb_profile_name": (C.c_char_p, [C.c_int]),
    "pnb_profile_get": (C.c_int, [C.c_int, C.POINTER(C.c_double), _pi64]),
}


def library_path() -> str:
    return _build.LIB


def lib() -> C.CDLL:
    """Load libpnb200.so (building it first if the sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.is_stale():
        # sources newer than the binary (or no binary): rebuild under a file lock so that the
        # ranks of a torchrun job do not compile into the same objects at the same time.  Fail
        # loudly when that is impossible: there is no silent fallback.
        try:
            _build.build_locked()
        except Exception as exc:  # pragma: no cover - depends on toolchain
            if not os.path.exists(path):
                raise ImportError(
                    f"libpnb200.so is missing and could not be built ({exc}); the CUDA library "
                    "is required, pnb200 has no CPU fallback") from exc
            import warnings
            warnings.warn(f"libpnb200.so is older than its sources and could not be rebuilt "
                          f"({exc}); running the stale binary")
    handle = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def last_error() -> str:
    return lib().pnb_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Translate a pnb_status into the exception the reference would raise."""
    if status == PNB_OK:
        return
    msg = last_error()
    if status == PNB_ERR_ARG:
        raise ArgumentError(msg)
    if status in (PNB_ERR_DOMAIN, PNB_ERR_LIST_FULL, PNB_ERR_STATE):
        raise PointNeighborsError(msg)
    if status == PNB_ERR_BOUNDS:
        raise BoundsError(msg)
    raise CudaError(msg)


def profile(enable=None, reset=False):
    """Per-kernel device times {name: (total_ms, launches)} accumulated since the last reset."""
    L = lib()
    if enable is not None:
        L.pnb_profile_enable(int(bool(enable)))
    out = {}
    for ph in range(L.pnb_profile_phases()):
        ms, n = C.c_double(), C.c_int64()
        check(L.pnb_profile_get(ph, C.byref(ms), C.byref(n)))
        out[L.pnb_profile_name(ph).decode()] = (ms.value, n.value)
    if reset:
        L.pnb_profile_reset()
    return out
