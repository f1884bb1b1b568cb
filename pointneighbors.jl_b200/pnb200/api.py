"""Host-side mirror of the PointNeighbors.jl interface for the hot path, on top of the C ABI.

The reference is Julia and there is no Julia toolchain in this image, so the host side is this
Python module; julia/PNB200.jl carries the same calls as `ccall`s.  Names,
keyword arguments, defaults, return values and error texts follow the reference:

    GridNeighborhoodSearch{NDIMS}(; ...)        ->  GridNeighborhoodSearch[NDIMS](...)      src/nhs_grid.jl:77-129
    FullGridCellList(; ...)                      ->  FullGridCellList(...)                   src/cell_lists/full_grid.jl:48-82
    PeriodicBox(; min_corner, max_corner)        ->  PeriodicBox(...)                        src/neighborhood_search.jl:129-143
    PrecomputedNeighborhoodSearch{NDIMS}(; ...)  ->  PrecomputedNeighborhoodSearch[NDIMS](...)  src/nhs_precomputed.jl:89-111
    initialize!(nhs, x, y; eachindex_y)          ->  initialize_(nhs, x, y, eachindex_y=)    src/nhs_grid.jl:220-225
    update!(nhs, x, y; points_moving, ...)       ->  update_(nhs, x, y, points_moving=)      src/nhs_grid.jl:283-292
    foreach_point_neighbor(f, x, y, nhs; points) ->  same                                    src/neighborhood_search.jl:183-201
    copy_neighborhood_search / freeze_neighborhood_search / requires_update / search_radius / ndims

Coordinates are torch CUDA tensors of shape (N, NDIMS), float32, contiguous: byte for byte the
memory of Julia's NDIMS x N column-major matrix.  Point indices are 0-based on this side (the
Julia glue passes index_base = 1).  torch is only the owner of device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (ArgumentError, BoundsError, PointNeighborsError, TlsphParams, WcsphParams,
                   check)

__all__ = [
    "foreach_neighbor", "foreach_neighbor_unsafe", "mapreduce_neighbor", "mapreduce_neighbor_unsafe",
    "foreach_point_neighbor_unsafe", "initialize_grid_", "update_grid_", "default_backend", "B200Backend",
    "ArgumentError", "PointNeighborsError", "BoundsError",
    "ParallelUpdate", "SerialUpdate", "ParallelIncrementalUpdate", "SemiParallelUpdate",
    "SerialIncrementalUpdate", "DynamicVectorOfVectors",
    "PeriodicBox", "FullGridCellList", "SpatialHashingCellList", "spatial_hash",
    "GridNeighborhoodSearch", "PrecomputedNeighborhoodSearch",
    "initialize_", "update_", "check_", "HostStepper", "initialize", "update", "foreach_point_neighbor",
    "copy_neighborhood_search", "freeze_neighborhood_search", "requires_update",
    "search_radius", "ndims", "CountNeighbors", "NBodyGravity", "WCSPHInteract",
    "TLSPHDeformationGradient", "TLSPHInteract", "compute_pk1_corrected_", "compute_pressure_",
    "wendland_c2_norm", "set_exact_arithmetic",
]

_EPS64 = 2.220446049250313e-16


def _torch():
    import torch
    return torch


# ---------------------------------------------------------------------------------------------
# update strategies (src/nhs_grid.jl:135-192).  The device build is a full counting-sort rebuild,
# i.e. ParallelUpdate semantics; SerialUpdate is the same rebuild in the reference
# (nhs_grid.jl:470-477).  The incremental strategies are CPU algorithms and not offered.
# ---------------------------------------------------------------------------------------------
class _Strategy:
    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return f"{type(self).__name__}()"


class ParallelUpdate(_Strategy):
    pass


class SerialUpdate(_Strategy):
    pass


class ParallelIncrementalUpdate(_Strategy):
    pass


class SemiParallelUpdate(_Strategy):
    pass


class SerialIncrementalUpdate(_Strategy):
    pass


class DynamicVectorOfVectors:
    """Type tag of the reference's list storage (src/vector_of_vectors.jl:3-31); `backend=` keyword."""

    def __init__(self, eltype=np.int32):
        self.eltype = eltype

    def __class_getitem__(cls, eltype):
        return cls(eltype)


class _ParametricBase(type):
    """`GridNeighborhoodSearch[3](...)` stands for Julia's `GridNeighborhoodSearch{3}(; ...)`."""

    def __getitem__(cls, ndims_):
        def ctor(**kwargs):
            return cls(int(ndims_), **kwargs)
        ctor.__name__ = f"{cls.__name__}[{ndims_}]"
        return ctor


def _as_real_vector(v, what):
    a = np.asarray(v)
    if a.ndim != 1:
        a = a.reshape(-1)
    return a


def _is_integer_scalar(v) -> bool:
    return isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_))


def _eltype_of(v):
    if isinstance(v, np.floating):
        return np.dtype(type(v))
    if isinstance(v, float):
        return np.dtype(np.float64)
    a = np.asarray(v)
    return a.dtype


class PeriodicBox:
    """PeriodicBox(; min_corner, max_corner)  (src/neighborhood_search.jl:129-143)."""

    def __init__(self, *, min_corner, max_corner):
        mn = np.asarray(min_corner)
        mx = np.asarray(max_corner)
        dt = np.result_type(mn, mx)
        if not np.issubdtype(dt, np.floating):
            dt = np.dtype(np.float64)
        self.min_corner = mn.astype(dt)
        self.max_corner = mx.astype(dt)
        self.size = self.max_corner - self.min_corner   # evaluated in the element type
        self.eltype = np.dtype(dt)

    def __len__(self):
        return self.min_corner.size


class FullGridCellList:
    """FullGridCellList(; min_corner, max_corner, search_radius, backend, max_points_per_cell)
    (src/cell_lists/full_grid.jl:48-82).  Stores the PADDED corners like the reference."""

    def __init__(self, *, min_corner, max_corner, search_radius=None,
                 backend=DynamicVectorOfVectors[np.int32], max_points_per_cell: int = 100,
                 mixed_precision: bool = False):
        """mixed_precision=True: Float64 corners / coordinates with a Float32 `search_radius`
        (docs/literate/src/tut_gpu_usage.jl:45-50).  Julia infers this from the argument types;
        here it is explicit because Python floats and arrays carry no such intent."""
        mn = _as_real_vector(min_corner, "min_corner")
        mx = _as_real_vector(max_corner, "max_corner")
        if mn.size != mx.size:
            raise ArgumentError("min_corner and max_corner must have the same length")
        if mn.size > 100:
            raise ArgumentError("FullGridCellList only supports up to 100 dimensions, "
                                "check your `min_corner` and `max_corner`")
        if search_radius is None:
            dt = mn.dtype if np.issubdtype(mn.dtype, np.floating) else np.dtype(np.float64)
            search_radius = dt.type(0)
        if not isinstance(backend, DynamicVectorOfVectors):
            raise ArgumentError("only the DynamicVectorOfVectors backend is GPU-compatible "
                                "(src/cell_lists/full_grid.jl:21-25)")
        self.backend = backend
        self.max_points_per_cell = int(max_points_per_cell)
        self.search_radius = search_radius
        self._ndims = int(mn.size)
        # Float64 search: corners AND radius given as Float64 (np.float64 scalar / float64 arrays),
        # like the reference, whose element type follows its arguments.  Everything else runs the
        # Float32 path (python floats are taken as Float32 values, as before).
        self.eltype = np.dtype(np.float64) if (isinstance(search_radius, np.float64)
                                                and mn.dtype == np.float64
                                                and mx.dtype == np.float64) else np.dtype(np.float32)
        self.mixed = bool(mixed_precision)
        if self.mixed:
            if self._ndims > 3:
                raise ArgumentError("`NDIMS` must be 1, 2, or 3")
            if isinstance(search_radius, np.float64):
                raise ArgumentError("mixed precision means a Float32 `search_radius` with Float64 "
                                    "corners; pass np.float32")
            self.eltype = np.dtype(np.float64)          # element type of corners / coordinates
            pmin64, pmax64, gsz64 = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int64 * 3)()
            mn64 = np.ascontiguousarray(mn, dtype=np.float64)
            mx64 = np.ascontiguousarray(mx, dtype=np.float64)
            check(_lib.lib().pnb_grid_params_mixed(
                self._ndims, np.float32(search_radius), mn64.ctypes.data_as(_lib._pd),
                mx64.ctypes.data_as(_lib._pd), None, None, pmin64, pmax64, gsz64, None, None))
            self.min_corner = np.array(pmin64[:self._ndims], dtype=np.float64)
            self.max_corner = np.array(pmax64[:self._ndims], dtype=np.float64)
            self.n_cells_per_dimension = tuple(int(v) for v in gsz64[:self._ndims])
            self._user_min, self._user_max = mn64, mx64
            return
        if self.eltype == np.float64:
            if self._ndims > 3:
                raise ArgumentError("`NDIMS` must be 1, 2, or 3")
            pmin64, pmax64, gsz64 = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int64 * 3)()
            mn64 = np.ascontiguousarray(mn, dtype=np.float64)
            mx64 = np.ascontiguousarray(mx, dtype=np.float64)
            check(_lib.lib().pnb_grid_params_f64(
                self._ndims, float(search_radius), mn64.ctypes.data_as(_lib._pd),
                mx64.ctypes.data_as(_lib._pd), None, None, pmin64, pmax64, gsz64, None, None))
            self.min_corner = np.array(pmin64[:self._ndims], dtype=np.float64)
            self.max_corner = np.array(pmax64[:self._ndims], dtype=np.float64)
            self.n_cells_per_dimension = tuple(int(v) for v in gsz64[:self._ndims])
            self._user_min, self._user_max = mn64, mx64
            return
        r = np.float32(search_radius)
        # padding and grid size through the library's host arithmetic (bit-identical to Julia)
        pmin = (C.c_float * 3)()
        pmax = (C.c_float * 3)()
        gsz = (C.c_int64 * 3)()
        mn32 = np.ascontiguousarray(mn, dtype=np.float32)
        mx32 = np.ascontiguousarray(mx, dtype=np.float32)
        if self._ndims > 3:
            raise ArgumentError("`NDIMS` must be 1, 2, or 3")
        check(_lib.lib().pnb_grid_params_f32(
            self._ndims, r, mn32.ctypes.data_as(_lib._pf), mx32.ctypes.data_as(_lib._pf), None,
            None, pmin, pmax, gsz, None, None))
        self.min_corner = np.array(pmin[:self._ndims], dtype=np.float32)
        self.max_corner = np.array(pmax[:self._ndims], dtype=np.float32)
        self.n_cells_per_dimension = tuple(int(v) for v in gsz[:self._ndims])
        # what was passed in, needed to re-create the device grid
        self._user_min = mn32
        self._user_max = mx32

    def ndims(self):
        return self._ndims

    @property
    def is_template(self):
        return float(self.search_radius) < _EPS64


def spatial_hash(cell, list_size) -> int:
    """spatial_hash(cell, list_size)  (src/cell_lists/spatial_hashing.jl:159-174), 1-based like the
    reference (the device table uses key - 1)."""
    c = (C.c_int64 * 3)(*([int(v) for v in cell] + [0] * (3 - len(cell))))
    k = int(_lib.lib().pnb_spatial_hash(len(cell), c, int(list_size)))
    if k < 0:
        raise ArgumentError("spatial_hash takes a 1-, 2- or 3-tuple and a positive list_size")
    return k + 1


class SpatialHashingCellList(metaclass=_ParametricBase):
    """SpatialHashingCellList{NDIMS}(; list_size, backend, max_points_per_cell)
    (src/cell_lists/spatial_hashing.jl:24-61): an unbounded domain hashed into `list_size` lists
    (about 2 * n_points is recommended, :9-10).  `SpatialHashingCellList[3](list_size=...)`."""

    def __init__(self, ndims_: int, *, list_size: int,
                 backend=DynamicVectorOfVectors[np.int32], max_points_per_cell: int = 100):
        if not isinstance(backend, DynamicVectorOfVectors):
            raise ArgumentError("only the DynamicVectorOfVectors backend is GPU-compatible "
                                "(src/cell_lists/spatial_hashing.jl:42-49)")
        if int(ndims_) not in (1, 2, 3):
            raise ArgumentError("`NDIMS` must be 1, 2, or 3")
        if int(list_size) < 1:
            raise ArgumentError("`list_size` must be positive")
        self._ndims = int(ndims_)
        self.list_size = int(list_size)
        self.backend = backend
        self.max_points_per_cell = int(max_points_per_cell)
        self.eltype = np.dtype(np.float32)
        self.search_radius = np.float32(0)
        self._user_min = np.zeros(3, dtype=np.float32)   # no corners: the domain is unbounded
        self._user_max = np.zeros(3, dtype=np.float32)

    def ndims(self):
        return self._ndims


def supported_update_strategies(cell_list):
    # src/cell_lists/full_grid.jl:39-42 lists five strategies for the CPU; on the device the two
    # full-rebuild strategies exist.
    return (ParallelUpdate, SerialUpdate)


def copy_cell_list(cell_list, search_radius, periodic_box):
    """src/cell_lists/full_grid.jl:179-185 -- re-pads from the STORED corners;
    src/cell_lists/spatial_hashing.jl:129-139 -- same list_size, backend and capacity."""
    if isinstance(cell_list, SpatialHashingCellList):
        return SpatialHashingCellList(cell_list._ndims, list_size=cell_list.list_size,
                                      backend=cell_list.backend,
                                      max_points_per_cell=cell_list.max_points_per_cell)
    return FullGridCellList(min_corner=cell_list.min_corner, max_corner=cell_list.max_corner,
                            search_radius=search_radius, backend=cell_list.backend,
                            max_points_per_cell=cell_list.max_points_per_cell,
                            mixed_precision=getattr(cell_list, "mixed", False))


_Parametric = _ParametricBase


class GridNeighborhoodSearch(metaclass=_Parametric):
    """GridNeighborhoodSearch{NDIMS}(; search_radius, n_points, periodic_box, cell_list,
    update_strategy)  (src/nhs_grid.jl:67-129) backed by a device CSR cell list."""

    def __init__(self, ndims_: int, *, search_radius=0.0, n_points: int = 0, periodic_box=None,
                 cell_list: Optional[FullGridCellList] = None, update_strategy=None):
        self._ndims = int(ndims_)
        if cell_list is None:
            raise ArgumentError("the default DictionaryCellList is not GPU-compatible "
                                "(src/cell_lists/dictionary.jl:8-10); pass a FullGridCellList "
                                "or a SpatialHashingCellList")
        if cell_list.ndims() != self._ndims:
            raise ArgumentError(f"a {self._ndims}D cell list is required for "
                                f"a GridNeighborhoodSearch{{{self._ndims}}}")
        if update_strategy is None:
            update_strategy = supported_update_strategies(cell_list)[0]()
        elif type(update_strategy) not in supported_update_strategies(cell_list):
            names = ", ".join(s.__name__ for s in supported_update_strategies(cell_list))
            raise ArgumentError(f"{update_strategy} is not a valid update strategy for "
                                f"this cell list. Available options are ({names})")
        if _is_integer_scalar(search_radius):
            raise ArgumentError("`search_radius` cannot be an integer type, since computed "
                                "distances will be converted to this type")
        self.cell_list = cell_list
        self.search_radius = search_radius
        self.periodic_box = periodic_box
        self.update_strategy = update_strategy
        self.update_buffer = None          # nhs_grid.jl:195-197
        self.n_points = int(n_points)
        is_template = float(search_radius) < _EPS64
        if is_template or periodic_box is None:
            self.n_cells = tuple(-1 for _ in range(self._ndims))
            self.cell_size = tuple(search_radius for _ in range(self._ndims))
        else:
            if _eltype_of(search_radius) != periodic_box.eltype:
                raise ArgumentError("the `search_radius` and the `PeriodicBox` must have "
                                    "the same element type")
            if getattr(cell_list, "mixed", False):
                if _eltype_of(search_radius) != np.dtype(np.float32):
                    raise ArgumentError("mixed precision needs a Float32 `search_radius`")
                nc = (C.c_int64 * 3)()
                cs = (C.c_float * 3)()
                bmn = np.ascontiguousarray(periodic_box.min_corner, dtype=np.float32)
                bmx = np.ascontiguousarray(periodic_box.max_corner, dtype=np.float32)
                check(_lib.lib().pnb_grid_params_mixed(
                    self._ndims, np.float32(search_radius),
                    cell_list._user_min.ctypes.data_as(_lib._pd),
                    cell_list._user_max.ctypes.data_as(_lib._pd),
                    bmn.ctypes.data_as(_lib._pf), bmx.ctypes.data_as(_lib._pf),
                    None, None, None, nc, cs))
                self.n_cells = tuple(int(v) for v in nc[:self._ndims])
                self.cell_size = tuple(np.float32(v) for v in cs[:self._ndims])
                self._handle = None
                self._window = None
                self._cell_list_radius = cell_list.search_radius
                self.eltype = np.dtype(np.float64)
                return
            if _eltype_of(search_radius) == np.dtype(np.float64) and \
                    cell_list.eltype == np.float64:
                nc = (C.c_int64 * 3)()
                cs64 = (C.c_double * 3)()
                bmn = np.ascontiguousarray(periodic_box.min_corner, dtype=np.float64)
                bmx = np.ascontiguousarray(periodic_box.max_corner, dtype=np.float64)
                check(_lib.lib().pnb_grid_params_f64(
                    self._ndims, float(search_radius),
                    cell_list._user_min.ctypes.data_as(_lib._pd),
                    cell_list._user_max.ctypes.data_as(_lib._pd),
                    bmn.ctypes.data_as(_lib._pd), bmx.ctypes.data_as(_lib._pd),
                    None, None, None, nc, cs64))
                self.n_cells = tuple(int(v) for v in nc[:self._ndims])
                self.cell_size = tuple(np.float64(v) for v in cs64[:self._ndims])
                self._handle = None
                self._window = None
                self._cell_list_radius = cell_list.search_radius
                self.eltype = np.dtype(np.float64)
                return
            if _eltype_of(search_radius) != np.dtype(np.float32):
                raise ArgumentError("the B200 path computes in Float32: pass a Float32 "
                                    "`search_radius` and `PeriodicBox`")
            nc = (C.c_int64 * 3)()
            cs = (C.c_float * 3)()
            bmn = np.ascontiguousarray(periodic_box.min_corner, dtype=np.float32)
            bmx = np.ascontiguousarray(periodic_box.max_corner, dtype=np.float32)
            check(_lib.lib().pnb_grid_params_f32(
                self._ndims, np.float32(search_radius),
                cell_list._user_min.ctypes.data_as(_lib._pf),
                cell_list._user_max.ctypes.data_as(_lib._pf),
                bmn.ctypes.data_as(_lib._pf), bmx.ctypes.data_as(_lib._pf),
                None, None, None, nc, cs))
            self.n_cells = tuple(int(v) for v in nc[:self._ndims])
            self.cell_size = tuple(np.float32(v) for v in cs[:self._ndims])
        self._handle = None
        self._window = None    # (lo, hi) global cell window for slab decomposition (slabs.py)
        self._cell_list_radius = cell_list.search_radius
        # Float64 search: a Float64 cell list with a Float64 radius (np.float64)
        self.eltype = np.dtype(np.float64) if (cell_list.eltype == np.float64 and
                                               (isinstance(search_radius, np.float64) or
                                                getattr(cell_list, "mixed", False))) \
            else np.dtype(np.float32)

    # -- device handle -----------------------------------------------------------------------
    def _grid(self):
        if self._handle is None and isinstance(self.cell_list, FullGridCellList) and \
                float(self.search_radius) >= _EPS64 and \
                float(self._cell_list_radius) != float(self.search_radius):
            # The reference keeps the cell list's own grid (padded and sized with ITS radius,
            # src/cell_lists/full_grid.jl:66-78) and only takes cell_size from the search; with
            # two different radii its sweeps index cells that do not exist (BoundsError / domain
            # error).  The device grid is derived from one radius, so refuse instead of silently
            # working on another grid than `cell_list` describes.
            raise ArgumentError(
                f"the cell list was created with search_radius = {self._cell_list_radius}, the "
                f"neighborhood search with {self.search_radius}: create the cell list with the "
                "search's radius, or pass a template and use copy_neighborhood_search "
                "(src/nhs_grid.jl:640-649)")
        if self._handle is None and self.eltype == np.float64:
            if self._window is not None:
                raise ArgumentError("slab windows exist for Float32 searches only")
            cl = self.cell_list
            h = C.c_void_p()
            bmn = bmx = None
            if getattr(cl, "mixed", False):
                if self.periodic_box is not None:
                    bmn_a = np.ascontiguousarray(self.periodic_box.min_corner, dtype=np.float32)
                    bmx_a = np.ascontiguousarray(self.periodic_box.max_corner, dtype=np.float32)
                    bmn, bmx = bmn_a.ctypes.data_as(_lib._pf), bmx_a.ctypes.data_as(_lib._pf)
                check(_lib.lib().pnb_grid_create_mixed(
                    self._ndims, np.float32(self.search_radius),
                    cl._user_min.ctypes.data_as(_lib._pd), cl._user_max.ctypes.data_as(_lib._pd),
                    bmn, bmx, C.byref(h)))
                self._handle = h
                return self._handle
            if self.periodic_box is not None:
                bmn_a = np.ascontiguousarray(self.periodic_box.min_corner, dtype=np.float64)
                bmx_a = np.ascontiguousarray(self.periodic_box.max_corner, dtype=np.float64)
                bmn, bmx = bmn_a.ctypes.data_as(_lib._pd), bmx_a.ctypes.data_as(_lib._pd)
            check(_lib.lib().pnb_grid_create_f64(
                self._ndims, float(self.search_radius), cl._user_min.ctypes.data_as(_lib._pd),
                cl._user_max.ctypes.data_as(_lib._pd), bmn, bmx, C.byref(h)))
            self._handle = h
        if self._handle is None:
            if float(self.search_radius) >= _EPS64 and \
                    _eltype_of(self.search_radius) != np.dtype(np.float32):
                raise ArgumentError("the B200 path computes in Float32: pass a Float32 "
                                    "`search_radius` (src/nhs_grid.jl:60-65)")
            cl = self.cell_list
            h = C.c_void_p()
            if isinstance(cl, SpatialHashingCellList):
                if self._window is not None:
                    raise ArgumentError("slab windows need a FullGridCellList")
                bmn = bmx = None
                if self.periodic_box is not None:
                    bmn_a = np.ascontiguousarray(self.periodic_box.min_corner, dtype=np.float32)
                    bmx_a = np.ascontiguousarray(self.periodic_box.max_corner, dtype=np.float32)
                    bmn, bmx = bmn_a.ctypes.data_as(_lib._pf), bmx_a.ctypes.data_as(_lib._pf)
                check(_lib.lib().pnb_grid_create_hashed_f32(
                    self._ndims, np.float32(self.search_radius), cl.list_size, bmn, bmx,
                    C.byref(h)))
                self._handle = h
                return self._handle
            bmn = bmx = None
            if self.periodic_box is not None:
                bmn_a = np.ascontiguousarray(self.periodic_box.min_corner, dtype=np.float32)
                bmx_a = np.ascontiguousarray(self.periodic_box.max_corner, dtype=np.float32)
                bmn, bmx = bmn_a.ctypes.data_as(_lib._pf), bmx_a.ctypes.data_as(_lib._pf)
            # The cell list was padded with ITS search radius (normally the same as the search's).
            wlo = whi = None
            if self._window is not None:
                wlo = (C.c_int64 * 3)(*[int(v) for v in self._window[0]] + [1] * (3 - self._ndims))
                whi = (C.c_int64 * 3)(*[int(v) for v in self._window[1]] + [1] * (3 - self._ndims))
            check(_lib.lib().pnb_grid_create_window_f32(
                self._ndims, np.float32(self.search_radius),
                cl._user_min.ctypes.data_as(_lib._pf), cl._user_max.ctypes.data_as(_lib._pf),
                bmn, bmx, wlo, whi, C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _lib.lib().pnb_grid_destroy(h)
            except Exception:
                pass
            self._handle = None

    # -- inspection (tests, exports) ------------------------------------------------------------
    def total_cells(self) -> int:
        return int(_lib.lib().pnb_grid_total_cells(self._grid()))

    def layout(self) -> str:
        """Layout the last initialize!/update! wrote: "csr" (two-pass counting sort) or "buckets"
        (one-pass update!); inspection only, every consumer handles both."""
        return {1: "buckets", 0: "csr"}.get(int(_lib.lib().pnb_grid_layout(self._grid())), "unbuilt")

    def export_csr(self):
        """(cell_start[C+1], cell_points[n]) int32 CUDA tensors, 0-based ids."""
        torch = _torch()
        g = self._grid()
        Cn = self.total_cells()
        n = int(_lib.lib().pnb_grid_n_points(g))
        cs = torch.empty(Cn + 1, dtype=torch.int32, device="cuda")
        cp = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
        check(_lib.lib().pnb_grid_export_csr(g, cs.data_ptr(), cp.data_ptr(), 0, _stream()))
        return cs, cp[:n]

    def export_dvov(self, max_points_per_cell=None, index_base=1):
        """The reference's cell storage: backend (C, max_points_per_cell) == Julia's
        max_points_per_cell x C column-major matrix, lengths (C,)."""
        torch = _torch()
        m = int(max_points_per_cell or self.cell_list.max_points_per_cell)
        Cn = self.total_cells()
        backend = torch.zeros((Cn, m), dtype=torch.int32, device="cuda")
        lengths = torch.zeros(Cn, dtype=torch.int32, device="cuda")
        check(_lib.lib().pnb_grid_export_dvov(self._grid(), backend.data_ptr(), lengths.data_ptr(),
                                              m, index_base, _stream()))
        return backend, lengths

    def export_hash_table(self):
        """SpatialHashingCellList only: (coords, collisions) = cell_list.coords as (list_size, 3)
        int32 cell coordinates (the three low words of the reference's UInt128; zeros = unused
        entry) and cell_list.collisions as bool (src/cell_lists/spatial_hashing.jl:24-29)."""
        torch = _torch()
        L = self.total_cells()
        words = torch.zeros((L, 4), dtype=torch.int32, device="cuda")
        coll = torch.zeros(L, dtype=torch.uint8, device="cuda")
        check(_lib.lib().pnb_grid_export_hash_table(self._grid(), words.data_ptr(),
                                                    coll.data_ptr(), _stream()))
        return words[:, :3].contiguous(), coll.bool()

    def points_in_cell(self, cell):
        """cell_list[cell] (spatial_hashing.jl:141-143 / full_grid.jl:163-169): 0-based ids."""
        cs, cp = self.export_csr()
        if isinstance(self.cell_list, SpatialHashingCellList):
            k = spatial_hash(cell, self.cell_list.list_size) - 1
        else:
            gs = self.cell_list.n_cells_per_dimension
            k, stride = 0, 1
            for d in range(self._ndims):
                k += (int(cell[d]) - 1) * stride
                stride *= gs[d]
        cs = cs.cpu().numpy()
        return cp.cpu().numpy()[cs[k]:cs[k + 1]].tolist()

    def point_cells(self, x):
        """0-based linear cell index of every point of x, -1 outside (cell_coords + cell_index)."""
        torch = _torch()
        x = _coords(x, self._ndims, self.eltype)
        out = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
        fn = _lib.lib().pnb_point_cells_f64 if self.eltype == np.float64 \
            else _lib.lib().pnb_point_cells_f32
        check(fn(self._grid(), x.data_ptr(), x.shape[0], out.data_ptr(), _stream()))
        return out


def _stream():
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _coords(x, nd, eltype=np.dtype(np.float32)):
    torch = _torch()
    if not isinstance(x, torch.Tensor):
        raise TypeError("coordinates must be a torch CUDA tensor of shape (N, NDIMS)")
    if not x.is_cuda:
        raise TypeError("coordinates must live on the GPU: pnb200 has no CPU path "
                        "(use adapt(...) / tensor.cuda())")
    want = torch.float64 if np.dtype(eltype) == np.float64 else torch.float32
    if x.dtype != want:
        raise TypeError(f"coordinates must be {str(want).split('.')[-1]} for this search "
                        "(the element type follows the search radius and the cell list)")
    if x.ndim != 2 or x.shape[1] != nd:
        raise ArgumentError(f"coordinates must have shape (N, {nd}) "
                            "(the memory of Julia's NDIMS x N matrix)")
    if not x.is_contiguous():
        raise TypeError("coordinates must be contiguous")
    return x


def _index_tensor(idx, n, what):
    """points / eachindex_y: None, range, sequence or int tensor -> (int32 CUDA tensor | None)."""
    torch = _torch()
    if idx is None:
        return None
    if isinstance(idx, range):
        if idx.step == 1 and idx.start == 0 and idx.stop == n:
            return None
        idx = list(idx)
    if isinstance(idx, torch.Tensor):
        t = idx.to(device="cuda", dtype=torch.int32).contiguous()
    else:
        t = torch.as_tensor(np.asarray(idx, dtype=np.int32), device="cuda")
    if t.numel() > 0:
        lo, hi = int(t.min()), int(t.max())
        if lo < 0 or hi >= n:   # @boundscheck checkbounds (neighborhood_search.jl:192, nhs_grid.jl:269)
            raise BoundsError(f"attempt to access {n} points at {what} index [{lo}, {hi}]")
    return t


def ndims(nhs) -> int:
    return nhs._ndims


def search_radius(nhs):
    return nhs.search_radius


def requires_update(nhs):
    if isinstance(nhs, PrecomputedNeighborhoodSearch):
        return (True, True)      # nhs_precomputed.jl:128
    return (False, True)         # nhs_grid.jl:133


# ---------------------------------------------------------------------------------------------
# initialize! / update!
# ---------------------------------------------------------------------------------------------
def initialize_(nhs, x, y, *, parallelization_backend=None, eachindex_y=None):
    """initialize!(nhs, x, y; eachindex_y)  (src/nhs_grid.jl:220-225, nhs_precomputed.jl:130-147)."""
    if isinstance(nhs, PrecomputedNeighborhoodSearch):
        return nhs._initialize(x, y, eachindex_y)
    y = _coords(y, nhs._ndims, nhs.eltype)
    idx = _index_tensor(eachindex_y, y.shape[0], "eachindex_y")
    if nhs.eltype == np.float64:
        check(_lib.lib().pnb_grid_build_f64(
            nhs._grid(), y.data_ptr(), y.shape[0], None if idx is None else idx.data_ptr(),
            0 if idx is None else idx.numel(), 0, _stream()))
        nhs._y_ref = y
        return nhs
    check(_lib.lib().pnb_grid_build_f32(
        nhs._grid(), y.data_ptr(), y.shape[0], None if idx is None else idx.data_ptr(),
        0 if idx is None else idx.numel(), 0, _stream()))
    nhs._y_ref = y   # keep the coordinates alive: the fast path is keyed on their address
    return nhs


def update_(nhs, x, y, *, points_moving=(True, True), parallelization_backend=None,
            eachindex_y=None, blocking=True):
    """update!(nhs, x, y; points_moving, eachindex_y)  (src/nhs_grid.jl:283-292).

    blocking=False (no reference counterpart; the reference synchronises after every launch,
    src/util.jl:166-170): the rebuild is only enqueued on the current stream.  A domain error of
    that build is raised by the next blocking call on the search (the sweep that follows, or
    `check_(nhs)`); Float32 FullGridCellList searches only, everything else stays blocking."""
    if isinstance(nhs, PrecomputedNeighborhoodSearch):
        return nhs._update(x, y, points_moving, eachindex_y)
    # "Only update when the second set is moving." (nhs_grid.jl:289)
    if not points_moving[1]:
        return nhs
    if not blocking and eachindex_y is None and nhs.eltype != np.float64:
        y = _coords(y, nhs._ndims, nhs.eltype)
        check(_lib.lib().pnb_grid_build_async_f32(nhs._grid(), y.data_ptr(), y.shape[0], _stream()))
        nhs._y_ref = y
        return nhs
    return initialize_(nhs, x, y, eachindex_y=eachindex_y)


class HostStepper:
    """The WCSPH step (update! + interact!) from HOST buffers, pipelined over three streams with
    double-buffered device arrays (pnb_hoststep_*, include/pnb200.h): what a host-side caller of
    the reference does per step -- copy coordinates and state to the device, update!, interact!,
    copy dv back -- with the copies of neighbouring steps overlapping the kernels.

        stepper = HostStepper(nhs, n_points)
        stepper.submit(y_host, v_host, pressure_host, dv_host, params, mass_host=mass_host)
        ...                       # pressure_host=None after set_state_equation(...): computed on the device
        stepper.wait()            # every dv_host passed so far is complete

    Host arrays: contiguous float32 numpy arrays or CPU torch tensors (pin them for overlap)."""

    def __init__(self, nhs, n_points: int):
        if isinstance(nhs, PrecomputedNeighborhoodSearch) or nhs.eltype == np.float64:
            raise ArgumentError("HostStepper needs a Float32 GridNeighborhoodSearch")
        self.nhs = nhs
        self.n = int(n_points)
        h = C.c_void_p()
        check(_lib.lib().pnb_hoststep_create(nhs._grid(), self.n, C.byref(h)))
        self._h = h
        self._keep = []

    @staticmethod
    def _hptr(a, n_elems, what):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            if a.is_cuda or a.dtype != torch.float32 or not a.is_contiguous() or a.numel() != n_elems:
                raise TypeError(f"{what}: contiguous float32 CPU tensor with {n_elems} elements expected")
            return a.data_ptr()
        a = np.asarray(a)
        if a.dtype != np.float32 or not a.flags.c_contiguous or a.size != n_elems:
            raise TypeError(f"{what}: contiguous float32 array with {n_elems} elements expected")
        return a.ctypes.data

    def set_state_equation(self, *, sound_speed, reference_density, exponent=1.0, background_pressure=0.0):
        """StateEquationCole of the system: `submit(..., pressure_host=None, ...)` then computes the
        pressure on the device (compute_pressure!) instead of copying it from the host."""
        check(_lib.lib().pnb_hoststep_set_state_equation(
            self._h, np.float32(sound_speed), np.float32(reference_density), np.float32(exponent),
            np.float32(background_pressure)))
        return self

    def submit(self, y_host, v_host, pressure_host, dv_host, params, mass_host=None):
        nd = self.nhs._ndims
        n = self.n
        prm = params.params if isinstance(params, WCSPHInteract) else params
        self._keep = (self._keep + [(y_host, v_host, pressure_host, dv_host, mass_host)])[-3:]
        check(_lib.lib().pnb_hoststep_wcsph_submit(
            self._h, self._hptr(y_host, n * nd, "y_host"), self._hptr(v_host, n * (nd + 1), "v_host"),
            None if mass_host is None else self._hptr(mass_host, n, "mass_host"),
            None if pressure_host is None else self._hptr(pressure_host, n, "pressure_host"), C.byref(prm),
            self._hptr(dv_host, n * (nd + 1), "dv_host")))

    def wait(self):
        check(_lib.lib().pnb_hoststep_wait(self._h))

    def host_times(self, since=None):
        """Host milliseconds spent inside the submits {enqueueing the H2D copies, waiting for the
        previous step, enqueueing update!, enqueueing interact!, enqueueing the D2H copy} and the number of submits,
        since the creation -- or, with since = an earlier result, per submit since then."""
        out, steps = (C.c_double * 5)(), C.c_int64()
        check(_lib.lib().pnb_hoststep_host_times(self._h, out, C.byref(steps)))
        names = ("h2d_enqueue_ms", "settle_wait_ms", "update_enqueue_ms", "interact_enqueue_ms",
                 "d2h_enqueue_ms")
        tot = dict(zip(names, (1e3 * v for v in out)))
        tot["submits"] = int(steps.value)
        if since is None:
            return tot
        k = max(tot["submits"] - since["submits"], 1)
        return {nm: (tot[nm] - since[nm]) / k for nm in names}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None:
            try:
                _lib.lib().pnb_hoststep_destroy(h)
            except Exception:
                pass
            self._h = None


def check_(nhs, settled=False):
    """Synchronise the current stream and raise what a blocking update! would have raised
    (settles `update_(..., blocking=False)`).  settled=True: the caller has already waited for an
    event recorded behind the update!; nothing is synchronised (later kernels keep running)."""
    if isinstance(nhs, PrecomputedNeighborhoodSearch) or nhs.eltype == np.float64:
        return nhs
    if settled:
        check(_lib.lib().pnb_grid_check_settled(nhs._grid()))
    else:
        check(_lib.lib().pnb_grid_check(nhs._grid(), _stream()))
    return nhs


initialize = initialize_
update = update_


# ---------------------------------------------------------------------------------------------
# closures that have a fused kernel
# ---------------------------------------------------------------------------------------------
class CountNeighbors:
    """`n_neighbors[i] += 1` (benchmarks/count_neighbors.jl:24-27); n_neighbors: int64 CUDA tensor.
    Like the benchmark, the array is zeroed first (`n_neighbors .= 0`, :22)."""

    def __init__(self, n_neighbors):
        self.n_neighbors = n_neighbors


class NBodyGravity:
    """benchmarks/n_body.jl:38-48: dv zeroed, then dv[:, i] += -G * mass[j] * pos_diff / distance^3."""

    def __init__(self, dv, mass, G):
        self.dv, self.mass, self.G = dv, mass, G


def wendland_c2_norm(ndims_: int, h) -> np.float32:
    """sigma_d / h^d of the Wendland C2 kernel (TrixiParticles kernel_normalization), in Float32."""
    h = np.float32(h)
    if ndims_ == 3:
        sigma = np.float32(21.0 / (16.0 * np.pi))
        return np.float32(sigma / (h * h * h))
    if ndims_ == 2:
        sigma = np.float32(7.0 / (4.0 * np.pi))
        return np.float32(sigma / (h * h))
    raise ArgumentError("WendlandC2Kernel is defined for 2 and 3 dimensions")


class WCSPHInteract:
    """TrixiParticles.interact!(dv, v, u, v, u, system, system, semi) as configured by
    benchmarks/smoothed_particle_hydrodynamics.jl:45-102.  v, dv: (N, NDIMS+1) = velocity, density."""

    def __init__(self, dv, v_x, v_y, mass_x, mass_y, pressure_x, pressure_y, *, smoothing_length,
                 sound_speed, alpha=0.02, beta=0.0, epsilon=0.01, delta=0.1, kernel_norm=None,
                 ndims_=3):
        self.dv = dv
        self.v_x, self.v_y = v_x, v_y
        self.mass_x, self.mass_y = mass_x, mass_y
        self.pressure_x, self.pressure_y = pressure_x, pressure_y
        if kernel_norm is None:
            kernel_norm = wendland_c2_norm(ndims_, smoothing_length)
        self.params = WcsphParams(np.float32(smoothing_length), np.float32(sound_speed),
                                  np.float32(alpha), np.float32(beta), np.float32(epsilon),
                                  np.float32(delta), np.float32(kernel_norm))
        # the same in Float64 (Float64 searches, pnb_wcsph_interact_f64)
        self.params64 = (float(smoothing_length), float(sound_speed), float(alpha), float(beta),
                         float(epsilon), float(delta), float(kernel_norm))

    def params_array(self):
        p = self.params
        return np.array([p.smoothing_length, p.sound_speed, p.alpha, p.beta, p.epsilon, p.delta,
                         p.kernel_norm], dtype=np.float32)


class TLSPHDeformationGradient:
    """TrixiParticles.calc_deformation_grad! over a PrecomputedNeighborhoodSearch
    (benchmarks/smoothed_particle_hydrodynamics.jl:136-189)."""

    def __init__(self, F, current_coordinates, mass, material_density, correction_matrix, *,
                 smoothing_length, kernel_norm=None, ndims_=3):
        self.F = F
        self.current_coordinates = current_coordinates
        self.mass = mass
        self.material_density = material_density
        self.correction_matrix = correction_matrix
        self.smoothing_length = np.float32(smoothing_length)
        self.kernel_norm = np.float32(kernel_norm if kernel_norm is not None
                                      else wendland_c2_norm(ndims_, smoothing_length))


class TLSPHInteract:
    """TrixiParticles.interact_structure_structure!(dv, v, system, semi) over a
    PrecomputedNeighborhoodSearch (benchmarks/smoothed_particle_hydrodynamics.jl:121, set-up
    :136-189): PK1 stress forces + PenaltyForceGanzenmueller.  dv (N, NDIMS) is overwritten;
    pk1_corrected and deformation_grad are (N, NDIMS*NDIMS), column-major per point."""

    def __init__(self, dv, current_coordinates, mass, material_density, pk1_corrected,
                 deformation_grad, *, smoothing_length, young_modulus, penalty_alpha=0.1,
                 kernel_norm=None, ndims_=3):
        self.dv = dv
        self.current_coordinates = current_coordinates
        self.mass = mass
        self.material_density = material_density
        self.pk1_corrected = pk1_corrected
        self.deformation_grad = deformation_grad
        if kernel_norm is None:
            kernel_norm = wendland_c2_norm(ndims_, smoothing_length)
        self.params = TlsphParams(np.float32(smoothing_length), np.float32(kernel_norm),
                                  np.float32(young_modulus), np.float32(penalty_alpha))


def compute_pk1_corrected_(pk1_corrected, deformation_grad, correction_matrix, *, young_modulus,
                           poisson_ratio):
    """TrixiParticles.compute_pk1_corrected!(system, semi)
    (benchmarks/smoothed_particle_hydrodynamics.jl:186), pointwise on the device:
    (N, NDIMS*NDIMS) float32 CUDA tensors, column-major per point."""
    n, nn = deformation_grad.shape
    nd = int(round(nn ** 0.5))
    check(_lib.lib().pnb_tlsph_pk1_corrected_f32(
        nd, n, deformation_grad.data_ptr(), correction_matrix.data_ptr(),
        np.float32(young_modulus), np.float32(poisson_ratio), pk1_corrected.data_ptr(), _stream()))
    return pk1_corrected


def compute_pressure_(pressure, v, *, sound_speed, reference_density, exponent=1.0,
                      background_pressure=0.0):
    """TrixiParticles.compute_pressure!(system, v, semi) for ContinuityDensity +
    StateEquationCole (benchmarks/smoothed_particle_hydrodynamics.jl:64-69, 99): v is the
    (N, NDIMS+1) state whose last column is the density."""
    n, ns = v.shape
    check(_lib.lib().pnb_wcsph_compute_pressure_f32(
        ns - 1, n, v.data_ptr(), np.float32(sound_speed), np.float32(reference_density),
        np.float32(exponent), np.float32(background_pressure), pressure.data_ptr(), _stream()))
    return pressure


def _ptr(t):
    return None if t is None else t.data_ptr()


def set_exact_arithmetic(on: bool) -> None:
    """Per-pair terms of the fused n-body / WCSPH closures: False (default) = MUFU + FMA fast
    arithmetic (within the 1e-5 bar), True = the reference's IEEE operation sequence (sums
    bit-identical to the CPU oracle).  The neighbour test itself is always exact."""
    _lib.lib().pnb_set_exact_arithmetic(int(bool(on)))


def foreach_point_neighbor(f, system_coords, neighbor_coords, neighborhood_search, *,
                           parallelization_backend=None, points=None, blocking=True):
    """foreach_point_neighbor(f, x, y, nhs; points)  (src/neighborhood_search.jl:183-201).

    f is one of the fused closures (CountNeighbors, NBodyGravity, WCSPHInteract,
    TLSPHDeformationGradient) or any Python callable f(i, j, pos_diff, distance); a callable is
    served from a device-built neighbour list (the pairs are computed on the GPU, the callable is
    host code and runs on the host).  Returns None like the reference.

    blocking=False (WCSPHInteract over all points with x === y only; no reference counterpart):
    gather + sweep are only enqueued on the current stream; `check_(nhs)` settles the chain."""
    nhs = neighborhood_search
    nd = nhs._ndims
    x = _coords(system_coords, nd, nhs.eltype)
    y = _coords(neighbor_coords, nd, nhs.eltype)
    if isinstance(nhs, PrecomputedNeighborhoodSearch):
        return nhs._foreach(f, x, y, points)
    L = _lib.lib()
    if nhs.eltype == np.float64:
        pts = _index_tensor(points, x.shape[0], "points")
        if isinstance(f, CountNeighbors):
            check(L.pnb_count_neighbors_f64(nhs._grid(), x.data_ptr(), x.shape[0], y.data_ptr(),
                                            y.shape[0], _ptr(pts), 0 if pts is None else pts.numel(),
                                            0, f.n_neighbors.data_ptr(), _stream()))
        elif isinstance(f, (NBodyGravity, WCSPHInteract)) and getattr(nhs.cell_list, "mixed", False):
            # mixed precision: Float64 coordinates, Float32 pos_diff / distance, Float32 state
            import torch
            arrs = [f.dv, f.mass] if isinstance(f, NBodyGravity) else \
                [f.dv, f.v_x, f.v_y, f.mass_x, f.mass_y, f.pressure_x, f.pressure_y]
            if any(a.dtype != torch.float32 for a in arrs):
                raise TypeError("a mixed-precision search hands Float32 pos_diff / distance to the closure: "
                                "its state arrays and dv must be float32")
            n_pts = 0 if pts is None else pts.numel()
            if isinstance(f, NBodyGravity):
                check(L.pnb_nbody_mixed(nhs._grid(), x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0],
                                        _ptr(pts), n_pts, 0, f.mass.data_ptr(), np.float32(f.G),
                                        f.dv.data_ptr(), _stream()))
            else:
                check(L.pnb_wcsph_interact_mixed(nhs._grid(), x.data_ptr(), x.shape[0], y.data_ptr(),
                                                 y.shape[0], _ptr(pts), n_pts, 0, f.v_x.data_ptr(),
                                                 f.v_y.data_ptr(), f.mass_x.data_ptr(), f.mass_y.data_ptr(),
                                                 f.pressure_x.data_ptr(), f.pressure_y.data_ptr(),
                                                 C.byref(f.params), f.dv.data_ptr(), _stream()))
        elif isinstance(f, NBodyGravity) and not getattr(nhs.cell_list, "mixed", False):
            check(L.pnb_nbody_f64(nhs._grid(), x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0],
                                  _ptr(pts), 0 if pts is None else pts.numel(), 0, f.mass.data_ptr(),
                                  float(f.G), f.dv.data_ptr(), _stream()))
        elif isinstance(f, WCSPHInteract) and not getattr(nhs.cell_list, "mixed", False):
            p64 = _lib.WcsphParams64(*f.params64)
            check(L.pnb_wcsph_interact_f64(nhs._grid(), x.data_ptr(), x.shape[0], y.data_ptr(),
                                           y.shape[0], _ptr(pts), 0 if pts is None else pts.numel(), 0,
                                           f.v_x.data_ptr(), f.v_y.data_ptr(), f.mass_x.data_ptr(),
                                           f.mass_y.data_ptr(), f.pressure_x.data_ptr(),
                                           f.pressure_y.data_ptr(), C.byref(p64), f.dv.data_ptr(),
                                           _stream()))
        elif callable(f) and not isinstance(f, (NBodyGravity, WCSPHInteract,
                                                 TLSPHDeformationGradient, TLSPHInteract)):
            lists = _NeighborLists.build(nhs, x, y, sort=False)
            lists.call_host(f, x, y, nhs, points, radius_test=True)
        else:
            raise TypeError("Float64 searches have the fused n-body / WCSPH closures (all arrays float64; "
                            "float32 state on a mixed-precision search); the TLSPH closures use Float32 "
                            "searches / neighbour lists")
        return None
    pts = _index_tensor(points, x.shape[0], "points")
    npts = 0 if pts is None else pts.numel()
    g = nhs._grid()
    if isinstance(f, CountNeighbors):
        check(L.pnb_count_neighbors_f32(g, x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0],
                                        _ptr(pts), npts, 0, f.n_neighbors.data_ptr(), _stream()))
    elif isinstance(f, NBodyGravity):
        check(L.pnb_nbody_f32(g, x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0], _ptr(pts),
                              npts, 0, f.mass.data_ptr(), np.float32(f.G), f.dv.data_ptr(),
                              _stream()))
    elif isinstance(f, WCSPHInteract) and not blocking and pts is None and \
            x.data_ptr() == y.data_ptr() and x.shape[0] == y.shape[0]:
        check(L.pnb_wcsph_interact_async_f32(g, y.data_ptr(), y.shape[0], f.v_y.data_ptr(),
                                             f.mass_y.data_ptr(), f.pressure_y.data_ptr(),
                                             C.byref(f.params), f.dv.data_ptr(), _stream()))
    elif isinstance(f, WCSPHInteract):
        check(L.pnb_wcsph_interact_f32(g, x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0],
                                       _ptr(pts), npts, 0, f.v_x.data_ptr(), f.v_y.data_ptr(),
                                       f.mass_x.data_ptr(), f.mass_y.data_ptr(),
                                       f.pressure_x.data_ptr(), f.pressure_y.data_ptr(),
                                       C.byref(f.params), f.dv.data_ptr(), _stream()))
    elif callable(f):
        lists = _NeighborLists.build(nhs, x, y, sort=False)
        lists.call_host(f, x, y, nhs, points, radius_test=True)
    else:
        raise TypeError("f must be a fused closure object or a callable f(i, j, pos_diff, d)")
    return None


def _check_point(point, n):
    i = int(point)
    if not 0 <= i < n:
        # the safe variants bounds-check `point` against system_coords (neighborhood_search.jl:250-262)
        raise BoundsError(f"BoundsError: attempt to access {n}-column coordinates at index [{point}]")
    return i


def foreach_neighbor(f, system_coords, neighbor_coords, neighborhood_search, point, *,
                     search_radius=None):
    """foreach_neighbor(f, x, y, nhs, point)  (src/neighborhood_search.jl:236-276): f(i, j,
    pos_diff, d) for every neighbour of ONE point, in the reference's visiting order.  The pairs
    come from the device (a one-point neighbour list); f is host code.  0-based `point`."""
    if search_radius is not None and np.float32(search_radius) != np.float32(
            neighborhood_search.search_radius):
        raise ArgumentError("a `search_radius` other than the one of the neighborhood search "
                            "is not supported")
    nhs = neighborhood_search
    x = _coords(system_coords, nhs._ndims, nhs.eltype)
    y = _coords(neighbor_coords, nhs._ndims, nhs.eltype)
    i = _check_point(point, x.shape[0])
    if isinstance(nhs, PrecomputedNeighborhoodSearch):
        return nhs._foreach(f, x, y, [i])
    # a one-point query set: list 0 of the device-built lists belongs to `point`
    x_one = x[i:i + 1].contiguous()
    lists = _NeighborLists.build(nhs, x_one, y, sort=False)
    lists.call_host(lambda _i, j, pd, d: f(i, j, pd, d), x_one, y, nhs, None, radius_test=True)
    return None


foreach_neighbor_unsafe = foreach_neighbor   # the bounds checks are host-side and cost nothing here


def foreach_point_neighbor_unsafe(f, system_coords, neighbor_coords, neighborhood_search, **kw):
    """foreach_point_neighbor_unsafe (src/neighborhood_search.jl:204-234): the reference skips the
    bounds checks INSIDE its kernel; the explicit check of `points` before the loop stays (:226).
    The device kernels here never index out of bounds on a search initialized with these arrays
    (stencil cells outside the grid are reported, not read), so both names run the same code."""
    return foreach_point_neighbor(f, system_coords, neighbor_coords, neighborhood_search, **kw)


def initialize_grid_(neighborhood_search, y, *, parallelization_backend=None, eachindex_y=None):
    """initialize_grid!(nhs, y; eachindex_y)  (src/nhs_grid.jl:227-281): the cell-list part of
    initialize! (which only forwards to it, :220-225)."""
    return initialize_(neighborhood_search, y, y, eachindex_y=eachindex_y)


def update_grid_(neighborhood_search, y, *, parallelization_backend=None, eachindex_y=None):
    """update_grid!(nhs, y; eachindex_y) of ParallelUpdate / SerialUpdate = a full rebuild
    (src/nhs_grid.jl:470-477)."""
    return update_(neighborhood_search, y, y, points_moving=(True, True), eachindex_y=eachindex_y)


class B200Backend:
    """What `default_backend(x)` returns for a CUDA tensor (src/util.jl:81-83 returns the
    KernelAbstractions backend of a GPU array): a marker, there is one device code path."""

    def __repr__(self):
        return "B200Backend()"


def default_backend(x):
    """default_backend(x) (src/util.jl:81-83).  Host arrays have no backend here: the CPU thread
    backends of the reference are out of scope and the library has no CPU path."""
    torch = _torch()
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return B200Backend()
    raise ArgumentError("pnb200 has a device code path only: pass CUDA tensors (the reference's CPU "
                        "thread backends are out of scope)")


def mapreduce_neighbor(f, op, system_coords, neighbor_coords, neighborhood_search, point, *, init,
                       search_radius=None):
    """mapreduce_neighbor(f, op, x, y, nhs, point; init)  (src/neighborhood_search.jl:325-357):
    op-reduction of f(i, j, pos_diff, d) over the neighbours of `point`, starting from `init`."""
    acc = [init]

    def g(i, j, pos_diff, d):
        acc[0] = op(acc[0], f(i, j, pos_diff, d))

    foreach_neighbor(g, system_coords, neighbor_coords, neighborhood_search, point,
                     search_radius=search_radius)
    return acc[0]


mapreduce_neighbor_unsafe = mapreduce_neighbor


# ---------------------------------------------------------------------------------------------
# neighbour lists
# ---------------------------------------------------------------------------------------------
class _NeighborLists:
    """Owner of a pnb_nlist handle (device CSR)."""

    def __init__(self, handle, nd):
        self._handle = handle
        self._nd = nd

    @classmethod
    def build(cls, grid_nhs, x, y, sort=True):
        h = C.c_void_p()
        fn = _lib.lib().pnb_nlist_build_f64 if grid_nhs.eltype == np.float64 \
            else _lib.lib().pnb_nlist_build_f32
        check(fn(grid_nhs._grid(), x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0],
                 int(bool(sort)), C.byref(h), _stream()))
        return cls(h, grid_nhs._ndims)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _lib.lib().pnb_nlist_destroy(h)
            except Exception:
                pass
            self._handle = None

    @property
    def n_points(self):
        return int(_lib.lib().pnb_nlist_n_points(self._handle))

    @property
    def n_pairs(self):
        return int(_lib.lib().pnb_nlist_n_pairs(self._handle))

    def export_csr(self, index_base=0):
        torch = _torch()
        off = torch.empty(self.n_points + 1, dtype=torch.int64, device="cuda")
        ids = torch.empty(max(self.n_pairs, 1), dtype=torch.int32, device="cuda")
        check(_lib.lib().pnb_nlist_export_csr(self._handle, off.data_ptr(), ids.data_ptr(),
                                              index_base, _stream()))
        return off, ids[:self.n_pairs]

    def export_dvov(self, max_neighbors, transposed=False, index_base=1):
        """(backend, lengths).  backend has the PARENT memory layout of the reference:
        regular: (n_points, max_neighbors) row-major == Julia max_neighbors x n_points;
        transposed: (max_neighbors, n_points) row-major == Julia parent n_points x max_neighbors
        (src/vector_of_vectors.jl:18-26)."""
        torch = _torch()
        n = self.n_points
        shape = (max_neighbors, n) if transposed else (n, max_neighbors)
        backend = torch.empty(shape, dtype=torch.int32, device="cuda")
        lengths = torch.empty(n, dtype=torch.int32, device="cuda")
        check(_lib.lib().pnb_nlist_export_dvov(self._handle, backend.data_ptr(), lengths.data_ptr(),
                                               int(max_neighbors), int(bool(transposed)),
                                               index_base, _stream()))
        return backend, lengths

    def pairs(self, grid_nhs, x, y):
        """pos_diff (P, nd) and distance (P,) of every listed pair, in list order."""
        torch = _torch()
        P = self.n_pairs
        f64 = grid_nhs.eltype == np.float64
        dt = torch.float64 if f64 else torch.float32
        pd = torch.empty((max(P, 1), self._nd), dtype=dt, device="cuda")
        dist = torch.empty(max(P, 1), dtype=dt, device="cuda")
        fn = _lib.lib().pnb_nlist_pairs_f64 if f64 else _lib.lib().pnb_nlist_pairs_f32
        check(fn(self._handle, grid_nhs._grid(), x.data_ptr(), y.data_ptr(), pd.data_ptr(),
                 dist.data_ptr(), _stream()))
        return pd[:P], dist[:P]

    def call_host(self, f: Callable, x, y, grid_nhs, points, radius_test: bool):
        off, ids = self.export_csr(0)
        pd, dist = self.pairs(grid_nhs, x, y)
        off = off.cpu().numpy()
        ids = ids.cpu().numpy()
        pd = pd.cpu().numpy()
        dist = dist.cpu().numpy()
        loop = range(x.shape[0]) if points is None else [int(p) for p in points]
        for i in loop:
            for k in range(off[i], off[i + 1]):
                f(i, int(ids[k]), pd[k], dist[k])


class PrecomputedNeighborhoodSearch(metaclass=_Parametric):
    """PrecomputedNeighborhoodSearch{NDIMS}(; search_radius, n_points, periodic_box,
    update_strategy, update_neighborhood_search, backend, transpose_backend, max_neighbors,
    sort_neighbor_lists)  (src/nhs_precomputed.jl:67-124).

    The lists live on the device as CSR; `neighbor_lists()` exports them in the reference's
    DynamicVectorOfVectors layout (regular or transposed) with `max_neighbors` rows."""

    @staticmethod
    def default_max_neighbors(nd):
        # nhs_precomputed.jl:114-124
        if nd == 1:
            return 32
        if nd == 2:
            return 64
        if nd == 3:
            return 320
        raise ArgumentError("`NDIMS` must be 1, 2, or 3")

    def __init__(self, ndims_: int, *, search_radius=0.0, n_points: int = 0, periodic_box=None,
                 update_strategy=None, update_neighborhood_search=None,
                 backend=DynamicVectorOfVectors[np.int32], transpose_backend: bool = False,
                 max_neighbors: Optional[int] = None, sort_neighbor_lists: bool = True):
        self._ndims = int(ndims_)
        if _is_integer_scalar(search_radius):
            raise ArgumentError("`search_radius` cannot be an integer type, since computed "
                                "distances will be converted to this type")
        if update_neighborhood_search is None:
            raise ArgumentError("the default update_neighborhood_search uses a DictionaryCellList, "
                                "which is not GPU-compatible; pass a GridNeighborhoodSearch with a "
                                "FullGridCellList (src/nhs_precomputed.jl:37-42)")
        self.search_radius = search_radius
        self.periodic_box = periodic_box
        self.neighborhood_search = update_neighborhood_search
        self.backend = backend
        self.transpose_backend = bool(transpose_backend)
        self.max_neighbors = int(max_neighbors if max_neighbors is not None
                                 else self.default_max_neighbors(self._ndims))
        self.sort_neighbor_lists = bool(sort_neighbor_lists)
        self.n_points = int(n_points)
        self._lists: Optional[_NeighborLists] = None
        self._grid_for_pairs = update_neighborhood_search
        self.eltype = update_neighborhood_search.eltype

    # initialize! (nhs_precomputed.jl:130-147)
    def _initialize(self, x, y, eachindex_y):
        nd = self._ndims
        x = _coords(x, nd, self.eltype)
        y = _coords(y, nd, self.eltype)
        if _index_tensor(eachindex_y, y.shape[0], "eachindex_y") is not None:
            raise PointNeighborsError("this neighborhood search does not support inactive points")
        if self.neighborhood_search is None:
            raise PointNeighborsError("this neighborhood search has been frozen and cannot be "
                                      "initialized or updated")
        initialize_(self.neighborhood_search, x, y)
        self._build_lists(x, y)
        return self

    # update! (nhs_precomputed.jl:149-169)
    def _update(self, x, y, points_moving, eachindex_y):
        nd = self._ndims
        x = _coords(x, nd, self.eltype)
        y = _coords(y, nd, self.eltype)
        if _index_tensor(eachindex_y, y.shape[0], "eachindex_y") is not None:
            raise PointNeighborsError("this neighborhood search does not support inactive points")
        if self.neighborhood_search is None:
            raise PointNeighborsError("this neighborhood search has been frozen and cannot be "
                                      "initialized or updated")
        update_(self.neighborhood_search, x, y, points_moving=points_moving)
        if any(points_moving):
            self._build_lists(x, y)
        return self

    def _build_lists(self, x, y):
        self._lists = _NeighborLists.build(self.neighborhood_search, x, y,
                                           sort=self.sort_neighbor_lists)
        # the reference errors when a list overflows `max_neighbors` (vector_of_vectors.jl:114-121)
        if int(_lib.lib().pnb_nlist_max_length(self._lists._handle)) > self.max_neighbors:
            raise PointNeighborsError("cell list is full. Use a larger `max_points_per_cell`.")

    def neighbor_lists(self, index_base=1):
        """(backend, lengths) in the reference layout, see _NeighborLists.export_dvov."""
        return self._lists.export_dvov(self.max_neighbors, self.transpose_backend, index_base)

    def export_csr(self, index_base=0):
        return self._lists.export_csr(index_base)

    def _foreach(self, f, x, y, points):
        if self._lists is None:
            raise PointNeighborsError("the neighborhood search has not been initialized")
        if isinstance(f, TLSPHDeformationGradient):
            if points is not None:
                raise ArgumentError("the fused TLSPH sweep loops over all points")
            if self.eltype == np.float64:
                raise TypeError("the fused TLSPH sweep exists in Float32 only")
            check(_lib.lib().pnb_tlsph_deformation_grad_f32(
                self._lists._handle, self._grid_for_pairs._grid(), x.data_ptr(),
                f.current_coordinates.data_ptr(), f.mass.data_ptr(),
                f.material_density.data_ptr(), f.correction_matrix.data_ptr(),
                f.smoothing_length, f.kernel_norm, f.F.data_ptr(), _stream()))
            return None
        if isinstance(f, TLSPHInteract):
            if points is not None:
                raise ArgumentError("the fused TLSPH sweep loops over all points")
            if self.eltype == np.float64:
                raise TypeError("the fused TLSPH sweep exists in Float32 only")
            check(_lib.lib().pnb_tlsph_interact_f32(
                self._lists._handle, self._grid_for_pairs._grid(), x.data_ptr(),
                f.current_coordinates.data_ptr(), f.mass.data_ptr(),
                f.material_density.data_ptr(), f.pk1_corrected.data_ptr(),
                f.deformation_grad.data_ptr(), C.byref(f.params), f.dv.data_ptr(), _stream()))
            return None
        if callable(f):
            self._lists.call_host(f, x, y, self._grid_for_pairs, points, radius_test=False)
            return None
        raise TypeError("f must be TLSPHDeformationGradient, TLSPHInteract or a callable "
                        "f(i, j, pos_diff, d)")


# ---------------------------------------------------------------------------------------------
# copy / freeze
# ---------------------------------------------------------------------------------------------
def copy_neighborhood_search(nhs, search_radius_, n_points, *, eachpoint=None):
    """copy_neighborhood_search(nhs, search_radius, n_points)  (src/nhs_grid.jl:640-649,
    src/nhs_precomputed.jl:249-266): a new, uninitialized search with the same options."""
    if isinstance(nhs, PrecomputedNeighborhoodSearch):
        inner = copy_neighborhood_search(nhs.neighborhood_search, search_radius_, n_points)
        return PrecomputedNeighborhoodSearch(
            nhs._ndims, search_radius=search_radius_, n_points=n_points,
            periodic_box=nhs.periodic_box, update_neighborhood_search=inner, backend=nhs.backend,
            transpose_backend=nhs.transpose_backend, max_neighbors=nhs.max_neighbors,
            sort_neighbor_lists=nhs.sort_neighbor_lists)
    cell_list = copy_cell_list(nhs.cell_list, search_radius_, nhs.periodic_box)
    return GridNeighborhoodSearch(nhs._ndims, search_radius=search_radius_, n_points=n_points,
                                  periodic_box=nhs.periodic_box, cell_list=cell_list,
                                  update_strategy=nhs.update_strategy)


def freeze_neighborhood_search(nhs):
    """src/neighborhood_search.jl:112-118, src/nhs_precomputed.jl:268-277.  The lists keep working;
    the inner grid search is no longer available for updates.  (The scalars needed by the list
    sweep -- periodic box, radius -- stay with the frozen object.)"""
    if isinstance(nhs, PrecomputedNeighborhoodSearch):
        frozen = PrecomputedNeighborhoodSearch.__new__(PrecomputedNeighborhoodSearch)
        frozen.__dict__.update(nhs.__dict__)
        frozen.neighborhood_search = None
        return frozen
    return nhs
