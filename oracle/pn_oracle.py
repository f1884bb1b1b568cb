"""ctypes front-end of the CPU oracle (oracle/pn_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.

All point ids are 0-based; cell coordinates are 1-based like the reference
(/root/reference/src/cell_lists/full_grid.jl:84-94).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpn_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (make -C oracle)."""
    src_mtime = max(os.path.getmtime(os.path.join(_HERE, f))
                    for f in ("pn_oracle.c", "pn_oracle_impl.h", "pn_oracle_mixed.h", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_mtime:
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.pno_max_threads.restype = C.c_int
        for suf in ("_f32", "_f64"):
            getattr(_lib, "pno_total_cells" + suf).restype = C.c_int64
            getattr(_lib, "pno_candidate_tests" + suf).restype = C.c_int64
            getattr(_lib, "pno_spatial_hash" + suf).restype = C.c_int64
        _lib.pno_total_cells_mix.restype = C.c_int64
    return _lib


def max_threads() -> int:
    return int(lib().pno_max_threads())


def set_threads(n: int) -> None:
    lib().pno_set_threads(int(n))


def _grid_struct(real):
    class Grid(C.Structure):
        _fields_ = [("ndims", C.c_int32), ("periodic", C.c_int32), ("search_radius", real),
                    ("min_corner", real * 3), ("max_corner", real * 3),
                    ("grid_size", C.c_int64 * 3), ("n_cells", C.c_int64 * 3),
                    ("cell_size", real * 3), ("box_min", real * 3), ("box_max", real * 3),
                    ("box_size", real * 3)]
    return Grid


_GridF32 = _grid_struct(C.c_float)
_GridF64 = _grid_struct(C.c_double)

ERR_TEXT = {
    1: "particle coordinates are NaN or outside the domain bounds of the cell list",
    2: "the `GridNeighborhoodSearch` needs at least 3 cells in each dimension when used with "
       "periodicity. Please use no NHS for very small problems.",
    3: "cell list is full. Use a larger `max_points_per_cell`.",
    4: "BoundsError: neighbouring cell outside the cell grid",
    5: "InexactError: a cell coordinate does not fit Int32 (coordinates_flattened)",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(ERR_TEXT.get(code, f"oracle error {code}"))
        self.code = code


def _ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


class Grid:
    """GridNeighborhoodSearch + FullGridCellList (+ PeriodicBox) parameters and cell list."""

    def __init__(self, ndims, search_radius, min_corner, max_corner, periodic_box=None,
                 dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.suf = "_f32" if self.dtype == np.float32 else "_f64"
        self.real = C.c_float if self.dtype == np.float32 else C.c_double
        self.g = (_GridF32 if self.dtype == np.float32 else _GridF64)()
        self.ndims = int(ndims)
        mn = np.ascontiguousarray(min_corner, dtype=self.dtype)
        mx = np.ascontiguousarray(max_corner, dtype=self.dtype)
        assert mn.size == self.ndims and mx.size == self.ndims
        if periodic_box is not None:
            bmn = np.ascontiguousarray(periodic_box[0], dtype=self.dtype)
            bmx = np.ascontiguousarray(periodic_box[1], dtype=self.dtype)
        else:
            bmn = bmx = None
        rc = self._fn("pno_grid_init")(C.byref(self.g), self.ndims, self.real(search_radius),
                                       _ptr(mn, self.real), _ptr(mx, self.real),
                                       int(periodic_box is not None), _ptr(bmn, self.real),
                                       _ptr(bmx, self.real))
        if rc:
            raise OracleError(rc)
        self.cell_start = None
        self.cell_points = None
        self.backend = None
        self.lengths = None
        self.max_inner = 0

    def _fn(self, name):
        return getattr(lib(), name + self.suf)

    # ---- scalars -------------------------------------------------------------------------
    @property
    def search_radius(self):
        return self.dtype.type(self.g.search_radius)

    @property
    def min_corner(self):
        return np.array(self.g.min_corner[:self.ndims], dtype=self.dtype)

    @property
    def max_corner(self):
        return np.array(self.g.max_corner[:self.ndims], dtype=self.dtype)

    @property
    def grid_size(self):
        return tuple(int(v) for v in self.g.grid_size[:self.ndims])

    @property
    def n_cells(self):
        return tuple(int(v) for v in self.g.n_cells[:self.ndims])

    @property
    def cell_size(self):
        return np.array(self.g.cell_size[:self.ndims], dtype=self.dtype)

    @property
    def total_cells(self):
        return int(self._fn("pno_total_cells")(C.byref(self.g)))

    @property
    def box(self):
        if not self.g.periodic:
            return None
        return (np.array(self.g.box_min[:self.ndims], dtype=self.dtype),
                np.array(self.g.box_max[:self.ndims], dtype=self.dtype))

    def _box_ptrs(self):
        b = self.box
        if b is None:
            return 0, None, None, None
        return 1, b, _ptr(b[0], self.real), _ptr(b[1], self.real)

    # ---- cells ---------------------------------------------------------------------------
    def _coords(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        assert x.ndim == 2 and x.shape[1] == self.ndims, "coordinates are (N, ndims) row-major"
        return x

    def cell_coords(self, point):
        p = np.ascontiguousarray(point, dtype=self.dtype)
        out = np.zeros(3, dtype=np.int64)
        self._fn("pno_cell_coords")(C.byref(self.g), _ptr(p, self.real), _ptr(out, C.c_int64))
        return tuple(int(v) for v in out[:self.ndims])

    def point_cells(self, x):
        x = self._coords(x)
        out = np.empty(x.shape[0], dtype=np.int64)
        self._fn("pno_point_cells")(C.byref(self.g), _ptr(x, self.real), C.c_int64(x.shape[0]),
                                    _ptr(out, C.c_int64))
        return out

    def build(self, y, eachindex_y=None):
        """initialize! / update! (full rebuild) -> deterministic CSR."""
        y = self._coords(y)
        idx = None if eachindex_y is None else np.ascontiguousarray(eachindex_y, dtype=np.int64)
        n_idx = y.shape[0] if idx is None else idx.size
        self.cell_start = np.zeros(self.total_cells + 1, dtype=np.int64)
        self.cell_points = np.zeros(max(n_idx, 1), dtype=np.int32)
        rc = self._fn("pno_build_csr")(C.byref(self.g), _ptr(y, self.real), C.c_int64(y.shape[0]),
                                       _ptr(idx, C.c_int64), C.c_int64(n_idx),
                                       _ptr(self.cell_start, C.c_int64),
                                       _ptr(self.cell_points, C.c_int32))
        if rc:
            raise OracleError(rc)
        self.cell_points = self.cell_points[:n_idx]
        self.backend = None
        return self

    def build_dvov(self, y, max_points_per_cell=100):
        """The reference's own layout + atomic-push build (timed CPU baseline)."""
        y = self._coords(y)
        Cn = self.total_cells
        if self.backend is None or self.backend.shape != (Cn, max_points_per_cell):
            self.backend = np.zeros((Cn, max_points_per_cell), dtype=np.int32)
            self.lengths = np.zeros(Cn, dtype=np.int32)
        self.max_inner = int(max_points_per_cell)
        rc = self._fn("pno_build_dvov")(C.byref(self.g), _ptr(y, self.real), C.c_int64(y.shape[0]),
                                        C.c_int32(self.max_inner), _ptr(self.backend, C.c_int32),
                                        _ptr(self.lengths, C.c_int32))
        if rc:
            raise OracleError(rc)
        return self

    def cells_as_lists(self):
        return [self.cell_points[self.cell_start[c]:self.cell_start[c + 1]].tolist()
                for c in range(self.total_cells)]

    def _cells_args(self, use_dvov):
        if use_dvov:
            assert self.backend is not None
            return (None, None, _ptr(self.backend, C.c_int32), _ptr(self.lengths, C.c_int32),
                    C.c_int32(self.max_inner))
        assert self.cell_start is not None, "call build() first"
        return (_ptr(self.cell_start, C.c_int64), _ptr(self.cell_points, C.c_int32), None, None,
                C.c_int32(0))

    # ---- sweeps --------------------------------------------------------------------------
    def _points(self, points):
        if points is None:
            return None, 0
        p = np.ascontiguousarray(points, dtype=np.int64)
        return p, p.size

    def count_neighbors(self, x, y, points=None, use_dvov=False, parallel=True):
        x, y = self._coords(x), self._coords(y)
        p, npnt = self._points(points)
        out = np.zeros(x.shape[0], dtype=np.int64)
        rc = self._fn("pno_count_neighbors")(C.byref(self.g), *self._cells_args(use_dvov),
                                             _ptr(x, self.real), C.c_int64(x.shape[0]),
                                             _ptr(y, self.real), _ptr(p, C.c_int64),
                                             C.c_int64(npnt), _ptr(out, C.c_int64),
                                             C.c_int(int(parallel)))
        if rc:
            raise OracleError(rc)
        return out

    def candidate_tests(self, x):
        x = self._coords(x)
        return int(self._fn("pno_candidate_tests")(C.byref(self.g),
                                                   _ptr(self.cell_start, C.c_int64),
                                                   _ptr(x, self.real), C.c_int64(x.shape[0])))

    def nbody(self, x, y, mass, G, points=None, use_dvov=False, parallel=True, wide=False):
        x, y = self._coords(x), self._coords(y)
        mass = np.ascontiguousarray(mass, dtype=self.dtype)
        p, npnt = self._points(points)
        dv = np.zeros_like(x)
        dv64 = np.zeros(x.shape, dtype=np.float64) if wide else None
        dvabs = np.zeros(x.shape, dtype=np.float64) if wide else None
        rc = self._fn("pno_nbody")(C.byref(self.g), *self._cells_args(use_dvov),
                                   _ptr(x, self.real), C.c_int64(x.shape[0]), _ptr(y, self.real),
                                   _ptr(p, C.c_int64), C.c_int64(npnt), _ptr(mass, self.real),
                                   self.real(G), _ptr(dv, self.real), _ptr(dv64, C.c_double),
                                   _ptr(dvabs, C.c_double), C.c_int(int(parallel)))
        if rc:
            raise OracleError(rc)
        return (dv, dv64, dvabs) if wide else dv

    def wcsph(self, x, y, v_x, v_y, mass_x, mass_y, pressure_x, pressure_y, params, points=None,
              use_dvov=False, parallel=True, wide=False):
        x, y = self._coords(x), self._coords(y)
        arrs = [np.ascontiguousarray(a, dtype=self.dtype)
                for a in (v_x, v_y, mass_x, mass_y, pressure_x, pressure_y)]
        params = np.ascontiguousarray(params, dtype=self.dtype)
        assert params.size == 7
        p, npnt = self._points(points)
        ns = self.ndims + 1
        dv = np.zeros((x.shape[0], ns), dtype=self.dtype)
        dv64 = np.zeros(dv.shape, dtype=np.float64) if wide else None
        dvabs = np.zeros(dv.shape, dtype=np.float64) if wide else None
        rc = self._fn("pno_wcsph")(C.byref(self.g), *self._cells_args(use_dvov),
                                   _ptr(x, self.real), C.c_int64(x.shape[0]), _ptr(y, self.real),
                                   _ptr(p, C.c_int64), C.c_int64(npnt),
                                   *[_ptr(a, self.real) for a in arrs],
                                   _ptr(params, self.real), _ptr(dv, self.real),
                                   _ptr(dv64, C.c_double), _ptr(dvabs, C.c_double),
                                   C.c_int(int(parallel)))
        if rc:
            raise OracleError(rc)
        return (dv, dv64, dvabs) if wide else dv

    def neighbor_lists(self, x, y, sort=True):
        """PrecomputedNeighborhoodSearch lists as CSR (offsets[nx+1], ids[P])."""
        x, y = self._coords(x), self._coords(y)
        offsets = np.zeros(x.shape[0] + 1, dtype=np.int64)
        fn = self._fn("pno_neighbor_lists")
        args = (C.byref(self.g), _ptr(self.cell_start, C.c_int64),
                _ptr(self.cell_points, C.c_int32), _ptr(x, self.real), C.c_int64(x.shape[0]),
                _ptr(y, self.real), _ptr(offsets, C.c_int64))
        rc = fn(*args, None, C.c_int(int(sort)))
        if rc:
            raise OracleError(rc)
        ids = np.zeros(max(int(offsets[-1]), 1), dtype=np.int32)
        rc = fn(*args, _ptr(ids, C.c_int32), C.c_int(int(sort)))
        if rc:
            raise OracleError(rc)
        return offsets, ids[:int(offsets[-1])]


class HashGrid:
    """GridNeighborhoodSearch + SpatialHashingCellList (+ PeriodicBox)
    (/root/reference/src/cell_lists/spatial_hashing.jl, hooks src/nhs_grid.jl:479-513).
    Keys are 0-based (the reference's hash key minus 1)."""

    def __init__(self, ndims, search_radius, list_size, periodic_box=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.suf = "_f32" if self.dtype == np.float32 else "_f64"
        self.real = C.c_float if self.dtype == np.float32 else C.c_double
        self.g = (_GridF32 if self.dtype == np.float32 else _GridF64)()
        self.ndims = int(ndims)
        self.list_size = int(list_size)
        if periodic_box is not None:
            bmn = np.ascontiguousarray(periodic_box[0], dtype=self.dtype)
            bmx = np.ascontiguousarray(periodic_box[1], dtype=self.dtype)
        else:
            bmn = bmx = None
        rc = self._fn("pno_hash_grid_init")(C.byref(self.g), self.ndims, self.real(search_radius),
                                            int(periodic_box is not None), _ptr(bmn, self.real),
                                            _ptr(bmx, self.real))
        if rc:
            raise OracleError(rc)
        self.key_start = self.key_points = self.coords = self.collisions = None

    def _fn(self, name):
        return getattr(lib(), name + self.suf)

    def _coords(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        assert x.ndim == 2 and x.shape[1] == self.ndims
        return x

    @property
    def n_cells(self):
        return tuple(int(v) for v in self.g.n_cells[:self.ndims])

    @property
    def cell_size(self):
        return np.array(self.g.cell_size[:self.ndims], dtype=self.dtype)

    def spatial_hash(self, cell):
        c = np.zeros(3, dtype=np.int64)
        c[:self.ndims] = cell
        return int(self._fn("pno_spatial_hash")(self.ndims, _ptr(c, C.c_int64),
                                                C.c_int64(self.list_size)))

    def cell_coords(self, point):
        p = np.ascontiguousarray(point, dtype=self.dtype)
        out = np.zeros(3, dtype=np.int64)
        self._fn("pno_hash_cell_coords")(C.byref(self.g), _ptr(p, self.real), _ptr(out, C.c_int64))
        return tuple(int(v) for v in out[:self.ndims])

    def build(self, y, eachindex_y=None):
        y = self._coords(y)
        idx = None if eachindex_y is None else np.ascontiguousarray(eachindex_y, dtype=np.int64)
        n_idx = y.shape[0] if idx is None else idx.size
        L = self.list_size
        self.key_start = np.zeros(L + 1, dtype=np.int64)
        self.key_points = np.zeros(max(n_idx, 1), dtype=np.int32)
        self.coords = np.zeros((L, 3), dtype=np.int32)
        self.collisions = np.zeros(L, dtype=np.uint8)
        rc = self._fn("pno_hash_build")(C.byref(self.g), C.c_int64(L), _ptr(y, self.real),
                                        C.c_int64(y.shape[0]), _ptr(idx, C.c_int64),
                                        C.c_int64(n_idx), _ptr(self.key_start, C.c_int64),
                                        _ptr(self.key_points, C.c_int32),
                                        _ptr(self.coords, C.c_int32),
                                        _ptr(self.collisions, C.c_uint8))
        if rc:
            raise OracleError(rc)
        self.key_points = self.key_points[:n_idx]
        return self

    def points_in_cell(self, cell):
        """cell_list[cell] (spatial_hashing.jl:141-143): the list of the cell's hash key."""
        k = self.spatial_hash(cell)
        return self.key_points[self.key_start[k]:self.key_start[k + 1]].tolist()

    def _table(self):
        assert self.key_start is not None, "call build() first"
        return (C.byref(self.g), C.c_int64(self.list_size), _ptr(self.key_start, C.c_int64),
                _ptr(self.key_points, C.c_int32), _ptr(self.coords, C.c_int32),
                _ptr(self.collisions, C.c_uint8))

    def count_neighbors(self, x, y, points=None):
        x, y = self._coords(x), self._coords(y)
        p = None if points is None else np.ascontiguousarray(points, dtype=np.int64)
        out = np.zeros(x.shape[0], dtype=np.int64)
        self._fn("pno_hash_count_neighbors")(*self._table(), _ptr(x, self.real),
                                             C.c_int64(x.shape[0]), _ptr(y, self.real),
                                             _ptr(p, C.c_int64), C.c_int64(0 if p is None else p.size),
                                             _ptr(out, C.c_int64))
        return out

    def neighbor_lists(self, x, y, sort=True):
        x, y = self._coords(x), self._coords(y)
        offsets = np.zeros(x.shape[0] + 1, dtype=np.int64)
        fn = self._fn("pno_hash_neighbor_lists")
        args = (*self._table(), _ptr(x, self.real), C.c_int64(x.shape[0]), _ptr(y, self.real),
                _ptr(offsets, C.c_int64))
        fn(*args, None, C.c_int(int(sort)))
        ids = np.zeros(max(int(offsets[-1]), 1), dtype=np.int32)
        fn(*args, _ptr(ids, C.c_int32), C.c_int(int(sort)))
        return offsets, ids[:int(offsets[-1])]

    def nbody(self, x, y, mass, G, wide=False):
        x, y = self._coords(x), self._coords(y)
        mass = np.ascontiguousarray(mass, dtype=self.dtype)
        dv = np.zeros_like(x)
        dv64 = np.zeros(x.shape, dtype=np.float64) if wide else None
        dvabs = np.zeros(x.shape, dtype=np.float64) if wide else None
        self._fn("pno_hash_nbody")(*self._table(), _ptr(x, self.real), C.c_int64(x.shape[0]),
                                   _ptr(y, self.real), _ptr(mass, self.real), self.real(G),
                                   _ptr(dv, self.real), _ptr(dv64, C.c_double),
                                   _ptr(dvabs, C.c_double))
        return (dv, dv64, dvabs) if wide else dv


class _GridMix(C.Structure):
    _fields_ = [("ndims", C.c_int32), ("periodic", C.c_int32), ("search_radius", C.c_float),
                ("min_corner", C.c_double * 3), ("max_corner", C.c_double * 3),
                ("grid_size", C.c_int64 * 3), ("n_cells", C.c_int64 * 3),
                ("cell_size", C.c_float * 3), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3),
                ("box_size", C.c_float * 3)]


class MixedGrid:
    """Float64 coordinates / corners with a Float32 search radius (and Float32 PeriodicBox):
    oracle/pn_oracle_mixed.h (docs/literate/src/tut_gpu_usage.jl:45-50)."""

    def __init__(self, ndims, search_radius, min_corner, max_corner, periodic_box=None):
        self.g = _GridMix()
        self.ndims = int(ndims)
        mn = np.ascontiguousarray(min_corner, dtype=np.float64)
        mx = np.ascontiguousarray(max_corner, dtype=np.float64)
        if periodic_box is not None:
            bmn = np.ascontiguousarray(periodic_box[0], dtype=np.float32)
            bmx = np.ascontiguousarray(periodic_box[1], dtype=np.float32)
        else:
            bmn = bmx = None
        rc = lib().pno_grid_init_mix(C.byref(self.g), self.ndims, C.c_float(search_radius),
                                     _ptr(mn, C.c_double), _ptr(mx, C.c_double),
                                     int(periodic_box is not None), _ptr(bmn, C.c_float),
                                     _ptr(bmx, C.c_float))
        if rc:
            raise OracleError(rc)
        self.cell_start = self.cell_points = None

    @property
    def min_corner(self):
        return np.array(self.g.min_corner[:self.ndims])

    @property
    def max_corner(self):
        return np.array(self.g.max_corner[:self.ndims])

    @property
    def grid_size(self):
        return tuple(int(v) for v in self.g.grid_size[:self.ndims])

    @property
    def n_cells(self):
        return tuple(int(v) for v in self.g.n_cells[:self.ndims])

    @property
    def cell_size(self):
        return np.array(self.g.cell_size[:self.ndims], dtype=np.float32)

    @property
    def total_cells(self):
        return int(lib().pno_total_cells_mix(C.byref(self.g)))

    def _coords(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.ndim == 2 and x.shape[1] == self.ndims
        return x

    def point_cells(self, x):
        x = self._coords(x)
        out = np.empty(x.shape[0], dtype=np.int64)
        lib().pno_point_cells_mix(C.byref(self.g), _ptr(x, C.c_double), C.c_int64(x.shape[0]),
                                  _ptr(out, C.c_int64))
        return out

    def build(self, y):
        y = self._coords(y)
        self.cell_start = np.zeros(self.total_cells + 1, dtype=np.int64)
        self.cell_points = np.zeros(max(y.shape[0], 1), dtype=np.int32)
        rc = lib().pno_build_csr_mix(C.byref(self.g), _ptr(y, C.c_double), C.c_int64(y.shape[0]),
                                     _ptr(self.cell_start, C.c_int64),
                                     _ptr(self.cell_points, C.c_int32))
        if rc:
            raise OracleError(rc)
        self.cell_points = self.cell_points[:y.shape[0]]
        return self

    def neighbor_lists(self, x, y, brute=False, sort=False):
        x, y = self._coords(x), self._coords(y)
        offsets = np.zeros(x.shape[0] + 1, dtype=np.int64)
        args = (C.byref(self.g), _ptr(self.cell_start, C.c_int64), _ptr(self.cell_points, C.c_int32),
                _ptr(x, C.c_double), C.c_int64(x.shape[0]), _ptr(y, C.c_double),
                C.c_int64(y.shape[0]), _ptr(offsets, C.c_int64))
        rc = lib().pno_neighbor_lists_mix(*args, None, int(brute))
        if rc:
            raise OracleError(rc)
        ids = np.zeros(max(int(offsets[-1]), 1), dtype=np.int32)
        lib().pno_neighbor_lists_mix(*args, _ptr(ids, C.c_int32), int(brute))
        ids = ids[:int(offsets[-1])]
        if sort:
            for i in range(x.shape[0]):
                ids[offsets[i]:offsets[i + 1]].sort()
        return offsets, ids

    def list_pairs(self, x, y, offsets, ids):
        x, y = self._coords(x), self._coords(y)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        P = int(offsets[-1])
        pd = np.zeros((max(P, 1), self.ndims), dtype=np.float32)
        dist = np.zeros(max(P, 1), dtype=np.float32)
        lib().pno_list_pairs_mix(C.byref(self.g), _ptr(x, C.c_double), C.c_int64(x.shape[0]),
                                 _ptr(y, C.c_double), _ptr(offsets, C.c_int64), _ptr(ids, C.c_int32),
                                 _ptr(pd, C.c_float), _ptr(dist, C.c_float))
        return pd[:P], dist[:P]

    # fused closures (Float32 closure arithmetic on the mixed pair geometry)
    def _points(self, points):
        if points is None:
            return None, 0
        pts = np.ascontiguousarray(points, dtype=np.int64)
        return pts, pts.shape[0]

    def nbody(self, x, y, mass, G, points=None):
        x, y = self._coords(x), self._coords(y)
        mass = np.ascontiguousarray(mass, dtype=np.float32)
        dv = np.zeros((x.shape[0], self.ndims), dtype=np.float32)
        pts, npts = self._points(points)
        rc = lib().pno_nbody_mix(C.byref(self.g), _ptr(self.cell_start, C.c_int64),
                                 _ptr(self.cell_points, C.c_int32), _ptr(x, C.c_double),
                                 C.c_int64(x.shape[0]), _ptr(y, C.c_double), _ptr(pts, C.c_int64),
                                 C.c_int64(npts), _ptr(mass, C.c_float), C.c_float(G), _ptr(dv, C.c_float))
        if rc:
            raise OracleError(rc)
        return dv

    def wcsph(self, x, y, v_x, v_y, mass_x, mass_y, pressure_x, pressure_y, params, points=None):
        x, y = self._coords(x), self._coords(y)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        v_x, v_y, mass_x, mass_y, pressure_x, pressure_y, params = (
            f(v_x), f(v_y), f(mass_x), f(mass_y), f(pressure_x), f(pressure_y), f(params))
        dv = np.zeros((x.shape[0], self.ndims + 1), dtype=np.float32)
        pts, npts = self._points(points)
        rc = lib().pno_wcsph_mix(C.byref(self.g), _ptr(self.cell_start, C.c_int64),
                                 _ptr(self.cell_points, C.c_int32), _ptr(x, C.c_double),
                                 C.c_int64(x.shape[0]), _ptr(y, C.c_double), _ptr(pts, C.c_int64),
                                 C.c_int64(npts), _ptr(v_x, C.c_float), _ptr(v_y, C.c_float),
                                 _ptr(mass_x, C.c_float), _ptr(mass_y, C.c_float),
                                 _ptr(pressure_x, C.c_float), _ptr(pressure_y, C.c_float),
                                 _ptr(params, C.c_float), _ptr(dv, C.c_float))
        if rc:
            raise OracleError(rc)
        return dv


def trivial_lists(x, y, search_radius, periodic_box=None, dtype=np.float32):
    """Brute force (TrivialNeighborhoodSearch) neighbour lists as CSR, ascending ids."""
    dtype = np.dtype(dtype)
    suf, real = ("_f32", C.c_float) if dtype == np.float32 else ("_f64", C.c_double)
    x = np.ascontiguousarray(x, dtype=dtype)
    y = np.ascontiguousarray(y, dtype=dtype)
    nd = x.shape[1]
    if periodic_box is not None:
        bmn = np.ascontiguousarray(periodic_box[0], dtype=dtype)
        bmx = np.ascontiguousarray(periodic_box[1], dtype=dtype)
    else:
        bmn = bmx = None
    offsets = np.zeros(x.shape[0] + 1, dtype=np.int64)
    fn = getattr(lib(), "pno_trivial_lists" + suf)
    args = (nd, real(search_radius), int(periodic_box is not None), _ptr(bmn, real),
            _ptr(bmx, real), _ptr(x, real), C.c_int64(x.shape[0]), _ptr(y, real),
            C.c_int64(y.shape[0]), _ptr(offsets, C.c_int64))
    fn(*args, None)
    ids = np.zeros(max(int(offsets[-1]), 1), dtype=np.int32)
    fn(*args, _ptr(ids, C.c_int32))
    return offsets, ids[:int(offsets[-1])]


def list_pairs(x, y, offsets, ids, search_radius, periodic_box=None, dtype=np.float32):
    """What the closure receives when sweeping precomputed lists: (pos_diff[P, nd], dist[P])."""
    dtype = np.dtype(dtype)
    suf, real = ("_f32", C.c_float) if dtype == np.float32 else ("_f64", C.c_double)
    x = np.ascontiguousarray(x, dtype=dtype)
    y = np.ascontiguousarray(y, dtype=dtype)
    nd = x.shape[1]
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    if periodic_box is not None:
        bmn = np.ascontiguousarray(periodic_box[0], dtype=dtype)
        bmx = np.ascontiguousarray(periodic_box[1], dtype=dtype)
    else:
        bmn = bmx = None
    P = int(offsets[-1])
    pd = np.zeros((max(P, 1), nd), dtype=dtype)
    dist = np.zeros(max(P, 1), dtype=dtype)
    getattr(lib(), "pno_list_pairs" + suf)(nd, real(search_radius), int(periodic_box is not None),
                                           _ptr(bmn, real), _ptr(bmx, real), _ptr(x, real),
                                           C.c_int64(x.shape[0]), _ptr(y, real), _ptr(offsets, C.c_int64),
                                           _ptr(ids, C.c_int32), _ptr(pd, real), _ptr(dist, real))
    return pd[:P], dist[:P]


def tlsph_deformation_grad(X0, xcur, offsets, ids, mass, rho0, L, h, kernel_norm, search_radius,
                           periodic_box=None, dtype=np.float32, wide=False):
    dtype = np.dtype(dtype)
    suf, real = ("_f32", C.c_float) if dtype == np.float32 else ("_f64", C.c_double)
    X0 = np.ascontiguousarray(X0, dtype=dtype)
    xcur = np.ascontiguousarray(xcur, dtype=dtype)
    n, nd = X0.shape
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    mass = np.ascontiguousarray(mass, dtype=dtype)
    rho0 = np.ascontiguousarray(rho0, dtype=dtype)
    L = np.ascontiguousarray(L, dtype=dtype)
    assert L.shape == (n, nd * nd)
    if periodic_box is not None:
        bmn = np.ascontiguousarray(periodic_box[0], dtype=dtype)
        bmx = np.ascontiguousarray(periodic_box[1], dtype=dtype)
    else:
        bmn = bmx = None
    F = np.zeros((n, nd * nd), dtype=dtype)
    F64 = np.zeros(F.shape, dtype=np.float64) if wide else None
    Fabs = np.zeros(F.shape, dtype=np.float64) if wide else None
    getattr(lib(), "pno_tlsph_deformation_grad" + suf)(
        nd, real(search_radius), int(periodic_box is not None), _ptr(bmn, real), _ptr(bmx, real),
        _ptr(X0, real), _ptr(xcur, real), C.c_int64(n), _ptr(offsets, C.c_int64), _ptr(ids, C.c_int32),
        _ptr(mass, real), _ptr(rho0, real), _ptr(L, real), real(h), real(kernel_norm),
        _ptr(F, real), _ptr(F64, C.c_double), _ptr(Fabs, C.c_double))
    return (F, F64, Fabs) if wide else F


def tlsph_pk1_corrected(F, L, young_modulus, poisson_ratio, dtype=np.float32):
    """compute_pk1_corrected! (unpinned; oracle pno_tlsph_pk1_corrected): (n, nd*nd) column-major."""
    dtype = np.dtype(dtype)
    suf, real = ("_f32", C.c_float) if dtype == np.float32 else ("_f64", C.c_double)
    F = np.ascontiguousarray(F, dtype=dtype)
    L = np.ascontiguousarray(L, dtype=dtype)
    n, nn = F.shape
    nd = int(round(nn ** 0.5))
    out = np.zeros_like(F)
    getattr(lib(), "pno_tlsph_pk1_corrected" + suf)(nd, C.c_int64(n), _ptr(F, real), _ptr(L, real),
                                                     real(young_modulus), real(poisson_ratio),
                                                     _ptr(out, real))
    return out


def tlsph_interact(X0, xcur, offsets, ids, mass, rho0, pk1c, F, h, kernel_norm, young_modulus,
                   alpha, search_radius, periodic_box=None, dtype=np.float32, wide=False):
    """interact_structure_structure! over precomputed lists (unpinned; oracle pno_tlsph_interact)."""
    dtype = np.dtype(dtype)
    suf, real = ("_f32", C.c_float) if dtype == np.float32 else ("_f64", C.c_double)
    X0 = np.ascontiguousarray(X0, dtype=dtype)
    xcur = np.ascontiguousarray(xcur, dtype=dtype)
    n, nd = X0.shape
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    mass = np.ascontiguousarray(mass, dtype=dtype)
    rho0 = np.ascontiguousarray(rho0, dtype=dtype)
    pk1c = np.ascontiguousarray(pk1c, dtype=dtype)
    F = np.ascontiguousarray(F, dtype=dtype)
    assert pk1c.shape == (n, nd * nd) and F.shape == (n, nd * nd)
    params = np.array([h, kernel_norm, young_modulus, alpha], dtype=dtype)
    if periodic_box is not None:
        bmn = np.ascontiguousarray(periodic_box[0], dtype=dtype)
        bmx = np.ascontiguousarray(periodic_box[1], dtype=dtype)
    else:
        bmn = bmx = None
    dv = np.zeros((n, nd), dtype=dtype)
    dv64 = np.zeros(dv.shape, dtype=np.float64) if wide else None
    dvabs = np.zeros(dv.shape, dtype=np.float64) if wide else None
    getattr(lib(), "pno_tlsph_interact" + suf)(
        nd, real(search_radius), int(periodic_box is not None), _ptr(bmn, real), _ptr(bmx, real),
        _ptr(X0, real), _ptr(xcur, real), C.c_int64(n), _ptr(offsets, C.c_int64),
        _ptr(ids, C.c_int32), _ptr(mass, real), _ptr(rho0, real), _ptr(pk1c, real), _ptr(F, real),
        _ptr(params, real), _ptr(dv, real), _ptr(dv64, C.c_double), _ptr(dvabs, C.c_double))
    return (dv, dv64, dvabs) if wide else dv


def wcsph_compute_pressure(v, sound_speed, reference_density, exponent=1.0, background_pressure=0.0,
                           dtype=np.float32):
    """compute_pressure! with StateEquationCole (unpinned; oracle pno_wcsph_compute_pressure)."""
    dtype = np.dtype(dtype)
    suf, real = ("_f32", C.c_float) if dtype == np.float32 else ("_f64", C.c_double)
    v = np.ascontiguousarray(v, dtype=dtype)
    n, ns = v.shape
    out = np.zeros(n, dtype=dtype)
    getattr(lib(), "pno_wcsph_compute_pressure" + suf)(ns - 1, C.c_int64(n), _ptr(v, real),
                                                        real(sound_speed), real(reference_density),
                                                        real(exponent), real(background_pressure),
                                                        _ptr(out, real))
    return out


def periodic_coords(x, box_min, box_max, dtype=np.float32):
    dtype = np.dtype(dtype)
    suf, real = ("_f32", C.c_float) if dtype == np.float32 else ("_f64", C.c_double)
    x = np.ascontiguousarray(x, dtype=dtype)
    bmn = np.ascontiguousarray(box_min, dtype=dtype)
    bmx = np.ascontiguousarray(box_max, dtype=dtype)
    out = np.zeros_like(x)
    getattr(lib(), "pno_periodic_coords" + suf)(x.size, _ptr(bmn, real), _ptr(bmx, real),
                                                _ptr(x, real), _ptr(out, real))
    return out
