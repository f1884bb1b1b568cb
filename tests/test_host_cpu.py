"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host-side
constructor arithmetic equals the oracle's (and hence the reference's KATs), the host mirror
raises the reference's errors, and the product refuses to compute without a CUDA device.
No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pn():
    import pnb200
    pnb200.build()
    return pnb200


def test_library_exports_every_declared_symbol(pn):
    header = open(os.path.join(REPO, "include", "pnb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pnb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = C.CDLL(pn.library_path())
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/pnb200.h but not exported"
    # and the ctypes binding covers all of them
    assert declared == set(pn._lib.SIGNATURES)
    assert pn._lib.lib().pnb_version() == 100


def test_library_is_sm100a_cuda(pn):
    """The .so carries sm_100a SASS (no PTX-only / other-arch fallback)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", pn.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def _params(pn, nd, r, mn, mx, box=None):
    L = pn._lib.lib()
    pf = pn._lib._pf
    mn = np.ascontiguousarray(mn, np.float32)
    mx = np.ascontiguousarray(mx, np.float32)
    pmin, pmax = (C.c_float * 3)(), (C.c_float * 3)()
    gs, nc = (C.c_int64 * 3)(), (C.c_int64 * 3)()
    cs = (C.c_float * 3)()
    bmn = bmx = None
    if box is not None:
        b0 = np.ascontiguousarray(box[0], np.float32)
        b1 = np.ascontiguousarray(box[1], np.float32)
        bmn, bmx = b0.ctypes.data_as(pf), b1.ctypes.data_as(pf)
    st = L.pnb_grid_params_f32(nd, np.float32(r), mn.ctypes.data_as(pf), mx.ctypes.data_as(pf),
                               bmn, bmx, pmin, pmax, gs, nc, cs)
    return st, (np.array(pmin[:nd], np.float32), np.array(pmax[:nd], np.float32),
                tuple(gs[:nd]), tuple(nc[:nd]), np.array(cs[:nd], np.float32))


def test_grid_params_match_oracle(pn, oracle):
    """pnb_grid_params_f32 is a pure host function: compare bit for bit with the oracle on the
    benchmark configurations and on random boxes."""
    expect = {64: 24, 101: 37, 254: 88, 504: 171}
    for n, cells in expect.items():
        r = np.float32(3.0) / np.float32(n + 1)
        st, (pmin, pmax, gs, nc, cs) = _params(pn, 3, r, np.zeros(3), np.ones(3))
        assert st == 0 and gs == (cells,) * 3 and nc == (-1,) * 3
        og = oracle.Grid(3, r, np.zeros(3), np.ones(3))
        assert og.grid_size == gs
        assert np.array_equal(og.min_corner, pmin) and np.array_equal(og.max_corner, pmax)
    rng = np.random.default_rng(0)
    for _ in range(300):
        nd = int(rng.integers(1, 4))
        mn = rng.normal(0, 3, nd).astype(np.float32)
        mx = (mn + rng.uniform(0.5, 4, nd)).astype(np.float32)
        r = np.float32(rng.uniform(0.02, 0.15))
        periodic = bool(rng.integers(0, 2))
        box = (mn, mx) if periodic else None
        st, (pmin, pmax, gs, nc, cs) = _params(pn, nd, r, mn, mx, box)
        og = oracle.Grid(nd, r, mn, mx, periodic_box=box)
        assert st == 0
        assert og.grid_size == gs and og.n_cells == nc
        assert np.array_equal(og.min_corner, pmin) and np.array_equal(og.max_corner, pmax)
        assert np.array_equal(og.cell_size, cs)


def test_grid_params_periodic_c4(pn, oracle):
    """200^3 lattice with a box of exactly 200 spacings -> 66 periodic cells in
    a 69^3 allocated grid; Float32 box 1 with r = 0.1f0 -> 9 cells (App. A.3)."""
    T = np.float32
    n = 200
    s = T(1) / T(n + 1)
    r = T(3) / T(n + 1)
    bmn = np.full(3, s / T(2), T)
    bmx = np.full(3, (T(n) + T(0.5)) * s, T)
    st, (_, _, gs, nc, cs) = _params(pn, 3, r, bmn, bmx, (bmn, bmx))
    assert st == 0 and nc == (66, 66, 66) and gs == (69, 69, 69)
    st, (_, _, gs, nc, cs) = _params(pn, 2, T(0.1), np.zeros(2), np.ones(2), (np.zeros(2), np.ones(2)))
    assert nc == (9, 9)
    # fewer than 3 periodic cells -> ArgumentError text of nhs_grid.jl:120-124
    st, _ = _params(pn, 2, T(0.4), np.zeros(2), np.ones(2), (np.zeros(2), np.ones(2)))
    assert st == pn._lib.PNB_ERR_ARG
    assert "needs at least 3 cells in each dimension" in pn._lib.last_error()


def test_constructor_errors(pn, kats):
    """test/nhs_grid.jl:2-19, test/cell_lists/full_grid.jl:3-16."""
    e = kats["error_texts"]
    with pytest.raises(pn.ArgumentError, match="FullGridCellList only supports up to 100 dimensions"):
        pn.FullGridCellList(min_corner=np.zeros(101), max_corner=np.ones(101))
    with pytest.raises(pn.ArgumentError, match=e["corner_length"]):
        pn.FullGridCellList(min_corner=np.zeros(3), max_corner=np.ones(2))
    cl3 = pn.FullGridCellList(min_corner=np.zeros(3), max_corner=np.ones(3))
    with pytest.raises(pn.ArgumentError, match="a 2D cell list is required for a GridNeighborhoodSearch\\{2\\}"):
        pn.GridNeighborhoodSearch[2](cell_list=cl3)
    cl2 = pn.FullGridCellList(min_corner=np.zeros(2), max_corner=np.ones(2))
    with pytest.raises(pn.ArgumentError, match="is not a valid update strategy"):
        pn.GridNeighborhoodSearch[2](cell_list=cl2, update_strategy="test")
    with pytest.raises(pn.ArgumentError, match="is not a valid update strategy"):
        pn.GridNeighborhoodSearch[2](cell_list=cl2, update_strategy=pn.SemiParallelUpdate())
    with pytest.raises(pn.ArgumentError, match="`search_radius` cannot be an integer type"):
        pn.GridNeighborhoodSearch[2](cell_list=cl2, search_radius=1)
    with pytest.raises(pn.ArgumentError, match="must have the same element type"):
        pn.GridNeighborhoodSearch[2](
            cell_list=pn.FullGridCellList(min_corner=np.zeros(2), max_corner=np.ones(2),
                                          search_radius=np.float32(0.1)),
            search_radius=np.float32(0.1),
            periodic_box=pn.PeriodicBox(min_corner=np.zeros(2), max_corner=np.ones(2)))
    with pytest.raises(pn.ArgumentError, match="`search_radius` cannot be an integer type"):
        pn.PrecomputedNeighborhoodSearch[2](search_radius=1)
    assert pn.PrecomputedNeighborhoodSearch.default_max_neighbors(3) == 320
    assert pn.PrecomputedNeighborhoodSearch.default_max_neighbors(2) == 64


def test_copy_neighborhood_search(pn):
    """test/nhs_grid.jl:21-60: strategy and max_points_per_cell survive; default is ParallelUpdate."""
    mn, mx = np.zeros(2, np.float32), np.ones(2, np.float32)
    nhs = pn.GridNeighborhoodSearch[2](cell_list=pn.FullGridCellList(min_corner=mn, max_corner=mx))
    assert nhs.update_strategy == pn.ParallelUpdate()
    assert nhs.n_cells == (-1, -1)
    assert pn.requires_update(nhs) == (False, True)
    cp = pn.copy_neighborhood_search(nhs, np.float32(1.0), 10)
    assert pn.ndims(cp) == 2 and pn.search_radius(cp) == np.float32(1.0)
    assert isinstance(cp.cell_list, pn.FullGridCellList)
    assert cp.update_strategy == pn.ParallelUpdate()
    # template corners are unpadded, the copy is padded by 1.001 r (full_grid.jl:66-67,179-185)
    assert np.array_equal(nhs.cell_list.min_corner, mn)
    assert np.array_equal(cp.cell_list.min_corner, mn - np.float32(1001.0 / 1000.0 * 1.0))
    nhs = pn.GridNeighborhoodSearch[2](
        cell_list=pn.FullGridCellList(min_corner=mn, max_corner=mx, max_points_per_cell=101),
        update_strategy=pn.SerialUpdate())
    cp = pn.copy_neighborhood_search(nhs, np.float32(1.0), 10)
    assert cp.update_strategy == pn.SerialUpdate()
    assert cp.cell_list.max_points_per_cell == 101
    pre = pn.PrecomputedNeighborhoodSearch[2](update_neighborhood_search=nhs, max_neighbors=77,
                                              transpose_backend=True, sort_neighbor_lists=False)
    assert pn.requires_update(pre) == (True, True)
    cp = pn.copy_neighborhood_search(pre, np.float32(0.5), 27)
    assert cp.max_neighbors == 77 and cp.transpose_backend and not cp.sort_neighbor_lists
    assert pn.search_radius(cp.neighborhood_search) == np.float32(0.5)
    assert pn.freeze_neighborhood_search(cp).neighborhood_search is None
    assert pn.freeze_neighborhood_search(nhs) is nhs


def test_padded_corners_kat(pn):
    """test/cell_lists/full_grid.jl:28-29 in Float32: min = 0 - 1.001f0*1, max = 10 + 1.001f0*1."""
    cl = pn.FullGridCellList(min_corner=np.zeros(3, np.float32), max_corner=np.full(3, 10, np.float32),
                             search_radius=np.float32(1.0))
    f = np.float32(1001) / np.float32(1000)
    assert np.array_equal(cl.min_corner, np.full(3, np.float32(0) - f, np.float32))
    assert np.array_equal(cl.max_corner, np.full(3, np.float32(10) + f, np.float32))
    assert cl.n_cells_per_dimension == (13, 13, 13)


def test_no_cpu_fallback(pn):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = pn._lib.lib()
    assert L.pnb_device_count() == 0
    nhs = pn.GridNeighborhoodSearch[3](
        search_radius=np.float32(0.1),
        cell_list=pn.FullGridCellList(min_corner=np.zeros(3, np.float32),
                                      max_corner=np.ones(3, np.float32),
                                      search_radius=np.float32(0.1)))
    with pytest.raises(pn._lib.CudaError, match="no CUDA device"):
        nhs._grid()
    y = torch.zeros((4, 3), dtype=torch.float32)
    with pytest.raises(TypeError, match="no CPU path"):
        pn.initialize_(nhs, y, y)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under pointneighbors.jl_b200/ may reference it."""
    root = os.path.join(REPO, "pointneighbors.jl_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for line in text.splitlines():
                    s = line.strip()
                    if s.startswith(("import ", "from ", "#include", "using ", "include(")):
                        assert "oracle" not in s, f"{f}: {s}"
                assert "libpn_oracle" not in text and "pn_oracle" not in text.replace(
                    "oracle/pn_oracle_impl.h", "").replace("oracle pno_", ""), f


def test_point_cloud_generator(pn):
    """test/point_cloud.jl:4-58: size, lattice + two perturbations, cell-sorted, dim 1 major."""
    c = pn.point_cloud((9, 10, 7), 2.5, seed=1)
    assert c.shape == (630, 3) and c.dtype == np.float64
    assert np.all(c.min(0) > 0.5) and np.all(c.max(0) < np.array([9.5, 10.5, 7.5]))
    assert not np.array_equal(c, pn.point_cloud((9, 10, 7), 2.5, seed=2))
    assert np.array_equal(c, pn.point_cloud((9, 10, 7), 2.5, seed=1))
    cells = np.floor(c / 2.5).astype(int)
    key = cells[:, 0] * 10000 + cells[:, 1] * 100 + cells[:, 2]
    # sorted by the once-perturbed cell; the second perturbation moves ~few % across cells
    assert np.mean(np.diff(key) < 0) < 0.25
    coords, r, mn, mx = pn.benchmark_cloud((8, 6, 4))
    assert coords.dtype == np.float32 and r == np.float32(3.0) / np.float32(9)
    assert np.array_equal(mx, np.array([1.0, 0.75, 0.5], np.float32))


def test_grid_params_f64_match_oracle(pn, oracle):
    """pnb_grid_params_f64 (host arithmetic of the Float64 constructors) against the Float64
    oracle on random boxes, incl. the reference's Float64 KAT -1.001 / 11.001
    (test/cell_lists/full_grid.jl:28-29) and the element-type rule of the host mirror."""
    import ctypes as C
    L = pn._lib.lib()
    pd = pn._lib._pd

    def params64(nd, r, mn, mx, box=None):
        mn = np.ascontiguousarray(mn, np.float64)
        mx = np.ascontiguousarray(mx, np.float64)
        pmin, pmax, cs = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
        gs, nc = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        bmn = bmx = None
        if box is not None:
            b0 = np.ascontiguousarray(box[0], np.float64)
            b1 = np.ascontiguousarray(box[1], np.float64)
            bmn, bmx = b0.ctypes.data_as(pd), b1.ctypes.data_as(pd)
        st = L.pnb_grid_params_f64(nd, float(r), mn.ctypes.data_as(pd), mx.ctypes.data_as(pd), bmn, bmx,
                                   pmin, pmax, gs, nc, cs)
        return st, (np.array(pmin[:nd]), np.array(pmax[:nd]), tuple(gs[:nd]), tuple(nc[:nd]),
                    np.array(cs[:nd]))

    st, (pmin, pmax, gs, _, _) = params64(2, 1.0, [0.0, 0.0], [10.0, 10.0])
    assert st == 0 and pmin.tolist() == [-1.001, -1.001] and pmax.tolist() == [11.001, 11.001]
    rng = np.random.default_rng(5)
    for _ in range(300):
        nd = int(rng.integers(1, 4))
        mn = rng.normal(0, 3, nd)
        mx = mn + rng.uniform(0.5, 4, nd)
        r = np.float64(rng.uniform(0.02, 0.15))
        box = (mn, mx) if rng.integers(0, 2) else None
        st, (pmin, pmax, gs, nc, cs) = params64(nd, r, mn, mx, box)
        og = oracle.Grid(nd, r, mn, mx, periodic_box=box, dtype=np.float64)
        assert st == 0
        assert og.grid_size == gs and og.n_cells == nc
        assert np.array_equal(og.min_corner, pmin) and np.array_equal(og.max_corner, pmax)
        assert np.array_equal(og.cell_size, cs)
    # host mirror: Float64 only when radius AND corners are Float64
    mn, mx = np.zeros(3), np.ones(3)
    assert pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=np.float64(0.1)).eltype == np.float64
    assert pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=np.float32(0.1)).eltype == np.float32
    assert pn.FullGridCellList(min_corner=mn.astype(np.float32), max_corner=mx.astype(np.float32),
                               search_radius=np.float64(0.1)).eltype == np.float32
    cl = pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=np.float64(0.1))
    nhs = pn.GridNeighborhoodSearch[3](search_radius=np.float64(0.1), cell_list=cl)
    assert nhs.eltype == np.float64 and pn.search_radius(nhs) == np.float64(0.1)


# ---------------------------------------------------------------------------------------------
# SURVEY 8f ranks 2-4 added after the first checkpoint: host logic of the hashed cell list, the
# mixed-precision constructors and properties of the TLSPH oracle definitions
# ---------------------------------------------------------------------------------------------
def test_spatial_hashing_host_logic():
    import pnb200 as pn
    from oracle import pn_oracle
    rng = np.random.default_rng(0)
    for nd in (1, 2, 3):
        oh = pn_oracle.HashGrid(nd, np.float32(0.1), 1000)
        for _ in range(200):
            cell = [int(v) for v in rng.integers(-2 ** 31, 2 ** 31 - 1, nd)]
            L = int(rng.integers(1, 10 ** 7))
            oh.list_size = L
            assert pn.spatial_hash(cell, L) == oh.spatial_hash(cell) + 1     # 1-based like Julia
    # the reference's own example: cells (-1, 0) and (-2, -1) collide in a table of 2
    assert pn.spatial_hash((-1, 0), 2) == pn.spatial_hash((-2, -1), 2)
    with pytest.raises(pn.ArgumentError):
        pn.SpatialHashingCellList[4](list_size=10)
    with pytest.raises(pn.ArgumentError):
        pn.SpatialHashingCellList[2](list_size=0)
    cl = pn.SpatialHashingCellList[3](list_size=77, max_points_per_cell=11)
    with pytest.raises(pn.ArgumentError, match="a 2D cell list is required"):
        pn.GridNeighborhoodSearch[2](search_radius=np.float32(0.1), cell_list=cl)
    box = pn.PeriodicBox(min_corner=np.zeros(3, np.float32), max_corner=np.ones(3, np.float32))
    nhs = pn.GridNeighborhoodSearch[3](search_radius=np.float32(0.1), n_points=5, cell_list=cl,
                                       periodic_box=box)
    oh = pn_oracle.HashGrid(3, np.float32(0.1), 77, periodic_box=(np.zeros(3, np.float32),
                                                                  np.ones(3, np.float32)))
    assert nhs.n_cells == oh.n_cells == (9, 9, 9)       # the Float64 `10eps()` quirk, as for the full grid
    assert np.array_equal(np.array(nhs.cell_size, np.float32), oh.cell_size)
    cp = pn.copy_neighborhood_search(nhs, np.float32(0.2), 9)
    assert isinstance(cp.cell_list, pn.SpatialHashingCellList)
    assert cp.cell_list.list_size == 77 and cp.cell_list.max_points_per_cell == 11
    assert cp.n_cells == (4, 4, 4)
    assert pn.requires_update(nhs) == (False, True)


def test_mixed_precision_constructor_arithmetic():
    import pnb200 as pn
    from oracle import pn_oracle
    rng = np.random.default_rng(3)
    for _ in range(200):
        nd = int(rng.integers(1, 4))
        mn = rng.normal(0, 100, nd)
        mx = mn + rng.random(nd) * 3 + 0.5
        r = np.float32(rng.random() * 0.15 + 0.01)
        periodic = bool(rng.integers(0, 2))
        box = (mn.astype(np.float32), mx.astype(np.float32)) if periodic else None
        try:
            og = pn_oracle.MixedGrid(nd, r, mn, mx, periodic_box=box)
        except pn_oracle.OracleError:
            continue
        cl = pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=r, mixed_precision=True)
        assert cl.eltype == np.float64 and cl.mixed
        assert np.array_equal(cl.min_corner, og.min_corner) and np.array_equal(cl.max_corner, og.max_corner)
        assert cl.n_cells_per_dimension == og.grid_size
        pb = None if box is None else pn.PeriodicBox(min_corner=box[0], max_corner=box[1])
        nhs = pn.GridNeighborhoodSearch[nd](search_radius=r, cell_list=cl, periodic_box=pb)
        assert nhs.eltype == np.float64
        if periodic:
            assert nhs.n_cells == og.n_cells
            assert np.array_equal(np.array(nhs.cell_size, np.float32), og.cell_size)
    with pytest.raises(pn.ArgumentError):
        pn.FullGridCellList(min_corner=np.zeros(2), max_corner=np.ones(2),
                            search_radius=np.float64(0.1), mixed_precision=True)


def test_tlsph_oracle_definitions():
    """Properties of the repo's definition of the TLSPH right-hand side (TrixiParticles is not
    vendored: parity unpinned): rigid rotations carry no stress, uniaxial stretch matches the
    St. Venant-Kirchhoff closed form, and the pair forces conserve linear momentum."""
    from oracle import pn_oracle as po
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    eye = np.eye(3).ravel()[None]
    assert np.abs(po.tlsph_pk1_corrected(R.T.ravel()[None], eye, 1.4e6, 0.4, dtype=np.float64)).max() < 1e-9
    lam, mu = 1.4e6 * 0.4 / (1.4 * 0.2), 1.4e6 / 2.8
    P = po.tlsph_pk1_corrected(np.diag([1.1, 1, 1]).T.ravel()[None], eye, 1.4e6, 0.4,
                               dtype=np.float64)[0].reshape(3, 3).T
    assert abs(P[0, 0] - 1.1 * (lam + 2 * mu) * 0.105) < 1e-6 and abs(P[1, 1] - lam * 0.105) < 1e-6
    # momentum conservation on a small random cloud (all pairs within the radius listed)
    rng = np.random.default_rng(2)
    n = 400
    X0 = rng.random((n, 3))
    xc = X0 + 0.01 * np.sin(6 * X0)
    r = 0.25
    off, ids = po.trivial_lists(X0, X0, r, dtype=np.float64)
    mass = 0.1 * (1 + 0.1 * rng.random(n))
    rho = 1000 + rng.random(n)
    L = np.tile(np.eye(3).ravel(), (n, 1))
    h = r / 2
    kn = 21 / (16 * np.pi) / h ** 3
    F = po.tlsph_deformation_grad(X0, xc, off, ids, mass, rho, L, h, kn, r, dtype=np.float64)
    pk1 = po.tlsph_pk1_corrected(F, L, 1.4e6, 0.4, dtype=np.float64)
    dv, dv64, dvabs = po.tlsph_interact(X0, xc, off, ids, mass, rho, pk1, F, h, kn, 1.4e6, 0.1, r,
                                        dtype=np.float64, wide=True)
    mom = (mass[:, None] * dv).sum(0)
    assert np.all(np.abs(mom) <= 1e-10 * (mass[:, None] * dvabs).sum(0))
    assert np.abs(dv).max() > 0
    # compute_pressure!: Cole with exponent 1 is linear in the density
    v = np.concatenate([np.zeros((5, 3)), 1000 + np.arange(5)[:, None]], axis=1)
    assert np.allclose(po.wcsph_compute_pressure(v, 10.0, 1000.0, dtype=np.float64), 100.0 * np.arange(5))


def test_ctypes_signatures_match_the_header(pn):
    """Every entry of pnb200/_lib.py's SIGNATURES has the arity of its C declaration, pointers
    where C has pointers, and scalars of the C width -- a stale binding (a parameter added in the
    header only) would otherwise shift every following argument silently."""
    import test_julia_glue as tj
    decls = tj.header_decls()
    scalar = {"int": (C.c_int,), "float": (C.c_float,), "double": (C.c_double,),
              "int64_t": (C.c_int64, C.c_longlong, C.c_long), "int32_t": (C.c_int32, C.c_int),
              "uint64_t": (C.c_uint64, C.c_ulonglong, C.c_ulong), "uint32_t": (C.c_uint32, C.c_uint)}
    checked = 0
    for name, (restype, argtypes) in pn._lib.SIGNATURES.items():
        cret, cparams = decls[name]
        assert len(argtypes) == len(cparams), f"{name}: {len(argtypes)} ctypes arguments, {len(cparams)} in C"
        if cret == "void":
            assert restype is None, name
        elif cret in ("pnb_status", "int"):
            assert restype is C.c_int, name
        for k, (at, cp) in enumerate(zip(argtypes, cparams)):
            is_ptr_c = "*" in cp
            is_ptr_py = at is C.c_void_p or at is C.c_char_p or hasattr(at, "_type_") and \
                isinstance(getattr(at, "_type_"), type)
            if is_ptr_c:
                assert is_ptr_py, f"{name} argument {k + 1}: C `{cp}` is a pointer, ctypes has {at}"
            else:
                base = cp.replace("const ", "").split()[0]
                assert at in scalar[base], f"{name} argument {k + 1}: C `{cp}` vs ctypes {at}"
            checked += 1
    assert checked > 400


def test_api_surface_mirrors_the_reference_exports(pn):
    """Every name src/PointNeighbors.jl exports that is in scope (SURVEY.md sections 2 / 8) exists in
    the host mirror (`!` -> trailing underscore); what is out of scope is absent, not stubbed."""
    in_scope = ["foreach_point_neighbor", "foreach_point_neighbor_unsafe", "foreach_neighbor",
                "foreach_neighbor_unsafe", "mapreduce_neighbor", "mapreduce_neighbor_unsafe",
                "GridNeighborhoodSearch", "PrecomputedNeighborhoodSearch", "FullGridCellList",
                "SpatialHashingCellList", "DynamicVectorOfVectors", "ParallelUpdate", "SemiParallelUpdate",
                "SerialIncrementalUpdate", "SerialUpdate", "ParallelIncrementalUpdate", "requires_update",
                "initialize_", "update_", "initialize_grid_", "update_grid_", "default_backend",
                "PeriodicBox", "copy_neighborhood_search"]
    for name in in_scope:
        assert hasattr(pn, name), name
    for name in ("TrivialNeighborhoodSearch", "DictionaryCellList", "PolyesterBackend", "SerialBackend"):
        assert not hasattr(pn, name), name          # CPU-only parts of the reference
    with pytest.raises(pn.ArgumentError):
        pn.default_backend(np.zeros((3, 2), np.float32))
