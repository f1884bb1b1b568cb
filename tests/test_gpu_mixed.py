"""GPU parity of mixed-precision searches: Float64 coordinates and corners
with a Float32 search radius (docs/literate/src/tut_gpu_usage.jl:45-50) against
oracle/pn_oracle_mixed.h.  Bar: cells, cell list, counts, neighbour lists and the pair geometry
(Float32 pos_diff / distance) bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from test_gpu_parity import dev, lattice, pn  # noqa: E402,F401


def make_mixed(pn, nd, r, mn, mx, box=None, n_points=0):
    cl = pn.FullGridCellList(min_corner=np.asarray(mn, np.float64), max_corner=np.asarray(mx, np.float64),
                             search_radius=np.float32(r), mixed_precision=True)
    pb = None if box is None else pn.PeriodicBox(min_corner=np.asarray(box[0], np.float32),
                                                 max_corner=np.asarray(box[1], np.float32))
    return pn.GridNeighborhoodSearch[nd](search_radius=np.float32(r), n_points=n_points,
                                         periodic_box=pb, cell_list=cl)


def test_gpu_tutorial_mixed(pn, kats, oracle):
    """test/gpu.jl:35 with the tutorial's Float64 coordinates and Float32 radius: (11, 29)."""
    k = kats["gpu_tutorial_count"]
    coords = lattice(k["lattice"], dtype=np.float64)
    nhs = make_mixed(pn, 2, k["search_radius"], coords.min(0), coords.max(0), n_points=len(coords))
    assert nhs.eltype == np.float64
    x = dev(coords)
    assert x.dtype == torch.float64
    pn.initialize_(nhs, x, x)
    cnt = torch.zeros(len(coords), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
    assert [int(cnt.min()), int(cnt.max())] == k["expected_extrema"]
    with pytest.raises(TypeError):
        pn.initialize_(nhs, x.float(), x.float())     # Float32 coordinates do not match this search


@pytest.mark.parametrize("nd", [1, 2, 3])
@pytest.mark.parametrize("periodic", [False, True])
def test_mixed_bit_exact(pn, oracle, nd, periodic):
    rng = np.random.default_rng(40 + nd)
    n = 3000
    # coordinates far from the origin: Float32 coordinates would lose the digits that decide
    # cells and neighbours, which is why SPH codes keep them in Float64
    shift = 1000.0
    y = rng.random((n, nd)) * 2 - 1 + shift
    xq = rng.random((257, nd)) * 2 - 1 + shift
    r = np.float32(0.13 if nd == 3 else 0.06)
    mn, mx = np.full(nd, shift - 1.0), np.full(nd, shift + 1.0)
    box = (mn.astype(np.float32), mx.astype(np.float32)) if periodic else None
    nhs = make_mixed(pn, nd, r, mn, mx, box=box, n_points=n)
    og = oracle.MixedGrid(nd, r, mn, mx, periodic_box=box)
    assert nhs.cell_list.n_cells_per_dimension == og.grid_size
    assert np.array_equal(nhs.cell_list.min_corner, og.min_corner)
    if periodic:
        assert nhs.n_cells == og.n_cells
        assert np.array_equal(np.array(nhs.cell_size, np.float32), og.cell_size)
    ty, tx = dev(y), dev(xq)
    pn.initialize_(nhs, ty, ty)
    og.build(y)
    assert np.array_equal(nhs.point_cells(ty).cpu().numpy(), og.point_cells(y))
    cs, cp = nhs.export_csr()
    assert np.array_equal(cs.cpu().numpy(), og.cell_start)
    assert np.array_equal(cp.cpu().numpy(), og.cell_points)
    for q, tq in ((y, ty), (xq, tx)):
        if q is xq and not periodic:
            # query points must keep their stencil inside the grid
            q = np.clip(q, mn, mx)
            tq = dev(q)
        roff, rids = og.neighbor_lists(q, y)
        boff, bids = og.neighbor_lists(q, y, brute=True)
        so, si = og.neighbor_lists(q, y, sort=True)
        assert np.array_equal(so, boff) and np.array_equal(si, bids)       # = brute force
        cnt = torch.zeros(len(q), dtype=torch.int64, device="cuda")
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), tq, ty, nhs)
        assert np.array_equal(cnt.cpu().numpy(), np.diff(roff))
        lists = pn.api._NeighborLists.build(nhs, tq, ty, sort=False)
        off, ids = (t.cpu().numpy() for t in lists.export_csr(0))
        assert np.array_equal(off, roff) and np.array_equal(ids, rids)     # reference visiting order
        pd, dist = lists.pairs(nhs, tq, ty)
        rpd, rdist = og.list_pairs(q, y, roff, rids)
        pd, dist = pd.cpu().numpy(), dist.cpu().numpy()
        # the library returns the Float32 values widened to Float64
        assert np.array_equal(pd.astype(np.float32).astype(np.float64), pd)
        assert np.array_equal(pd.astype(np.float32), rpd) and np.array_equal(dist.astype(np.float32), rdist)


def test_mixed_from_padded_corners(pn, oracle):
    """The Julia glue adapts a FullGridCellList from its STORED (already padded) corners:
    pnb_grid_create_padded_mixed must give the same grid as the constructor from user corners."""
    import ctypes as C
    L = pn._lib.lib()
    rng = np.random.default_rng(4)
    y = rng.random((2000, 3)) * 3 + 50.0
    r = np.float32(0.21)
    mn, mx = y.min(0), y.max(0)
    nhs = make_mixed(pn, 3, r, mn, mx, n_points=len(y))
    cl = nhs.cell_list
    h = C.c_void_p()
    pmn = np.ascontiguousarray(cl.min_corner, dtype=np.float64)
    pmx = np.ascontiguousarray(cl.max_corner, dtype=np.float64)
    pn._lib.check(L.pnb_grid_create_padded_mixed(3, r, pmn.ctypes.data_as(pn._lib._pd),
                                                 pmx.ctypes.data_as(pn._lib._pd), None, None,
                                                 C.byref(h)))
    try:
        assert L.pnb_grid_total_cells(h) == nhs.total_cells() == int(np.prod(cl.n_cells_per_dimension))
        ty = dev(y)
        a = torch.empty(len(y), dtype=torch.int32, device="cuda")
        pn._lib.check(L.pnb_point_cells_f64(h, ty.data_ptr(), len(y), a.data_ptr(), None))
        assert torch.equal(a, nhs.point_cells(ty))
        og = oracle.MixedGrid(3, r, mn, mx)
        assert np.array_equal(a.cpu().numpy(), og.point_cells(y))
    finally:
        L.pnb_grid_destroy(h)


@pytest.mark.parametrize("nd", [2, 3])
@pytest.mark.parametrize("periodic", [False, True])
def test_mixed_fused_closures(pn, oracle, nd, periodic):
    """n-body and WCSPH closures on a mixed-precision search: Float32 pos_diff / distance reach the
    closure (nhs_grid.jl:547-555), Float32 state, Float32 closure arithmetic -- sums bit-identical
    to the mixed oracle (np.array_equal), with a `points` subset and two point sets, beta != 0."""
    rng = np.random.default_rng(70 + nd)
    n, T = 2500, np.float32
    shift = 1000.0
    y = rng.random((n, nd)) * 2 - 1 + shift
    xq = rng.random((301, nd)) * 2 - 1 + shift
    r = T(0.16 if nd == 3 else 0.07)
    mn, mx = np.full(nd, shift - 1.0), np.full(nd, shift + 1.0)
    box = (mn.astype(T), mx.astype(T)) if periodic else None
    nhs = make_mixed(pn, nd, r, mn, mx, box=box, n_points=n)
    og = oracle.MixedGrid(nd, r, mn, mx, periodic_box=box)
    ty, tx = dev(y), dev(xq)
    pn.initialize_(nhs, tx, ty)
    og.build(y)
    mass = (T(0.5) + rng.random(n).astype(T)).astype(T)
    # ---- n-body: two point sets, then a subset of the points --------------------------------
    dv = torch.full((len(xq), nd), 7.0, dtype=torch.float32, device="cuda")
    pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), T(1.5)), tx, ty, nhs)
    assert np.array_equal(dv.cpu().numpy(), og.nbody(xq, y, mass, T(1.5)))
    pts = rng.choice(len(xq), 77, replace=False)
    dv.fill_(3.0)
    pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), T(1.5)), tx, ty, nhs, points=pts)
    assert np.array_equal(dv.cpu().numpy(), og.nbody(xq, y, mass, T(1.5), points=pts))
    # ---- WCSPH, x === y -------------------------------------------------------------------------
    pn.initialize_(nhs, ty, ty)
    v = np.concatenate([rng.normal(0, 0.1, (n, nd)), 1000.0 + rng.random((n, 1))], axis=1).astype(T)
    pres = (T(100.0) * (v[:, nd] - T(1000.0))).astype(T)
    dvw = torch.zeros((n, nd + 1), dtype=torch.float32, device="cuda")
    f = pn.WCSPHInteract(dvw, dev(v), dev(v), dev(mass), dev(mass), dev(pres), dev(pres),
                         smoothing_length=T(r / T(2)), sound_speed=T(10.0), alpha=T(0.02), beta=T(0.3),
                         delta=T(0.1), ndims_=nd)
    pn.foreach_point_neighbor(f, ty, ty, nhs)
    ref = og.wcsph(y, y, v, v, mass, mass, pres, pres, f.params_array())
    assert np.array_equal(dvw.cpu().numpy(), ref)
    assert np.abs(ref).max() > 0
    # Float64 state on a mixed search is a type error, not a silent conversion
    with pytest.raises(TypeError):
        pn.foreach_point_neighbor(pn.NBodyGravity(dv.double(), dev(mass).double(), 1.5), tx, ty, nhs)
