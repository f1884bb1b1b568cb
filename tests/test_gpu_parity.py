"""Parity of the CUDA path (through the C ABI / host mirror) against the CPU oracle and the
reference's golden vectors.  Needs a B200: every test is marked gpu.

Bar (BASELINE.json north_star): cell assignments and neighbour sets bit-exact; summed
forces/densities within 1e-5 relative (tolerance written at each assert).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def pn():
    import pnb200
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    assert pnb200._lib.lib().pnb_device_count() >= 1
    return pnb200


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def make_grid(pn, nd, r, mn, mx, box=None, n_points=0):
    cl = pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=np.float32(r))
    pb = None if box is None else pn.PeriodicBox(min_corner=np.asarray(box[0], np.float32),
                                                 max_corner=np.asarray(box[1], np.float32))
    return pn.GridNeighborhoodSearch[nd](search_radius=np.float32(r), n_points=n_points,
                                         periodic_box=pb, cell_list=cl)


def lattice(dims, dtype=np.float32):
    grids = np.meshgrid(*[np.arange(1, n + 1) for n in dims], indexing="ij")
    return np.stack([g.ravel(order="F") for g in grids], axis=1).astype(dtype)


# ------------------------------------------------------------------------------------------
# golden vectors of the reference
# ------------------------------------------------------------------------------------------
def test_gpu_tutorial_extrema(pn, kats):
    """test/gpu.jl:35: extrema(n_neighbors_gpu) == (11, 29)."""
    k = kats["gpu_tutorial_count"]
    coords = lattice(k["lattice"])
    nhs = make_grid(pn, 2, k["search_radius"], coords.min(0), coords.max(0), n_points=len(coords))
    x = dev(coords)
    pn.initialize_(nhs, x, x)
    n_neighbors = torch.zeros(len(coords), dtype=torch.int64, device="cuda")
    assert pn.foreach_point_neighbor(pn.CountNeighbors(n_neighbors), x, x, nhs) is None
    assert [int(n_neighbors.min()), int(n_neighbors.max())] == k["expected_extrema"]


@pytest.mark.parametrize("case", [0, 1, 2])
@pytest.mark.parametrize("copied", [False, True])
def test_periodic_neighbors(pn, kats, oracle, case, copied):
    """test/neighborhood_search.jl:59-184, in Float32 (the coordinates are far from any
    rounding boundary, so the expected lists hold in Float32 too; checked with the oracle)."""
    k = kats["periodic_neighbors"]
    c = k["cases"][case]
    coords = np.array(c["coordinates_rows"], dtype=np.float32).T.copy()
    nd = coords.shape[1]
    r = np.float32(k["search_radius"])
    bmn, bmx = np.array(c["box_min"], np.float32), np.array(c["box_max"], np.float32)
    try:
        oracle.Grid(nd, r, bmn, bmx, periodic_box=(bmn, bmx))
    except oracle.OracleError as exc:
        # Float32 quirk of nhs_grid.jl:117 (Float64 `10eps()` on a Float32 box):
        # 0.35f0 - 0.05f0 < 3 * 0.1f0, so the reference itself rejects this box in Float32.
        assert exc.code == 2 and case == 2
        with pytest.raises(pn.ArgumentError, match="needs at least 3 cells in each dimension"):
            make_grid(pn, nd, r, bmn, bmx, box=(bmn, bmx))
        bmx = bmx.copy()
        bmx[2] = np.float32(0.3500001)   # derived case: widen z by 1 ulp-ish so Float32 gets 3 cells
    if copied:
        box = pn.PeriodicBox(min_corner=bmn, max_corner=bmx)
        template = pn.GridNeighborhoodSearch[nd](periodic_box=box, cell_list=pn.FullGridCellList(
            min_corner=bmn, max_corner=bmx))
        nhs = pn.copy_neighborhood_search(template, r, len(coords))
    else:
        nhs = make_grid(pn, nd, r, bmn, bmx, box=(bmn, bmx), n_points=len(coords))
    x = dev(coords)
    pn.initialize_(nhs, x, x)
    neighbors = [[] for _ in range(len(coords))]
    pn.foreach_point_neighbor(lambda i, j, pos_diff, d: neighbors[i].append(j + 1), x, x, nhs,
                              points=range(len(coords)))
    assert [sorted(v) for v in neighbors] == k["expected_neighbors"]
    # PrecomputedNeighborhoodSearch on top of the same grid
    pre = pn.PrecomputedNeighborhoodSearch[nd](search_radius=r, n_points=len(coords),
                                               periodic_box=nhs.periodic_box,
                                               update_neighborhood_search=nhs)
    pn.initialize_(pre, x, x)
    neighbors = [[] for _ in range(len(coords))]
    pn.foreach_point_neighbor(lambda i, j, pos_diff, d: neighbors[i].append(j + 1), x, x, pre)
    assert neighbors == k["expected_neighbors"]     # sorted lists are deterministic


def test_lattice3d_eachindex_y(pn, kats, oracle):
    """test/nhs_grid.jl:197-225 in Float32; candidates compared with the Float32 oracle."""
    k = kats["lattice3d_eachindex_y"]
    rng_ = np.array(k["range"])
    a, b, c = np.meshgrid(rng_, rng_, rng_, indexing="ij")
    coords1 = np.stack([a.ravel(order="F"), b.ravel(order="F"), c.ravel(order="F")], axis=1)
    coords2 = (coords1 + np.array(k["shift"])).astype(np.float32)
    coords1 = coords1.astype(np.float32)
    mn = np.minimum(coords1.min(0), coords2.min(0))
    mx = np.maximum(coords1.max(0), coords2.max(0))
    r = np.float32(k["search_radius"])
    nhs = make_grid(pn, 3, r, mn, mx, n_points=len(coords1))
    x1, x2 = dev(coords1), dev(coords2)
    pn.initialize_(nhs, x1, x1)
    lo, hi = k["eachindex_y"]
    idx = np.arange(lo - 1, hi)
    pn.update_(nhs, x2, x2, eachindex_y=idx)
    cs, cp = nhs.export_csr()
    og = oracle.Grid(3, r, mn, mx)
    og.build(coords2, eachindex_y=idx)
    assert (cs.cpu().numpy() == og.cell_start).all()
    assert (cp.cpu().numpy() == og.cell_points).all()
    assert sorted(cp.cpu().numpy().tolist()) == list(range(lo - 1, hi))


@pytest.mark.parametrize("nd", [1, 2, 3])
def test_full_grid_bounds(pn, kats, nd):
    """test/cell_lists/full_grid.jl:19-66 in Float32: NaN / outside -> the reference's error."""
    k = kats["full_grid_bounds"]
    mn = np.zeros(nd, np.float32)
    mx = np.full(nd, 10.0, np.float32)
    nhs = make_grid(pn, nd, 1.0, mn, mx)
    y = np.random.default_rng(0).random((k["n_points"], nd)).astype(np.float32)
    for bad in k["error_values"]:
        y[k["bad_point"] - 1, 0] = np.float32(float(bad))
        t = dev(y)
        with pytest.raises(pn.PointNeighborsError, match=k["error_text"]):
            pn.initialize_(nhs, t, t)
        with pytest.raises(pn.PointNeighborsError, match=k["error_text"]):
            pn.update_(nhs, t, t)
    for ok in k["ok_values"]:
        y[k["bad_point"] - 1, 0] = np.float32(ok)
        t = dev(y)
        pn.initialize_(nhs, t, t)
        pn.update_(nhs, t, t)


def test_periodic_face_rounding(pn, kats, oracle):
    """test/neighborhood_search.jl:4-57 (Float32): cell of prevfloat/nextfloat face points."""
    k = kats["periodic_face_rounding"]
    T = np.float32
    zero, one, half, inf = T(0), T(1), T(0.5), T(np.inf)
    vals = [np.nextafter(zero, -inf), zero, np.nextafter(zero, inf),
            np.nextafter(one, -inf), one, np.nextafter(one, inf)]
    pts = np.array([(v, half) for v in vals] + [(half, v) for v in vals], dtype=T)
    box = (np.array(k["box_min"], T), np.array(k["box_max"], T))
    nhs = make_grid(pn, 2, T(k["search_radius"]), box[0], box[1], box=box)
    assert nhs.n_cells == (9, 9)
    og = oracle.Grid(2, T(k["search_radius"]), box[0], box[1], periodic_box=box)
    cells = nhs.point_cells(dev(pts)).cpu().numpy()
    assert (cells == og.point_cells(pts)).all()
    assert (cells >= 0).all()


# ------------------------------------------------------------------------------------------
# oracle parity on seeded perturbed lattices
# ------------------------------------------------------------------------------------------
CLOUDS = [((10, 11), 2.5), ((100, 90), 2.5), ((9, 10, 7), 2.5), ((39, 40, 41), 2.5), ((300,), 2.5)]


def _cloud(pn, size, r, seed):
    c = pn.point_cloud(size, r, seed=seed).astype(np.float32)
    return c


@pytest.mark.parametrize("size,r", CLOUDS)
@pytest.mark.parametrize("seed", [1, 2])
def test_cells_and_neighbor_sets_bit_exact(pn, oracle, size, r, seed):
    """test/neighborhood_search.jl:186-337: initialize! with seed 1, update! with seed 2; cell
    assignment, CSR cell list and sorted neighbour lists must equal the oracle's bit for bit, and
    the oracle equals brute force."""
    r = np.float32(r)
    coords = _cloud(pn, size, r, seed)
    coords_init = _cloud(pn, size, r, 1)
    nd = len(size)
    mn = np.minimum(coords.min(0), coords_init.min(0)) - r
    mx = np.maximum(coords.max(0), coords_init.max(0)) + r
    nhs = make_grid(pn, nd, r, mn, mx, n_points=len(coords))
    x0, x = dev(coords_init), dev(coords)
    pn.initialize_(nhs, x0, x0)
    if seed != 1:
        pn.update_(nhs, x, x)
    og = oracle.Grid(nd, r, mn, mx)
    og.build(coords)
    assert nhs.cell_list.n_cells_per_dimension == og.grid_size
    assert (nhs.point_cells(x).cpu().numpy() == og.point_cells(coords)).all()
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all()
    assert (cp.cpu().numpy() == og.cell_points).all()
    # neighbour counts (fused kernel)
    cnt = torch.zeros(len(coords), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
    off_o, ids_o = og.neighbor_lists(coords, coords, sort=True)
    assert (cnt.cpu().numpy() == np.diff(off_o)).all()
    # sorted neighbour lists
    pre = pn.PrecomputedNeighborhoodSearch[nd](search_radius=r, n_points=len(coords),
                                               update_neighborhood_search=nhs)
    pn.update_(pre, x, x) if seed != 1 else pn.initialize_(pre, x, x)
    off, ids = pre.export_csr()
    assert (off.cpu().numpy() == off_o).all()
    assert (ids.cpu().numpy() == ids_o).all()
    off_t, ids_t = oracle.trivial_lists(coords, coords, r)
    assert (off_o == off_t).all() and (ids_o == ids_t).all()


def test_two_point_sets_and_points_subset(pn, oracle):
    """x != y (docs tut_basic_usage.jl:69-85) and the `points` keyword (neighborhood_search.jl:185)."""
    r = np.float32(2.5)
    y = _cloud(pn, (20, 18, 16), r, 3)
    rng = np.random.default_rng(5)
    x = (y[rng.choice(len(y), 777, replace=False)] +
         rng.normal(0, 0.3, (777, 3))).astype(np.float32)
    mn, mx = y.min(0) - 2 * r, y.max(0) + 2 * r
    nhs = make_grid(pn, 3, r, mn, mx)
    tx, ty = dev(x), dev(y)
    pn.initialize_(nhs, tx, ty)
    og = oracle.Grid(3, r, mn, mx)
    og.build(y)
    cnt = torch.full((len(x),), -7, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), tx, ty, nhs)
    assert (cnt.cpu().numpy() == og.count_neighbors(x, y)).all()
    pts = np.arange(5, 700, 7)
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), tx, ty, nhs, points=pts)
    assert (cnt.cpu().numpy() == og.count_neighbors(x, y, points=pts)).all()
    # a query point whose stencil leaves the grid -> BoundsError like the safe variant
    x_bad = x.copy()
    x_bad[3] = mx + 10 * r
    with pytest.raises(pn.BoundsError):
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), dev(x_bad), ty, nhs)


def _nbody_inputs(n, seed=11):
    rng = np.random.default_rng(seed)
    mass = (np.float32(1e10) * (rng.random(n).astype(np.float32) + np.float32(1))).astype(np.float32)
    return mass, np.float32(6.6743e-11)


@pytest.fixture
def arithmetic(pn, request):
    """exact=True: IEEE operation sequence of the oracle; exact=False: the default fast terms."""
    pn.set_exact_arithmetic(request.param)
    yield request.param
    pn.set_exact_arithmetic(False)


@pytest.mark.parametrize("arithmetic", [False, True], indirect=True, ids=["fast", "exact"])
@pytest.mark.parametrize("size", [(24, 24, 24), (40, 30), (200,)])
def test_nbody_parity(pn, oracle, size, arithmetic):
    """benchmarks/n_body.jl closure.  Tolerance: |gpu - ref64| <= 1e-5 * sum_j |term_ij| per
    component and global relative L2 <= 1e-5 (SURVEY.md section 7 hard parts).  In exact mode
    the kernel visits pairs in the oracle's order with the oracle's operations, so the Float32
    sums must be IDENTICAL."""
    nd = len(size)
    c, r, mn, mx = pn.benchmark_cloud(size, seed=4)
    mass, G = _nbody_inputs(len(c))
    nhs = make_grid(pn, nd, r, mn, mx)
    x = dev(c)
    pn.initialize_(nhs, x, x)
    dv = torch.full((len(c), nd), 3.0, dtype=torch.float32, device="cuda")
    pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), G), x, x, nhs)
    og = oracle.Grid(nd, r, mn, mx)
    og.build(c)
    ref, ref64, refabs = og.nbody(c, c, mass, G, wide=True)
    got = dv.cpu().numpy()
    assert np.all(np.abs(got - ref64) <= 1e-5 * refabs + 1e-30)
    assert np.linalg.norm(got - ref64) <= 1e-5 * np.linalg.norm(ref64)
    if arithmetic:
        assert np.array_equal(got, ref)
    # general path (points subset) gives the same numbers for the looped points
    pts = np.arange(0, len(c), 3)
    dv2 = torch.zeros_like(dv)
    pn.foreach_point_neighbor(pn.NBodyGravity(dv2, dev(mass), G), x, x, nhs, points=pts)
    got2 = dv2.cpu().numpy()
    assert np.all(np.abs(got2[pts] - ref64[pts]) <= 1e-5 * refabs[pts] + 1e-30)
    if arithmetic:
        assert np.array_equal(got2[pts], ref[pts])
    assert not got2[np.setdiff1d(np.arange(len(c)), pts)].any()   # dv .= 0 for the others


def _wcsph_inputs(pn, c, r, nd, seed=21, moving=True):
    rng = np.random.default_rng(seed)
    n = len(c)
    T = np.float32
    rho = (T(1000.0) + rng.random(n).astype(T)).astype(T)          # density + rand (:60-62)
    vel = (rng.normal(0, 0.1, (n, nd)).astype(T) if moving else np.zeros((n, nd), T))
    v = np.concatenate([vel, rho[:, None]], axis=1).astype(T)       # vcat(velocity, density') (:92)
    spacing = T(r / T(3))
    mass = np.full(n, T(0.1) * spacing, dtype=T)                    # mass = 0.1 * particle_spacing (:56)
    c0 = T(10.0)
    pressure = (c0 * c0 * (rho - T(1000.0))).astype(T)              # StateEquationCole, exponent 1
    h = T(r / T(2))
    return v, mass, pressure, dict(smoothing_length=h, sound_speed=c0, alpha=T(0.02), beta=T(0.0),
                                   epsilon=T(0.01), delta=T(0.1), ndims_=nd)


@pytest.mark.parametrize("arithmetic", [False, True], indirect=True, ids=["fast", "exact"])
@pytest.mark.parametrize("size", [(24, 24, 24), (40, 30)])
@pytest.mark.parametrize("moving", [True, False])
def test_wcsph_parity(pn, oracle, size, moving, arithmetic):
    """WCSPH continuity + momentum (formulas: oracle pno_cl_wcsph; TrixiParticles parity is
    unpinned).  Tolerance 1e-5 as for n-body; in exact mode bit identity with the oracle's Float32 sums."""
    nd = len(size)
    c, r, mn, mx = pn.benchmark_cloud(size, seed=6)
    v, mass, pressure, kw = _wcsph_inputs(pn, c, r, nd, moving=moving)
    nhs = make_grid(pn, nd, r, mn, mx)
    x = dev(c)
    pn.initialize_(nhs, x, x)
    dv = torch.full((len(c), nd + 1), 5.0, dtype=torch.float32, device="cuda")
    tv, tm, tp = dev(v), dev(mass), dev(pressure)
    f = pn.WCSPHInteract(dv, tv, tv, tm, tm, tp, tp, **kw)
    pn.foreach_point_neighbor(f, x, x, nhs)
    og = oracle.Grid(nd, r, mn, mx)
    og.build(c)
    ref, ref64, refabs = og.wcsph(c, c, v, v, mass, mass, pressure, pressure, f.params_array(),
                                  wide=True)
    got = dv.cpu().numpy()
    assert np.all(np.abs(got - ref64) <= 1e-5 * refabs + 1e-30)
    assert np.linalg.norm(got - ref64) <= 1e-5 * np.linalg.norm(ref64)
    if arithmetic:
        assert np.array_equal(got, ref)
    pts = np.arange(1, len(c), 5)
    dv2 = torch.zeros_like(dv)
    f2 = pn.WCSPHInteract(dv2, tv, tv, tm, tm, tp, tp, **kw)
    pn.foreach_point_neighbor(f2, x, x, nhs, points=pts)
    got2 = dv2.cpu().numpy()
    assert np.all(np.abs(got2[pts] - ref64[pts]) <= 1e-5 * refabs[pts] + 1e-30)
    if arithmetic:
        assert np.array_equal(got2[pts], ref[pts])


def _periodic_case(pn, n, nd, seed=2):
    """SURVEY.md 8d C4 in small: lattice n^d with spacing s, box = n*s so it is exactly periodic."""
    T = np.float32
    s = T(1.0) / T(n + 1)
    r = T(3.0) / T(n + 1)
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.arange(1, n + 1) for _ in range(nd)], indexing="ij")
    lat = np.stack([g.ravel(order="F") for g in grids], axis=1).astype(np.float64)
    lat += rng.normal(0, 0.0707, lat.shape)
    c = (lat / (n + 1)).astype(T)
    bmn = np.full(nd, s / T(2), T)
    bmx = np.full(nd, (T(n) + T(0.5)) * s, T)
    # keep every point strictly inside the box
    c = np.clip(c, bmn + T(1e-6), bmx - T(1e-6)).astype(T)
    return c, r, bmn, bmx


@pytest.mark.parametrize("nd,n", [(3, 24), (2, 40), (1, 100)])
def test_periodic_lists_and_tlsph(pn, oracle, nd, n):
    """PeriodicBox + PrecomputedNeighborhoodSearch: lists bit-exact, pair geometry bit-exact,
    TLSPH deformation gradient within 1e-5 (warp-tree summation order differs from the oracle)."""
    c, r, bmn, bmx = _periodic_case(pn, n, nd)
    nhs = make_grid(pn, nd, r, bmn, bmx, box=(bmn, bmx))
    og = oracle.Grid(nd, r, bmn, bmx, periodic_box=(bmn, bmx))
    assert nhs.n_cells == og.n_cells
    assert tuple(np.float32(v) for v in nhs.cell_size) == tuple(og.cell_size)
    x = dev(c)
    pre = pn.PrecomputedNeighborhoodSearch[nd](search_radius=r, n_points=len(c),
                                               periodic_box=nhs.periodic_box,
                                               update_neighborhood_search=nhs, max_neighbors=128,
                                               transpose_backend=True)
    pn.initialize_(pre, x, x)
    og.build(c)
    off_o, ids_o = og.neighbor_lists(c, c, sort=True)
    off, ids = pre.export_csr()
    assert (off.cpu().numpy() == off_o).all() and (ids.cpu().numpy() == ids_o).all()
    off_t, ids_t = oracle.trivial_lists(c, c, r, periodic_box=(bmn, bmx))
    assert (off_o == off_t).all() and (ids_o == ids_t).all()
    # what the closure receives (nhs_precomputed.jl:230-238)
    pd, dist = pre._lists.pairs(nhs, x, x)
    pd_o, dist_o = oracle.list_pairs(c, c, off_o, ids_o, r, periodic_box=(bmn, bmx))
    assert np.array_equal(pd.cpu().numpy(), pd_o) and np.array_equal(dist.cpu().numpy(), dist_o)
    assert float(dist.max()) <= float(r)
    # the reference's layouts (test/nhs_precomputed.jl:9-37)
    backend_t, lengths = pre.neighbor_lists(index_base=1)
    assert backend_t.shape == (128, len(c))
    lengths_h = lengths.cpu().numpy()
    assert (lengths_h == np.diff(off_o)).all()
    bt = backend_t.cpu().numpy()
    for i in (0, len(c) // 2, len(c) - 1):
        assert (bt[:lengths_h[i], i] == ids_o[off_o[i]:off_o[i + 1]] + 1).all()
        assert (bt[lengths_h[i]:, i] == np.iinfo(np.int32).max).all()
    backend_r, _ = pre._lists.export_dvov(128, transposed=False, index_base=1)
    assert np.array_equal(backend_r.cpu().numpy(), bt.T)
    if nd == 1:
        return
    # TLSPH deformation gradient
    rng = np.random.default_rng(9)
    T = np.float32
    disp = (0.01 * np.sin(2 * np.pi * c)).astype(T)
    xcur = (c + disp * r).astype(T)
    mass = np.full(len(c), 0.1, T)
    rho0 = np.full(len(c), 1000.0, T)
    Lm = (np.eye(nd, dtype=T)[None] + 0.05 * rng.normal(size=(len(c), nd, nd))).astype(T)
    Lm = Lm.reshape(len(c), nd * nd)
    h = T(r / T(2))
    F = torch.zeros((len(c), nd * nd), dtype=torch.float32, device="cuda")
    f = pn.TLSPHDeformationGradient(F, dev(xcur), dev(mass), dev(rho0), dev(Lm),
                                    smoothing_length=h, ndims_=nd)
    pn.foreach_point_neighbor(f, x, x, pre)
    ref, ref64, refabs = oracle.tlsph_deformation_grad(c, xcur, off_o, ids_o, mass, rho0, Lm, h,
                                                       f.kernel_norm, r, periodic_box=(bmn, bmx),
                                                       wide=True)
    got = F.cpu().numpy()
    assert np.all(np.abs(got - ref64) <= 1e-5 * refabs + 1e-30)
    assert np.linalg.norm(got - ref64) <= 1e-5 * np.linalg.norm(ref64)


def test_edge_cases(pn, oracle):
    """SURVEY.md Appendix E."""
    T = np.float32
    mn, mx = np.zeros(3, T), np.ones(3, T)
    # template / unused search: zero radius is a legal no-op (nhs_grid.jl:263-267)
    tmpl = pn.GridNeighborhoodSearch[3](cell_list=pn.FullGridCellList(min_corner=mn, max_corner=mx))
    y = dev(np.random.default_rng(0).random((50, 3)).astype(T))
    pn.initialize_(tmpl, y, y)
    cnt = torch.ones(50, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), y, y, tmpl)
    assert int(cnt.sum()) == 0
    # empty y: every neighbourhood is empty (test/neighborhood_search.jl:445-465)
    nhs = make_grid(pn, 3, 0.1, mn, mx)
    empty = torch.zeros((0, 3), dtype=torch.float32, device="cuda")
    pn.initialize_(nhs, y, empty)
    cnt = torch.ones(50, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), y, empty, nhs)
    assert int(cnt.sum()) == 0
    # points_moving = (true, false): update! is a no-op for the grid search (nhs_grid.jl:289)
    pn.initialize_(nhs, y, y)
    cs0, cp0 = nhs.export_csr()
    y2 = dev(np.random.default_rng(1).random((50, 3)).astype(T))
    pn.update_(nhs, y2, y2, points_moving=(True, False))
    cs1, cp1 = nhs.export_csr()
    assert torch.equal(cs0, cs1) and torch.equal(cp0, cp1)
    # cells with more than 32 points (multi-pass warps) and more than max_points_per_cell
    dense = (0.5 + 0.01 * np.random.default_rng(2).random((300, 3))).astype(T)
    td = dev(dense)
    pn.initialize_(nhs, td, td)
    cnt = torch.zeros(300, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), td, td, nhs)
    og = oracle.Grid(3, T(0.1), mn, mx)
    og.build(dense)
    assert (cnt.cpu().numpy() == og.count_neighbors(dense, dense)).all()
    with pytest.raises(pn.PointNeighborsError, match="cell list is full"):
        nhs.export_dvov(max_points_per_cell=100)
    backend, lengths = nhs.export_dvov(max_points_per_cell=400, index_base=1)
    assert int(lengths.sum()) == 300
    # a tile denser than the staging buffer of the tile kernel (> 1920 candidates): overflow path
    rngd = np.random.default_rng(12)
    dense2 = np.concatenate([(0.5 + 0.05 * rngd.random((3000, 3))), rngd.random((2000, 3))]).astype(T)
    td2 = dev(dense2)
    pn.initialize_(nhs, td2, td2)
    cnt2 = torch.zeros(len(dense2), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt2), td2, td2, nhs)
    og2 = oracle.Grid(3, T(0.1), mn, mx)
    og2.build(dense2)
    assert (cnt2.cpu().numpy() == og2.count_neighbors(dense2, dense2)).all()
    mass2 = (1e10 * (1 + rngd.random(len(dense2)))).astype(T)
    dvd = torch.zeros((len(dense2), 3), dtype=torch.float32, device="cuda")
    pn.foreach_point_neighbor(pn.NBodyGravity(dvd, dev(mass2), T(6.6743e-11)), td2, td2, nhs)
    _, r64, rabs = og2.nbody(dense2, dense2, mass2, T(6.6743e-11), wide=True)
    assert np.all(np.abs(dvd.cpu().numpy() - r64) <= 1e-5 * rabs + 1e-30)
    # long neighbour lists: > 512 neighbours (global odd-even sort) and 128..512 (shared memory
    # rank sort) must come out ascending and equal to the oracle's
    pre_big = pn.PrecomputedNeighborhoodSearch[3](search_radius=T(0.1), n_points=len(dense2),
                                                  update_neighborhood_search=nhs, max_neighbors=4000)
    pn.initialize_(pre_big, td2, td2)
    off_b, ids_b = pre_big.export_csr()
    off_o2, ids_o2 = og2.neighbor_lists(dense2, dense2, sort=True)
    assert (off_b.cpu().numpy() == off_o2).all() and (ids_b.cpu().numpy() == ids_o2).all()
    assert int(np.diff(off_o2).max()) > 512
    pn.initialize_(nhs, td, td)
    pre_mid = pn.PrecomputedNeighborhoodSearch[3](search_radius=T(0.1), n_points=300,
                                                  update_neighborhood_search=nhs, max_neighbors=640)
    pn.initialize_(pre_mid, td, td)
    off_m, ids_m = pre_mid.export_csr()
    off_o1, ids_o1 = og.neighbor_lists(dense, dense, sort=True)
    assert (off_m.cpu().numpy() == off_o1).all() and (ids_m.cpu().numpy() == ids_o1).all()
    assert 128 < int(np.diff(off_o1).max()) <= 512
    # determinism: two builds give identical structures
    pn.update_(nhs, td, td)
    cs_a, cp_a = nhs.export_csr()
    pn.update_(nhs, td, td)
    cs_b, cp_b = nhs.export_csr()
    assert torch.equal(cs_a, cs_b) and torch.equal(cp_a, cp_b)
    # Precomputed rejects inactive points (nhs_precomputed.jl:136-138)
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=T(0.1), n_points=300,
                                              update_neighborhood_search=nhs)
    with pytest.raises(pn.PointNeighborsError, match="does not support inactive points"):
        pn.initialize_(pre, td, td, eachindex_y=np.arange(10))
    # list overflow -> the reference's error text (vector_of_vectors.jl:119)
    pre_small = pn.PrecomputedNeighborhoodSearch[3](search_radius=T(0.1), n_points=300,
                                                    update_neighborhood_search=nhs, max_neighbors=8)
    with pytest.raises(pn.PointNeighborsError, match="cell list is full"):
        pn.initialize_(pre_small, td, td)


def test_count_full_config1(pn, oracle):
    """BASELINE config 1: count_neighbors on the 64^3 perturbed lattice, full size, vs the oracle."""
    c, r, mn, mx = pn.benchmark_cloud((64, 64, 64), seed=1)
    nhs = make_grid(pn, 3, r, mn, mx, n_points=len(c))
    assert nhs.cell_list.n_cells_per_dimension == (24, 24, 24)
    x = dev(c)
    pn.initialize_(nhs, x, x)
    cnt = torch.zeros(len(c), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
    og = oracle.Grid(3, r, mn, mx)
    og.build(c)
    ref = og.count_neighbors(c, c)
    assert (cnt.cpu().numpy() == ref).all()
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()


def test_full_size_properties_16m(pn):
    """BASELINE config 3 size (254^3 = 16 387 064 points): size-independent properties.
    cell list is a permutation; per-cell ids ascending; every point sits in the cell the
    point_cells kernel assigns; neighbour relation is symmetric (sum of counts is even after
    removing self pairs) and every point counts itself; update! is idempotent."""
    n = 254
    N = n ** 3
    T = np.float32
    gen = torch.Generator(device="cuda").manual_seed(7)
    k = torch.arange(N, device="cuda", dtype=torch.int64)
    idx = torch.stack([k % n, (k // n) % n, k // (n * n)], dim=1).to(torch.float32) + 1.0
    coords = (idx + 0.0707 * torch.randn(N, 3, device="cuda", generator=gen)) / T(n + 1)
    coords = coords.to(torch.float32).contiguous()
    r = T(3.0) / T(n + 1)
    mn, mx = np.zeros(3, T), np.ones(3, T)
    nhs = make_grid(pn, 3, r, mn, mx, n_points=N)
    assert nhs.cell_list.n_cells_per_dimension == (88, 88, 88)
    pn.initialize_(nhs, coords, coords)
    cs, cp = nhs.export_csr()
    assert int(cs[-1]) == N
    assert torch.equal(torch.sort(cp.to(torch.int64)).values, k)
    cells = nhs.point_cells(coords).to(torch.int64)
    counts = torch.bincount(cells, minlength=nhs.total_cells())
    assert torch.equal(counts, (cs[1:] - cs[:-1]).to(torch.int64))
    # ids ascending inside each cell: diffs are positive except at cell starts
    d = cp[1:].to(torch.int64) - cp[:-1].to(torch.int64)
    is_start = torch.zeros(N, dtype=torch.bool, device="cuda")
    is_start[cs[:-1][counts > 0].to(torch.int64)] = True
    assert bool(((d > 0) | is_start[1:]).all())
    # the cell of the k-th listed point is the cell whose range contains k
    cell_of_slot = torch.repeat_interleave(torch.arange(nhs.total_cells(), device="cuda"), counts)
    assert torch.equal(cells[cp.to(torch.int64)], cell_of_slot)
    cnt = torch.zeros(N, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), coords, coords, nhs)
    assert int(cnt.min()) >= 1
    assert (int(cnt.sum()) - N) % 2 == 0
    mean = float(cnt.double().mean())
    assert 95.0 < mean < 115.0            # ~108 in the bulk, SURVEY.md section 8
    pn.update_(nhs, coords, coords)
    cs2, cp2 = nhs.export_csr()
    assert torch.equal(cs, cs2) and torch.equal(cp, cp2)


# ------------------------------------------------------------------------------------------
# the fp16 pre-filter of the tile sweep must never change a neighbour set
# ------------------------------------------------------------------------------------------
def _shell_cloud(n_centres, r, seed, nd=3):
    """Adversarial cloud: for many centres scattered over the cells of a grid, partners at
    distance r * (1 + k ulp) for k in -4..4 in random directions, i.e. pairs whose d^2 straddles
    r^2 by a few units in the last place -- the cases a sloppy pre-filter would get wrong."""
    T = np.float32
    rng = np.random.default_rng(seed)
    centres = (rng.uniform(2.0, 10.0, (n_centres, nd)) * r).astype(T)
    pts = [centres]
    for k in range(-4, 5):
        d = rng.normal(size=(n_centres, nd))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rad = np.float64(r) * (1.0 + k * 2.0 ** -23)
        pts.append((centres.astype(np.float64) + rad * d).astype(T))
    # axis-aligned partners: the worst case for a pre-filter built on cell-local coordinates
    for ax in range(nd):
        e = np.zeros(nd)
        e[ax] = 1.0
        for sgn in (-1.0, 1.0):
            pts.append((centres.astype(np.float64) + sgn * np.float64(r) * e).astype(T))
    return np.concatenate(pts).astype(T)


@pytest.mark.parametrize("nd", [3, 2])
def test_prefilter_never_changes_neighbor_sets(pn, oracle, nd):
    """Pairs at distance r (1 +- few ulp): counts, sorted lists and the delivered pos_diff /
    distance equal the oracle's bit for bit with the fp16 pre-filter on (default) and off, for
    both warps-per-cell settings."""
    r = np.float32(0.37)
    c = _shell_cloud(1500, r, 17, nd)
    mn, mx = c.min(0) - r, c.max(0) + r
    og = oracle.Grid(nd, r, mn, mx)
    og.build(c)
    off_o, ids_o = og.neighbor_lists(c, c, sort=True)
    off_t, ids_t = oracle.trivial_lists(c, c, r)
    assert (off_o == off_t).all() and (ids_o == ids_t).all()
    x = dev(c)
    L = pn._lib.lib()
    try:
        for wpc, half in ((0, -1), (2, -1), (4, -1), (2, 0), (4, 0)):
            L.pnb_set_tuning(wpc, half)
            nhs = make_grid(pn, nd, r, mn, mx)
            pn.initialize_(nhs, x, x)
            cnt = torch.zeros(len(c), dtype=torch.int64, device="cuda")
            pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
            assert (cnt.cpu().numpy() == np.diff(off_o)).all(), (wpc, half)
            pre = pn.PrecomputedNeighborhoodSearch[nd](search_radius=r, n_points=len(c),
                                                       update_neighborhood_search=nhs,
                                                       max_neighbors=4096)
            pn.initialize_(pre, x, x)
            off, ids = pre.export_csr()
            assert (off.cpu().numpy() == off_o).all() and (ids.cpu().numpy() == ids_o).all(), (wpc, half)
    finally:
        L.pnb_set_tuning(0, -1)
    # n-body over the same cloud: the fused closure sees exactly the oracle's pairs
    mass, G = _nbody_inputs(len(c))
    dv = torch.zeros((len(c), nd), dtype=torch.float32, device="cuda")
    pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), G), x, x, nhs)
    ref, ref64, refabs = og.nbody(c, c, mass, G, wide=True)
    assert np.all(np.abs(dv.cpu().numpy() - ref64) <= 1e-5 * refabs + 1e-30)


@pytest.mark.parametrize("ratio", [1.0005, 1.17, 1.3329])
def test_periodic_prefilter_cell_size_ratios(pn, oracle, ratio):
    """Periodic boxes whose cell_size / r ranges up to the maximum 4/3 (3 cells per dimension is
    the smallest legal grid): the pre-filter works on minimum-image copies staged next to the tile;
    lists and counts must equal the oracle's and brute force."""
    T = np.float32
    r = T(0.11)
    rng = np.random.default_rng(23)
    size = np.array([3 * ratio, 4 * ratio, 5 * ratio]) * np.float64(r)
    bmn = np.array([0.1, -0.2, 0.05], T)
    bmx = (bmn.astype(np.float64) + size).astype(T)
    n = 6000
    c = (bmn.astype(np.float64) + rng.random((n, 3)) * (bmx.astype(np.float64) - bmn)).astype(T)
    c = np.clip(c, bmn, np.nextafter(bmx, bmn)).astype(T)
    og = oracle.Grid(3, r, bmn, bmx, periodic_box=(bmn, bmx))
    nhs = make_grid(pn, 3, r, bmn, bmx, box=(bmn, bmx))
    assert nhs.n_cells == og.n_cells
    og.build(c)
    x = dev(c)
    pn.initialize_(nhs, x, x)
    off_o, ids_o = og.neighbor_lists(c, c, sort=True)
    off_t, ids_t = oracle.trivial_lists(c, c, r, periodic_box=(bmn, bmx))
    assert (off_o == off_t).all() and (ids_o == ids_t).all()
    cnt = torch.zeros(n, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
    assert (cnt.cpu().numpy() == np.diff(off_o)).all()
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=n,
                                              periodic_box=nhs.periodic_box,
                                              update_neighborhood_search=nhs, max_neighbors=4096)
    pn.initialize_(pre, x, x)
    off, ids = pre.export_csr()
    assert (off.cpu().numpy() == off_o).all() and (ids.cpu().numpy() == ids_o).all()
    # fused closure over the periodic grid
    mass, G = _nbody_inputs(n)
    dv = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
    pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), G), x, x, nhs)
    ref, ref64, refabs = og.nbody(c, c, mass, G, wide=True)
    assert np.all(np.abs(dv.cpu().numpy() - ref64) <= 1e-5 * refabs + 1e-30)


def test_build_variants_identical(pn, oracle):
    """Every variant of the counting-sort kernels (pnb_set_build_tuning) yields the oracle's CSR,
    for a cell-sorted and for a shuffled cloud (thread-local runs vs. lane runs vs. no runs)."""
    c, r, mn, mx = pn.benchmark_cloud((33, 31, 29), seed=12)
    rng = np.random.default_rng(1)
    L = pn._lib.lib()
    try:
        for cloud in (c, c[rng.permutation(len(c))]):
            og = oracle.Grid(3, r, mn, mx)
            og.build(cloud)
            x = dev(cloud)
            for variant in (0, 1, 2, 3, 4, 6, 8, 9, 11, 16, 24, 25):
                L.pnb_set_build_tuning(variant)
                nhs = make_grid(pn, 3, r, mn, mx)
                pn.initialize_(nhs, x, x)
                cs, cp = nhs.export_csr()
                assert (cs.cpu().numpy() == og.cell_start).all(), variant
                assert (cp.cpu().numpy() == og.cell_points).all(), variant
            # one-pass update! into buckets: match.any lane groups (default) and runs of
            # adjacent lanes (bit 2048) fill the same buckets
            for variant in (25, 25 | 2048):
                L.pnb_set_build_tuning(variant)
                nhs = make_grid(pn, 3, r, mn, mx)
                pn.initialize_(nhs, x, x)
                pn.update_(nhs, x, x)
                pn.update_(nhs, x, x)
                cs, cp = nhs.export_csr()
                assert (cs.cpu().numpy() == og.cell_start).all(), variant
                assert (cp.cpu().numpy() == og.cell_points).all(), variant
    finally:
        L.pnb_set_build_tuning(25)


def test_per_point_api(pn, oracle):
    """foreach_neighbor / mapreduce_neighbor (src/neighborhood_search.jl:236-383) for single points
    of a second point set: neighbours, pos_diff and distance equal the oracle's, in the reference's
    visiting order; the reduction equals the host reduction; out-of-range point -> BoundsError."""
    r = np.float32(2.5)
    y = _cloud(pn, (14, 13, 12), r, 5)
    rng = np.random.default_rng(2)
    x = (y[rng.choice(len(y), 50, replace=False)] + rng.normal(0, 0.2, (50, 3))).astype(np.float32)
    mn, mx = y.min(0) - 2 * r, y.max(0) + 2 * r
    nhs = make_grid(pn, 3, r, mn, mx)
    tx, ty = dev(x), dev(y)
    pn.initialize_(nhs, tx, ty)
    og = oracle.Grid(3, r, mn, mx)
    og.build(y)
    off_o, ids_o = og.neighbor_lists(x, y, sort=False)
    pd_o, dist_o = oracle.list_pairs(x, y, off_o, ids_o, r)
    for i in (0, 17, 49):
        got = []
        pn.foreach_neighbor(lambda a, j, pd, d: got.append((a, j, tuple(pd), float(d))), tx, ty, nhs, i)
        lo, hi = off_o[i], off_o[i + 1]
        assert [g[1] for g in got] == ids_o[lo:hi].tolist()
        assert all(g[0] == i for g in got)
        assert np.array_equal(np.array([g[2] for g in got], np.float32).reshape(-1, 3), pd_o[lo:hi])
        assert np.array_equal(np.array([g[3] for g in got], np.float32), dist_o[lo:hi])
        total = pn.mapreduce_neighbor(lambda a, j, pd, d: float(d), lambda u, v: u + v, tx, ty, nhs,
                                      i, init=0.0)
        assert total == sum(float(d) for d in dist_o[lo:hi])
        nmax = pn.mapreduce_neighbor(lambda a, j, pd, d: j, max, tx, ty, nhs, i, init=-1)
        assert nmax == (int(ids_o[lo:hi].max()) if hi > lo else -1)
    with pytest.raises(pn.BoundsError):
        pn.foreach_neighbor(lambda *a: None, tx, ty, nhs, 50)


@pytest.mark.parametrize("periodic", [False, True])
def test_two_sets_tile_path(pn, oracle, periodic):
    """x != y with many query points: the query points are binned into the grid's cells and swept
    by the tile kernel (pnb_set_twoset_tiles).  Counts / unsorted-then-sorted lists bit-exact,
    n-body and WCSPH within 1e-5, identical neighbour sets with the tile path on and off; periodic
    case with points of both sets up to two periods outside the box."""
    T = np.float32
    rng = np.random.default_rng(31)
    if periodic:
        c, r, bmn, bmx = _periodic_case(pn, 26, 3)
        y = c
        size = (bmx - bmn).astype(np.float64)
        x = (bmn + rng.random((9000, 3)) * size).astype(T)
        # the reference wraps cells, not coordinates: move some points to other periods
        y = (y + (rng.integers(-2, 3, y.shape) * (rng.random(y.shape) < 0.2)) * size).astype(T)
        x = (x + (rng.integers(-2, 3, x.shape) * (rng.random(x.shape) < 0.2)) * size).astype(T)
        mn, mx, box = bmn, bmx, (bmn, bmx)
    else:
        r = T(2.5)
        y = _cloud(pn, (30, 28, 26), r, 4)
        x = (y[rng.choice(len(y), 9000, replace=False)] + rng.normal(0, 0.4, (9000, 3))).astype(T)
        mn, mx, box = y.min(0) - 2 * r, y.max(0) + 2 * r, None
    og = oracle.Grid(3, r, mn, mx, periodic_box=box)
    og.build(y)
    cnt_o = og.count_neighbors(x, y)
    off_o, ids_o = og.neighbor_lists(x, y, sort=True)
    off_t, ids_t = oracle.trivial_lists(x, y, r, periodic_box=box)
    assert (off_o == off_t).all() and (ids_o == ids_t).all()
    tx, ty = dev(x), dev(y)
    nhs = make_grid(pn, 3, r, mn, mx, box=box)
    pn.initialize_(nhs, tx, ty)
    mass, G = _nbody_inputs(len(y))
    ref_nb, ref64_nb, refabs_nb = og.nbody(x, y, mass, G, wide=True)
    L = pn._lib.lib()
    try:
        for tiles_on in (2, 1, 0):     # 2 = tile kernel always, 1 = by occupancy, 0 = per point
            L.pnb_set_twoset_tiles(tiles_on)
            cnt = torch.full((len(x),), -1, dtype=torch.int64, device="cuda")
            pn.foreach_point_neighbor(pn.CountNeighbors(cnt), tx, ty, nhs)
            assert (cnt.cpu().numpy() == cnt_o).all(), tiles_on
            lists = pn.api._NeighborLists.build(nhs, tx, ty, sort=True)
            off, ids = lists.export_csr(0)
            assert (off.cpu().numpy() == off_o).all() and (ids.cpu().numpy() == ids_o).all(), tiles_on
            dv = torch.zeros((len(x), 3), dtype=torch.float32, device="cuda")
            pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), G), tx, ty, nhs)
            assert np.all(np.abs(dv.cpu().numpy() - ref64_nb) <= 1e-5 * refabs_nb + 1e-30), tiles_on
    finally:
        L.pnb_set_twoset_tiles(1)
    if not periodic:
        # WCSPH between two systems (fluid-boundary style)
        vy, my, py_, kw = _wcsph_inputs(pn, y, r, 3, seed=3)
        vx, mxx, px_, _ = _wcsph_inputs(pn, x, r, 3, seed=4)
        dvw = torch.zeros((len(x), 4), dtype=torch.float32, device="cuda")
        f = pn.WCSPHInteract(dvw, dev(vx), dev(vy), dev(mxx), dev(my), dev(px_), dev(py_), **kw)
        ref, ref64, refabs = og.wcsph(x, y, vx, vy, mxx, my, px_, py_, f.params_array(), wide=True)
        try:
            for tiles_on in (2, 0):
                L.pnb_set_twoset_tiles(tiles_on)
                pn.foreach_point_neighbor(f, tx, ty, nhs)
                assert np.all(np.abs(dvw.cpu().numpy() - ref64) <= 1e-5 * refabs + 1e-30), tiles_on
            L.pnb_set_twoset_tiles(2)
            # a query point whose stencil leaves the grid -> BoundsError, as on the per-point path
            x_bad = x.copy()
            x_bad[11] = mx + 10 * r
            with pytest.raises(pn.BoundsError):
                pn.foreach_point_neighbor(pn.CountNeighbors(cnt), dev(x_bad), ty, nhs)
        finally:
            L.pnb_set_twoset_tiles(1)
        # a query point whose stencil leaves the grid -> BoundsError, as on the per-point path
        x_bad = x.copy()
        x_bad[11] = mx + 10 * r
        with pytest.raises(pn.BoundsError):
            pn.foreach_point_neighbor(pn.CountNeighbors(cnt), dev(x_bad), ty, nhs)


def test_two_sets_edge_cases(pn, oracle):
    """Two-set tile path on awkward inputs: query cells with more than 32 points (several passes),
    candidate tiles beyond the staging capacity (overflow kernel with a separate query list),
    queries in cells that hold no candidates, empty and tiny query sets."""
    T = np.float32
    mn, mx = np.zeros(3, T), np.ones(3, T)
    r = T(0.1)
    rng = np.random.default_rng(44)
    # candidates: a dense blob (tiles overflow) + a sparse background
    y = np.concatenate([0.5 + 0.05 * rng.random((3000, 3)), rng.random((3000, 3))]).astype(T)
    # queries: a blob overlapping the candidates' blob (> 32 per cell), a blob in an empty corner
    # of the background, and scattered points
    x = np.concatenate([0.48 + 0.08 * rng.random((2500, 3)), 0.05 + 0.02 * rng.random((1500, 3)),
                        rng.random((2000, 3))]).astype(T)
    og = oracle.Grid(3, r, mn, mx)
    og.build(y)
    cnt_o = og.count_neighbors(x, y)
    off_o, ids_o = og.neighbor_lists(x, y, sort=True)
    mass, G = _nbody_inputs(len(y))
    _, r64, rabs = og.nbody(x, y, mass, G, wide=True)
    nhs = make_grid(pn, 3, r, mn, mx)
    tx, ty = dev(x), dev(y)
    pn.initialize_(nhs, tx, ty)
    L = pn._lib.lib()
    try:
        for mode in (2, 0):
            L.pnb_set_twoset_tiles(mode)
            cnt = torch.full((len(x),), -3, dtype=torch.int64, device="cuda")
            pn.foreach_point_neighbor(pn.CountNeighbors(cnt), tx, ty, nhs)
            assert (cnt.cpu().numpy() == cnt_o).all(), mode
            lists = pn.api._NeighborLists.build(nhs, tx, ty, sort=True)
            off, ids = lists.export_csr(0)
            assert (off.cpu().numpy() == off_o).all() and (ids.cpu().numpy() == ids_o).all(), mode
            dv = torch.zeros((len(x), 3), dtype=torch.float32, device="cuda")
            pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), G), tx, ty, nhs)
            assert np.all(np.abs(dv.cpu().numpy() - r64) <= 1e-5 * rabs + 1e-30), mode
        L.pnb_set_twoset_tiles(2)
        # empty and tiny query sets (below the tile path's threshold: per-point kernel)
        e = torch.zeros((0, 3), dtype=torch.float32, device="cuda")
        pn.foreach_point_neighbor(pn.CountNeighbors(torch.zeros(0, dtype=torch.int64, device="cuda")),
                                  e, ty, nhs)
        few = dev(x[:7])
        c7 = torch.zeros(7, dtype=torch.int64, device="cuda")
        pn.foreach_point_neighbor(pn.CountNeighbors(c7), few, ty, nhs)
        assert (c7.cpu().numpy() == cnt_o[:7]).all()
    finally:
        L.pnb_set_twoset_tiles(1)


# ------------------------------------------------------------------------------------------
# Float64 searches: everything in Float64, bit-exact vs the Float64 oracle
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [(14, 13, 12), (40, 30), (200,)])
def test_float64_search(pn, oracle, size):
    """Float64 coordinates, corners and radius: cells, CSR cell list, counts, sorted and unsorted
    neighbour lists and the pair geometry (pos_diff, distance) equal the Float64 oracle's bit for
    bit; x != y and `points` subsets included; Float32 tensors are rejected."""
    T = np.float64
    nd = len(size)
    r = T(2.5)
    y = pn.point_cloud(size, 2.5, seed=3).astype(T)
    rng = np.random.default_rng(8)
    mn, mx = y.min(0) - 2 * r, y.max(0) + 2 * r
    cl = pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=r)
    assert cl.eltype == np.float64
    nhs = pn.GridNeighborhoodSearch[nd](search_radius=r, n_points=len(y), cell_list=cl)
    og = oracle.Grid(nd, r, mn, mx, dtype=np.float64)
    assert cl.n_cells_per_dimension == og.grid_size
    assert np.array_equal(cl.min_corner, og.min_corner) and np.array_equal(cl.max_corner, og.max_corner)
    ty = torch.as_tensor(y, device="cuda")
    assert ty.dtype == torch.float64
    pn.initialize_(nhs, ty, ty)
    og.build(y)
    assert (nhs.point_cells(ty).cpu().numpy() == og.point_cells(y)).all()
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()
    cnt = torch.zeros(len(y), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), ty, ty, nhs)
    off_o, ids_o = og.neighbor_lists(y, y, sort=True)
    assert (cnt.cpu().numpy() == np.diff(off_o)).all()
    off_t, ids_t = oracle.trivial_lists(y, y, r, dtype=np.float64)
    assert (off_o == off_t).all() and (ids_o == ids_t).all()
    for sort in (True, False):
        pre = pn.PrecomputedNeighborhoodSearch[nd](search_radius=r, n_points=len(y),
                                                   update_neighborhood_search=nhs, max_neighbors=4096,
                                                   sort_neighbor_lists=sort)
        pn.initialize_(pre, ty, ty)
        off, ids = pre.export_csr()
        off_s, ids_s = og.neighbor_lists(y, y, sort=sort)
        assert (off.cpu().numpy() == off_s).all() and (ids.cpu().numpy() == ids_s).all(), sort
    pd, dist = pre._lists.pairs(nhs, ty, ty)
    pd_o, dist_o = oracle.list_pairs(y, y, off_s, ids_s, r, dtype=np.float64)
    assert pd.dtype == torch.float64
    assert np.array_equal(pd.cpu().numpy(), pd_o) and np.array_equal(dist.cpu().numpy(), dist_o)
    # second point set + points subset + callable
    x = (y[rng.choice(len(y), 60, replace=False)] + rng.normal(0, 0.3, (60, nd))).astype(T)
    tx = torch.as_tensor(x, device="cuda")
    cx = torch.zeros(60, dtype=torch.int64, device="cuda")
    pts = np.arange(3, 60, 4)
    pn.foreach_point_neighbor(pn.CountNeighbors(cx), tx, ty, nhs, points=pts)
    assert (cx.cpu().numpy() == og.count_neighbors(x, y, points=pts)).all()
    got = []
    pn.foreach_neighbor(lambda i, j, pdv, d: got.append((j, float(d))), tx, ty, nhs, 5)
    off_x, ids_x = og.neighbor_lists(x, y, sort=False)
    pdx, dx = oracle.list_pairs(x, y, off_x, ids_x, r, dtype=np.float64)
    assert [g[0] for g in got] == ids_x[off_x[5]:off_x[6]].tolist()
    assert [g[1] for g in got] == [float(v) for v in dx[off_x[5]:off_x[6]]]
    with pytest.raises(TypeError):
        pn.initialize_(nhs, ty.float(), ty.float())
    # fused closures in Float64: the oracle's IEEE double operation sequence, candidates in the
    # reference's order -> bit-identical sums (n-body in every dimension, WCSPH in 2-D / 3-D)
    mass = (1e10 * (rng.random(len(y)) + 1)).astype(T)
    G = T(6.6743e-11)
    dv = torch.full((len(y), nd), 3.0, dtype=torch.float64, device="cuda")
    pn.foreach_point_neighbor(pn.NBodyGravity(dv, torch.as_tensor(mass, device="cuda"), G), ty, ty, nhs)
    assert np.array_equal(dv.cpu().numpy(), og.nbody(y, y, mass, G))
    dvx = torch.zeros((60, nd), dtype=torch.float64, device="cuda")
    pn.foreach_point_neighbor(pn.NBodyGravity(dvx, torch.as_tensor(mass, device="cuda"), G), tx, ty, nhs,
                              points=pts)
    assert np.array_equal(dvx.cpu().numpy()[pts], og.nbody(x, y, mass, G, points=pts)[pts])
    if nd >= 2:
        rho = (1000.0 + rng.random(len(y))).astype(T)
        v = np.concatenate([rng.normal(0, 0.1, (len(y), nd)), rho[:, None]], axis=1).astype(T)
        m = np.full(len(y), 0.1 * (r / 3), T)
        p = (100.0 * (rho - 1000.0)).astype(T)
        h = T(r / 2)
        sigma = 21.0 / (16.0 * np.pi) / h ** 3 if nd == 3 else 7.0 / (4.0 * np.pi) / h ** 2
        tv, tm, tp = (torch.as_tensor(a, device="cuda") for a in (v, m, p))
        dvw = torch.zeros((len(y), nd + 1), dtype=torch.float64, device="cuda")
        f = pn.WCSPHInteract(dvw, tv, tv, tm, tm, tp, tp, smoothing_length=h, sound_speed=T(10.0),
                             alpha=0.02, beta=0.3, epsilon=0.01, delta=0.1, kernel_norm=T(sigma), ndims_=nd)
        pn.foreach_point_neighbor(f, ty, ty, nhs)
        prm = np.array(f.params64, dtype=T)
        assert np.array_equal(dvw.cpu().numpy(), og.wcsph(y, y, v, v, m, m, p, p, prm))
    with pytest.raises(TypeError):
        pn.foreach_point_neighbor(pn.TLSPHDeformationGradient(None, None, None, None, None,
                                                              smoothing_length=1.0), ty, ty, nhs)


def test_float64_periodic(pn, oracle):
    """Float64 PeriodicBox: n_cells / cell_size from the Float64 constructor arithmetic, lists equal
    to the oracle and to brute force, points outside the box included."""
    T = np.float64
    rng = np.random.default_rng(12)
    r = T(0.11)
    bmn = np.array([0.1, -0.2, 0.05], T)
    bmx = bmn + np.array([0.37, 0.52, 0.61], T)
    n = 4000
    y = bmn + rng.random((n, 3)) * (bmx - bmn)
    y = y + (rng.integers(-1, 2, y.shape) * (rng.random(y.shape) < 0.15)) * (bmx - bmn)
    box = pn.PeriodicBox(min_corner=bmn, max_corner=bmx)
    nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=n, periodic_box=box,
                                       cell_list=pn.FullGridCellList(min_corner=bmn, max_corner=bmx,
                                                                     search_radius=r))
    og = oracle.Grid(3, r, bmn, bmx, periodic_box=(bmn, bmx), dtype=np.float64)
    assert nhs.n_cells == og.n_cells
    assert tuple(nhs.cell_size) == tuple(og.cell_size)
    ty = torch.as_tensor(y, device="cuda")
    pn.initialize_(nhs, ty, ty)
    og.build(y)
    off_o, ids_o = og.neighbor_lists(y, y, sort=True)
    off_t, ids_t = oracle.trivial_lists(y, y, r, periodic_box=(bmn, bmx), dtype=np.float64)
    assert (off_o == off_t).all() and (ids_o == ids_t).all()
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=n, periodic_box=box,
                                              update_neighborhood_search=nhs, max_neighbors=4096)
    pn.initialize_(pre, ty, ty)
    off, ids = pre.export_csr()
    assert (off.cpu().numpy() == off_o).all() and (ids.cpu().numpy() == ids_o).all()
    pd, dist = pre._lists.pairs(nhs, ty, ty)
    pd_o, dist_o = oracle.list_pairs(y, y, off_o, ids_o, r, periodic_box=(bmn, bmx), dtype=np.float64)
    assert np.array_equal(pd.cpu().numpy(), pd_o) and np.array_equal(dist.cpu().numpy(), dist_o)


def test_grid_from_padded_corners(pn, oracle):
    """pnb_grid_create_padded_f32/_f64 (what the Julia glue calls with the corners stored in a
    FullGridCellList): same grid, same cell list as the constructor path from the user corners."""
    import ctypes as C
    L = pn._lib.lib()
    for T, create, build, pf in ((np.float32, L.pnb_grid_create_padded_f32, L.pnb_grid_build_f32, pn._lib._pf),
                                 (np.float64, L.pnb_grid_create_padded_f64, L.pnb_grid_build_f64, pn._lib._pd)):
        r = T(2.5)
        y = pn.point_cloud((12, 11, 10), 2.5, seed=7).astype(T)
        mn, mx = (y.min(0) - r).astype(T), (y.max(0) + r).astype(T)
        cl = pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=r)
        assert cl.eltype == np.dtype(T)
        pmn = np.ascontiguousarray(cl.min_corner, dtype=T)
        pmx = np.ascontiguousarray(cl.max_corner, dtype=T)
        h = C.c_void_p()
        pn._lib.check(create(3, r, pmn.ctypes.data_as(pf), pmx.ctypes.data_as(pf), None, None, C.byref(h)))
        try:
            assert int(L.pnb_grid_total_cells(h)) == int(np.prod(cl.n_cells_per_dimension))
            ty = torch.as_tensor(y, device="cuda")
            pn._lib.check(build(h, ty.data_ptr(), len(y), None, 0, 0, None))
            Cn = int(L.pnb_grid_total_cells(h))
            cs = torch.empty(Cn + 1, dtype=torch.int32, device="cuda")
            cp = torch.empty(len(y), dtype=torch.int32, device="cuda")
            pn._lib.check(L.pnb_grid_export_csr(h, cs.data_ptr(), cp.data_ptr(), 0, None))
            og = oracle.Grid(3, r, mn, mx, dtype=T)
            og.build(y)
            assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()
        finally:
            L.pnb_grid_destroy(h)


def test_update_sequence_layouts(pn, oracle):
    """A sequence of update! calls on ONE search that walks through every build path: first build
    (CSR), one-pass bucket builds, a cloud whose crowded cells overflow the buckets (fallback to
    the CSR build and a new K), growing and shrinking point counts, an `eachindex_y` subset, a
    domain error in the middle, and CSR consumers (exports, two-set sweep) between the builds.
    After every step counts, CSR cell list and WCSPH sums are checked against the oracle."""
    T = np.float32
    r = T(0.1)
    mn, mx = np.zeros(3, T), np.ones(3, T)
    rng = np.random.default_rng(77)
    lat = pn.benchmark_cloud((28, 28, 28), seed=3)[0]           # ~27 points per cell at r = 3/29
    r = pn.benchmark_cloud((28, 28, 28), seed=3)[1]
    mx = pn.benchmark_cloud((28, 28, 28), seed=3)[3]
    clouds = [
        ("lattice", lat),
        ("perturbed", (lat + T(4e-4) * r * rng.standard_normal(lat.shape).astype(T)).astype(T)),
        ("perturbed2", (lat + T(4e-3) * r * rng.standard_normal(lat.shape).astype(T)).astype(T)),
        ("blob", np.concatenate([lat, (0.5 + 0.02 * rng.random((2500, 3))).astype(T)])),   # overflows K
        ("lattice-again", lat),
        ("fewer", lat[: len(lat) // 2]),
        ("more", np.concatenate([lat, np.clip(lat[:5000] + T(0.3) * r, 0, mx - 1e-6).astype(T)])),
    ]
    nhs = make_grid(pn, 3, r, mn, mx)
    og = oracle.Grid(3, r, mn, mx)
    for k, (name, c) in enumerate(clouds):
        c = np.clip(c, 0, mx).astype(T)
        x = dev(c)
        pn.update_(nhs, x, x) if k else pn.initialize_(nhs, x, x)
        og.build(c)
        cnt = torch.zeros(len(c), dtype=torch.int64, device="cuda")
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
        assert (cnt.cpu().numpy() == og.count_neighbors(c, c)).all(), name
        v, m, p, kw = _wcsph_inputs(pn, c, r, 3, seed=k)
        dv = torch.zeros((len(c), 4), dtype=torch.float32, device="cuda")
        f = pn.WCSPHInteract(dv, dev(v), dev(v), dev(m), dev(m), dev(p), dev(p), **kw)
        pn.foreach_point_neighbor(f, x, x, nhs)
        _, r64, rabs = og.wcsph(c, c, v, v, m, m, p, p, f.params_array(), wide=True)
        assert np.all(np.abs(dv.cpu().numpy() - r64) <= 1e-5 * rabs + 1e-30), name
        if k % 2 == 1:
            cs, cp = nhs.export_csr()                      # CSR + canonical order on demand
            assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all(), name
            q = dev(c[::7] + T(0.2) * r)                   # two-set sweep between the builds
            cq = torch.zeros(q.shape[0], dtype=torch.int64, device="cuda")
            pn.foreach_point_neighbor(pn.CountNeighbors(cq), q, x, nhs)
            assert (cq.cpu().numpy() == og.count_neighbors(q.cpu().numpy(), c)).all(), name
            # the fused sweep still works after the layout was switched to CSR
            pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
            assert (cnt.cpu().numpy() == og.count_neighbors(c, c)).all(), name
    # a domain error in the middle leaves the search rebuildable
    bad = lat.copy()
    bad[17, 1] = np.nan
    with pytest.raises(pn.PointNeighborsError, match="NaN or outside the domain"):
        pn.update_(nhs, dev(bad), dev(bad))
    x = dev(lat)
    pn.update_(nhs, x, x)
    og.build(lat)
    idx = np.arange(100, 9000)
    pn.update_(nhs, x, x, eachindex_y=idx)
    og.build(lat, eachindex_y=idx)
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()


@pytest.mark.parametrize("order", [0, 1, -1])
@pytest.mark.parametrize("size", [(26, 26, 26), (60, 50)])
def test_bucket_order_variants(pn, oracle, order, size):
    """The buckets of the one-pass update! may be numbered in linear or in transposed cell order
    (chosen from the order of the input, pnb_set_bucket_order): every consumer of the layout --
    tile sweeps, payload gathers, CSR conversion, exports, two-set sweeps, neighbour lists -- must
    give the same results for both."""
    L = pn._lib.lib()
    nd = len(size)
    c, r, mn, mx = pn.benchmark_cloud(size, seed=6)
    rng = np.random.default_rng(1)
    T = np.float32
    L.pnb_set_bucket_order(order)
    L.pnb_set_build_layout(2)          # every build ends in the bucket layout
    try:
        nhs = make_grid(pn, nd, r, mn, mx)
        og = oracle.Grid(nd, r, mn, mx)
        x = dev(c)
        pn.initialize_(nhs, x, x)
        for step in range(3):
            if step:
                c = np.clip(c + T(4e-3) * r * rng.standard_normal(c.shape).astype(T), 0, mx).astype(T)
                x = dev(c)
                pn.update_(nhs, x, x)
            og.build(c)
            cnt = torch.zeros(len(c), dtype=torch.int64, device="cuda")
            pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
            assert np.array_equal(cnt.cpu().numpy(), og.count_neighbors(c, c))
            v, m, p, kw = _wcsph_inputs(pn, c, r, nd, seed=step)
            dv = torch.zeros((len(c), nd + 1), dtype=torch.float32, device="cuda")
            f = pn.WCSPHInteract(dv, dev(v), dev(v), dev(m), dev(m), dev(p), dev(p), **kw)
            pn.foreach_point_neighbor(f, x, x, nhs)
            _, r64, rabs = og.wcsph(c, c, v, v, m, m, p, p, f.params_array(), wide=True)
            assert np.all(np.abs(dv.cpu().numpy() - r64) <= 1e-5 * rabs + 1e-30)
            if step == 1:
                # sorted neighbour lists straight from the bucket layout (tile kernels)
                lists = pn.api._NeighborLists.build(nhs, x, x, sort=True)
                off, ids = (t.cpu().numpy() for t in lists.export_csr(0))
                roff, rids = og.neighbor_lists(c, c, sort=True)
                assert np.array_equal(off, roff) and np.array_equal(ids, rids)
            if step == 2:
                cs, cp = nhs.export_csr()              # buckets -> CSR -> canonical order
                assert np.array_equal(cs.cpu().numpy(), og.cell_start)
                assert np.array_equal(cp.cpu().numpy(), og.cell_points)
                q = dev(np.clip(c[::5] + T(0.2) * r, 0, mx).astype(T))
                cq = torch.zeros(q.shape[0], dtype=torch.int64, device="cuda")
                pn.foreach_point_neighbor(pn.CountNeighbors(cq), q, x, nhs)
                assert np.array_equal(cq.cpu().numpy(), og.count_neighbors(q.cpu().numpy(), c))
    finally:
        L.pnb_set_bucket_order(0)
        L.pnb_set_build_layout(1)


@pytest.mark.parametrize("layout", [0, 2], ids=["csr", "buckets"])
@pytest.mark.parametrize("periodic", [False, True])
def test_surplus_points_kernel(pn, oracle, layout, periodic):
    """Cells with 33 .. 40 points hand their surplus points to k_sweep_left (default only for large
    clouds; forced here with pnb_set_sweep_left(2)).  The benchmark lattice has such cells (4.8 %);
    a blob adds cells far above 40 points, which keep their additional batches.  n-body and WCSPH
    against the oracle, x === y and two point sets, both cell-list layouts."""
    L = pn._lib.lib()
    T = np.float32
    if periodic:
        c, r, bmn, bmx = _periodic_case(pn, 26, 3, seed=8)
        box = (bmn, bmx)
        mn, mx = bmn, bmx
    else:
        c, r, mn, mx = pn.benchmark_cloud((26, 26, 26), seed=8)
        rng = np.random.default_rng(0)
        c = np.concatenate([c, (0.5 + 0.03 * rng.random((900, 3))).astype(T)])
        box = None
    L.pnb_set_sweep_left(2)
    L.pnb_set_build_layout(layout)
    try:
        nhs = make_grid(pn, 3, r, mn, mx, box=box)
        og = oracle.Grid(3, r, mn, mx, periodic_box=box)
        x = dev(c)
        pn.initialize_(nhs, x, x)
        og.build(c)
        cells = np.diff(og.cell_start)
        assert ((cells > 32) & (cells <= 40)).sum() > 10      # the kernel has work to do
        if not periodic:
            assert (cells > 40).sum() > 0
        mass, G = _nbody_inputs(len(c))
        dv = torch.zeros((len(c), 3), dtype=torch.float32, device="cuda")
        pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), G), x, x, nhs)
        _, r64, rabs = og.nbody(c, c, mass, G, wide=True)
        assert np.all(np.abs(dv.cpu().numpy() - r64) <= 1e-5 * rabs + 1e-30)
        v, m, p, kw = _wcsph_inputs(pn, c, r, 3, seed=3)
        dvw = torch.full((len(c), 4), 2.0, dtype=torch.float32, device="cuda")
        f = pn.WCSPHInteract(dvw, dev(v), dev(v), dev(m), dev(m), dev(p), dev(p), **kw)
        pn.foreach_point_neighbor(f, x, x, nhs)
        _, r64, rabs = og.wcsph(c, c, v, v, m, m, p, p, f.params_array(), wide=True)
        assert np.all(np.abs(dvw.cpu().numpy() - r64) <= 1e-5 * rabs + 1e-30)
        # two point sets through the tile kernel (forced), surplus QUERY points
        q = np.clip(c + T(0.15) * r, mn, mx).astype(T) if not periodic else c[::-1].copy()
        L.pnb_set_twoset_tiles(2)
        vq, mq, pq, _ = _wcsph_inputs(pn, q, r, 3, seed=4)
        dq = torch.zeros((len(q), 4), dtype=torch.float32, device="cuda")
        f2 = pn.WCSPHInteract(dq, dev(vq), dev(v), dev(mq), dev(m), dev(pq), dev(p), **kw)
        pn.foreach_point_neighbor(f2, dev(q), x, nhs)
        _, r64, rabs = og.wcsph(q, c, vq, v, mq, m, pq, p, f2.params_array(), wide=True)
        assert np.all(np.abs(dq.cpu().numpy() - r64) <= 1e-5 * rabs + 1e-30)
    finally:
        L.pnb_set_sweep_left(1)
        L.pnb_set_build_layout(1)
        L.pnb_set_twoset_tiles(1)


@pytest.mark.parametrize("periodic", [False, True])
def test_neighbor_lists_one_pass_rebuilds(pn, oracle, periodic):
    """Repeated builds of sorted x === y lists fill fixed-capacity rows in ONE test pass (capacity =
    longest list of the previous build + margin) and sort + compact them into the CSR list; a list
    that outgrows the capacity falls back to count + fill.  Every rebuild must give the oracle's
    lists; PNB_NLIST_ONE_PASS-independent."""
    T = np.float32
    rng = np.random.default_rng(12)
    if periodic:
        c, r, bmn, bmx = _periodic_case(pn, 22, 3, seed=4)
        box, mn, mx = (bmn, bmx), bmn, bmx
    else:
        c, r, mn, mx = pn.benchmark_cloud((24, 24, 24), seed=4)
        box = None
    nhs = make_grid(pn, 3, r, mn, mx, box=box)
    og = oracle.Grid(3, r, mn, mx, periodic_box=box)
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=len(c),
                                              periodic_box=nhs.periodic_box,
                                              update_neighborhood_search=nhs, max_neighbors=640)
    clouds = [c]
    for k in range(2):
        clouds.append(np.clip(clouds[-1] + T(0.02) * r * rng.standard_normal(c.shape).astype(T),
                              mn + T(1e-6), mx - T(1e-6)).astype(T))
    # a blob: the longest list grows far beyond the capacity derived from the previous build
    blob = clouds[-1].copy()
    blob[:300] = (T(0.5) * (mn + mx) + T(0.3) * r * rng.random((300, 3)).astype(T)).astype(T)
    clouds.append(blob)
    clouds.append(clouds[1])          # and back: one pass again with the larger capacity
    for k, cl in enumerate(clouds):
        x = dev(cl)
        if k == 0:
            pn.initialize_(pre, x, x)
        else:
            pn.update_(pre, x, x, points_moving=(True, True))
        og.build(cl)
        roff, rids = og.neighbor_lists(cl, cl, sort=True)
        off, ids = (t.cpu().numpy() for t in pre.export_csr())
        assert np.array_equal(off, roff) and np.array_equal(ids, rids), k
        backend, lengths = pre.neighbor_lists(index_base=1)
        assert np.array_equal(lengths.cpu().numpy(), np.diff(roff)), k


def test_live_neighbor_coords_semantics(pn, oracle):
    """The reference's sweep reads neighbor_coords LIVE with the cell list of the last
    initialize!/update! (src/nhs_grid.jl:543-548).  Passing another array than the one the search
    was built from (same length, slightly moved points, no update!) must therefore give: old
    cell list + new coordinates -- what the oracle computes when it is built from the old cloud
    and swept with the new one.  A different number of points is a call-order error."""
    c, r, mn, mx = pn.benchmark_cloud((18, 18, 18), seed=12)
    rng = np.random.default_rng(4)
    c2 = (c + np.float32(0.02) * r * rng.standard_normal(c.shape).astype(np.float32)).astype(np.float32)
    nhs = make_grid(pn, 3, r, mn, mx, n_points=len(c))
    x, x2 = dev(c), dev(c2)
    pn.initialize_(nhs, x, x)
    pn.update_(nhs, x, x, points_moving=(True, True))        # bucket layout
    og = oracle.Grid(3, r, mn, mx)
    og.build(c)
    cnt = torch.zeros(len(c), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x2, x2, nhs)
    assert (cnt.cpu().numpy() == og.count_neighbors(c2, c2)).all()
    # same contents in a new buffer: identical to the original array
    x_copy = x.clone()
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x_copy, x_copy, nhs)
    assert (cnt.cpu().numpy() == og.count_neighbors(c, c)).all()
    # two-set: queries from one array, neighbours live from another
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x2, nhs)
    assert (cnt.cpu().numpy() == og.count_neighbors(c, c2)).all()
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=len(c),
                                              update_neighborhood_search=nhs, max_neighbors=200)
    with pytest.raises(pn.PointNeighborsError, match="call update! first"):
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x[:100].contiguous(), nhs)
    del pre


def test_stream_ordered_update(pn, oracle):
    """update_(..., blocking=False): the one-pass build is only enqueued; results of the sweep
    that follows equal the blocking path; a domain error is raised by the next blocking call;
    a bucket overflow (a cell that outgrows the capacity chosen from the previous build) is
    repaired invisibly: the library rebuilds and repeats the sweep."""
    c, r, mn, mx = pn.benchmark_cloud((20, 20, 20), seed=3)
    nhs = make_grid(pn, 3, r, mn, mx, n_points=len(c))
    x = dev(c)
    pn.initialize_(nhs, x, x)
    og = oracle.Grid(3, r, mn, mx)
    rng = np.random.default_rng(5)
    cnt = torch.zeros(len(c), dtype=torch.int64, device="cuda")
    for step in range(4):
        c = (c + np.float32(0.05) * r * rng.standard_normal(c.shape).astype(np.float32)).astype(np.float32)
        c = np.clip(c, mn, mx).astype(np.float32)
        x = dev(c)
        pn.update_(nhs, x, x, points_moving=(True, True), blocking=False)
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
        og.build(c)
        assert (cnt.cpu().numpy() == og.count_neighbors(c, c)).all()
        assert nhs.layout() == "buckets"
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()
    # bucket overflow: 200 points moved into one cell
    c2 = c.copy()
    c2[:200] = c[0] + np.float32(1e-4) * r * rng.standard_normal((200, 3)).astype(np.float32)
    c2 = np.clip(c2, mn, mx).astype(np.float32)
    x2 = dev(c2)
    pn.update_(nhs, x2, x2, points_moving=(True, True), blocking=False)
    v, mass, pressure, kw = _wcsph_inputs(pn, c2, r, 3)
    tv, tm, tp = dev(v), dev(mass), dev(pressure)
    dv = torch.zeros((len(c2), 4), dtype=torch.float32, device="cuda")
    f = pn.WCSPHInteract(dv, tv, tv, tm, tm, tp, tp, **kw)
    pn.foreach_point_neighbor(f, x2, x2, nhs)          # runs on the overflowed list, is repeated
    og.build(c2)
    _, ref64, refabs = og.wcsph(c2, c2, v, v, mass, mass, pressure, pressure, f.params_array(), wide=True)
    assert np.all(np.abs(dv.cpu().numpy() - ref64) <= 1e-5 * refabs + 1e-30)
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x2, x2, nhs)
    assert (cnt.cpu().numpy() == og.count_neighbors(c2, c2)).all()
    # the same through an export instead of a sweep
    pn.update_(nhs, x, x, points_moving=(True, True))
    pn.update_(nhs, x2, x2, points_moving=(True, True), blocking=False)
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()
    # domain error: raised by the next blocking call
    c3 = c.copy()
    c3[7] = mx + np.float32(5.0) * r
    x3 = dev(c3)
    pn.update_(nhs, x, x, points_moving=(True, True))
    pn.update_(nhs, x3, x3, points_moving=(True, True), blocking=False)
    with pytest.raises(pn.PointNeighborsError, match="outside the domain bounds"):
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x3, x3, nhs)
    pn.update_(nhs, x, x, points_moving=(True, True))
    pn.update_(nhs, x3, x3, points_moving=(True, True), blocking=False)
    with pytest.raises(pn.PointNeighborsError, match="outside the domain bounds"):
        pn.check_(nhs)


def test_host_stepper(pn, oracle):
    """pnb200.HostStepper (pnb_hoststep_*): the WCSPH step from HOST buffers, pipelined inside the
    library.  Every submitted step's dv equals the oracle's (1e-5 of sum |term|), with the pressure
    copied from the host and with the pressure computed on the device from the state equation
    (compute_pressure!, oracle: pno_wcsph_compute_pressure); a point outside the grid is reported
    by the call that settles its step."""
    c0, r, mn, mx = pn.benchmark_cloud((18, 20, 16), seed=5)
    n = len(c0)
    nhs = make_grid(pn, 3, r, mn, mx, n_points=n)
    x0 = dev(c0)
    pn.initialize_(nhs, x0, x0)
    og = oracle.Grid(3, r, mn, mx)
    rng = np.random.default_rng(9)
    v, mass, pressure, kw = _wcsph_inputs(pn, c0, r, 3)
    f = pn.WCSPHInteract(None, None, None, None, None, None, None, **kw)
    stepper = pn.HostStepper(nhs, n)
    clouds, outs = [], []
    for s in range(5):
        c = np.clip(c0 + np.float32(0.1) * r * rng.standard_normal(c0.shape).astype(np.float32), mn, mx)
        clouds.append(np.ascontiguousarray(c, np.float32))
        outs.append(np.full((n, 4), np.nan, np.float32))
    with pytest.raises(pn.ArgumentError, match="state equation"):
        stepper.submit(clouds[0], v, None, outs[0], f, mass_host=mass)
    for s in range(3):                       # pressure from the host
        stepper.submit(clouds[s], v, pressure, outs[s], f, mass_host=mass if s == 0 else None)
    stepper.set_state_equation(sound_speed=10.0, reference_density=1000.0)
    for s in range(3, 5):                    # pressure computed on the device
        stepper.submit(clouds[s], v, None, outs[s], f)
    stepper.wait()
    p_eos = oracle.wcsph_compute_pressure(v, 10.0, 1000.0)
    for s in range(5):
        og.build(clouds[s])
        p = pressure if s < 3 else p_eos
        _, ref64, refabs = og.wcsph(clouds[s], clouds[s], v, v, mass, mass, p, p, f.params_array(), wide=True)
        assert np.all(np.abs(outs[s] - ref64) <= 1e-5 * refabs + 1e-30), s
    # a point outside the grid: reported when its step is settled
    bad = clouds[0].copy()
    bad[3] = mx + np.float32(4.0) * r
    stepper.submit(bad, v, pressure, outs[0], f)
    with pytest.raises(pn.PointNeighborsError, match="outside the domain bounds"):
        stepper.wait()


def test_grid_level_calls_and_unsafe_variant(pn, oracle):
    """initialize_grid! / update_grid! (src/nhs_grid.jl:227-281, 470-477) and
    foreach_point_neighbor_unsafe (src/neighborhood_search.jl:204-234) of the host mirror: same
    cell list and counts as initialize! / update! / foreach_point_neighbor and as the oracle."""
    c, r, mn, mx = pn.benchmark_cloud((12, 10, 9), seed=4)
    nhs = make_grid(pn, 3, r, mn, mx, n_points=len(c))
    x = dev(c)
    og = oracle.Grid(3, r, mn, mx)
    assert repr(pn.default_backend(x)) == "B200Backend()"
    pn.initialize_grid_(nhs, x)
    og.build(c)
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()
    cnt = torch.zeros(len(c), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor_unsafe(pn.CountNeighbors(cnt), x, x, nhs)
    assert (cnt.cpu().numpy() == og.count_neighbors(c, c)).all()
    c2 = np.clip(c + np.float32(0.2) * r * np.random.default_rng(1).standard_normal(c.shape).astype(np.float32),
                 mn, mx).astype(np.float32)
    x2 = dev(c2)
    pn.update_grid_(nhs, x2)
    og.build(c2)
    pn.foreach_point_neighbor_unsafe(pn.CountNeighbors(cnt), x2, x2, nhs, points=range(5, 50))
    ref = og.count_neighbors(c2, c2)
    got = cnt.cpu().numpy()
    assert (got[5:50] == ref[5:50]).all() and got[:5].sum() == 0 and got[50:].sum() == 0
    with pytest.raises(pn.BoundsError):
        pn.foreach_point_neighbor_unsafe(pn.CountNeighbors(cnt), x2, x2, nhs, points=[len(c2)])
