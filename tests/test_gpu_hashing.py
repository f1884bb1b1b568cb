"""GPU parity of GridNeighborhoodSearch + SpatialHashingCellList against the
oracle's restatement of src/cell_lists/spatial_hashing.jl and the reference's own tests
(test/cell_lists/spatial_hashing.jl, test/neighborhood_search.jl:121-123).  Bar: the hash table
(lists, coords, collisions), counts and neighbour lists bit-exact; n-body sums bit-exact in exact
mode and within 1e-5 relative (per-component bound |gpu - ref64| <= 1e-5 * sum |term|) otherwise.
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
    return pnb200


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def make_hashed(pn, nd, r, list_size, box=None, n_points=0):
    pb = None if box is None else pn.PeriodicBox(min_corner=np.asarray(box[0], np.float32),
                                                 max_corner=np.asarray(box[1], np.float32))
    return pn.GridNeighborhoodSearch[nd](
        search_radius=np.float32(r), n_points=n_points, periodic_box=pb,
        cell_list=pn.SpatialHashingCellList[nd](list_size=list_size))


def lists_of(off, ids):
    off, ids = off.cpu().numpy(), ids.cpu().numpy()
    return off, ids


@pytest.mark.parametrize("case", [0, 1])
def test_reference_collision_cases(pn, kats, case):
    """test/cell_lists/spatial_hashing.jl:2-67 in Float32 (same cells: -0.05f0/0.1f0 etc.)."""
    k = kats["spatial_hashing"]
    c = k["cases"][case]
    coords = np.array(c["coordinates_rows"], dtype=np.float32).T.copy()
    r = np.float32(k["search_radius"])
    nhs = make_hashed(pn, 2, r, c["list_size"], n_points=len(coords))
    x = dev(coords)
    pn.initialize_(nhs, x, x)
    assert pn.spatial_hash(c["cell1"], c["list_size"]) == pn.spatial_hash(c["cell2"], c["list_size"])
    want = c.get("expected_points_in_cell1", c.get("expected_points_in_cell1_sorted"))
    assert sorted(v + 1 for v in nhs.points_in_cell(c["cell1"])) == want
    assert sorted(v + 1 for v in nhs.points_in_cell(c["cell2"])) == want
    if case == 0:
        neighbors = []
        pn.foreach_neighbor(lambda i, j, pd, d: neighbors.append(j + 1), x, x, nhs,
                            c["point_index"] - 1)
        assert [neighbors] == c["expected_neighbors"]
    else:
        neighbors = [[] for _ in range(len(coords))]
        pn.foreach_point_neighbor(lambda i, j, pd, d: neighbors[i].append(j + 1), x, x, nhs,
                                  points=range(len(coords)))
        assert neighbors == c["expected_neighbors"]
        _, coll = nhs.export_hash_table()
        assert bool(coll[pn.spatial_hash(c["cell1"], c["list_size"]) - 1])


def test_coordinates_flattened_layout(pn, kats):
    """test/cell_lists/spatial_hashing.jl:69-111: the exported `coords` words are the reference's
    UInt128 (one point per cell, table large enough that the chosen cells do not share a key)."""
    r = np.float32(1.0)
    for item in kats["spatial_hashing"]["coordinates_flattened"]:
        cell = item["cell"]
        if any(abs(v) > 2 ** 24 for v in cell) or not any(cell):
            continue     # Float32 coordinates cannot address cells near +-2^31; (0,..) is "unused"
        nd = len(cell)
        coords = (np.array([cell], dtype=np.float32) + np.float32(0.5))
        nhs = make_hashed(pn, nd, r, 16, n_points=1)
        x = dev(coords)
        pn.initialize_(nhs, x, x)
        words, coll = nhs.export_hash_table()
        key = pn.spatial_hash(cell, 16) - 1
        got = words[key].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        assert got.tolist() == item["words"][:3]
        assert not bool(coll.any())


@pytest.mark.parametrize("nd", [1, 2, 3])
@pytest.mark.parametrize("list_size", [1, 13, 600])
@pytest.mark.parametrize("periodic", [False, True])
def test_table_counts_lists_bit_exact(pn, oracle, nd, list_size, periodic):
    rng = np.random.default_rng(nd * 1000 + list_size)
    n = 2500
    y = (rng.random((n, nd)) * 2 - 1).astype(np.float32)
    xq = (rng.random((301, nd)) * 2.4 - 1.2).astype(np.float32)
    r = np.float32(0.11 if nd == 3 else 0.05)
    box = (np.full(nd, -1, np.float32), np.full(nd, 1, np.float32)) if periodic else None
    nhs = make_hashed(pn, nd, r, list_size, box=box, n_points=n)
    oh = oracle.HashGrid(nd, r, list_size, periodic_box=box)
    assert nhs.n_cells == (oh.n_cells if periodic else tuple([-1] * nd))
    assert np.array_equal(np.array(nhs.cell_size, np.float32), oh.cell_size)
    ty = dev(y)
    pn.initialize_(nhs, ty, ty)
    y2 = (y + np.float32(0.01) * rng.standard_normal(y.shape).astype(np.float32)).astype(np.float32)
    for step, cloud in enumerate((y, y2)):
        t = dev(cloud)
        if step:
            pn.update_(nhs, t, t, points_moving=(True, True))
        oh.build(cloud)
        # the table: lists (ids ascending), coords, collisions
        cs, cp = nhs.export_csr()
        assert np.array_equal(cs.cpu().numpy(), oh.key_start)
        assert np.array_equal(cp.cpu().numpy(), oh.key_points)
        words, coll = nhs.export_hash_table()
        assert np.array_equal(words.cpu().numpy(), oh.coords)
        assert np.array_equal(coll.cpu().numpy().astype(np.uint8), oh.collisions)
        assert nhs.point_cells(t).cpu().numpy()[:50].tolist() == \
            [oh.spatial_hash(oh.cell_coords(p)) for p in cloud[:50]]
        for q in (cloud, xq):
            tq = t if q is cloud else dev(q)
            cnt = torch.zeros(len(q), dtype=torch.int64, device="cuda")
            pn.foreach_point_neighbor(pn.CountNeighbors(cnt), tq, t, nhs)
            ref_cnt = oh.count_neighbors(q, cloud)
            assert np.array_equal(cnt.cpu().numpy(), ref_cnt)
            off_t, ids_t = oracle.trivial_lists(q, cloud, r, periodic_box=box)
            assert np.array_equal(ref_cnt, np.diff(off_t))          # = brute force
            for sort in (True, False):
                lists = pn.api._NeighborLists.build(nhs, tq, t, sort=sort)
                off, ids = lists_of(*lists.export_csr(0))
                roff, rids = oh.neighbor_lists(q, cloud, sort=sort)
                assert np.array_equal(off, roff) and np.array_equal(ids, rids)
                if sort:
                    assert np.array_equal(ids, ids_t)
        # points subset
        pts = np.arange(0, len(cloud), 7)
        cnt = torch.full((len(cloud),), 5, dtype=torch.int64, device="cuda")
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), t, t, nhs, points=pts)
        ref = np.zeros(len(cloud), np.int64)
        ref[pts] = oh.count_neighbors(cloud, cloud)[pts]
        assert np.array_equal(cnt.cpu().numpy(), ref)


@pytest.mark.parametrize("exact", [False, True])
def test_nbody_on_hashed_list(pn, oracle, exact):
    c, r, mn, mx = pn.benchmark_cloud((20, 20, 20), seed=9)
    rng = np.random.default_rng(3)
    mass = (np.float32(1e10) * (rng.random(len(c)).astype(np.float32) + np.float32(1))).astype(np.float32)
    G = np.float32(6.6743e-11)
    nhs = make_hashed(pn, 3, r, 2 * len(c), n_points=len(c))
    x = dev(c)
    pn.initialize_(nhs, x, x)
    oh = oracle.HashGrid(3, r, 2 * len(c)).build(c)
    ref, ref64, refabs = oh.nbody(c, c, mass, G, wide=True)
    pn.set_exact_arithmetic(exact)
    try:
        dv = torch.full((len(c), 3), 7.0, dtype=torch.float32, device="cuda")
        pn.foreach_point_neighbor(pn.NBodyGravity(dv, dev(mass), G), x, x, nhs)
    finally:
        pn.set_exact_arithmetic(False)
    got = dv.cpu().numpy()
    assert np.all(np.abs(got - ref64) <= 1e-5 * refabs + 1e-30)
    if exact:
        assert np.array_equal(got, ref)
    # the same cloud through a FullGridCellList gives the same neighbour sets
    cl = pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=r)
    full = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=len(c), cell_list=cl)
    pn.initialize_(full, x, x)
    a = torch.zeros(len(c), dtype=torch.int64, device="cuda")
    b = torch.zeros(len(c), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(a), x, x, full)
    pn.foreach_point_neighbor(pn.CountNeighbors(b), x, x, nhs)
    assert torch.equal(a, b)


def test_wcsph_on_hashed_list_matches_full_grid(pn):
    """The fused WCSPH closure over the hash table against the same closure over the full grid
    (whose parity with the oracle is tests/test_gpu_parity.py::test_wcsph_parity): exact mode
    visits the same pairs with the same operations; only the order of the neighbour cells'
    candidates can differ, so compare within the 1e-5 bar."""
    c, r, mn, mx = pn.benchmark_cloud((18, 18, 18), seed=2)
    n = len(c)
    T = np.float32
    rng = np.random.default_rng(8)
    rho = (T(1000) + rng.random(n).astype(T)).astype(T)
    v = np.concatenate([rng.normal(0, 0.1, (n, 3)).astype(T), rho[:, None]], axis=1)
    m = np.full(n, T(0.1) * (r / T(3)), T)
    p = (T(100) * (rho - T(1000))).astype(T)
    x, tv, tm, tp = dev(c), dev(v), dev(m), dev(p)
    outs = []
    for hashed in (False, True):
        if hashed:
            nhs = make_hashed(pn, 3, r, 2 * n, n_points=n)
        else:
            nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=n, cell_list=pn.FullGridCellList(
                min_corner=mn, max_corner=mx, search_radius=r))
        pn.initialize_(nhs, x, x)
        dv = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        f = pn.WCSPHInteract(dv, tv, tv, tm, tm, tp, tp, smoothing_length=r / T(2), sound_speed=T(10))
        pn.foreach_point_neighbor(f, x, x, nhs)
        outs.append(dv.cpu().numpy().astype(np.float64))
    scale = np.abs(outs[0]).max(axis=0)
    assert np.all(np.abs(outs[0] - outs[1]) <= 2e-5 * scale)


def test_hashed_edge_cases(pn, oracle):
    rng = np.random.default_rng(11)
    y = rng.random((400, 3)).astype(np.float32)
    r = np.float32(0.15)
    t = dev(y)
    # eachindex_y subset (src/nhs_grid.jl:271)
    idx = np.arange(50, 300)
    nhs = make_hashed(pn, 3, r, 800, n_points=len(y))
    pn.initialize_(nhs, t, t, eachindex_y=idx)
    oh = oracle.HashGrid(3, r, 800).build(y, eachindex_y=idx)
    cs, cp = nhs.export_csr()
    assert np.array_equal(cs.cpu().numpy(), oh.key_start) and np.array_equal(cp.cpu().numpy(), oh.key_points)
    cnt = torch.zeros(len(y), dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), t, t, nhs)
    assert np.array_equal(cnt.cpu().numpy(), oh.count_neighbors(y, y))
    # a coordinate whose cell does not fit Int32: the reference's InexactError
    bad = y.copy()
    bad[7, 2] = np.float32(1e12)
    with pytest.raises(pn.PointNeighborsError, match="InexactError"):
        pn.initialize_(nhs, dev(bad), dev(bad))
    bad[7, 2] = np.float32("nan")
    with pytest.raises(pn.PointNeighborsError, match="InexactError"):
        pn.initialize_(nhs, dev(bad), dev(bad))
    # the handle is usable again after a failed build
    pn.initialize_(nhs, t, t)
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), t, t, nhs)
    assert np.array_equal(cnt.cpu().numpy(), oracle.HashGrid(3, r, 800).build(y).count_neighbors(y, y))
    # zero radius: legal no-op search; empty y
    z = make_hashed(pn, 3, 0.0, 10)
    pn.initialize_(z, t, t)
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), t, t, z)
    assert not cnt.any()
    empty = torch.zeros((0, 3), dtype=torch.float32, device="cuda")
    pn.initialize_(nhs, t, empty)
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), t, empty, nhs)
    assert not cnt.any()
    # copy_neighborhood_search keeps list_size (spatial_hashing.jl:129-139); unbounded domain:
    # points far away from the origin are fine
    far = (y * np.float32(50) - np.float32(1000)).astype(np.float32)
    cp_nhs = pn.copy_neighborhood_search(nhs, np.float32(3.0), len(far))
    assert cp_nhs.cell_list.list_size == 800
    tf = dev(far)
    pn.initialize_(cp_nhs, tf, tf)
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), tf, tf, cp_nhs)
    off, _ = oracle.trivial_lists(far, far, np.float32(3.0))
    assert np.array_equal(cnt.cpu().numpy(), np.diff(off))
    # PrecomputedNeighborhoodSearch over a hashed search
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=np.float32(3.0), n_points=len(far),
                                              update_neighborhood_search=cp_nhs)
    pn.initialize_(pre, tf, tf)
    o, i = lists_of(*pre.export_csr(0))
    ro, ri = oracle.trivial_lists(far, far, np.float32(3.0))
    assert np.array_equal(o, ro) and np.array_equal(i, ri)


def test_hashed_benchmark_cloud_262k(pn, oracle):
    """config 1 (64^3 lattice) through the hash table, list_size = 2 n_points: counts equal the
    full grid's and the oracle's."""
    c, r, mn, mx = pn.benchmark_cloud((64, 64, 64), seed=1)
    n = len(c)
    nhs = make_hashed(pn, 3, r, 2 * n, n_points=n)
    x = dev(c)
    pn.initialize_(nhs, x, x)
    cnt = torch.zeros(n, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
    og = oracle.Grid(3, r, mn, mx)
    og.build(c)
    assert np.array_equal(cnt.cpu().numpy(), og.count_neighbors(c, c))
