"""Parity at the sizes BASELINE.json names (configs 2, 3, 4 at full size, config 5's single-GPU
cloud for the > 2^31-byte arrays), against the CPU oracle.  Property to match:
test/neighborhood_search.jl:186-337 (neighbour sets equal the reference's after initialize! and
after update!).  The clouds are generated on the device (pnb200.benchmark_cloud_torch: the
distribution of test/point_cloud.jl) and copied to the host once for the oracle; the oracle's
sweeps take `points`, so per-pair arithmetic is compared on samples of >= 100 000 ids spread over
interior, faces and the cells with more than 32 points, cell lists and counts on ALL points.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

T = np.float32


@pytest.fixture(scope="module")
def pn():
    import pnb200
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    return pnb200


def _grid(pn, r, mn, mx, n, box=None):
    cl = pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=T(r))
    pb = None if box is None else pn.PeriodicBox(min_corner=box[0], max_corner=box[1])
    return pn.GridNeighborhoodSearch[3](search_radius=T(r), n_points=n, periodic_box=pb, cell_list=cl)


def _sample(og, n_total, n_sample, seed, gs):
    """ids spread over the cloud: random interior points, every point of the cells on the six
    faces of a coarse subsample, and every point of a cell with more than 32 points (the cells
    whose lanes do not fit one warp in the tile kernel)."""
    rng = np.random.default_rng(seed)
    cs, cp = og.cell_start, og.cell_points
    counts = np.diff(cs)
    big = np.nonzero(counts > 32)[0]
    big = big[rng.permutation(len(big))[:2000]]
    ids = [cp[cs[c]:cs[c + 1]] for c in big]
    # cells of the first / last occupied layer in every dimension
    lin = np.nonzero(counts > 0)[0]
    c0 = lin % gs[0]
    c1 = (lin // gs[0]) % gs[1]
    c2 = lin // (gs[0] * gs[1])
    face = lin[(c0 == c0.min()) | (c0 == c0.max()) | (c1 == c1.min()) | (c1 == c1.max()) |
               (c2 == c2.min()) | (c2 == c2.max())]
    face = face[rng.permutation(len(face))[:1500]]
    ids += [cp[cs[c]:cs[c + 1]] for c in face]
    ids.append(rng.integers(0, n_total, n_sample).astype(np.int32))
    out = np.unique(np.concatenate(ids)).astype(np.int64)
    return out, len(big), len(face)


def _wcsph_state(n, r, seed, moving, beta=0.0):
    rng = np.random.default_rng(seed)
    rho = (T(1000.0) + rng.random(n, dtype=np.float32)).astype(T)
    vel = rng.normal(0, 0.1, (n, 3)).astype(T) if moving else np.zeros((n, 3), T)
    v = np.concatenate([vel, rho[:, None]], axis=1).astype(T)
    mass = np.full(n, T(0.1) * T(r / T(3)), dtype=T)
    c0 = T(10.0)
    pressure = (c0 * c0 * (rho - T(1000.0))).astype(T)
    kw = dict(smoothing_length=T(r / T(2)), sound_speed=c0, alpha=T(0.02), beta=T(beta),
              epsilon=T(0.01), delta=T(0.1), ndims_=3)
    return v, mass, pressure, kw


def test_config2_nbody_all_points(pn, oracle):
    """BASELINE config 2: n-body on the 101^3 = 1 030 301 point cloud, EVERY point against the
    oracle (|gpu - ref64| <= 1e-5 * sum_j |term_ij| per component, relative L2 <= 1e-5)."""
    xc, r, mn, mx = pn.benchmark_cloud_torch((101, 101, 101), seed=4)
    c = xc.cpu().numpy()
    N = len(c)
    nhs = _grid(pn, r, mn, mx, N)
    assert nhs.cell_list.n_cells_per_dimension == (37, 37, 37)
    pn.initialize_(nhs, xc, xc)
    rng = np.random.default_rng(5)
    mass = (T(1e10) * (rng.random(N, dtype=np.float32) + T(1))).astype(T)
    G = T(6.6743e-11)
    dv = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
    pn.foreach_point_neighbor(pn.NBodyGravity(dv, torch.as_tensor(mass, device="cuda"), G), xc, xc, nhs)
    og = oracle.Grid(3, r, mn, mx)
    og.build(c)
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()
    _, ref64, refabs = og.nbody(c, c, mass, G, wide=True)
    got = dv.cpu().numpy()
    assert np.all(np.abs(got - ref64) <= 1e-5 * refabs + 1e-30)
    assert np.linalg.norm(got - ref64) <= 1e-5 * np.linalg.norm(ref64)
    cnt = torch.zeros(N, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), xc, xc, nhs)
    assert (cnt.cpu().numpy() == og.count_neighbors(c, c)).all()


@pytest.fixture(scope="module")
def config3(pn, oracle):
    """254^3 cloud + its update! target, searches after initialize! and update!, oracle grids."""
    xa, r, mn, mx = pn.benchmark_cloud_torch((254, 254, 254), seed=1)
    N = xa.shape[0]
    gen = torch.Generator(device="cuda").manual_seed(2)
    xb = (xa + (T(4e-4) * r) * torch.randn(N, 3, device="cuda", generator=gen)).contiguous()
    return dict(xa=xa, xb=xb, r=r, mn=mn, mx=mx, N=N)


def test_config3_cells_and_counts_all_points(pn, oracle, config3):
    """BASELINE config 3 (254^3 = 16 387 064 points): after initialize! (two-pass CSR build) and
    after update! on the sigma = 4e-4 r perturbed coordinates (one-pass bucket build, the
    benchmarked path) the CSR cell list is the oracle's bit for bit, and the neighbour count of
    EVERY point equals the oracle's."""
    k = config3
    N, r = k["N"], k["r"]
    nhs = _grid(pn, r, k["mn"], k["mx"], N)
    assert nhs.cell_list.n_cells_per_dimension == (88, 88, 88)
    og = oracle.Grid(3, r, k["mn"], k["mx"])
    cnt = torch.zeros(N, dtype=torch.int64, device="cuda")
    for step, x in enumerate((k["xa"], k["xb"], k["xa"])):
        if step == 0:
            pn.initialize_(nhs, x, x)
        else:
            pn.update_(nhs, x, x, points_moving=(True, True))
        # the benchmarked order: sweep straight after the build (bucket layout from step 1 on) ...
        cnt.zero_()
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), x, x, nhs)
        c = x.cpu().numpy()
        og.build(c)
        ref = og.count_neighbors(c, c)
        assert (cnt.cpu().numpy() == ref).all(), f"counts differ after build {step}"
        # ... then the cell list itself
        cs, cp = nhs.export_csr()
        assert (cs.cpu().numpy() == og.cell_start).all(), f"cell_start differs after build {step}"
        assert (cp.cpu().numpy() == og.cell_points).all(), f"cell_points differ after build {step}"
        if step == 2:
            break
    k["pairs_a"] = int(ref.sum())


@pytest.mark.parametrize("moving,beta", [(True, 0.0), (False, 0.0), (True, 0.5)],
                         ids=["v!=0", "v=0 (benchmark)", "beta=0.5"])
def test_config3_wcsph_sampled(pn, oracle, config3, moving, beta):
    """The benchmarked sweep itself (tile kernel over all 16.4 M points straight after the one-pass
    update!) against the oracle on a sample of >= 100 000 ids: interior, the six faces, and cells
    with more than 32 points.  1e-5 * sum |term| per component, relative L2 <= 1e-5.  beta != 0
    exercises the quadratic term of the artificial viscosity (sign: DESIGN.md 3)."""
    k = config3
    N, r = k["N"], k["r"]
    nhs = _grid(pn, r, k["mn"], k["mx"], N)
    pn.initialize_(nhs, k["xa"], k["xa"])
    x = k["xb"]
    pn.update_(nhs, x, x, points_moving=(True, True))
    v, mass, pressure, kw = _wcsph_state(N, r, 31, moving, beta)
    tv, tm, tp = (torch.as_tensor(a, device="cuda") for a in (v, mass, pressure))
    dv = torch.full((N, 4), 7.0, dtype=torch.float32, device="cuda")
    f = pn.WCSPHInteract(dv, tv, tv, tm, tm, tp, tp, **kw)
    pn.foreach_point_neighbor(f, x, x, nhs)
    c = x.cpu().numpy()
    og = oracle.Grid(3, r, k["mn"], k["mx"])
    og.build(c)
    pts, n_big, n_face = _sample(og, N, 110_000, 3, og.grid_size)
    assert len(pts) >= 100_000 and n_big >= 1000 and n_face >= 1000
    _, ref64, refabs = og.wcsph(c, c, v, v, mass, mass, pressure, pressure, f.params_array(),
                                points=pts, wide=True)
    got = dv[torch.as_tensor(pts, device="cuda")].cpu().numpy()
    assert np.all(np.abs(got - ref64[pts]) <= 1e-5 * refabs[pts] + 1e-30)
    assert np.linalg.norm(got - ref64[pts]) <= 1e-5 * np.linalg.norm(ref64[pts])
    if moving:
        assert np.abs(ref64[pts][:, 3]).max() > 0      # continuity term is exercised
    # the per-point kernel (`points=` subsets) on the same sample
    dv2 = torch.zeros_like(dv)
    f2 = pn.WCSPHInteract(dv2, tv, tv, tm, tm, tp, tp, **kw)
    pn.foreach_point_neighbor(f2, x, x, nhs, points=pts[::4])
    got2 = dv2[torch.as_tensor(pts[::4], device="cuda")].cpu().numpy()
    assert np.all(np.abs(got2 - ref64[pts[::4]]) <= 1e-5 * refabs[pts[::4]] + 1e-30)


def test_config3_neighbor_lists_beyond_2g(pn, oracle, config3):
    """Neighbour lists of all 16.4 M points: P = 1.75 G ids = 7 GB, i.e. byte offsets far beyond
    2^31 (and id positions beyond 2^30).  List lengths of EVERY point equal the oracle's counts,
    full sorted lists equal the oracle's on a sample that includes the last points (the largest
    offsets)."""
    k = config3
    N, r = k["N"], k["r"]
    x = k["xa"]
    nhs = _grid(pn, r, k["mn"], k["mx"], N)
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=N,
                                              update_neighborhood_search=nhs, max_neighbors=160)
    pn.initialize_(pre, x, x)
    off, ids = pre.export_csr()
    P = int(off[-1])
    assert P * 4 > 2 ** 32 and ids.numel() >= P
    c = x.cpu().numpy()
    og = oracle.Grid(3, r, k["mn"], k["mx"])
    og.build(c)
    ref_cnt = og.count_neighbors(c, c)
    off_h = off.cpu().numpy()
    assert (np.diff(off_h) == ref_cnt).all()
    assert P == int(ref_cnt.sum())
    rng = np.random.default_rng(8)
    pts = np.unique(np.concatenate([rng.integers(0, N, 20000), np.arange(N - 3000, N),
                                    np.arange(0, 3000)])).astype(np.int64)
    ro, ri = og.neighbor_lists(c[pts], c, sort=True)
    sel = torch.as_tensor(pts, device="cuda")
    starts = off[sel]
    lens = (off[sel + 1] - starts)
    assert (lens.cpu().numpy() == np.diff(ro)).all()
    # gather the sampled lists on the device
    rep = torch.repeat_interleave(torch.arange(len(pts), device="cuda"), lens)
    pos = torch.arange(int(lens.sum()), device="cuda") - torch.as_tensor(ro[:-1], device="cuda")[rep]
    got = ids[starts[rep] + pos].cpu().numpy()
    assert int(starts.max()) * 4 > 2 ** 32
    assert (got == ri).all()
    del pre, off, ids


def test_config4_periodic_lists_and_tlsph(pn, oracle):
    """BASELINE config 4: 200^3 = 8 M points in a PeriodicBox (66 periodic cells per dimension),
    PrecomputedNeighborhoodSearch: list lengths of EVERY point equal the oracle's counts; full
    sorted lists and the TLSPH deformation gradient on a sample of >= 100 000 points."""
    n = 200
    s = T(1.0) / T(n + 1)
    r = T(3.0) / T(n + 1)
    xc, _, _, _ = pn.benchmark_cloud_torch((n, n, n), seed=9)
    bmn = np.full(3, s / T(2), T)
    bmx = np.full(3, (T(n) + T(0.5)) * s, T)
    N = xc.shape[0]
    nhs = _grid(pn, r, bmn, bmx, N, box=(bmn, bmx))
    assert nhs.n_cells == (66, 66, 66)
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=N,
                                              periodic_box=nhs.periodic_box,
                                              update_neighborhood_search=nhs, max_neighbors=160,
                                              transpose_backend=True)
    pn.initialize_(pre, xc, xc)
    off, ids = pre.export_csr()
    c = xc.cpu().numpy()
    og = oracle.Grid(3, r, bmn, bmx, periodic_box=(bmn, bmx))
    og.build(c)
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()
    ref_cnt = og.count_neighbors(c, c)
    off_h = off.cpu().numpy()
    assert (np.diff(off_h) == ref_cnt).all()
    pts, _, n_face = _sample(og, N, 100_000, 4, og.grid_size)
    assert len(pts) >= 100_000 and n_face >= 1000
    ro, ri = og.neighbor_lists(c[pts], c, sort=True)
    sel = torch.as_tensor(pts, device="cuda")
    starts, lens = off[sel], off[sel + 1] - off[sel]
    rep = torch.repeat_interleave(torch.arange(len(pts), device="cuda"), lens)
    pos = torch.arange(int(lens.sum()), device="cuda") - torch.as_tensor(ro[:-1], device="cuda")[rep]
    assert (ids[starts[rep] + pos].cpu().numpy() == ri).all()
    # a second build goes through the one-pass (fixed-capacity rows) path: same lists
    pn.update_(pre, xc, xc)
    off2, ids2 = pre.export_csr()
    assert torch.equal(off, off2) and torch.equal(ids[:int(off[-1])], ids2[:int(off[-1])])
    # TLSPH deformation gradient over the lists (sampled rows vs the oracle)
    rng = np.random.default_rng(10)
    xcur = (c + (0.01 * np.sin(2 * np.pi * c)).astype(T) * r).astype(T)
    mass = np.full(N, 0.1, T)
    rho0 = np.full(N, 1000.0, T)
    Lm = (np.eye(3, dtype=T)[None] + 0.05 * rng.normal(size=(N, 3, 3))).astype(T).reshape(N, 9)
    h = T(r / T(2))
    F = torch.zeros((N, 9), dtype=torch.float32, device="cuda")
    f = pn.TLSPHDeformationGradient(F, *(torch.as_tensor(a, device="cuda") for a in (xcur, mass, rho0, Lm)),
                                    smoothing_length=h, ndims_=3)
    pn.foreach_point_neighbor(f, xc, xc, pre)
    # oracle on the sample: full-length offsets with empty lists outside the sample
    lens_full = np.zeros(N, np.int64)
    lens_full[pts] = np.diff(ro)
    off_sparse = np.concatenate([[0], np.cumsum(lens_full)])
    _, ref64, refabs = oracle.tlsph_deformation_grad(c, xcur, off_sparse, ri, mass, rho0, Lm, h,
                                                     f.kernel_norm, r, periodic_box=(bmn, bmx),
                                                     wide=True)
    got = F[sel].cpu().numpy()
    assert np.all(np.abs(got - ref64[pts]) <= 1e-5 * refabs[pts] + 1e-30)
    assert np.linalg.norm(got - ref64[pts]) <= 1e-5 * np.linalg.norm(ref64[pts])


def test_config5_cloud_one_gpu_buckets_beyond_2g(pn, oracle):
    """BASELINE config 5's cloud (504^3 = 128 024 064 points, 171^3 cells) on ONE GPU: the bucket
    layout of the one-pass update! is 5 M cells x 64 slots x 16 B = 5.1 GB (slot byte offsets
    beyond 2^32).  Cell list after update! bit-exact vs the oracle for all points; counts and
    WCSPH sums on a sample."""
    n = 504
    xa, r, mn, mx = pn.benchmark_cloud_torch((n, n, n), seed=11)
    N = xa.shape[0]
    assert N == 128_024_064
    nhs = _grid(pn, r, mn, mx, N)
    assert nhs.cell_list.n_cells_per_dimension == (171, 171, 171)
    pn.initialize_(nhs, xa, xa)
    gen = torch.Generator(device="cuda").manual_seed(12)
    xa += (T(4e-4) * r) * torch.randn(N, 3, device="cuda", generator=gen)   # in place: update! target
    pn.update_(nhs, xa, xa, points_moving=(True, True))
    assert nhs.layout() == "buckets"
    c = xa.cpu().numpy()
    og = oracle.Grid(3, r, mn, mx)
    og.build(c)
    v, mass, pressure, kw = _wcsph_state(N, r, 13, True)
    tv, tm, tp = (torch.as_tensor(a, device="cuda") for a in (v, mass, pressure))
    dv = torch.zeros((N, 4), dtype=torch.float32, device="cuda")
    f = pn.WCSPHInteract(dv, tv, tv, tm, tm, tp, tp, **kw)
    pn.foreach_point_neighbor(f, xa, xa, nhs)          # reads the 5.1 GB bucket array
    cnt = torch.zeros(N, dtype=torch.int64, device="cuda")
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), xa, xa, nhs)
    rng = np.random.default_rng(14)
    cs_o, cp_o = og.cell_start, og.cell_points
    last_cells = np.nonzero(np.diff(cs_o) > 0)[0][-400:]          # highest bucket addresses
    pts = np.unique(np.concatenate([rng.integers(0, N, 60_000)] +
                                   [cp_o[cs_o[q]:cs_o[q + 1]] for q in last_cells])).astype(np.int64)
    sel = torch.as_tensor(pts, device="cuda")
    assert (cnt[sel].cpu().numpy() == og.count_neighbors(c, c, points=pts)[pts]).all()
    _, ref64, refabs = og.wcsph(c, c, v, v, mass, mass, pressure, pressure, f.params_array(),
                                points=pts, wide=True)
    got = dv[sel].cpu().numpy()
    assert np.all(np.abs(got - ref64[pts]) <= 1e-5 * refabs[pts] + 1e-30)
    del dv, tv, tm, tp, cnt
    cs, cp = nhs.export_csr()
    assert (cs.cpu().numpy() == cs_o).all() and (cp.cpu().numpy() == cp_o).all()
