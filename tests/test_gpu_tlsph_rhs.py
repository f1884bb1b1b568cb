"""GPU parity of the steps either side of the neighbour sweep in a TLSPH / WCSPH right-hand side
(SURVEY.md 8f rank 2): compute_pk1_corrected!, interact_structure_structure! over a
PrecomputedNeighborhoodSearch, compute_pressure!.  The arithmetic belongs to TrixiParticles.jl
(not vendored): parity is against the oracle's definition (UNPINNED).
Tolerances: pointwise kernels bit-exact (same IEEE operation sequence); force sums
|gpu - ref64| <= 1e-5 * sum_j |term_ij| per component and relative L2 <= 1e-5.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from test_gpu_parity import _periodic_case, dev, make_grid, pn  # noqa: E402,F401


def _tlsph_state(c, r, nd, seed=5):
    rng = np.random.default_rng(seed)
    T = np.float32
    n = len(c)
    disp = (0.02 * np.sin(2 * np.pi * c) + 0.01 * np.cos(4 * np.pi * c[:, ::-1])).astype(T)
    xcur = (c + disp * r).astype(T)
    mass = (T(0.1) * (T(1) + T(0.05) * rng.random(n).astype(T))).astype(T)
    rho0 = (T(1000.0) + rng.random(n).astype(T)).astype(T)
    L = (np.eye(nd, dtype=T)[None] + 0.05 * rng.normal(size=(n, nd, nd))).astype(T).reshape(n, nd * nd)
    return xcur, mass, rho0, L


@pytest.mark.parametrize("exact", [False, True], ids=["fast", "exact"])
@pytest.mark.parametrize("nd,n,periodic", [(3, 22, True), (3, 20, False), (2, 40, True)])
def test_tlsph_rhs(pn, oracle, nd, n, periodic, exact):
    T = np.float32
    if periodic:
        c, r, bmn, bmx = _periodic_case(pn, n, nd)
        box = (bmn, bmx)
        nhs = make_grid(pn, nd, r, bmn, bmx, box=box)
    else:
        c, r, mn, mx = pn.benchmark_cloud((n,) * nd, seed=3)
        box = None
        nhs = make_grid(pn, nd, r, mn, mx)
    N = len(c)
    x = dev(c)
    pre = pn.PrecomputedNeighborhoodSearch[nd](search_radius=r, n_points=N,
                                               periodic_box=nhs.periodic_box,
                                               update_neighborhood_search=nhs, max_neighbors=160)
    pn.initialize_(pre, x, x)
    off, ids = (t.cpu().numpy() for t in pre.export_csr())
    xcur, mass, rho0, L = _tlsph_state(c, r, nd)
    h = T(r / T(2))
    young, nu, alpha = T(1.4e6), T(0.4), T(0.1)
    txc, tm, trho, tL = dev(xcur), dev(mass), dev(rho0), dev(L)
    pn.set_exact_arithmetic(exact)
    try:
        # 1. deformation gradient (fused sweep, already covered by test_gpu_parity; needed as input)
        F = torch.zeros((N, nd * nd), dtype=torch.float32, device="cuda")
        fdef = pn.TLSPHDeformationGradient(F, txc, tm, trho, tL, smoothing_length=h, ndims_=nd)
        pn.foreach_point_neighbor(fdef, x, x, pre)
        Fh = F.cpu().numpy()
        # 2. compute_pk1_corrected!: pointwise, same IEEE operation sequence -> bit-exact
        pk1 = torch.zeros_like(F)
        pn.compute_pk1_corrected_(pk1, F, tL, young_modulus=young, poisson_ratio=nu)
        pk1_ref = oracle.tlsph_pk1_corrected(Fh, L, young, nu)
        assert np.array_equal(pk1.cpu().numpy(), pk1_ref)
        # 3. interact_structure_structure!
        dv = torch.full((N, nd), 9.0, dtype=torch.float32, device="cuda")
        fint = pn.TLSPHInteract(dv, txc, tm, trho, pk1, F, smoothing_length=h, young_modulus=young,
                                penalty_alpha=alpha, ndims_=nd)
        assert pn.foreach_point_neighbor(fint, x, x, pre) is None
    finally:
        pn.set_exact_arithmetic(False)
    ref, ref64, refabs = oracle.tlsph_interact(c, xcur, off, ids, mass, rho0, pk1_ref, Fh, h,
                                               fint.params.kernel_norm, young, alpha, r,
                                               periodic_box=box, wide=True)
    got = dv.cpu().numpy()
    assert np.isfinite(got).all() and np.abs(ref64).max() > 0
    assert np.all(np.abs(got - ref64) <= 1e-5 * refabs + 1e-30)
    assert np.linalg.norm(got - ref64) <= 1e-5 * np.linalg.norm(ref64)


def test_compute_pressure(pn, oracle):
    rng = np.random.default_rng(1)
    T = np.float32
    n = 100_003
    for nd in (2, 3):
        v = np.concatenate([rng.normal(0, 1, (n, nd)).astype(T),
                            (T(1000) + rng.random((n, 1)).astype(T) * T(20) - T(10))], axis=1).astype(T)
        tv = dev(v)
        p = torch.zeros(n, dtype=torch.float32, device="cuda")
        # the benchmark's state equation: exponent 1 (smoothed_particle_hydrodynamics.jl:64-69)
        pn.compute_pressure_(p, tv, sound_speed=T(10), reference_density=T(1000))
        assert np.array_equal(p.cpu().numpy(), oracle.wcsph_compute_pressure(v, T(10), T(1000)))
        # Cole with gamma = 7 and a background pressure: pow is not correctly rounded on either
        # side, 1 ulp of the power is amplified by B = rho0 c^2 / 7
        pn.compute_pressure_(p, tv, sound_speed=T(10), reference_density=T(1000), exponent=T(7),
                             background_pressure=T(100))
        ref = oracle.wcsph_compute_pressure(v, T(10), T(1000), exponent=T(7), background_pressure=T(100))
        B = 1000.0 * 100.0 / 7.0
        assert np.all(np.abs(p.cpu().numpy() - ref) <= 2 * B * np.finfo(T).eps * 1.2)


def test_tlsph_interact_full_size_properties(pn):
    """Config 4 size (8 M points, PeriodicBox) is exercised by tools/config_times.py; here a
    1 M periodic cloud checks a size-independent property of the force sweep: with F = I
    (undeformed body, current == initial coordinates) the PK1 stress and the penalty term vanish,
    so dv must be exactly zero."""
    T = np.float32
    c, r, bmn, bmx = _periodic_case(pn, 100, 3)
    nhs = make_grid(pn, 3, r, bmn, bmx, box=(bmn, bmx))
    N = len(c)
    x = dev(c)
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=N,
                                              periodic_box=nhs.periodic_box,
                                              update_neighborhood_search=nhs, max_neighbors=160)
    pn.initialize_(pre, x, x)
    eye = torch.eye(3, dtype=torch.float32, device="cuda").reshape(1, 9).repeat(N, 1).contiguous()
    pk1 = torch.ones_like(eye)
    pn.compute_pk1_corrected_(pk1, eye, eye, young_modulus=T(1.4e6), poisson_ratio=T(0.4))
    assert not pk1.any()
    dv = torch.ones((N, 3), dtype=torch.float32, device="cuda")
    mass = torch.full((N,), 0.1, dtype=torch.float32, device="cuda")
    rho = torch.full((N,), 1000.0, dtype=torch.float32, device="cuda")
    f = pn.TLSPHInteract(dv, x, mass, rho, pk1, eye, smoothing_length=T(r / T(2)),
                         young_modulus=T(1.4e6))
    pn.foreach_point_neighbor(f, x, x, pre)
    # eps = 2 X_ij - 2 x_ij: zero except across the periodic faces, where x_ij is not wrapped
    # (the reference wraps only the initial pos_diff); interior points see exactly zero
    inner = ((x > torch.as_tensor(bmn + r, device="cuda")) &
             (x < torch.as_tensor(bmx - r, device="cuda"))).all(dim=1)
    assert int(inner.sum()) > N // 2
    assert not dv[inner].any()
