"""CPU: the mixed-precision closures of the oracle (oracle/pn_oracle_mixed.h) against its Float32
restatement."""
import numpy as np


def test_mixed_closures_reduce_to_float32(oracle):
    """With coordinates that are exactly representable in Float32 the mixed pair geometry
    (Float64 subtraction, one conversion) equals the Float32 one, so the mixed closures must equal
    the Float32 oracle's bit for bit -- this pins pno_nbody_mix / pno_wcsph_mix to the restatement
    that the reference's golden vectors pin."""
    rng = np.random.default_rng(5)
    T = np.float32
    for nd in (2, 3):
        n = 600
        x32 = rng.random((n, nd)).astype(T)
        r = T(0.2)
        mg = oracle.MixedGrid(nd, r, x32.min(0).astype(np.float64), x32.max(0).astype(np.float64))
        mg.build(x32.astype(np.float64))
        g = oracle.Grid(nd, r, x32.min(0), x32.max(0))
        g.build(x32)
        mass = rng.random(n).astype(T)
        ref = g.nbody(x32, x32, mass, T(2.0))
        ref = ref[0] if isinstance(ref, tuple) else ref
        assert np.array_equal(mg.nbody(x32.astype(np.float64), x32.astype(np.float64), mass, T(2.0)), ref)
        v = np.concatenate([rng.normal(0, 0.1, (n, nd)), 1000 + rng.random((n, 1))], axis=1).astype(T)
        p = (T(100) * (v[:, nd] - T(1000))).astype(T)
        h = T(r / T(2))
        sigma = T(21.0 / (16.0 * np.pi)) / (h * h * h) if nd == 3 else T(7.0 / (4.0 * np.pi)) / (h * h)
        prm = np.array([h, 10.0, 0.02, 0.1, 0.01, 0.1, sigma], T)
        ref = g.wcsph(x32, x32, v, v, mass, mass, p, p, prm)
        ref = ref[0] if isinstance(ref, tuple) else ref
        pts = rng.choice(n, 100, replace=False)
        got = mg.wcsph(x32.astype(np.float64), x32.astype(np.float64), v, v, mass, mass, p, p, prm)
        assert np.array_equal(got, ref)
        got_p = mg.wcsph(x32.astype(np.float64), x32.astype(np.float64), v, v, mass, mass, p, p, prm, points=pts)
        assert np.array_equal(got_p[pts], ref[pts]) and not got_p[np.setdiff1d(np.arange(n), pts)].any()
