"""The vectorised shape-profile interpolation of the ellipsoidal density path must be bit-identical to the
reference's per-object scipy.interpolate.interp1d(..., fill_value='extrapolate') calls
(cosmic_profiles/dens_profs/dens_profs_algos.pyx:169-195), including NaN abscissae (unconverged shells),
unsorted abscissae and extrapolation on both sides."""
import numpy as np

from cosmic_profiles_b200 import dens_profs_algos as D
from oracle import cp_oracle as O


def _case(seed, n=40, R=17, nan_frac=0.15, shuffle=False):
    rng = np.random.default_rng(seed)
    r200 = rng.uniform(0.2, 2.0, size=n)
    rr_shape = np.logspace(-1.5, 0, R)
    d = r200[:, None] * rr_shape[None, :]
    q = rng.uniform(0.4, 1.0, size=(n, R))
    s = q * rng.uniform(0.4, 1.0, size=(n, R))
    vec = rng.normal(size=(3, n, R, 3))
    bad = rng.uniform(size=(n, R)) < nan_frac
    bad[3] = True          # an object without a single converged shell
    bad[4, 1:] = True      # exactly one converged shell
    d[bad] = np.nan
    q[bad] = np.nan
    s[bad] = np.nan
    vec[:, bad] = np.nan
    if shuffle:
        perm = rng.permutation(R)
        d, q, s, vec = d[:, perm], q[:, perm], s[:, perm], vec[:, :, perm]
    return r200, d, d * q, d * s, vec[0], vec[1], vec[2]


def test_bit_identical_to_scipy():
    rb = np.logspace(-2, 0.3, 23)   # reaches below and above every object's shape profile: extrapolation
    for seed, shuffle in ((1, False), (2, True), (3, False)):
        r200, a, b, c, mj, it, mn = _case(seed, shuffle=shuffle)
        with np.errstate(all="ignore"):
            want = O.interpolate_shapes(r200, rb, a, b, c, mj, it, mn)     # scipy, one object at a time
            got = D.interpolate_shapes(r200, rb, a, b, c, mj, it, mn)
        for w, g, nm in zip(want, got, "a b c major inter minor".split()):
            assert w.shape == g.shape, nm
            assert np.array_equal(w, g, equal_nan=True), "%s differs from scipy (seed %d)" % (nm, seed)


def test_clean_profiles():
    r200, a, b, c, mj, it, mn = _case(9, nan_frac=0.0)
    rb = np.logspace(-1.8, 0, 12)
    want = O.interpolate_shapes(r200[5:], rb, a[5:], b[5:], c[5:], mj[5:], it[5:], mn[5:])
    got = D.interpolate_shapes(r200[5:], rb, a[5:], b[5:], c[5:], mj[5:], it[5:], mn[5:])
    for w, g in zip(want, got):
        assert np.array_equal(w, g)
