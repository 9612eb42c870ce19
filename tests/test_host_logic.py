"""Host-side helpers of the drop-in shim that must reproduce the reference's NumPy arithmetic bit for bit."""
import numpy as np
from numpy.random import default_rng

from cosmic_profiles_b200 import centers as C


def test_vectorised_reference_points_equal_the_reference_loop():
    """python_routines.py:512-514 / :540-542: mean of <= 50 seeded random members, per object."""
    rng = np.random.default_rng(1)
    sizes = np.concatenate((rng.integers(1, 49, size=20), rng.integers(50, 5000, size=300), [0, 50, 51]))
    n_tot = int(sizes.sum())
    xyz = rng.uniform(0, 100, size=(2 * n_tot, 3))
    idx = rng.permutation(2 * n_tot)[:n_tot].astype(np.int32)
    got = C.reference_points(xyz, idx, sizes)
    off = np.concatenate(([0], np.cumsum(sizes)))
    for p, n in enumerate(sizes):
        if n == 0:
            assert (got[p] == 0).all()
            continue
        pts = xyz[idx[off[p]:off[p + 1]]]
        choose = default_rng(seed=0).choice(np.arange(len(pts)), (min(50, len(pts)),), replace=False)
        assert np.array_equal(got[p], np.average(pts[choose], axis=0)), (p, n)
