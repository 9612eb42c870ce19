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


def test_unpack_shape_equals_the_numpy_masking():
    """cpg_unpack_shape (host threads) == the reference's post-processing: slices of the (n, R, 12) result with every
    exact zero replaced by NaN (shape_profs_algos.pyx:616-624)."""
    from cosmic_profiles_b200 import _lib
    rng = np.random.default_rng(3)
    out = rng.normal(size=(37, 11, 12))
    out[rng.uniform(size=out.shape) < 0.2] = 0.0
    out[5] = 0.0
    got = _lib.unpack_shape(out)
    ref = out.copy()
    ref[ref == 0.0] = np.nan
    want = (ref[:, :, 0], ref[:, :, 1], ref[:, :, 2], ref[:, :, 3:6], ref[:, :, 6:9], ref[:, :, 9:12])
    for a, b in zip(got, want):
        assert a.flags["C_CONTIGUOUS"] and np.array_equal(a, b, equal_nan=True)


def test_catalogue_ranges_and_host_token():
    from cosmic_profiles_b200 import _lib
    sizes = np.array([5, 0, 3, 7], dtype=np.int32)
    idx = np.concatenate([np.arange(100, 105), np.arange(3, 6), np.arange(40, 47)]).astype(np.int32)
    ok, start = _lib.catalogue_ranges(idx, sizes)
    assert ok and start.tolist() == [100, 0, 3, 40]
    idx2 = idx.copy()
    idx2[6] = 99
    assert not _lib.catalogue_ranges(idx2, sizes)[0]
    with np.errstate(all="ignore"):
        import pytest
        with pytest.raises(_lib.CosmicGpuError):
            _lib.catalogue_ranges(idx, sizes[:3])
    a = np.arange(300000, dtype=np.float64)
    t0, f0 = _lib.host_token(a), _lib.host_token(a, full=True)
    assert t0 == _lib.host_token(a.copy()) and f0 == _lib.host_token(a.copy(), full=True)
    b = a * 1.5
    assert _lib.host_token(b) != t0 and _lib.host_token(b, full=True) != f0
    a[123457] += 1.0
    assert _lib.host_token(a, full=True) != f0
