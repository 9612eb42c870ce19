"""Parity gates shared by the oracle and CUDA tests (SURVEY.md section 8(d) "Parity gates").

ints (iteration counts, particle counts): bit-exact.  d, q, s, densities: <= RTOL relative, identical
NaN pattern.  Eigenvectors: <= RTOL after sign alignment, skipped where the neighbouring eigenvalue
gap is small (the problem itself is ill-conditioned there: error ~ eps / gap)."""
import numpy as np

RTOL = 1e-9          # BASELINE.json north_star: "within 1e-9 relative in fp64"
GAP_MIN = 1e-4       # relative eigenvalue gap below which axes are not compared


def assert_close(got, ref, name, rtol=RTOL):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (name, got.shape, ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "%s: NaN pattern differs" % name
    ok = ~np.isnan(ref)
    if ok.any():
        err = np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
        assert err.max() <= rtol, "%s: max rel err %.3e > %.1e" % (name, err.max(), rtol)


def assert_axes(got, ref, q, s, which, name, rtol=RTOL):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, name
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "%s: NaN pattern differs" % name
    q2, s2 = np.asarray(q) ** 2, np.asarray(s) ** 2
    gap = {"major": 1 - q2, "inter": np.minimum(1 - q2, q2 - s2), "minor": q2 - s2}[which]
    ok = np.isfinite(ref).all(-1) & (gap > GAP_MIN)
    if ok.any():
        a, b = got[ok], ref[ok]
        sign = np.sign(np.sum(a * b, -1))[:, None]
        err = np.abs(a - sign * b).max()
        assert err <= rtol * 10, "%s: max axis err %.3e" % (name, err)  # axes are O(1); 1e-8 absolute


def assert_shape_tuple(got, ref, name, rtol=RTOL):
    """got/ref: (d, q, s, minor, inter, major, centers, m) as the reference returns them."""
    for k, nm in enumerate(["d", "q", "s"]):
        assert_close(got[k], ref[k], name + "/" + nm, rtol)
    for k, nm in zip((3, 4, 5), ["minor", "inter", "major"]):
        assert_axes(got[k], ref[k], ref[1], ref[2], nm, name + "/" + nm, rtol)
    assert_close(got[6], ref[6], name + "/centers", rtol)
    assert_close(got[7], ref[7], name + "/m", rtol)


def assert_ints(got, ref, name):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, name
    bad = int((got != ref).sum())
    assert bad == 0, "%s: %d of %d integers differ" % (name, bad, ref.size)


def load_golden(kind):
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_%s.npz" % kind))
    meta = json.loads(str(z["meta"]))
    cat = {k[3:]: z[k] for k in z.files if k.startswith("in/")}
    return z, meta, cat


def golden_tuple(z, tag):
    return tuple(z[tag + "/" + nm] for nm in ["d", "q", "s", "minor", "inter", "major", "centers", "m"])
