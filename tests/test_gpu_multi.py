"""Sharded execution on the device: ONE catalogue dealt over several cpg_multi slots must reproduce the single-device
results -- integers exactly, floats to 1e-12 (the per-object arithmetic is the same; only the scheduling choices that
depend on the number of objects per device, warp / CTA kernel and shells per pass, may differ in summation order).
On a single-GPU box the slots share device 0 (every slot is a context of its own); with >= 2 GPUs the real thing runs too.
Also: segment uploads, contiguous-range catalogues and the snapshot cache."""
import numpy as np
import pytest

from tests import parity as P

pytestmark = pytest.mark.gpu
K = (1e-2, 100, 10)


@pytest.fixture()
def sess():
    from cosmic_profiles_b200 import session
    session.options.clear()
    keep = (session.devices, session.allow_duplicate_devices)
    yield session
    session.devices, session.allow_duplicate_devices = keep
    session.options.clear()
    session.close_all()


def _device_sets():
    from cosmic_profiles_b200 import _lib
    sets = [[0, 0, 0]]
    if _lib.device_count() >= 2:
        sets.append([0, 1])
    return sets


def _cat(shuffle, vel=True, seed=4):
    from cosmic_profiles_b200.mock import make_catalogue
    return make_catalogue([900, 300, 1500, 700, 1100, 450, 2000, 5200, 640, 130000, 2500], seed=seed, with_velocities=vel,
                          shuffle_particles=shuffle)


def _same(got, ref, name):
    for a, b, nm in zip(got, ref, ["d", "q", "s", "minor", "inter", "major", "centers", "m"]):
        assert np.array_equal(np.isnan(a), np.isnan(b)), name + "/" + nm
        ok = ~np.isnan(b)
        if nm in ("minor", "inter", "major"):
            sign = np.sign(np.nansum(a * b, -1))[..., None]
            a = a * sign
        assert np.allclose(a[ok], b[ok], rtol=1e-12, atol=1e-13), name + "/" + nm


@pytest.mark.parametrize("shuffle", [False, True])
def test_sharded_shapes_equal_single_device(sess, shuffle):
    from cosmic_profiles_b200 import shape_profs_algos as S
    cat = _cat(shuffle)
    rr = np.logspace(-1.3, 0, 9)
    calls = {
        "local": lambda c: S.calcMorphLocal(cat["xyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"], 0.0, rr,
                                            *K, c, False, False),
        "local_shell_reduced": lambda c: S.calcMorphLocal(cat["xyz"], cat["masses"], cat["r200"], cat["idx_cat"],
                                                          cat["obj_size"], 0.0, rr, *K, c, True, True),
        "global": lambda c: S.calcMorphGlobal(cat["xyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"], 0.0,
                                              *K, c, 6.0, False),
        "vlocal": lambda c: S.calcMorphLocalVelDisp(cat["xyz"], cat["vxyz"], cat["masses"], cat["r200"], cat["idx_cat"],
                                                    cat["obj_size"], 0.0, rr, *K, c, False, False),
    }
    sess.devices = [0]
    ref = {}
    for name, fn in calls.items():
        for c in ("com", "mode"):
            ref[name, c] = (fn(c), dict(S.last_info))
    for devs in _device_sets():
        sess.devices = devs
        sess.allow_duplicate_devices = True
        for (name, c), (r, info) in ref.items():
            got = calls[name](c)
            P.assert_ints(S.last_info["n_iter"], info["n_iter"], "%s/%s n_iter %r" % (name, c, devs))
            P.assert_ints(S.last_info["n_pts"], info["n_pts"], "%s/%s n_pts %r" % (name, c, devs))
            _same(got, r, "%s/%s %r" % (name, c, devs))
            assert np.allclose(S.last_info["vcenters"], info["vcenters"], rtol=1e-12, atol=0)
            multi = S.last_info["multi"]
            assert multi is not None and len(multi["devices"]) == len(devs)
            assert sum(d["n_obj"] for d in multi["devices"]) == len(cat["obj_size"])
            parts = np.array([d["n_part"] for d in multi["devices"]], dtype=float)
            if not shuffle:   # a contiguous catalogue ships only each shard's particles
                assert parts.sum() == cat["obj_size"].sum()
            else:             # a general catalogue ships the whole snapshot to every device
                assert (parts == len(cat["masses"])).all()


def test_sharded_densities_centres_and_fits_equal_single_device(sess):
    from cosmic_profiles_b200 import dens_profs_algos as D, dens_profs_tools as T, shape_profs_algos as S
    cat = _cat(False, vel=False, seed=6)
    a4 = (cat["xyz"], cat["masses"])
    ic = (cat["idx_cat"], cat["obj_size"])
    rb = np.logspace(-2, 0, 24)

    def everything():
        out = {}
        out["mc_com"] = D.calcMassesCenters(*a4, *ic, 0.0, "com")
        out["mc_mode"] = D.calcMassesCenters(*a4, *ic, 0.0, "mode")
        out["sph"] = D.calcDensProfsSphDirectBinning(*a4, cat["r200"], rb, *ic, 0.0, "com")
        out["sph_counts"] = D.last_info["counts"].copy()
        out["kde"] = D.calcDensProfsKernelBased(*a4, cat["r200"], rb, *ic, 0.0, "com")
        d, q, s, mn, it, mj, _, _ = S.calcMorphLocal(*a4, cat["r200"], *ic, 0.0, np.logspace(-1.5, 0, 20), *K, "com", False, False)
        out["ell"] = D.calcDensProfsEllDirectBinning(*a4, cat["r200"], rb, d, d * q, d * s, mj, it, mn, *ic, 0.0, "com")
        out["ell_counts"] = D.last_info["counts"].copy()
        out["fit"] = T.fitDensProfHelper(out["sph"][:, 4:], rb[4:], cat["r200"], {"profile": "nfw"})
        out["fit_nfev"] = T.last_info["nfev"].copy()
        return out
    sess.devices = [0]
    ref = everything()
    for devs in _device_sets():
        sess.devices = devs
        sess.allow_duplicate_devices = True
        got = everything()
        for key in ("sph_counts", "ell_counts"):
            P.assert_ints(got[key], ref[key], "%s %r" % (key, devs))
        for key in ("sph", "kde", "ell"):
            P.assert_close(got[key], ref[key], "%s %r" % (key, devs), rtol=1e-12)
        # the fits start from densities that may differ in the last bit: same valley, optimiser tolerance 1e-4
        P.assert_close(got["fit"], ref["fit"], "fit %r" % (devs,), rtol=1e-5)
        for key in ("mc_com", "mc_mode"):
            P.assert_close(got[key][0], ref[key][0], key + " centres", rtol=1e-12)
            P.assert_close(got[key][1], ref[key][1], key + " masses", rtol=1e-13)


def test_segments_and_ranges_equal_the_flat_catalogue(sess):
    """cpg_set_particles_segments + cpg_set_catalogue_ranges (what a shard receives) against cpg_set_particles +
    cpg_set_catalogue of the same objects: bit-identical results."""
    from cosmic_profiles_b200 import _lib
    cat = _cat(False, vel=True, seed=8)
    sizes = cat["obj_size"].astype(np.int64)
    off = np.concatenate(([0], np.cumsum(sizes)))
    pick = np.array([1, 2, 5, 9, 10])
    rr = np.logspace(-1.3, 0, 8)
    cen = np.stack([np.average(cat["xyz"][off[h]:off[h + 1]], axis=0, weights=cat["masses"][off[h]:off[h + 1]]) for h in pick])
    ctx = _lib.Context(0)
    # reference: flat catalogue of the picked objects over the whole snapshot
    sub_idx = np.concatenate([cat["idx_cat"][off[h]:off[h + 1]] for h in pick])
    ctx.set_option("detect_ranges", 0)
    ctx.set_particles(cat["xyz"], cat["vxyz"], cat["masses"])
    ctx.set_catalogue(sub_idx, cat["obj_size"][pick])
    assert ctx.stat("catalogue_as_ranges") == 0
    ref = ctx.shape(cen, cat["r200"][pick], rr, True, 0.0, 0.0, *K, False, False, True)
    # the same catalogue, recognised as ranges
    ctx.set_option("detect_ranges", 1)
    ctx.set_catalogue(sub_idx, cat["obj_size"][pick])
    assert ctx.stat("catalogue_as_ranges") == 1
    got = ctx.shape(cen, cat["r200"][pick], rr, True, 0.0, 0.0, *K, False, False, True)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    # only the picked objects' particles on the device
    ctx.set_particles_segments(cat["xyz"], cat["vxyz"], cat["masses"], off[pick], sizes[pick])
    assert ctx.stat("n_part") == sizes[pick].sum()
    with pytest.raises(_lib.CosmicGpuError, match="cpg_set_catalogue again"):
        ctx.shape(cen, cat["r200"][pick], rr, True, 0.0, 0.0, *K, False, False, True)   # stale catalogue refused
    local = np.concatenate(([0], np.cumsum(sizes[pick])))[:-1]
    ctx.set_catalogue_ranges(local, cat["obj_size"][pick])
    got = ctx.shape(cen, cat["r200"][pick], rr, True, 0.0, 0.0, *K, False, False, True)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    xyz_dev, v_dev, m_dev = ctx.get_particles(int(sizes[pick].sum()), with_vel=True)
    sel = np.concatenate([np.arange(off[h], off[h + 1]) for h in pick])
    assert np.array_equal(xyz_dev, cat["xyz"][sel]) and np.array_equal(m_dev, cat["masses"][sel]) and np.array_equal(v_dev, cat["vxyz"][sel])
    ctx.close()


def test_snapshot_cache(sess):
    """Same arrays again -> no upload; changed content -> upload; snapshot_cache = 2 hashes every byte."""
    from cosmic_profiles_b200 import _lib
    cat = _cat(True, vel=False, seed=9)
    ctx = _lib.Context(0)
    xyz, m = cat["xyz"].copy(), cat["masses"].copy()
    ctx.set_particles(xyz, None, m)
    ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
    v0, s0 = ctx.stat("snapshot_version"), ctx.stat("uploads_skipped")
    ctx.set_particles(xyz, None, m)
    ctx.set_catalogue(cat["idx_cat"].copy(), cat["obj_size"])      # a fresh array with the same content (getSubSetIdxCat)
    assert ctx.stat("snapshot_version") == v0 and ctx.stat("uploads_skipped") == s0 + 2
    xyz *= 1.5                                                     # rescaled in place: the sampled token notices
    ctx.set_particles(xyz, None, m)
    assert ctx.stat("snapshot_version") == v0 + 1
    idx2 = cat["idx_cat"].copy()
    idx2[[3, 4]] = idx2[[4, 3]]                                    # any change of the flat catalogue is noticed (full hash)
    c0 = ctx.stat("catalogue_version")
    ctx.set_catalogue(idx2, cat["obj_size"])
    assert ctx.stat("catalogue_version") == c0 + 1
    ctx.set_option("snapshot_cache", 2)
    ctx.set_particles(xyz, None, m)
    v1 = ctx.stat("snapshot_version")
    xyz[len(xyz) // 3 + 7, 1] += 1.0                               # one coordinate: only the full hash notices
    ctx.set_particles(xyz, None, m)
    assert ctx.stat("snapshot_version") == v1 + 1
    ctx.set_option("snapshot_cache", 0)
    ctx.set_particles(xyz, None, m)
    assert ctx.stat("snapshot_version") == v1 + 2
    ctx.close()
