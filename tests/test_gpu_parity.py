"""CUDA path (through the ctypes shim -> C ABI of libcosmicgpu.so) against
 (a) golden vectors produced by the compiled reference (tests/golden/*.npz), and
 (b) the CPU oracle on larger seeded catalogues, for every scheduling variant
     (shells per pass 1/2/4, warp kernel, CTA kernel, thread-block clusters of 2..16).
Gates (tests/parity.py): iteration and particle counts bit-exact, q/s/d/densities <= 1e-9 relative,
axes <= 1e-8 absolute after sign alignment, identical NaN pattern."""
import numpy as np
import pytest

from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from cosmic_profiles_b200 import dens_profs_algos, session, shape_profs_algos
    session.options.clear()
    yield shape_profs_algos, dens_profs_algos, session
    session.options.clear()
    session.close_all()


def _args(cat):
    return cat["xyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"]


@pytest.mark.parametrize("kind", ["main", "pbc"])
def test_golden_shapes(gpu, kind):
    S, D, sess = gpu
    z, meta, cat = P.load_golden(kind)
    X, M, R2, I, Z = _args(cat)
    k = meta["katz"]
    L = meta["L_BOX"]
    for center in (("com", "mode") if kind == "main" else ("com",)):
        for reduced in (0, 1):
            for shell in (0, 1):
                tag = "local_%s_r%d_s%d" % (center, reduced, shell)
                got = S.calcMorphLocal(X, M, R2, I, Z, L, cat["rr_shape"], k["IT_TOL"], k["IT_WALL"], k["IT_MIN"],
                                       center, bool(reduced), bool(shell))
                P.assert_ints(S.last_info["n_iter"], z[tag + "/n_iter"], tag + "/n_iter")
                P.assert_ints(S.last_info["n_pts"], z[tag + "/n_pts"], tag + "/n_pts")
                P.assert_shape_tuple(got, P.golden_tuple(z, tag), tag)
            tag = "global_%s_r%d" % (center, reduced)
            got = S.calcMorphGlobal(X, M, R2, I, Z, L, k["IT_TOL"], k["IT_WALL"], k["IT_MIN"], center, meta["SAFE"],
                                    bool(reduced))
            P.assert_ints(S.last_info["n_iter"], z[tag + "/n_iter"][:, 0], tag + "/n_iter")
            P.assert_ints(S.last_info["n_pts"], z[tag + "/n_pts"][:, 0], tag + "/n_pts")
            P.assert_shape_tuple(got, P.golden_tuple(z, tag), tag)


def test_golden_velocity_dispersion(gpu):
    S, D, sess = gpu
    z, meta, cat = P.load_golden("main")
    X, M, R2, I, Z = _args(cat)
    V = cat["vxyz"]
    k = meta["katz"]
    for reduced in (0, 1):
        tag = "vglobal_com_r%d" % reduced
        got = S.calcMorphGlobalVelDisp(X, V, M, R2, I, Z, 0.0, k["IT_TOL"], k["IT_WALL"], k["IT_MIN"], "com",
                                       meta["SAFE"], bool(reduced))
        P.assert_ints(S.last_info["n_iter"], z[tag + "/n_iter"][:, 0], tag + "/n_iter")
        P.assert_shape_tuple(got, P.golden_tuple(z, tag), tag)
        for shell in (0, 1):
            tag = "vlocal_com_r%d_s%d" % (reduced, shell)   # reference patched for defect (ii)
            got = S.calcMorphLocalVelDisp(X, V, M, R2, I, Z, 0.0, cat["rr_shape"], k["IT_TOL"], k["IT_WALL"],
                                          k["IT_MIN"], "com", bool(reduced), bool(shell))
            P.assert_ints(S.last_info["n_iter"], z[tag + "/n_iter"], tag + "/n_iter")
            P.assert_ints(S.last_info["n_pts"], z[tag + "/n_pts"], tag + "/n_pts")
            P.assert_shape_tuple(got, P.golden_tuple(z, tag), tag)


@pytest.mark.parametrize("kind", ["main", "pbc"])
def test_golden_densities(gpu, kind):
    S, D, sess = gpu
    z, meta, cat = P.load_golden(kind)
    X, M, R2, I, Z = _args(cat)
    L = meta["L_BOX"]
    rb = cat["rr_dens"]
    for center in ("com", "mode"):
        P.assert_close(D.calcDensProfsSphDirectBinning(X, M, R2, rb, I, Z, L, center), z["dens_sph_" + center],
                       "sph_" + center)
        P.assert_close(D.calcDensProfsKernelBased(X, M, R2, rb, I, Z, L, center), z["dens_kernel_" + center],
                       "kernel_" + center)
        cen, m = D.calcMassesCenters(X, M, I, Z, L, center)
        P.assert_close(cen, z["masses_centers_%s/centers" % center], "centers")
        P.assert_close(m, z["masses_centers_%s/m" % center], "m")
    tag = "local_com_r0_s0"
    d, q, s = z[tag + "/d"], z[tag + "/q"], z[tag + "/s"]
    got = D.calcDensProfsEllDirectBinning(X, M, R2, rb, d, d * q, d * s, z[tag + "/major"], z[tag + "/inter"],
                                          z[tag + "/minor"], I.copy(), Z, L, "com")
    P.assert_close(got, z["dens_ell_com"], "ell_com (reference patched for defect i)")


@pytest.fixture(scope="module")
def big():
    from cosmic_profiles_b200.mock import make_catalogue
    from oracle import cp_oracle as O
    sizes = [30000, 12000, 9001, 7000, 5000, 4000, 3000, 2500, 2000, 1500, 1200, 1000, 700, 300, 120, 40, 9, 1]
    cat = make_catalogue(sizes, seed=5, with_velocities=True, shuffle_particles=True)
    rr = np.logspace(-1.5, 0, 13)
    cen = O.centres(cat["xyz"], cat["masses"], cat["idx_cat"], cat["obj_size"], 0.0, "com", True)
    return cat, rr, cen, {}


def _oracle(big, vel, reduced, shell, local):
    from oracle import cp_oracle as O
    cat, rr, cen, cache = big
    key = (vel, reduced, shell, local)
    if key not in cache:
        cache[key] = O.morph(cat["xyz"], cat["vxyz"] if vel else None, cat["masses"], cat["r200"], cat["idx_cat"],
                             cat["obj_size"], 0.0, rr if local else None, 1e-2, 100, 10, "com", reduced, shell, local,
                             SAFE=6.0, centers=cen)
    return cache[key]


SCHED = [dict(shells_per_pass=1, large_halo_min=1 << 40, cluster_size=0),     # warp kernel only
         dict(shells_per_pass=2, large_halo_min=1 << 40, cluster_size=0),
         dict(shells_per_pass=4, large_halo_min=1 << 40, cluster_size=0),
         dict(shells_per_pass=4, large_halo_min=1, cluster_size=1),           # CTA per task
         dict(shells_per_pass=1, large_halo_min=1, cluster_size=2),           # clusters + DSMEM
         dict(shells_per_pass=2, large_halo_min=2000, cluster_size=4),
         dict(shells_per_pass=4, large_halo_min=1, cluster_size=8),
         dict(shells_per_pass=4, large_halo_min=5000, cluster_size=16),
         dict(shells_per_pass=0, large_halo_min=0, cluster_size=0),           # auto
         dict(shells_per_pass=4, large_halo_min=1 << 40, cluster_size=0, radial_order=0),   # catalogue order
         dict(shells_per_pass=2, large_halo_min=3000, cluster_size=2, radial_order=0),
         dict(shells_per_pass=4, shell_groups=2, large_halo_min=1 << 40, cluster_size=0),   # 8 shells per warp task
         dict(shells_per_pass=4, shell_groups=2, large_halo_min=8000, cluster_size=2, radial_order=0),
         dict(shells_per_pass=4, large_halo_min=1 << 40, cluster_size=0, grid_halo_min=2000),   # grid-per-iteration path
         dict(shells_per_pass=2, large_halo_min=1 << 40, cluster_size=0, grid_halo_min=6000, radial_order=0),
         dict(shells_per_pass=4, large_halo_min=4000, cluster_size=2, prefix_table=0),   # radial order, no prefix table
         # warp kernel 1d (two task slots per warp share one step; forced: the default takes it for long queues only)
         dict(shells_per_pass=4, large_halo_min=1 << 40, cluster_size=0, async_step=3),
         dict(shells_per_pass=2, large_halo_min=1 << 40, cluster_size=0, async_step=3, radial_order=0),
         dict(shells_per_pass=1, large_halo_min=1 << 40, cluster_size=0, async_step=3),
         dict(shells_per_pass=4, large_halo_min=5000, cluster_size=2, async_step=3, prefix_table=0),
         # warp kernel 1c (pass warps + asynchronous step warp), kernel 1 (a warp per task, its own step)
         dict(shells_per_pass=4, large_halo_min=1 << 40, cluster_size=0, async_step=1),
         dict(shells_per_pass=2, large_halo_min=1 << 40, cluster_size=0, async_step=1, radial_order=0),
         dict(shells_per_pass=4, large_halo_min=1 << 40, cluster_size=0, async_step=0)]


@pytest.mark.parametrize("sched", SCHED, ids=lambda s: "S%d_L%d_C%d_R%d_G%d_grid%d" % (s["shells_per_pass"], min(s["large_halo_min"], 99999), s["cluster_size"], s.get("radial_order", 1), s.get("shell_groups", 0), s.get("grid_halo_min", 0)) + ("_notable" if s.get("prefix_table", 1) == 0 else "") + ("_async%d" % s["async_step"] if "async_step" in s else ""))
def test_oracle_shapes_all_schedules(gpu, big, sched):
    S, D, sess = gpu
    cat, rr, cen, _ = big
    X, M, R2, I, Z = _args(cat)
    sess.options.clear()
    sess.options.update(dict(radial_order=1), **sched) if "radial_order" not in sched else sess.options.update(sched)
    try:
        for reduced in (False, True):
            for shell in (False, True):
                ref, (nit, npt, _, _) = _oracle(big, False, reduced, shell, True)
                got = S.calcMorphLocal(X, M, R2, I, Z, 0.0, rr, 1e-2, 100, 10, "com", reduced, shell, centers=cen)
                tag = "local r%d s%d %s" % (reduced, shell, sched)
                P.assert_ints(S.last_info["n_iter"], nit, tag + " n_iter")
                P.assert_ints(S.last_info["n_pts"], npt, tag + " n_pts")
                P.assert_shape_tuple(got, ref, tag)
            ref, (nit, npt, _, _) = _oracle(big, False, reduced, False, False)
            got = S.calcMorphGlobal(X, M, R2, I, Z, 0.0, 1e-2, 100, 10, "com", 6.0, reduced, centers=cen)
            P.assert_ints(S.last_info["n_iter"], nit, "global n_iter")
            P.assert_shape_tuple(got, ref, "global r%d %s" % (reduced, sched))
    finally:
        sess.options.clear()


@pytest.mark.parametrize("sched", [SCHED[2], SCHED[5], SCHED[8], SCHED[10], SCHED[11], SCHED[13]],
                         ids=["warpS4", "clusterS2C4", "auto", "clusterS2C2_catalogue_order", "warpS4G2", "gridS4"])
def test_oracle_velocity_dispersion(gpu, big, sched):
    S, D, sess = gpu
    cat, rr, cen, _ = big
    X, M, R2, I, Z = _args(cat)
    V = cat["vxyz"]
    sess.options.clear()
    sess.options.update(sched)
    try:
        for reduced in (False, True):
            for shell in (False, True):
                ref, (nit, npt, vc, _) = _oracle(big, True, reduced, shell, True)
                got = S.calcMorphLocalVelDisp(X, V, M, R2, I, Z, 0.0, rr, 1e-2, 100, 10, "com", reduced, shell,
                                              centers=cen)
                P.assert_ints(S.last_info["n_iter"], nit, "vlocal n_iter")
                P.assert_ints(S.last_info["n_pts"], npt, "vlocal n_pts")
                P.assert_shape_tuple(got, ref, "vlocal r%d s%d" % (reduced, shell))
                ok = Z > 0
                P.assert_close(S.last_info["vcenters"][ok], vc[ok], "vcenters")
            ref, (nit, npt, _, _) = _oracle(big, True, reduced, False, False)
            got = S.calcMorphGlobalVelDisp(X, V, M, R2, I, Z, 0.0, 1e-2, 100, 10, "com", 6.0, reduced, centers=cen)
            P.assert_ints(S.last_info["n_iter"], nit, "vglobal n_iter")
            P.assert_shape_tuple(got, ref, "vglobal r%d" % reduced)
    finally:
        sess.options.clear()


def test_oracle_densities(gpu, big):
    from oracle import cp_oracle as O
    S, D, sess = gpu
    cat, rr, cen, _ = big
    X, M, R2, I, Z = _args(cat)
    rb = np.logspace(-2, 0, 40)
    ref, cnt = O.calcDensProfsSphDirectBinning(X, M, R2, rb, I, Z, 0.0, "com", centers=cen, return_counts=True)
    got = D.calcDensProfsSphDirectBinning(X, M, R2, rb, I, Z, 0.0, "com", centers=cen)
    P.assert_ints(D.last_info["counts"], cnt, "sph counts")
    P.assert_close(got, ref, "sph")
    ref = O.calcDensProfsKernelBased(X, M, R2, rb, I, Z, 0.0, "com", centers=cen)
    got = D.calcDensProfsKernelBased(X, M, R2, rb, I, Z, 0.0, "com", centers=cen)
    P.assert_close(got, ref, "kernel")
    sh, _ = _oracle(big, False, False, False, True)
    d, q, s, minor, inter, major = sh[:6]
    ref, cnt = O.calcDensProfsEllDirectBinning(X, M, R2, rb, d, d * q, d * s, major, inter, minor, I, Z, 0.0, "com",
                                               centers=cen, return_counts=True)
    got = D.calcDensProfsEllDirectBinning(X, M, R2, rb, d, d * q, d * s, major, inter, minor, I, Z, 0.0, "com",
                                          centers=cen)
    P.assert_ints(D.last_info["counts"], cnt, "ell counts")
    P.assert_close(got, ref, "ell")


def test_edge_cases(gpu):
    S, D, sess = gpu
    rng = np.random.default_rng(0)
    X = rng.normal(size=(300, 3)) + 50.0
    M = np.ones(300)
    I = np.arange(300, dtype=np.int32)
    Z = np.array([5, 195, 0, 100], dtype=np.int32)
    R2 = np.array([1.0, 0.0, 1.0, 1.0])
    rr = np.logspace(-1, 0, 4)
    first = None
    for opts in ({}, dict(async_step=3), dict(async_step=1), dict(async_step=3, shells_per_pass=1), dict(async_step=1, shells_per_pass=2)):
        # (every warp kernel: dead tasks -- empty object, r200 == 0 -- are skipped inside the task queue)
        sess.options.clear()
        sess.options.update(opts)
        try:
            got = S.calcMorphLocal(X, M, R2, I, Z, 0.0, rr, 1e-2, 100, 10, "com", False, False)
        finally:
            sess.options.clear()
        assert np.isnan(got[1][:3]).all() and np.isnan(got[0][:3]).all()
        assert np.isfinite(got[1][3]).any()
        assert (S.last_info["n_iter"][0] == 0).all()        # 5 particles < IT_MIN
        assert (S.last_info["n_iter"][1] == -1).all()       # r200 == 0
        assert (S.last_info["n_iter"][2] == -1).all()       # empty object
        np.testing.assert_allclose(got[7], [5.0, 195.0, 0.0, 100.0])
        if first is None:
            first = (got, S.last_info["n_iter"].copy(), S.last_info["n_pts"].copy())
        else:   # same kernels' arithmetic whatever the schedule: identical integers, values to rounding
            assert np.array_equal(S.last_info["n_iter"], first[1]) and np.array_equal(S.last_info["n_pts"], first[2]), opts
            np.testing.assert_allclose(got[1], first[0][1], rtol=1e-12, equal_nan=True)
    assert D.calcDensProfsSphDirectBinning(X, M, R2, rr, I[:0], Z[:0], 0.0, "com").shape == (0, 4)
    # catalogue index out of range must be refused, not dereferenced
    from cosmic_profiles_b200._lib import CosmicGpuError
    bad = I.copy()
    bad[7] = 300
    with pytest.raises(CosmicGpuError):
        S.calcMorphLocal(X, M, R2, bad, Z, 0.0, rr, 1e-2, 100, 10, "com", False, False,
                         centers=np.zeros((4, 3)))


@pytest.mark.parametrize("kind", ["main", "pbc"])
def test_device_centres(gpu, kind):
    """cpg_centers ('com' and shrinking-sphere 'mode', shape-path and density-path start radii, periodic box)
    against the reference's numpy centre finding (golden calcMassesCenters + the oracle's restatement)."""
    from cosmic_profiles_b200 import centers as C
    from cosmic_profiles_b200 import session as sess_mod
    from oracle import cp_oracle as O
    S, D, sess = gpu
    z, meta, cat = P.load_golden(kind)
    X, M, R2, I, Z = _args(cat)
    L = meta["L_BOX"]
    ctx = sess_mod.get_context(0)
    ctx.set_particles(X, None, M)
    ctx.set_catalogue(I, Z)
    ref = C.reference_points(X, I, Z) if L != 0.0 else None
    for center in ("com", "mode"):
        for shape_path in (True, False):
            want = O.centres(X, M, I, Z, L, center, shape_path)
            got = ctx.centers(1 if center == "mode" else 0, shape_path, L, ref)
            assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), (center, shape_path)
        # the public density-path driver (recentred centres + masses) against the compiled reference
        cen, m = D.calcMassesCenters(X, M, I, Z, L, center)
        P.assert_close(cen, z["masses_centers_%s/centers" % center], "centers " + center, rtol=1e-12)
        P.assert_close(m, z["masses_centers_%s/m" % center], "m", rtol=1e-12)


def test_unusual_radii_and_limits(gpu, big):
    """Shell lists the radial ordering cannot use (non-ascending, a single shell, more than 255 shells), a tight
    IT_WALL and a large IT_MIN: same integers and values as the oracle."""
    from oracle import cp_oracle as O
    S, D, sess = gpu
    cat, _, cen, _ = big
    X, M, R2, I, Z = _args(cat)
    cases = [
        (np.array([0.5, 0.1, 1.0, 0.3]), 1e-2, 100, 10),            # not ascending
        (np.array([0.7]), 1e-2, 100, 10),                             # one shell
        (np.logspace(-1.2, 0, 300), 1e-2, 100, 10),                   # R > 255
        (np.logspace(-1.5, 0, 9), 1e-3, 4, 10),                       # IT_WALL hit by most shells
        (np.logspace(-1.5, 0, 9), 1e-2, 100, 2500),                   # IT_MIN rejects the inner shells
        (np.logspace(-1.2, 0, 200), 1e-2, 100, 10, dict(async_step=3)),  # kernel 1d with its largest radius / prefix tables
        (np.array([0.7]), 1e-2, 100, 10, dict(async_step=3)),         # ... and one shell per task slot
    ]
    sub = slice(3, 12)   # a few mid-size objects keep the R=300 case cheap for the CPU oracle
    off = cat["offsets"]
    idx = I[off[3]:off[12]]
    for case in cases:
        rr, tol, wall, itmin = case[:4]
        sess.options.clear()
        sess.options.update(case[4] if len(case) > 4 else {})
        for shell in (False, True):
            ref, (nit, npt, _, _) = O.morph(X, None, M, R2[sub], idx, Z[sub], 0.0, rr, tol, wall, itmin, "com", False,
                                            shell, True, centers=cen[sub])
            got = S.calcMorphLocal(X, M, R2[sub], idx, Z[sub], 0.0, rr, tol, wall, itmin, "com", False, shell,
                                   centers=cen[sub])
            tag = "R=%d tol=%g wall=%d itmin=%d shell=%d" % (len(rr), tol, wall, itmin, shell)
            P.assert_ints(S.last_info["n_iter"], nit, tag + " n_iter")
            P.assert_ints(S.last_info["n_pts"], npt, tag + " n_pts")
            P.assert_shape_tuple(got, ref, tag)
    sess.options.clear()


def test_centres_cache_keeps_the_gather_resident(gpu, big):
    """The reference's class recomputes the centres in every driver from the same arrays; the device returns the cached
    ones, so getShapeCatGlobal right after getShapeCatLocal neither recomputes centres nor gathers again -- and a new
    snapshot invalidates the cache."""
    S, D, sess = gpu
    cat, rr, _, _ = big
    X, M, R2, I, Z = _args(cat)
    a = S.calcMorphLocal(X, M, R2, I, Z, 0.0, rr, 1e-2, 100, 10, "com", False, False)
    t_local = S.last_info["timing"][0]
    b = S.calcMorphGlobal(X, M, R2, I, Z, 0.0, 1e-2, 100, 10, "com", 6.0, False)
    t_global = S.last_info["timing"][0]
    assert t_local["gather_ms"] > 0.0
    assert t_global["gather_ms"] < 0.02, t_global            # resident SoA reused (an event pair around nothing)
    assert np.array_equal(a[6], b[6])                        # same centres
    X2 = X + 0.25                                            # another snapshot: new centres, new gather
    c = S.calcMorphGlobal(X2, M, R2, I, Z, 0.0, 1e-2, 100, 10, "com", 6.0, False)
    assert S.last_info["timing"][0]["gather_ms"] > 0.0
    np.testing.assert_allclose(c[6], b[6] + 0.25, rtol=0, atol=1e-12)
    np.testing.assert_allclose(c[1], b[1], rtol=1e-9)        # q of a translated catalogue


def test_device_centres_large_objects(gpu):
    """Objects above the cluster threshold of cpg_centers (8 CTAs per object, slice-local compaction, partial sums
    exchanged through distributed shared memory): same 'mode' / 'com' centres as the numpy restatement, with and
    without a periodic box."""
    from cosmic_profiles_b200 import centers as C
    from cosmic_profiles_b200 import session as sess_mod
    from oracle import cp_oracle as O
    rng = np.random.default_rng(77)
    sizes = np.array([40000, 700, 131075, 32768, 32767, 5, 90001], dtype=np.int32)
    L = 50.0
    parts = []
    for n in sizes:
        c = rng.uniform(0.0, L, 3)
        r = 0.4 * rng.random(n) ** 2.0           # cuspy: many shrinking steps before fewer than 5 are left
        u = rng.normal(size=(n, 3))
        u /= np.linalg.norm(u, axis=1)[:, None]
        parts.append(c + r[:, None] * u * np.array([1.0, 0.7, 0.5]))
    X = np.ascontiguousarray(np.concatenate(parts))
    M = np.ascontiguousarray(rng.uniform(0.5, 1.5, len(X)))
    I = rng.permutation(len(X)).astype(np.int32)
    # catalogue order = a permutation of the particle array, objects stay contiguous in idx_cat
    owner = np.repeat(np.arange(len(sizes)), sizes)
    I = np.concatenate([rng.permutation(np.flatnonzero(owner == h)) for h in range(len(sizes))]).astype(np.int32)
    ctx = sess_mod.get_context(0)
    for box in (0.0, L):
        Xb = np.ascontiguousarray(np.mod(X, L)) if box else X
        ctx.set_particles(Xb, None, M)
        ctx.set_catalogue(I, sizes)
        ref = C.reference_points(Xb, I, sizes) if box else None
        for center in ("mode", "com"):
            for shape_path in (True, False):
                want = O.centres(Xb, M, I, sizes, box, center, shape_path)
                got = ctx.centers(1 if center == "mode" else 0, shape_path, box, ref)
                assert np.abs(got - want).max() <= 1e-11 * max(np.abs(want).max(), 1.0), (box, center, shape_path)


def test_device_shape_interpolation(gpu):
    """cpg_shape_interp against its NumPy twin (itself pinned bit-for-bit against scipy.interpolate.interp1d in
    tests/test_interp_rows.py): NaN abscissae from unconverged shells, unsorted abscissae, extrapolation on both
    sides, an object without a converged shell -- bit-identical, NaN pattern included -- and the ellipsoidal
    binning fed from the device-resident profiles equals the binning fed from the host-interpolated ones."""
    from cosmic_profiles_b200 import _lib, dens_profs_algos as D
    from cosmic_profiles_b200.mock import make_catalogue
    from tests.test_interp_rows import _case
    rb = np.logspace(-2, 0.3, 23)
    for seed, shuffle, nanf in ((1, False, 0.15), (2, True, 0.15), (9, False, 0.0)):
        r200, a, b, c, mj, it, mn = _case(seed, shuffle=shuffle, nan_frac=nanf)
        n = len(r200)
        cat = make_catalogue([300] * n, seed=seed)
        ctx = _lib.Context(0)
        ctx.set_particles(cat["xyz"], None, cat["masses"])
        ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
        with np.errstate(all="ignore"):
            want = D.interpolate_shapes(r200, rb, a, b, c, mj, it, mn)
        got = ctx.shape_interp(r200, rb, D.ell_bin_edges(rb), a, b, c, mj, it, mn, fetch=True)
        for w, g, nm in zip(want, got, "a b c major inter minor".split()):
            assert w.shape == g.shape, nm
            assert np.array_equal(w, g, equal_nan=True), "%s differs (seed %d)" % (nm, seed)
        cen = np.array([cat["xyz"][cat["idx_cat"][300 * h:300 * (h + 1)]].mean(0) for h in range(n)])
        d_dev, c_dev = ctx.dens_ell(cen, None, None, None, None, None, None, 0.0)
        d_host, c_host = ctx.dens_ell(cen, *want, 0.0)
        # (bin masses are accumulated with shared-memory atomics: the last bit depends on the arrival order)
        assert np.array_equal(c_dev, c_host) and np.allclose(d_dev, d_host, rtol=1e-12, atol=0.0, equal_nan=True)
        with pytest.raises(_lib.CosmicGpuError):     # the host call above replaced the resident profiles
            ctx.dens_ell(cen, None, None, None, None, None, None, 0.0)
        ctx.close()


def test_direct_result_stores_equal_the_copied_results(gpu, big):
    """cpg_shape with page-locked result buffers lets the kernels store results straight into them; with the option
    off (or pageable buffers) the results are copied after the kernels.  Both must deliver the same bytes."""
    from cosmic_profiles_b200 import _lib
    cat, _, cen, _ = big
    rr = np.logspace(-1.5, 0, 9)
    ctx = _lib.Context(0)
    ctx.set_particles(cat["xyz"], None, cat["masses"])
    ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
    res = {}
    for direct in (1, 0):
        ctx.set_option("direct_results", direct)
        out, nit, npt, m, _ = ctx.shape(cen, cat["r200"], rr, True, 0.0, 0.0, 1e-2, 100, 10, True, True, False)
        res[direct] = (out.copy(), nit.copy(), npt.copy(), m.copy())
    for a, b in zip(res[1], res[0]):
        assert np.array_equal(a, b)
    assert (res[1][1] > 0).any() and (res[1][0] == 0).any()      # converged and failed halo-shells both present
    ctx.close()


def test_densities_are_bit_reproducible(gpu):
    """Bin masses are accumulated with integer atomics on a fixed-point grid (cpg_dens.cuh): the binned densities do not
    depend on the order in which the threads of a chunk find the members.  The same call twice on one context and once
    on a fresh context: bit-identical.  With the gathered particles in a different order (radial ordering left behind by
    a local-shape call) the chunks hold other particles, so only the order of the final per-chunk sums may differ:
    equal counts, densities within a few ulp."""
    S, D, sess = gpu
    from cosmic_profiles_b200 import _lib
    from cosmic_profiles_b200.mock import make_catalogue
    cat = make_catalogue([20000, 3000, 150000, 800, 41000], seed=12, shuffle_particles=True)
    rb = np.logspace(-2, 0, 40)
    cen = np.stack([np.average(cat["xyz"][cat["idx_cat"][a:b]], axis=0) for a, b in
                    zip(np.cumsum(cat["obj_size"]) - cat["obj_size"], np.cumsum(cat["obj_size"]))])
    a = np.ones((5, 41)) * D.ell_bin_edges(rb)[None, :] * cat["r200"][:, None]
    ax = np.tile(np.eye(3)[None, None], (5, 40, 1, 1))
    axes = [np.ascontiguousarray(ax[:, :, k]) for k in range(3)]

    def both(ctx):
        d_s, c_s = ctx.dens_sph(cen, cat["r200"], D.sph_bin_edges(rb), 0.0)
        d_e, c_e = ctx.dens_ell(cen, a, 0.8 * a, 0.6 * a, *axes, 0.0)
        return d_s, c_s, d_e, c_e
    runs = []
    for fresh in range(2):
        ctx = _lib.Context(0)
        try:
            ctx.set_particles(cat["xyz"], None, cat["masses"])
            ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
            runs.append(both(ctx))
            runs.append(both(ctx))
            if fresh:   # leave a radially ordered gather behind: the density kernels then see another particle order
                ctx.shape(cen, cat["r200"], np.logspace(-1.5, 0, 12), True, 0.0, 0.0, 1e-2, 100, 10, False, False, False)
                other = both(ctx)
        finally:
            ctx.close()
    for r in runs[1:]:
        for x, y in zip(r, runs[0]):
            assert np.array_equal(x, y, equal_nan=True)
    assert runs[0][1].sum() > 100000 and np.isfinite(runs[0][0]).all()
    assert np.array_equal(other[1], runs[0][1]) and np.array_equal(other[3], runs[0][3])
    for k in (0, 2):
        ok = np.isfinite(runs[0][k])
        assert np.array_equal(np.isfinite(other[k]), ok)
        assert np.max(np.abs(other[k][ok] - runs[0][k][ok]) / np.maximum(np.abs(runs[0][k][ok]), 1e-300)) < 1e-14
