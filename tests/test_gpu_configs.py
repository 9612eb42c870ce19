"""BASELINE.json configuration flavours on the GPU, and size-independent properties at a size the oracle could
not cover in seconds: schedule invariance, permutation invariance, rigid-motion
invariance."""
import numpy as np
import pytest

from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from cosmic_profiles_b200 import session, shape_profs_algos
    session.options.clear()
    yield shape_profs_algos, session
    session.options.clear()
    session.close_all()


def _args(cat):
    return cat["xyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"]


def test_config1_like(gpu):
    """configs[0]: 100 NFW halos x 10^4 particles, getShapeCatLocal over 20 shells, reduced tensor."""
    from cosmic_profiles_b200.mock import make_catalogue
    from oracle import cp_oracle as O
    S, sess = gpu
    cat = make_catalogue([10000] * 100, seed=20260101, profile="nfw", q_range=(0.55, 0.65), s_range=(0.18, 0.25))
    rr = np.logspace(-1.5, 0, 20)
    cen = O.centres(cat["xyz"], cat["masses"], cat["idx_cat"], cat["obj_size"], 0.0, "com", True)
    for shell in (False, True):
        ref, (nit, npt, _, _) = O.morph(cat["xyz"], None, cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"],
                                        0.0, rr, 1e-2, 100, 10, "com", True, shell, True, centers=cen)
        got = S.calcMorphLocal(*_args(cat), 0.0, rr, 1e-2, 100, 10, "com", True, shell, centers=cen)
        P.assert_ints(S.last_info["n_iter"], nit, "config1 n_iter shell=%d" % shell)
        P.assert_ints(S.last_info["n_pts"], npt, "config1 n_pts shell=%d" % shell)
        P.assert_shape_tuple(got, ref, "config1 shell=%d" % shell)


@pytest.mark.parametrize("cluster", [0, 16, -1])
def test_config5_like_large_halo(gpu, cluster):
    """configs[4] flavour: one large halo, shell-based non-reduced profile on the cluster/DSMEM path."""
    from cosmic_profiles_b200.mock import make_catalogue
    from oracle import cp_oracle as O
    S, sess = gpu
    cat = make_catalogue([400000], seed=77, profile="nfw", q_range=(0.7, 0.7001), s_range=(0.5, 0.5001), equal_mass=True)
    rr = np.logspace(-2, 0, 20)
    cen = O.centres(cat["xyz"], cat["masses"], cat["idx_cat"], cat["obj_size"], 0.0, "com", True)
    ref, (nit, npt, _, _) = O.morph(cat["xyz"], None, cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"], 0.0,
                                    rr, 1e-2, 100, 10, "com", False, True, True, centers=cen)
    sess.options.clear()
    if cluster > 0:
        sess.options.update(cluster_size=cluster, large_halo_min=1)
    elif cluster < 0:   # the whole-grid-per-iteration path that objects of >= 2^22 particles take
        sess.options.update(grid_halo_min=100000)
    try:
        got = S.calcMorphLocal(*_args(cat), 0.0, rr, 1e-2, 100, 10, "com", False, True, centers=cen)
    finally:
        sess.options.clear()
    P.assert_ints(S.last_info["n_iter"], nit, "config5 n_iter")
    P.assert_ints(S.last_info["n_pts"], npt, "config5 n_pts")
    P.assert_shape_tuple(got, ref, "config5")


@pytest.fixture(scope="module")
def large():
    from cosmic_profiles_b200.mock import make_catalogue, power_law_sizes
    sizes = power_law_sizes(1200, 1000, 300000, -1.9, seed=11)
    cat = make_catalogue(sizes, seed=12, profile="einasto")
    off = cat["offsets"]
    # exact-enough centres, shared by every variant below (centres are an input of the path)
    cen = np.array([np.average(cat["xyz"][off[h]:off[h + 1]], axis=0, weights=cat["masses"][off[h]:off[h + 1]])
                    for h in range(len(sizes))])
    return cat, cen


def test_large_schedule_and_permutation_invariance(gpu, large):
    S, sess = gpu
    cat, cen = large
    rr = np.logspace(-1.5, 0, 70)
    sess.options.clear()
    base = S.calcMorphLocal(*_args(cat), 0.0, rr, 1e-2, 100, 10, "com", False, False, centers=cen)
    nit0, npt0 = S.last_info["n_iter"].copy(), S.last_info["n_pts"].copy()
    assert (nit0 > 0).mean() > 0.9
    # (1) a completely different schedule: one shell per pass, catalogue order, clusters of 4 for objects >= 20000
    sess.options.update(shells_per_pass=1, radial_order=0, cluster_size=4, large_halo_min=20000)
    try:
        alt = S.calcMorphLocal(*_args(cat), 0.0, rr, 1e-2, 100, 10, "com", False, False, centers=cen)
        P.assert_ints(S.last_info["n_iter"], nit0, "schedule invariance n_iter")
        P.assert_ints(S.last_info["n_pts"], npt0, "schedule invariance n_pts")
        P.assert_shape_tuple(alt, base, "schedule invariance", rtol=1e-11)
    finally:
        sess.options.clear()
    # (2) permuting the catalogue order inside every object changes nothing but fp64 summation order
    rng = np.random.default_rng(3)
    off = cat["offsets"]
    idx = cat["idx_cat"].copy()
    for h in range(len(cat["obj_size"])):
        idx[off[h]:off[h + 1]] = rng.permutation(idx[off[h]:off[h + 1]])
    perm = S.calcMorphLocal(cat["xyz"], cat["masses"], cat["r200"], idx, cat["obj_size"], 0.0, rr, 1e-2, 100, 10, "com",
                            False, False, centers=cen)
    P.assert_ints(S.last_info["n_iter"], nit0, "permutation invariance n_iter")
    P.assert_ints(S.last_info["n_pts"], npt0, "permutation invariance n_pts")
    P.assert_shape_tuple(perm, base, "permutation invariance", rtol=1e-11)


def test_large_rigid_motion_invariance(gpu, large):
    """Axis ratios are invariant under a rigid rotation + translation of the snapshot; axes co-rotate."""
    from cosmic_profiles_b200.mock import random_rotations
    S, sess = gpu
    cat, cen = large
    rr = np.logspace(-1.5, 0, 24)
    base = S.calcMorphLocal(*_args(cat), 0.0, rr, 1e-2, 100, 10, "com", True, False, centers=cen)
    nit0 = S.last_info["n_iter"].copy()
    Rm = random_rotations(np.random.default_rng(5), 1)[0]
    shift = np.array([3.0, -2.0, 7.0])
    xyz2 = cat["xyz"] @ Rm.T + shift
    cen2 = cen @ Rm.T + shift
    rot = S.calcMorphLocal(xyz2, cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"], 0.0, rr, 1e-2, 100, 10, "com",
                           True, False, centers=cen2)
    nit1 = S.last_info["n_iter"]
    # a rotated snapshot rounds differently: allow the (rare) boundary particle to flip an iteration count
    assert (nit1 != nit0).mean() < 1e-3
    same = (nit1 == nit0) & np.isfinite(base[1]) & np.isfinite(rot[1])
    assert np.abs(rot[1][same] / base[1][same] - 1).max() < 1e-9      # q
    assert np.abs(rot[2][same] / base[2][same] - 1).max() < 1e-9      # s
    major0 = base[5][same] @ Rm.T
    dots = np.abs(np.sum(major0 * rot[5][same], axis=-1))
    gap = 1 - base[1][same] ** 2
    assert (1 - dots[gap > 1e-3]).max() < 1e-12


def test_velocity_centres_equal_mass_closed_form(gpu):
    """Dark-matter-only snapshots (all masses equal) take the closed-form route of the sequential fp32 mass
    normaliser of helper_class.calcCoM (cpg_gather.cuh: mass32_equal); it must agree with the oracle's literal
    sequential loop, including for an object large enough that fp32 accumulation visibly drifts."""
    from cosmic_profiles_b200.mock import make_catalogue
    from oracle import cp_oracle as O
    S, sess = gpu
    cat = make_catalogue([300000, 70000, 5000, 1234, 77], seed=21, equal_mass=True, with_velocities=True)
    cen = O.centres(cat["xyz"], cat["masses"], cat["idx_cat"], cat["obj_size"], 0.0, "com", True)
    ref, (nit, npt, vc, _) = O.morph(cat["xyz"], cat["vxyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"],
                                     0.0, None, 1e-2, 100, 10, "com", False, False, False, SAFE=6.0, centers=cen)
    got = S.calcMorphGlobalVelDisp(cat["xyz"], cat["vxyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"],
                                   0.0, 1e-2, 100, 10, "com", 6.0, False, centers=cen)
    # the fp32 normaliser differs from the fp64 mass by ~1e-7..1e-5 relative here; the closed form must hit it
    m64 = np.array([cat["masses"][0] * n for n in cat["obj_size"]])
    drift = np.abs(np.linalg.norm(vc, axis=1) / np.linalg.norm(S.last_info["vcenters"], axis=1) - 1)
    assert drift.max() < 1e-13, drift
    P.assert_ints(S.last_info["n_iter"], nit, "vglobal equal-mass n_iter")
    P.assert_shape_tuple(got, ref, "vglobal equal-mass")
    assert m64.shape == (5,)
