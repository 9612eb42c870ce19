"""The CPU restatement (oracle/cp_oracle.c) against golden vectors produced by the compiled reference
(tests/golden/make_golden.py).  CPU-only; this is what pins the oracle."""
import numpy as np
import pytest

from oracle import cp_oracle as O
from tests import parity as P


@pytest.fixture(scope="module", params=["main", "pbc"])
def gold(request):
    return (request.param,) + P.load_golden(request.param)


def _args(cat):
    return cat["xyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"]


def test_local_shapes(gold):
    kind, z, meta, cat = gold
    X, M, R2, I, S = _args(cat)
    k = meta["katz"]
    for center in (("com", "mode") if kind == "main" else ("com",)):
        for reduced in (0, 1):
            for shell in (0, 1):
                tag = "local_%s_r%d_s%d" % (center, reduced, shell)
                got, (nit, npt, _, _) = O.morph(X, None, M, R2, I, S, meta["L_BOX"], cat["rr_shape"], k["IT_TOL"],
                                                k["IT_WALL"], k["IT_MIN"], center, reduced, shell, True)
                P.assert_shape_tuple(got, P.golden_tuple(z, tag), tag)
                P.assert_ints(nit, z[tag + "/n_iter"], tag + "/n_iter")
                P.assert_ints(npt, z[tag + "/n_pts"], tag + "/n_pts")


def test_global_shapes(gold):
    kind, z, meta, cat = gold
    X, M, R2, I, S = _args(cat)
    k = meta["katz"]
    for center in (("com", "mode") if kind == "main" else ("com",)):
        for reduced in (0, 1):
            tag = "global_%s_r%d" % (center, reduced)
            got, (nit, npt, _, _) = O.morph(X, None, M, R2, I, S, meta["L_BOX"], None, k["IT_TOL"], k["IT_WALL"],
                                            k["IT_MIN"], center, reduced, False, False, SAFE=meta["SAFE"])
            P.assert_shape_tuple(got, P.golden_tuple(z, tag), tag)
            P.assert_ints(nit, z[tag + "/n_iter"][:, 0], tag + "/n_iter")
            P.assert_ints(npt, z[tag + "/n_pts"][:, 0], tag + "/n_pts")


def test_velocity_dispersion(gold):
    kind, z, meta, cat = gold
    if "vxyz" not in cat:
        pytest.skip("no velocities in this fixture")
    X, M, R2, I, S = _args(cat)
    V = cat["vxyz"]
    k = meta["katz"]
    for reduced in (0, 1):
        tag = "vglobal_com_r%d" % reduced
        got, (nit, npt, _, _) = O.morph(X, V, M, R2, I, S, meta["L_BOX"], None, k["IT_TOL"], k["IT_WALL"],
                                        k["IT_MIN"], "com", reduced, False, False, SAFE=meta["SAFE"])
        P.assert_shape_tuple(got, P.golden_tuple(z, tag), tag)
        P.assert_ints(nit, z[tag + "/n_iter"][:, 0], tag + "/n_iter")
        for shell in (0, 1):
            tag = "vlocal_com_r%d_s%d" % (reduced, shell)  # vs PATCHED reference (defect ii)
            got, (nit, npt, _, _) = O.morph(X, V, M, R2, I, S, meta["L_BOX"], cat["rr_shape"], k["IT_TOL"],
                                            k["IT_WALL"], k["IT_MIN"], "com", reduced, shell, True)
            P.assert_shape_tuple(got, P.golden_tuple(z, tag), tag)
            P.assert_ints(nit, z[tag + "/n_iter"], tag + "/n_iter")
            P.assert_ints(npt, z[tag + "/n_pts"], tag + "/n_pts")


def test_densities(gold):
    kind, z, meta, cat = gold
    X, M, R2, I, S = _args(cat)
    L = meta["L_BOX"]
    rb = cat["rr_dens"]
    for center in ("com", "mode"):
        P.assert_close(O.calcDensProfsSphDirectBinning(X, M, R2, rb, I, S, L, center), z["dens_sph_" + center],
                       "sph_" + center, rtol=1e-13)
        P.assert_close(O.calcDensProfsKernelBased(X, M, R2, rb, I, S, L, center), z["dens_kernel_" + center],
                       "kernel_" + center, rtol=1e-13)
        cen, m = O.calcMassesCenters(X, M, I, S, L, center)
        P.assert_close(cen, z["masses_centers_%s/centers" % center], "centers", rtol=1e-15)
        P.assert_close(m, z["masses_centers_%s/m" % center], "m", rtol=1e-15)
    tag = "local_com_r0_s0"
    d, q, s = z[tag + "/d"], z[tag + "/q"], z[tag + "/s"]
    got = O.calcDensProfsEllDirectBinning(X, M, R2, rb, d, d * q, d * s, z[tag + "/major"], z[tag + "/inter"],
                                          z[tag + "/minor"], I.copy(), S, L, "com")
    P.assert_close(got, z["dens_ell_com"], "ell_com (vs PATCHED reference, defect i)", rtol=1e-13)


def test_edge_cases():
    """Empty catalogue, object below IT_MIN, r200 == 0."""
    rng = np.random.default_rng(0)
    X = rng.normal(size=(300, 3)) + 50.0
    M = np.ones(300)
    I = np.arange(300, dtype=np.int32)
    S = np.array([5, 195, 100], dtype=np.int32)
    R2 = np.array([1.0, 0.0, 1.0])
    rr = np.logspace(-1, 0, 4)
    got, (nit, npt, _, raw) = O.morph(X[:200], None, M[:200], R2[:2], I[:200], S[:2], 0.0, rr, 1e-2, 100, 10, "com",
                                      0, 0, True)
    assert np.isnan(got[1]).all()          # 5 particles < IT_MIN, and r200 == 0
    assert (nit[0] == 0).all() and (nit[1] == -1).all()
    assert O.calcDensProfsSphDirectBinning(X, M, R2, rr, I[:0], S[:0], 0.0, "com").shape == (0, 4)
