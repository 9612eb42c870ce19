"""The drop-in claim end to end: the reference's own public classes (compiled from /root/reference into
oracle/_ref -- test infrastructure) are driven through DensShapeProfs.getShapeCatLocal / getShapeCatGlobal /
getMassesCenters / estDensProfs / fitDensProfs, once as shipped and once after cosmic_profiles_b200.install() +
install_fits(); the two sets of public outputs must agree (ints exact, floats 1e-9, fits per tests/test_fit.py's gate).
Mirrors cosmic_profiles/tests/test_shapes.py:27-192 and tests/test_densities.py:27-95 with seeded data and value checks."""
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import parity  # noqa: E402


def _reference():
    from oracle import ref_loader
    if not ref_loader.available("pristine"):
        pytest.skip("oracle/_ref/pristine not built")
    ref_loader.load("pristine")
    import cosmic_profiles
    return cosmic_profiles, ref_loader


def _public_run(cp, ref_loader, tmp, cat, katz, rbins, center):
    cp.updateInUnitSystem(length_in_cm="Mpc/h", mass_in_g="E+10Msun/h", velocity_in_cm_per_s=1e5, little_h=0.6774)
    cp.updateOutUnitSystem(length_in_cm="kpc/h", mass_in_g="Msun/h", velocity_in_cm_per_s=1e5, little_h=0.6774)
    off = np.concatenate(([0], np.cumsum(cat["obj_size"])))
    idx_in = [cat["idx_cat"][off[h]:off[h + 1]].tolist() for h in range(len(cat["obj_size"]))]
    os.makedirs(os.path.join(tmp, "viz"), exist_ok=True)
    os.makedirs(os.path.join(tmp, "cat"), exist_ok=True)
    with ref_loader.quiet_stdout(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c = cp.DensShapeProfs(cat["xyz"], cat["masses"], idx_in, cat["r200"], 0.0, "016", os.path.join(tmp, "viz"),
                              os.path.join(tmp, "cat"), MIN_NUMBER_PTCS=1000, CENTER=center)
        objs = list(range(len(c.getIdxCat()[1])))
        res = dict(objs=objs, mc=c.getMassesCenters(objs), loc=c.getShapeCatLocal(objs, katz_config=katz),
                   glob=c.getShapeCatGlobal(objs, katz_config=katz),
                   sph=c.estDensProfs(rbins, objs, direct_binning=True, spherical=True, katz_config=katz),
                   kde=c.estDensProfs(rbins, objs, direct_binning=False, spherical=True, katz_config=katz))
        res["fit_nfw"] = c.fitDensProfs(res["sph"][:, 10:], rbins[10:], {"profile": "nfw"}, objs)
        res["fit_einasto"] = c.fitDensProfs(res["sph"][:, 10:], rbins[10:], {"profile": "einasto"}, objs)
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("center,reduced,shell", [("mode", False, False), ("com", True, True)])
def test_public_classes_before_and_after_install(tmp_path, center, reduced, shell):
    cp, ref_loader = _reference()
    import cosmic_profiles_b200 as cpb
    from cosmic_profiles_b200.mock import make_catalogue
    cat = make_catalogue([4000, 1500, 2500, 800, 6000, 3000], seed=21, profile="einasto")   # the 800 one is filtered out
    katz = {"ROverR200": np.logspace(-1.5, 0, 24), "IT_TOL": 1e-2, "IT_WALL": 100, "IT_MIN": 10, "REDUCED": reduced,
            "SHELL_BASED": shell}
    rbins = np.logspace(-2, 0, 30)
    ref = _public_run(cp, ref_loader, str(tmp_path / "ref"), cat, katz, rbins, center)
    cpb.install()
    cpb.install_fits()
    try:
        got = _public_run(cp, ref_loader, str(tmp_path / "gpu"), cat, katz, rbins, center)
    finally:
        cpb.uninstall()
    from cosmic_profiles_b200 import dens_profs_tools as DT, shape_profs_algos as SA
    assert SA.last_info["timing"][0]["hot_kernel_launches"] > 0 and len(DT.last_info["nfev"]) == 5   # the device ran
    assert got["objs"] == ref["objs"] and len(ref["objs"]) == 5
    for key in ("loc", "glob"):
        g, r = got[key], ref[key]
        assert g.dtype == r.dtype and g.shape == r.shape, key
        assert np.array_equal(g["is_conv"], r["is_conv"]), key
        for f in ("d", "q", "s"):
            parity.assert_close(g[f], r[f], key + "/" + f)
        for f in ("minor", "inter", "major"):
            parity.assert_axes(g[f], r[f], r["q"], r["s"], f, key + "/" + f)
    parity.assert_close(got["mc"]["mass"], ref["mc"]["mass"], "mass")
    parity.assert_close(got["mc"]["centre"], ref["mc"]["centre"], "centre")
    parity.assert_close(got["sph"], ref["sph"], "estDensProfs spherical")
    parity.assert_close(got["kde"], ref["kde"], "estDensProfs kernel")
    for k in ("fit_nfw", "fit_einasto"):
        g, r = got[k], ref[k]
        assert g.dtype == r.dtype and g.shape == r.shape, k
        for f in r.dtype.names:
            if r[f].dtype == bool:
                assert np.array_equal(g[f], r[f]), (k, f)
                continue
            fin = np.isfinite(r[f])
            assert np.array_equal(np.isfinite(g[f]), fin), (k, f)
            rel = np.abs(g[f][fin] - r[f][fin]) / np.abs(r[f][fin])
            assert rel.max() < 1e-2 and np.median(rel) < 1e-9, (k, f, rel)   # tests/test_fit.py's gate
