#!/usr/bin/env python3
"""Subprocess body of tests/test_public_api_gadget.py (one compiled-reference flavour per process: this one needs the
PATCHED flavour, whose calcMorphLocalVelDisp and ellipsoidal binning work -- SURVEY.md 8(c) defects (i), (ii)).

Builds a small Gadget-HDF5-layout snapshot + FoF/SUBFIND group table on disk in the format oracle/stubs/h5py serves,
drives the reference's own DensShapeProfsGadget (reader, FoF table loader, calcObjCat, classes) through
getShapeCatLocal / getShapeCatGlobal / getShapeCatVelLocal / estDensProfs(spherical=False) / estDensProfs(kernel), once
as shipped and once after cosmic_profiles_b200.install() + install_gadget(), and compares the public outputs.
Prints 'OK <n objects>' on success."""
import os
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from tests import parity  # noqa: E402


def build_files(tmp):
    import h5py
    from cosmic_profiles_b200.mock import make_catalogue
    cat = make_catalogue([4000, 1500, 2500, 800, 6000, 3000], seed=33, profile="einasto", with_velocities=True)
    rng = np.random.default_rng(5)
    sizes = cat["obj_size"].astype(np.int64)
    off = np.concatenate(([0], np.cumsum(sizes)))
    pos, vel, fof, nsub, shlen = [], [], [], [], []
    for g, n in enumerate(sizes):
        c = cat["xyz"][off[g]:off[g + 1]].mean(axis=0)
        extra = 50 + 10 * g                                # a small second subhalo + unbound fuzz after the central one
        pos += [cat["xyz"][off[g]:off[g + 1]], c + rng.normal(size=(extra, 3)) * 0.3]
        vel += [cat["vxyz"][off[g]:off[g + 1]], rng.normal(size=(extra, 3)) * 200.0]
        fof.append(n + extra)
        nsub.append(2)
        shlen += [n, 30]
    # a group without subhalos and one whose central subhalo is below MIN_NUMBER_PTCS
    pos.append(rng.uniform(10, 20, size=(40, 3))); vel.append(rng.normal(size=(40, 3))); fof.append(40); nsub.append(0)
    pos = np.concatenate(pos) * 1000.0                      # kpc/h in the file
    vel = np.concatenate(vel)
    npart = len(pos)
    time_a = 0.64
    table = np.zeros(6); table[1] = 0.00123456
    npt = np.zeros(6, dtype=np.int32); npt[1] = npart
    nall = np.zeros(6, dtype=np.uint32); nall[1] = npart
    snap = os.path.join(tmp, "snap_016")
    h5py.write_stub_file(snap + ".hdf5", {"PartType1/Coordinates": pos.astype(np.float32), "PartType1/Velocities": vel.astype(np.float32)},
                         {"Time": np.float64(time_a), "Redshift": np.float64(1 / time_a - 1), "BoxSize": np.float64(100000.0),
                          "NumFilesPerSnapshot": np.int32(1), "Omega0": np.float64(0.3089), "OmegaLambda": np.float64(0.6911),
                          "HubbleParam": np.float64(0.6774), "MassTable": table, "NumPart_ThisFile": npt, "NumPart_Total": nall,
                          "Flag_Cooling": np.int32(0)})
    gdir = os.path.join(tmp, "groups_016")
    os.makedirs(gdir)
    glen = np.zeros((len(fof), 6), dtype=np.int32); glen[:, 1] = fof
    slen = np.zeros((len(shlen), 6), dtype=np.int32); slen[:, 1] = shlen
    r200 = np.concatenate((cat["r200"], [0.5])).astype(np.float32) * 1000.0
    # two group files, as SUBFIND writes them
    cut_g, cut_s = 4, 8
    for i, (ga, gb, sa, sb) in enumerate(((0, cut_g, 0, cut_s), (cut_g, len(fof), cut_s, len(shlen)))):
        h5py.write_stub_file(os.path.join(gdir, "fof_subhalo_tab_016.%d.hdf5" % i),
                             {"Group/GroupLenType": glen[ga:gb], "Group/Group_R_TopHat200": r200[ga:gb], "Group/Group_R_Mean200": r200[ga:gb],
                              "Group/GroupNsubs": np.asarray(nsub[ga:gb], dtype=np.int32), "Subhalo/SubhaloLenType": slen[sa:sb]}, {})
    return snap, gdir


class lapack_signs:
    """estDensProfs(spherical=False) interpolates the eigenvector COMPONENTS of the local shape profile between shells
    (dens_profs_algos.pyx:176-195).  LAPACK's eigenvectors carry arbitrary signs, so where two neighbouring shells
    come out with opposite signs the interpolated frame is an artefact of that choice, and the binned densities change
    with it (by factors, driver-level check in DESIGN.md section 6): through the public class, parity is defined up to
    the eigenvector signs, like the shapes themselves.  This context manager gives the device driver the reference's
    sign choice (it runs the reference's calcMorphLocal on the same arguments and flips the device vectors to match),
    so that everything else on that path -- interpolation, binning, units -- is compared value for value."""

    def __init__(self, module, reference_fn):
        self.module, self.reference_fn = module, reference_fn

    def __enter__(self):
        self.gpu_fn = self.module.calcMorphLocal
        if self.gpu_fn is self.reference_fn:
            return self

        def aligned(*args):
            got = list(self.gpu_fn(*args))
            ref = self.reference_fn(*args)
            for k in (3, 4, 5):
                sign = np.sign(np.nansum(got[k] * ref[k], axis=-1))[..., None]
                got[k] = got[k] * np.where(sign == 0, 1.0, sign)
            return tuple(got)
        self.module.calcMorphLocal = aligned
        return self

    def __exit__(self, *exc):
        self.module.calcMorphLocal = self.gpu_fn
        return False


def public_run(cp, tmp, snap, gdir, katz, rbins, reference_morph_local=None):
    cp.updateInUnitSystem(length_in_cm="kpc/h", mass_in_g="E+10Msun/h", velocity_in_cm_per_s=1e5, little_h=0.6774)
    cp.updateOutUnitSystem(length_in_cm="kpc/h", mass_in_g="Msun/h", velocity_in_cm_per_s=1e5, little_h=0.6774)
    os.makedirs(os.path.join(tmp, "viz"), exist_ok=True)
    os.makedirs(os.path.join(tmp, "cat"), exist_ok=True)
    with ref_loader.quiet_stdout(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c = cp.DensShapeProfsGadget(snap, gdir, "dm", "016", os.path.join(tmp, "viz"), os.path.join(tmp, "cat"), RVIR_OR_R200="Rvir",
                                    MIN_NUMBER_PTCS=1000, CENTER="com")
        objs = list(range(len(c.getIdxCat()[1])))
        res = dict(objs=objs, sizes=np.asarray(c.getIdxCat()[1]), r200=np.asarray(c.getR200()),
                   loc=c.getShapeCatLocal(objs, katz_config=katz), glob=c.getShapeCatGlobal(objs, katz_config=katz),
                   vloc=c.getShapeCatVelLocal(objs, katz_config=katz),
                   kde=c.estDensProfs(rbins, objs, direct_binning=False, spherical=True, katz_config=katz))
        import cosmic_profiles.common.cosmic_base_class as cbc
        with lapack_signs(cbc, reference_morph_local or cbc.calcMorphLocal):
            res["ell"] = c.estDensProfs(rbins, objs, direct_binning=True, spherical=False, katz_config=katz)
    return res


def main():
    ref_loader.load("patched")
    import cosmic_profiles as cp
    import cosmic_profiles_b200 as cpb
    katz = {"ROverR200": np.logspace(-1.5, 0, 24), "IT_TOL": 1e-2, "IT_WALL": 100, "IT_MIN": 10, "REDUCED": False, "SHELL_BASED": False}
    rbins = np.logspace(-1.7, 0, 20)
    with tempfile.TemporaryDirectory() as tmp:
        snap, gdir = build_files(tmp)
        ref = public_run(cp, os.path.join(tmp, "ref"), snap, gdir, katz, rbins)
        import cosmic_profiles.common.cosmic_base_class as cbc
        reference_morph_local = cbc.calcMorphLocal
        cpb.install()
        cpb.install_gadget()
        try:
            got = public_run(cp, os.path.join(tmp, "gpu"), snap, gdir, katz, rbins, reference_morph_local)
        finally:
            cpb.uninstall()
    from cosmic_profiles_b200 import shape_profs_algos as SA
    assert SA.last_info["timing"][0]["hot_kernel_launches"] > 0          # the device ran
    assert got["objs"] == ref["objs"] and len(ref["objs"]) == 5           # the 800-particle halo and the two odd groups are out
    assert np.array_equal(got["sizes"], ref["sizes"]) and np.array_equal(got["r200"], ref["r200"])
    for key in ("loc", "glob", "vloc"):
        g, r = got[key], ref[key]
        assert g.dtype == r.dtype and g.shape == r.shape, key
        assert np.array_equal(g["is_conv"], r["is_conv"]), key
        assert r["is_conv"].any(), key
        for f in ("d", "q", "s"):
            parity.assert_close(g[f], r[f], key + "/" + f)
        for f in ("minor", "inter", "major"):
            parity.assert_axes(g[f], r[f], r["q"], r["s"], f, key + "/" + f)
    bad = np.argwhere(np.isnan(got["ell"]) != np.isnan(ref["ell"]))
    if len(bad):
        print("ell NaN pattern differs at", bad.tolist(), "got", got["ell"][tuple(bad.T)], "ref", ref["ell"][tuple(bad.T)], file=sys.stderr)
        print("got row", got["ell"][bad[0][0]], "\nref row", ref["ell"][bad[0][0]], file=sys.stderr)
        print("loc d row", ref["loc"]["d"][bad[0][0]], "\nq", ref["loc"]["q"][bad[0][0]], file=sys.stderr)
    parity.assert_close(got["ell"], ref["ell"], "estDensProfs(spherical=False)")
    parity.assert_close(got["kde"], ref["kde"], "estDensProfs kernel")
    assert np.isfinite(ref["ell"]).any()
    print("OK %d" % len(ref["objs"]))


if __name__ == "__main__":
    main()
