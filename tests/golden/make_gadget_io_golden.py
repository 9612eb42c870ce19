#!/usr/bin/env python3
"""Golden vectors for the snapshot-block arithmetic of the Gadget row: the reference's OWN
reader, readgadget.read_block / read_field / header (gadget/readgadget.py:34-52, :75-124, :127-194), runs unmodified
on snapshot files served by the in-memory-format h5py stand-in (oracle/stubs/h5py: h5py itself is absent here), and
its float64 outputs are stored next to the raw datasets in tests/golden/gadget_io_golden.npz.  Cases: float32 and
float64 datasets, Masses dataset present / MassTable constant, expansion factor != 1 (velocity sqrt(a) factor), a unit
system whose factors are not representable in float32, and a snapshot split over two files.
Run in the build container (needs oracle/_ref)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def header(npart_this, npart_total, nfiles, mass_table, time):
    return {"Time": np.float64(time), "Redshift": np.float64(1.0 / time - 1.0), "BoxSize": np.float64(40000.0),
            "NumFilesPerSnapshot": np.int32(nfiles), "Omega0": np.float64(0.3089), "OmegaLambda": np.float64(0.6911),
            "HubbleParam": np.float64(0.6774), "MassTable": np.asarray(mass_table, dtype=np.float64),
            "NumPart_ThisFile": np.asarray(npart_this, dtype=np.int32), "NumPart_Total": np.asarray(npart_total, dtype=np.uint32),
            "Flag_Cooling": np.int32(0)}


def main():
    ref_loader.load("pristine")
    import h5py
    from cosmic_profiles.common import config
    from cosmic_profiles.gadget import readgadget
    rng = np.random.default_rng(20261020)
    out, cases = {}, []
    units = {"kpc_h": (3.085678e21 / 0.6774, 1.989e43 / 0.6774, 1e5, 0.6774),           # Gadget default: kpc/h, 1e10 Msun/h, km/s
             "odd": (3.085678e21 * 1.2345, 1.989e33 * 7.7e9, 0.98765e5, 0.7)}             # factors with long mantissas
    with tempfile.TemporaryDirectory() as tmp:
        for name, dt, with_mass, time, unit, nfiles in (("f32_masses", np.float32, True, 1.0, "kpc_h", 1),
                                                         ("f32_masstable_a05", np.float32, False, 0.5, "kpc_h", 1),
                                                         ("f64_masses_a037", np.float64, True, 0.37, "odd", 1),
                                                         ("f32_two_files", np.float32, False, 0.8, "odd", 2)):
            n = 2000
            L, M, V, h = units[unit]
            config.initialize(2, 128, L, M, V, h, 3.085678e24 / h, 1.989e43 / h, 1e5)
            pos = rng.uniform(0, 40000.0, size=(n, 3)).astype(dt)
            vel = (rng.standard_normal(size=(n, 3)) * 300.0).astype(dt)
            mass = (rng.uniform(0.5, 1.5, size=n) * 0.0123).astype(dt)
            table = np.zeros(6)
            table[1] = 0.0 if with_mass else 0.012345678
            cuts = [0, n] if nfiles == 1 else [0, 850, n]
            snap = os.path.join(tmp, name)
            for i in range(nfiles):
                a, b = cuts[i], cuts[i + 1]
                ds = {"PartType1/Coordinates": pos[a:b], "PartType1/Velocities": vel[a:b]}
                if with_mass:
                    ds["PartType1/Masses"] = mass[a:b]
                npt, nall = np.zeros(6, dtype=np.int32), np.zeros(6, dtype=np.uint32)
                npt[1], nall[1] = b - a, n
                fn = snap + (".hdf5" if nfiles == 1 else ".%d.hdf5" % i)
                h5py.write_stub_file(fn, ds, header(npt, nall, nfiles, table, time))
            got = {blk: readgadget.read_block(snap, blk, [1]) for blk in ("POS ", "VEL ", "MASS")}
            lf, mf, vf = (config.getLMVInternal()[0] / config.InUnitLength_in_cm, config.getLMVInternal()[1] / config.InUnitMass_in_g,
                          config.getLMVInternal()[2] / config.InUnitVelocity_in_cm_per_s)
            out.update({name + "/pos": pos, name + "/vel": vel, name + "/time": np.float64(time),
                        name + "/mass_table": np.float64(table[1]), name + "/unit_fac": np.array([lf, mf, vf]),
                        name + "/cuts": np.asarray(cuts), name + "/ref_pos": got["POS "], name + "/ref_vel": got["VEL "],
                        name + "/ref_mass": got["MASS"]})
            if with_mass:
                out[name + "/mass"] = mass
            cases.append(name)
            print(name, got["POS "].dtype, got["POS "].shape, float(got["MASS"][0]))
    out["cases"] = np.array(cases)
    out["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gadget_io_golden.npz"), **out)


if __name__ == "__main__":
    main()
