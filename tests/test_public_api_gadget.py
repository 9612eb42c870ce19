"""The HDF5-variant public class end to end: DensShapeProfsGadget of the compiled reference (PATCHED flavour: the
velocity-dispersion driver and the ellipsoidal binning work, SURVEY.md 8(c) defects (i), (ii)) reads a Gadget-layout
snapshot and a two-file FoF/SUBFIND group table through its own readers (h5py served by oracle/stubs/h5py), builds its
catalogue with calcObjCat, and runs getShapeCatLocal / getShapeCatGlobal / getShapeCatVelLocal /
estDensProfs(spherical=False) / estDensProfs(kernel) -- once as shipped, once after install() + install_gadget().
Runs in a subprocess: the rest of the suite holds the pristine flavour, and only one flavour fits in a process.
(DensShapeProfsGadget.getShapeCatVelGlobal raises TypeError in the reference itself -- defect (iii), above the boundary.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_gadget_class_before_and_after_install():
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    if not ref_loader.available("patched"):
        pytest.skip("oracle/_ref/patched not built")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "gadget_public_run.py")], capture_output=True, text=True,
                         timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.strip().splitlines()[-1] == "OK 5"
