#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the compiled reference (oracle/_ref, built by oracle/build_ref.py).

Run in the build container (needs /root/reference once, to build oracle/_ref):
    python oracle/build_ref.py && python tests/golden/make_golden.py
The fixtures pin the oracle AND the CUDA path to the reference's own outputs.  Cases that exercise a
shipped reference defect (ellipsoidal binning, local velocity dispersion)
come from the 'patched' flavour and are labelled so; everything else is produced by the patched
flavour too (it is the only one exporting iteration/particle counts) and cross-checked bit-for-bit
against the pristine flavour in a subprocess.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

KATZ = dict(IT_TOL=1e-2, IT_WALL=100, IT_MIN=10)
RR_SHAPE = np.logspace(-1.5, 0, 10)
RR_DENS = np.logspace(-2, 0, 16)
SAFE = 6.0


def inputs(kind):
    from cosmic_profiles_b200.mock import make_catalogue
    if kind == "main":
        return make_catalogue([400, 900, 1500, 650, 1200, 333, 2000, 64], seed=11, with_velocities=True,
                              shuffle_particles=True), 0.0
    if kind == "pbc":
        cat = make_catalogue([700, 1100, 500], seed=12, with_velocities=False, shuffle_particles=True, box=10.0)
        L = 10.0
        # push the halos across the box faces and wrap
        shift = np.array([[4.9, 0.0, 0.0], [0.0, -4.8, 4.7], [0.0, 0.0, 0.0]])
        off = cat["offsets"]
        xyz = cat["xyz"].copy()
        cen = np.array([xyz[cat["idx_cat"][off[p]:off[p + 1]]].mean(0) for p in range(3)])
        for p in range(3):
            sel = cat["idx_cat"][off[p]:off[p + 1]]
            xyz[sel] += (np.array([5.0, 5.0, 5.0]) - cen[p]) + shift[p]
        cat["xyz"] = np.ascontiguousarray(np.mod(xyz, L))
        return cat, L
    raise ValueError(kind)


def run_reference(flavour, kind):
    from oracle import ref_loader
    ns = ref_loader.load(flavour)
    cat, L = inputs(kind)
    X, M, R2, I, S = cat["xyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"]
    V = cat.get("vxyz")
    n, R = len(S), len(RR_SHAPE)
    out = {}
    instr = flavour == "patched"

    def shapes(tag, fn, *args, Rk=R):
        it = np.full(n * Rk, -1, np.int32)
        pts = np.full(n * Rk, -1, np.int32)
        if instr:
            ns.shape.dbg_setup(it, pts, Rk)
        with ref_loader.quiet_stdout():
            res = fn(*args)
        if instr:
            ns.shape.dbg_clear()
            out[tag + "/n_iter"] = it.reshape(n, Rk)
            out[tag + "/n_pts"] = pts.reshape(n, Rk)
        for k, nm in enumerate(["d", "q", "s", "minor", "inter", "major", "centers", "m"]):
            out[tag + "/" + nm] = np.asarray(res[k])

    k = (KATZ["IT_TOL"], KATZ["IT_WALL"], KATZ["IT_MIN"])
    for center in (("com", "mode") if kind == "main" else ("com",)):
        for reduced in (False, True):
            for shell in (False, True):
                shapes("local_%s_r%d_s%d" % (center, reduced, shell), ns.calcMorphLocal, X, M, R2, I, S, L,
                       RR_SHAPE, *k, center, reduced, shell)
            shapes("global_%s_r%d" % (center, reduced), ns.calcMorphGlobal, X, M, R2, I, S, L, *k, center, SAFE,
                   reduced, Rk=1)
    if V is not None:
        for reduced in (False, True):
            shapes("vglobal_com_r%d" % reduced, ns.calcMorphGlobalVelDisp, X, V, M, R2, I, S, L, *k, "com", SAFE,
                   reduced, Rk=1)
            if flavour == "patched":  # pristine raises ZeroDivisionError (defect ii)
                for shell in (False, True):
                    shapes("vlocal_com_r%d_s%d" % (reduced, shell), ns.calcMorphLocalVelDisp, X, V, M, R2, I, S, L,
                           RR_SHAPE, *k, "com", reduced, shell)
    for center in ("com", "mode"):
        with ref_loader.quiet_stdout():
            out["dens_sph_" + center] = ns.calcDensProfsSphDirectBinning(X, M, R2, RR_DENS, I, S, L, center)
            out["dens_kernel_" + center] = ns.calcDensProfsKernelBased(X, M, R2, RR_DENS, I, S, L, center)
            mc = ns.calcMassesCenters(X, M, I, S, L, center)
        out["masses_centers_%s/centers" % center] = mc[0]
        out["masses_centers_%s/m" % center] = mc[1]
    if flavour == "patched":  # pristine returns all-NaN (defect i)
        d, q, s = out["local_com_r0_s0/d"], out["local_com_r0_s0/q"], out["local_com_r0_s0/s"]
        with ref_loader.quiet_stdout():
            out["dens_ell_com"] = ns.calcDensProfsEllDirectBinning(
                X, M, R2, RR_DENS, d, d * q, d * s, out["local_com_r0_s0/major"], out["local_com_r0_s0/inter"],
                out["local_com_r0_s0/minor"], I.copy(), S, L, "com")
    return cat, L, out


def main():
    if len(sys.argv) == 4 and sys.argv[1] == "--child":
        _, _, out = run_reference(sys.argv[2], sys.argv[3])
        np.savez(os.path.join("/tmp", "golden_%s_%s.npz" % (sys.argv[2], sys.argv[3])), **out)
        return
    for kind in ("main", "pbc"):
        cat, L, out = run_reference("patched", kind)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "pristine", kind], check=True)
        pr = np.load("/tmp/golden_pristine_%s.npz" % kind)
        n_chk = 0
        for key in pr.files:
            assert np.array_equal(pr[key], out[key], equal_nan=True), "pristine != patched for " + key
            n_chk += 1
        with open(os.path.join(ROOT, "oracle", "_ref", "patched", "BUILD_INFO.json")) as f:
            info = json.load(f)
        meta = dict(build=info, L_BOX=L, SAFE=SAFE, katz=KATZ, pristine_checked=n_chk,
                    patched_only=sorted(k for k in out if k not in pr.files))
        arrays = {"in/" + k: v for k, v in cat.items() if isinstance(v, np.ndarray)}
        arrays.update({"in/rr_shape": RR_SHAPE, "in/rr_dens": RR_DENS})
        arrays.update(out)
        path = os.path.join(HERE, "ref_%s.npz" % kind)
        np.savez_compressed(path, meta=json.dumps(meta), **arrays)
        print(kind, "->", path, os.path.getsize(path) // 1024, "KiB;", n_chk, "arrays identical to pristine;",
              len(meta["patched_only"]), "patched-only")


if __name__ == "__main__":
    main()
