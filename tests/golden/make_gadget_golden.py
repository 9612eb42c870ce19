#!/usr/bin/env python3
"""Golden vectors for the Gadget / FoF catalogue row: runs the COMPILED reference's
calcObjCat (oracle/_ref/pristine, gadget/gen_catalogues.pyx:39-95) on seeded FoF/SUBFIND-style group tables and
stores inputs + outputs in tests/golden/gadget_golden.npz.  Run in the build container (needs oracle/_ref)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def group_table(rng, n_fof, p_nosub=0.15, max_sub=5, lo=5, hi=4000):
    nb_shs = rng.integers(1, max_sub + 1, size=n_fof).astype(np.int32)
    nb_shs[rng.uniform(size=n_fof) < p_nosub] = 0
    fof = np.zeros(n_fof, dtype=np.int32)
    sh_len = []
    for p in range(n_fof):
        tot = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        if nb_shs[p] > 0:
            w = np.sort(rng.uniform(size=nb_shs[p]))[::-1] + 0.05
            bound = int(tot * rng.uniform(0.6, 1.0))       # SUBFIND: part of the FoF group is unbound "fuzz"
            lens = np.maximum((w / w.sum() * bound).astype(np.int64), 0)
            sh_len.extend(lens.tolist())
        fof[p] = tot
    return nb_shs, np.asarray(sh_len, dtype=np.int32), fof, rng.uniform(0.05, 2.0, size=n_fof)


def main():
    ref_loader.load("pristine")
    from cosmic_profiles.gadget.gen_catalogues import calcObjCat
    rng = np.random.default_rng(20260320)
    out = {}
    cases = [("a", 300, 200), ("b", 64, 20), ("c", 500, 100000), ("d", 7, 1), ("e", 3000, 300)]
    for name, n_fof, minp in cases:
        nb, sl, fof, r200 = group_table(rng, n_fof)
        if name == "d":
            nb[:] = np.array([0, 2, 0, 1, 0, 0, 3], dtype=np.int32)
            sl = np.array([9, 1, 4, 0, 7, 2], dtype=np.int32)
            fof[:] = np.array([3, 12, 5, 4, 1, 1, 9], dtype=np.int32)
        cat, o_r200, o_size = calcObjCat(nb, sl, fof, r200, np.int32(minp))
        out.update({name + "_nb_shs": nb, name + "_sh_len": sl, name + "_fof_sizes": fof, name + "_group_r200": r200,
                    name + "_min": np.int32(minp), name + "_obj_cat": np.asarray(cat, dtype=np.int32),
                    name + "_obj_r200": np.asarray(o_r200), name + "_obj_size": np.asarray(o_size, dtype=np.int32)})
        print(name, n_fof, "->", len(o_size), "objects,", len(cat), "indices")
    out["cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gadget_golden.npz"), **out)


if __name__ == "__main__":
    main()
