#!/usr/bin/env python3
"""Parity of the three density estimators at BASELINE.json configs[2] scale (measurement record; the CPU side takes
a minute): the full configs[1] catalogue on the GPU, the oracle's C restatement on a seeded random sample of whole
objects.  Integer bin counts must match exactly, densities to 1e-9.

    python profiles/parity_dens_at_scale.py [--budget 2e7]
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import _lib, dens_profs_algos as D  # noqa: E402
from oracle import cp_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--budget", type=float, default=2e7)
a = ap.parse_args()
warnings.filterwarnings("ignore")
K = (1e-2, 100, 10)
tab = bench.halo_table(10000, 20260101)
cat = bench.generate_on_device(tab, 20260101, 0)
sizes = cat["obj_size"].astype(np.int64)
rb = np.logspace(-2, 0, 50)
ctx = _lib.Context(0)
ctx.set_particles(cat["xyz"], None, cat["masses"])
ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
cen, r200 = cat["centers"], cat["r200"]
d_sph, c_sph = ctx.dens_sph(cen, r200, D.sph_bin_edges(rb), 0.0)
d_kde = ctx.dens_kernel(cen, r200, rb, 0.0)
o, _, _, _, _ = ctx.shape(cen, r200, bench.RR, True, 0.0, 0.0, *K, False, False, False)
o = o.copy()
o[o == 0.0] = np.nan
sa, sb, sc = o[:, :, 0], o[:, :, 0] * o[:, :, 1], o[:, :, 0] * o[:, :, 2]
smj, sit, smn = np.ascontiguousarray(o[:, :, 3:6]), np.ascontiguousarray(o[:, :, 6:9]), np.ascontiguousarray(o[:, :, 9:12])
ctx.shape_interp(r200, rb, D.ell_bin_edges(rb), sa, sb, sc, smj, sit, smn)
d_ell, c_ell = ctx.dens_ell(cen, None, None, None, None, None, None, 0.0)

rng = np.random.default_rng(4)
pick, tot = [], 0
for h in rng.permutation(10000):
    if tot + sizes[h] <= a.budget:
        pick.append(int(h))
        tot += int(sizes[h])
pick = np.sort(np.asarray(pick))
xyz, m, idx, sz, r2 = bench.reference_sample_inputs(cat, pick)
cores = bench.use_all_host_cores()
t0 = time.perf_counter()
r_sph, rc_sph = O.calcDensProfsSphDirectBinning(xyz, m, r2, rb, idx, sz, 0.0, "com", nthreads=cores, centers=cen[pick], return_counts=True)
r_kde = O.calcDensProfsKernelBased(xyz, m, r2, rb, idx, sz, 0.0, "com", nthreads=cores, centers=cen[pick])
# the oracle interpolates with scipy itself, one object at a time (dens_profs_algos.pyx:163-195)
r_ell, rc_ell = O.calcDensProfsEllDirectBinning(xyz, m, r2, rb, sa[pick], sb[pick], sc[pick], smj[pick], sit[pick], smn[pick], idx, sz,
                                                0.0, "com", nthreads=cores, centers=cen[pick], return_counts=True)
cpu_s = time.perf_counter() - t0


def rel(got, ref):
    nanpat = int((np.isnan(got) != np.isnan(ref)).sum())
    ok = np.isfinite(ref) & (ref != 0)
    return (float(np.max(np.abs(got[ok] - ref[ok]) / np.abs(ref[ok]))) if ok.any() else 0.0), nanpat


res = dict(workload="configs[2]: 10000 objects, %d particles, 50 bins logspace(-2,0); oracle sample %d objects / %d particles"
                    % (int(sizes.sum()), len(pick), tot),
           bins_compared=int(rc_sph.size),
           sph_count_mismatches=int((c_sph[pick] != rc_sph).sum()), sph_max_rel=rel(d_sph[pick], r_sph)[0],
           ell_count_mismatches=int((c_ell[pick] != rc_ell).sum()), ell_max_rel=rel(d_ell[pick], r_ell)[0],
           ell_nan_pattern_mismatches=rel(d_ell[pick], r_ell)[1], ell_nan_bins=int(np.isnan(r_ell).sum()),
           kernel_max_rel=rel(d_kde[pick], r_kde)[0],
           members_binned_sph=int(rc_sph.sum()), members_binned_ell=int(rc_ell.sum()),
           oracle="oracle/cp_oracle.c + scipy interp1d per object, %d threads, %.1f s" % (cores, cpu_s))
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "parity_dens_at_scale.json"), "w"), indent=1)
