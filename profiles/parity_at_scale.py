#!/usr/bin/env python3
"""Parity at BASELINE.json configs[1] scale (measurement record, not a pytest: the CPU side takes minutes).

Runs the full configs[1] step on the GPU (1e4 objects, 9.4e7 particles, 70 shells E1 non-reduced + global shape),
then the oracle (compiled reference where present for the floats, C restatement for the integers) on a seeded random
sample of whole objects, and reports how many of the compared halo-shells differ in the two integers the contract
pins (Katz iteration count, particle count of the last tensor evaluation) and the largest relative q / s difference.

    python profiles/parity_at_scale.py [--budget 2.5e7] [--reduced] [--shell]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import _lib  # noqa: E402
from oracle import cp_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--budget", type=float, default=2.5e7, help="particles in the oracle sample")
ap.add_argument("--halos", type=int, default=10000)
ap.add_argument("--reduced", action="store_true")
ap.add_argument("--shell", action="store_true")
ap.add_argument("--vel", action="store_true", help="velocity-dispersion tensor (calcMorphLocalVelDisp / GlobalVelDisp)")
ap.add_argument("--tag", default="")
a = ap.parse_args()
K = (1e-2, 100, 10)

tab = bench.halo_table(a.halos, 20260101)
cat = bench.generate_on_device(tab, 20260101, 0)
sizes = cat["obj_size"].astype(np.int64)
vxyz = None
if a.vel:   # anisotropic dispersion + a bulk flow per object, Gadget MassTable-style constant mass
    rngv = np.random.default_rng(99)
    vxyz = rngv.standard_normal(size=cat["xyz"].shape) * np.array([300.0, 200.0, 100.0])
    vxyz += np.repeat(rngv.normal(0, 500.0, size=(a.halos, 3)), sizes, axis=0)
    cat["masses"] = np.full(cat["masses"].shape, 0.0123456789)
ctx = _lib.Context(0)
ctx.set_particles(cat["xyz"], vxyz, cat["masses"])
ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
out_l, nit_l, npt_l, _, _ = ctx.shape(cat["centers"], cat["r200"], bench.RR, True, 0.0, 0.0, *K, a.reduced, a.shell, a.vel)
out_l, nit_l, npt_l = out_l.copy(), nit_l.copy(), npt_l.copy()
out_g, nit_g, npt_g, _, _ = ctx.shape(cat["centers"], cat["r200"], None, False, bench.SAFE, 0.0, *K, a.reduced, False, a.vel)
out_g, nit_g, npt_g = out_g.copy(), nit_g.copy(), npt_g.copy()

# sample: seeded random objects, any size (the big ones matter: they take the cluster kernel)
rng = np.random.default_rng(4)
pick, tot = [], 0
for h in rng.permutation(a.halos):
    if tot + sizes[h] <= a.budget:
        pick.append(int(h))
        tot += int(sizes[h])
pick = np.sort(np.asarray(pick))
xyz, m, idx, sz, r200 = bench.reference_sample_inputs(cat, pick)
vsub = None
if a.vel:
    off_all = np.concatenate(([0], np.cumsum(sizes)))
    vsub = np.ascontiguousarray(np.concatenate([vxyz[off_all[h]:off_all[h + 1]] for h in pick]))
cen = cat["centers"][pick]
cores = bench.use_all_host_cores()
t0 = time.perf_counter()
ref_l, (rit_l, rpt_l, _, _) = O.morph(xyz, vsub, m, r200, idx, sz, 0.0, bench.RR, *K, "com", a.reduced, a.shell, True,
                                      nthreads=cores, centers=cen)
ref_g, (rit_g, rpt_g, _, _) = O.morph(xyz, vsub, m, r200, idx, sz, 0.0, None, *K, "com", a.reduced, False, False,
                                      SAFE=bench.SAFE, nthreads=cores, centers=cen)
cpu_s = time.perf_counter() - t0


def rel(got, ref):
    ok = np.isfinite(ref) & (ref != 0)
    g = got.copy()
    g[g == 0.0] = np.nan
    nanpat = int((np.isnan(g) != np.isnan(ref)).sum())
    return float(np.nanmax(np.abs(g[ok] - ref[ok]) / np.abs(ref[ok]))) if ok.any() else 0.0, nanpat


ql, nan_q = rel(out_l[pick][:, :, 1], ref_l[1])
sl, nan_s = rel(out_l[pick][:, :, 2], ref_l[2])
qg, _ = rel(out_g[pick][:, 0, 1], ref_g[1])
sg, _ = rel(out_g[pick][:, 0, 2], ref_g[2])
res = dict(workload="configs[1], %d objects, %d particles; variant reduced=%s shell_based=%s velocity_dispersion=%s" % (a.halos, int(sizes.sum()), a.reduced, a.shell, a.vel),
           sample_objects=int(len(pick)), sample_particles=int(tot), largest_sampled_object=int(sz.max()),
           halo_shells_compared=int(rit_l.size + rit_g.size),
           tensor_evaluations_compared=int(rit_l.clip(0).sum() + rit_g.clip(0).sum()),
           particle_iterations_compared=float((rit_l.clip(0).sum(1) * sz).sum() + (rit_g.clip(0) * sz).sum()),
           n_iter_mismatches_local=int((nit_l[pick] != rit_l).sum()), n_pts_mismatches_local=int((npt_l[pick] != rpt_l).sum()),
           n_iter_mismatches_global=int((nit_g[pick][:, 0] != rit_g).sum()), n_pts_mismatches_global=int((npt_g[pick][:, 0] != rpt_g).sum()),
           nan_pattern_mismatches=nan_q + nan_s, max_rel_q_local=ql, max_rel_s_local=sl, max_rel_q_global=qg, max_rel_s_global=sg,
           oracle="oracle/cp_oracle.c (C restatement, pinned against the compiled reference), %d threads, %.1f s" % (cores, cpu_s))
import zlib  # noqa: E402
if len(pick) == a.halos and not (a.reduced or a.shell or a.vel) and a.halos == 10000:
    # the whole catalogue was compared: record checksums of the (oracle-verified) integer outputs for
    # tests/test_gpu_full_size.py
    clean = all(res[k] == 0 for k in res if "mismatches" in k)
    res["verified_checksums"] = dict(
        clean=clean, particle_iterations=int((nit_l.clip(0).sum(1).astype(np.int64) * sizes).sum() + (nit_g.clip(0)[:, 0].astype(np.int64) * sizes).sum()),
        sum_n_pts_local=int(npt_l.clip(0).astype(np.int64).sum()), sum_n_pts_global=int(npt_g.clip(0).astype(np.int64).sum()),
        crc_n_iter_local=zlib.crc32(np.ascontiguousarray(nit_l).tobytes()), crc_n_pts_local=zlib.crc32(np.ascontiguousarray(npt_l).tobytes()),
        crc_n_iter_global=zlib.crc32(np.ascontiguousarray(nit_g).tobytes()), crc_n_pts_global=zlib.crc32(np.ascontiguousarray(npt_g).tobytes()))
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "parity_at_scale%s.json" % a.tag), "w"), indent=1)
