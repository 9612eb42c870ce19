#!/usr/bin/env python3
"""Where the host->device leg of a step goes (configs[1] catalogue, page-locked source arrays):
cpg_set_particles, cpg_set_catalogue (host pass over idx_cat + install), cpg_set_snapshot (both, overlapped), with the
contiguous-range detection on and off.  One JSON object on stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import _lib  # noqa: E402

tab = bench.halo_table(10000, 20260101)
cat = bench.generate_on_device(tab, 20260101, 0)
ctx = _lib.Context(0)
ctx.set_option("snapshot_cache", 0)
out = {}


def t(fn, n=4):
    fn()
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


for det in (1, 0):
    ctx.set_option("detect_ranges", det)
    k = "ranges" if det else "flat"
    out["set_particles_ms"] = t(lambda: ctx.set_particles(cat["xyz"], None, cat["masses"]))
    out["set_catalogue_%s_ms" % k] = t(lambda: ctx.set_catalogue(cat["idx_cat"], cat["obj_size"]))
    out["set_snapshot_%s_ms" % k] = t(lambda: ctx.set_snapshot(cat["xyz"], None, cat["masses"], cat["idx_cat"], cat["obj_size"]))
t0 = time.perf_counter()
ok, start = _lib.catalogue_ranges(cat["idx_cat"], cat["obj_size"])
out["host_scan_alone_ms"] = (time.perf_counter() - t0) * 1e3
out["bytes"] = dict(xyz=int(cat["xyz"].nbytes), masses=int(cat["masses"].nbytes), idx_cat=int(cat["idx_cat"].nbytes))
# pageable source
xyz_p, m_p, idx_p = np.array(cat["xyz"]), np.array(cat["masses"]), np.array(cat["idx_cat"])
ctx.set_option("detect_ranges", 1)
out["pageable_set_snapshot_ms"] = t(lambda: ctx.set_snapshot(xyz_p, None, m_p, idx_p, cat["obj_size"]), 3)
out["pageable_set_particles_ms"] = t(lambda: ctx.set_particles(xyz_p, None, m_p), 3)
print(json.dumps(out))
