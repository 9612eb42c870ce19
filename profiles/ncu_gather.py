#!/usr/bin/env python3
"""One local-shape call on 3000 objects of the configs[1] catalogue, for `ncu -k regex:gather_bucketed|bucket_hist`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import _lib  # noqa: E402

tab = bench.halo_table(3000, 20260101)
cat = bench.generate_on_device(tab, 20260101, 0)
ctx = _lib.Context(0)
K = (1e-2, 100, 10)
ctx.set_particles(cat["xyz"], None, cat["masses"]); ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
for _ in range(2):
    ctx.set_option("invalidate_gather", 1)
    ctx.shape(cat["centers"], cat["r200"], bench.RR, True, 0.0, 0.0, *K, False, False, False)
