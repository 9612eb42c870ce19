#!/usr/bin/env python3
"""Wall time of the public drop-in drivers on the configs[1] catalogue, called the way the reference calls them:
plain (pageable) numpy arrays in, the reference's return tuple out.  Written to gpurun_out/dropin_timing.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import dens_profs_algos as D, shape_profs_algos as S  # noqa: E402

from cosmic_profiles_b200 import session  # noqa: E402
if "--upload-threads" in sys.argv:
    session.options["upload_threads"] = int(sys.argv[sys.argv.index("--upload-threads") + 1])
tab = bench.halo_table(10000, 20260101)
cat = bench.generate_on_device(tab, 20260101, 0)
xyz, m, idx = np.array(cat["xyz"]), np.array(cat["masses"]), np.array(cat["idx_cat"])     # pageable copies
sz, r200 = np.array(cat["obj_size"]), np.array(cat["r200"])
out = {}


def timed(name, fn, n=3):
    fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    out[name] = dict(best_s=min(t), all_s=t)
    print(name, "%.3f s" % min(t), flush=True)


timed("calcMorphLocal_com", lambda: S.calcMorphLocal(xyz, m, r200, idx, sz, 0.0, bench.RR, 1e-2, 100, 10, "com", False, False))
timed("calcMorphLocal_mode", lambda: S.calcMorphLocal(xyz, m, r200, idx, sz, 0.0, bench.RR, 1e-2, 100, 10, "mode", False, False))
timed("calcMorphLocal_com_periodic_box_L100", lambda: S.calcMorphLocal(xyz, m, r200, idx, sz, 100.0, bench.RR, 1e-2, 100, 10, "com", False, False))
timed("calcMorphGlobal_com", lambda: S.calcMorphGlobal(xyz, m, r200, idx, sz, 0.0, 1e-2, 100, 10, "com", 6.0, False))
timed("calcDensProfsSphDirectBinning_com", lambda: D.calcDensProfsSphDirectBinning(xyz, m, r200, np.logspace(-2, 0, 50), idx, sz, 0.0, "com"))
out["particles"] = int(sz.sum())
out["h2d_bytes"] = int(xyz.nbytes + m.nbytes + idx.nbytes)
out["upload_threads"] = session.options.get("upload_threads", 8)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dropin_timing_t%d.json" % out["upload_threads"]), "w"), indent=1)
