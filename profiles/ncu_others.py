#!/usr/bin/env python3
"""One invocation of every non-morphology kernel family on the configs[1]/[2] catalogue (5000 objects), for
`ncu --set full -k regex:...` (profiles/capture_others.sh).  Measures nothing itself."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import _lib, dens_profs_algos as D, dens_profs_tools as DT, gen_catalogues as G  # noqa: E402

warnings.filterwarnings("ignore")
nh = 3000
rounds = 1 if '--once' in sys.argv else 2
tab = bench.halo_table(nh, 20260101)
cat = bench.generate_on_device(tab, 20260101, 0)
ctx = _lib.Context(0)
K = (1e-2, 100, 10)
rb = np.logspace(-2, 0, 50)
for rep in range(rounds):
    ctx.set_particles(cat["xyz"], None, cat["masses"]); ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
    ctx.centers(0, True, 0.0, None)
    ctx.centers(1, True, 0.0, None)
    o, nit, _, _, _ = ctx.shape(cat["centers"], cat["r200"], bench.RR, True, 0.0, 0.0, *K, False, False, False)
    o = o.copy(); o[o == 0.0] = np.nan
    dens, _ = ctx.dens_sph(cat["centers"], cat["r200"], D.sph_bin_edges(rb), 0.0)
    ctx.dens_kernel(cat["centers"], cat["r200"], rb, 0.0)
    d_, q_, s_ = o[:, :, 0], o[:, :, 1], o[:, :, 2]
    ai, bi, ci, mj, it, mn = D.interpolate_shapes(cat["r200"], rb, d_, d_ * q_, d_ * s_, o[:, :, 3:6], o[:, :, 6:9], o[:, :, 9:12])
    ctx.dens_ell(cat["centers"], ai, bi, ci, mj, it, mn, 0.0)
    sl = cat["obj_size"].astype(np.int32)
    ctx.obj_cat(np.ones(nh, np.int32), sl, sl, cat["r200"], 200, fetch_idx=False)
    dd = dens[:, 10:] * 1e10
    dd[dd == 0.0] = np.nan
    from cosmic_profiles_b200 import session
    session.get_context(0)
    DT.fitDensProfHelper(dd, rb[10:], cat["r200"], {"profile": "einasto"})
print("done")
