#!/usr/bin/env python3
"""One host process driving N GPUs through cpg_multi (what DensShapeProfs.getShapeCatLocal does after install() with
session.devices = [0..N-1]): ONE catalogue of N x 10^4 configs[1]-distribution objects in pageable host memory,
public drivers calcMorphLocal + calcMorphGlobal.  Prints one JSON object: partition / upload / scatter times, per-device
kernel times (imbalance), cold (snapshot distributed) and warm (snapshot cached) wall times.

    python profiles/multi_single_process.py N [halos_per_gpu]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import session, shape_profs_algos as S  # noqa: E402

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
K = (1e-2, 100, 10)
tab = bench.halo_table(per * n_dev, 20260101)
# generate the catalogue in pieces on device 0 (plumbing), keep it in pageable host memory
parts = []
for k in range(n_dev):
    rows = np.arange(k * per, (k + 1) * per)
    c = bench.generate_on_device(bench.table_rows(tab, rows), 20260101 + 1000 * k, 0)
    parts.append({key: np.array(c[key], copy=True) for key in ("xyz", "masses", "obj_size", "r200")})
    del c
xyz = np.concatenate([p["xyz"] for p in parts])
masses = np.concatenate([p["masses"] for p in parts])
size = np.concatenate([p["obj_size"] for p in parts]).astype(np.int32)
r200 = np.concatenate([p["r200"] for p in parts])
idx = np.arange(len(masses), dtype=np.int32)
del parts
session.devices = list(range(n_dev))
out = {"n_devices": n_dev, "objects": int(len(size)), "particles": int(len(masses))}


def pair():
    t0 = time.perf_counter()
    S.calcMorphLocal(xyz, masses, r200, idx, size, 0.0, bench.RR, *K, "com", False, False)
    info_l = dict(S.last_info)
    t1 = time.perf_counter()
    S.calcMorphGlobal(xyz, masses, r200, idx, size, 0.0, *K, "com", bench.SAFE, False)
    t2 = time.perf_counter()
    return 1e3 * (t1 - t0), 1e3 * (t2 - t1), info_l


pair()                                  # first use: buffers, pinned staging
session.close_all_snapshots()
cold = pair()
warm = min((pair() for _ in range(3)), key=lambda r: r[0] + r[1])
nit = warm[2]["n_iter"]
pi = float((nit.clip(0).sum(1).astype(np.int64) * size).sum())
multi = warm[2]["multi"]
kern = [d["timing"]["kernel_ms"] for d in multi["devices"]] if multi else [warm[2]["timing"][0]["kernel_ms"]]
gath = [d["timing"]["gather_ms"] for d in multi["devices"]] if multi else [warm[2]["timing"][0]["gather_ms"]]
out.update(cold_ms={"calcMorphLocal": cold[0], "calcMorphGlobal": cold[1]}, warm_ms={"calcMorphLocal": warm[0], "calcMorphGlobal": warm[1]},
           cold_host=(cold[2]["multi"] or {}).get("host"), warm_host=(multi or {}).get("host"),
           local_kernel_ms_per_device=kern, local_gather_ms_per_device=gath,
           kernel_imbalance_max_over_mean=max(kern) / (sum(kern) / len(kern)),
           particles_per_device=[d["n_part"] for d in multi["devices"]] if multi else [int(len(masses))],
           objects_per_device=[d["n_obj"] for d in multi["devices"]] if multi else [int(len(size))],
           local_particle_iterations=pi, local_PI_per_s_warm=pi / (warm[0] * 1e-3))
print(json.dumps(out))
