"""Per-task timeline of the local-shape morphology kernels on bench.py's workload (diagnostics build:
make -C cosmic_profiles_b200/csrc OUT=... EXTRA=-DCPG_TASK_TRACE; run with COSMICGPU_LIB pointing at it).
Writes gpurun_out/<tag>_trace.npz: start/end (ns), sm, tiles, steps, N per task for the warp and cluster kernels,
and prints how many warps are busy over time."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from cosmic_profiles_b200 import _lib  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "trace"
opts = sys.argv[2:]
tab = B.halo_table(10000, 20260101)
cat = B.generate_on_device(tab, 20260101, 0)
ctx = _lib.Context(0)
for kv in opts:
    n, v = kv.split("=")
    ctx.set_option(n, int(v))
ctx.set_snapshot(cat["xyz"], None, cat["masses"], cat["idx_cat"], cat["obj_size"])
k = (B.KATZ["IT_TOL"], B.KATZ["IT_WALL"], B.KATZ["IT_MIN"])
for _ in range(3):
    ctx.set_option("invalidate_gather", 1)
    ctx.shape(cat["centers"], cat["r200"], B.RR, True, 0.0, 0.0, *k, False, False, False, zero_copy=True)
    t = ctx.timing()
print("kernel_ms", t["kernel_ms"], "gather_ms", t["gather_ms"])
L = _lib.load()
nw = 2 + 4 * (1 << 20)
buf = np.zeros(nw, dtype=np.uint64)
L.cpg_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
rc = L.cpg_debug_trace(ctx._h if hasattr(ctx, "_h") else ctx.handle, buf.ctypes.data_as(ctypes.c_void_p), nw)
assert rc == 0, rc
tr = buf[2:].reshape(-1, 4)
out = {}
for name, sl in (("warp", slice(0, 1 << 19)), ("cluster", slice(1 << 19, 1 << 20))):
    a = tr[sl]
    ok = a[:, 1] > 0
    a = a[ok]
    out[name] = dict(t0=a[:, 0].astype(np.int64), t1=a[:, 1].astype(np.int64), sm=(a[:, 2] >> np.uint64(32)).astype(np.int32),
                     tiles=(a[:, 2] & np.uint64(0xffffffff)).astype(np.int64), steps=(a[:, 3] >> np.uint64(32)).astype(np.int32),
                     N=(a[:, 3] & np.uint64(0xffffffff)).astype(np.int64), idx=np.nonzero(ok)[0])
t_origin = min(out["warp"]["t0"].min(), out["cluster"]["t0"].min() if len(out["cluster"]["t0"]) else 1 << 62)
for name in out:
    o = out[name]
    if not len(o["t0"]):
        continue
    s = (o["t0"] - t_origin) / 1e6
    e = (o["t1"] - t_origin) / 1e6
    d = e - s
    print("%s kernel: %d tasks, first start %.3f ms, last end %.3f ms, sum of task durations %.1f ms (x warps), longest %.3f ms"
          % (name, len(s), s.min(), e.max(), d.sum(), d.max()))
    A = np.stack([o["tiles"].astype(float), o["steps"].astype(float), np.ones(len(d))], 1)
    coef = np.linalg.lstsq(A, d * 1e3, rcond=None)[0]
    print("   fit of the task durations: %.3f us/tile  %.3f us/step  %.3f us/task   (tiles %.3g, steps %.3g)"
          % (coef[0], coef[1], coef[2], o["tiles"].sum(), o["steps"].sum()))
    order = np.argsort(-d)[:3]
    for i in order:
        print("   task %6d N=%7d tiles=%8d steps=%3d start %.3f dur %.3f ms  (%.1f ns/tile)" % (
            o["idx"][i], o["N"][i], o["tiles"][i], o["steps"][i], s[i], d[i], 1e6 * d[i] / max(1, o["tiles"][i])))
    # busy count over time
    edges = np.linspace(0, e.max(), 41)
    busy = [(np.minimum(e, b1) - np.maximum(s, b0)).clip(0).sum() / (b1 - b0) for b0, b1 in zip(edges[:-1], edges[1:])]
    print("   busy tasks per time slice (%.2f ms each):" % (edges[1] - edges[0]), " ".join("%d" % round(x) for x in busy))
np.savez_compressed(os.path.join(ROOT, "gpurun_out", tag + "_trace.npz"),
                    **{n + "_" + k_: v for n, o in out.items() for k_, v in o.items()})
