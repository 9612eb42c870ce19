"""Cycle accounting of kernel 1c (morph_wasync_kernel) on bench.py's workload (diagnostics build:
make -C cosmic_profiles_b200/csrc OUT=... EXTRA="-DCPG_TASK_TRACE -DCPG_ASYNC_DEFAULT=1"; COSMICGPU_LIB pointing at it)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from cosmic_profiles_b200 import _lib  # noqa: E402

opts = sys.argv[1:]
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
buf = np.zeros(16, dtype=np.uint64)
L.cpg_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
rc = L.cpg_debug_trace(ctx._h if hasattr(ctx, "_h") else ctx.handle, buf.ctypes.data_as(ctypes.c_void_p), 16)
assert rc == 0, rc
w = [int(x) for x in buf]
print("tiles", w[0], "tile-shells", w[1])
print("pass warps: total %.3g cycles, in passes %.1f %%, waiting / fetching %.1f %%; %d passes (%.0f cycles each), %d on resident tiles"
      % (w[2], 100.0 * w[3] / max(w[2], 1), 100.0 * w[4] / max(w[2], 1), w[10], w[3] / max(w[10], 1), w[11]))
print("step warps: total %.3g cycles, busy %.1f %%; %d batches (%.0f cycles each), %.2f slots per batch"
      % (w[6], 100.0 * w[7] / max(w[6], 1), w[8], w[7] / max(w[8], 1), w[9] / max(w[8], 1)))
if w[5]:
    print("kernel 1d: total %.3g warp-cycles: passes %.1f %%, steps %.1f %%, refill %.1f %%; %d rounds, %.2f passes and %.2f shells per round; "
          "%.0f cycles per pass, %.0f per step, %.0f refill per round"
          % (w[2], 100.0 * w[3] / w[2], 100.0 * w[5] / w[2], 100.0 * w[4] / w[2], w[8], w[10] / w[8], w[12] / w[8], w[3] / w[10], w[5] / w[8], w[4] / w[8]))
