import sys, numpy as np
sys.path.insert(0, '.')
import bench
from cosmic_profiles_b200 import _lib
tab = bench.halo_table(10000, 20260101)
cat = bench.generate_on_device(tab, 20260101, 0)
ctx = _lib.Context(0)
ctx.set_particles(cat["xyz"], None, cat["masses"]); ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
K = (1e-2, 100, 10)
_, nit, npt, _, _ = ctx.shape(cat["centers"], cat["r200"], bench.RR, True, 0.0, 0.0, *K, False, False, False)
nit = nit.copy(); npt = npt.copy()
sizes = cat["obj_size"].astype(np.int64)
print("halo-shells", nit.size, "hit the wall (n_iter==100):", int((nit == 100).sum()), " n_iter>=50:", int((nit >= 50).sum()))
h = np.bincount(nit.clip(0).ravel(), minlength=101)
print("histogram of n_iter (0..20):", h[:21].tolist())
pi = (nit.clip(0).astype(np.int64) * sizes[:, None])
print("particle-iterations total %.3e, in wall shells %.3e (%.1f %%)" % (pi.sum(), pi[nit == 100].sum(), 100.0 * pi[nit == 100].sum() / pi.sum()))
# iterations weighted by prefix size (n_pts as proxy)
w = nit.clip(0).astype(np.int64) * npt.clip(0).astype(np.int64)
print("n_iter*n_pts total %.3e, wall shells %.3e (%.1f %%)" % (w.sum(), w[nit == 100].sum(), 100.0 * w[nit == 100].sum() / w.sum()))
# which objects / shells
idx = np.argwhere(nit == 100)
print("first few (object, shell, N, n_pts):", [(int(a), int(b), int(sizes[a]), int(npt[a, b])) for a, b in idx[:12]])
