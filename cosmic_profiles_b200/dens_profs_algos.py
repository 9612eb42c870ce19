"""Drop-in replacements for the four density drivers of the reference, same positional signatures:
calcMassesCenters / calcDensProfsSphDirectBinning / calcDensProfsEllDirectBinning / calcDensProfsKernelBased
(cosmic_profiles/dens_profs/dens_profs_algos.pyx:15, :54, :112, :221).

Host side keeps only what the reference does in Python: centres (:39-44, :96-101), bin-edge construction
(:83, :155-156) and the scipy interp1d of the shape profiles onto the bins (:163-195, same scipy).  Binning
and the kernel estimator run in libcosmicgpu.so.  No CPU fallback.
"""
import numpy as np

from . import _lib, centers as _centers, session as _session

last_info = {}


def _prep(xyz, masses, idx_cat, obj_size):
    return (_lib.f64(np.asarray(xyz)), _lib.f64(np.asarray(masses)), _lib.i32(np.asarray(idx_cat)),
            _lib.i32(np.asarray(obj_size)))


def _context(xyz, masses, idx_cat, obj_size):
    ctx = _session.get_context((_session.devices or [0])[0])
    ctx.set_particles(xyz, None, masses)
    ctx.set_catalogue(idx_cat, obj_size)
    return ctx


def _centres(ctx, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, centers):
    """Density paths: calcMode starts from rad = 1000 (dens_profs_algos.pyx:42)."""
    if centers is not None:
        return _lib.f64(centers)
    if _session.gpu_centers:
        ref = _centers.reference_points(xyz, idx_cat, obj_size) if L_BOX != 0.0 else None
        return ctx.centers(1 if CENTER == 'mode' else 0, False, L_BOX, ref)
    return _centers.compute_centers(xyz, masses, idx_cat, obj_size, L_BOX, CENTER, False)


def calcMassesCenters(xyz, masses, idx_cat, obj_size, L_BOX, CENTER):
    xyz, masses, idx_cat, obj_size = _prep(xyz, masses, idx_cat, obj_size)
    if len(obj_size) == 0:
        return np.zeros((0, 3)), np.zeros(0)
    ctx = _context(xyz, masses, idx_cat, obj_size)
    cen = _centres(ctx, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, None)
    m = ctx.obj_masses()
    return _centers.recentre(cen, L_BOX), m


def sph_bin_edges(r_over_r200):
    r = _lib.f64(np.asarray(r_over_r200))
    return np.hstack(([np.float64(1e-8), (r[:-1] + r[1:]) / 2., r[-1]]))


def ell_bin_edges(r_over_r200):
    r = _lib.f64(np.asarray(r_over_r200))
    mid = (r[:-1] + r[1:]) / 2
    return np.hstack(([np.float64(mid[0] * mid[0] / mid[1]), mid, r[-1]]))


def calcDensProfsSphDirectBinning(xyz, masses, r200s, r_over_r200, idx_cat, obj_size, L_BOX, CENTER, centers=None):
    xyz, masses, idx_cat, obj_size = _prep(xyz, masses, idx_cat, obj_size)
    R = len(r_over_r200)
    if len(obj_size) == 0:
        return np.zeros((0, R), dtype=np.float64)
    ctx = _context(xyz, masses, idx_cat, obj_size)
    cen = _centres(ctx, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, centers)
    dens, cnt = ctx.dens_sph(cen, r200s, sph_bin_edges(r_over_r200), L_BOX)
    last_info.clear()
    last_info.update(counts=cnt, timing=[ctx.timing()])
    return dens


def interpolate_shapes(r200s, r_over_r200, a, b, c, major, inter, minor):
    """dens_profs_algos.pyx:163-195, with the very scipy the reference would call."""
    from scipy.interpolate import interp1d
    r = _lib.f64(np.asarray(r_over_r200))
    edges = ell_bin_edges(r)
    n, R = a.shape[0], len(r)
    ai, bi, ci = (np.zeros((n, R + 1)) for _ in range(3))
    mj, it, mn = (np.zeros((n, R, 3)) for _ in range(3))
    kw = dict(bounds_error=False, fill_value='extrapolate')
    for p in range(n):
        r_vec = np.float64(a[p])
        ai[p] = interp1d(r_vec, a[p], **kw)(edges * r200s[p])
        bi[p] = interp1d(r_vec, b[p], **kw)(edges * r200s[p])
        ci[p] = interp1d(r_vec, c[p], **kw)(edges * r200s[p])
        for k in range(3):
            mj[p, :, k] = interp1d(r_vec, major[p, :, k], **kw)(r * r200s[p])
            it[p, :, k] = interp1d(r_vec, inter[p, :, k], **kw)(r * r200s[p])
            mn[p, :, k] = interp1d(r_vec, minor[p, :, k], **kw)(r * r200s[p])
    return ai, bi, ci, mj, it, mn


def calcDensProfsEllDirectBinning(xyz, masses, r200s, r_over_r200, a, b, c, major, inter, minor, idx_cat, obj_size,
                                  L_BOX, CENTER, centers=None):
    xyz, masses, idx_cat, obj_size = _prep(xyz, masses, idx_cat, obj_size)
    R = len(r_over_r200)
    if len(obj_size) == 0:
        return np.zeros((0, R), dtype=np.float64)
    r200s = _lib.f64(np.asarray(r200s))
    a, b, c = (_lib.f64(np.asarray(v)) for v in (a, b, c))
    major, inter, minor = (_lib.f64(np.asarray(v)) for v in (major, inter, minor))
    assert a.shape[0] == b.shape[0] and b.shape[0] == c.shape[0]
    ai, bi, ci, mj, it, mn = interpolate_shapes(r200s, r_over_r200, a, b, c, major, inter, minor)
    ctx = _context(xyz, masses, idx_cat, obj_size)
    cen = _centres(ctx, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, centers)
    dens, cnt = ctx.dens_ell(cen, ai, bi, ci, mj, it, mn, L_BOX)
    last_info.clear()
    last_info.update(counts=cnt, timing=[ctx.timing()])
    return dens


def calcDensProfsKernelBased(xyz, masses, r200s, r_over_r200, idx_cat, obj_size, L_BOX, CENTER, centers=None):
    xyz, masses, idx_cat, obj_size = _prep(xyz, masses, idx_cat, obj_size)
    R = len(r_over_r200)
    if len(obj_size) == 0:
        return np.zeros((0, R), dtype=np.float64)
    ctx = _context(xyz, masses, idx_cat, obj_size)
    cen = _centres(ctx, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, centers)
    dens = ctx.dens_kernel(cen, r200s, r_over_r200, L_BOX)
    last_info.clear()
    last_info.update(timing=[ctx.timing()])
    return dens
