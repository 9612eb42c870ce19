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


def _rows(a, objs):
    """Rows of a per-object array that belong to a shard (objs None: the single-device case, all rows)."""
    return a if objs is None else np.ascontiguousarray(a[objs])


def _centres(grp, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, centers):
    """Density paths: calcMode starts from rad = 1000 (dens_profs_algos.pyx:42)."""
    if centers is not None:
        return _lib.f64(centers)
    if _session.gpu_centers:
        ref = _centers.reference_points(xyz, idx_cat, obj_size) if L_BOX != 0.0 else None
        mode = 1 if CENTER == 'mode' else 0
        return grp.rows(lambda ctx, objs: ctx.centers(mode, False, L_BOX, None if ref is None else _rows(ref, objs)))
    return _centers.compute_centers(xyz, masses, idx_cat, obj_size, L_BOX, CENTER, False)


def calcMassesCenters(xyz, masses, idx_cat, obj_size, L_BOX, CENTER):
    xyz, masses, idx_cat, obj_size = _prep(xyz, masses, idx_cat, obj_size)
    if len(obj_size) == 0:
        return np.zeros((0, 3)), np.zeros(0)
    with _session.locked_group() as grp:
        grp.set_snapshot(xyz, None, masses, idx_cat, obj_size)
        cen = _centres(grp, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, None)
        m = grp.rows(lambda ctx, objs: ctx.obj_masses())
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
    r200s = _lib.f64(np.asarray(r200s))
    edges = sph_bin_edges(r_over_r200)
    with _session.locked_group() as grp:
        grp.set_snapshot(xyz, None, masses, idx_cat, obj_size)
        cen = _centres(grp, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, centers)
        dens, cnt = grp.rows(lambda ctx, objs: ctx.dens_sph(_rows(cen, objs), _rows(r200s, objs), edges, L_BOX))
        timing = grp.timings()
    last_info.clear()
    last_info.update(counts=cnt, timing=timing)
    return dens


device_interp = True               # interpolate the shape profiles on the device (cpg_shape_interp) ...
MAX_DEVICE_INTERP_SHELLS = 256     # ... for profiles of up to this many shells; beyond that the NumPy twin below


def _interp_rows(X, Y, Xnew):
    """Row-wise scipy.interpolate.interp1d(X[p], Y[p], bounds_error=False, fill_value='extrapolate')(Xnew[p]),
    vectorised over the rows p.  Operation for operation what scipy 1.x does for kind='linear' with
    extrapolation (interp1d.__init__: stable argsort of x, NaNs last; interp1d._call_linear: searchsorted,
    clip to [1, len-1], two-term de Boor formula), so the result is bit-identical to the per-object loop of the
    reference (dens_profs_algos.pyx:169-195) -- tests/test_interp_rows.py checks that against scipy itself,
    including NaN abscissae from unconverged shells.  Y may carry trailing axes."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    Xnew = np.asarray(Xnew, dtype=np.float64)
    n, L = X.shape
    ind = np.argsort(X, axis=1, kind="mergesort")
    Xs = np.take_along_axis(X, ind, axis=1)
    Ys = np.take_along_axis(Y, ind.reshape(ind.shape + (1,) * (Y.ndim - 2)), axis=1)
    # searchsorted(side='left') per row: number of entries strictly below x_new; NaN entries sort last and
    # compare False, a NaN query lands behind every finite entry
    idx = (Xs[:, None, :] < Xnew[:, :, None]).sum(axis=2)
    nan_q = np.isnan(Xnew)
    if nan_q.any():
        idx = np.where(nan_q, (~np.isnan(Xs)).sum(axis=1)[:, None], idx)
    idx = idx.clip(1, L - 1).astype(int)
    lo, hi = idx - 1, idx
    x_lo = np.take_along_axis(Xs, lo, axis=1)
    x_hi = np.take_along_axis(Xs, hi, axis=1)
    ex = (1,) * (Y.ndim - 2)
    y_lo = np.take_along_axis(Ys, lo.reshape(lo.shape + ex), axis=1)
    y_hi = np.take_along_axis(Ys, hi.reshape(hi.shape + ex), axis=1)
    w_hi = ((Xnew - x_lo) / (x_hi - x_lo)).reshape(lo.shape + ex)
    w_lo = ((x_hi - Xnew) / (x_hi - x_lo)).reshape(lo.shape + ex)
    return w_hi * y_hi + w_lo * y_lo


def interpolate_shapes(r200s, r_over_r200, a, b, c, major, inter, minor):
    """Host half of calcDensProfsEllDirectBinning (dens_profs_algos.pyx:163-195): a, b, c onto the bin edges,
    eigenvectors onto the bin centres, abscissae r_vec = a[p]; SURVEY.md section 8(f) next-4 (the reference loops
    over objects calling scipy twelve times each: 3.1 s for 10^4 objects, against 22 ms for the binning kernel)."""
    r = _lib.f64(np.asarray(r_over_r200))
    r200s = _lib.f64(np.asarray(r200s))
    edges = ell_bin_edges(r)
    x_edges = edges[None, :] * r200s[:, None]
    x_cent = r[None, :] * r200s[:, None]
    ai = _interp_rows(a, a, x_edges)
    bi = _interp_rows(a, b, x_edges)
    ci = _interp_rows(a, c, x_edges)
    mj = _interp_rows(a, major, x_cent)
    it = _interp_rows(a, inter, x_cent)
    mn = _interp_rows(a, minor, x_cent)
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
    r = _lib.f64(np.asarray(r_over_r200))
    on_device = device_interp and 2 <= a.shape[1] <= MAX_DEVICE_INTERP_SHELLS
    if not on_device:
        prof = interpolate_shapes(r200s, r_over_r200, a, b, c, major, inter, minor)

    def shard(ctx, objs):
        if on_device:
            # the interpolation of :163-195 on the device; the profiles never come back to the host
            ctx.shape_interp(_rows(r200s, objs), r, ell_bin_edges(r), *[_rows(v, objs) for v in (a, b, c, major, inter, minor)])
            return ctx.dens_ell(_rows(cen, objs), None, None, None, None, None, None, L_BOX)
        return ctx.dens_ell(_rows(cen, objs), *[_rows(v, objs) for v in prof], L_BOX)

    with _session.locked_group() as grp:
        grp.set_snapshot(xyz, None, masses, idx_cat, obj_size)
        cen = _centres(grp, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, centers)
        dens, cnt = grp.rows(shard)
        timing = grp.timings()
    last_info.clear()
    last_info.update(counts=cnt, timing=timing)
    return dens


def calcDensProfsKernelBased(xyz, masses, r200s, r_over_r200, idx_cat, obj_size, L_BOX, CENTER, centers=None):
    xyz, masses, idx_cat, obj_size = _prep(xyz, masses, idx_cat, obj_size)
    R = len(r_over_r200)
    if len(obj_size) == 0:
        return np.zeros((0, R), dtype=np.float64)
    r200s = _lib.f64(np.asarray(r200s))
    with _session.locked_group() as grp:
        grp.set_snapshot(xyz, None, masses, idx_cat, obj_size)
        cen = _centres(grp, xyz, masses, idx_cat, obj_size, L_BOX, CENTER, centers)
        dens = grp.rows(lambda ctx, objs: ctx.dens_kernel(_rows(cen, objs), _rows(r200s, objs), r_over_r200, L_BOX))
        timing = grp.timings()
    last_info.clear()
    last_info.update(timing=timing)
    return dens
