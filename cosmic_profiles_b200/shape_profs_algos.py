"""Drop-in replacements for the four shape drivers of the reference, same positional signatures and return
tuples:  calcMorphLocal / calcMorphGlobal / calcMorphLocalVelDisp / calcMorphGlobalVelDisp
(cosmic_profiles/shape_profs/shape_profs_algos.pyx:517, :631, :739, :859).

Everything numeric runs in libcosmicgpu.so (sm_100a); this module only does what the reference does in
Python around its prange: centres (:586-591), zero -> NaN masking (:616-624), recentring (:627) and the
(d, q, s, minor, inter, major, centers, m) packing (:628).  No CPU fallback.
"""
import numpy as np

from . import _lib, centers as _centers, session as _session

last_info = {}   # n_iter / n_pts / timing of the most recent call (the reference keeps these as locals)


def _morph(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN, CENTER, SAFE,
           reduced, shell_based, local, centers=None):
    xyz = _lib.f64(np.asarray(xyz))
    masses = _lib.f64(np.asarray(masses))
    vxyz = None if vxyz is None else _lib.f64(np.asarray(vxyz))
    r200 = _lib.f64(np.asarray(r200))
    idx_cat = _lib.i32(np.asarray(idx_cat))
    obj_size = _lib.i32(np.asarray(obj_size))
    rr = _lib.f64(np.asarray(r_over_r200)) if local else None
    n = int(obj_size.shape[0])
    if centers is not None:
        cen = _lib.f64(centers)
    elif _session.gpu_centers:
        cen = None          # computed on the device(s) that hold the objects (cpg_centers)
    else:
        cen = _centers.compute_centers(xyz, masses, idx_cat, obj_size, L_BOX, CENTER, shape_path=True)
    mode = -1 if cen is not None else (1 if CENTER == 'mode' else 0)
    # the reference point of python_routines.respectPBCNoRef decides which particles move in a periodic box
    ref = _centers.reference_points(xyz, idx_cat, obj_size) if (cen is None and L_BOX != 0.0) else None
    args = (rr, local, SAFE, L_BOX, IT_TOL, IT_WALL, IT_MIN, reduced, shell_based, vxyz is not None)
    # objects are independent (shape_profs_algos.pyx:592: prange over objects): with several devices in
    # session.devices the library deals them out by particle count (cpg_multi); one lock for the whole sequence
    with _session.locked_group() as grp:
        if grp.multi is not None and n >= 2 * len(grp.devices):
            grp.set_snapshot(xyz, vxyz, masses, idx_cat, obj_size)
            out, nit, npt, m, vc, cen = grp.multi.shape(cen, mode, ref, r200, *args)
            timings = grp.timings()
            extra = grp.multi.info()
        else:   # one device (or too few objects to shard: everything on the first one)
            ctx = grp.contexts[0]
            ctx.set_snapshot(xyz, vxyz, masses, idx_cat, obj_size)
            if cen is None:   # shape path: calcMode starts from the bounding-box extent
                cen = ctx.centers(mode, True, L_BOX, ref)
            # out / nit / npt: views of the context's page-locked scratch, consumed below while the lock is held
            out, nit, npt, m, vc = ctx.shape(cen, r200, *args, zero_copy=True)
            nit, npt = nit.copy(), npt.copy()
            timings = [ctx.timing()]
            extra = None
        # zeros -> NaN (:616-624) and the split into the six returned arrays, on a few host threads
        d, q, s, major, inter, minor = _lib.unpack_shape(out)
    cen_out = _centers.recentre(cen, L_BOX)
    if not local:
        d, q, s = d[:, 0], q[:, 0], s[:, 0]
        major, inter, minor = major[:, 0], inter[:, 0], minor[:, 0]
        nit, npt = nit[:, 0], npt[:, 0]
    last_info.clear()
    last_info.update(n_iter=nit, n_pts=npt, vcenters=vc, timing=timings, multi=extra)
    return d, q, s, minor, inter, major, cen_out, m


def calcMorphLocal(xyz, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN, CENTER, reduced,
                   shell_based, centers=None):
    return _morph(xyz, None, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN, CENTER,
                  0.0, reduced, shell_based, True, centers)


def calcMorphGlobal(xyz, masses, r200, idx_cat, obj_size, L_BOX, IT_TOL, IT_WALL, IT_MIN, CENTER, SAFE, reduced,
                    centers=None):
    return _morph(xyz, None, masses, r200, idx_cat, obj_size, L_BOX, None, IT_TOL, IT_WALL, IT_MIN, CENTER, SAFE,
                  reduced, False, False, centers)


def calcMorphLocalVelDisp(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN,
                          CENTER, reduced, shell_based, centers=None):
    return _morph(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN, CENTER,
                  0.0, reduced, shell_based, True, centers)


def calcMorphGlobalVelDisp(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, IT_TOL, IT_WALL, IT_MIN, CENTER, SAFE,
                           reduced, centers=None):
    return _morph(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, None, IT_TOL, IT_WALL, IT_MIN, CENTER, SAFE,
                  reduced, False, False, centers)
