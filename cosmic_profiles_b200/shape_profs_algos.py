"""Drop-in replacements for the four shape drivers of the reference, same positional signatures and return
tuples:  calcMorphLocal / calcMorphGlobal / calcMorphLocalVelDisp / calcMorphGlobalVelDisp
(cosmic_profiles/shape_profs/shape_profs_algos.pyx:517, :631, :739, :859).

Everything numeric runs in libcosmicgpu.so (sm_100a); this module only does what the reference does in
Python around its prange: centres (:586-591), zero -> NaN masking (:616-624), recentring (:627) and the
(d, q, s, minor, inter, major, centers, m) packing (:628).  No CPU fallback.
"""
import threading

import numpy as np

from . import _lib, centers as _centers, partition as _partition, session as _session

last_info = {}   # n_iter / n_pts / timing of the most recent call (the reference keeps these as locals)


def _nan0(a):
    a = np.array(a, dtype=np.float64, copy=True)
    a[a == 0.0] = np.nan
    return a


def _run_one(device, xyz, vxyz, masses, idx_cat, obj_size, cen, r200, rr, local, SAFE, L_BOX, IT_TOL, IT_WALL,
             IT_MIN, reduced, shell_based, CENTER="com"):
    """Returns (out, nit, npt, m, vc, centres, timing)."""
    ctx = _session.get_context(device)
    ctx.set_particles(xyz, vxyz, masses)
    ctx.set_catalogue(idx_cat, obj_size)
    if cen is None:   # centres on the device (shape path: calcMode starts from the bounding-box extent)
        ref = _centers.reference_points(xyz, idx_cat, obj_size) if L_BOX != 0.0 else None
        cen = ctx.centers(1 if CENTER == 'mode' else 0, True, L_BOX, ref)
    # out/nit/npt are views into the context's page-locked scratch, valid until its next shape() call;
    # every caller below copies out of them before touching the context again
    return ctx.shape(cen, r200, rr, local, SAFE, L_BOX, IT_TOL, IT_WALL, IT_MIN, reduced, shell_based,
                     vxyz is not None) + (cen, ctx.timing())


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
    R = int(rr.shape[0]) if local else 1
    if centers is not None:
        cen = _lib.f64(centers)
    elif _session.gpu_centers:
        cen = None          # computed on the device by every shard (cpg_centers)
    else:
        cen = _centers.compute_centers(xyz, masses, idx_cat, obj_size, L_BOX, CENTER, shape_path=True)
    devs = _session.devices or [0]
    args = (rr, local, SAFE, L_BOX, IT_TOL, IT_WALL, IT_MIN, reduced, shell_based, CENTER)
    if len(devs) == 1 or n < 2 * len(devs):
        out, nit, npt, m, vc, cen, tim = _run_one(devs[0], xyz, vxyz, masses, idx_cat, obj_size, cen, r200, *args)
        nit, npt = nit.copy(), npt.copy()
        timings = [tim]
    else:
        # objects are independent: shard them, one host thread per device, disjoint output slices
        parts = _partition.lpt_partition(obj_size.astype(np.float64) * R, len(devs))
        out = np.zeros((n, R, 12))
        nit = np.full((n, R), -1, dtype=np.int32)
        npt = np.full((n, R), -1, dtype=np.int32)
        m = np.zeros(n)
        vc = np.zeros((n, 3))
        cen_all = np.zeros((n, 3))
        timings = [None] * len(devs)
        errors = []

        def work(k):
            try:
                objs = parts[k]
                if len(objs) == 0:
                    return
                sub_idx, sub_size = _partition.sub_catalogue(idx_cat, obj_size, objs)
                # ship only this shard's particles; the shard catalogue becomes 0..n_sub-1
                local_idx = np.arange(len(sub_idx), dtype=np.int32)
                res = _run_one(devs[k], xyz[sub_idx], None if vxyz is None else vxyz[sub_idx], masses[sub_idx],
                               local_idx, sub_size, None if cen is None else cen[objs], r200[objs], *args)
                out[objs], nit[objs], npt[objs], m[objs], vc[objs], cen_all[objs] = res[:6]
                timings[k] = res[6]
            except Exception as e:   # surfaced below, in the caller's thread
                errors.append(e)

        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(devs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        cen = cen_all
    d = _nan0(out[:, :, 0])
    q = _nan0(out[:, :, 1])
    s = _nan0(out[:, :, 2])
    major = _nan0(out[:, :, 3:6])
    inter = _nan0(out[:, :, 6:9])
    minor = _nan0(out[:, :, 9:12])
    cen_out = _centers.recentre(cen, L_BOX)
    if not local:
        d, q, s = d[:, 0], q[:, 0], s[:, 0]
        major, inter, minor = major[:, 0], inter[:, 0], minor[:, 0]
        nit, npt = nit[:, 0], npt[:, 0]
    last_info.clear()
    last_info.update(n_iter=nit, n_pts=npt, vcenters=vc, timing=timings)
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
