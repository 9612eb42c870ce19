"""ctypes front-end of the CPU restatement (oracle/cp_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Mirrors the eight reference entry points (same positional arguments, same return tuples) so that
parity tests read like calls into the reference:
    cosmic_profiles/shape_profs/shape_profs_algos.pyx:517,631,739,859
    cosmic_profiles/dens_profs/dens_profs_algos.pyx:15,54,112,221
plus ``*_ints`` variants that also return the per halo-shell iteration / particle counts.
Centre finding restates cosmic_profiles/common/python_routines.py:287-306,499-565 in numpy.
"""
import ctypes
import os
import subprocess

import numpy as np
from numpy.random import default_rng

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "_build", "libcp_oracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "cp_oracle.c")):
            build()
        _LIB = ctypes.CDLL(so)
        for f in ("cpo_morph", "cpo_dens_sph", "cpo_dens_ell", "cpo_dens_kernel", "cpo_max_threads"):
            getattr(_LIB, f).restype = ctypes.c_int
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _offsets(obj_size):
    return np.concatenate(([0], np.cumsum(np.asarray(obj_size, dtype=np.int64)))).astype(np.int64)


# ---- centres: python_routines.py ------------------------------------------------------------

def respectPBCNoRef(xyz, L_BOX=None):
    """python_routines.py:499-528 (translation relative to the mean of <=50 seeded random members)."""
    if L_BOX != 0.0:
        out = xyz.copy()
        rng = default_rng(seed=0)
        choose = rng.choice(np.arange(len(xyz)), (min(50, len(xyz)),), replace=False)
        ref = np.average(out[choose], axis=0)
        for ax in range(3):
            dist = out[:, ax] - ref[ax]
            hi = dist > L_BOX / 2
            lo = dist < -L_BOX / 2
            out[:, ax][hi] = out[:, ax][hi] - L_BOX
            out[:, ax][lo] = out[:, ax][lo] + L_BOX
        return out
    return xyz


def calcCoM(xyz, masses):
    """python_routines.py:530-547."""
    rng = default_rng(seed=0)
    choose = rng.choice(np.arange(len(xyz)), (min(50, len(xyz)),), replace=False)
    ref = np.average(xyz[choose], axis=0)
    delta = xyz.copy() - ref
    mass_total = np.sum(masses)
    com = np.sum(np.reshape(masses, (len(masses), 1)) * delta / mass_total, axis=0)
    return com + ref


def calcMode(xyz, masses, rad):
    """python_routines.py:287-306 (shrinking sphere, radius x0.83 until < 5 particles remain)."""
    while True:
        com = np.sum(xyz * np.reshape(masses, (masses.shape[0], 1)), axis=0) / masses.sum()
        dist = np.linalg.norm(xyz - com, axis=1)
        keep = dist < rad
        xyz_c = xyz[keep]
        if xyz_c.shape[0] < 5:
            return com
        xyz, masses, rad = xyz_c, masses[keep], rad * 0.83


def recentreObject(xyz, L_BOX):
    """python_routines.py:549-565."""
    out = xyz.copy()
    for ax in range(3):
        hi = out[:, ax] >= L_BOX
        out[:, ax][hi] = out[:, ax][hi] - L_BOX
        lo = out[:, ax] < 0.0
        out[:, ax][lo] = out[:, ax][lo] + L_BOX
    return out


def centres(xyz, masses, idx_cat, obj_size, L_BOX, CENTER, shape_path):
    """Serial centre loop of the drivers (shape_profs_algos.pyx:586-591, dens_profs_algos.pyx:96-101).
    shape_path: 'mode' starts from the bounding-box extent; density paths start from rad=1000."""
    off = _offsets(obj_size)
    cen = np.zeros((len(obj_size), 3))
    for p in range(len(obj_size)):
        sel = idx_cat[off[p]:off[p + 1]]
        x = respectPBCNoRef(xyz[sel], L_BOX)
        if CENTER == 'mode':
            rad = max(x[:, 0].max() - x[:, 0].min(), x[:, 1].max() - x[:, 1].min(),
                      x[:, 2].max() - x[:, 2].min()) if shape_path else 1000
            cen[p] = calcMode(x, masses[sel], rad)
        else:
            cen[p] = calcCoM(x, masses[sel])
    return cen


# ---- shapes --------------------------------------------------------------------------------

def _nan0(a):
    a = a.copy()
    a[a == 0.0] = np.nan
    return a


def morph(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN, CENTER,
          reduced, shell_based, local, SAFE=0.0, nthreads=0, centers=None):
    xyz = _f64(xyz)
    masses = _f64(masses)
    vx = None if vxyz is None else _f64(vxyz)
    r200 = _f64(r200)
    idx = np.ascontiguousarray(idx_cat, dtype=np.int32)
    off = _offsets(obj_size)
    n = len(obj_size)
    rr = _f64(r_over_r200) if local else np.zeros(1)
    R = len(rr) if local else 1
    cen = centres(xyz, masses, idx, obj_size, L_BOX, CENTER, True) if centers is None else _f64(centers)
    out = np.zeros((n, R, 12))
    nit = np.zeros((n, R), dtype=np.int32)
    npt = np.zeros((n, R), dtype=np.int32)
    m = np.zeros(n)
    vc = np.zeros((n, 3))
    D, I = ctypes.c_double, ctypes.c_int32
    rc = lib().cpo_morph(_p(xyz, D), _p(vx, D), _p(masses, D), _p(idx, I), _p(off, ctypes.c_int64), n,
                         _p(cen, D), _p(r200, D), _p(rr, D), R, int(bool(local)), D(SAFE),
                         ctypes.c_float(L_BOX), D(IT_TOL), int(IT_WALL), int(IT_MIN), int(bool(reduced)),
                         int(bool(shell_based)), _p(out, D), _p(nit, I), _p(npt, I), _p(m, D), _p(vc, D),
                         int(nthreads))
    assert rc == 0
    d = _nan0(out[:, :, 0])
    q = _nan0(out[:, :, 1])
    s = _nan0(out[:, :, 2])
    major = _nan0(out[:, :, 3:6])
    inter = _nan0(out[:, :, 6:9])
    minor = _nan0(out[:, :, 9:12])
    cen_out = recentreObject(cen, L_BOX)
    if not local:
        d, q, s = d[:, 0], q[:, 0], s[:, 0]
        major, inter, minor = major[:, 0], inter[:, 0], minor[:, 0]
        nit, npt = nit[:, 0], npt[:, 0]
    return (d, q, s, minor, inter, major, cen_out, m), (nit, npt, vc, out)


def calcMorphLocal(xyz, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN, CENTER,
                   reduced, shell_based, **kw):
    return morph(xyz, None, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN,
                 CENTER, reduced, shell_based, True, **kw)[0]


def calcMorphGlobal(xyz, masses, r200, idx_cat, obj_size, L_BOX, IT_TOL, IT_WALL, IT_MIN, CENTER, SAFE, reduced,
                    **kw):
    return morph(xyz, None, masses, r200, idx_cat, obj_size, L_BOX, None, IT_TOL, IT_WALL, IT_MIN, CENTER,
                 reduced, False, False, SAFE=SAFE, **kw)[0]


def calcMorphLocalVelDisp(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL,
                          IT_MIN, CENTER, reduced, shell_based, **kw):
    return morph(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, r_over_r200, IT_TOL, IT_WALL, IT_MIN,
                 CENTER, reduced, shell_based, True, **kw)[0]


def calcMorphGlobalVelDisp(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, IT_TOL, IT_WALL, IT_MIN, CENTER,
                           SAFE, reduced, **kw):
    return morph(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, None, IT_TOL, IT_WALL, IT_MIN, CENTER,
                 reduced, False, False, SAFE=SAFE, **kw)[0]


# ---- densities -----------------------------------------------------------------------------

def calcMassesCenters(xyz, masses, idx_cat, obj_size, L_BOX, CENTER):
    """dens_profs_algos.pyx:15-51."""
    xyz = _f64(xyz)
    masses = _f64(masses)
    off = _offsets(obj_size)
    cen = centres(xyz, masses, idx_cat, obj_size, L_BOX, CENTER, False)
    m = np.zeros(len(obj_size))
    for p in range(len(obj_size)):
        acc = 0.0
        for v in masses[idx_cat[off[p]:off[p + 1]]]:
            acc = acc + v
        m[p] = acc
    return recentreObject(cen, L_BOX), m


def sph_bin_edges(r_over_r200):
    """dens_profs_algos.pyx:83."""
    r = _f64(r_over_r200)
    return np.hstack(([np.float64(1e-8), (r[:-1] + r[1:]) / 2., r[-1]]))


def ell_bin_edges(r_over_r200):
    """dens_profs_algos.pyx:155-156."""
    r = _f64(r_over_r200)
    mid = (r[:-1] + r[1:]) / 2
    return np.hstack(([np.float64(mid[0] * mid[0] / mid[1]), mid, r[-1]]))


def calcDensProfsSphDirectBinning(xyz, masses, r200s, r_over_r200, idx_cat, obj_size, L_BOX, CENTER, nthreads=0,
                                  centers=None, return_counts=False):
    xyz = _f64(xyz)
    masses = _f64(masses)
    r200s = _f64(r200s)
    idx = np.ascontiguousarray(idx_cat, dtype=np.int32)
    n, R = len(obj_size), len(r_over_r200)
    if n == 0:
        return np.zeros((0, R))
    off = _offsets(obj_size)
    edges = sph_bin_edges(r_over_r200)
    cen = centres(xyz, masses, idx, obj_size, L_BOX, CENTER, False) if centers is None else _f64(centers)
    dens = np.zeros((n, R))
    cnt = np.zeros((n, R), dtype=np.int32)
    D, I = ctypes.c_double, ctypes.c_int32
    rc = lib().cpo_dens_sph(_p(xyz, D), _p(masses, D), _p(idx, I), _p(off, ctypes.c_int64), n, _p(cen, D),
                            _p(r200s, D), _p(edges, D), R, ctypes.c_float(L_BOX), _p(dens, D), _p(cnt, I),
                            int(nthreads))
    assert rc == 0
    return (dens, cnt) if return_counts else dens


def interpolate_shapes(r200s, r_over_r200, a, b, c, major, inter, minor):
    """Host half of calcDensProfsEllDirectBinning (dens_profs_algos.pyx:163-195): scipy interp1d with
    extrapolation of a, b, c onto bin edges and of the eigenvectors onto bin centres."""
    from scipy.interpolate import interp1d
    r = _f64(r_over_r200)
    edges = ell_bin_edges(r)
    n, R = a.shape[0], len(r)
    ai = np.zeros((n, R + 1))
    bi = np.zeros((n, R + 1))
    ci = np.zeros((n, R + 1))
    mj = np.zeros((n, R, 3))
    it = np.zeros((n, R, 3))
    mn = np.zeros((n, R, 3))
    for p in range(n):
        r_vec = np.float64(a[p])
        kw = dict(bounds_error=False, fill_value='extrapolate')
        ai[p] = interp1d(r_vec, a[p], **kw)(edges * r200s[p])
        bi[p] = interp1d(r_vec, b[p], **kw)(edges * r200s[p])
        ci[p] = interp1d(r_vec, c[p], **kw)(edges * r200s[p])
        for k in range(3):
            mj[p, :, k] = interp1d(r_vec, major[p, :, k], **kw)(r * r200s[p])
            it[p, :, k] = interp1d(r_vec, inter[p, :, k], **kw)(r * r200s[p])
            mn[p, :, k] = interp1d(r_vec, minor[p, :, k], **kw)(r * r200s[p])
    return ai, bi, ci, mj, it, mn


def calcDensProfsEllDirectBinning(xyz, masses, r200s, r_over_r200, a, b, c, major, inter, minor, idx_cat,
                                  obj_size, L_BOX, CENTER, nthreads=0, centers=None, return_counts=False):
    xyz = _f64(xyz)
    masses = _f64(masses)
    r200s = _f64(r200s)
    idx = np.ascontiguousarray(idx_cat, dtype=np.int32)
    n, R = len(obj_size), len(r_over_r200)
    if n == 0:
        return np.zeros((0, R))
    off = _offsets(obj_size)
    ai, bi, ci, mj, it, mn = interpolate_shapes(r200s, r_over_r200, _f64(a), _f64(b), _f64(c), _f64(major),
                                                _f64(inter), _f64(minor))
    cen = centres(xyz, masses, idx, obj_size, L_BOX, CENTER, False) if centers is None else _f64(centers)
    dens = np.zeros((n, R))
    cnt = np.zeros((n, R), dtype=np.int32)
    D, I = ctypes.c_double, ctypes.c_int32
    rc = lib().cpo_dens_ell(_p(xyz, D), _p(masses, D), _p(idx, I), _p(off, ctypes.c_int64), n, _p(cen, D),
                            _p(ai, D), _p(bi, D), _p(ci, D), _p(mj, D), _p(it, D), _p(mn, D), R,
                            ctypes.c_float(L_BOX), _p(dens, D), _p(cnt, I), int(nthreads))
    assert rc == 0
    return (dens, cnt) if return_counts else dens


def calcDensProfsKernelBased(xyz, masses, r200s, r_over_r200, idx_cat, obj_size, L_BOX, CENTER, nthreads=0,
                             centers=None):
    xyz = _f64(xyz)
    masses = _f64(masses)
    r200s = _f64(r200s)
    rr = _f64(r_over_r200)
    idx = np.ascontiguousarray(idx_cat, dtype=np.int32)
    n, R = len(obj_size), len(rr)
    if n == 0:
        return np.zeros((0, R))
    off = _offsets(obj_size)
    cen = centres(xyz, masses, idx, obj_size, L_BOX, CENTER, False) if centers is None else _f64(centers)
    dens = np.zeros((n, R))
    D, I = ctypes.c_double, ctypes.c_int32
    rc = lib().cpo_dens_kernel(_p(xyz, D), _p(masses, D), _p(idx, I), _p(off, ctypes.c_int64), n, _p(cen, D),
                               _p(r200s, D), _p(rr, D), R, ctypes.c_float(L_BOX), _p(dens, D), int(nthreads))
    assert rc == 0
    return dens


# ---- Gadget / FoF ingest ---------------------------------------------------

def calcObjCat(nb_shs, sh_len, fof_sizes, group_r200, MIN_NUMBER_PTCS):
    """gadget/gen_catalogues.pyx:39-95 restated in numpy (pinned against the compiled reference by
    tests/golden/gadget_golden.npz): group p passes when nb_shs[p] > 0 and the length of its first subhalo,
    sh_len[sum(nb_shs[:p])], is >= MIN_NUMBER_PTCS (:66-73); its object is the index range starting at
    sum(fof_sizes[:p]) (:80-85, calcCSHIdxs :29-35).  Returns (obj_cat, obj_r200, obj_size) as :91."""
    nb_shs = np.asarray(nb_shs, dtype=np.int64)
    sh_len = np.asarray(sh_len, dtype=np.int64)
    fof_sizes = np.asarray(fof_sizes, dtype=np.int64)
    if len(nb_shs) == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")   # np.max at :76
    sh_off = np.concatenate(([0], np.cumsum(nb_shs)[:-1]))
    start = np.concatenate(([0], np.cumsum(fof_sizes)[:-1]))
    first_len = np.where(nb_shs > 0, sh_len[np.minimum(sh_off, max(len(sh_len) - 1, 0))] if len(sh_len) else 0, 0)
    ok = (nb_shs > 0) & (first_len >= MIN_NUMBER_PTCS)
    size = first_len[ok]
    parts = [np.arange(s0, s0 + n, dtype=np.int64) for s0, n in zip(start[ok], size)]
    idx = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
    return idx.astype(np.int32), np.asarray(group_r200, dtype=np.float64)[ok], size.astype(np.int32)


def read_field_arith(block, dataset, unit_fac, time=1.0, mass_table=0.0, n_part=0):
    """The arithmetic of readgadget.read_field + read_block on one dataset (gadget/readgadget.py:106-120, :148-169),
    restated with the very NumPy expressions of the reference (h5py, which only supplies `dataset`, is absent here,
    so this restatement is the pin: "parity unpinned" against a running reference for this function).
    block in {"POS ", "VEL ", "MASS"}; dataset None (MASS only) = no Masses dataset, MassTable constant."""
    if dataset is None:
        if mass_table == 0.0:
            raise Exception("Problem reading the block %s" % block)
        array = np.ones(n_part, np.float64) * mass_table / unit_fac          # :108
    else:
        array = dataset[:] / unit_fac                                           # :112
    if block == "VEL ":
        array *= np.sqrt(time)                                                  # :116
    out = np.zeros(array.shape, dtype=np.float64)                              # :164 (float64 destination)
    out[:] = array                                                              # :169
    return out
