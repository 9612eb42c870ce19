"""Host-side centre finding, as the reference does it before entering the hot path.

The reference computes object centres in a serial Python loop ahead of every driver
(cosmic_profiles/shape_profs/shape_profs_algos.pyx:586-591, dens_profs/dens_profs_algos.pyx:39-44,96-101)
with the helpers of cosmic_profiles/common/python_routines.py (calcMode :287-306, respectPBCNoRef :499-528,
calcCoM :530-547, recentreObject :549-565).  Centres are an INPUT of the GPU path (BASELINE.json
north_star).  By default the drop-in drivers compute them on the device (cpg_centers, csrc/cpg_centers.cuh; agreement
with the reference's centres <= 1e-12, tests/test_gpu_scale_parity.py); this module is the host alternative
(`session.gpu_centers = False`): the same numpy operations as the reference, bit-identical centres.  It also supplies
the seeded reference points the device centre finder needs in a periodic box, and recentreObject.
"""
import numpy as np
from numpy.random import default_rng


def _reference_points(n):
    rng = default_rng(seed=0)
    # == rng.choice(np.arange(n), ...) of python_routines.py:512-513, without materialising arange(n)
    return rng.choice(n, (min(50, n),), replace=False)


_choice_cache = {}


def _choices(n):
    choose = _choice_cache.get(n)
    if choose is None:
        choose = _reference_points(n)
        if len(_choice_cache) < 100000:
            _choice_cache[n] = choose
    return choose


def reference_points(xyz, idx_cat, obj_size):
    """The reference point respectPBCNoRef / calcCoM average from <= 50 seeded random members of every object
    (python_routines.py:512-514, :540-542).  Needed by the device centre finder when L_BOX != 0, where the
    reference point decides which particles are translated by +-L_BOX.  Objects of >= 50 members are handled in
    one vectorised gather + mean (same summation order as np.average over the 50 rows of one object)."""
    obj_size = np.asarray(obj_size, dtype=np.int64)
    off = np.concatenate(([0], np.cumsum(obj_size)))
    ref = np.zeros((len(obj_size), 3), dtype=np.float64)
    big = np.flatnonzero(obj_size >= 50)
    if len(big):
        pick = np.empty((len(big), 50), dtype=np.int64)
        for k, p in enumerate(big):
            pick[k] = _choices(int(obj_size[p]))
        members = np.asarray(idx_cat)[pick + off[big][:, None]]
        ref[big] = np.average(xyz[members], axis=1)
    for p in np.flatnonzero((obj_size > 0) & (obj_size < 50)):
        n = int(obj_size[p])
        ref[p] = np.average(xyz[idx_cat[off[p]:off[p + 1]][_choices(n)]], axis=0)
    return ref


def respect_pbc_noref(xyz, L_BOX):
    if L_BOX == 0.0:
        return xyz
    out = xyz.copy()
    ref = np.average(out[_reference_points(len(xyz))], axis=0)
    for ax in range(3):
        dist = out[:, ax] - ref[ax]
        far = dist > L_BOX / 2
        near = dist < -L_BOX / 2
        out[:, ax][far] = out[:, ax][far] - L_BOX
        out[:, ax][near] = out[:, ax][near] + L_BOX
    return out


def calc_com(xyz, masses):
    ref = np.average(xyz[_reference_points(len(xyz))], axis=0)
    delta = xyz.copy() - ref
    mass_total = np.sum(masses)
    com = np.sum(np.reshape(masses, (len(masses), 1)) * delta / mass_total, axis=0)
    return com + ref


def calc_mode(xyz, masses, rad):
    while True:
        com = np.sum(xyz * np.reshape(masses, (masses.shape[0], 1)), axis=0) / masses.sum()
        keep = np.linalg.norm(xyz - com, axis=1) < rad
        if np.count_nonzero(keep) < 5:
            return com
        xyz, masses, rad = xyz[keep], masses[keep], rad * 0.83


def recentre(centers, L_BOX):
    out = centers.copy()
    for ax in range(3):
        hi = out[:, ax] >= L_BOX
        out[:, ax][hi] = out[:, ax][hi] - L_BOX
        lo = out[:, ax] < 0.0
        out[:, ax][lo] = out[:, ax][lo] + L_BOX
    return out


def compute_centers(xyz, masses, idx_cat, obj_size, L_BOX, CENTER, shape_path):
    """shape_path=True: 'mode' starts from the object's bounding-box extent (shape_profs_algos.pyx:589);
    False: from rad=1000 (dens_profs_algos.pyx:42)."""
    off = np.concatenate(([0], np.cumsum(np.asarray(obj_size, dtype=np.int64))))
    cen = np.zeros((len(obj_size), 3), dtype=np.float64)
    for p in range(len(obj_size)):
        sel = idx_cat[off[p]:off[p + 1]]
        x = respect_pbc_noref(xyz[sel], L_BOX)
        if CENTER == 'mode':
            if shape_path:
                rad = max(x[:, 0].max() - x[:, 0].min(), x[:, 1].max() - x[:, 1].min(), x[:, 2].max() - x[:, 2].min())
            else:
                rad = 1000
            cen[p] = calc_mode(x, masses[sel], rad)
        else:
            cen[p] = calc_com(x, masses[sel])
    return cen
