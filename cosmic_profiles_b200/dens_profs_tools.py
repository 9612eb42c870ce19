"""Drop-in for cosmic_profiles.dens_profs.dens_profs_tools.fitDensProfHelper (dens_profs_tools.py:325-359) on the
device.

The reference fits every object's density profile with scipy.optimize.minimize(method='Powell', bounds=...) in a
pathos process pool (:347-348), the objective being a Python list comprehension over the bins (:227-230).  Here
the per-object preparation is the reference's own NumPy arithmetic, vectorised over objects -- NaN bins dropped
(:254-258), median / np.average(median) (:277), its logarithm (the constant half of every objective term) -- and the
minimisation itself (scipy 1.18.1's Powell with bounded Forsythe-Malcolm-Moler line searches, restated) runs as one
warp per object in csrc/cpg_fit.cuh.  Same signature, same (N, n) return array.
"""
import numpy as np

from . import session
from ._lib import CosmicGpuError

PROFILE_ID = {"nfw": 0, "hernquist": 1, "einasto": 2, "alpha_beta_gamma": 3}
NPARS = {"nfw": 2, "hernquist": 2, "einasto": 3, "alpha_beta_gamma": 5}
MAX_BINS = 1024
last_info = {}


def _bounds(method):
    """dens_profs_tools.py:262-318: (lb, ub) per parameter; a parameter named in `method` is pinned."""
    def b(name, lo):
        return (method[name], method[name]) if name in method else (lo, np.inf)
    prof = method["profile"]
    if prof == "einasto":
        return [(1e-7, np.inf), b("alpha", 1e-10), b("r_s", 1e-5)]
    if prof == "alpha_beta_gamma":
        return [(1e-7, np.inf), b("alpha", 1e-5), b("beta", 1e-5), b("gamma", 1e-5), b("r_s", 1e-5)]
    return [(1e-7, np.inf), b("r_s", 1e-5)]


def _initial_guess(prof, r200s):
    n = len(r200s)
    if prof == "einasto":
        x0 = np.tile(np.array([0.1, 0.18, 0.0]), (n, 1))                    # :276
    elif prof == "alpha_beta_gamma":
        x0 = np.tile(np.array([0.1, 1.0, 1.0, 1.0, 0.0]), (n, 1))           # :304
    else:
        x0 = np.tile(np.array([0.1, 0.0]), (n, 1))                          # :317
    x0[:, -1] = r200s / 5
    return x0


def prepare(dens_profs, ROverR200, r200s):
    """Per-object inputs of the objective, with the reference's NumPy operations: R_to_min = ROverR200*r200 (:253),
    NaN bins removed (:255-256), median/np.average(median) (:277) and np.log of it (:229)."""
    dens = np.ascontiguousarray(dens_profs, dtype=np.float64)
    rr = np.asarray(ROverR200, dtype=np.float64)
    r200s = np.asarray(r200s, dtype=np.float64)
    n, R = dens.shape
    keep = ~np.isnan(dens)
    n_valid = keep.sum(axis=1).astype(np.int32)
    radii = rr[None, :] * r200s[:, None]
    med = dens
    if (n_valid < R).any():
        # objects with NaN bins (empty shells): stable partition, kept bins first, for all rows at once
        order = np.argsort(~keep, axis=1, kind="stable")
        radii = np.take_along_axis(radii, order, axis=1)
        med = np.take_along_axis(dens, order, axis=1)
    avg = np.full(n, np.nan)
    with np.errstate(all="ignore"):
        for m in np.unique(n_valid):          # np.average(row[:m]) == mean over a contiguous row of m values (same pairwise sum)
            rows = np.flatnonzero(n_valid == m)
            if m > 0:
                avg[rows] = np.ascontiguousarray(med[rows, :m]).mean(axis=1)
        lmed = np.log(med / avg[:, None])
    return np.ascontiguousarray(radii), np.ascontiguousarray(lmed), n_valid, avg


def fitDensProfHelper(dens_profs, ROverR200, r200s, method):
    prof = method["profile"]
    if prof not in PROFILE_ID:
        raise KeyError(prof)
    if method.get("min_method", "Powell") != "Powell":
        raise CosmicGpuError("only the reference's default minimiser (Powell) runs on the device; got min_method=%r "
                             "(there is no CPU fallback)" % (method["min_method"],))
    dens = np.asarray(dens_profs, dtype=np.float64)
    if dens.ndim != 2 or dens.shape[1] > MAX_BINS:
        raise ValueError("dens_profs must be (N, r_res) with r_res <= %d" % MAX_BINS)
    r200s = np.asarray(r200s, dtype=np.float64)
    radii, lmed, n_valid, avg = prepare(dens, ROverR200, r200s)
    bnds = _bounds(method)
    lb = np.array([b[0] for b in bnds], dtype=np.float64)
    ub = np.array([b[1] for b in bnds], dtype=np.float64)
    x0 = _initial_guess(prof, r200s)
    # objects are independent (the reference maps them over a process pool, :347-348): equal slices per device
    with session.locked_group() as grp:
        ctxs = grp.contexts
        n = dens.shape[0]
        cuts = [n * k // len(ctxs) for k in range(len(ctxs) + 1)] if n >= 2 * len(ctxs) else [0] + [n] * len(ctxs)
        parts, errs = [None] * len(ctxs), []

        def work(k):
            a, b = cuts[k], cuts[k + 1]
            if b > a:
                try:
                    parts[k] = ctxs[k].fit_profiles(radii[a:b], lmed[a:b], n_valid[a:b], avg[a:b], x0[a:b], lb, ub,
                                                    PROFILE_ID[prof]) + (ctxs[k].timing(),)
                except Exception as e:   # re-raised below, in the caller's thread
                    errs.append(e)
        import threading
        th = [threading.Thread(target=work, args=(k,)) for k in range(1, len(ctxs))]
        for t in th:
            t.start()
        work(0)
        for t in th:
            t.join()
        if errs:
            raise errs[0]
    parts = [p for p in parts if p is not None]
    if parts:
        best, fun, nfev, nit = (np.concatenate([p[j] for p in parts]) for j in range(4))
    else:
        best, fun = np.zeros((0, NPARS[prof])), np.zeros(0)
        nfev = nit = np.zeros(0, dtype=np.int32)
    last_info.clear()
    last_info.update(fun=fun, nfev=nfev, nit=nit, timing=parts[0][4] if parts else None, timings=[p[4] for p in parts])
    return best
