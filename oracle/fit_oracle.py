"""CPU restatement of the density-profile fit -- TEST INFRASTRUCTURE ONLY.

Restates, in plain Python floats,
  * fitDensProf / toMinimize and the four profile models of the reference
    (cosmic_profiles/dens_profs/dens_profs_tools.py:22-79, :227-323), and
  * the algorithm the reference delegates to, which is NOT under /root/reference: scipy 1.18.1
    ``scipy.optimize.minimize(method='Powell', bounds=...)`` = ``_minimize_powell`` with ``_linesearch_powell``,
    ``_line_for_search`` and ``_minimize_scalar_bounded`` (scipy/optimize/_optimize.py; Powell's conjugate-direction
    method with Forsythe-Malcolm-Moler ``fmin`` line searches, infinite one-sided bounds mapped through arctan/tan).
    With the reference's bounds every parameter has a finite lower bound, so the unbounded (Brent) branch of
    ``_linesearch_powell`` is unreachable and is not restated.

Pinned by tests/test_fit_oracle.py against the reference's own fitDensProf running the installed scipy
(same iterates, same number of function evaluations) and by tests/golden/fit_golden.npz.
The CUDA kernel (cosmic_profiles_b200/csrc/cpg_fit.cuh) follows this file statement by statement.
"""
import math

import numpy as np

INF = float("inf")
PROFILES = ("nfw", "hernquist", "einasto", "alpha_beta_gamma")
NPARS = {"nfw": 2, "hernquist": 2, "einasto": 3, "alpha_beta_gamma": 5}


def pairwise_sum(a):
    """numpy's float64 add.reduce over a contiguous 1-D array (pairwise_sum in numpy/_core/src/umath/loops_utils.h.src):
    n < 8 sequential; n <= 128 (PW_BLOCKSIZE) 8 accumulators, a fixed combine tree and a sequential remainder; longer
    arrays are split at n2 = n / 2 rounded down to a multiple of 8 and the two halves summed recursively."""
    n = len(a)
    if n > 128:
        n2 = n // 2
        n2 -= n2 % 8
        return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])
    if n < 8:
        res = 0.0 if n == 0 else a[0]     # reduce starts from the first element ... (-0.0 + x identities aside)
        for i in range(1, n):
            res += a[i]
        return res
    r = [a[j] for j in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] += a[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res += a[i]
        i += 1
    return res


def model(profile, r, x):
    """dens_profs_tools.py:22-79 on np.float64 scalars (so that ** and exp are NumPy's, as in the reference)."""
    r = np.float64(r)
    x = [np.float64(v) for v in x]
    if profile == "nfw":            # rho_s/((r/r_s)*(1+r/r_s)**2)
        rho_s, r_s = x
        return rho_s / ((r / r_s) * (1 + r / r_s) ** 2)
    if profile == "hernquist":      # rho_s/((r/r_s)*(1+r/r_s)**3)
        rho_s, r_s = x
        return rho_s / ((r / r_s) * (1 + r / r_s) ** 3)
    if profile == "einasto":        # rho_s*np.exp(-2/alpha*((r/r_s)**alpha-1))
        rho_s, alpha, r_s = x
        return rho_s * np.exp(-2 / alpha * ((r / r_s) ** alpha - 1))
    rho_s, alpha, beta, gamma, r_s = x   # rho_s/((r/r_s)**gamma*(1+(r/r_s)**alpha)**((beta-gamma)/alpha))
    return rho_s / ((r / r_s) ** gamma * (1 + (r / r_s) ** alpha) ** ((beta - gamma) / alpha))


def _log(v):
    with np.errstate(all="ignore"):
        return float(np.log(np.float64(v)))   # NumPy's own (SIMD) log, not libm's: they differ in the last bit


class _MaxFun(Exception):
    pass


class Objective:
    """toMinimize (dens_profs_tools.py:227-230) with the call counter of scipy's
    _wrap_scalar_function_maxfun_validation."""

    def __init__(self, profile, lmed, radii, maxfun):
        self.profile, self.lmed, self.radii, self.maxfun = profile, lmed, radii, maxfun
        self.ncalls = 0

    def __call__(self, x):
        if self.ncalls >= self.maxfun:
            raise _MaxFun()
        self.ncalls += 1
        n = len(self.radii)
        terms = []
        for i in range(n):
            with np.errstate(all="ignore"):
                m = model(self.profile, self.radii[i], x)
                t = np.float64(self.lmed[i]) - np.log(m)
                terms.append(float(t ** 2 / n))
        return pairwise_sum(terms)


def line_for_search(x0, alpha, lb, ub):
    """scipy _line_for_search."""
    lmin, lmax = -INF, INF
    first = True
    for i in range(len(x0)):
        if alpha[i] == 0.0:
            continue
        low = (lb[i] - x0[i]) / alpha[i]
        high = (ub[i] - x0[i]) / alpha[i]
        if alpha[i] > 0:
            lo_i, hi_i = low + 0, high + 0
        else:
            lo_i, hi_i = 0 + high, 0 + low
        lmin = lo_i if first else max(lmin, lo_i)
        lmax = hi_i if first else min(lmax, hi_i)
        first = False
    return (lmin, lmax) if lmax >= lmin else (0.0, 0.0)


def _sign(v):
    return (v > 0) - (v < 0)


def fminbound(func, x1, x2, xatol, maxfun=500):
    """scipy _minimize_scalar_bounded; returns (xf, fx)."""
    sqrt_eps = math.sqrt(2.2e-16)
    golden_mean = 0.5 * (3.0 - math.sqrt(5.0))
    a, b = x1, x2
    fulc = a + golden_mean * (b - a)
    nfc, xf = fulc, fulc
    rat = e = 0.0
    x = xf
    fx = func(x)
    num = 1
    ffulc = fnfc = fx
    xm = 0.5 * (a + b)
    tol1 = sqrt_eps * abs(xf) + xatol / 3.0
    tol2 = 2.0 * tol1
    while abs(xf - xm) > (tol2 - 0.5 * (b - a)):
        golden = 1
        if abs(e) > tol1:
            golden = 0
            r = (xf - nfc) * (fx - ffulc)
            q = (xf - fulc) * (fx - fnfc)
            p = (xf - fulc) * q - (xf - nfc) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = e
            e = rat
            if (abs(p) < abs(0.5 * q * r)) and (p > q * (a - xf)) and (p < q * (b - xf)):
                rat = (p + 0.0) / q
                x = xf + rat
                if ((x - a) < tol2) or ((b - x) < tol2):
                    si = _sign(xm - xf) + ((xm - xf) == 0)
                    rat = tol1 * si
            else:
                golden = 1
        if golden:
            if xf >= xm:
                e = a - xf
            else:
                e = b - xf
            rat = golden_mean * e
        si = _sign(rat) + (rat == 0)
        x = xf + si * max(abs(rat), tol1)
        fu = func(x)
        num += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            fulc, ffulc = nfc, fnfc
            nfc, fnfc = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if (fu <= fnfc) or (nfc == xf):
                fulc, ffulc = nfc, fnfc
                nfc, fnfc = x, fu
            elif (fu <= ffulc) or (fulc == xf) or (fulc == nfc):
                fulc, ffulc = x, fu
        xm = 0.5 * (a + b)
        tol1 = sqrt_eps * abs(xf) + xatol / 3.0
        tol2 = 2.0 * tol1
        if num >= maxfun:
            break
    return xf, fx


def linesearch(func, p, xi, tol, lb, ub, fval):
    """scipy _linesearch_powell (bounded branches)."""
    if not any(v != 0.0 for v in xi):
        return fval, p, xi
    b0, b1 = line_for_search(p, xi, lb, ub)
    if b0 == -INF and b1 == INF:
        raise NotImplementedError("unbounded line search (Brent) is unreachable with finite lower bounds")

    def along(alpha):
        return func([p[i] + alpha * xi[i] for i in range(len(p))])

    if b0 != -INF and b1 != INF:
        amin, fret = fminbound(along, b0, b1, xatol=tol / 100)
    else:
        amin, fret = fminbound(lambda u: along(float(np.tan(u))), float(np.arctan(b0)), float(np.arctan(b1)), xatol=tol / 100)
        amin = float(np.tan(amin))
    xi = [amin * v for v in xi]
    return fret, [p[i] + xi[i] for i in range(len(p))], xi


def powell(func, x0, lb, ub, xtol=1e-4, ftol=1e-4):
    """scipy _minimize_powell with bounds, default options (maxiter = maxfev = N*1000).  func: an Objective."""
    x = list(x0)
    N = len(x)
    maxiter = N * 1000
    direc = [[1.0 if i == j else 0.0 for j in range(N)] for i in range(N)]
    fval = func(x)
    x1 = list(x)
    it = 0
    while True:
        try:
            fx = fval
            bigind = 0
            delta = 0.0
            for i in range(N):
                direc1 = direc[i]
                fx2 = fval
                fval, x, direc1 = linesearch(func, x, direc1, xtol * 100, lb, ub, fval)
                if (fx2 - fval) > delta:
                    delta = fx2 - fval
                    bigind = i
            it += 1
            bnd = ftol * (abs(fx) + abs(fval)) + 1e-20
            if 2.0 * (fx - fval) <= bnd:
                break
            if func.ncalls >= func.maxfun:
                break
            if it >= maxiter:
                break
            if math.isnan(fx) and math.isnan(fval):
                break
            direc1 = [x[i] - x1[i] for i in range(N)]
            x1 = list(x)
            _, lmax = line_for_search(x, direc1, lb, ub)
            step = min(lmax, 1)
            x2 = [x[i] + step * direc1[i] for i in range(N)]
            fx2 = func(x2)
            if fx > fx2:
                t = 2.0 * (fx + fx2 - 2.0 * fval)
                temp = fx - fval - delta
                t *= temp * temp
                temp = fx - fx2
                t -= delta * temp * temp
                if t < 0.0:
                    fval, x, direc1 = linesearch(func, x, direc1, xtol * 100, lb, ub, fval)
                    if any(v != 0.0 for v in direc1):
                        direc[bigind] = direc[-1]
                        direc[-1] = direc1
        except _MaxFun:
            break
    return x, fval, it, func.ncalls


def bounds_and_guess(method, r200):
    """dens_profs_tools.py:262-318."""
    prof = method["profile"]

    def b(name, lo):
        return (method[name], method[name]) if name in method else (lo, INF)
    if prof == "einasto":
        return [0.1, 0.18, r200 / 5], [(1e-7, INF), b("alpha", 1e-10), b("r_s", 1e-5)]
    if prof == "alpha_beta_gamma":
        return [0.1, 1.0, 1.0, 1.0, r200 / 5], [(1e-7, INF), b("alpha", 1e-5), b("beta", 1e-5), b("gamma", 1e-5), b("r_s", 1e-5)]
    return [0.1, r200 / 5], [(1e-7, INF), b("r_s", 1e-5)]


def fitDensProf(ROverR200, method, median, r200):
    """dens_profs_tools.py:232-323 for one object; returns (best_fit, nfev, nit)."""
    if method.get("min_method", "Powell") != "Powell":
        raise NotImplementedError("only the reference's default minimiser (Powell) is restated")
    median = np.asarray(median, dtype=np.float64)
    keep = ~np.isnan(median)
    radii = (np.asarray(ROverR200, dtype=np.float64) * r200)[keep]
    med = median[keep]
    x0, bnds = bounds_and_guess(method, float(r200))
    if len(med) == 0:
        avg = float("nan")
    else:
        avg = pairwise_sum([float(v) for v in med]) / len(med)
    scaled = [float(v) / avg for v in med]
    lmed = [_log(v) if v == v else float("nan") for v in scaled]
    f = Objective(method["profile"], lmed, [float(v) for v in radii], len(x0) * 1000)
    x, fval, it, nfev = powell(f, x0, [lo for lo, _ in bnds], [hi for _, hi in bnds])
    x = list(x)
    x[0] *= avg
    return np.asarray(x), nfev, it


def fitDensProfHelper(dens_profs, ROverR200, r200s, method):
    """dens_profs_tools.py:325-359 (the pool map is a plain loop here)."""
    dens_profs = np.asarray(dens_profs, dtype=np.float64)
    out = np.zeros((dens_profs.shape[0], NPARS[method["profile"]]))
    info = np.zeros((dens_profs.shape[0], 2), dtype=np.int64)
    for p in range(dens_profs.shape[0]):
        out[p], info[p, 0], info[p, 1] = fitDensProf(ROverR200, method, dens_profs[p], r200s[p])
    return out, info
