#!/usr/bin/env python3
"""Golden vectors for the profile-fit row: runs the reference's own fitDensProf
(dens_profs/dens_profs_tools.py:232-323, imported from oracle/_ref/pristine) with the installed scipy on seeded
mock density profiles and stores inputs, best fits and scipy's nfev / nit in tests/golden/fit_golden.npz.
Run in the build container (needs oracle/_ref)."""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

CASES = [("nfw", {"profile": "nfw"}), ("hernquist", {"profile": "hernquist"}), ("einasto", {"profile": "einasto"}),
         ("einasto_alpha018", {"profile": "einasto", "alpha": 0.18}), ("abg", {"profile": "alpha_beta_gamma"}),
         ("nfw_rs", {"profile": "nfw", "r_s": 0.08})]


def mock_profiles(rng, n, rr):
    """Noisy NFW / Einasto / Hernquist profiles in Msun h^2 / Mpc^3, some with NaN (empty) bins, one with an
    exact zero (log -> -inf) and one all-NaN row."""
    P, R = [], []
    for i in range(n):
        r200 = rng.uniform(0.2, 1.5)
        rs = r200 / rng.uniform(3, 12)
        rho = 10 ** rng.uniform(13, 16)
        r = rr * r200
        kind = i % 3
        if kind == 0:
            d = rho / ((r / rs) * (1 + r / rs) ** 2)
        elif kind == 1:
            d = rho * np.exp(-2 / 0.18 * ((r / rs) ** 0.18 - 1))
        else:
            d = rho / ((r / rs) * (1 + r / rs) ** 3)
        d = d * np.exp(rng.normal(0, 0.15, size=len(r)))
        if i % 7 == 3:
            d[rng.integers(len(r), size=3)] = np.nan
        if i == 10:
            d[5] = 0.0
        if i == 11:
            d[:] = np.nan
        P.append(d)
        R.append(r200)
    return np.array(P), np.array(R)


def main():
    warnings.filterwarnings("ignore")
    ref_loader.load("pristine")
    import scipy
    from scipy import optimize
    from cosmic_profiles.dens_profs import dens_profs_tools as T
    from oracle import fit_oracle as F
    rng = np.random.default_rng(20260401)
    rr = np.logspace(-2, 0, 50)[10:]          # tests/test_densities.py:91-93 fits on [:, 10:]
    dens, r200 = mock_profiles(rng, 24, rr)
    out = {"dens": dens, "rr": rr, "r200": r200,
           "meta": json.dumps({"scipy": scipy.__version__, "numpy": np.__version__, "cases": [c[0] for c in CASES],
                               "methods": [c[1] for c in CASES]})}
    for name, method in CASES:
        best = np.array([T.fitDensProf(rr, method, (dens[i], r200[i], i))[0] for i in range(len(dens))])
        counts = np.zeros((len(dens), 2), dtype=np.int64)
        for i in range(len(dens)):
            keep = ~np.isnan(dens[i])
            med, rad = dens[i][keep], (rr * r200[i])[keep]
            x0, b = F.bounds_and_guess(method, float(r200[i]))
            res = optimize.minimize(T.toMinimize, np.array(x0), method="Powell",
                                    args=(med / np.average(med), rad, method["profile"]), bounds=b)
            counts[i] = res.nfev, res.nit
        out[name + "_best"] = best
        out[name + "_counts"] = counts
        print(name, "mean nfev %.0f" % counts[:, 0].mean(), "finite rows", int(np.isfinite(best).all(1).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fit_golden.npz"), **out)


if __name__ == "__main__":
    main()
