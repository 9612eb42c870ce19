"""Seeded, vectorised synthetic halo catalogues for tests and benchmarks.

Statistically equivalent to the reference's ``genHalo`` onion-shell sampler
(cosmic_profiles/mock_tools/mock_halo_gen.py:16: ellipsoidal shells with axes a, b=q*a, c=s*a
filled according to a spherical profile rho(a)) but vectorised: genHalo is a per-particle Python
loop (common/python_routines.py:460-463) and cannot produce 1e7-1e9 particles.  Inputs follow
SURVEY.md section 8(d): internal units (Mpc/h, 1e10 Msun/h, km/s), fp64, FoF-style contiguous
index catalogue (gadget/gen_catalogues.pyx:29-35).
"""
import numpy as np


def _rho(profile, x, alpha=0.18):
    """Dimensionless density vs x = r/r_s (dens_profs/dens_profs_tools.py:22-80 model forms)."""
    if profile == "nfw":
        return 1.0 / (x * (1.0 + x) ** 2)
    if profile == "hernquist":
        return 1.0 / (x * (1.0 + x) ** 3)
    if profile == "einasto":
        return np.exp(-2.0 / alpha * (x ** alpha - 1.0))
    if profile == "alpha_beta_gamma":
        a, b, g = 1.0, 3.0, 1.2
        return 1.0 / (x ** g * (1.0 + x ** a) ** ((b - g) / a))
    raise ValueError(profile)


def _radius_sampler(profile, conc, n_grid=2048):
    """Inverse CDF of M(<a)/M(<r_vir) on a log grid of a/r_vir in [1e-4, 1]."""
    u = np.logspace(-4, 0, n_grid)
    x = u * conc
    dm = _rho(profile, x) * x * x * x  # dM/dln a
    cdf = np.concatenate(([0.0], np.cumsum(0.5 * (dm[1:] + dm[:-1]) * np.diff(np.log(u)))))
    cdf /= cdf[-1]
    return u, cdf


def random_rotations(rng, n):
    """n uniformly distributed rotation matrices (QR of Gaussian matrices, sign-fixed)."""
    a = rng.normal(size=(n, 3, 3))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.einsum("nii->ni", r))[:, None, :]
    det = np.linalg.det(q)
    q[:, :, 0] *= det[:, None]
    return q


def make_catalogue(sizes, seed=20260101, profile="nfw", q_range=(0.5, 0.9), s_range=(0.3, None), conc=5.0,
                   box=None, equal_mass=False, with_velocities=False, shuffle_particles=False):
    """Return a dict with xyz (N,3), masses (N,), [vxyz (N,3)], idx_cat (N,) int32, obj_size (n,) int32,
    r200 (n,), plus the generating axis ratios q_true, s_true.

    Each halo h: r_vir ~ U(0.5,1.5) scaled with N^(1/3), q ~ U(q_range), s ~ U(s_range[0], q),
    random orientation, random centre inside ``box`` (default: far from 0 so that L_BOX=0 is safe).
    """
    rng = np.random.default_rng(seed)
    sizes = np.asarray(sizes, dtype=np.int64)
    n = len(sizes)
    ntot = int(sizes.sum())
    off = np.concatenate(([0], np.cumsum(sizes)))
    halo = np.repeat(np.arange(n), sizes)
    q_true = rng.uniform(q_range[0], q_range[1], size=n)
    s_hi = q_true if s_range[1] is None else np.minimum(q_true, s_range[1])
    s_true = rng.uniform(s_range[0], 1.0, size=n) * 0 + (s_range[0] + (s_hi - s_range[0]) * rng.uniform(size=n))
    r_vir = rng.uniform(0.5, 1.5, size=n) * (sizes / 1.0e4) ** (1.0 / 3.0) * 0.3
    u, cdf = _radius_sampler(profile, conc)
    a = np.interp(rng.uniform(size=ntot), cdf, u) * r_vir[halo]
    z = rng.uniform(-1.0, 1.0, size=ntot)
    ph = rng.uniform(0.0, 2.0 * np.pi, size=ntot)
    st = np.sqrt(1.0 - z * z)
    local = np.empty((ntot, 3))
    local[:, 0] = a * st * np.cos(ph)
    local[:, 1] = a * st * np.sin(ph) * q_true[halo]
    local[:, 2] = a * z * s_true[halo]
    rot = random_rotations(rng, n)
    L = 100.0 if box is None else float(box)
    centre = rng.uniform(0.2 * L, 0.8 * L, size=(n, 3))
    xyz = np.einsum("nij,nj->ni", rot[halo], local) + centre[halo]
    if equal_mass:
        masses = np.full(ntot, 0.0123456789)
    else:
        masses = rng.uniform(0.5, 1.5, size=ntot) * 0.01
    out = dict(q_true=q_true, s_true=s_true, r200=np.ascontiguousarray(r_vir), obj_size=sizes.astype(np.int32),
               offsets=off)
    if with_velocities:
        sig = 150.0 * np.sqrt(sizes / 1.0e4)[halo]
        vloc = rng.normal(size=(ntot, 3)) * np.array([3.0, 2.0, 1.0]) / 3.0 * sig[:, None]
        bulk = rng.normal(size=(n, 3)) * 300.0
        out["vxyz"] = np.ascontiguousarray(np.einsum("nij,nj->ni", rot[halo], vloc) + bulk[halo])
    idx = np.arange(ntot, dtype=np.int32)
    if shuffle_particles:
        # scatter the particles over the snapshot so that idx_cat is a genuine gather
        perm = rng.permutation(ntot)
        inv = np.empty(ntot, dtype=np.int64)
        inv[perm] = np.arange(ntot)
        xyz = xyz[perm]
        masses = masses[perm]
        if with_velocities:
            out["vxyz"] = np.ascontiguousarray(out["vxyz"][perm])
        idx = inv.astype(np.int32)
    out.update(xyz=np.ascontiguousarray(xyz), masses=np.ascontiguousarray(masses), idx_cat=idx)
    return out


def power_law_sizes(n_halo, n_min=1000, n_max=1000000, slope=-1.9, seed=7):
    """Halo sizes drawn from p(N) ~ N^slope on [n_min, n_max] (BASELINE.json configs[1])."""
    rng = np.random.default_rng(seed)
    g = slope + 1.0
    uu = rng.uniform(size=n_halo)
    sizes = (n_min ** g + uu * (n_max ** g - n_min ** g)) ** (1.0 / g)
    return np.maximum(n_min, np.minimum(n_max, sizes.astype(np.int64)))
