#!/usr/bin/env python3
"""Timings of the BASELINE.json configurations that are NOT the bench line (configs[0], [2], [3]-flavour, [4]),
on one B200, written to profiles/r1_other_configs.json.  Parity for these flavours is covered by
tests/test_gpu_configs.py and tests/test_gpu_parity.py; this script only measures.

    python profiles/other_configs.py [--quick] [--only=config0|config2|config3|config4|gadget ...]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import _lib  # noqa: E402

K = (1e-2, 100, 10)
quick = "--quick" in sys.argv
only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
out = {}


def want(name):
    return not only or name in only


def timed(fn, n=3):
    fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        r = fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return min(t), r


def catalogue(sizes, seed, profile="einasto", vel=False):
    tab = bench.halo_table(len(sizes), seed)
    tab["sizes"] = np.asarray(sizes, dtype=np.int64)
    tab["r_vir"] = tab["r_vir"] * 0 + 0.3 * (tab["sizes"] / 1e4) ** (1 / 3.0)
    if profile != "einasto":
        from cosmic_profiles_b200.mock import _radius_sampler
        tab["u"], tab["cdf"] = _radius_sampler(profile, 5.0)
    cat = bench.generate_on_device(tab, seed, 0)
    if vel:
        rng = np.random.default_rng(seed)
        cat["vxyz"] = np.ascontiguousarray(rng.normal(size=cat["xyz"].shape) * np.array([300.0, 200.0, 100.0]))
    return cat


ctx = _lib.Context(0)


def sec_config0():
    # ---- configs[0]: 100 NFW halos x 1e4, 20 shells, reduced, E1 and S1 -------------------------------------
    cat = catalogue([10000] * 100, 1, "nfw")
    ctx.set_particles(cat["xyz"], None, cat["masses"]); ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
    rr20 = np.logspace(-1.5, 0, 20)
    for shell in (False, True):
        def run():
            ctx.set_option("invalidate_gather", 1)
            o, nit, _, _, _ = ctx.shape(cat["centers"], cat["r200"], rr20, True, 0.0, 0.0, *K, True, shell, False)
            return nit.copy(), ctx.timing()
        ms, (nit, tim) = timed(run)
        pi = float((nit.clip(0).sum(1) * cat["obj_size"]).sum())
        out["config0_100x1e4_R20_reduced_%s" % ("S1" if shell else "E1")] = dict(
            wall_ms=ms, kernel_ms=tim["kernel_ms"], gather_ms=tim["gather_ms"], particle_iterations=pi,
            PI_per_s_kernel=pi / (tim["kernel_ms"] * 1e-3), halo_shells_per_s=2000 / (ms * 1e-3))



def sec_config2():
    # ---- configs[2]: densities on the configs[1] catalogue ---------------------------------------------------
    nh = 2000 if quick else 10000
    tab = bench.halo_table(nh, 20260101)
    cat = bench.generate_on_device(tab, 20260101, 0)
    npart = int(cat["obj_size"].sum())
    ctx.set_particles(cat["xyz"], None, cat["masses"]); ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
    rb = np.logspace(-2, 0, 50)
    from cosmic_profiles_b200 import dens_profs_algos as D  # noqa: E402
    edges = D.sph_bin_edges(rb)


    def sph():
        ctx.set_option("invalidate_gather", 1)
        d, c = ctx.dens_sph(cat["centers"], cat["r200"], edges, 0.0)
        return ctx.timing()


    ms, tim = timed(sph)
    out["config2_dens_sph_R50"] = dict(halos=nh, particles=npart, wall_ms=ms, kernel_ms=tim["kernel_ms"], gather_ms=tim["gather_ms"],
                                       GBps_kernel=npart * 32 / (tim["kernel_ms"] * 1e-3) / 1e9,
                                       halo_bins_per_s=nh * 50 / (ms * 1e-3))


    def kde():
        d = ctx.dens_kernel(cat["centers"], cat["r200"], rb, 0.0)
        return ctx.timing()


    ms, tim = timed(kde, 2)
    out["config2_dens_kernel_R50"] = dict(halos=nh, particles=npart, wall_ms=ms, kernel_ms=tim["kernel_ms"],
                                          evaluations_per_s=npart * 50 / (tim["kernel_ms"] * 1e-3))
    # section 8(f) next-1: the centre finding the drivers do before the hot path, on the device
    for label, mode in (("com", 0), ("mode", 1)):
        ms, tim = timed(lambda: (ctx.centers(mode, True, 0.0, None), ctx.timing())[1])
        out["centers_%s_configs1_catalogue" % label] = dict(halos=nh, particles=npart, wall_ms=ms, kernel_ms=tim["kernel_ms"])
    # ellipsoidal binning needs the local shape profile first (default 70-shell config), then scipy interpolation
    o, nit, _, _, _ = ctx.shape(cat["centers"], cat["r200"], bench.RR, True, 0.0, 0.0, *K, False, False, False)
    o = o.copy()
    o[o == 0.0] = np.nan
    d_, q_, s_ = o[:, :, 0], o[:, :, 1], o[:, :, 2]
    t0 = time.perf_counter()
    ai, bi, ci, mj, it, mn = D.interpolate_shapes(cat["r200"], rb, d_, d_ * q_, d_ * s_, o[:, :, 3:6], o[:, :, 6:9], o[:, :, 9:12])
    interp_s = time.perf_counter() - t0


    def ell():
        d, c = ctx.dens_ell(cat["centers"], ai, bi, ci, mj, it, mn, 0.0)
        return ctx.timing()


    ms, tim = timed(ell, 2)
    # the same interpolation on the device (cpg_shape_interp), profiles staying resident for the binning
    ms_ip, tim_ip = timed(lambda: (ctx.shape_interp(cat["r200"], rb, D.ell_bin_edges(rb), d_, d_ * q_, d_ * s_, o[:, :, 3:6],
                                                    o[:, :, 6:9], o[:, :, 9:12]), ctx.timing())[1], 3)
    ms_dev, _ = timed(lambda: ctx.dens_ell(cat["centers"], None, None, None, None, None, None, 0.0), 2)
    out["config2_dens_ell_R50"] = dict(halos=nh, particles=npart, wall_ms=ms, kernel_ms=tim["kernel_ms"],
                                       host_numpy_interp_s=interp_s, device_interp_wall_ms=ms_ip,
                                       device_interp_kernel_ms=tim_ip["kernel_ms"], wall_ms_resident_profiles=ms_dev,
                                       particle_bin_tests_per_s=npart * 50 / (tim["kernel_ms"] * 1e-3))



def sec_config3():
    # ---- configs[3] flavour: velocity-dispersion local + global on the same catalogue (one GPU's share) ----
    nv = 500 if quick else 3000
    tabv = bench.halo_table(nv, 5)
    catv = bench.generate_on_device(tabv, 5, 0)
    rng = np.random.default_rng(5)
    vx = np.ascontiguousarray(rng.normal(size=catv["xyz"].shape) * np.array([300.0, 200.0, 100.0]))


    def vd():
        ctx.set_option("invalidate_gather", 1)
        o, nit, _, _, _ = ctx.shape(catv["centers"], catv["r200"], bench.RR, True, 0.0, 0.0, *K, False, False, True)
        return nit.copy(), ctx.timing()


    # Gadget MassTable-style constant particle mass (BASELINE configs[3]); with unequal masses the sequential fp32
    # mass normaliser of calcCoM cannot take the closed-form route and dominates the gather (see "unequal" entry)
    mconst = np.full(catv["masses"].shape, 0.0123456789)
    for label, mm in (("", mconst), ("_unequal_masses", catv["masses"])):
        ctx.set_particles(catv["xyz"], vx, mm); ctx.set_catalogue(catv["idx_cat"], catv["obj_size"])
        ms, (nit, tim) = timed(lambda: vd(), 2)
        pi = float((nit.clip(0).sum(1) * catv["obj_size"]).sum())
        out["config3_vdisp_local_R70" + label] = dict(halos=nv, particles=int(catv["obj_size"].sum()), wall_ms=ms,
                                                      kernel_ms=tim["kernel_ms"], gather_ms=tim["gather_ms"], particle_iterations=pi,
                                                      PI_per_s_kernel=pi / (tim["kernel_ms"] * 1e-3),
                                                      algorithmic_GBps_56B=pi * 56 / (tim["kernel_ms"] * 1e-3) / 1e9)




def sec_config4():
    # ---- configs[4]: one large halo, 50 shells, shell-based non-reduced (cluster / DSMEM path) -------------
    nbig = 10_000_000 if quick else 100_000_000
    catb = catalogue([nbig], 9, "nfw")
    ctx.set_particles(catb["xyz"], None, catb["masses"]); ctx.set_catalogue(catb["idx_cat"], catb["obj_size"])
    rr50 = np.logspace(-2, 0, 50)
    for opts in ({}, {"shells_per_pass": 4, "cluster_size": 16}, {"shells_per_pass": 2, "cluster_size": 8}):
        for kk in ("shells_per_pass", "cluster_size", "large_halo_min"):
            ctx.set_option(kk, opts.get(kk, 0))

        def big():
            ctx.set_option("invalidate_gather", 1)
            o, nit, npt, _, _ = ctx.shape(catb["centers"], catb["r200"], rr50, True, 0.0, 0.0, *K, False, True, False)
            return nit.copy(), ctx.timing()
        ms, (nit, tim) = timed(big, 2)
        pi = float(nit.clip(0).sum() * nbig)
        ne = ctx.prefix_lengths(50).astype(np.int64)
        out["config4_single_halo_%g_R50_S1shell_%s" % (nbig, "auto" if not opts else "S%dC%d" % (opts["shells_per_pass"], opts["cluster_size"]))] = dict(
            wall_ms=ms, kernel_ms=tim["kernel_ms"], gather_ms=tim["gather_ms"], iterations=nit.ravel().tolist(),
            particle_iterations=pi, PI_per_s_kernel=pi / (tim["kernel_ms"] * 1e-3),
            algorithmic_GBps=pi * 32 / (tim["kernel_ms"] * 1e-3) / 1e9,
            streamed_GBps=float((ne * nit.clip(0)).sum()) * 32 / (tim["kernel_ms"] * 1e-3) / 1e9)


def sec_gadget():
    """configs[3] as ONE GPU's share of the 1e9-particle / 1e5-halo Gadget-layout snapshot (1.25e4 FoF groups, ~1.2e8
    particles, float32 Coordinates/Velocities, MassTable mass): device catalogue + raw-block ingest + local
    velocity-dispersion shapes, against the fp64-host-array route through cpg_set_particles/cpg_set_catalogue."""
    import torch
    from cosmic_profiles_b200 import gen_catalogues as G, readgadget as RG
    ng = 1500 if quick else 12500
    tabg = bench.halo_table(ng, 77)
    catg = bench.generate_on_device(tabg, 77, 0)
    n = int(catg["obj_size"].sum())
    rng = np.random.default_rng(77)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    pos32 = pin(catg["xyz"].astype(np.float32))
    vel32 = pin((rng.standard_normal(size=(n, 3), dtype=np.float32) * np.array([300.0, 200.0, 100.0], dtype=np.float32)))
    nb = np.ones(ng, dtype=np.int32)
    sl = catg["obj_size"].astype(np.int32)
    fof = sl.copy()
    mtab = 0.0123456789
    # the fp64 host arrays the reference's read_block would hand over (unit factors 1, a = 1)
    xyz64 = pin(pos32.astype(np.float64)); vel64 = pin(vel32.astype(np.float64)); m64 = pin(np.full(n, mtab))
    idx = catg["idx_cat"]
    cen = catg["centers"]
    g = _lib.Context(0)

    def route_f64():
        g.set_particles(xyz64, vel64, m64); g.set_catalogue(idx, sl)
        o, nit, _, _, _ = g.shape(cen, catg["r200"], bench.RR, True, 0.0, 0.0, *K, False, False, True)
        return nit.copy(), g.timing()

    def route_gadget():
        RG.upload_snapshot(g, pos32, vel32, None, mtab, 1.0, 1.0, 1.0, 1.0)
        G.device_catalogue(g, nb, sl, fof, catg["r200"], 200)
        o, nit, _, _, _ = g.shape(cen, catg["r200"], bench.RR, True, 0.0, 0.0, *K, False, False, True)
        return nit.copy(), g.timing()

    ms64, (nit64, tim64) = timed(route_f64, 2)
    ms32, (nit32, tim32) = timed(route_gadget, 2)
    assert np.array_equal(nit64, nit32)
    pi = float((nit32.clip(0).sum(1) * sl).sum())

    def oc():
        g.obj_cat(nb, sl, fof, catg["r200"], 200, fetch_idx=False)
        return g.timing()
    ms_oc, tim_oc = timed(oc, 3)
    ms_oc_fetch, _ = timed(lambda: g.obj_cat(nb, sl, fof, catg["r200"], 200, fetch_idx=True), 2)
    # the reference's calcObjCat on the first groups of the same table (O(n^2) prefix sums + one hstack per object)
    ref = None
    try:
        from oracle import ref_loader
        if ref_loader.available("pristine"):
            ref_loader.load("pristine")
            from cosmic_profiles.gadget.gen_catalogues import calcObjCat
            nr = min(ng, 1500)
            t0 = time.perf_counter()
            with ref_loader.quiet_stdout():
                calcObjCat(nb[:nr].copy(), sl[:nr].copy(), fof[:nr].copy(), catg["r200"][:nr].copy(), np.int32(200))
            ref = dict(groups=nr, particles=int(sl[:nr].sum()), seconds=time.perf_counter() - t0)
    except Exception as e:   # measurement only
        ref = dict(error=repr(e))
    out["config3_gadget_layout_one_gpu_share"] = dict(
        groups=ng, particles=n, particle_iterations=pi,
        e2e_ms_fp64_host_arrays=ms64, e2e_ms_raw_float32_blocks_device_catalogue=ms32,
        h2d_bytes_fp64_route=int(n * 60), h2d_bytes_gadget_route=int(n * 24),
        kernel_ms=tim32["kernel_ms"], gather_ms=tim32["gather_ms"],
        PI_per_s_e2e_gadget_route=pi / (ms32 * 1e-3), PI_per_s_e2e_fp64_route=pi / (ms64 * 1e-3),
        obj_cat_device_ms=tim_oc["kernel_ms"], obj_cat_wall_ms=ms_oc, obj_cat_wall_ms_with_idx_fetch=ms_oc_fetch,
        reference_calcObjCat=ref)
    g.close()


def _ref_fit_one(args):
    import warnings
    warnings.filterwarnings("ignore")
    from cosmic_profiles.dens_profs import dens_profs_tools as T
    rr, method, tup = args
    return T.fitDensProf(rr, method, tup)[0]


def sec_fits():
    """configs[2]: NFW / Hernquist / Einasto / alpha-beta-gamma fits of the 1e4 spherical direct-binning profiles
    (bins [10:] as tests/test_densities.py:91-93), device Powell against the reference's fitDensProf on the host."""
    import multiprocessing as mp
    import warnings
    from cosmic_profiles_b200 import dens_profs_algos as D, dens_profs_tools as DT
    nh = 2000 if quick else 10000
    tab = bench.halo_table(nh, 20260101)
    cat = bench.generate_on_device(tab, 20260101, 0)
    ctx.set_particles(cat["xyz"], None, cat["masses"]); ctx.set_catalogue(cat["idx_cat"], cat["obj_size"])
    rb = np.logspace(-2, 0, 50)
    dens, _ = ctx.dens_sph(cat["centers"], cat["r200"], D.sph_bin_edges(rb), 0.0)
    dens = dens[:, 10:] * 1e10
    dens[dens == 0.0] = np.nan     # empty inner bins of small objects (log(0)); the reference would fit -inf there
    rr = rb[10:]
    ref_ok = False
    try:
        from oracle import ref_loader
        ref_ok = ref_loader.available("pristine")
        if ref_ok:
            ref_loader.load("pristine")
    except Exception:
        ref_ok = False
    cores = len(os.sched_getaffinity(0))
    for prof in ("nfw", "hernquist", "einasto", "alpha_beta_gamma"):
        method = {"profile": prof}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t_prep0 = time.perf_counter()
            DT.prepare(dens, rr, cat["r200"])
            prep_ms = (time.perf_counter() - t_prep0) * 1e3
            ms, best = timed(lambda: DT.fitDensProfHelper(dens, rr, cat["r200"], method), 2)
        info = DT.last_info
        entry = dict(objects=nh, bins=int(dens.shape[1]), wall_ms=ms, kernel_ms=info["timing"]["kernel_ms"], host_prepare_ms=prep_ms,
                     fits_per_s=nh / (ms * 1e-3), mean_nfev=float(info["nfev"].mean()), max_nfev=int(info["nfev"].max()),
                     evaluations_per_s_kernel=float(info["nfev"].sum()) / (info["timing"]["kernel_ms"] * 1e-3),
                     finite_rows=int(np.isfinite(best).all(1).sum()))
        if ref_ok:
            ns = min(nh, int(os.environ.get("CP_FIT_SAMPLE", 8 * cores)))
            pick = np.random.default_rng(1).choice(nh, ns, replace=False)
            jobs = [(rr, method, (dens[i], cat["r200"][i], int(i))) for i in pick]
            with mp.get_context("fork").Pool(cores) as pool:
                pool.map(_ref_fit_one, jobs[:cores])          # warm the workers
                t0 = time.perf_counter()
                ref = np.array(pool.map(_ref_fit_one, jobs))
                dt = time.perf_counter() - t0
            fin = np.isfinite(ref).all(1) & np.isfinite(best[pick]).all(1)
            rel = np.max(np.abs(best[pick][fin] - ref[fin]) / np.abs(ref[fin]), axis=1)
            # where the two optimisers ended in different points: how do the objective values compare?
            radii, lmed, n_valid, avg = DT.prepare(dens[pick], rr, cat["r200"][pick])

            def psi2(x, j):   # toMinimize on the normalised profile (dens_profs_tools.py:227-230), x[0] un-scaled again
                m = int(n_valid[j])
                xx = np.array(x, dtype=np.float64)
                xx[0] = xx[0] / avg[j]
                r = radii[j, :m]
                if prof == "nfw":
                    mod = xx[0] / ((r / xx[1]) * (1 + r / xx[1]) ** 2)
                elif prof == "hernquist":
                    mod = xx[0] / ((r / xx[1]) * (1 + r / xx[1]) ** 3)
                elif prof == "einasto":
                    mod = xx[0] * np.exp(-2 / xx[1] * ((r / xx[2]) ** xx[1] - 1))
                else:
                    mod = xx[0] / ((r / xx[4]) ** xx[3] * (1 + (r / xx[4]) ** xx[1]) ** ((xx[2] - xx[3]) / xx[1]))
                return float(np.sum((lmed[j, :m] - np.log(mod)) ** 2 / m))
            far = np.flatnonzero(fin)[rel > 1e-6]
            dpsi = [(psi2(best[pick][j], j), psi2(ref[j], j)) for j in far]
            entry.update(sample_objects_beyond_1e6=int(len(far)),
                         psi2_gpu_over_ref_of_those=[round(a_ / b_, 6) if b_ > 0 else None for a_, b_ in dpsi][:8])
            entry.update(reference_fits_per_s=ns / dt, reference_cores=cores, reference_sample=ns,
                         speedup=(nh / (ms * 1e-3)) / (ns / dt), sample_within_1e9=float((rel < 1e-9).mean()),
                         sample_within_1e4=float((rel < 1e-4).mean()), sample_max_rel=float(rel.max()),
                         sample_finite_pattern_equal=bool(np.array_equal(np.isfinite(ref).all(1), np.isfinite(best[pick]).all(1))))
        out["config2_fit_%s" % prof] = entry


SECTIONS = dict(fits=sec_fits, config0=sec_config0, config2=sec_config2, config3=sec_config3, config4=sec_config4, gadget=sec_gadget)
for name, fn in SECTIONS.items():
    if want(name):
        fn()
dst = os.path.join(ROOT, "gpurun_out", "r1_other_configs.json")
try:   # keep the entries of sections that were not re-run
    prev = json.load(open(os.path.join(ROOT, "profiles", "r1_other_configs.json")))
except Exception:
    prev = {}
prev.update(out)
json.dump(prev, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
