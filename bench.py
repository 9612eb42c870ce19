#!/usr/bin/env python3
"""bench.py -- headline benchmark of the shape-profiling hot path (BASELINE.json configs[1]).

A step = one pass of the hot path over one synthetic catalogue:
    calcMorphLocal  (ellipsoid-based, non-reduced, default Katz config: 70 shells logspace(-1.5,0), IT_TOL 1e-2,
                     IT_WALL 100, IT_MIN 10; python_routines.py:698)  +  calcMorphGlobal (SAFE=6)
on mock Einasto halos whose sizes follow p(N) ~ N^-1.9 on [1e3, 1e6]: 10^4 of them per GPU (~9e7 particles, 2.9 GB
of fp64 x,y,z,m per GPU: larger than the 126 MB L2, so no L2 flush is needed between steps).
Centres are an input of the path (north_star); they are computed once outside the timed region.

Several GPUs (torchrun, one rank per GPU): ONE catalogue of N x 10^4 objects -- one draw of the size table for the whole
catalogue -- is split over the ranks by the library's particle-count-balanced LPT partition (cpg_partition_lpt, the
function cpg_multi shards with); every rank generates and owns only its shard.  No data-path collective; the per-object
results are gathered to rank 0 once, outside the timed steps, and that gather is reported (multi_gpu.host_gather_ms).
`--scaling strong --halos-total H` keeps the catalogue fixed (H objects) as N grows instead.

metric  = particle-iterations/s,  PI = sum_h N_h * sum_k (shape-tensor evaluations of halo-shell (h,k))
value   = PI of all ranks / max over ranks of the device time of the cpg_shape calls, snapshot + catalogue resident
e2e     = PI / time of the full C-ABI sequence from pinned HOST buffers (cpg_set_particles + cpg_set_catalogue +
          cpg_shape x2, results copied back) per step
roofline= the local-shape morphology kernels against the MEASURED fp64 multiply-add peak of the device (fp64 is the
          resource they use most -- 48 % of what the pipe sustains; HBM is at 0.2 of its peak; what bounds them is the
          per-warp dependency chain at 12 resident warps per SM): fp64 work from the kernels'
          own tile counters / CUDA-event time; the contract's HBM figures (algorithmic bytes, physical DRAM traffic
          under ncu) are reported beside it
cpu_baseline = the reference's own compiled calcMorphLocal+calcMorphGlobal (oracle/_ref) on a bounded,
          seeded sample of the same catalogue, all host cores; calcMassesCenters timed separately ("centres excluded").

`--impl reference` times only that reference arm (rank 0).  `--config {0,2,3,4}` measures one of the other BASELINE
configurations on one GPU instead (its own JSON line, same keys).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KATZ = dict(IT_TOL=1e-2, IT_WALL=100, IT_MIN=10)
RR = np.logspace(-1.5, 0, 70)
SAFE = 6.0
METRIC = "particle-iterations/sec"
UNIT = "particle-iterations/s"
BYTES_PER_PI = 32.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------- workload

def halo_table(n_halos, seed, size_seed=None):
    """size_seed: seed of the object sizes (default seed + 1).  Under torchrun every rank draws its own shapes,
    orientations and particles but the SAME sizes, so that the per-GPU work is fixed as N grows (weak scaling)."""
    from cosmic_profiles_b200.mock import _radius_sampler, power_law_sizes, random_rotations
    rng = np.random.default_rng(seed)
    sizes = power_law_sizes(n_halos, 1000, 1000000, -1.9, seed=(seed + 1) if size_seed is None else size_seed)
    q = rng.uniform(0.5, 0.9, size=n_halos)
    s = 0.3 + (q - 0.3) * rng.uniform(size=n_halos)
    r_vir = rng.uniform(0.5, 1.5, size=n_halos) * (sizes / 1.0e4) ** (1.0 / 3.0) * 0.3
    rot = random_rotations(rng, n_halos)
    centre = rng.uniform(20.0, 80.0, size=(n_halos, 3))
    u, cdf = _radius_sampler("einasto", 5.0)
    return dict(sizes=sizes, q=q, s=s, r_vir=r_vir, rot=rot, centre=centre, u=u, cdf=cdf)


def global_halos(halos_per_gpu, world):
    """Weak scaling: the ONE catalogue that N ranks share holds N x halos_per_gpu objects."""
    return int(halos_per_gpu) * int(world)


def rank_objects(sizes, rank, world):
    """Catalogue rows owned by `rank`: the library's particle-count-balanced LPT partition (cpg_partition_lpt)."""
    from cosmic_profiles_b200 import _lib
    if world == 1:
        return np.arange(len(sizes))
    owner = _lib.partition_lpt(np.asarray(sizes, dtype=np.float64), world)
    return np.flatnonzero(owner == rank)


def table_rows(tab, rows):
    """The object table restricted to `rows` (a rank's shard)."""
    out = dict(tab)
    for k in ("sizes", "q", "s", "r_vir", "rot", "centre"):
        out[k] = tab[k][rows]
    return out


def generate_on_device(tab, seed, device):
    """Sample the catalogue with torch on the GPU (plumbing only) and return pinned host numpy arrays."""
    import torch
    dev = torch.device("cuda", device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    sizes = torch.as_tensor(tab["sizes"], device=dev)
    n = int(sizes.numel())
    ntot = int(sizes.sum().item())
    halo = torch.repeat_interleave(torch.arange(n, device=dev), sizes)
    f64 = dict(dtype=torch.float64, device=dev)
    cdf = torch.as_tensor(tab["cdf"], **f64)
    u = torch.as_tensor(tab["u"], **f64)
    p = torch.rand(ntot, generator=g, **f64)
    j = torch.clamp(torch.searchsorted(cdf, p), 1, cdf.numel() - 1)
    w = (p - cdf[j - 1]) / (cdf[j] - cdf[j - 1])
    a = (u[j - 1] + w * (u[j] - u[j - 1])) * torch.as_tensor(tab["r_vir"], **f64)[halo]
    del p, j, w
    z = torch.rand(ntot, generator=g, **f64) * 2 - 1
    ph = torch.rand(ntot, generator=g, **f64) * (2 * np.pi)
    st = torch.sqrt(1 - z * z)
    lx = a * st * torch.cos(ph)
    ly = a * st * torch.sin(ph) * torch.as_tensor(tab["q"], **f64)[halo]
    lz = a * z * torch.as_tensor(tab["s"], **f64)[halo]
    del a, z, ph, st
    rot = torch.as_tensor(tab["rot"], **f64)
    cen = torch.as_tensor(tab["centre"], **f64)
    xyz = torch.empty((ntot, 3), **f64)
    for c in range(3):
        xyz[:, c] = rot[:, c, 0][halo] * lx + rot[:, c, 1][halo] * ly + rot[:, c, 2][halo] * lz + cen[:, c][halo]
    del lx, ly, lz, halo
    masses = (torch.rand(ntot, generator=g, **f64) + 0.5) * 0.01
    xyz_h = torch.empty((ntot, 3), dtype=torch.float64, pin_memory=True)
    m_h = torch.empty((ntot,), dtype=torch.float64, pin_memory=True)
    idx_h = torch.empty((ntot,), dtype=torch.int32, pin_memory=True)
    xyz_h.copy_(xyz)
    m_h.copy_(masses)
    idx_h.copy_(torch.arange(ntot, dtype=torch.int32))   # FoF-style contiguous catalogue
    torch.cuda.synchronize(dev)
    # exact centres of mass (an input of the path), computed once outside every timed region
    hid = torch.repeat_interleave(torch.arange(n, device=dev), sizes)
    msum = torch.zeros(n, **f64).index_add_(0, hid, masses)
    com = torch.zeros((n, 3), **f64).index_add_(0, hid, (xyz - cen[hid]) * masses[:, None]) / msum[:, None] + cen
    com_h = com.cpu().numpy()
    del xyz, masses, hid
    torch.cuda.empty_cache()
    return dict(xyz=xyz_h.numpy(), masses=m_h.numpy(), idx_cat=idx_h.numpy(),
                obj_size=np.asarray(tab["sizes"], dtype=np.int32), r200=np.ascontiguousarray(tab["r_vir"]),
                centers=np.ascontiguousarray(com_h), _keep=(xyz_h, m_h, idx_h))


def stratified_sample(sizes, budget_particles, seed=1):
    """A seeded random subset of the objects (so the size distribution is that of the catalogue), leaving out
    objects larger than 1/16 of the budget so that the reference's OpenMP prange over objects has enough
    independent objects to keep every host core busy; filled up to the particle budget."""
    sizes = np.asarray(sizes, dtype=np.int64)
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(sizes))
    cap = max(int(budget_particles // 16), int(sizes.min()))
    pick, tot = [], 0
    for h in order:
        if sizes[h] > cap or tot + sizes[h] > budget_particles:
            continue
        pick.append(h)
        tot += int(sizes[h])
    if not pick:
        pick = [int(np.argmin(sizes))]
    return np.sort(np.asarray(pick))


# --------------------------------------------------------------------------------------- reference arm

def reference_sample_inputs(cat, pick):
    off = np.concatenate(([0], np.cumsum(cat["obj_size"].astype(np.int64))))
    sel = np.concatenate([np.arange(off[h], off[h + 1]) for h in pick])
    xyz = np.ascontiguousarray(cat["xyz"][sel])
    m = np.ascontiguousarray(cat["masses"][sel])
    sizes = cat["obj_size"][pick].astype(np.int32)
    return xyz, m, np.arange(len(sel), dtype=np.int32), sizes, np.ascontiguousarray(cat["r200"][pick])


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must use every host core.  Set the variable
    (for a libgomp that is not loaded yet) and the ICV of an already loaded libgomp."""
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except Exception:
        pass
    return cores


def time_reference(cat, pick, steps, warmup):
    """Time the reference's own drivers (compiled from /root/reference into oracle/_ref; falls back to the
    C port of the oracle if that build is absent) on the sampled objects.  Returns (PI/s, info)."""
    from oracle import cp_oracle as O
    from oracle import ref_loader
    xyz, m, idx, sizes, r200 = reference_sample_inputs(cat, pick)
    cores = use_all_host_cores()
    k = (KATZ["IT_TOL"], KATZ["IT_WALL"], KATZ["IT_MIN"])
    # particle-iterations of the sample, from the oracle's integer outputs (same algorithm, same counts)
    _, (nit_l, _, _, _) = O.morph(xyz, None, m, r200, idx, sizes, 0.0, RR, *k, "com", False, False, True)
    _, (nit_g, _, _, _) = O.morph(xyz, None, m, r200, idx, sizes, 0.0, None, *k, "com", False, False, False, SAFE=SAFE)
    pi = float((nit_l.clip(0).sum(1) * sizes).sum() + (nit_g.clip(0) * sizes).sum())
    if ref_loader.available("pristine"):
        ns = ref_loader.load("pristine")
        kind = "reference"

        def step():
            with ref_loader.quiet_stdout():
                ns.calcMorphLocal(xyz, m, r200, idx, sizes, 0.0, RR, *k, "com", False, False)
                ns.calcMorphGlobal(xyz, m, r200, idx, sizes, 0.0, *k, "com", SAFE, False)
    else:
        kind = "port"

        def step():
            O.morph(xyz, None, m, r200, idx, sizes, 0.0, RR, *k, "com", False, False, True, nthreads=cores)
            O.morph(xyz, None, m, r200, idx, sizes, 0.0, None, *k, "com", False, False, False, SAFE=SAFE, nthreads=cores)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    # "centres excluded" (SURVEY.md 8(d)): both reference drivers run the serial Python centre loop once each
    # (shape_profs_algos.pyx:586-591, :696-701); the GPU arm takes centres as given.  Timed on its own through
    # calcMassesCenters, which runs the same loop (dens_profs_algos.pyx:39-44).
    cen_s = None
    if kind == "reference":
        c0 = time.perf_counter()
        with ref_loader.quiet_stdout():
            ns.calcMassesCenters(xyz, m, idx, sizes, 0.0, "com")
        cen_s = time.perf_counter() - c0
    info = dict(kind=kind, cores=cores,
                sample="%d of %d objects (seeded random subset, objects <= budget/16, %d particles, %.3g particle-iterations), "
                       "calcMorphLocal 70 shells + calcMorphGlobal incl. the reference's own centre loop, "
                       "OMP_NUM_THREADS=%d" % (len(pick), len(cat["obj_size"]), int(sizes.sum()), pi, cores),
                seconds_per_step=dt, centres_seconds=cen_s,
                centres_excluded_value=(pi / max(dt - 2.0 * cen_s, 1e-9)) if cen_s is not None else None)
    return pi / dt, info


def time_dropin(cat, pi_step):
    """e2e_dropin: the call a user of the reference makes -- the public drivers calcMorphLocal + calcMorphGlobal of
    cosmic_profiles_b200 (what install() binds into the reference's classes) with plain PAGEABLE numpy arrays and no
    centres given ('com' centres on the device, as the reference's drivers compute them inside).  cold: first call on
    a fresh session, snapshot upload included; warm: the same arrays again (the reference's class hands the same xyz /
    masses to every driver, cosmic_base_class.pyx:155,232: the snapshot cache skips the upload)."""
    from cosmic_profiles_b200 import session, shape_profs_algos as S
    xyz = np.array(cat["xyz"], copy=True)          # pageable copies: what a numpy user holds
    masses = np.array(cat["masses"], copy=True)
    idx = np.array(cat["idx_cat"], copy=True)
    size = np.array(cat["obj_size"], copy=True)
    r200 = np.array(cat["r200"], copy=True)
    k = (KATZ["IT_TOL"], KATZ["IT_WALL"], KATZ["IT_MIN"])

    def pair():
        t0 = time.perf_counter()
        S.calcMorphLocal(xyz, masses, r200, idx, size, 0.0, RR, *k, "com", False, False)
        t1 = time.perf_counter()
        S.calcMorphGlobal(xyz, masses, r200, idx, size, 0.0, *k, "com", SAFE, False)
        t2 = time.perf_counter()
        return 1e3 * (t1 - t0), 1e3 * (t2 - t1)
    session.close_all()
    pair()                      # first use of the session: device buffers, pinned scratch (not a steady-state cost)
    session.close_all_snapshots()
    cold = pair()               # buffers exist, snapshot not resident: uploads included
    warm = min((pair() for _ in range(3)), key=sum)
    session.close_all()
    return {"unit": UNIT, "cold_ms": {"calcMorphLocal": cold[0], "calcMorphGlobal": cold[1]},
            "warm_ms": {"calcMorphLocal": warm[0], "calcMorphGlobal": warm[1]},
            "cold_value": pi_step / (sum(cold) * 1e-3), "warm_value": pi_step / (sum(warm) * 1e-3),
            "h2d_bytes_cold": int(xyz.nbytes + masses.nbytes + 12 * len(size)),
            "note": "public drop-in drivers, pageable numpy in, reference tuple out (NaN masking, recentring included); "
                    "centres 'com' computed on the device inside the call"}


# ------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled while the timed regions run (B200_PROFILING.md's clocks
    line).  Primary source: NVML in-process (the library nvidia-smi itself reads; a sample every 20 ms, so even a
    200 ms timed region is covered); fallback: an `nvidia-smi -lms 100` child.  Every sample carries its host
    timestamp; stop(t0, t1) keeps the samples taken inside [t0, t1]."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []       # (t, sm_mhz, sm_max_mhz, [reason names])
        self.proc = None
        self.source = None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.device
            if vis:
                ent = vis.split(",")[self.device].strip()
                phys = int(ent) if ent.isdigit() else self.device
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                     ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                     ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((time.perf_counter(), sm, mx, [n for n, b in names if bits & b]))
                    except Exception:
                        pass
                    self._stop.wait(0.02)
            self.source = "nvml"
            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                why = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8])
                       if v.lower().startswith("active")]
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]), why))
            except Exception:
                pass

    def stop(self, t0=None, t1=None):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"], "samples": 0}
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.th.join(timeout=1)
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        window = "timed regions"
        if not rows:   # region shorter than the sampling period: fall back to everything seen since start()
            rows, window = list(self.rows), "whole run (no sample fell inside the timed regions)"
        sm = [r[1] for r in rows if r[1] > 0]
        reasons = sorted({n for r in rows for n in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "reasons": reasons, "samples": len(rows), "source": self.source, "window": window}


# ------------------------------------------------------------------------------------ rank reduction

def reduce_over_ranks(times_ms, work, world, device):
    """Multi-GPU bookkeeping (no data-path collective exists): every time-like entry is the MAX over ranks,
    every work-like entry the SUM.  Uses whatever torch.distributed backend is initialised (nccl on the GPU
    box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [float(x) for x in w.tolist()]


def rank_seed(seed, rank):
    """Each rank samples its own catalogue of the per-GPU size (weak scaling)."""
    return seed + 1000 * rank


# --------------------------------------------------------------------------------------------- main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--halos", type=int, default=10000, help="objects per GPU (BASELINE configs[1]: 10^4)")
    ap.add_argument("--cpu-budget", type=float, default=8.0e6, help="particles in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e: only the strictly serial measurement")
    ap.add_argument("--no-dropin", action="store_true", help="skip the e2e_dropin measurement (public drivers, pageable numpy)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: N x --halos objects in the one catalogue; strong: --halos-total objects whatever N")
    ap.add_argument("--halos-total", type=int, default=100000, help="objects of the catalogue under --scaling strong")
    ap.add_argument("--config", type=int, default=1, choices=[0, 1, 2, 3, 4],
                    help="BASELINE.json configuration to measure (1 = the headline; the others: one GPU, own JSON line)")
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--opt", action="append", default=[], help="scheduling tunable name=value (cpg_set_option)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: whatever libraries write to fd 1 on the way (NCCL prints its version
    # banner there) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    config = {"workload": "BASELINE configs[1]: %d mock Einasto halos/GPU, sizes p(N)~N^-1.9 on [1e3,1e6], "
                          "ellipsoid-based local shapes (70 shells, non-reduced, IT_TOL 1e-2) + global shape, "
                          "CENTER=com given" % a.halos,
              "halos_per_gpu": a.halos, "shells": 70, "l2": "inputs_larger_than_L2", "sharding": "objects"}

    if a.impl == "reference":
        if rank != 0:
            return
        use_all_host_cores()
        import torch
        tab = halo_table(a.halos, a.seed)   # the per-GPU catalogue size: the sample is drawn from 10^4 objects at every N
        if torch.cuda.is_available():
            cat = generate_on_device(tab, a.seed, 0)
        else:   # CPU-only container: same generator family on the host, small sizes only
            from cosmic_profiles_b200.mock import make_catalogue
            cat = make_catalogue(tab["sizes"], seed=a.seed, profile="einasto")
        pick = stratified_sample(np.asarray(cat["obj_size"], dtype=np.int64), a.cpu_budget)
        v, info = time_reference(cat, pick, max(1, a.steps), a.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * info["seconds_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                                 "sample": info["sample"], "centres_excluded_value": info.get("centres_excluded_value"),
                                 "centres_seconds": info.get("centres_seconds")},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if a.config != 1:
        # the other BASELINE configurations: one GPU, one JSON line each (bench_configs.py)
        if rank != 0:
            return
        import bench_configs
        emit(bench_configs.RUN[a.config](sys.modules[__name__], a, ClockSampler(local_rank)))
        return
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from cosmic_profiles_b200 import _lib

    # ---- ONE catalogue, partitioned over the ranks -----------------------------------------------------------
    t0 = time.perf_counter()
    n_global = a.halos_total if a.scaling == "strong" else global_halos(a.halos, world)
    tab = halo_table(n_global, a.seed)
    tp0 = time.perf_counter()
    mine = rank_objects(tab["sizes"], rank, world)
    partition_ms = 1e3 * (time.perf_counter() - tp0)
    cat = generate_on_device(table_rows(tab, mine), rank_seed(a.seed, rank), local_rank)
    n_part = int(cat["obj_size"].sum())
    n_mine = len(mine)
    log("[rank %d] shard: %d of %d objects, %d particles (%.2f GB x,y,z,m), partition %.1f ms, generated in %.1f s"
        % (rank, n_mine, n_global, n_part, n_part * 32 / 1e9, partition_ms, time.perf_counter() - t0))
    config["halos_total"] = n_global
    config["sharding"] = "one catalogue, objects split by the library's LPT partition on particle counts (cpg_partition_lpt)"

    ctx = _lib.Context(local_rank)
    ctx.set_option("snapshot_cache", 0)   # every upload() of the e2e phases must really move its bytes
    for kv in a.opt:
        name, val = kv.split("=")
        ctx.set_option(name, int(val))
    k = (KATZ["IT_TOL"], KATZ["IT_WALL"], KATZ["IT_MIN"])

    def upload(ctx=ctx):
        # cpg_set_snapshot = cpg_set_particles + cpg_set_catalogue (the host-side pass over idx_cat overlaps the DMA)
        ctx.set_snapshot(cat["xyz"], None, cat["masses"], cat["idx_cat"], cat["obj_size"])

    def shapes(ctx=ctx):
        # a step always redoes the catalogue gather ("includes the on-device gather"); the
        # global-shape call of the same step then reuses it, as a getShapeCatLocal + getShapeCatGlobal
        # sequence on one catalogue would
        ctx.set_option("invalidate_gather", 1)
        out_l, nit_l, _, _, _ = ctx.shape(cat["centers"], cat["r200"], RR, True, 0.0, 0.0, *k, False, False, False,
                                          zero_copy=True)
        nit_l = nit_l.copy()   # a view into the context's page-locked result scratch
        t_l = ctx.timing()
        out_g, nit_g, _, _, _ = ctx.shape(cat["centers"], cat["r200"], None, False, SAFE, 0.0, *k, False, False, False,
                                          zero_copy=True)
        nit_g = nit_g.copy()
        t_g = ctx.timing()
        return nit_l, nit_g, t_l, t_g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing (value) --------------------------------------------------------
    upload()
    for _ in range(a.warmup):
        nit_l, nit_g, _, _ = shapes()
    sizes = cat["obj_size"].astype(np.int64)
    pi_local = float((nit_l.clip(0).sum(1) * sizes).sum())
    pi_global = float((nit_g.clip(0)[:, 0] * sizes).sum())
    pi_step = pi_local + pi_global
    try:   # particles the local-shape kernels really stream (radial prefixes), for the roofline discussion
        n_eff = ctx.prefix_lengths(len(RR)).astype(np.int64)
        visited_local = float((n_eff * nit_l.clip(0)).sum())
    except Exception:
        visited_local = pi_local
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    t_clk0 = time.perf_counter()
    dev_ms = 0.0
    brk = {}
    kern_local_ms = 0.0
    launches = 0
    hot_launches_local = 0
    fma_local = 0.0
    per_ts = 1.0 if ctx.stat("f32_test") else 2.0
    w0 = time.perf_counter()
    for _ in range(a.steps):
        _, _, t_l, t_g = shapes()
        for nm, t in (("local", t_l), ("global", t_g)):
            dev_ms += t["h2d_ms"] + t["gather_ms"] + t["kernel_ms"] + t["d2h_ms"]
            launches += t["kernel_launches"]
            for kk in ("h2d_ms", "gather_ms", "kernel_ms", "d2h_ms"):
                brk[nm + "_" + kk] = brk.get(nm + "_" + kk, 0.0) + t[kk] / a.steps
        kern_local_ms += t_l["kernel_ms"]
        hot_launches_local += t_l["hot_kernel_launches"]
        # fp64 work of the local-shape kernels' streaming loop from their own counters: per 128-particle tile 6 products
        # per particle; per (tile, shell) 6 multiply-adds of the membership form + 6 tensor updates per particle
        fma_local += 128.0 * 6.0 * (t_l["tile_visits"] + per_ts * t_l["tile_shell_visits"])
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - w0)
    fp64_peak = ctx.measure_fp64_peak()   # TFMA/s, outside every timed region

    # ---- end-to-end timing from pinned host buffers (e2e) ---------------------------------------
    # Streaming configuration of the public option "direct_results" (include/cosmic_gpu.h): the result copy of a step
    # hides behind the next step's upload anyway, so the kernels keep their results on the device and the host does
    # not have to initialise the result buffers (measured: ~1 % at N = 1, more where the host's DMA is the bottleneck).
    E2E_OPTIONS = {"direct_results": 0}
    for name, val in E2E_OPTIONS.items():
        ctx.set_option(name, val)
    # (1) strictly serial: every step uploads its inputs, computes, reads its results back, then the next
    upload()
    shapes()   # untimed: the option change above (re)allocates the device-side result buffers
    barrier()
    e0 = time.perf_counter()
    e2e_steps = max(1, min(a.steps, 3))
    for _ in range(e2e_steps):
        upload()
        shapes()
    barrier()
    e2e_serial_ms = 1e3 * (time.perf_counter() - e0) / e2e_steps
    # (2) double-buffered: two contexts (two sets of device buffers), two host threads; step i+1's host->device
    # copies overlap step i's kernels.  Every step still uploads all of its inputs and downloads all of its
    # results inside the timed region; this is the sustained rate for a stream of catalogues (snapshots).
    e2e_ms, e2e_mode = e2e_serial_ms, "serial"
    if not a.no_pipeline:
        ctx2 = _lib.Context(local_rank)
        ctx2.set_option("snapshot_cache", 0)
        for kv in a.opt:
            name, val = kv.split("=")
            ctx2.set_option(name, int(val))
        for name, val in E2E_OPTIONS.items():
            ctx2.set_option(name, val)
        upload(ctx2)
        shapes(ctx2)                       # warm its buffers and scratch
        # long enough that the one compute the last upload cannot hide behind is a small share of the run
        n_pipe = 2 * max(2, min(2 * a.steps, 10))
        errs = []

        copy_lock, compute_lock = threading.Lock(), threading.Lock()

        def worker(c, n):
            # one upload and one compute in flight at a time: the two threads fall into the natural
            # copy(i+1) || compute(i) rhythm instead of fighting over PCIe and the SMs in lock-step
            try:
                for _ in range(n):
                    with copy_lock:
                        upload(c)
                    with compute_lock:
                        shapes(c)
            except Exception as e:      # re-raised below
                errs.append(e)

        barrier()
        e0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(c, n_pipe // 2)) for c in (ctx, ctx2)]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        barrier()
        if errs:
            raise errs[0]
        pipe_ms = 1e3 * (time.perf_counter() - e0) / n_pipe
        if pipe_ms < e2e_serial_ms:
            e2e_ms, e2e_mode = pipe_ms, "double-buffered (2 contexts, uploads of step i+1 overlap kernels of step i)"
        ctx2.close()
    # host->device ceiling of this box with all ranks copying at once (what bounds e2e): plain pinned copies of the
    # same size as one step's inputs, CUDA events, no kernels
    h2d_probe = torch.from_numpy(cat["xyz"])
    dev_probe = torch.empty_like(h2d_probe, device="cuda")
    barrier()
    p0 = time.perf_counter()
    for _ in range(2):
        dev_probe.copy_(h2d_probe, non_blocking=True)
    barrier()
    h2d_gbs = 2 * cat["xyz"].nbytes / (time.perf_counter() - p0) / 1e9
    del dev_probe
    clk = clocks.stop(t_clk0, time.perf_counter())   # the resident-input steps and both e2e measurements
    cat_ranges = bool(_lib.catalogue_ranges(cat["idx_cat"], cat["obj_size"])[0])
    # bytes that really cross PCIe per step: a contiguous-range catalogue travels as (start, size) per object
    idx_bytes = (8 + 4) * n_mine if cat_ranges else cat["idx_cat"].nbytes + cat["obj_size"].nbytes
    h2d = int(cat["xyz"].nbytes + cat["masses"].nbytes + idx_bytes
              + 2 * (cat["centers"].nbytes + cat["r200"].nbytes) + RR.nbytes)
    d2h = int(t_l["d2h_bytes"] + t_g["d2h_bytes"])

    # ---- host gather of the per-object results to rank 0 (outside the timed steps; N > 1 only) ---------------
    gather_ms = 0.0
    if world > 1:
        ctx.set_option("direct_results", 1)
        out_l, nit_all, npt_all, _, _ = ctx.shape(cat["centers"], cat["r200"], RR, True, 0.0, 0.0, *k, False, False, False)
        barrier()
        g0 = time.perf_counter()
        n_all = [None] * world
        dist.all_gather_object(n_all, n_mine)
        n_max = max(n_all)
        send = torch.zeros((n_max, len(RR), 12), dtype=torch.float64, device="cuda")
        send[:n_mine].copy_(torch.from_numpy(out_l))
        recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
        dist.gather(send, recv, dst=0)
        if rank == 0:
            full = np.zeros((n_global, len(RR), 12))
            owner = _lib.partition_lpt(np.asarray(tab["sizes"], dtype=np.float64), world)
            for r in range(world):
                full[np.flatnonzero(owner == r)] = recv[r][:n_all[r]].cpu().numpy()
        barrier()
        gather_ms = 1e3 * (time.perf_counter() - g0)

    # ---- reduce over ranks: time = max, work = sum -------------------------------------------------
    (step_ms, wall_step_ms, e2e_step_ms, kern_ms, e2e_serial_step_ms, part_ms), \
        (pi_all, pi_local_all, launches_all, h2d_all, d2h_all, visited_all, fma_all, kern_sum, step_sum, npart_all, h2d_gbs_all) = \
        reduce_over_ranks([dev_ms / a.steps, wall_ms / a.steps, e2e_ms, kern_local_ms / a.steps, e2e_serial_ms, partition_ms],
                          [pi_step, pi_local, float(launches), float(h2d), float(d2h), visited_local, fma_local / a.steps,
                           kern_local_ms / a.steps, dev_ms / a.steps, float(n_part), h2d_gbs], world, "cuda")
    per_rank = None
    if world > 1:
        mine_stats = [rank, n_mine, n_part, dev_ms / a.steps, kern_local_ms / a.steps, pi_step, clk]
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine_stats)
        # the clocks line covers every GPU of the run: lowest median SM clock, union of the throttle reasons
        meds = [r[6].get("sm_mhz") for r in per_rank if r[6].get("sm_mhz")]
        clk = dict(clk, sm_mhz=min(meds) if meds else clk.get("sm_mhz"),
                   reasons=sorted({x for r in per_rank for x in (r[6].get("reasons") or [])}),
                   sm_mhz_per_rank=[r[6].get("sm_mhz") for r in per_rank], window="timed regions, all ranks")

    cpu = None
    dropin = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        pick = stratified_sample(sizes, a.cpu_budget)
        v, info = time_reference(cat, pick, 1, 0)
        cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
               "centres_excluded_value": info.get("centres_excluded_value"),
               "centres_seconds": info.get("centres_seconds")}
    if rank == 0 and world == 1 and not a.no_dropin:
        dropin = time_dropin(cat, pi_step)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        traffic_src = None
        for fn in ("r3_traffic.json", "r2_traffic.json", "r1d_traffic.json"):   # DRAM bytes of the local-shape kernels under ncu, same workload
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", fn)))
                if int(tr.get("workload_halos", -1)) == a.halos and world == 1:
                    traffic = float(tr["dram_bytes_per_local_shape_launch_set"])
                    traffic_src = "profiles/%s (ncu --set full, same workload)" % fn
                    break
            except Exception:
                pass
        pipe_occ = None
        try:   # ncu: sm__inst_executed_pipe_fp64 of the two local-shape kernels / the microbenchmark's ceiling in the same units
            pf = "r3_fp64_pipe.json" if os.path.exists(os.path.join(ROOT, "profiles", "r3_fp64_pipe.json")) else "r2_fp64_pipe.json"
            po = json.load(open(os.path.join(ROOT, "profiles", pf)))
            pipe_occ = {"warp_kernel": po["warp_kernel_pct"] / po["dfma_ceiling_pct"],
                        "cluster_kernel": po["cluster_kernel_pct"] / po["dfma_ceiling_pct"], "source": "profiles/" + pf}
        except Exception:
            pass
        kern_mean = kern_sum / world
        algorithmic_gbs = pi_local_all / world * BYTES_PER_PI / (kern_ms * 1e-3) / 1e9   # per GPU
        achieved_tflops = 2.0 * fma_all / world / (kern_mean * 1e-3) / 1e12               # per GPU, 2 flop per FMA
        peak_tflops = 2.0 * fp64_peak
        line = {"metric": METRIC, "value": pi_all / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": a.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "halo_shells_per_s": n_global * 71 / (step_ms * 1e-3),
                "wall_ms_per_step": wall_step_ms,
                "e2e": {"value": pi_all / (e2e_step_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                        "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_step_ms, "mode": e2e_mode,
                        "serial_ms_per_step": e2e_serial_step_ms, "options": E2E_OPTIONS,
                        "serial_value": pi_all / (e2e_serial_step_ms * 1e-3),
                        "catalogue_as_ranges": cat_ranges,
                        "h2d_ceiling_GBps_all_ranks": h2d_gbs_all,
                        "h2d_ceiling_ms_per_step": h2d_all / (h2d_gbs_all * 1e9) * 1e3 if h2d_gbs_all else None,
                        "note": "h2d_ceiling: plain pinned host->device copies issued by all ranks at once on this box; "
                                "e2e cannot be faster than h2d_bytes / that rate"},
                "e2e_dropin": dropin,
                "gpu_launches": int(launches_all),
                "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                             "frac": achieved_tflops / peak_tflops if peak_tflops else None, "traffic": traffic,
                             "traffic_source": traffic_src,
                             "kernel": "morph_wduo_kernel/morph_cluster_kernel (local shapes)",
                             "peak_source": "measured in this run: cpg_measure_fp64_peak (independent DFMA chains on every SM, "
                                            "CUDA events), x2 flop per multiply-add (of measured)",
                             "fp64_fma_per_launch_set": fma_all / world,
                             "kernel_ms": kern_ms, "kernel_ms_mean_over_ranks": kern_mean,
                             "hbm": {"peak_GBps": hbm_peak,
                                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                                     "physical_GBps": traffic / (kern_ms * 1e-3) / 1e9 if traffic else None,
                                     "physical_frac": traffic / (kern_ms * 1e-3) / 1e9 / hbm_peak if traffic else None,
                                     "algorithmic_bytes_per_launch_set": pi_local_all / world * BYTES_PER_PI,
                                     "algorithmic_GBps": algorithmic_gbs,
                                     "algorithmic_frac": algorithmic_gbs / hbm_peak,
                                     "streamed_particle_iterations": visited_all / world,
                                     "streamed_GBps": visited_all / world * BYTES_PER_PI / (kern_ms * 1e-3) / 1e9},
                             "fp64_pipe_occupancy_ncu": pipe_occ,
                             "note": "fp64 is the resource the morphology loop uses most (HBM: hbm.physical_frac of its peak): achieved = fp64 "
                                     "multiply-adds of the streaming loop (the kernels' own tile counters: 128 particles x 6 products per "
                                     "tile + 128 x 12 per (tile, shell): membership form + tensor update) / CUDA-event kernel time, against "
                                     "the fp64 FMA rate measured on this GPU in this run.  fp64_pipe_occupancy_ncu = the pipe counter of the "
                                     "two kernels under ncu / what a pure DFMA stream reads on the same counter (92 %): about half; the "
                                     "per-iteration eigen-solves and the compares share that half with the streaming loop.  The kernels are "
                                     "bound by the dependency chain of each warp (12 resident warps per SM at 168 registers, stall reason "
                                     "'wait' 51 %, instruction fetch 10 %; every warp of the default kernel 1d works on two tasks and pays "
                                     "the eigen-solve step once for both), not by the pipe: profiles/README.md.  hbm.physical = DRAM bytes under ncu / kernel time; "
                                     "hbm.algorithmic = 32 B x particle-iterations (the contract's SURVEY 8(d) figure), > peak because the "
                                     "radial ordering and the prefix tensor table let an iteration skip most of an object's particles."},
                "cpu_baseline": cpu, "clocks": clk, "breakdown_ms_per_step_rank0": {k_: round(v_, 3) for k_, v_ in brk.items()},
                "particle_iterations_per_step": pi_all, "particles_total": int(npart_all)}
        if world > 1:
            steps_r = [r[3] for r in per_rank]
            kerns_r = [r[4] for r in per_rank]
            line["multi_gpu"] = {"partition": "LPT on particle counts (cpg_partition_lpt), one catalogue of %d objects" % n_global,
                                 "partition_ms": part_ms, "host_gather_ms": gather_ms,
                                 "host_gather": "per-object results (n, 70, 12) of every rank gathered to rank 0 over NCCL and "
                                                "scattered into catalogue order on the host; once, outside the timed steps",
                                 "per_rank": [{"rank": r[0], "objects": r[1], "particles": r[2], "ms_per_step": r[3],
                                               "local_kernel_ms": r[4], "particle_iterations": r[5]} for r in per_rank],
                                 "imbalance_step_max_over_mean": max(steps_r) / (sum(steps_r) / world),
                                 "imbalance_kernel_max_over_mean": max(kerns_r) / (sum(kerns_r) / world),
                                 "imbalance_particles_max_over_mean": max(r[2] for r in per_rank) / (npart_all / world)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
