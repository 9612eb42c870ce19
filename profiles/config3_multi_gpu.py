#!/usr/bin/env python3
"""BASELINE.json configs[3]: a Gadget-layout snapshot of 1e9 DM particles in 1e5 FoF groups, halo-sharded over the
GPUs of one box (1.25e4 groups / ~1.2e8 particles per GPU), shape profiles + velocity-dispersion tensor.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        profiles/config3_multi_gpu.py            (or plain `python profiles/config3_multi_gpu.py` for N = 1)

Every rank owns one GPU and one shard: float32 Coordinates / Velocities blocks + MassTable mass (the layout
readgadget.read_block reads, gadget/readgadget.py:87-124), FoF group tables (read_fof.getFoFSHData, :69-81).  Per
step: raw-block ingest (cpg_set_particles_gadget), device catalogue (cpg_obj_cat + cpg_use_obj_cat), local position
shapes (70 shells, E1 non-reduced) and local velocity-dispersion shapes.  No data-path collective: torch.distributed
only reduces the timings (max) and the work (sum).  One JSON line on rank 0 -> gpurun_out/config3_n<N>.json."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cosmic_profiles_b200 import _lib, gen_catalogues as G, readgadget as RG  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
groups = int(os.environ.get("CP_GROUPS_PER_GPU", "12500"))
steps = int(os.environ.get("CP_STEPS", "3"))
K = (1e-2, 100, 10)

tab = bench.halo_table(groups, 777 + 1000 * rank, size_seed=778)      # same group sizes on every rank
cat = bench.generate_on_device(tab, 777 + 1000 * rank, local_rank)
n = int(cat["obj_size"].sum())
rng = np.random.default_rng(777 + rank)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
pos32 = pin(cat["xyz"].astype(np.float32))
vel32 = pin(rng.standard_normal(size=(n, 3), dtype=np.float32) * np.array([300.0, 200.0, 100.0], dtype=np.float32))
nb = np.ones(groups, dtype=np.int32)
sl = cat["obj_size"].astype(np.int32)
fof = sl.copy()
mtab = 0.0123456789
cen, r200 = cat["centers"], cat["r200"]
ctx = _lib.Context(local_rank)


def step():
    t0 = time.perf_counter()
    RG.upload_snapshot(ctx, pos32, vel32, None, mtab, 1.0, 1.0, 1.0, 1.0)
    G.device_catalogue(ctx, nb, sl, fof, r200, 200)
    t1 = time.perf_counter()
    _, nit_p, _, _, _ = ctx.shape(cen, r200, bench.RR, True, 0.0, 0.0, *K, False, False, False)
    nit_p = nit_p.copy()
    tp = ctx.timing()
    _, nit_v, _, _, _ = ctx.shape(cen, r200, bench.RR, True, 0.0, 0.0, *K, False, False, True)
    nit_v = nit_v.copy()
    tv = ctx.timing()
    t2 = time.perf_counter()
    return nit_p, nit_v, tp, tv, (t1 - t0) * 1e3, (t2 - t1) * 1e3


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


step()
step()
barrier()
acc = np.zeros(6)
w0 = time.perf_counter()
for _ in range(steps):
    nit_p, nit_v, tp, tv, ms_in, ms_sh = step()
    acc += [ms_in, ms_sh, tp["kernel_ms"], tp["gather_ms"], tv["kernel_ms"], tv["gather_ms"]]
barrier()
wall = (time.perf_counter() - w0) * 1e3 / steps
acc /= steps
sizes = sl.astype(np.int64)
pi = float((nit_p.clip(0).sum(1) * sizes).sum() + (nit_v.clip(0).sum(1) * sizes).sum())
t = torch.tensor(list(acc) + [wall], dtype=torch.float64, device="cuda")
w = torch.tensor([pi, float(n), float(groups)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
if rank == 0:
    tt, ww = t.tolist(), w.tolist()
    kern = tt[2] + tt[4]
    line = dict(config="BASELINE configs[3]: Gadget-layout shards, %d groups / GPU, float32 blocks + MassTable, local position "
                       "shapes + local velocity-dispersion shapes (70 shells, E1 non-reduced)" % groups,
                n_gpus=world, groups=int(ww[2]), particles=int(ww[1]), particle_iterations_per_step=ww[0],
                ms_per_step_wall_max_over_ranks=tt[6], ingest_and_catalogue_ms=tt[0], shapes_ms=tt[1],
                position_kernel_ms=tt[2], position_gather_ms=tt[3], vdisp_kernel_ms=tt[4], vdisp_gather_ms=tt[5],
                PI_per_s_end_to_end=ww[0] / (tt[6] * 1e-3), PI_per_s_kernels=ww[0] / (kern * 1e-3),
                h2d_bytes_per_step=int(ww[1] * 24))
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", "config3_n%d.json" % world), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
