"""CPU tests of the N>1 plumbing (there is no data-path collective: objects are sharded, results are
gathered on the host; bench.py only reduces timings).
 * world_size-2 gloo run of bench.reduce_over_ranks (time = max over ranks, work = sum);
 * the shim's device sharding: with the executor replaced by the CPU oracle, sharding objects over
   2 and 3 "devices" must reproduce the unsharded result exactly, object order and all."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        times, work = bench.reduce_over_ranks([10.0 + rank, 3.0 - rank], [100.0 * (rank + 1), 1.0], world, "cpu")
        seeds = bench.rank_seed(7, rank)
        q.put((rank, times, work, seeds))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_reduction():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, times, work, seed in got:
        assert times == [11.0, 3.0]          # max over ranks
        assert work == [300.0, 2.0]          # sum over ranks
    assert sorted(g[3] for g in got) == [7, 1007]   # distinct per-rank catalogues


def test_shim_sharding_matches_unsharded(monkeypatch):
    from cosmic_profiles_b200 import session, shape_profs_algos as S
    from cosmic_profiles_b200.mock import make_catalogue
    from oracle import cp_oracle as O

    calls = []

    def fake_run_one(device, xyz, vxyz, masses, idx_cat, obj_size, cen, r200, rr, local, SAFE, L_BOX, IT_TOL, IT_WALL,
                     IT_MIN, reduced, shell_based, CENTER="com"):
        calls.append((device, len(obj_size)))
        if cen is None:   # what cpg_centers would do on the shard
            cen = O.centres(xyz, masses, idx_cat, obj_size, L_BOX, CENTER, True)
        _, (nit, npt, vc, out) = O.morph(xyz, vxyz, masses, r200, idx_cat, obj_size, L_BOX, rr, IT_TOL, IT_WALL, IT_MIN,
                                         "com", reduced, shell_based, local, SAFE=SAFE, centers=cen)
        m = np.array([masses[idx_cat[a:b]].sum() for a, b in zip(np.cumsum(obj_size) - obj_size, np.cumsum(obj_size))])
        R = out.shape[1]
        return out, nit.reshape(len(obj_size), R), npt.reshape(len(obj_size), R), m, vc, cen, {"kernel_ms": 0.0}

    monkeypatch.setattr(S, "_run_one", fake_run_one)
    cat = make_catalogue([900, 300, 1500, 700, 1100, 450, 2000], seed=4, with_velocities=True, shuffle_particles=True)
    rr = np.logspace(-1.2, 0, 6)
    args = (cat["xyz"], cat["vxyz"], cat["masses"], cat["r200"], cat["idx_cat"], cat["obj_size"], 0.0, rr, 1e-2, 100,
            10, "com", True, False)
    monkeypatch.setattr(session, "devices", [0])
    ref = S.calcMorphLocalVelDisp(*args)
    ref_info = dict(S.last_info)
    for devs in ([0, 1], [0, 1, 2]):
        calls.clear()
        monkeypatch.setattr(session, "devices", devs)
        got = S.calcMorphLocalVelDisp(*args)
        assert sorted(c[0] for c in calls) == devs and sum(c[1] for c in calls) == 7
        for a, b in zip(got, ref):
            assert np.array_equal(a, b, equal_nan=True)
        assert np.array_equal(S.last_info["n_iter"], ref_info["n_iter"])
        assert np.array_equal(S.last_info["n_pts"], ref_info["n_pts"])


def test_weak_scaling_tables_share_sizes():
    """Every rank draws its own shapes and orientations but the same object sizes (fixed per-GPU work)."""
    import numpy as np
    import bench
    t0 = bench.halo_table(500, bench.rank_seed(11, 0), size_seed=12)
    t3 = bench.halo_table(500, bench.rank_seed(11, 3), size_seed=12)
    assert np.array_equal(t0["sizes"], t3["sizes"])
    assert not np.array_equal(t0["q"], t3["q"]) and not np.array_equal(t0["centre"], t3["centre"])
    assert np.array_equal(bench.halo_table(500, 11)["sizes"], t0["sizes"])       # rank 0 == the single-GPU catalogue
