"""CPU tests of the N>1 plumbing (there is no data-path collective: objects are sharded, results are
gathered on the host; bench.py only reduces timings).
 * world_size-2 gloo run of bench.reduce_over_ranks (time = max over ranks, work = sum);
 * world_size-2 gloo run of bench's partition of ONE catalogue over the ranks;
 * the library's LPT partition against its plain-Python statement;
 * the shim's scatter of per-shard results into catalogue order.
The sharded execution itself (cpg_multi) is tested on the device: tests/test_gpu_multi.py."""
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


def test_library_partition_equals_the_python_statement():
    """cpg_partition_lpt (what cpg_multi and bench.py shard with) against the plain-Python statement of the same
    greedy rule in cosmic_profiles_b200/partition.py, on a power-law size table and on ties."""
    from cosmic_profiles_b200 import _lib
    from cosmic_profiles_b200.mock import power_law_sizes
    from cosmic_profiles_b200.partition import lpt_partition
    for sizes, parts in ((power_law_sizes(3000, seed=5), 8), (np.array([5, 3, 3, 2, 2, 2, 1]), 3), (np.ones(17), 4),
                         (np.array([7.0]), 3)):
        owner = _lib.partition_lpt(np.asarray(sizes, dtype=np.float64), parts)
        ref = lpt_partition(sizes, parts)
        for p in range(parts):
            assert np.array_equal(np.flatnonzero(owner == p), ref[p])
        loads = np.array([np.asarray(sizes)[owner == p].sum() for p in range(parts)])
        if len(sizes) >= 100:
            assert loads.max() / loads.mean() < 1.02


def test_device_group_rows_scatter_in_catalogue_order():
    """DeviceGroup.rows: per-shard results (one host thread per shard) land in the catalogue rows of the shard."""
    from cosmic_profiles_b200 import _lib

    class FakeCtx:
        def __init__(self, tag):
            self.tag = tag

    grp = _lib.DeviceGroup.__new__(_lib.DeviceGroup)
    grp.devices = [0, 1, 2]
    grp.single = None
    grp.n_obj = 10
    owner = np.array([0, 1, 2, 1, 0, 0, 2, 1, 0, 2])
    grp.multi = type("M", (), {})()
    grp.multi.contexts = [FakeCtx(k) for k in range(3)]
    grp.multi.shards = [np.flatnonzero(owner == k).astype(np.int32) for k in range(3)]
    val = np.arange(10) * 10.0

    def fn(ctx, objs):
        return val[objs] + ctx.tag, np.stack([objs, objs], axis=1)
    a, b = grp.rows(fn)
    assert np.array_equal(a, val + owner) and np.array_equal(b[:, 0], np.arange(10)) and b.shape == (10, 2)
    one = grp.rows(lambda ctx, objs: val[objs])
    assert np.array_equal(one, val)

    def boom(ctx, objs):
        raise RuntimeError("shard %d failed" % ctx.tag)
    with pytest.raises(RuntimeError, match="failed"):
        grp.rows(boom)


def _partition_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        tab = bench.halo_table(bench.global_halos(400, world), 9)
        mine = bench.rank_objects(tab["sizes"], rank, world)
        q.put((rank, mine.tolist(), int(tab["sizes"][mine].sum()), int(tab["sizes"].sum())))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_partition_of_one_catalogue():
    """Under torchrun every rank derives the same global object table and takes its own shard of the library's LPT
    partition: shards are disjoint, cover the catalogue and balance the particle counts."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_partition_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = got
    assert not set(a[1]) & set(b[1]) and sorted(a[1] + b[1]) == list(range(800))
    assert a[3] == b[3] == a[2] + b[2]
    assert abs(a[2] - b[2]) / a[3] < 0.01


def test_weak_scaling_tables_share_sizes():
    """Every rank draws its own shapes and orientations but the same object sizes (fixed per-GPU work)."""
    import numpy as np
    import bench
    t0 = bench.halo_table(500, bench.rank_seed(11, 0), size_seed=12)
    t3 = bench.halo_table(500, bench.rank_seed(11, 3), size_seed=12)
    assert np.array_equal(t0["sizes"], t3["sizes"])
    assert not np.array_equal(t0["q"], t3["q"]) and not np.array_equal(t0["centre"], t3["centre"])
    assert np.array_equal(bench.halo_table(500, 11)["sizes"], t0["sizes"])       # rank 0 == the single-GPU catalogue
