"""Gadget / FoF ingest row: calcObjCat (gadget/gen_catalogues.pyx:39-95) and the block
arithmetic of readgadget.read_field (gadget/readgadget.py:106-120).

CPU: the numpy restatement against golden vectors produced by the compiled reference.
GPU: the CUDA path (through the C ABI) against the golden vectors and the restatement -- integers bit-exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cp_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "gadget_golden.npz")


def _cases():
    g = np.load(GOLD)
    for name in g["cases"]:
        name = str(name)
        yield name, {k: g[name + "_" + k] for k in ("nb_shs", "sh_len", "fof_sizes", "group_r200", "min", "obj_cat",
                                                    "obj_r200", "obj_size")}


def _table(rng, n_fof, p_nosub=0.1):
    nb = rng.integers(0, 6, size=n_fof).astype(np.int32)
    nb[rng.uniform(size=n_fof) < p_nosub] = 0
    sh_len = rng.integers(0, 3000, size=int(nb.sum())).astype(np.int32)
    off = np.concatenate(([0], np.cumsum(nb)[:-1]))
    first = np.where(nb > 0, sh_len[np.minimum(off, len(sh_len) - 1)], 0)
    fof = (first + rng.integers(0, 500, size=n_fof)).astype(np.int32)
    return nb, sh_len, fof, rng.uniform(0.05, 2.0, size=n_fof)


def test_oracle_obj_cat_matches_compiled_reference():
    n = 0
    for name, c in _cases():
        cat, r200, size = O.calcObjCat(c["nb_shs"], c["sh_len"], c["fof_sizes"], c["group_r200"], int(c["min"]))
        assert np.array_equal(cat, c["obj_cat"]) and cat.dtype == np.int32, name
        assert np.array_equal(r200, c["obj_r200"]), name
        assert np.array_equal(size, c["obj_size"]) and size.dtype == np.int32, name
        n += 1
    assert n == 5


def test_oracle_read_field_arith_dtypes():
    rng = np.random.default_rng(1)
    x32 = rng.uniform(0, 1e5, size=(1000, 3)).astype(np.float32)
    u = 3.085678e24 / 0.6774 / 3.085678e21
    y = O.read_field_arith("POS ", x32, u)
    assert y.dtype == np.float64
    assert np.array_equal(y, (x32 / np.float32(u)).astype(np.float64))          # fp32 division (NEP 50)
    y64 = O.read_field_arith("POS ", x32.astype(np.float64), u)
    assert np.array_equal(y64, x32.astype(np.float64) / u)
    v = O.read_field_arith("VEL ", x32, 1.0, time=0.37)
    assert np.array_equal(v, (x32.astype(np.float64) * np.sqrt(0.37)).astype(np.float32).astype(np.float64))
    m = O.read_field_arith("MASS", None, 1.989e43 / 0.6774 / 1.989e43, mass_table=0.0123, n_part=10)
    assert np.all(m == 0.0123 / (1.989e43 / 0.6774 / 1.989e43))
    with pytest.raises(Exception):
        O.read_field_arith("MASS", None, 1.0, mass_table=0.0, n_part=10)


def test_drop_in_names():
    from cosmic_profiles_b200 import gen_catalogues as G
    assert G.NAMES == ("calcObjCat",) and callable(G.calcObjCat)
    with pytest.raises(ValueError):   # the reference's np.max of an empty array (gen_catalogues.pyx:76)
        G.calcObjCat(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), 10)


# ------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_gpu_obj_cat_golden():
    from cosmic_profiles_b200 import gen_catalogues as G
    for name, c in _cases():
        cat, r200, size = G.calcObjCat(c["nb_shs"], c["sh_len"], c["fof_sizes"], c["group_r200"], c["min"])
        assert cat.dtype == np.int32 and size.dtype == np.int32 and r200.dtype == np.float64
        assert np.array_equal(cat, c["obj_cat"]), name
        assert np.array_equal(r200, c["obj_r200"]), name
        assert np.array_equal(size, c["obj_size"]), name


@pytest.mark.gpu
@pytest.mark.parametrize("n_fof,minp", [(1, 1), (2047, 50), (2048, 50), (2049, 50), (100000, 200), (700001, 1000)])
def test_gpu_obj_cat_vs_oracle(n_fof, minp):
    """Sizes around the 2048-element scan tile and beyond one top-level scan block; bit-exact."""
    from cosmic_profiles_b200 import gen_catalogues as G
    nb, sl, fof, r200 = _table(np.random.default_rng(n_fof), n_fof)
    if n_fof == 1:
        nb[:] = 1
        sl = np.array([5], dtype=np.int32)
        fof[:] = 9
    want = O.calcObjCat(nb, sl, fof, r200, minp)
    got = G.calcObjCat(nb, sl, fof, r200, np.int32(minp))
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    # properties that hold at any size: sizes sum to the catalogue length, every object is a contiguous range
    assert int(got[2].sum()) == len(got[0])
    if len(got[0]):
        brk = np.flatnonzero(np.diff(got[0]) != 1) + 1
        assert set(brk.tolist()) <= set(np.cumsum(got[2])[:-1].tolist())


@pytest.mark.gpu
def test_gpu_obj_cat_none_pass_and_errors():
    from cosmic_profiles_b200 import _lib, gen_catalogues as G
    nb, sl, fof, r200 = _table(np.random.default_rng(5), 1000)
    cat, o_r200, size = G.calcObjCat(nb, sl, fof, r200, 10 ** 6)
    assert len(cat) == 0 and len(o_r200) == 0 and len(size) == 0
    big = np.full(3, 2 ** 30, dtype=np.int32)          # third group starts at 2^31: int32 catalogue overflows
    with pytest.raises(_lib.CosmicGpuError):
        G.calcObjCat(np.ones(3, np.int32), np.full(3, 100, np.int32), big, np.ones(3), 10)


@pytest.mark.gpu
def test_gpu_device_catalogue_equals_host_catalogue():
    """Shapes computed on the device-built catalogue (no idx_cat upload) == shapes on the host-built one."""
    from cosmic_profiles_b200 import _lib, gen_catalogues as G
    from cosmic_profiles_b200.mock import make_catalogue
    sizes = [1500, 800, 2600, 300, 1200]
    cat = make_catalogue(sizes, seed=11)
    # FoF layout: every group = its central subhalo + some unbound particles that follow it
    rng = np.random.default_rng(3)
    extra = rng.integers(0, 40, size=len(sizes))
    fof = (np.asarray(sizes) + extra).astype(np.int32)
    n_tot = int(fof.sum())
    xyz = np.zeros((n_tot, 3))
    m = np.full(n_tot, 0.01)
    src = np.concatenate(([0], np.cumsum(sizes)))
    dst = np.concatenate(([0], np.cumsum(fof)))
    for h, n in enumerate(sizes):
        xyz[dst[h]:dst[h] + n] = cat["xyz"][cat["idx_cat"][src[h]:src[h + 1]]]
        xyz[dst[h] + n:dst[h + 1]] = cat["xyz"][cat["idx_cat"][src[h]]] + rng.normal(size=(extra[h], 3))
    nb = np.array([2, 1, 3, 1, 1], dtype=np.int32)
    sl = []
    for h, n in enumerate(sizes):
        sl.extend([n] + [3] * (nb[h] - 1))
    sl = np.asarray(sl, dtype=np.int32)
    minp = 500                                         # drops the 300-particle group
    idx, o_r200, o_size = O.calcObjCat(nb, sl, fof, cat["r200"], minp)
    rr = np.logspace(-1.5, 0, 12)
    cen = O.centres(xyz, m, idx, o_size, 0.0, "com", True)
    a = _lib.Context(0)
    a.set_particles(xyz, None, m)
    a.set_catalogue(idx, o_size)
    out_a, nit_a, npt_a, _, _ = a.shape(cen, o_r200, rr, True, 0.0, 0.0, 1e-2, 100, 10, False, False, False)
    out_a, nit_a, npt_a = out_a.copy(), nit_a.copy(), npt_a.copy()
    b = _lib.Context(0)
    b.set_particles(xyz, None, m)
    r2, s2 = G.device_catalogue(b, nb, sl, fof, cat["r200"], minp)
    assert np.array_equal(r2, o_r200) and np.array_equal(s2, o_size) and b.n_obj == len(o_size)
    out_b, nit_b, npt_b, _, _ = b.shape(cen, o_r200, rr, True, 0.0, 0.0, 1e-2, 100, 10, False, False, False)
    assert np.array_equal(nit_a, nit_b) and np.array_equal(npt_a, npt_b) and np.array_equal(out_a, out_b)
    assert (nit_a > 0).any()
    a.close()
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gpu_block_ingest_bit_exact(dt):
    from cosmic_profiles_b200 import _lib, readgadget as RG
    rng = np.random.default_rng(7)
    n = 200003
    pos = rng.uniform(0, 75000.0, size=(n, 3)).astype(dt)          # kpc/h
    vel = rng.normal(0, 300.0, size=(n, 3)).astype(dt)
    mass = rng.uniform(1e-4, 3e-4, size=n).astype(dt)
    lf, mf, vf = RG.unit_factors(3.085678e21 / 0.6774, 1.989e43 / 0.6774, 1e5, 0.6774)
    time = 0.3333
    ctx = _lib.Context(0)
    RG.upload_snapshot(ctx, pos, vel, mass, 0.0, lf, mf, vf, time)
    x, v, m = ctx.get_particles(n, with_vel=True)
    assert np.array_equal(x, O.read_field_arith("POS ", pos, lf))
    assert np.array_equal(v, O.read_field_arith("VEL ", vel, vf, time=time))
    assert np.array_equal(m, O.read_field_arith("MASS", mass, mf))
    RG.upload_snapshot(ctx, pos, None, None, 0.0123, lf, mf, vf, time)
    x, _, m = ctx.get_particles(n)
    assert np.array_equal(m, O.read_field_arith("MASS", None, mf, mass_table=0.0123, n_part=n))
    with pytest.raises(_lib.CosmicGpuError):
        RG.upload_snapshot(ctx, pos, None, None, 0.0, lf, mf, vf, time)
    ctx.close()


@pytest.mark.gpu
def test_gpu_pageable_upload_round_trip():
    """Plain (pageable) numpy arrays of >= 64 MB take the threaded bounce-buffer upload (8 MB pieces, ragged last
    piece), smaller ones and option upload_threads=0 the direct path: what arrives on the device must be the bytes
    that were sent (profiles/r1c_dropin_timing.json: 2.5-4x on the drop-in drivers' wall time)."""
    from cosmic_profiles_b200 import _lib
    rng = np.random.default_rng(3)
    n = 3_000_017
    xyz = rng.normal(size=(n, 3))
    v = rng.normal(size=(n, 3))
    m = rng.uniform(size=n)
    ctx = _lib.Context(0)
    ctx.set_particles(xyz, v, m)
    x2, v2, m2 = ctx.get_particles(n, with_vel=True)
    assert np.array_equal(x2, xyz) and np.array_equal(v2, v) and np.array_equal(m2, m)
    ctx.set_particles(xyz[::-1].copy(), None, m)      # second upload reuses the bounce buffers
    x3, _, _ = ctx.get_particles(n)
    assert np.array_equal(x3, xyz[::-1])
    ctx.set_option("upload_threads", 0)
    ctx.set_particles(xyz, None, m)
    x4, _, _ = ctx.get_particles(n)
    assert np.array_equal(x4, xyz)
    ctx.close()


# ---- the reference's own reader as the oracle of the block arithmetic (tests/golden/make_gadget_io_golden.py) --------

def _io_cases():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gadget_io_golden.npz"))
    return z, [str(c) for c in z["cases"]]


def test_oracle_read_field_arith_matches_the_reference_reader():
    """readgadget.read_block ran unmodified (h5py served by oracle/stubs/h5py) on float32 / float64 snapshots with and
    without a Masses dataset, a != 1, two unit systems, one- and two-file snapshots: the restatement reproduces its
    float64 outputs bit for bit -- this pins the oracle of the ingest row."""
    z, cases = _io_cases()
    assert len(cases) == 4
    for c in cases:
        # Python floats, as the reference's config module holds them: under NumPy >= 2 promotion a Python-float divisor
        # keeps the dataset's precision (a float32 block is divided in float32), an np.float64 one would not
        lf, mf, vf = (float(v) for v in z[c + "/unit_fac"])
        t = float(z[c + "/time"])
        n = z[c + "/pos"].shape[0]
        assert z[c + "/ref_pos"].dtype == np.float64
        assert np.array_equal(O.read_field_arith("POS ", z[c + "/pos"], lf), z[c + "/ref_pos"]), c
        assert np.array_equal(O.read_field_arith("VEL ", z[c + "/vel"], vf, time=t), z[c + "/ref_vel"]), c
        mass = z[c + "/mass"] if (c + "/mass") in z.files else None
        got = O.read_field_arith("MASS", mass, mf, mass_table=float(z[c + "/mass_table"]), n_part=n)
        assert np.array_equal(got, z[c + "/ref_mass"]), c


@pytest.mark.gpu
def test_gpu_block_ingest_matches_the_reference_reader():
    """cpg_set_particles_gadget on the raw datasets against what the reference's reader returned for the same files."""
    from cosmic_profiles_b200 import _lib, readgadget as RG
    z, cases = _io_cases()
    ctx = _lib.Context(0)
    try:
        for c in cases:
            lf, mf, vf = (float(v) for v in z[c + "/unit_fac"])
            mass = z[c + "/mass"] if (c + "/mass") in z.files else None
            n = z[c + "/pos"].shape[0]
            RG.upload_snapshot(ctx, z[c + "/pos"], z[c + "/vel"], mass, float(z[c + "/mass_table"]), lf, mf, vf, float(z[c + "/time"]))
            x, v, m = ctx.get_particles(n, with_vel=True)
            assert np.array_equal(x, z[c + "/ref_pos"]), c
            assert np.array_equal(v, z[c + "/ref_vel"]), c
            assert np.array_equal(m, z[c + "/ref_mass"]), c
    finally:
        ctx.close()
