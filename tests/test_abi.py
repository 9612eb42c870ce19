"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/cosmic_gpu.h declares, refuses to run without a device, and install() rebinds the 8 driver names."""
import os
import re
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cosmic_gpu.h")).read()
    return sorted(set(re.findall(r"CPG_API\s+[\w\s\*]+?\b(cpg_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cosmic_profiles_b200 import _lib
    L = _lib.load()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), "libcosmicgpu.so does not export " + n
    assert sorted(_lib.EXPORTS) == names
    assert b"sm_100a" in L.cpg_version()


def test_no_cpu_fallback():
    from cosmic_profiles_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(_lib.CosmicGpuError, match="no CPU fallback"):
        _lib.Context(0)
    import numpy as np
    from cosmic_profiles_b200 import shape_profs_algos as S
    with pytest.raises(_lib.CosmicGpuError):
        S.calcMorphGlobal(np.zeros((20, 3)), np.ones(20), np.ones(1), np.arange(20, dtype=np.int32),
                          np.array([20], dtype=np.int32), 0.0, 1e-2, 100, 10, "com", 6.0, False)


def test_install_rebinds_the_eight_drivers():
    import cosmic_profiles_b200 as cpb
    target = types.SimpleNamespace(**{n: ("ref", n) for n in cpb.DRIVER_NAMES})
    cpb.install(target)
    for n in cpb.DRIVER_NAMES:
        fn = getattr(target, n)
        assert callable(fn) and fn.__module__.startswith("cosmic_profiles_b200."), n
    cpb.uninstall()
    for n in cpb.DRIVER_NAMES:
        assert getattr(target, n) == ("ref", n)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "cosmic_profiles_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), fn
                assert "cp_oracle" not in txt and "oracle/_ref" not in txt, fn


def test_lpt_partition_balances():
    import numpy as np
    from cosmic_profiles_b200.partition import lpt_partition, sub_catalogue
    from cosmic_profiles_b200.mock import power_law_sizes
    sizes = power_law_sizes(2000, seed=3)
    parts = lpt_partition(sizes, 8)
    loads = np.array([sizes[p].sum() for p in parts])
    assert sorted(np.concatenate(parts).tolist()) == list(range(2000))
    assert loads.max() / loads.mean() < 1.02
    idx = np.arange(int(sizes.sum()), dtype=np.int32)
    sub_idx, sub_size = sub_catalogue(idx, sizes, parts[3])
    assert sub_size.sum() == len(sub_idx) == sizes[parts[3]].sum()


def test_install_gadget_rebinds_calcObjCat():
    import cosmic_profiles_b200 as cpb
    a, b = types.SimpleNamespace(calcObjCat="ref"), types.SimpleNamespace(calcObjCat="ref")
    cpb.install_gadget([a, b])
    assert a.calcObjCat is b.calcObjCat and a.calcObjCat.__module__ == "cosmic_profiles_b200.gen_catalogues"
    cpb.uninstall()
    assert a.calcObjCat == "ref" and b.calcObjCat == "ref"
