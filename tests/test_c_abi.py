"""The boundary is a C ABI: a plain-C program (tests/c_abi/c_abi_smoke.c) includes include/cosmic_gpu.h, links
libcosmicgpu.so, and its output is checked against the oracle.  CPU: it compiles and links (every prototype it uses
exists with the declared signature).  GPU: it runs."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = os.path.join(ROOT, "tests", "c_abi", "c_abi_smoke.c")


def _build(tmp):
    exe = os.path.join(str(tmp), "c_abi_smoke")
    lib = os.path.join(ROOT, "cosmic_profiles_b200")
    cmd = ["/usr/bin/gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", lib, "-l:libcosmicgpu.so", "-Wl,-rpath," + lib, "-lm"]
    subprocess.run(cmd, check=True)
    return exe


def test_c_program_compiles_and_links(tmp_path):
    if not os.path.exists(os.path.join(ROOT, "cosmic_profiles_b200", "libcosmicgpu.so")):
        pytest.skip("libcosmicgpu.so not built")
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_c_program_matches_the_oracle(tmp_path):
    from oracle import cp_oracle as O
    exe = _build(tmp_path)
    n = 20000
    res = subprocess.run([exe, str(n)], check=True, capture_output=True, text=True).stdout.split()
    assert int(res[0]) == n
    q, s, n_iter, n_pts, mass = float(res[1]), float(res[2]), int(res[3]), int(res[4]), float(res[5])
    counts = np.array([int(v) for v in res[6:10]])
    dens = np.array([float(v) for v in res[10:14]])
    # the same particles, regenerated with the C program's LCG
    state = 12345
    mask = (1 << 64) - 1

    def u01():
        nonlocal state
        state = (state * 6364136223846793005 + 1442695040888963407) & mask
        return ((state >> 11) & ((1 << 53) - 1)) / 9007199254740992.0
    xyz = np.zeros((n, 3))
    for i in range(n):
        while True:
            x, y, z = 2 * u01() - 1, 2 * u01() - 1, 2 * u01() - 1
            if x * x + y * y + z * z < 1.0:
                break
        xyz[i] = 50.0 + x, 50.0 + 0.7 * y, 50.0 + 0.4 * z
    m = np.full(n, 0.01)
    idx = np.arange(n, dtype=np.int32)
    size = np.array([n], dtype=np.int32)
    cen = np.array([[50.0, 50.0, 50.0]])
    ref, (nit, npt, _, _) = O.morph(xyz, None, m, np.ones(1), idx, size, 0.0, None, 1e-2, 100, 10, "com", False, False, False,
                                    SAFE=6.0, centers=cen)
    assert n_iter == int(nit[0]) and n_pts == int(npt[0])
    assert abs(q / ref[1][0] - 1) < 1e-9 and abs(s / ref[2][0] - 1) < 1e-9
    assert abs(mass / m.sum() - 1) < 1e-12
    assert 0.65 < q < 0.75 and 0.35 < s < 0.45          # the generating axis ratios
    r = np.linalg.norm(xyz - cen, axis=1)
    edges = np.array([1e-8, 0.25, 0.5, 0.75, 1.0])
    want_counts = np.array([np.count_nonzero((r * r < edges[k + 1] ** 2) & (r * r >= edges[k] ** 2)) for k in range(4)])
    assert np.array_equal(counts, want_counts)
    vol = 4.0 / 3.0 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
    assert np.allclose(dens, want_counts * 0.01 / vol, rtol=1e-12)
