#!/usr/bin/env python3
"""Build the UNMODIFIED reference (and a minimally patched twin) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cosmic_profiles_b200/`` may import or
call what this script produces; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs do.

What it does:

* copies ``/root/reference/cosmic_profiles`` (minus ``for_docs``/``tests``) to a
  scratch directory under ``/tmp`` (the reference tree is read-only),
* runs ``cython`` on the seven ``.pyx`` files and compiles each generated ``.c``
  with ``/usr/bin/gcc -O2 -fopenmp`` -- the same flags the reference's
  ``setup_compile.py:37-140`` ends up with (python's default ``-O2`` plus
  ``-fopenmp``); the reference's own build system is NOT run,
* installs ONLY the built ``.so`` files plus the package's ``.py`` modules
  (needed at import) into ``oracle/_ref/<flavour>/cosmic_profiles`` -- a
  git-ignored install target, like ``pip install --target``; no reference
  source enters the git history.

Flavours
--------
``pristine``  the reference exactly as shipped.
``patched``   three documented one-line defect fixes + instrumentation that
              exports the integers the parity gate needs:
  (i)   ``dens_profs_algos.pyx:170`` overwrites each object's index slice with
        one scalar -> ellipsoidal binning returns all-NaN.  Fix: delete line.
  (ii)  ``calcMorphLocalVelDisp`` never fills ``vxyz_obj``
        (``shape_profs_algos.pyx:820-826``) -> 0/0.  Fix: fill it like
        ``calcMorphGlobalVelDisp`` does at ``:939``.
  (iii) instrumentation: ``dbg_setup(it, pts, R)`` registers two int32 arrays;
        every ``run*Algo`` exit records the number of shape-tensor evaluations
        and the particle count used by the last evaluation (or the count that
        failed the ``IT_MIN`` test) for (object, shell).

``/root/reference`` only exists in the build container; on the GPU box the
prebuilt ``oracle/_ref`` travels with the snapshot and this script is a no-op
(``--if-missing``).
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("CP_REFERENCE_ROOT", "/root/reference")
OUT_ROOT = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")
GCC = "/usr/bin/gcc"  # $CC=/opt/gcc/bin/gcc cannot find libgomp.spec

PYX = [
    "cosmic_profiles/cython_helpers/helper_class.pyx",
    "cosmic_profiles/gadget/gen_catalogues.pyx",
    "cosmic_profiles/dens_profs/dens_profs_algos.pyx",
    "cosmic_profiles/shape_profs/shape_profs_algos.pyx",
    "cosmic_profiles/common/cosmic_base_class.pyx",
    "cosmic_profiles/dens_profs/dens_profs_classes.pyx",
    "cosmic_profiles/shape_profs/shape_profs_classes.pyx",
]

DBG_HEADER = '''
# ---- instrumentation added by oracle/build_ref.py (patched flavour only) ----
cdef int _dbg_p[4096]
cdef int _dbg_k[4096]
cdef int* _dbg_iter = NULL
cdef int* _dbg_pts = NULL
cdef int _dbg_R = 1
_dbg_keep = []

def dbg_setup(int[::1] it, int[::1] pts, int R):
    global _dbg_iter, _dbg_pts, _dbg_R
    _dbg_keep[:] = [it, pts]
    _dbg_iter = &it[0]
    _dbg_pts = &pts[0]
    _dbg_R = R

def dbg_clear():
    global _dbg_iter, _dbg_pts
    _dbg_iter = NULL
    _dbg_pts = NULL
    _dbg_keep[:] = []

cdef void _dbg_rec(int n_iter, int n_pts) noexcept nogil:
    cdef int t = openmp.omp_get_thread_num()
    if _dbg_iter != NULL:
        _dbg_iter[_dbg_p[t]*_dbg_R + _dbg_k[t]] = n_iter
        _dbg_pts[_dbg_p[t]*_dbg_R + _dbg_k[t]] = n_pts
# ---- end instrumentation ----
'''


def _must_replace(src, old, new, count=None, what=""):
    n = src.count(old)
    if n == 0 or (count is not None and n != count):
        raise RuntimeError("patch anchor %r found %d times (expected %s)" % (what or old[:50], n, count))
    return src.replace(old, new)


def patch_shape_algos(src):
    # header after the imports
    src = _must_replace(src, "from libc.math cimport sqrt\n", "from libc.math cimport sqrt\n" + DBG_HEADER, 1, "imports")
    # split into top-level function chunks so that each patch is local
    parts = re.split(r"(?m)^(?=@cython\.embedsignature\(True\)\n)", src)
    out = []
    for chunk in parts:
        m = re.search(r"^(?:cdef [^\n]*? |def )(\w+)\(", chunk, re.M)
        name = m.group(1) if m else ""
        if name in ("runShellAlgo", "runEllAlgo", "runEllVDispAlgo", "runShellVDispAlgo"):
            var = "pts_in_shell" if "pts_in_shell" in chunk else "pts_in_ell"
            chunk = _must_replace(chunk, "    cdef int iteration = 1\n",
                                  "    cdef int iteration = 1\n    cdef int dbg_last = 0\n", 1, name + ":decl")
            chunk = _must_replace(chunk,
                                  "            morph_info[:] = 0.0\n            return morph_info\n",
                                  "            morph_info[:] = 0.0\n            _dbg_rec(iteration-1, %s)\n            return morph_info\n" % var,
                                  2, name + ":fail")
            chunk = _must_replace(chunk, "        shape_tensor = CythonHelpers.calcShapeTensor(",
                                  "        dbg_last = %s\n        shape_tensor = CythonHelpers.calcShapeTensor(" % var,
                                  1, name + ":tensor")
            chunk = _must_replace(chunk, "        iteration += 1\n    return morph_info\n",
                                  "        iteration += 1\n    _dbg_rec(iteration-1, dbg_last)\n    return morph_info\n",
                                  1, name + ":exit")
        elif name in ("calcObjMorphLocal", "calcObjMorphLocalVelDisp"):
            chunk = _must_replace(chunk, "    for shell_nb in range(nb_shells):\n",
                                  "    for shell_nb in range(nb_shells):\n        _dbg_k[openmp.omp_get_thread_num()] = shell_nb\n",
                                  1, name + ":k")
        elif name in ("calcMorphLocal", "calcMorphGlobal", "calcMorphLocalVelDisp", "calcMorphGlobalVelDisp"):
            chunk = _must_replace(chunk, "    for p in prange(nb_objs, schedule = 'dynamic', nogil = True):\n",
                                  "    for p in prange(nb_objs, schedule = 'dynamic', nogil = True):\n"
                                  "        _dbg_p[openmp.omp_get_thread_num()] = p\n"
                                  "        _dbg_k[openmp.omp_get_thread_num()] = 0\n",
                                  1, name + ":p")
            if name == "calcMorphLocalVelDisp":
                # defect (ii): velocities never gathered
                anchor = "            xyz_obj[openmp.omp_get_thread_num(),n,2] = xyz[idx_cat[offsets[p]+n],2]\n"
                fill = "".join(
                    "            vxyz_obj[openmp.omp_get_thread_num(),n,%d] = vxyz[idx_cat[offsets[p]+n],%d]\n" % (c, c)
                    for c in range(3))
                chunk = _must_replace(chunk, anchor, anchor + fill, 1, name + ":vfill")
        out.append(chunk)
    return "".join(out)


def patch_dens_algos(src):
    # defect (i)
    return _must_replace(src, "        idx_cat.base[offsets[p]:offsets[p+1]] = np.array(idx_cat[p])\n", "", 1, "dens:170")


def run(cmd, cwd, env=None):
    r = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:])
        raise RuntimeError("command failed: %s" % " ".join(cmd))
    return r.stdout


def build(flavour, jobs=1):
    import numpy as np
    import Cython
    import scipy

    work = "/tmp/cp_ref_build_%s" % flavour
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    shutil.copytree(os.path.join(REF_ROOT, "cosmic_profiles"), os.path.join(work, "cosmic_profiles"),
                    ignore=shutil.ignore_patterns("for_docs", "tests", "*.so", "*.c", "__pycache__"))
    if flavour == "patched":
        for rel, fn in (("cosmic_profiles/shape_profs/shape_profs_algos.pyx", patch_shape_algos),
                        ("cosmic_profiles/dens_profs/dens_profs_algos.pyx", patch_dens_algos)):
            p = os.path.join(work, rel)
            with open(p) as f:
                s = f.read()
            with open(p, "w") as f:
                f.write(fn(s))

    env = dict(os.environ)
    env["PYTHONPATH"] = STUBS + os.pathsep + work
    inc = [sysconfig.get_config_var("INCLUDEPY"), np.get_include(), work]
    cflags = ["-fno-strict-overflow", "-DNDEBUG", "-g0", "-O2", "-fwrapv", "-fPIC", "-fopenmp", "-w"]
    procs = []
    for rel in PYX:
        run([sys.executable, "-m", "cython", "-3", "-I", work, rel], cwd=work, env=env)
    for rel in PYX:
        c = rel[:-4] + ".c"
        so = rel[:-4] + ".so"
        cmd = [GCC] + cflags + ["-I" + i for i in inc] + [c, "-shared", "-o", so, "-fopenmp", "-lm"]
        procs.append((rel, subprocess.Popen(cmd, cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        if len(procs) >= jobs:
            rel0, p0 = procs.pop(0)
            o, _ = p0.communicate()
            if p0.returncode:
                sys.stderr.write(o[-4000:])
                raise RuntimeError("gcc failed on " + rel0)
    for rel0, p0 in procs:
        o, _ = p0.communicate()
        if p0.returncode:
            sys.stderr.write(o[-4000:])
            raise RuntimeError("gcc failed on " + rel0)

    dst = os.path.join(OUT_ROOT, flavour)
    shutil.rmtree(dst, ignore_errors=True)
    for root, dirs, files in os.walk(os.path.join(work, "cosmic_profiles")):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for fn in files:
            if fn.endswith((".py", ".so")):
                relp = os.path.relpath(os.path.join(root, fn), work)
                os.makedirs(os.path.dirname(os.path.join(dst, relp)), exist_ok=True)
                shutil.copy2(os.path.join(root, fn), os.path.join(dst, relp))
    info = {
        "flavour": flavour,
        "reference_root": REF_ROOT,
        "cython": Cython.__version__,
        "numpy": np.__version__,
        "scipy": scipy.__version__,
        "gcc": run([GCC, "--version"], cwd=work).splitlines()[0],
        "cflags": " ".join(cflags),
        "python": sys.version.split()[0],
    }
    with open(os.path.join(dst, "BUILD_INFO.json"), "w") as f:
        json.dump(info, f, indent=1)
    shutil.rmtree(work, ignore_errors=True)
    return dst


def have(flavour):
    return os.path.exists(os.path.join(OUT_ROOT, flavour, "BUILD_INFO.json"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flavours", default="pristine,patched")
    ap.add_argument("--if-missing", action="store_true")
    ap.add_argument("--jobs", type=int, default=4)
    a = ap.parse_args()
    for fl in a.flavours.split(","):
        if a.if_missing and have(fl):
            print("oracle/_ref/%s: present" % fl)
            continue
        if not os.path.isdir(REF_ROOT):
            print("oracle/_ref/%s: reference tree %s absent, cannot build (prebuilt copy travels with the snapshot)" % (fl, REF_ROOT))
            continue
        print("building oracle/_ref/%s ..." % fl, flush=True)
        print("  ->", build(fl, a.jobs))


if __name__ == "__main__":
    main()
