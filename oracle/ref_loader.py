"""Loader for the compiled reference under oracle/_ref (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  One flavour per process (both are called ``cosmic_profiles``).
"""
import contextlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def available(flavour="pristine"):
    return os.path.exists(os.path.join(HERE, "_ref", flavour, "BUILD_INFO.json"))


def load(flavour="pristine"):
    """Import the compiled reference; returns a namespace with its 8 hot-path entry points."""
    if not available(flavour):
        raise ImportError("oracle/_ref/%s missing: run python oracle/build_ref.py" % flavour)
    if "cosmic_profiles" in sys.modules:
        got = getattr(sys.modules["cosmic_profiles"], "__file__", "") or ""
        if os.path.join("_ref", flavour) not in got:
            raise ImportError("another cosmic_profiles flavour is already imported: " + got)
    for p in (os.path.join(HERE, "_ref", flavour), os.path.join(HERE, "stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    ns = types.SimpleNamespace()
    from cosmic_profiles.shape_profs import shape_profs_algos as spa
    from cosmic_profiles.dens_profs import dens_profs_algos as dpa
    from cosmic_profiles.common import python_routines as pr
    ns.shape = spa
    ns.dens = dpa
    ns.routines = pr
    ns.flavour = flavour
    for n in ("calcMorphLocal", "calcMorphGlobal", "calcMorphLocalVelDisp", "calcMorphGlobalVelDisp"):
        setattr(ns, n, getattr(spa, n))
    for n in ("calcMassesCenters", "calcDensProfsSphDirectBinning", "calcDensProfsEllDirectBinning",
              "calcDensProfsKernelBased"):
        setattr(ns, n, getattr(dpa, n))
    return ns


@contextlib.contextmanager
def quiet_stdout():
    """The reference printf()s one line per object from C; silence fd 1 around a call."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
