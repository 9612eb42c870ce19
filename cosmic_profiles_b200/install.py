"""Rebind the reference's eight hot-path driver names to the GPU implementations.

cosmic_profiles/common/cosmic_base_class.pyx:11-12 imports the drivers by name and looks them up as module
globals at call time (call sites :80, :155, :232, :324, :404, :825, :863, :887), so replacing those
attributes redirects DensShapeProfs / DensShapeProfsGadget .getShapeCatLocal / getShapeCatGlobal /
getShapeCatVel* / estDensProfs / fitDensProfs to this package without touching the reference.
"""
DRIVER_NAMES = ("calcMorphLocal", "calcMorphGlobal", "calcMorphLocalVelDisp", "calcMorphGlobalVelDisp",
                "calcMassesCenters", "calcDensProfsSphDirectBinning", "calcDensProfsEllDirectBinning",
                "calcDensProfsKernelBased")
_saved = {}


def _implementations():
    from . import dens_profs_algos as d, shape_profs_algos as s
    return {n: getattr(s, n, None) or getattr(d, n) for n in DRIVER_NAMES}


def install(target=None):
    """target: the module whose globals hold the driver names; default
    ``cosmic_profiles.common.cosmic_base_class`` (must be importable)."""
    from . import _lib
    _lib.load()   # fail now, loudly, if the CUDA library is missing
    if target is None:
        import importlib
        target = importlib.import_module("cosmic_profiles.common.cosmic_base_class")
    impl = _implementations()
    for n in DRIVER_NAMES:
        key = (n, id(target))   # per target module: a second install(target=other) records that module's originals too
        if key not in _saved:
            _saved[key] = (target, getattr(target, n, None))
        setattr(target, n, impl[n])
    return target


def install_fits(target=None):
    """Rebind ``fitDensProfHelper`` (dens_profs/dens_profs_tools.py:325) where fitDensProfs / estConcentrations look
    it up: a module global of cosmic_base_class (:10, call sites :780, :800)."""
    from . import _lib, dens_profs_tools
    _lib.load()
    if target is None:
        import importlib
        target = importlib.import_module("cosmic_profiles.common.cosmic_base_class")
    key = ("fitDensProfHelper", id(target))
    if key not in _saved:
        _saved[key] = (target, getattr(target, "fitDensProfHelper", None))
    setattr(target, "fitDensProfHelper", dens_profs_tools.fitDensProfHelper)
    return target


def install_gadget(targets=None):
    """Rebind ``calcObjCat`` (gadget/gen_catalogues.pyx:39) where DensShapeProfsGadget.__init__ looks it up: the
    module globals of shape_profs_classes (:24, call :580) and dens_profs_classes (:10).  targets: modules to patch
    (default: those two, which must be importable)."""
    from . import _lib, gen_catalogues
    _lib.load()
    if targets is None:
        import importlib
        targets = [importlib.import_module(m) for m in gen_catalogues.TARGET_MODULES]
    for t in targets:
        key = ("calcObjCat", id(t))
        if key not in _saved:
            _saved[key] = (t, getattr(t, "calcObjCat", None))
        setattr(t, "calcObjCat", gen_catalogues.calcObjCat)
    return targets


def uninstall():
    for key, (target, fn) in list(_saved.items()):
        if fn is not None:
            setattr(target, key[0], fn)
        del _saved[key]
