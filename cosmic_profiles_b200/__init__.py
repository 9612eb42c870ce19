"""cosmic_profiles_b200 -- B200 (sm_100a) drop-in for the per-halo shape / density hot path of
tibordome/cosmic_profiles.  See DESIGN.md and INTEGRATION.md.

    import cosmic_profiles_b200 as cpb
    cpb.install()          # rebinds the 8 driver names inside cosmic_profiles.common.cosmic_base_class
    cpb.install_fits()     # rebinds fitDensProfHelper (fitDensProfs / estConcentrations: batched Powell fits)
    cpb.install_gadget()   # rebinds calcObjCat for DensShapeProfsGadget (FoF/SUBFIND group tables -> index catalogue)
"""
from .install import install, install_fits, install_gadget, uninstall, DRIVER_NAMES  # noqa: F401

__all__ = ["install", "install_fits", "install_gadget", "uninstall", "DRIVER_NAMES"]
