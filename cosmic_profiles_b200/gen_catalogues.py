"""Drop-in for cosmic_profiles.gadget.gen_catalogues.calcObjCat (gadget/gen_catalogues.pyx:39-95) on the device
(SURVEY.md section 8(f) next-2).

The reference builds the central-subhalo (CSH) catalogue of a FoF/SUBFIND group table with two O(n^2) prefix sums
inside a prange (:69-70, :82-84) and one np.hstack per object (:87-89); here it is two device-wide exclusive scans,
a compaction and a fill (csrc/cpg_gadget.cuh).  Same positional signature, same return tuple
``(obj_cat int32 (N3,), obj_r200 float64 (n,), obj_size int32 (n,))``.  The reference's docstring says the indices
are 'true index + 1'; its code (calcCSHIdxs, :29-35) stores the true index, and so does this.
"""
import numpy as np

from . import session

NAMES = ("calcObjCat",)
# the two reference modules that bind calcObjCat by name (shape_profs_classes.pyx:24, dens_profs_classes.pyx:10)
TARGET_MODULES = ("cosmic_profiles.shape_profs.shape_profs_classes", "cosmic_profiles.dens_profs.dens_profs_classes")


def calcObjCat(nb_shs, sh_len, fof_sizes, group_r200, MIN_NUMBER_PTCS):
    nb_shs = np.asarray(nb_shs)
    if nb_shs.shape[0] == 0:
        # the reference reaches np.max of an empty array (:76)
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    # one device: the scans are a millisecond of work; the catalogue is sharded later, with its particles
    with session.locked_group() as grp:
        idx, r200, size = grp.contexts[0].obj_cat(nb_shs, sh_len, fof_sizes, group_r200, int(MIN_NUMBER_PTCS))
    return idx, r200, size


def device_catalogue(ctx, nb_shs, sh_len, fof_sizes, group_r200, MIN_NUMBER_PTCS):
    """Build the CSH catalogue on `ctx`'s device and make it the context's catalogue without the flat index array
    crossing PCIe in either direction.  Returns (obj_r200, obj_size)."""
    _, r200, size = ctx.obj_cat(nb_shs, sh_len, fof_sizes, group_r200, int(MIN_NUMBER_PTCS), fetch_idx=False)
    ctx.use_obj_cat()
    return r200, size
