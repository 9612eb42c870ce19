"""Device ingest of Gadget snapshot blocks: the arithmetic of readgadget.read_field / read_block
(cosmic_profiles/gadget/readgadget.py:76-124, :127-194) on arrays as they are stored on disk.

The reference turns every block into a float64 host array before anything else happens
(``array = f[prefix+suffix][:]/unit_fac``, velocities ``*= np.sqrt(head.time)``, then assignment into a float64
array).  Gadget snapshots usually store float32 Coordinates / Velocities and no Masses dataset at all for dark
matter (MassTable constant), so the float64 host copy doubles (positions, velocities) or creates (masses) the bytes
that must cross PCIe.  `upload_snapshot` sends the raw blocks and does the unit division, the sqrt(time) factor and
the widening on the device, bit-identical to the NumPy (>= 2, NEP 50) expressions above.

File parsing (h5py / binary formats, multi-file loops) stays with the reference's reader: this module takes the
datasets it yields.
"""
import numpy as np


def unit_factors(in_length_cm, in_mass_g, in_velocity_cm_s, little_h):
    """l/m/v_target_over_curr of readgadget.py:82-85; internal units Mpc/h, 1e10 Msun/h, km/s
    (common/config.py:10-19)."""
    return (3.085678e24 / little_h) / in_length_cm, (1.989e43 / little_h) / in_mass_g, 1e5 / in_velocity_cm_s


def upload_snapshot(ctx, coordinates, velocities=None, masses=None, mass_table=0.0, length_unit_fac=1.0,
                    mass_unit_fac=1.0, velocity_unit_fac=1.0, time=1.0):
    """coordinates (N,3), velocities (N,3) or None, masses (N,) or None (then mass_table = header MassTable entry).
    dtype float32 or float64 as stored in the file."""
    ctx.set_particles_gadget(coordinates, velocities, masses, mass_table, length_unit_fac, mass_unit_fac,
                             velocity_unit_fac, time)
