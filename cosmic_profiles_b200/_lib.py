"""ctypes binding of libcosmicgpu.so (include/cosmic_gpu.h).  No CPU fallback: a missing library or a
missing CUDA device raises, loudly."""
import ctypes
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COSMICGPU_LIB", os.path.join(HERE, "libcosmicgpu.so"))

_D = ctypes.POINTER(ctypes.c_double)
_I = ctypes.POINTER(ctypes.c_int32)
_lib = None


class CosmicGpuError(RuntimeError):
    pass


class Timing(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_double), ("gather_ms", ctypes.c_double), ("kernel_ms", ctypes.c_double),
                ("d2h_ms", ctypes.c_double), ("kernel_launches", ctypes.c_int32),
                ("hot_kernel_launches", ctypes.c_int32), ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64),
                ("tile_visits", ctypes.c_int64), ("tile_shell_visits", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["cpg_device_count", "cpg_last_error", "cpg_version", "cpg_create", "cpg_destroy", "cpg_set_particles",
           "cpg_set_catalogue", "cpg_shape", "cpg_dens_sph", "cpg_dens_ell", "cpg_dens_kernel", "cpg_obj_masses",
           "cpg_get_timing", "cpg_set_option", "cpg_host_alloc", "cpg_host_free", "cpg_get_prefix_lengths", "cpg_centers",
           "cpg_obj_cat", "cpg_obj_cat_fetch", "cpg_use_obj_cat", "cpg_set_particles_gadget", "cpg_get_particles", "cpg_fit_profiles", "cpg_shape_interp",
           "cpg_set_particles_segments", "cpg_set_catalogue_ranges", "cpg_catalogue_ranges", "cpg_host_token",
           "cpg_measure_fp64_peak", "cpg_partition_lpt", "cpg_multi_create", "cpg_multi_destroy", "cpg_multi_devices",
           "cpg_multi_context", "cpg_multi_set_option", "cpg_multi_set_snapshot", "cpg_multi_shard", "cpg_multi_shape",
           "cpg_multi_get_info", "cpg_get_stat", "cpg_set_snapshot", "cpg_unpack_shape"]
_I64 = ctypes.POINTER(ctypes.c_int64)
_NO_INT_RESTYPE = ("cpg_last_error", "cpg_version", "cpg_destroy", "cpg_host_free", "cpg_multi_destroy", "cpg_multi_context")


def load():
    """dlopen the library and declare every prototype of include/cosmic_gpu.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CosmicGpuError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                             "`make -C cosmic_profiles_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    L.cpg_device_count.restype = ctypes.c_int
    L.cpg_last_error.restype = ctypes.c_char_p
    L.cpg_version.restype = ctypes.c_char_p
    L.cpg_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.cpg_destroy.argtypes = [vp]
    L.cpg_destroy.restype = None
    L.cpg_set_particles.argtypes = [vp, _D, _D, _D, ctypes.c_int64]
    L.cpg_set_catalogue.argtypes = [vp, _I, ctypes.c_int64, _I, ctypes.c_int32]
    L.cpg_shape.argtypes = [vp, _D, _D, _D, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.c_float,
                            ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                            ctypes.c_int32, _D, _I, _I, _D, _D]
    L.cpg_dens_sph.argtypes = [vp, _D, _D, _D, ctypes.c_int32, ctypes.c_float, _D, _I]
    L.cpg_dens_ell.argtypes = [vp, _D, _D, _D, _D, _D, _D, _D, ctypes.c_int32, ctypes.c_float, _D, _I]
    L.cpg_dens_kernel.argtypes = [vp, _D, _D, _D, ctypes.c_int32, ctypes.c_float, _D]
    L.cpg_obj_masses.argtypes = [vp, _D]
    L.cpg_get_timing.argtypes = [vp, ctypes.POINTER(Timing)]
    L.cpg_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
    L.cpg_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    L.cpg_host_free.argtypes = [vp]
    L.cpg_host_free.restype = None
    L.cpg_get_prefix_lengths.argtypes = [vp, _I, ctypes.c_int32]
    L.cpg_centers.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, _D, _D]
    L.cpg_obj_cat.argtypes = [vp, _I, _I, _I, _D, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                              ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64)]
    L.cpg_obj_cat_fetch.argtypes = [vp, _I, _D, _I]
    L.cpg_use_obj_cat.argtypes = [vp]
    L.cpg_set_particles_gadget.argtypes = [vp, vp, ctypes.c_int32, vp, ctypes.c_int32, vp, ctypes.c_int32, ctypes.c_double,
                                           ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_double]
    L.cpg_get_particles.argtypes = [vp, _D, _D, _D]
    L.cpg_shape_interp.argtypes = [vp, _D, _D, _D, ctypes.c_int32, _D, _D, _D, _D, _D, _D, ctypes.c_int32, _D, _D, _D, _D, _D,
                                   _D]
    L.cpg_fit_profiles.argtypes = [vp, _D, _D, _I, _D, _D, _D, _D, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                   ctypes.c_double, ctypes.c_double, _D, _D, _I, _I]
    L.cpg_set_particles_segments.argtypes = [vp, _D, _D, _D, _I64, _I64, ctypes.c_int64]
    L.cpg_set_catalogue_ranges.argtypes = [vp, _I64, _I, ctypes.c_int32]
    L.cpg_catalogue_ranges.argtypes = [_I, ctypes.c_int64, _I, ctypes.c_int32, _I64]
    L.cpg_host_token.argtypes = [vp, ctypes.c_size_t, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64)]
    L.cpg_measure_fp64_peak.argtypes = [vp, _D]
    L.cpg_partition_lpt.argtypes = [_D, ctypes.c_int32, ctypes.c_int32, _I]
    L.cpg_multi_create.argtypes = [_I, ctypes.c_int32, ctypes.POINTER(vp)]
    L.cpg_multi_destroy.argtypes = [vp]
    L.cpg_multi_destroy.restype = None
    L.cpg_multi_devices.argtypes = [vp]
    L.cpg_multi_context.argtypes = [vp, ctypes.c_int32]
    L.cpg_multi_context.restype = vp
    L.cpg_multi_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
    L.cpg_multi_set_snapshot.argtypes = [vp, _D, _D, _D, ctypes.c_int64, _I, ctypes.c_int64, _I, ctypes.c_int32]
    L.cpg_multi_shard.argtypes = [vp, ctypes.c_int32, _I, _I64, _I]
    L.cpg_multi_shape.argtypes = [vp, _D, ctypes.c_int32, _D, _D, _D, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                  ctypes.c_float, ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                  ctypes.c_int32, ctypes.c_int32, _D, _I, _I, _D, _D, _D]
    L.cpg_get_stat.argtypes = [vp, ctypes.c_char_p, _I64]
    L.cpg_set_snapshot.argtypes = [vp, _D, _D, _D, ctypes.c_int64, _I, ctypes.c_int64, _I, ctypes.c_int32]
    L.cpg_unpack_shape.argtypes = [_D, ctypes.c_int64, _D, _D, _D, _D, _D, _D]
    L.cpg_multi_get_info.argtypes = [vp, ctypes.c_int32, _I, _I, _I64, ctypes.POINTER(Timing), _D]
    for f in EXPORTS:
        if f not in _NO_INT_RESTYPE:
            getattr(L, f).restype = ctypes.c_int
    _lib = L
    return L


def device_count():
    return int(load().cpg_device_count())


def _dp(a):
    return None if a is None else a.ctypes.data_as(_D)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_I)


def _check(rc, what):
    if rc != 0:
        raise CosmicGpuError("%s failed (%d): %s" % (what, rc, load().cpg_last_error().decode()))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def partition_lpt(cost, n_parts):
    """owner[h] = part of object h (cpg_partition_lpt: greedy longest-processing-time on `cost`)."""
    cost = f64(np.asarray(cost))
    owner = np.zeros(cost.shape[0], dtype=np.int32)
    _check(load().cpg_partition_lpt(_dp(cost), cost.shape[0], int(n_parts), _ip(owner)), "cpg_partition_lpt")
    return owner


def catalogue_ranges(idx_cat, obj_size):
    """(contiguous?, obj_start int64 (n_obj,)): whether every object of the flat catalogue is a contiguous ascending
    run of snapshot indices (cpg_catalogue_ranges)."""
    idx_cat, obj_size = i32(np.asarray(idx_cat)), i32(np.asarray(obj_size))
    start = np.zeros(obj_size.shape[0], dtype=np.int64)
    rc = load().cpg_catalogue_ranges(_ip(idx_cat), idx_cat.shape[0], _ip(obj_size), obj_size.shape[0],
                                     start.ctypes.data_as(_I64))
    if rc < 0:
        _check(rc, "cpg_catalogue_ranges")
    return bool(rc), start


def host_token(a, full=False):
    a = np.ascontiguousarray(a)
    tok = ctypes.c_uint64()
    _check(load().cpg_host_token(ctypes.c_void_p(a.ctypes.data), a.nbytes, int(bool(full)), ctypes.byref(tok)), "cpg_host_token")
    return int(tok.value)


def unpack_shape(out12):
    """(n, R, 12) rows of cpg_shape -> d, q, s (n, R) and major, inter, minor (n, R, 3), zeros -> NaN (cpg_unpack_shape)."""
    out12 = np.ascontiguousarray(out12, dtype=np.float64)
    n, R = out12.shape[:2]
    d, q, s = (np.empty((n, R)) for _ in range(3))
    major, inter, minor = (np.empty((n, R, 3)) for _ in range(3))
    _check(load().cpg_unpack_shape(_dp(out12), n * R, _dp(d), _dp(q), _dp(s), _dp(major), _dp(inter), _dp(minor)), "cpg_unpack_shape")
    return d, q, s, major, inter, minor


class PinnedBuffer:
    """Grow-only page-locked host scratch (cpg_host_alloc) handed out as numpy views."""

    def __init__(self):
        self.ptr = ctypes.c_void_p()
        self.cap = 0

    def view(self, shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if nbytes > self.cap:
            self.release()
            cap = max(nbytes, 1 << 16)
            _check(load().cpg_host_alloc(cap, ctypes.byref(self.ptr)), "cpg_host_alloc")
            self.cap = cap
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(self.ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def release(self):
        if self.ptr:
            load().cpg_host_free(self.ptr)
            self.ptr = ctypes.c_void_p()
            self.cap = 0


class Context:
    """One CUDA device with a resident snapshot + catalogue (cpg_ctx).  One caller at a time: `lock` is held by the
    drop-in drivers for a whole set-snapshot -> compute -> copy-out sequence."""

    def __init__(self, device=0, _borrowed=None):
        self._L = load()
        self._owned = _borrowed is None
        if _borrowed is None:
            self._h = ctypes.c_void_p()
            _check(self._L.cpg_create(int(device), ctypes.byref(self._h)), "cpg_create")
        else:   # a slot of a cpg_multi handle, destroyed with it
            self._h = ctypes.c_void_p(_borrowed)
        self.device = int(device)
        self.n_obj = 0
        self.have_vel = False
        self.lock = threading.RLock()
        self._pin = [PinnedBuffer() for _ in range(3)]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            if self._owned:
                self._L.cpg_destroy(self._h)   # drains the context's streams before anything is freed
            self._h = ctypes.c_void_p()
            for b in self._pin:
                b.release()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _check(self._L.cpg_set_option(self._h, name.encode(), int(value)), "cpg_set_option")

    def set_particles(self, xyz, vxyz, masses):
        xyz = f64(xyz)
        masses = f64(masses)
        vxyz = None if vxyz is None else f64(vxyz)
        if xyz.ndim != 2 or xyz.shape[1] != 3 or masses.shape[0] != xyz.shape[0]:
            raise ValueError("xyz must be (N,3) and masses (N,)")
        if vxyz is not None and vxyz.shape != xyz.shape:
            raise ValueError("vxyz must match xyz")
        _check(self._L.cpg_set_particles(self._h, _dp(xyz), _dp(vxyz), _dp(masses), xyz.shape[0]), "cpg_set_particles")
        self.have_vel = vxyz is not None

    def set_catalogue(self, idx_cat, obj_size):
        idx_cat = i32(idx_cat)
        obj_size = i32(obj_size)
        _check(self._L.cpg_set_catalogue(self._h, _ip(idx_cat), idx_cat.shape[0], _ip(obj_size), obj_size.shape[0]),
               "cpg_set_catalogue")
        self.n_obj = int(obj_size.shape[0])

    def set_snapshot(self, xyz, vxyz, masses, idx_cat, obj_size):
        """set_particles + set_catalogue in one call (cpg_set_snapshot): the host-side pass over the flat catalogue
        overlaps the transfer of the particle arrays."""
        xyz, masses = f64(xyz), f64(masses)
        vxyz = None if vxyz is None else f64(vxyz)
        idx_cat, obj_size = i32(idx_cat), i32(obj_size)
        if xyz.ndim != 2 or xyz.shape[1] != 3 or masses.shape[0] != xyz.shape[0]:
            raise ValueError("xyz must be (N,3) and masses (N,)")
        if vxyz is not None and vxyz.shape != xyz.shape:
            raise ValueError("vxyz must match xyz")
        _check(self._L.cpg_set_snapshot(self._h, _dp(xyz), _dp(vxyz), _dp(masses), xyz.shape[0], _ip(idx_cat), idx_cat.shape[0],
                                        _ip(obj_size), obj_size.shape[0]), "cpg_set_snapshot")
        self.have_vel = vxyz is not None
        self.n_obj = int(obj_size.shape[0])

    def set_catalogue_ranges(self, obj_start, obj_size):
        """Objects as contiguous snapshot ranges [obj_start[h], obj_start[h] + obj_size[h]): no flat index array."""
        obj_start = np.ascontiguousarray(obj_start, dtype=np.int64)
        obj_size = i32(obj_size)
        if obj_start.shape != obj_size.shape:
            raise ValueError("obj_start and obj_size must have the same length")
        _check(self._L.cpg_set_catalogue_ranges(self._h, obj_start.ctypes.data_as(_I64), _ip(obj_size),
                                                obj_size.shape[0]), "cpg_set_catalogue_ranges")
        self.n_obj = int(obj_size.shape[0])

    def set_particles_segments(self, xyz, vxyz, masses, seg_start, seg_len):
        """Snapshot := concatenation of the particle ranges [seg_start[i], seg_start[i] + seg_len[i]) of the host arrays."""
        xyz, masses = f64(xyz), f64(masses)
        vxyz = None if vxyz is None else f64(vxyz)
        seg_start = np.ascontiguousarray(seg_start, dtype=np.int64)
        seg_len = np.ascontiguousarray(seg_len, dtype=np.int64)
        if seg_start.shape != seg_len.shape or seg_start.ndim != 1:
            raise ValueError("seg_start and seg_len must be 1-d arrays of the same length")
        if len(seg_start) and (int((seg_start + seg_len).max()) > xyz.shape[0] or masses.shape[0] != xyz.shape[0]):
            raise ValueError("a segment reaches past the end of the particle arrays")
        _check(self._L.cpg_set_particles_segments(self._h, _dp(xyz), _dp(vxyz), _dp(masses), seg_start.ctypes.data_as(_I64),
                                                  seg_len.ctypes.data_as(_I64), seg_start.shape[0]),
               "cpg_set_particles_segments")
        self.have_vel = vxyz is not None

    def stat(self, name):
        v = ctypes.c_int64()
        _check(self._L.cpg_get_stat(self._h, name.encode(), ctypes.byref(v)), "cpg_get_stat")
        return int(v.value)

    def measure_fp64_peak(self):
        """Measured fp64 multiply-add rate of the device in 1e12 FMA/s (cpg_measure_fp64_peak)."""
        v = ctypes.c_double()
        _check(self._L.cpg_measure_fp64_peak(self._h, ctypes.byref(v)), "cpg_measure_fp64_peak")
        return float(v.value)

    def shape(self, centers, r200, r_over_r200, local, SAFE, L_BOX, IT_TOL, IT_WALL, IT_MIN, reduced, shell_based,
              use_vel, zero_copy=False):
        """Returns (out12, n_iter, n_pts, obj_mass, vcenters) as fresh arrays.  zero_copy=True returns out12 / n_iter /
        n_pts as views of the context's page-locked result scratch instead: valid only until the next shape() call
        on this context or its close() (bench.py uses that to keep the copy out of its timed region)."""
        n = self.n_obj
        centers = f64(centers)
        r200 = f64(r200)
        rr = f64(r_over_r200) if local else None
        R = int(rr.shape[0]) if local else 1
        # results land in page-locked scratch owned by this context: the returned out/nit/npt are views that
        # stay valid until the next shape() call (the drop-in drivers copy out of them right away)
        out = self._pin[0].view((n, R, 12), np.float64)
        nit = self._pin[1].view((n, R), np.int32)
        npt = self._pin[2].view((n, R), np.int32)
        m = np.zeros(n)
        vc = np.zeros((n, 3))
        _check(self._L.cpg_shape(self._h, _dp(centers), _dp(r200), _dp(rr), R, int(bool(local)), float(SAFE),
                                 float(L_BOX), float(IT_TOL), int(IT_WALL), int(IT_MIN), int(bool(reduced)),
                                 int(bool(shell_based)), int(bool(use_vel)), _dp(out), _ip(nit), _ip(npt), _dp(m),
                                 _dp(vc)), "cpg_shape")
        if not zero_copy:
            out, nit, npt = out.copy(), nit.copy(), npt.copy()
        return out, nit, npt, m, vc

    def dens_sph(self, centers, r200, bin_edges, L_BOX):
        n, R = self.n_obj, len(bin_edges) - 1
        dens = np.zeros((n, R))
        cnt = np.zeros((n, R), dtype=np.int32)
        _check(self._L.cpg_dens_sph(self._h, _dp(f64(centers)), _dp(f64(r200)), _dp(f64(bin_edges)), R, float(L_BOX),
                                    _dp(dens), _ip(cnt)), "cpg_dens_sph")
        return dens, cnt

    def shape_interp(self, r200, r_over_r200, bin_edges, a, b, c, major, inter, minor, fetch=False):
        """Interpolate the shape profiles onto the density bins on the device (cpg_shape_interp).  The result stays
        resident for dens_ell(centers, None, ...); fetch=True also returns (ai, bi, ci, mj, it, mn)."""
        n, R, Rs = self.n_obj, len(r_over_r200), a.shape[1]
        out = [None] * 6
        if fetch:
            out = [np.zeros((n, R + 1)) for _ in range(3)] + [np.zeros((n, R, 3)) for _ in range(3)]
        _check(self._L.cpg_shape_interp(self._h, _dp(f64(r200)), _dp(f64(r_over_r200)), _dp(f64(bin_edges)), R, _dp(f64(a)),
                                        _dp(f64(b)), _dp(f64(c)), _dp(f64(major)), _dp(f64(inter)), _dp(f64(minor)), Rs,
                                        *[_dp(o) for o in out]), "cpg_shape_interp")
        self._interp_R = R
        return tuple(out) if fetch else None

    def dens_ell(self, centers, a, b, c, major, inter, minor, L_BOX):
        if a is None:   # profiles interpolated on the device by shape_interp()
            n, R = self.n_obj, self._interp_R
            dens = np.zeros((n, R))
            cnt = np.zeros((n, R), dtype=np.int32)
            _check(self._L.cpg_dens_ell(self._h, _dp(f64(centers)), None, None, None, None, None, None, R, float(L_BOX),
                                        _dp(dens), _ip(cnt)), "cpg_dens_ell")
            return dens, cnt
        n, R = self.n_obj, major.shape[1]
        dens = np.zeros((n, R))
        cnt = np.zeros((n, R), dtype=np.int32)
        _check(self._L.cpg_dens_ell(self._h, _dp(f64(centers)), _dp(f64(a)), _dp(f64(b)), _dp(f64(c)), _dp(f64(major)),
                                    _dp(f64(inter)), _dp(f64(minor)), R, float(L_BOX), _dp(dens), _ip(cnt)),
               "cpg_dens_ell")
        return dens, cnt

    def dens_kernel(self, centers, r200, r_over_r200, L_BOX):
        n, R = self.n_obj, len(r_over_r200)
        dens = np.zeros((n, R))
        _check(self._L.cpg_dens_kernel(self._h, _dp(f64(centers)), _dp(f64(r200)), _dp(f64(r_over_r200)), R,
                                       float(L_BOX), _dp(dens)), "cpg_dens_kernel")
        return dens

    def obj_masses(self):
        m = np.zeros(self.n_obj)
        _check(self._L.cpg_obj_masses(self._h, _dp(m)), "cpg_obj_masses")
        return m

    def centers(self, mode, shape_path, L_BOX, ref_points=None):
        cen = np.zeros((self.n_obj, 3))
        ref = None if ref_points is None else f64(ref_points)
        _check(self._L.cpg_centers(self._h, int(mode), int(bool(shape_path)), float(L_BOX), _dp(ref), _dp(cen)),
               "cpg_centers")
        return cen

    def obj_cat(self, nb_shs, sh_len, fof_sizes, group_r200, min_number_ptcs, fetch_idx=True):
        """calcObjCat on the device (gadget/gen_catalogues.pyx:39-95): returns (idx_cat or None, obj_r200, obj_size);
        the result also stays on the device for use_obj_cat()."""
        nb_shs, sh_len, fof_sizes = i32(nb_shs), i32(sh_len), i32(fof_sizes)
        group_r200 = f64(group_r200)
        if not (nb_shs.shape == fof_sizes.shape == group_r200.shape) or nb_shs.ndim != 1 or sh_len.ndim != 1:
            raise ValueError("nb_shs, fof_sizes, group_r200 must be (N1,) and sh_len (N2,)")
        n_obj, n_idx = ctypes.c_int32(), ctypes.c_int64()
        _check(self._L.cpg_obj_cat(self._h, _ip(nb_shs), _ip(sh_len), _ip(fof_sizes), _dp(group_r200), nb_shs.shape[0],
                                   sh_len.shape[0], int(min_number_ptcs), ctypes.byref(n_obj), ctypes.byref(n_idx)),
               "cpg_obj_cat")
        idx = np.empty(n_idx.value, dtype=np.int32) if fetch_idx else None
        r200 = np.empty(n_obj.value, dtype=np.float64)
        size = np.empty(n_obj.value, dtype=np.int32)
        _check(self._L.cpg_obj_cat_fetch(self._h, _ip(idx), _dp(r200), _ip(size)), "cpg_obj_cat_fetch")
        self._oc_n_obj = int(n_obj.value)
        return idx, r200, size

    def use_obj_cat(self):
        """Make the catalogue built by obj_cat() this context's catalogue (device-resident, no idx_cat upload)."""
        _check(self._L.cpg_use_obj_cat(self._h), "cpg_use_obj_cat")
        self.n_obj = self._oc_n_obj

    def set_particles_gadget(self, pos, vel, mass, mass_table, length_unit_fac, mass_unit_fac, velocity_unit_fac, time):
        """Raw snapshot blocks (float32 or float64, as stored on disk) -> fp64 device arrays, with the arithmetic of
        readgadget.read_field (gadget/readgadget.py:76-124)."""
        def prep(a, cols):
            if a is None:
                return None, 0, None
            a = np.asarray(a)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            a = np.ascontiguousarray(a)
            if (cols == 3 and (a.ndim != 2 or a.shape[1] != 3)) or (cols == 1 and a.ndim != 1):
                raise ValueError("positions/velocities must be (N,3), masses (N,)")
            return a, int(a.dtype == np.float32), ctypes.c_void_p(a.ctypes.data)
        pos, pf, pp = prep(pos, 3)
        vel, vf, vp_ = prep(vel, 3)
        mass, mf, mp = prep(mass, 1)
        n = pos.shape[0]
        if (vel is not None and vel.shape[0] != n) or (mass is not None and mass.shape[0] != n):
            raise ValueError("block lengths differ")
        _check(self._L.cpg_set_particles_gadget(self._h, pp, pf, vp_, vf, mp, mf, float(mass_table), n,
                                                float(length_unit_fac), float(mass_unit_fac), float(velocity_unit_fac),
                                                float(time)), "cpg_set_particles_gadget")
        self.have_vel = vel is not None
        self.n_part = n

    def get_particles(self, n_part, with_vel=False):
        xyz = np.empty((n_part, 3))
        m = np.empty(n_part)
        v = np.empty((n_part, 3)) if with_vel else None
        _check(self._L.cpg_get_particles(self._h, _dp(xyz), _dp(v), _dp(m)), "cpg_get_particles")
        return xyz, v, m

    def fit_profiles(self, radii, lmed, n_valid, avg, x0, lb, ub, profile, xtol=1e-4, ftol=1e-4):
        """Batched Powell fits (cpg_fit_profiles); returns (best (n,npar), fun, nfev, nit)."""
        radii, lmed, avg, x0, lb, ub = f64(radii), f64(lmed), f64(avg), f64(x0), f64(lb), f64(ub)
        n_valid = i32(n_valid)
        n, R = radii.shape
        npar = x0.shape[1]
        best = np.zeros((n, npar))
        fun = np.zeros(n)
        nfev = np.zeros(n, dtype=np.int32)
        nit = np.zeros(n, dtype=np.int32)
        _check(self._L.cpg_fit_profiles(self._h, _dp(radii), _dp(lmed), _ip(n_valid), _dp(avg), _dp(x0), _dp(lb), _dp(ub),
                                        n, R, int(profile), float(xtol), float(ftol), _dp(best), _dp(fun), _ip(nfev),
                                        _ip(nit)), "cpg_fit_profiles")
        return best, fun, nfev, nit

    def prefix_lengths(self, R):
        ne = np.zeros((self.n_obj, R), dtype=np.int32)
        _check(self._L.cpg_get_prefix_lengths(self._h, _ip(ne), int(R)), "cpg_get_prefix_lengths")
        return ne

    def timing(self):
        t = Timing()
        _check(self._L.cpg_get_timing(self._h, ctypes.byref(t)), "cpg_get_timing")
        return t.as_dict()



class MultiContext:
    """Several devices behind one handle (cpg_multi): the objects of ONE catalogue are dealt to the devices by a
    particle-count-balanced LPT partition; every device receives only its own objects' particles."""

    def __init__(self, devices):
        devices = [int(d) for d in devices]
        self._L = load()
        self._h = ctypes.c_void_p()
        arr = np.asarray(devices, dtype=np.int32)
        _check(self._L.cpg_multi_create(_ip(arr), len(devices), ctypes.byref(self._h)), "cpg_multi_create")
        self.devices = devices
        self.contexts = [Context(d, _borrowed=self._L.cpg_multi_context(self._h, k)) for k, d in enumerate(devices)]
        self.shards = [np.zeros(0, dtype=np.int32) for _ in devices]
        self.n_obj = 0
        self.lock = threading.RLock()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            for c in self.contexts:
                c.close()
            self._L.cpg_multi_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _check(self._L.cpg_multi_set_option(self._h, name.encode(), int(value)), "cpg_multi_set_option")

    def set_snapshot(self, xyz, vxyz, masses, idx_cat, obj_size):
        xyz, masses = f64(xyz), f64(masses)
        vxyz = None if vxyz is None else f64(vxyz)
        idx_cat, obj_size = i32(idx_cat), i32(obj_size)
        if xyz.ndim != 2 or xyz.shape[1] != 3 or masses.shape[0] != xyz.shape[0]:
            raise ValueError("xyz must be (N,3) and masses (N,)")
        if vxyz is not None and vxyz.shape != xyz.shape:
            raise ValueError("vxyz must match xyz")
        _check(self._L.cpg_multi_set_snapshot(self._h, _dp(xyz), _dp(vxyz), _dp(masses), xyz.shape[0], _ip(idx_cat),
                                              idx_cat.shape[0], _ip(obj_size), obj_size.shape[0]), "cpg_multi_set_snapshot")
        self.n_obj = int(obj_size.shape[0])
        for k, c in enumerate(self.contexts):
            nd, npart = ctypes.c_int32(), ctypes.c_int64()
            _check(self._L.cpg_multi_shard(self._h, k, ctypes.byref(nd), ctypes.byref(npart), None), "cpg_multi_shard")
            objs = np.zeros(nd.value, dtype=np.int32)
            _check(self._L.cpg_multi_shard(self._h, k, None, None, _ip(objs)), "cpg_multi_shard")
            self.shards[k] = objs
            c.n_obj = int(nd.value)
            c.have_vel = vxyz is not None

    def shape(self, centers, center_mode, ref_points, r200, r_over_r200, local, SAFE, L_BOX, IT_TOL, IT_WALL, IT_MIN,
              reduced, shell_based, use_vel):
        """cpg_multi_shape.  centers given (center_mode = -1) or computed per device (0 'com', 1 'mode').
        Returns (out12, n_iter, n_pts, obj_mass, vcenters, centers)."""
        n = self.n_obj
        r200 = f64(r200)
        rr = f64(r_over_r200) if local else None
        R = int(rr.shape[0]) if local else 1
        cen_in = None if centers is None else f64(centers)
        ref = None if ref_points is None else f64(ref_points)
        out = np.zeros((n, R, 12))
        nit = np.full((n, R), -1, dtype=np.int32)
        npt = np.full((n, R), -1, dtype=np.int32)
        m = np.zeros(n)
        vc = np.zeros((n, 3))
        cen_out = np.zeros((n, 3))
        _check(self._L.cpg_multi_shape(self._h, _dp(cen_in), int(center_mode), _dp(ref), _dp(r200), _dp(rr), R,
                                       int(bool(local)), float(SAFE), float(L_BOX), float(IT_TOL), int(IT_WALL),
                                       int(IT_MIN), int(bool(reduced)), int(bool(shell_based)), int(bool(use_vel)),
                                       _dp(out), _ip(nit), _ip(npt), _dp(m), _dp(vc), _dp(cen_out)), "cpg_multi_shape")
        return out, nit, npt, m, vc, (cen_in if center_mode < 0 else cen_out)

    def info(self):
        """Per slot: device, objects, resident particles, timing of its last call; plus the host-side times (ms)."""
        per, host = [], None
        for k in range(len(self.contexts)):
            dev, nd, npart, t = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64(), Timing()
            hm = np.zeros(4)
            _check(self._L.cpg_multi_get_info(self._h, k, ctypes.byref(dev), ctypes.byref(nd), ctypes.byref(npart),
                                              ctypes.byref(t), _dp(hm)), "cpg_multi_get_info")
            per.append(dict(device=int(dev.value), n_obj=int(nd.value), n_part=int(npart.value), timing=t.as_dict()))
            host = dict(partition_ms=float(hm[0]), upload_ms=float(hm[1]), scatter_ms=float(hm[2]), call_ms=float(hm[3]))
        return dict(devices=per, host=host)


class DeviceGroup:
    """What the drop-in drivers talk to: one device (a Context) or several (a MultiContext) behind the same few
    operations.  rows(fn) runs fn(ctx, objs) for every shard on its own host thread (objs = catalogue rows of the
    shard, None for the single-device case = all rows in order) and scatters the per-shard result arrays into
    catalogue order."""

    def __init__(self, devices):
        self.devices = [int(d) for d in devices]
        self.multi = MultiContext(self.devices) if len(self.devices) > 1 else None
        self.single = Context(self.devices[0]) if self.multi is None else None
        self.lock = self.multi.lock if self.multi else self.single.lock
        self.n_obj = 0

    @property
    def contexts(self):
        return self.multi.contexts if self.multi else [self.single]

    def close(self):
        (self.multi or self.single).close()

    def set_option(self, name, value):
        (self.multi or self.single).set_option(name, value)

    def set_snapshot(self, xyz, vxyz, masses, idx_cat, obj_size):
        if self.multi:
            self.multi.set_snapshot(xyz, vxyz, masses, idx_cat, obj_size)
        else:
            self.single.set_snapshot(xyz, vxyz, masses, idx_cat, obj_size)
        self.n_obj = int(np.asarray(obj_size).shape[0])

    def shards(self):
        if self.multi:
            return [(c, o) for c, o in zip(self.multi.contexts, self.multi.shards) if len(o)]
        return [(self.single, None)]

    def rows(self, fn):
        """fn(ctx, objs) -> array or tuple of arrays whose first axis runs over the shard's objects."""
        shards = self.shards()
        if len(shards) == 1 and shards[0][1] is None:
            return fn(self.single, None)
        res, errs = [None] * len(shards), []

        def work(k):
            try:
                res[k] = fn(*shards[k])
            except Exception as e:   # re-raised in the caller's thread
                errs.append(e)
        th = [threading.Thread(target=work, args=(k,)) for k in range(len(shards))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        first = res[0]
        many = isinstance(first, tuple)
        outs = []
        for j in range(len(first) if many else 1):
            a0 = first[j] if many else first
            full = np.zeros((self.n_obj,) + a0.shape[1:], dtype=a0.dtype)
            for (ctx, objs), r in zip(shards, res):
                full[objs] = r[j] if many else r
            outs.append(full)
        return tuple(outs) if many else outs[0]

    def timings(self):
        return [c.timing() for c, _ in self.shards()]
