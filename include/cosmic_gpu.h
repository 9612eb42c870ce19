/*
 * cosmic_gpu.h -- C ABI of libcosmicgpu.so, the B200 (sm_100a) replacement for the
 * per-halo shape / density hot path of tibordome/cosmic_profiles.
 *
 * The reference has no FFI for this path: its orchestration layer calls eight Cython `def`
 * functions by module-global name (cosmic_profiles/common/cosmic_base_class.pyx:11-12, call
 * sites :80,:155,:232,:324,:404,:825,:863,:887).  A ctypes shim
 * (cosmic_profiles_b200/{shape_profs_algos,dens_profs_algos}.py) keeps those eight Python
 * signatures and binds the entry points below; INTEGRATION.md shows the rebind.
 *
 * Conventions: plain pointers and sizes, no exceptions, every function returns 0 on success
 * or a negative cpg_status and leaves a message for cpg_last_error().  All array arguments
 * are HOST pointers owned by the caller (borrowed for the duration of the call, never
 * mutated); results are written into caller-allocated host buffers.  Units are the
 * reference's internal ones (Mpc/h, 1e10 Msun/h, km/s).  One caller thread per context.
 * There is no CPU fallback: without a CUDA device cpg_create fails with CPG_ERR_NO_DEVICE.
 */
#ifndef COSMIC_GPU_H
#define COSMIC_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CPG_API __attribute__((visibility("default")))
#else
#define CPG_API
#endif

typedef enum cpg_status {
    CPG_OK = 0,
    CPG_ERR_NO_DEVICE = -1,   /* no usable CUDA device / driver */
    CPG_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed; see cpg_last_error() */
    CPG_ERR_ARGUMENT = -3,    /* bad argument (NULL, negative size, catalogue out of range ...) */
    CPG_ERR_STATE = -4,       /* call order: particles / catalogue not set */
    CPG_ERR_NOMEM = -5
} cpg_status;

typedef struct cpg_ctx cpg_ctx; /* opaque: one device, one stream, device-resident snapshot + catalogue */

/* Timing of the most recent compute call on a context, measured with CUDA events on the
 * context's own stream (bench.py's roofline numbers come from here). */
typedef struct cpg_timing {
    double h2d_ms;          /* host->device copies issued by the call (centres, radii ...) */
    double gather_ms;       /* catalogue gather + PBC + centre-relative SoA + per-halo statistics kernels */
    double kernel_ms;       /* the hot kernels (morphology loop / binning / kernel estimator) */
    double d2h_ms;          /* device->host copy of the results */
    int32_t kernel_launches;        /* number of library kernels launched by the call */
    int32_t hot_kernel_launches;    /* of which hot kernels (the ones kernel_ms covers) */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    /* cpg_shape only: 128-particle tiles the morphology kernels passed over, and (tile, shell) pairs among them --
     * the fp64 work of the streaming loop is 128 * (6 * tile_visits + 12 * tile_shell_visits) multiply-adds: 6 products
     * per particle, 6 for the membership form and 6 tensor updates per particle and shell (bench.py's fp64 roofline;
     * 6 instead of 12 when the library is built with the fp32 pre-classification) */
    int64_t tile_visits;
    int64_t tile_shell_visits;
} cpg_timing;

CPG_API int cpg_device_count(void);
CPG_API const char *cpg_last_error(void);
CPG_API const char *cpg_version(void);

/* Create / destroy a context bound to CUDA device `device`. */
CPG_API int cpg_create(int device, cpg_ctx **out);
CPG_API void cpg_destroy(cpg_ctx *ctx);

/* Upload the snapshot: xyz (n_part,3) C-contiguous fp64, optional vxyz (n_part,3) or NULL,
 * masses (n_part).  Replaces: the `xyz`, `vxyz`, `masses` arguments of every calcMorph* /
 * calcDensProfs* call (shape_profs_algos.pyx:517,631,739,859; dens_profs_algos.pyx:15,54,112,221). */
CPG_API int cpg_set_particles(cpg_ctx *ctx, const double *xyz, const double *vxyz, const double *masses,
                              int64_t n_part);

/* Upload the index catalogue: idx_cat (n_idx) int32 particle indices, obj_size (n_obj) int32.
 * Replaces the `idx_cat`, `obj_size` arguments and the int32 `offsets` the reference derives from
 * them (shape_profs_algos.pyx:557); offsets are 64-bit here. */
CPG_API int cpg_set_catalogue(cpg_ctx *ctx, const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size,
                              int32_t n_obj);

/* cpg_set_particles + cpg_set_catalogue in one call, the host-side pass over the flat catalogue (contiguity test,
 * index range, content hash) overlapped with the host->device transfer of the particle arrays. */
CPG_API int cpg_set_snapshot(cpg_ctx *ctx, const double *xyz, const double *vxyz, const double *masses, int64_t n_part,
                             const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size, int32_t n_obj);

/* Snapshot as the concatenation of element ranges of the host arrays: particles [seg_start[i], seg_start[i] +
 * seg_len[i]) of xyz / vxyz / masses, i = 0..n_seg-1, become the device snapshot (host threads pack the ranges
 * straight into page-locked bounce buffers).  Used to ship one shard of a halo-partitioned catalogue to its GPU
 * (replaces the fancy-indexed copy xyz[sub_idx] a host-side sharding would need); cpg_multi_set_snapshot builds on it. */
CPG_API int cpg_set_particles_segments(cpg_ctx *ctx, const double *xyz, const double *vxyz, const double *masses,
                                       const int64_t *seg_start, const int64_t *seg_len, int64_t n_seg);

/* Index catalogue whose objects are contiguous snapshot ranges (FoF / SUBFIND order, what calcCSHIdxs produces,
 * gadget/gen_catalogues.pyx:29-35): object h = particles [obj_start[h], obj_start[h] + obj_size[h]).  No flat index
 * array exists on either side (4 bytes per particle less over PCIe and in HBM) and the gather reads the snapshot
 * coalesced.  cpg_set_catalogue recognises such catalogues by itself (option "detect_ranges").
 * cpg_catalogue_ranges is the host-side test it uses: returns 1 and fills obj_start_out (n_obj) when every object of
 * the flat catalogue is a contiguous ascending run, 0 when not, < 0 on bad arguments.  No device needed. */
CPG_API int cpg_set_catalogue_ranges(cpg_ctx *ctx, const int64_t *obj_start, const int32_t *obj_size, int32_t n_obj);
CPG_API int cpg_catalogue_ranges(const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size, int32_t n_obj,
                                 int64_t *obj_start_out);

/* Content token of a host array as the snapshot cache computes it (option "snapshot_cache"): full == 0: first and
 * last 4 KB plus 4096 evenly spaced 64-byte blocks; full != 0: every byte.  No device needed. */
CPG_API int cpg_host_token(const void *p, size_t bytes, int32_t full, uint64_t *token_out);

/*
 * Katz-Dubinski iterative morphology for every object of the catalogue.
 * Replaces calcMorphLocal (shape_profs_algos.pyx:517-628), calcMorphGlobal (:631-736),
 * calcMorphLocalVelDisp (:739-856), calcMorphGlobalVelDisp (:859-972) and everything below them
 * (runEllAlgo :148, runShellAlgo :16, run*VDispAlgo :265,:386, calcObjMorph* :975-1240,
 * CythonHelpers.calcShapeTensor / ZHEEVR / respectPBCNoRef / calcCoM / calcLocalSpread).
 *
 *   centers      (n_obj,3)  object centres (the reference computes them in Python, :586-591)
 *   r200         (n_obj)
 *   r_over_r200  (R)        local != 0: shells at d = r200*r_over_r200[k] (:1035-1044);
 *                           local == 0: R must be 1 and d = r200 + SAFE (:1096), pointer may be NULL
 *   use_vel      != 0: tensor from velocities about vcenter = calcCoM(v, m) (helper_class.pyx:85-105,
 *                fp32 sequential mass normaliser reproduced); selection stays positional
 *   out12        (n_obj,R,12): d, q, s, major xyz, inter xyz, minor xyz; all zeros where the
 *                reference returns zeros (IT_WALL / IT_MIN / r200==0 / degenerate object)
 *   n_iter,n_pts (n_obj,R) int32: shape-tensor evaluations performed; particle count of the last
 *                evaluation (or the count that failed the test); -1 where the per-object exit hit
 *   obj_mass     (n_obj) sum of member masses (:598); may be NULL
 *   vcenters     (n_obj,3) output, only written when use_vel; may be NULL
 */
CPG_API int cpg_shape(cpg_ctx *ctx, const double *centers, const double *r200, const double *r_over_r200, int32_t R,
                      int32_t local, double SAFE, float L_BOX, double IT_TOL, int32_t IT_WALL, int32_t IT_MIN,
                      int32_t reduced, int32_t shell_based, int32_t use_vel, double *out12, int32_t *n_iter,
                      int32_t *n_pts, double *obj_mass, double *vcenters);

/* Spherical direct binning.  Replaces calcDensProfsSphDirectBinning (dens_profs_algos.pyx:54-109) and
 * CythonHelpers.calcDensProfBruteForceSph (helper_class.pyx:204-241).  bin_edges (R+1) in units of r200
 * (built by the shim exactly as :83).  dens (n_obj,R); counts (n_obj,R) int32 may be NULL. */
CPG_API int cpg_dens_sph(cpg_ctx *ctx, const double *centers, const double *r200, const double *bin_edges, int32_t R,
                         float L_BOX, double *dens, int32_t *counts);

/* Ellipsoidal-shell direct binning.  Replaces the prange part of calcDensProfsEllDirectBinning
 * (dens_profs_algos.pyx:196-218) and CythonHelpers.calcDensProfBruteForceEll (helper_class.pyx:246-306).
 * a,b,c (n_obj,R+1) and major,inter,minor (n_obj,R,3) are the scipy-interpolated profiles of :163-195
 * (computed by the shim with the same scipy), or all six NULL to use the profiles cpg_shape_interp left on the
 * device.  Empty bin -> NaN. */
CPG_API int cpg_dens_ell(cpg_ctx *ctx, const double *centers, const double *a, const double *b, const double *c,
                         const double *major, const double *inter, const double *minor, int32_t R, float L_BOX,
                         double *dens, int32_t *counts);

/* Shape-profile interpolation onto the density bins, on the device.  Replaces the
 * per-object loop of calcDensProfsEllDirectBinning that calls scipy.interpolate.interp1d(a[p], y, bounds_error=False,
 * fill_value='extrapolate') twelve times (dens_profs_algos.pyx:163-195): a, b, c (n_obj,Rs) and major, inter, minor
 * (n_obj,Rs,3) are the local shape profiles (NaN where a shell did not converge), r_over_r200 (R) the bin centres and
 * bin_edges (R+1) the edges of :155-156, both in units of r200 (n_obj).  2 <= Rs <= 256.  Bit-identical to scipy's
 * linear interp1d.  The interpolated profiles stay on the device: a following cpg_dens_ell call with all six profile
 * pointers NULL uses them.  ai, bi, ci (n_obj,R+1), mj, it, mn (n_obj,R,3): optional host copies (each may be NULL). */
CPG_API int cpg_shape_interp(cpg_ctx *ctx, const double *r200, const double *r_over_r200, const double *bin_edges,
                             int32_t R, const double *a, const double *b, const double *c, const double *major,
                             const double *inter, const double *minor, int32_t Rs, double *ai, double *bi, double *ci,
                             double *mj, double *it, double *mn);

/* Kernel-based estimator.  Replaces calcDensProfsKernelBased (dens_profs_algos.pyx:221-280) and
 * CythonHelpers.calcKTilde (helper_class.pyx:417-428), mixed fp32/fp64 expression tree reproduced. */
CPG_API int cpg_dens_kernel(cpg_ctx *ctx, const double *centers, const double *r200, const double *r_over_r200,
                            int32_t R, float L_BOX, double *dens);

/* Per-object total mass in catalogue order.  Replaces the prange of calcMassesCenters
 * (dens_profs_algos.pyx:45-47). */
CPG_API int cpg_obj_masses(cpg_ctx *ctx, double *obj_mass);

/* Object centres.  Replaces the serial Python centre loops of the drivers
 * (shape_profs_algos.pyx:586-591, dens_profs_algos.pyx:39-44, :96-101) and python_routines.calcCoM (:530-547),
 * calcMode (:287-306), respectPBCNoRef (:499-528).  mode 0 = 'com', 1 = 'mode'; shape_path selects the initial
 * radius of calcMode (1: bounding-box extent, :589; 0: 1000, dens_profs_algos.pyx:42).  ref_points (n_obj,3):
 * the reference points respectPBCNoRef / calcCoM average from 50 seeded random members -- required when
 * L_BOX != 0 (they decide which particles are translated), optional otherwise.  centers_out (n_obj,3) is the
 * centre BEFORE recentreObject; an empty object yields NaN like the reference's division by zero. */
CPG_API int cpg_centers(cpg_ctx *ctx, int32_t mode, int32_t shape_path, double L_BOX, const double *ref_points,
                        double *centers_out);

/*
 * Gadget / FoF-layout ingest: the two steps DensShapeProfsGadget.__init__ runs on the
 * host before any driver is called (shape_profs_classes.pyx:571-580).
 *
 * cpg_obj_cat: central-subhalo index catalogue from the FoF / SUBFIND group tables.  Replaces calcObjCat and
 * calcCSHIdxs (gadget/gen_catalogues.pyx:39-95, :12-35).  nb_shs (n_fof) subhalos per group, sh_len (n_sh)
 * particles per subhalo, fof_sizes (n_fof) particles per group, group_r200 (n_fof).  A group passes when it has a
 * subhalo and its first one holds >= min_number_ptcs particles (:66-73); its object is the index range
 * [sum_{q<p} fof_sizes[q], + sh_len[sum_{q<p} nb_shs[q]]) (:80-85).  The result stays on the device;
 * *n_obj_out / *n_idx_out give the sizes of the arrays cpg_obj_cat_fetch copies out:
 * idx_cat (n_idx) int32 (may be NULL), obj_r200 (n_obj), obj_size (n_obj) -- calcObjCat's return tuple (:91).
 * cpg_use_obj_cat makes the device-resident catalogue the context's catalogue (what cpg_set_catalogue does
 * with host arrays) without the flat index array ever crossing PCIe.
 */
CPG_API int cpg_obj_cat(cpg_ctx *ctx, const int32_t *nb_shs, const int32_t *sh_len, const int32_t *fof_sizes,
                        const double *group_r200, int32_t n_fof, int64_t n_sh, int32_t min_number_ptcs,
                        int32_t *n_obj_out, int64_t *n_idx_out);
CPG_API int cpg_obj_cat_fetch(cpg_ctx *ctx, int32_t *idx_cat, double *obj_r200, int32_t *obj_size);
CPG_API int cpg_use_obj_cat(cpg_ctx *ctx);

/*
 * Upload the snapshot blocks as stored in the Gadget file and convert them on the device to the fp64 arrays
 * readgadget.read_block returns (gadget/readgadget.py:76-124, :127-194): pos / vel (n_part,3) and mass (n_part) are
 * float32 (*_is_f32 != 0) or float64 datasets; array = dataset / unit_fac in the dataset's own precision (NumPy
 * promotion of a Python-float divisor, :112), velocities then *= sqrt(time) (:116), all widened exactly to fp64
 * (:169).  mass == NULL: constant MassTable entry, np.ones(N) * mass_table / mass_unit_fac (:108).  vel may be NULL.
 * unit_fac = internal unit / input unit as computed at readgadget.py:82-85.  Replaces cpg_set_particles for
 * snapshots that are float32 on disk: 12 instead of 24 bytes per particle and block cross PCIe.
 * cpg_get_particles copies the converted fp64 arrays back (tests, or callers that also need them on the host).
 */
CPG_API int cpg_set_particles_gadget(cpg_ctx *ctx, const void *pos, int32_t pos_is_f32, const void *vel,
                                     int32_t vel_is_f32, const void *mass, int32_t mass_is_f32, double mass_table,
                                     int64_t n_part, double length_unit_fac, double mass_unit_fac,
                                     double velocity_unit_fac, double time);
CPG_API int cpg_get_particles(cpg_ctx *ctx, double *xyz, double *vxyz, double *masses);

/*
 * Batched density-profile fits.  Replaces fitDensProfHelper / fitDensProf /
 * toMinimize and the four profile models (dens_profs/dens_profs_tools.py:325-359, :232-323, :227-230, :22-79):
 * per object, scipy.optimize.minimize(method='Powell', bounds=...) of psi^2 = sum_i (ln med_i - ln model(r_i))^2 / n
 * (scipy 1.18.1 _minimize_powell + _linesearch_powell + _minimize_scalar_bounded, default options), one warp per
 * object.  profile 0 nfw, 1 hernquist (rho_s, r_s), 2 einasto (rho_s, alpha, r_s), 3 alpha_beta_gamma
 * (rho_s, alpha, beta, gamma, r_s).  Inputs as the reference prepares them in NumPy (:254-258, :277):
 * radii (n_obj,R) = ROverR200 * r200 and lmed (n_obj,R) = log(median / average(median)), NaN bins removed and the
 * n_valid[p] kept values stored first; avg (n_obj) = average(median); x0 (n_obj,npar) initial guesses; lb / ub
 * (npar) bounds, ub = +inf where free (every lb finite).  R <= 1024 (NumPy's pairwise summation order restated: blocks of 128, halved recursively).
 * Outputs: best (n_obj,npar) with best[:,0] multiplied by avg (:320); fun (n_obj) final psi^2, nfev, nit (n_obj)
 * function evaluations / Powell iterations as scipy counts them (any may be NULL except best).
 */
CPG_API int cpg_fit_profiles(cpg_ctx *ctx, const double *radii, const double *lmed, const int32_t *n_valid,
                             const double *avg, const double *x0, const double *lb, const double *ub, int32_t n_obj,
                             int32_t R, int32_t profile, double xtol, double ftol, double *best, double *fun,
                             int32_t *nfev, int32_t *nit);

/* Host-side helper (no device): the (n_rows,12) out12 rows of cpg_shape split into the six arrays the reference's drivers
 * return -- d, q, s (n_rows) and major, inter, minor (n_rows,3) -- with the reference's zeros -> NaN masking
 * (shape_profs_algos.pyx:616-624: d, q, s and every eigenvector component that is exactly 0.0). */
CPG_API int cpg_unpack_shape(const double *out12, int64_t n_rows, double *d, double *q, double *s, double *major,
                             double *inter, double *minor);

/* Page-locked host buffers for results.  When out12, n_iter and n_pts of cpg_shape all live in such buffers, the
 * kernels store every finished halo-shell straight into them while they run (no result copy after the kernels; option
 * "direct_results" = 0 restores the copy); pageable result buffers are filled by an ordinary device->host copy, which
 * runs at a fraction of the PCIe rate.  No counterpart in the reference, whose drivers np.zeros() their outputs
 * (shape_profs_algos.pyx:560-572). */
CPG_API int cpg_host_alloc(size_t bytes, void **out);
CPG_API void cpg_host_free(void *p);

/* Radial-prefix lengths n_eff (n_obj,R) of the most recent local cpg_shape call (diagnostics: how many
 * particles shell k actually streams per Katz iteration; see cpg_gather.cuh "Radial ordering"). */
CPG_API int cpg_get_prefix_lengths(cpg_ctx *ctx, int32_t *n_eff, int32_t R);

/* Counters of a context (diagnostics, tests): "uploads_skipped" (calls the snapshot cache answered), "catalogue_as_ranges"
 * (1: the resident catalogue is kept as contiguous ranges), "n_part", "n_obj", "snapshot_version", "catalogue_version",
 * "f32_test" (build option CPG_F32_TEST: 1 = membership tests pre-classified in fp32). */
CPG_API int cpg_get_stat(cpg_ctx *ctx, const char *name, int64_t *out);

/* Timing breakdown of the most recent compute call on this context. */
CPG_API int cpg_get_timing(cpg_ctx *ctx, cpg_timing *out);

/* Measured fp64 multiply-add peak of the context's device in 1e12 FMA/s (independent FMA chains on every SM, CUDA
 * events, best of a few repetitions): the denominator of bench.py's fp64 roofline. */
CPG_API int cpg_measure_fp64_peak(cpg_ctx *ctx, double *tfma_per_s);

/*
 * Several GPUs of one box.  The reference spreads the objects of ONE catalogue over the
 * host's cores (prange, shape_profs_algos.pyx:592, dens_profs_algos.pyx:102); cpg_multi spreads them over devices:
 * a particle-count-balanced LPT partition of the objects (cpg_partition_lpt: descending cost, each object to the least
 * loaded part; owner_out[h] = part of object h), every device receives only its own objects' particles and runs the
 * single-device path on its own host thread, results land in the rows of the caller's arrays.  No collective.
 *   cpg_multi_set_snapshot  = cpg_set_particles + cpg_set_catalogue for the whole catalogue (contiguous-range
 *                             catalogues ship each shard's ranges only; others ship the snapshot to every device)
 *   cpg_multi_shape         = cpg_shape; center_mode < 0: centers (n_obj,3) given; 0 / 1: 'com' / 'mode' centres
 *                             computed per device (cpg_centers, shape path; ref_points as there) -> centers_out
 *   cpg_multi_context       = the single-device context of a slot (owned by the handle) for the other entry points,
 *   cpg_multi_shard           with cpg_multi_shard giving the catalogue rows (ascending) that slot holds
 *   cpg_multi_get_info      = per slot: device ordinal, objects, resident particles, timing of its last call;
 *                             host_ms[4] = partition, sharded upload, result scatter (mean per device), last call (wall)
 */
typedef struct cpg_multi cpg_multi;
CPG_API int cpg_partition_lpt(const double *cost, int32_t n, int32_t n_parts, int32_t *owner_out);
CPG_API int cpg_multi_create(const int32_t *devices, int32_t n_dev, cpg_multi **out);
CPG_API void cpg_multi_destroy(cpg_multi *m);
CPG_API int cpg_multi_devices(cpg_multi *m);
CPG_API cpg_ctx *cpg_multi_context(cpg_multi *m, int32_t slot);
CPG_API int cpg_multi_set_option(cpg_multi *m, const char *name, int64_t value);
CPG_API int cpg_multi_set_snapshot(cpg_multi *m, const double *xyz, const double *vxyz, const double *masses,
                                   int64_t n_part, const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size,
                                   int32_t n_obj);
CPG_API int cpg_multi_shard(cpg_multi *m, int32_t slot, int32_t *n_obj_out, int64_t *n_part_out, int32_t *objs_out);
CPG_API int cpg_multi_shape(cpg_multi *m, const double *centers, int32_t center_mode, const double *ref_points,
                            const double *r200, const double *r_over_r200, int32_t R, int32_t local, double SAFE,
                            float L_BOX, double IT_TOL, int32_t IT_WALL, int32_t IT_MIN, int32_t reduced,
                            int32_t shell_based, int32_t use_vel, double *out12, int32_t *n_iter, int32_t *n_pts,
                            double *obj_mass, double *vcenters, double *centers_out);
CPG_API int cpg_multi_get_info(cpg_multi *m, int32_t slot, int32_t *device, int32_t *n_obj, int64_t *n_part,
                               cpg_timing *timing, double *host_ms);

/* Tunables (scheduling only; results do not depend on them beyond fp64 summation order):
 *   "shells_per_pass"  1,2,4 or 0=auto   how many shells of one object share one read of its particles
 *   "shell_groups"     1 (default, also 0) or 2 groups of 4 shells per warp task: the groups are passed over one after
 *                      the other and share one reduction / eigen-solve step per Katz iteration (slower on configs[1])
 *   "async_step"       which warp kernel serves the objects below "large_halo_min" (position shapes; same results):
 *                      2 (default): kernel 1d when the call has >= 8 x 12 x SMs tasks -- every warp works on TWO tasks at a
 *                      time and runs one eigen-solve / convergence step for the shells of both (one lane per shell), so the
 *                      serial chain of the step is paid once per two task-iterations (configs[1]: local-shape kernels
 *                      17.5 -> 17.0 ms) -- and kernel 1 (a warp per task, its own step) for shorter queues and for the
 *                      velocity-dispersion variant; 3: kernel 1d whatever the queue length; 1: kernel 1c, five "pass" warps
 *                      per CTA that only stream particles and hand their sums to one "step" warp through shared memory
 *                      (measured 17.45 - 18.1 ms: the step leaves the pass warps' chain, but a twelfth of the warps no longer
 *                      streams and the pass warps wait 19 % of the time for the step warp); 0: kernel 1 always
 *   "batched_step"     1: the warps of a CTA of the warp kernel meet once per Katz iteration and ONE of them runs the
 *                      eigen-solve / convergence step for all their shells (16 useful lanes per fp64 instruction instead
 *                      of 4).  Same results; measured slower (22.2 against 19.2 ms on configs[1]: the wait for one another
 *                      lengthens the per-warp dependency chain that bounds the kernel).  0 (default): every warp steps its own 4 shells
 *   "large_halo_min"   particles from which a thread-block cluster works on one object
 *   "grid_halo_min"    particles from which every Katz iteration of an object is one launch over the whole grid
 *                      (default 2^22)
 *   "cluster_size"     CTAs per cluster for large objects (1,2,4,8,16; 0=auto)
 *   "invalidate_gather" (any value) one-shot: forget the resident gathered catalogue, so that the next compute
 *                      call redoes the gather even if snapshot, catalogue, centres and L_BOX are unchanged
 *   "radial_order"     1 (default): gather every object in radial-bucket order and let shell k stream only
 *                      the particles with |x-c| < d_k (they are the only possible members); the shell-based variant
 *                      also skips the particles with |x-c| < s*d - delta (inside the inner ellipsoid whatever its
 *                      orientation).  0: catalogue order, every particle tested
 *   "warp_ctas_per_sm" diagnostics: resident CTAs (4 warps each) per SM of the warp kernel, 0 (default) = as many as fit
 *   "direct_results"   1 (default): cpg_shape stores results directly into page-locked caller buffers (see cpg_host_alloc)
 *   "upload_threads"   host threads (0..8, default 8) that stage PAGEABLE host arrays of >= 64 MB through page-locked bounce
 *                      buffers in cpg_set_particles / cpg_set_catalogue; 0 leaves them to cudaMemcpyAsync.  Page-locked
 *                      sources are always copied directly.
 *   "snapshot_cache"   1 (default): cpg_set_particles / cpg_set_catalogue return at once when called with the arrays the
 *                      device already holds -- same pointers, same lengths, same content token (the reference's class
 *                      hands the same xyz / masses / idx_cat to every driver, common/cosmic_base_class.pyx:155,232,863);
 *                      the token samples the arrays (cpg_host_token), the flat catalogue is hashed in full.  2: full
 *                      hash of the particle arrays too; 0: always upload
 *   "detect_ranges"    1 (default): cpg_set_catalogue keeps a contiguous-range catalogue as ranges (no index array)
 *   "prefix_table"     1 (default): non-reduced ellipsoid-based shapes on radially ordered particles take the tensor
 *                      sums of the whole 128-particle tiles at |x-c| < s*d (certain members) from a per-object prefix
 *                      table built once per gather instead of streaming them every Katz iteration; 0: no inner prefix
 */
CPG_API int cpg_set_option(cpg_ctx *ctx, const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* COSMIC_GPU_H */
