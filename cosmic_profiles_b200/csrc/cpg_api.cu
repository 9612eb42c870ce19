// cpg_api.cu -- C ABI (include/cosmic_gpu.h) and host-side orchestration of libcosmicgpu.
//
// Host logic mirrors the reference drivers (cosmic_profiles/shape_profs/shape_profs_algos.pyx:517-972,
// cosmic_profiles/dens_profs/dens_profs_algos.pyx:15-280) minus everything that is arithmetic: that all
// happens in the kernels of cpg_gather.cuh / cpg_morph.cuh / cpg_dens.cuh.  There is no CPU fallback.
#include <math.h>
#include <stdarg.h>

#include <algorithm>
#include <climits>
#include <limits>
#include <atomic>
#include <chrono>
#include <new>
#include <numeric>
#include <thread>
#include <vector>

#include "cpg_centers.cuh"
#include "cpg_dens.cuh"
#include "cpg_fit.cuh"
#include "cpg_gadget.cuh"
#include "cpg_gather.cuh"
#include "cpg_interp.cuh"
#include "cpg_morph.cuh"

#ifndef CPG_ASYNC_DEFAULT
#define CPG_ASYNC_DEFAULT 2   // default of the "async_step" option: 0 kernel 1, 1 kernel 1c (step warp), 2 kernel 1d (two task slots per warp)
#endif

namespace cpg {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap && p) return CPG_OK;
        release();
        const size_t want = (n ? n : 1) + (16 + sizeof(T) - 1) / sizeof(T);   // 16 bytes of slack (small_upload pads)
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e));
            (void)cudaGetLastError();
            return CPG_ERR_NOMEM;
        }
        cap = want;
        return CPG_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

#define CPG_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != CPG_OK) return rc__; \
    } while (0)

}  // namespace cpg

using namespace cpg;

struct cpg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: large-object kernel, concurrent with the warp kernel
    cudaEvent_t ev[6] = {};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // page-locked, device-mapped staging for the small per-call inputs (centres, radii, task list): they are
    // pulled by a copy kernel instead of queueing on the host->device copy engine behind another context's
    // multi-GB snapshot upload
    unsigned char *stage = nullptr;
    size_t stage_cap = 0, stage_used = 0;
    // pageable host arrays (the drop-in drivers receive plain numpy arrays): worker threads copy pieces into
    // page-locked bounce buffers and queue the host->device copies on their own streams (upload_pageable)
    static constexpr int UP_THREADS = 8, UP_BUFS = 2;
    static constexpr size_t UP_PIECE = (size_t)8 << 20;
    unsigned char *up_buf[UP_THREADS][UP_BUFS] = {};
    cudaStream_t up_stream[UP_THREADS] = {};
    cudaEvent_t up_ev[UP_THREADS][UP_BUFS] = {};
    bool up_ready = false;
    int opt_upload_threads = UP_THREADS;   // 0: leave pageable sources to cudaMemcpyAsync
    int opt_direct_results = 1;            // cpg_shape: kernels store results straight into page-locked caller buffers
    // snapshot
    DevBuf<double> xyz, vxyz, masses;
    int64_t n_part = 0;
    bool have_vel = false;
    // catalogue: flat int32 index array, or (cat_ranges) one contiguous snapshot range per object and no index array
    DevBuf<int32_t> idx, size;
    DevBuf<int64_t> off, poff, start;
    bool cat_ranges = false;
    int64_t cat_max_idx = -1;     // largest snapshot index the catalogue refers to
    bool cat_stale = false;       // a later, smaller snapshot no longer covers the catalogue: compute calls refuse
    // what the device currently holds, as seen from the host (cpg_set_particles / cpg_set_catalogue skip the upload
    // when asked for the same thing again: the reference's class passes the same arrays to every driver,
    // common/cosmic_base_class.pyx:155,232,863)
    int opt_snapshot_cache = 1;   // 0 off, 1 pointer + length + sampled content token, 2 pointer + length + full hash
    int opt_detect_ranges = 1;    // cpg_set_catalogue: recognise contiguous-range catalogues, keep no index array
    struct {
        bool valid = false;
        const void *xyz = nullptr, *vxyz = nullptr, *masses = nullptr;
        int64_t n = 0;
        int mode = 0;
        uint64_t tok[3] = {0, 0, 0};
    } snap_key;
    struct {
        bool valid = false, ranges = false;
        int64_t n_idx = 0;
        uint64_t hash = 0;
        std::vector<int64_t> start;
        std::vector<int32_t> size;
    } cat_key;
    uint64_t uploads_skipped = 0;
    DevBuf<unsigned long long> morph_stats;   // [0] tiles, [1] tile-shells visited by the morphology kernels of a call
    DevBuf<double> peak_scratch;
    int64_t n_idx = 0, n_pad = 0;
    int32_t n_obj = 0;
    std::vector<int32_t> h_size;
    std::vector<int32_t> order;   // objects sorted by size, largest first
    // chunk lists
    std::vector<Chunk> h_chunks, h_kchunks, h_schunks;   // CHUNK (statistics, binning), KDE_CHUNK, SORT_CHUNK particles
    std::vector<int32_t> h_chunk_first, h_kchunk_first, h_schunk_first;
    std::vector<int32_t> h_big_sort, h_big_stat;   // objects of more than BIG_OBJECT_CHUNKS sort / statistics chunks
    DevBuf<int32_t> big_stat, seg_tot;
    DevBuf<Chunk> chunks, kchunks, schunks;
    DevBuf<int32_t> chunk_first, kchunk_first, schunk_first;
    // gathered SoA
    DevBuf<double> g_dx, g_dy, g_dz, g_m, g_vx, g_vy, g_vz;
    bool mass_stream_valid = false;
    // per-object scratch
    DevBuf<double> centers, vcenters, r200, rr, obj_mass, part_mass, vpart, part_mm, obj_eqm;
    DevBuf<int32_t> part_flag;
    DevBuf<float> mass32;
    DevBuf<uint8_t> skip, bucket;
    DevBuf<int32_t> tile_hist, dest, n_eff, cidxA, cidxB, d_order;
    DevBuf<double> cref, ptab;
    bool sorted_valid = false;    // the SoA streams are in radial-bucket order and `dest` maps catalogue -> slot
    // what the gathered SoA currently holds (skip the gather when a call asks for the same thing again)
    uint64_t snapshot_version = 0, catalogue_version = 0;
    struct {
        bool valid = false, stats = false, vel = false, mmax = false;
        bool parts_from_sort = false;   // the per-chunk statistics were produced by the ordering gather (sort chunks)
        int ptab = 0;   // prefix tensor table: 0 none, 1 positions, 2 velocities
        uint64_t snap = 0, cat = 0;
        float L_BOX = 0.f;
        int sort_R = 0;
        std::vector<double> centers, r200, rr;
    } have;
    // task lists are cached per (catalogue, R, S, thresholds): a getShapeCatLocal + getShapeCatGlobal sequence
    // alternates between two of them, and rebuilding + uploading the 1.8e5 tasks of configs[1] costs ~1.5 ms
    struct TaskKey {
        bool valid = false;
        uint64_t cat = 0, used = 0;
        int S = 0, R = 0, nw = 0, ncl = 0, ngrid = 0, cluster_size = 1;
        int64_t grid_min = 0;
        int64_t large_min = 0;
        DevBuf<Task> tasks;
    } task_cache[4];
    uint64_t task_clock = 0;
    // morphology
    DevBuf<GridState> grid_states;
    DevBuf<double> grid_partials;
    int64_t opt_grid_halo_min = 0;
    DevBuf<int32_t> task_counter;
    DevBuf<double> out12;
    DevBuf<int32_t> n_iter, n_pts;
    // densities
    DevBuf<double> d_a, d_b, d_c, d_major, d_inter, d_minor, d_edges, dens, bin_mass, chunk_mmax;
    DevBuf<int32_t> bin_cnt, counts;
    // Gadget / FoF ingest (cpg_gadget.cuh): group tables, scans, the compacted central-subhalo catalogue
    DevBuf<int32_t> oc_nb, oc_shlen, oc_fof, oc_size, oc_idx, oc_err;
    DevBuf<double> oc_r200in, oc_r200;
    DevBuf<int64_t> oc_shoff, oc_start, oc_passoff, oc_idxoff, oc_cstart, oc_coff;
    DevBuf<Pair64> oc_tiles;
    DevBuf<unsigned char> raw_block;   // a snapshot block as stored on disk, before conversion to fp64
    DevBuf<double> ip_a, ip_b, ip_c, ip_major, ip_inter, ip_minor, ip_rr, ip_edges;   // raw shape profiles for cpg_shape_interp
    int interp_R = 0;          // > 0: d_a ... d_minor hold profiles interpolated on the device for this many bins
    int32_t interp_n = 0;
    DevBuf<double> fit_radii, fit_lmed, fit_avg, fit_x0, fit_best, fit_fun;
    DevBuf<int32_t> fit_nvalid, fit_nfev, fit_nit;
    bool oc_valid = false;
    int64_t oc_n_pass = 0, oc_n_idx = 0;
    // options
    int opt_shells_per_pass = 0;
    int64_t opt_large_halo_min = 0;
    int opt_cluster_size = 0;
    int opt_radial_order = 1;
    int opt_prefix_table = 1;
    int opt_groups = 0;   // 0 auto, 1 one group of S shells per warp task, 2 two groups
    int opt_warp_ctas_per_sm = 0;   // diagnostics: resident CTAs per SM of the warp kernel (0 = as many as fit)
    struct CenterCache {   // the centres of the last cpg_centers call and what they were computed from
        bool valid = false;
        int64_t snap = -1, cat = -1;
        int mode = -1, shape_path = -1;
        double L_BOX = 0.0;
        std::vector<double> ref, centers;
    } cen_cache;
    int opt_async_step = CPG_ASYNC_DEFAULT;     // warp kernel 1c: pass warps + one asynchronous step warp per CTA (cpg_morph.cuh)
    int opt_batched_step = 0;   // warp kernel: one warp per CTA runs the per-iteration step of all the CTA's warps (cpg_morph.cuh, kernel 1b)
    cpg_timing timing = {};
};

namespace {

int use_device(cpg_ctx *c)
{
    if (!c) {
        set_error("null context");
        return CPG_ERR_ARGUMENT;
    }
    CPG_CUDA(cudaSetDevice(c->device));
    return CPG_OK;
}

// Entry check of every compute call: device current, catalogue still covered by the snapshot.
int compute_ready(cpg_ctx *c)
{
    CPG_TRY(use_device(c));
    if (c->cat_stale) {
        set_error("the catalogue indexes particles up to %lld but the snapshot uploaded after it has only %lld: "
                  "call cpg_set_catalogue again", (long long)c->cat_max_idx, (long long)c->n_part);
        return CPG_ERR_STATE;
    }
    return CPG_OK;
}

// Large host->device transfers are issued in pieces: the copy engine serves streams copy by copy, so a small
// upload of another context (centres, radii) is not stuck behind a multi-GB transfer for tens of ms.
// Pageable source: cudaMemcpyAsync bounces every piece through the driver's staging buffer from the calling thread
// alone (12 GB/s on the 16-core host measured here, a fifth of the PCIe rate).  Here up to UP_THREADS host threads
// each own two page-locked bounce buffers and a stream: thread t copies pieces t, t + T, ... into its buffers and queues
// their host->device copies, so that the host-side memcpy of one piece overlaps the PCIe transfer of others.
int ensure_upload_buffers(cpg_ctx *c)
{
    if (c->up_ready) return CPG_OK;
    for (int t = 0; t < cpg_ctx::UP_THREADS; t++) {
        CPG_CUDA(cudaStreamCreateWithFlags(&c->up_stream[t], cudaStreamNonBlocking));
        for (int b = 0; b < cpg_ctx::UP_BUFS; b++) {
            CPG_CUDA(cudaHostAlloc((void **)&c->up_buf[t][b], cpg_ctx::UP_PIECE, cudaHostAllocDefault));
            CPG_CUDA(cudaEventCreateWithFlags(&c->up_ev[t][b], cudaEventDisableTiming));
        }
    }
    c->up_ready = true;
    return CPG_OK;
}

int upload_pageable(cpg_ctx *c, void *dst, const void *src, size_t bytes)
{
    constexpr int NB = cpg_ctx::UP_BUFS;
    const int T = std::min(c->opt_upload_threads, (int)cpg_ctx::UP_THREADS);
    const size_t piece = cpg_ctx::UP_PIECE;
    CPG_TRY(ensure_upload_buffers(c));
    const size_t n_pieces = (bytes + piece - 1) / piece;
    std::atomic<int> failed{0};
    auto work = [&](int t) {
        if (cudaSetDevice(c->device) != cudaSuccess) { failed = 1; return; }
        int turn = 0;
        for (size_t i = (size_t)t; i < n_pieces && !failed; i += T, turn++) {
            const int b = turn % NB;
            const size_t o = i * piece, nb = std::min(piece, bytes - o);
            if (turn >= NB && cudaEventSynchronize(c->up_ev[t][b]) != cudaSuccess) { failed = 1; return; }
            memcpy(c->up_buf[t][b], (const char *)src + o, nb);
            if (cudaMemcpyAsync((char *)dst + o, c->up_buf[t][b], nb, cudaMemcpyHostToDevice, c->up_stream[t]) != cudaSuccess ||
                cudaEventRecord(c->up_ev[t][b], c->up_stream[t]) != cudaSuccess) { failed = 1; return; }
        }
        // the bounce buffers are reused by the next upload, and the caller's stream is not ordered behind these
        // streams: drain before returning
        if (cudaStreamSynchronize(c->up_stream[t]) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>((size_t)T, n_pieces);
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (failed) {
        set_error("upload_pageable: a CUDA call failed: %s", cudaGetErrorString(cudaGetLastError()));
        return CPG_ERR_CUDA;
    }
    return CPG_OK;
}

bool is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

int upload_chunked(cpg_ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t stream)
{
    if (c->opt_upload_threads > 0 && bytes >= ((size_t)64 << 20) && is_pageable(src)) {
        // earlier work on `stream` may still be reading dst's previous contents
        CPG_CUDA(cudaStreamSynchronize(stream));
        return upload_pageable(c, dst, src, bytes);
    }
    const size_t piece = (size_t)32 << 20;
    for (size_t o = 0; o < bytes; o += piece) {
        const size_t nb = std::min(piece, bytes - o);
        CPG_CUDA(cudaMemcpyAsync((char *)dst + o, (const char *)src + o, nb, cudaMemcpyHostToDevice, stream));
    }
    return CPG_OK;
}


// ---------------------------------------------------------------------------------------------------
// Host-side helpers of the snapshot / catalogue cache and of the sharded uploads.
// ---------------------------------------------------------------------------------------------------
inline uint64_t mix64(uint64_t h, uint64_t w)
{
    h ^= w;
    h *= 0x9E3779B97F4A7C15ull;
    return h ^ (h >> 29);
}

uint64_t hash_bytes(const unsigned char *p, size_t n, uint64_t seed)
{
    // four independent lanes (the multiply chain is the bottleneck of one), folded at the end
    uint64_t h0 = seed, h1 = seed ^ 0x1111111111111111ull, h2 = seed ^ 0x2222222222222222ull, h3 = seed ^ 0x3333333333333333ull;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        uint64_t w[4];
        memcpy(w, p + i, 32);
        h0 = mix64(h0, w[0]); h1 = mix64(h1, w[1]); h2 = mix64(h2, w[2]); h3 = mix64(h3, w[3]);
    }
    uint64_t tail = 0;
    if (i < n) {
        unsigned char t[32] = {0};
        memcpy(t, p + i, n - i);
        uint64_t w[4];
        memcpy(w, t, 32);
        h0 = mix64(h0, w[0]); h1 = mix64(h1, w[1]); h2 = mix64(h2, w[2]); h3 = mix64(h3, w[3]);
        tail = n - i;
    }
    return mix64(mix64(mix64(mix64(h0, h1), h2), h3), (uint64_t)n ^ (tail << 56));
}

int host_threads(size_t bytes)
{
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t want = bytes / ((size_t)16 << 20) + 1;   // one thread per 16 MB
    return (int)std::min<size_t>(std::min<size_t>(want, 12), std::max(1u, hw - 2));
}

// Content token of a host array.  full: hash of every byte (multi-threaded); else: the first and last 4 KB and 4096
// evenly spaced 64-byte blocks -- enough to notice a rescaled, recentred, re-read or reordered snapshot, not a
// deliberate edit of a few particles in place (callers that do that use snapshot_cache = 2 or 0).
uint64_t host_token(const void *ptr, size_t bytes, bool full)
{
    const unsigned char *p = (const unsigned char *)ptr;
    if (!p || bytes == 0) return 0x5eedull;
    if (!full) {
        uint64_t h = hash_bytes(p, std::min<size_t>(bytes, 4096), 1);
        if (bytes > 4096) h = mix64(h, hash_bytes(p + bytes - 4096, 4096, 2));
        if (bytes > 8192) {
            const size_t nblk = 4096, span = bytes - 64;
            for (size_t k = 0; k < nblk; k++) h = mix64(h, hash_bytes(p + (size_t)((double)span * (double)k / (double)nblk), 64, 3));
        }
        return h;
    }
    const int T = host_threads(bytes);
    std::vector<uint64_t> part(T, 0);
    auto work = [&](int t) {
        const size_t a = bytes * (size_t)t / T, b = bytes * (size_t)(t + 1) / T;
        part[t] = hash_bytes(p + a, b - a, 7 + t);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    uint64_t h = 11;
    for (int t = 0; t < T; t++) h = mix64(h, part[t]);
    return h;
}

// One pass over a flat index catalogue: is every object a contiguous ascending run of snapshot indices (the FoF /
// SUBFIND layout, gadget/gen_catalogues.pyx:29-35)?  Also the largest / smallest index and a hash of all entries.
struct CatScan {
    bool contiguous = true;
    int64_t min_idx = INT64_MAX, max_idx = -1;
    uint64_t hash = 0;
    std::vector<int64_t> start;   // first snapshot index of every object (objects of size 0: 0)
};

void scan_catalogue(const int32_t *idx, const int64_t *off, int32_t n_obj, CatScan &out)
{
    const int64_t n_idx = off[n_obj];
    out.start.assign((size_t)n_obj, 0);
    const int T = host_threads((size_t)n_idx * 4);
    std::vector<CatScan> part(T);
    auto work = [&](int t) {
        CatScan &r = part[t];
        // objects [h0, h1): split by entries
        const int64_t e0 = n_idx * t / T, e1 = n_idx * (t + 1) / T;
        int h0 = (int)(std::upper_bound(off, off + n_obj + 1, e0) - off) - 1;
        int h1 = (t == T - 1) ? n_obj : (int)(std::upper_bound(off, off + n_obj + 1, e1) - off) - 1;
        if (t == 0) h0 = 0;
        h0 = std::max(h0, 0);
        uint64_t h = 17 + (uint64_t)t;
        for (int o = h0; o < h1; o++) {
            const int64_t a = off[o], b = off[o + 1];
            if (b == a) continue;
            const int32_t *q = idx + a;
            const int64_t first = q[0], n = b - a;
            out.start[o] = first;
            bool run = true;
            int32_t lo = q[0], hi = q[0];
            // four independent multiply-add chains (one would bound the pass at ~1 element per multiply latency)
            uint64_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
            int64_t i = 0;
            for (; i + 4 <= n; i += 4) {
                const int32_t v0 = q[i], v1 = q[i + 1], v2 = q[i + 2], v3 = q[i + 3];
                const int32_t e = (int32_t)(first + i);
                run &= (v0 == e) & (v1 == e + 1) & (v2 == e + 2) & (v3 == e + 3);
                lo = std::min(std::min(lo, v0), std::min(std::min(v1, v2), v3));
                hi = std::max(std::max(hi, v0), std::max(std::max(v1, v2), v3));
                h0 = h0 * 0x100000001B3ull + (uint32_t)v0; h1 = h1 * 0x100000001B3ull + (uint32_t)v1;
                h2 = h2 * 0x100000001B3ull + (uint32_t)v2; h3 = h3 * 0x100000001B3ull + (uint32_t)v3;
            }
            for (; i < n; i++) {
                const int32_t v = q[i];
                run &= (v == (int32_t)(first + i));
                lo = std::min(lo, v); hi = std::max(hi, v);
                h0 = h0 * 0x100000001B3ull + (uint32_t)v;
            }
            const uint64_t hh = mix64(mix64(mix64(h0, h1), h2), h3);
            if (first + n - 1 > 0x7fffffffLL) run = false;
            r.contiguous &= run;
            r.min_idx = std::min<int64_t>(r.min_idx, lo);
            r.max_idx = std::max<int64_t>(r.max_idx, hi);
            h = mix64(h, hh ^ (uint64_t)n);
        }
        r.hash = h;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    uint64_t h = 23;
    for (auto &r : part) {
        out.contiguous &= r.contiguous;
        out.min_idx = std::min(out.min_idx, r.min_idx);
        out.max_idx = std::max(out.max_idx, r.max_idx);
        h = mix64(h, r.hash);
    }
    out.hash = h;
}

// Upload the concatenation of the element ranges [seg_start[i], seg_start[i] + seg_len[i]) of a host array (elements of
// `esz` bytes) to dst.  Always staged through the page-locked bounce buffers by the upload threads (the packing of
// the segments is a host memcpy either way); prefix[i] = elements before segment i, prefix[n_seg] = total.
int upload_segments(cpg_ctx *c, void *dst, const void *src, size_t esz, const int64_t *seg_start, const int64_t *seg_len,
                    int64_t n_seg, const std::vector<int64_t> &prefix)
{
    const size_t bytes = (size_t)prefix[n_seg] * esz;
    if (bytes == 0) return CPG_OK;
    CPG_TRY(ensure_upload_buffers(c));
    constexpr int NB = cpg_ctx::UP_BUFS;
    const int T = std::max(1, std::min(c->opt_upload_threads > 0 ? c->opt_upload_threads : 4, (int)cpg_ctx::UP_THREADS));
    const size_t piece = cpg_ctx::UP_PIECE;
    const size_t n_pieces = (bytes + piece - 1) / piece;
    std::atomic<int> failed{0};
    auto work = [&](int t) {
        if (cudaSetDevice(c->device) != cudaSuccess) { failed = 1; return; }
        int turn = 0;
        for (size_t i = (size_t)t; i < n_pieces && !failed; i += T, turn++) {
            const int b = turn % NB;
            const size_t o = i * piece, nb = std::min(piece, bytes - o);
            if (turn >= NB && cudaEventSynchronize(c->up_ev[t][b]) != cudaSuccess) { failed = 1; return; }
            // pack [o, o + nb) of the concatenated stream
            size_t done = 0;
            int64_t sg = (int64_t)(std::upper_bound(prefix.begin(), prefix.end(), (int64_t)(o / esz)) - prefix.begin()) - 1;
            while (done < nb) {
                const size_t so = (size_t)prefix[sg] * esz, sl = (size_t)seg_len[sg] * esz;
                const size_t within = (o + done) - so;
                const size_t take = std::min(nb - done, sl - within);
                if (take) memcpy(c->up_buf[t][b] + done, (const char *)src + (size_t)seg_start[sg] * esz + within, take);
                done += take;
                if (within + take == sl) sg++;
            }
            if (cudaMemcpyAsync((char *)dst + o, c->up_buf[t][b], nb, cudaMemcpyHostToDevice, c->up_stream[t]) != cudaSuccess ||
                cudaEventRecord(c->up_ev[t][b], c->up_stream[t]) != cudaSuccess) { failed = 1; return; }
        }
        if (cudaStreamSynchronize(c->up_stream[t]) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>((size_t)T, n_pieces);
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (failed) {
        set_error("upload_segments: a CUDA call failed: %s", cudaGetErrorString(cudaGetLastError()));
        return CPG_ERR_CUDA;
    }
    return CPG_OK;
}

__global__ void stage_copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
}

int grid_for(int64_t n, int block, int cap);

// Host -> device copy of a small array through the mapped staging buffer (dst must be 16-byte aligned, which
// every cudaMalloc'ed buffer is; the tail is padded inside the staging area, so dst needs 16-byte slack:
// DevBuf allocations are rounded up accordingly).
int small_upload(cpg_ctx *c, void *dst, const void *src, size_t bytes)
{
    const size_t padded = (bytes + 15) & ~(size_t)15;
    if (c->stage_used + padded > c->stage_cap) {
        // grow: wait for earlier staged copies, then reallocate
        CPG_CUDA(cudaStreamSynchronize(c->stream));
        const size_t want = std::max<size_t>((c->stage_used + padded) * 2, (size_t)4 << 20);
        if (c->stage) cudaFreeHost(c->stage);
        c->stage = nullptr; c->stage_cap = 0; c->stage_used = 0;
        CPG_CUDA(cudaHostAlloc((void **)&c->stage, want, cudaHostAllocMapped | cudaHostAllocPortable));
        c->stage_cap = want;
    }
    unsigned char *h = c->stage + c->stage_used;
    memcpy(h, src, bytes);
    if (padded > bytes) memset(h + bytes, 0, padded - bytes);
    c->stage_used += padded;
    const size_t n16 = padded / 16;
    stage_copy_kernel<<<grid_for((int64_t)n16, 256, c->sm_count * 4), 256, 0, c->stream>>>((const uint4 *)h, (uint4 *)dst, n16);
    CPG_CUDA(cudaGetLastError());
    return CPG_OK;
}

int grid_for(int64_t n, int block, int cap)
{
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

SoA soa_view(cpg_ctx *c)
{
    SoA s;
    s.dx = c->g_dx.p; s.dy = c->g_dy.p; s.dz = c->g_dz.p; s.m = c->g_m.p;
    s.vx = c->g_vx.p; s.vy = c->g_vy.p; s.vz = c->g_vz.p;
    s.poff = c->poff.p;
    s.size = c->size.p;
    return s;
}

GatherParams gather_view(cpg_ctx *c, float L_BOX)
{
    GatherParams G;
    G.xyz = c->xyz.p; G.vxyz = c->vxyz.p; G.masses = c->masses.p;
    G.idx = c->cat_ranges ? nullptr : c->idx.p; G.start = c->start.p; G.off = c->off.p; G.poff = c->poff.p;
    G.centers = c->centers.p; G.vcenters = c->vcenters.p;
    G.n_obj = c->n_obj; G.n_idx = c->n_idx; G.L_BOX = L_BOX;
    G.dx = c->g_dx.p; G.dy = c->g_dy.p; G.dz = c->g_dz.p; G.m = c->g_m.p;
    G.vx = c->g_vx.p; G.vy = c->g_vy.p; G.vz = c->g_vz.p;
    G.dest = c->sorted_valid ? c->dest.p : nullptr;
    return G;
}

__global__ void idx_range_kernel(const int32_t *idx, int64_t n, int32_t *minmax)
{
    int lo = 0x7fffffff, hi = -0x7fffffff - 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int v = idx[j];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[0], lo);
        atomicMax(&minmax[1], hi);
    }
}

// Result arrays of cpg_shape: out12 zeros (what the reference returns for a failed halo-shell), counters -1
// (= the per-object early exit), work-queue heads 0.
__global__ void init_results_kernel(double2 *__restrict__ out12, int32_t *__restrict__ n_iter, int32_t *__restrict__ n_pts,
                                    int32_t *__restrict__ counters, unsigned long long *__restrict__ stats, int64_t nR)
{
    if (blockIdx.x == 0 && threadIdx.x < 2) stats[threadIdx.x] = 0ull;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < nR * 6; i += stride) out12[i] = make_double2(0.0, 0.0);
    for (int64_t i = i0; i < nR; i += stride) { n_iter[i] = -1; n_pts[i] = -1; }
    if (i0 < 4) counters[i0] = 0;
}

// Device-side alias of a page-locked host buffer (nullptr if `p` is pageable or device memory).
void *pinned_device_alias(const void *p)
{
    cudaPointerAttributes at;
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

struct Timer {
    cpg_ctx *c;
    explicit Timer(cpg_ctx *ctx) : c(ctx) { memset(&c->timing, 0, sizeof c->timing); }
    void mark(int i) { cudaEventRecord(c->ev[i], c->stream); }
    static double ms(cudaEvent_t a, cudaEvent_t b)
    {
        float t = 0.f;
        cudaEventElapsedTime(&t, a, b);
        return (double)t;
    }
};

// Gather positions and masses; requires centres (and, for the radial ordering, r200 / rr) on the device.
// sort_R > 0: write every object's particles in radial-bucket order of the sort_R shell radii and fill
// n_eff (cpg_gather.cuh "Radial ordering").  Skipped when the SoA already holds exactly this (same snapshot,
// catalogue, centres, L_BOX and -- if an ordering is requested -- the same radii): the order of the
// particles inside an object never matters to a kernel that does not use n_eff.
int run_gather(cpg_ctx *c, const double *h_centers, const double *h_r200, const double *h_rr, float L_BOX,
               int sort_R = 0)
{
    const int64_t n = c->n_idx;
    const size_t n3 = (size_t)c->n_obj * 3;
    auto &H = c->have;
    bool same = H.valid && H.snap == c->snapshot_version && H.cat == c->catalogue_version && H.L_BOX == L_BOX &&
                H.centers.size() == n3 && !memcmp(H.centers.data(), h_centers, n3 * sizeof(double));
    if (same && sort_R > 0)
        same = H.sort_R == sort_R && H.rr.size() == (size_t)sort_R && !memcmp(H.rr.data(), h_rr, sort_R * sizeof(double)) &&
               H.r200.size() == (size_t)c->n_obj && !memcmp(H.r200.data(), h_r200, c->n_obj * sizeof(double));
    if (same) return CPG_OK;
    H.valid = false; H.stats = false; H.vel = false; H.mmax = false; H.ptab = 0; H.parts_from_sort = false;
    CPG_TRY(c->g_dx.ensure(c->n_pad)); CPG_TRY(c->g_dy.ensure(c->n_pad));
    CPG_TRY(c->g_dz.ensure(c->n_pad)); CPG_TRY(c->g_m.ensure(c->n_pad));
    c->sorted_valid = false;
    const int nc = (int)c->h_schunks.size();   // ordering works on SORT_CHUNK-particle chunks
    if (sort_R > 0 && n > 0 && nc > 0) {
        const int B = sort_R + 1;
        CPG_TRY(c->bucket.ensure(n)); CPG_TRY(c->tile_hist.ensure((size_t)nc * B));
        CPG_TRY(c->n_eff.ensure((size_t)c->n_obj * sort_R));
        CPG_TRY(c->dest.ensure(n));
        SortParams P;
        P.G = gather_view(c, L_BOX);
        P.chunks = c->schunks.p; P.chunk_first = c->schunk_first.p; P.size = c->size.p;
        P.r200 = c->r200.p; P.rr = c->rr.p; P.n_chunks = nc; P.R = sort_R;
        P.bucket = c->bucket.p; P.tile_hist = c->tile_hist.p;
        P.dest = c->have_vel ? c->dest.p : nullptr;   // catalogue -> slot map: only the velocity gather needs it
        P.n_eff = c->n_eff.p;
        CPG_TRY(c->part_mass.ensure(nc)); CPG_TRY(c->part_flag.ensure(nc)); CPG_TRY(c->part_mm.ensure((size_t)2 * nc));
        P.part_mass = c->part_mass.p; P.part_spread = c->part_flag.p; P.part_mminmax = c->part_mm.p;
        H.parts_from_sort = true;
        CPG_CUDA(cudaFuncSetAttribute(gather_bucketed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)GATHER_BUCKETED_SMEM));
        bucket_hist_kernel<<<nc, 256, 0, c->stream>>>(P);
        bucket_scan_kernel<<<c->n_obj, 256, 0, c->stream>>>(P, c->n_obj);
        for (int32_t hb : c->h_big_sort) {   // (a handful at most: objects of > 2e6 particles)
            const int nchunk = c->h_schunk_first[hb + 1] - c->h_schunk_first[hb];
            const int nseg = (nchunk + SCAN_SEG - 1) / SCAN_SEG;
            CPG_TRY(c->seg_tot.ensure((size_t)nseg * B));
            bucket_segsum_kernel<<<nseg, 256, 0, c->stream>>>(P, hb, c->seg_tot.p);
            bucket_segscan_kernel<<<1, 256, 0, c->stream>>>(P, hb, nseg, c->seg_tot.p);
            bucket_segapply_kernel<<<nseg, 256, 0, c->stream>>>(P, hb, c->seg_tot.p);
            c->timing.kernel_launches += 3;
        }
        if (CPG_GATHER_STAGED) gather_bucketed_kernel<<<nc, 256, GATHER_BUCKETED_SMEM, c->stream>>>(P);
        else gather_bucketed2_kernel<<<nc, 256, 0, c->stream>>>(P);
        c->timing.kernel_launches += 3;
        c->sorted_valid = true;
    } else if (n > 0) {
        gather_pos_kernel<<<grid_for(n, 256, c->sm_count * 32), 256, 0, c->stream>>>(gather_view(c, L_BOX));
        c->timing.kernel_launches++;
    }
    CPG_CUDA(cudaGetLastError());
    c->mass_stream_valid = true;
    H.valid = true; H.snap = c->snapshot_version; H.cat = c->catalogue_version; H.L_BOX = L_BOX;
    H.sort_R = sort_R;
    H.centers.assign(h_centers, h_centers + n3);
    if (sort_R > 0) {
        H.rr.assign(h_rr, h_rr + sort_R);
        H.r200.assign(h_r200, h_r200 + c->n_obj);
    }
    return CPG_OK;
}

int run_stats(cpg_ctx *c)
{
    if (c->have.valid && c->have.stats) return CPG_OK;
    const bool fused = c->have.valid && c->have.parts_from_sort;   // chunk partials already written by the ordering gather
    const int nc = (int)c->h_chunks.size();
    if (!fused) {
        CPG_TRY(c->part_mass.ensure(nc)); CPG_TRY(c->part_flag.ensure(nc)); CPG_TRY(c->part_mm.ensure((size_t)2 * nc));
    }
    CPG_TRY(c->obj_mass.ensure(c->n_obj)); CPG_TRY(c->skip.ensure(c->n_obj)); CPG_TRY(c->obj_eqm.ensure(c->n_obj));
    if (nc > 0 && !fused) {
        stats_chunk_kernel<<<nc, 256, 0, c->stream>>>(soa_view(c), c->chunks.p, nc, c->part_mass.p, c->part_flag.p,
                                                     c->part_mm.p);
        c->timing.kernel_launches++;
    }
    if (c->n_obj > 0) {
        stats_final_kernel<<<grid_for((int64_t)c->n_obj * 32, 128, 1 << 30), 128, 0, c->stream>>>(
            fused ? c->schunk_first.p : c->chunk_first.p, c->size.p, c->n_obj, c->part_mass.p, c->part_flag.p, c->part_mm.p,
            c->obj_mass.p, c->skip.p, c->obj_eqm.p);
        c->timing.kernel_launches++;
        const std::vector<int32_t> &big = fused ? c->h_big_sort : c->h_big_stat;
        if (!big.empty()) {
            CPG_TRY(c->big_stat.ensure(big.size()));
            CPG_CUDA(cudaMemcpyAsync(c->big_stat.p, big.data(), sizeof(int32_t) * big.size(), cudaMemcpyHostToDevice, c->stream));
            stats_final_big_kernel<<<(int)big.size(), 256, 0, c->stream>>>(
                fused ? c->schunk_first.p : c->chunk_first.p, c->size.p, c->big_stat.p, c->part_mass.p, c->part_flag.p,
                c->part_mm.p, c->obj_mass.p, c->skip.p, c->obj_eqm.p);
            c->timing.kernel_launches++;
        }
    }
    CPG_CUDA(cudaGetLastError());
    c->have.stats = c->have.valid;
    return CPG_OK;
}

int run_vcenters(cpg_ctx *c, float L_BOX)
{
    if (c->have.valid && c->have.vel) return CPG_OK;
    const int nc = (int)c->h_chunks.size();
    CPG_TRY(c->mass32.ensure(c->n_obj)); CPG_TRY(c->vpart.ensure((size_t)nc * 3));
    CPG_TRY(c->vcenters.ensure((size_t)c->n_obj * 3));
    CPG_TRY(c->g_vx.ensure(c->n_pad)); CPG_TRY(c->g_vy.ensure(c->n_pad)); CPG_TRY(c->g_vz.ensure(c->n_pad));
    if (c->n_obj > 0) {
        mass32_kernel<<<grid_for((int64_t)c->n_obj * 32, 128, c->sm_count * 16), 128, 0, c->stream>>>(
            gather_view(c, L_BOX), c->n_obj, c->obj_eqm.p, c->mass32.p);
        if (nc > 0)
            vsum_chunk_kernel<<<nc, 256, 0, c->stream>>>(gather_view(c, L_BOX), soa_view(c), c->chunks.p, nc, c->vpart.p);
        vcenter_final_kernel<<<grid_for(c->n_obj, 128, 1 << 30), 128, 0, c->stream>>>(c->chunk_first.p, c->n_obj,
                                                                                       c->vpart.p, c->mass32.p,
                                                                                       c->vcenters.p);
        if (c->n_idx > 0)
            gather_vel_kernel<<<grid_for(c->n_idx, 256, c->sm_count * 32), 256, 0, c->stream>>>(gather_view(c, L_BOX));
        c->timing.kernel_launches += 4;
    }
    CPG_CUDA(cudaGetLastError());
    c->have.vel = c->have.valid;
    return CPG_OK;
}

// Prefix tensor table of the radially ordered SoA (cpg_gather.cuh): built once per gather, reused by every
// non-reduced ellipsoid-based cpg_shape call on it.
int run_ptab(cpg_ctx *c, bool use_vel)
{
    const int want = use_vel ? 2 : 1;
    if (c->have.valid && c->have.ptab == want) return CPG_OK;
    const int nc = (int)c->h_kchunks.size();
    CPG_TRY(c->ptab.ensure(((size_t)(c->n_pad >> 7) + (size_t)c->n_obj + 1) * 6));
    if (nc > 0) {
        tile_tensor_kernel<<<nc, 256, 0, c->stream>>>(soa_view(c), c->kchunks.p, nc, use_vel ? 1 : 0, c->ptab.p);
        tile_tensor_scan_kernel<<<c->n_obj, 256, 0, c->stream>>>(soa_view(c), c->n_obj, c->ptab.p);
        c->timing.kernel_launches += 2;
    }
    CPG_CUDA(cudaGetLastError());
    c->have.ptab = c->have.valid ? want : 0;
    return CPG_OK;
}

template <int S, int G, bool SHELL, bool REDUCED, bool VDISP>
int launch_morph(cpg_ctx *c, MorphParams P, int n_warp_tasks, int n_cluster_tasks, int cluster_size)
{
    // The few large objects (cluster kernel) and the many small ones (warp kernel) are independent: run the two
    // kernels concurrently (fork/join on a second stream) so that each fills the other's tail.
    const bool both = n_warp_tasks > 0 && n_cluster_tasks > 0;
    cudaStream_t cl_stream = both ? c->stream2 : c->stream;
    if (both) {
        CPG_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        CPG_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    }
    if (n_cluster_tasks > 0) {
        auto kern = morph_cluster_kernel<S, SHELL, REDUCED, VDISP>;
        const int smem = CLUSTER_KERNEL_WARPS * RingCfg<VDISP>::WARP_BYTES;
        CPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (cluster_size > 8)
            CPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        MorphParams Pc = P;
        Pc.tasks = P.tasks + n_warp_tasks;   // cluster tasks are stored behind the warp tasks
        Pc.n_tasks = n_cluster_tasks;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_cluster_tasks * cluster_size);
        cfg.blockDim = dim3(CLUSTER_KERNEL_WARPS * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = cl_stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster_size;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CPG_CUDA(cudaLaunchKernelEx(&cfg, kern, Pc));
        c->timing.kernel_launches++;
        c->timing.hot_kernel_launches++;
    }
    // kernel 1d keeps two tasks per warp: only worth it when the queue is long enough to keep every slot of the GPU busy
    bool duo = n_warp_tasks > 0 && G == 1 && !VDISP &&
               ((c->opt_async_step == 2 && n_warp_tasks >= 8 * c->sm_count * 12) || c->opt_async_step == 3);
    int duo_per_sm = 0;
    const int duo_smem = WD_WARPS * RingCfg<false, RING_MAX_STAGES>::WARP_BYTES + ((P.R + 7) & ~7) * (int)(sizeof(float) + WD_NSLOT * sizeof(uint16_t));
    if (duo) {
        auto kern = morph_wduo_kernel<S, SHELL, REDUCED, false>;
        CPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, duo_smem));
        CPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&duo_per_sm, kern, WD_WARPS * 32, duo_smem));
        if (getenv("CPG_DEBUG")) fprintf(stderr, "[cpg] morph_wduo_kernel: %d CTAs per SM (smem %d dynamic)\n", duo_per_sm, duo_smem);
        // the radii / prefix-length tables grow with R: above R ~ 100 only two CTAs (8 warps) fit an SM, and kernel 1 with its 12
        // warps is the faster one (measured: 19.7 against 17.5 ms on configs[1])
        if (duo_per_sm < CPG_WD_MINB && c->opt_async_step == 2) duo = false;
    }
    if (duo) {
        auto kern = morph_wduo_kernel<S, SHELL, REDUCED, false>;
        const int smem = duo_smem;
        int per_sm = duo_per_sm;
        if (per_sm < 1) per_sm = 1;
        if (c->opt_warp_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_warp_ctas_per_sm);
        const int grid = std::min((n_warp_tasks + WD_NSLOT - 1) / WD_NSLOT, c->sm_count * per_sm);
        MorphParams Pw = P;
        Pw.n_tasks = n_warp_tasks;
        kern<<<grid, WD_WARPS * 32, smem, c->stream>>>(Pw);
        c->timing.kernel_launches++;
        c->timing.hot_kernel_launches++;
        CPG_CUDA(cudaGetLastError());
    } else if (n_warp_tasks > 0 && G == 1 && !VDISP && c->opt_async_step == 1) {
        auto kern = morph_wasync_kernel<S, SHELL, REDUCED, false>;
        const int smem = WA_NSLOT * RingCfg<false, WA_STAGES>::WARP_BYTES;
        CPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        CPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WA_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        if (c->opt_warp_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_warp_ctas_per_sm);
        const int grid = std::min((n_warp_tasks + WA_NSLOT - 1) / WA_NSLOT, c->sm_count * per_sm);
        MorphParams Pw = P;
        Pw.n_tasks = n_warp_tasks;
        kern<<<grid, WA_THREADS, smem, c->stream>>>(Pw);
        c->timing.kernel_launches++;
        c->timing.hot_kernel_launches++;
        CPG_CUDA(cudaGetLastError());
    } else if (n_warp_tasks > 0 && G == 1 && c->opt_batched_step) {
        auto kern = morph_wbatch_kernel<S, SHELL, REDUCED, VDISP>;
        const int smem = WB_WARPS * RingCfg<VDISP>::WARP_BYTES;
        CPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        CPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WB_WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        const int grid = std::min((n_warp_tasks + WB_WARPS - 1) / WB_WARPS, c->sm_count * per_sm);
        MorphParams Pw = P;
        Pw.n_tasks = n_warp_tasks;
        kern<<<grid, WB_WARPS * 32, smem, c->stream>>>(Pw);
        c->timing.kernel_launches++;
        c->timing.hot_kernel_launches++;
        CPG_CUDA(cudaGetLastError());
    } else if (n_warp_tasks > 0) {
        auto kern = morph_warp_kernel<S, G, SHELL, REDUCED, VDISP>;
        const int smem = WARP_KERNEL_WARPS * RingCfg<VDISP>::WARP_BYTES;
        CPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        CPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARP_KERNEL_WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        if (c->opt_warp_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_warp_ctas_per_sm);
        const int grid = std::min((n_warp_tasks + WARP_KERNEL_WARPS - 1) / WARP_KERNEL_WARPS, c->sm_count * per_sm);
        MorphParams Pw = P;
        Pw.n_tasks = n_warp_tasks;
        kern<<<grid, WARP_KERNEL_WARPS * 32, smem, c->stream>>>(Pw);
        c->timing.kernel_launches++;
        c->timing.hot_kernel_launches++;
        CPG_CUDA(cudaGetLastError());
    }
    if (both) {
        CPG_CUDA(cudaEventRecord(c->ev_join, c->stream2));
        CPG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    return CPG_OK;
}

template <int S, int G>
int dispatch_morph(cpg_ctx *c, const MorphParams &P, bool shell, bool reduced, bool vdisp, int nw, int ncl, int cs)
{
#define CPG_CASE(SH, RE, VD) \
    if (shell == SH && reduced == RE && vdisp == VD) return launch_morph<S, G, SH, RE, VD>(c, P, nw, ncl, cs);
    CPG_CASE(false, false, false)
#ifndef CPG_DEV_MIN   // (development builds: only the E1 non-reduced position variant, a tenth of the compile time)
    CPG_CASE(false, true, false) CPG_CASE(true, false, false) CPG_CASE(true, true, false)
    CPG_CASE(false, false, true) CPG_CASE(false, true, true) CPG_CASE(true, false, true) CPG_CASE(true, true, true)
#endif
#undef CPG_CASE
    return CPG_ERR_ARGUMENT;
}

// Kernel 3 driver: tasks (object, S shells) of huge objects, one after the other; per Katz iteration one pass
// over the whole grid + one step warp.  The host looks at the task's active mask every few iterations.
template <int S, bool SHELL, bool REDUCED, bool VDISP>
int launch_grid(cpg_ctx *c, MorphParams P, int first_task, int n_tasks)
{
    auto pass = grid_pass_kernel<S, SHELL, REDUCED, VDISP>;
    const int smem = GRID_KERNEL_WARPS * RingCfg<VDISP>::WARP_BYTES;
    CPG_CUDA(cudaFuncSetAttribute(pass, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    CPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pass, GRID_KERNEL_WARPS * 32, smem));
    const int n_cta = c->sm_count * std::max(1, per_sm);
    CPG_TRY(c->grid_states.ensure(n_tasks));
    CPG_TRY(c->grid_partials.ensure((size_t)n_cta * 8 * S));
    const Task *tasks = P.tasks + first_task;
    grid_init_kernel<S, SHELL><<<n_tasks, 32, 0, c->stream>>>(P, tasks, c->grid_states.p, n_tasks);
    c->timing.kernel_launches++;
    unsigned *h_amask = nullptr;
    CPG_CUDA(cudaHostAlloc((void **)&h_amask, sizeof(unsigned), cudaHostAllocDefault));
    int rc = CPG_OK;
    for (int t = 0; t < n_tasks && rc == CPG_OK; t++) {
        GridState *gs = c->grid_states.p + t;
        int done = 0;
        while (done <= P.wall + 1) {
            const int batch = std::min(8, P.wall + 2 - done);
            for (int i = 0; i < batch; i++) {
                pass<<<n_cta, GRID_KERNEL_WARPS * 32, smem, c->stream>>>(P, gs, c->grid_partials.p);
                grid_step_kernel<S, SHELL><<<1, 32, 0, c->stream>>>(P, gs, c->grid_partials.p, n_cta);
            }
            c->timing.kernel_launches += 2 * batch;
            c->timing.hot_kernel_launches += batch;
            done += batch;
            if (cudaMemcpyAsync(h_amask, &gs->amask, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess) {
                set_error("grid morphology path failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = CPG_ERR_CUDA;
                break;
            }
            if (*h_amask == 0) break;
        }
    }
    cudaFreeHost(h_amask);
    return rc;
}

template <int S>
int dispatch_grid(cpg_ctx *c, const MorphParams &P, bool shell, bool reduced, bool vdisp, int first_task, int n_tasks)
{
#define CPG_CASE(SH, RE, VD) \
    if (shell == SH && reduced == RE && vdisp == VD) return launch_grid<S, SH, RE, VD>(c, P, first_task, n_tasks);
    CPG_CASE(false, false, false)
#ifndef CPG_DEV_MIN   // (development builds: only the E1 non-reduced position variant, a tenth of the compile time)
    CPG_CASE(false, true, false) CPG_CASE(true, false, false) CPG_CASE(true, true, false)
    CPG_CASE(false, false, true) CPG_CASE(false, true, true) CPG_CASE(true, false, true) CPG_CASE(true, true, true)
#endif
#undef CPG_CASE
    return CPG_ERR_ARGUMENT;
}

int init_context(cpg_ctx *c)
{
    CPG_CUDA(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CPG_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 9) {
        set_error("cpg_create: device %d is sm_%d%d; this build targets sm_100a (thread-block clusters required)",
                  c->device, prop.major, prop.minor);
        return CPG_ERR_NO_DEVICE;
    }
    c->sm_count = prop.multiProcessorCount;
    CPG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CPG_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    for (auto &e : c->ev) CPG_CUDA(cudaEventCreate(&e));
    CPG_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CPG_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    return CPG_OK;
}

int pow2_floor(int64_t v)
{
    int p = 1;
    while ((int64_t)p * 2 <= v) p *= 2;
    return p;
}

}  // namespace

// ------------------------------------------------------------------------------------------ C ABI

extern "C" {

CPG_API int cpg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

CPG_API const char *cpg_last_error(void) { return g_err; }

CPG_API const char *cpg_version(void) { return "cosmicgpu 0.1 (sm_100a)"; }

CPG_API int cpg_create(int device, cpg_ctx **out)
{
    if (!out) {
        set_error("cpg_create: out is NULL");
        return CPG_ERR_ARGUMENT;
    }
    *out = nullptr;
    const int n = cpg_device_count();
    if (n <= 0 || device < 0 || device >= n) {
        set_error("cpg_create: CUDA device %d not available (%d visible); this library has no CPU fallback", device, n);
        return CPG_ERR_NO_DEVICE;
    }
    cpg_ctx *c = new (std::nothrow) cpg_ctx();
    if (!c) return CPG_ERR_NOMEM;
    c->device = device;
    const int rc = init_context(c);
    if (rc != CPG_OK) {   // release whatever was created (streams, events) together with the context
        cpg_destroy(c);
        return rc;
    }
    *out = c;
    return CPG_OK;
}

CPG_API void cpg_destroy(cpg_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    // kernels of a failed call may still be storing into page-locked caller memory: drain every stream first
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    for (auto &us : c->up_stream)
        if (us) cudaStreamSynchronize(us);
    (void)cudaGetLastError();
    DevBuf<double> *dbl[] = {&c->xyz, &c->vxyz, &c->masses, &c->g_dx, &c->g_dy, &c->g_dz, &c->g_m, &c->g_vx, &c->g_vy,
                             &c->g_vz, &c->centers, &c->vcenters, &c->r200, &c->rr, &c->obj_mass, &c->part_mass,
                             &c->vpart, &c->part_mm, &c->obj_eqm, &c->out12, &c->d_a, &c->d_b, &c->d_c, &c->d_major, &c->d_inter, &c->d_minor,
                             &c->d_edges, &c->dens, &c->bin_mass, &c->chunk_mmax};
    for (auto *b : dbl) b->release();
    DevBuf<int32_t> *i32[] = {&c->idx, &c->size, &c->chunk_first, &c->kchunk_first, &c->part_flag, &c->task_counter,
                              &c->n_iter, &c->n_pts, &c->bin_cnt, &c->counts};
    for (auto *b : i32) b->release();
    c->off.release(); c->poff.release(); c->chunks.release(); c->kchunks.release(); c->schunks.release(); c->schunk_first.release();
    c->mass32.release(); c->skip.release(); c->bucket.release();
    for (auto &tk : c->task_cache) tk.tasks.release();
    c->grid_states.release(); c->grid_partials.release();
    c->tile_hist.release(); c->dest.release(); c->n_eff.release(); c->cidxA.release(); c->cidxB.release(); c->d_order.release(); c->ptab.release();
    c->cref.release();
    c->oc_nb.release(); c->oc_shlen.release(); c->oc_fof.release(); c->oc_size.release(); c->oc_idx.release();
    c->oc_err.release(); c->oc_r200in.release(); c->oc_r200.release(); c->oc_shoff.release(); c->oc_start.release();
    c->oc_passoff.release(); c->oc_idxoff.release(); c->oc_cstart.release(); c->oc_coff.release();
    c->oc_tiles.release(); c->raw_block.release();
    c->ip_a.release(); c->ip_b.release(); c->ip_c.release(); c->ip_major.release(); c->ip_inter.release();
    c->ip_minor.release(); c->ip_rr.release(); c->ip_edges.release();
    c->fit_radii.release(); c->fit_lmed.release(); c->fit_avg.release(); c->fit_x0.release(); c->fit_best.release();
    c->fit_fun.release(); c->fit_nvalid.release(); c->fit_nfev.release(); c->fit_nit.release();
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stage) cudaFreeHost(c->stage);
    if (c->up_ready)
        for (int t = 0; t < cpg_ctx::UP_THREADS; t++) {
            for (int b = 0; b < cpg_ctx::UP_BUFS; b++) {
                if (c->up_buf[t][b]) cudaFreeHost(c->up_buf[t][b]);
                if (c->up_ev[t][b]) cudaEventDestroy(c->up_ev[t][b]);
            }
            if (c->up_stream[t]) cudaStreamDestroy(c->up_stream[t]);
        }
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

CPG_API int cpg_set_option(cpg_ctx *c, const char *name, int64_t value)
{
    if (!c || !name) return CPG_ERR_ARGUMENT;
    if (!strcmp(name, "shells_per_pass")) {
        if (value != 0 && value != 1 && value != 2 && value != 4) {
            set_error("shells_per_pass must be 0, 1, 2 or 4");
            return CPG_ERR_ARGUMENT;
        }
        c->opt_shells_per_pass = (int)value;
    } else if (!strcmp(name, "grid_halo_min")) {
        c->opt_grid_halo_min = value;
    } else if (!strcmp(name, "large_halo_min")) {
        c->opt_large_halo_min = value;
    } else if (!strcmp(name, "invalidate_gather")) {
        c->have.valid = false;   // one-shot: the next compute call redoes the catalogue gather
    } else if (!strcmp(name, "shell_groups")) {
        if (value < 0 || value > 2) {
            set_error("shell_groups must be 0, 1 or 2");
            return CPG_ERR_ARGUMENT;
        }
        c->opt_groups = (int)value;
    } else if (!strcmp(name, "upload_threads")) {
        if (value < 0 || value > cpg_ctx::UP_THREADS) {
            set_error("upload_threads must be 0..%d", (int)cpg_ctx::UP_THREADS);
            return CPG_ERR_ARGUMENT;
        }
        c->opt_upload_threads = (int)value;
    } else if (!strcmp(name, "snapshot_cache")) {
        if (value < 0 || value > 2) {
            set_error("snapshot_cache must be 0 (off), 1 (sampled content token) or 2 (full hash)");
            return CPG_ERR_ARGUMENT;
        }
        c->opt_snapshot_cache = (int)value;
        if (value == 0) { c->snap_key.valid = false; c->cat_key.valid = false; }
    } else if (!strcmp(name, "detect_ranges")) {
        c->opt_detect_ranges = value != 0;
    } else if (!strcmp(name, "async_step")) {
        c->opt_async_step = (int)value;   // 0: kernel 1, 1: kernel 1c (step warp), 2: kernel 1d for long queues, 3: kernel 1d always
    } else if (!strcmp(name, "batched_step")) {
        c->opt_batched_step = value != 0;
    } else if (!strcmp(name, "warp_ctas_per_sm")) {
        c->opt_warp_ctas_per_sm = (int)value;
    } else if (!strcmp(name, "direct_results")) {
        c->opt_direct_results = value != 0;
    } else if (!strcmp(name, "radial_order")) {
        c->opt_radial_order = value != 0;
    } else if (!strcmp(name, "prefix_table")) {
        c->opt_prefix_table = value != 0;
    } else if (!strcmp(name, "cluster_size")) {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8 && value != 16) {
            set_error("cluster_size must be 0, 1, 2, 4, 8 or 16");
            return CPG_ERR_ARGUMENT;
        }
        c->opt_cluster_size = (int)value;
    } else {
        set_error("unknown option %s", name);
        return CPG_ERR_ARGUMENT;
    }
    return CPG_OK;
}

// Snapshot bookkeeping shared by the three upload entry points.
static void snapshot_begin(cpg_ctx *c)
{
    // a failed upload must not leave a usable but inconsistent snapshot behind
    c->have.valid = false; c->mass_stream_valid = false;
    c->n_part = 0;
    c->snap_key.valid = false;
}

static void snapshot_done(cpg_ctx *c, int64_t n_part, bool have_vel)
{
    c->n_part = n_part;
    c->have_vel = have_vel;
    c->snapshot_version++;
    // a catalogue installed for an earlier, larger snapshot must not be used against this one (the device buffers
    // never shrink, so the gather would not fault: it would read stale particles)
    c->cat_stale = c->n_obj > 0 && c->cat_max_idx >= n_part;
}

CPG_API int cpg_set_particles(cpg_ctx *c, const double *xyz, const double *vxyz, const double *masses, int64_t n_part)
{
    CPG_TRY(use_device(c));
    if (n_part < 0 || (n_part > 0 && (!xyz || !masses))) {
        set_error("cpg_set_particles: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    // same arrays as last time?  (pointer, length and a content token; see "snapshot_cache")
    uint64_t tok[3] = {0, 0, 0};
    const int mode = c->opt_snapshot_cache;
    if (mode > 0) {
        tok[0] = host_token(xyz, sizeof(double) * 3 * (size_t)n_part, mode == 2);
        tok[1] = host_token(masses, sizeof(double) * (size_t)n_part, mode == 2);
        tok[2] = vxyz ? host_token(vxyz, sizeof(double) * 3 * (size_t)n_part, mode == 2) : 0;
        auto &K = c->snap_key;
        if (K.valid && K.mode >= mode && K.xyz == xyz && K.vxyz == vxyz && K.masses == masses && K.n == n_part &&
            K.tok[0] == tok[0] && K.tok[1] == tok[1] && K.tok[2] == tok[2] && c->n_part == n_part) {
            c->uploads_skipped++;
            return CPG_OK;
        }
    }
    snapshot_begin(c);
    CPG_TRY(c->xyz.ensure((size_t)n_part * 3));
    CPG_TRY(c->masses.ensure((size_t)n_part));
    CPG_TRY(upload_chunked(c, c->xyz.p, xyz, sizeof(double) * 3 * n_part, c->stream));
    CPG_TRY(upload_chunked(c, c->masses.p, masses, sizeof(double) * n_part, c->stream));
    if (vxyz) {
        CPG_TRY(c->vxyz.ensure((size_t)n_part * 3));
        CPG_TRY(upload_chunked(c, c->vxyz.p, vxyz, sizeof(double) * 3 * n_part, c->stream));
    }
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    snapshot_done(c, n_part, vxyz != nullptr);
    if (mode > 0) {
        auto &K = c->snap_key;
        K.valid = true; K.mode = mode; K.xyz = xyz; K.vxyz = vxyz; K.masses = masses; K.n = n_part;
        K.tok[0] = tok[0]; K.tok[1] = tok[1]; K.tok[2] = tok[2];
    }
    return CPG_OK;
}

CPG_API int cpg_set_particles_segments(cpg_ctx *c, const double *xyz, const double *vxyz, const double *masses,
                                       const int64_t *seg_start, const int64_t *seg_len, int64_t n_seg)
{
    CPG_TRY(use_device(c));
    if (n_seg < 0 || (n_seg > 0 && (!xyz || !masses || !seg_start || !seg_len))) {
        set_error("cpg_set_particles_segments: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    std::vector<int64_t> prefix((size_t)n_seg + 1, 0);
    for (int64_t i = 0; i < n_seg; i++) {
        if (seg_start[i] < 0 || seg_len[i] < 0) {
            set_error("cpg_set_particles_segments: segment %lld has a negative start or length", (long long)i);
            return CPG_ERR_ARGUMENT;
        }
        prefix[i + 1] = prefix[i] + seg_len[i];
    }
    const int64_t n_part = prefix[n_seg];
    snapshot_begin(c);
    CPG_TRY(c->xyz.ensure((size_t)n_part * 3));
    CPG_TRY(c->masses.ensure((size_t)n_part));
    if (vxyz) CPG_TRY(c->vxyz.ensure((size_t)n_part * 3));
    CPG_CUDA(cudaStreamSynchronize(c->stream));   // earlier work may still read the old snapshot
    CPG_TRY(upload_segments(c, c->xyz.p, xyz, 3 * sizeof(double), seg_start, seg_len, n_seg, prefix));
    CPG_TRY(upload_segments(c, c->masses.p, masses, sizeof(double), seg_start, seg_len, n_seg, prefix));
    if (vxyz) CPG_TRY(upload_segments(c, c->vxyz.p, vxyz, 3 * sizeof(double), seg_start, seg_len, n_seg, prefix));
    snapshot_done(c, n_part, vxyz != nullptr);
    return CPG_OK;
}

// Catalogue bookkeeping shared by cpg_set_catalogue (idx_cat on the host), cpg_set_catalogue_ranges (no index array)
// and cpg_use_obj_cat (idx_cat already on the device in c->idx: idx_cat == nullptr, obj_start == nullptr).
// scan: what scan_catalogue found out about a host idx_cat (nullptr: not scanned).
static int install_catalogue(cpg_ctx *c, const int32_t *idx_cat, const int64_t *obj_start, int64_t n_idx,
                             const int32_t *obj_size, int32_t n_obj, const CatScan *scan = nullptr)
{
    // the flat catalogue is by far the largest piece: start its upload first, the host-side bookkeeping below then
    // runs while the copy engine works (an inconsistent obj_size is found afterwards and invalidates the catalogue)
    c->n_idx = 0; c->n_obj = 0; c->have.valid = false; c->cat_key.valid = false; c->cat_stale = false;
    for (auto &tk : c->task_cache) tk.valid = false;
    const bool ranges = obj_start != nullptr;
    if (idx_cat && !ranges) {
        CPG_TRY(c->idx.ensure(n_idx));
        if (n_idx) CPG_TRY(upload_chunked(c, c->idx.p, idx_cat, sizeof(int32_t) * n_idx, c->stream));
    }
    std::vector<int64_t> off(n_obj + 1, 0), poff(n_obj + 1, 0);
    c->h_size.assign(obj_size, obj_size + n_obj);
    c->h_chunks.clear(); c->h_kchunks.clear(); c->h_schunks.clear(); c->h_big_sort.clear(); c->h_big_stat.clear();
    c->h_chunk_first.assign(n_obj + 1, 0); c->h_kchunk_first.assign(n_obj + 1, 0); c->h_schunk_first.assign(n_obj + 1, 0);
    int64_t max_idx = -1, min_idx = INT64_MAX;
    for (int h = 0; h < n_obj; h++) {
        if (obj_size[h] < 0) {
            set_error("cpg_set_catalogue: obj_size[%d] < 0", h);
            return CPG_ERR_ARGUMENT;
        }
        off[h + 1] = off[h] + obj_size[h];
        poff[h + 1] = poff[h] + ((obj_size[h] + 1) & ~1);
        for (int s = 0; s < obj_size[h]; s += CHUNK) c->h_chunks.push_back(Chunk{h, s});
        for (int s = 0; s < obj_size[h]; s += KDE_CHUNK) c->h_kchunks.push_back(Chunk{h, s});
        for (int s = 0; s < obj_size[h]; s += SORT_CHUNK) c->h_schunks.push_back(Chunk{h, s});
        if ((int)c->h_chunks.size() - c->h_chunk_first[h] > BIG_OBJECT_CHUNKS) c->h_big_stat.push_back(h);
        if ((int)c->h_schunks.size() - c->h_schunk_first[h] > BIG_OBJECT_CHUNKS) c->h_big_sort.push_back(h);
        c->h_chunk_first[h + 1] = (int32_t)c->h_chunks.size();
        c->h_kchunk_first[h + 1] = (int32_t)c->h_kchunks.size();
        c->h_schunk_first[h + 1] = (int32_t)c->h_schunks.size();
        if (ranges && obj_size[h] > 0) {
            min_idx = std::min(min_idx, obj_start[h]);
            max_idx = std::max(max_idx, obj_start[h] + obj_size[h] - 1);
        }
    }
    if (off[n_obj] != n_idx) {
        set_error("cpg_set_catalogue: sum(obj_size) = %lld but idx_cat has %lld entries", (long long)off[n_obj],
                  (long long)n_idx);
        return CPG_ERR_ARGUMENT;
    }
    c->order.resize(n_obj);
    std::iota(c->order.begin(), c->order.end(), 0);
    std::stable_sort(c->order.begin(), c->order.end(), [&](int a, int b) { return obj_size[a] > obj_size[b]; });

    CPG_TRY(c->size.ensure(n_obj));
    CPG_TRY(c->off.ensure(n_obj + 1)); CPG_TRY(c->poff.ensure(n_obj + 1));
    CPG_TRY(c->chunks.ensure(c->h_chunks.size())); CPG_TRY(c->kchunks.ensure(c->h_kchunks.size()));
    CPG_TRY(c->schunks.ensure(c->h_schunks.size()));
    CPG_TRY(c->chunk_first.ensure(n_obj + 1)); CPG_TRY(c->kchunk_first.ensure(n_obj + 1)); CPG_TRY(c->schunk_first.ensure(n_obj + 1));
    CPG_TRY(c->task_counter.ensure(4));
    CPG_TRY(c->d_order.ensure(n_obj));
    if (ranges) {
        CPG_TRY(c->start.ensure(n_obj));
        if (n_obj)
            CPG_CUDA(cudaMemcpyAsync(c->start.p, obj_start, sizeof(int64_t) * n_obj, cudaMemcpyHostToDevice, c->stream));
    }
    if (n_obj)
        CPG_CUDA(cudaMemcpyAsync(c->d_order.p, c->order.data(), sizeof(int32_t) * n_obj, cudaMemcpyHostToDevice, c->stream));
    if (n_obj) CPG_CUDA(cudaMemcpyAsync(c->size.p, obj_size, sizeof(int32_t) * n_obj, cudaMemcpyHostToDevice, c->stream));
    CPG_CUDA(cudaMemcpyAsync(c->off.p, off.data(), sizeof(int64_t) * (n_obj + 1), cudaMemcpyHostToDevice, c->stream));
    CPG_CUDA(cudaMemcpyAsync(c->poff.p, poff.data(), sizeof(int64_t) * (n_obj + 1), cudaMemcpyHostToDevice, c->stream));
    if (!c->h_chunks.empty())
        CPG_CUDA(cudaMemcpyAsync(c->chunks.p, c->h_chunks.data(), sizeof(Chunk) * c->h_chunks.size(),
                                 cudaMemcpyHostToDevice, c->stream));
    if (!c->h_kchunks.empty())
        CPG_CUDA(cudaMemcpyAsync(c->kchunks.p, c->h_kchunks.data(), sizeof(Chunk) * c->h_kchunks.size(),
                                 cudaMemcpyHostToDevice, c->stream));
    CPG_CUDA(cudaMemcpyAsync(c->chunk_first.p, c->h_chunk_first.data(), sizeof(int32_t) * (n_obj + 1),
                             cudaMemcpyHostToDevice, c->stream));
    CPG_CUDA(cudaMemcpyAsync(c->kchunk_first.p, c->h_kchunk_first.data(), sizeof(int32_t) * (n_obj + 1),
                             cudaMemcpyHostToDevice, c->stream));
    if (!c->h_schunks.empty())
        CPG_CUDA(cudaMemcpyAsync(c->schunks.p, c->h_schunks.data(), sizeof(Chunk) * c->h_schunks.size(),
                                 cudaMemcpyHostToDevice, c->stream));
    CPG_CUDA(cudaMemcpyAsync(c->schunk_first.p, c->h_schunk_first.data(), sizeof(int32_t) * (n_obj + 1),
                             cudaMemcpyHostToDevice, c->stream));
    // index range check: an out-of-range index would be an illegal address in the gather
    if (scan) {
        min_idx = scan->min_idx; max_idx = scan->max_idx;
    } else if (!ranges && n_idx > 0) {   // on the device (the flat array is already there)
        const int32_t init[2] = {0x7fffffff, -0x7fffffff - 1};
        int32_t got[2];
        CPG_CUDA(cudaMemcpyAsync(c->task_counter.p, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
        idx_range_kernel<<<grid_for(n_idx, 256, c->sm_count * 16), 256, 0, c->stream>>>(c->idx.p, n_idx, c->task_counter.p);
        CPG_CUDA(cudaGetLastError());
        CPG_CUDA(cudaMemcpyAsync(got, c->task_counter.p, sizeof got, cudaMemcpyDeviceToHost, c->stream));
        CPG_CUDA(cudaStreamSynchronize(c->stream));
        min_idx = got[0]; max_idx = got[1];
    }
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    if (n_idx > 0 && (min_idx < 0 || max_idx >= c->n_part)) {
        set_error("cpg_set_catalogue: idx_cat range [%lld, %lld] outside the %lld uploaded particles", (long long)min_idx,
                  (long long)max_idx, (long long)c->n_part);
        return CPG_ERR_ARGUMENT;
    }
    c->n_idx = n_idx;
    c->n_pad = poff[n_obj];
    c->n_obj = n_obj;
    c->cat_ranges = ranges;
    c->cat_max_idx = n_idx > 0 ? max_idx : -1;
    c->mass_stream_valid = false;
    c->catalogue_version++;
    c->have.valid = false;
    for (auto &tk : c->task_cache) tk.valid = false;
    return CPG_OK;
}

static bool same_sizes(const std::vector<int32_t> &a, const int32_t *b, int32_t n)
{
    return a.size() == (size_t)n && (n == 0 || !memcmp(a.data(), b, sizeof(int32_t) * (size_t)n));
}

// Offsets of a flat catalogue (validated); false + error message on inconsistent sizes.
static bool catalogue_offsets(const char *who, const int32_t *obj_size, int32_t n_obj, int64_t n_idx, std::vector<int64_t> &off)
{
    off.assign((size_t)n_obj + 1, 0);
    for (int h = 0; h < n_obj; h++) {
        if (obj_size[h] < 0) {
            set_error("%s: obj_size[%d] < 0", who, h);
            return false;
        }
        off[h + 1] = off[h] + obj_size[h];
    }
    if (off[n_obj] != n_idx) {
        set_error("%s: sum(obj_size) = %lld but idx_cat has %lld entries", who, (long long)off[n_obj], (long long)n_idx);
        return false;
    }
    return true;
}

// cpg_set_catalogue proper; `sc` = the result of scan_catalogue on idx_cat (contiguity, index range, content hash).
static int set_catalogue_scanned(cpg_ctx *c, const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size, int32_t n_obj,
                                 const CatScan &sc)
{
    const bool ranges = c->opt_detect_ranges && sc.contiguous;
    auto &K = c->cat_key;
    if (c->opt_snapshot_cache && K.valid && !c->cat_stale && K.ranges == ranges && K.n_idx == n_idx && K.hash == sc.hash &&
        same_sizes(K.size, obj_size, n_obj) && c->n_obj == n_obj && c->cat_max_idx < c->n_part) {
        c->uploads_skipped++;
        return CPG_OK;
    }
    const int rc = install_catalogue(c, idx_cat, ranges ? sc.start.data() : nullptr, n_idx, obj_size, n_obj, &sc);
    if (rc == CPG_OK && c->opt_snapshot_cache) {
        K.valid = true; K.ranges = ranges; K.n_idx = n_idx; K.hash = sc.hash;
        K.size.assign(obj_size, obj_size + n_obj);
    }
    return rc;
}

CPG_API int cpg_set_catalogue(cpg_ctx *c, const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size, int32_t n_obj)
{
    CPG_TRY(use_device(c));
    if (n_idx < 0 || n_obj < 0 || (n_idx > 0 && !idx_cat) || (n_obj > 0 && !obj_size)) {
        set_error("cpg_set_catalogue: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    if (!c->opt_detect_ranges && !c->opt_snapshot_cache) return install_catalogue(c, idx_cat, nullptr, n_idx, obj_size, n_obj);
    // one multi-threaded pass over the flat array: contiguous-range detection, index range, content hash
    std::vector<int64_t> off;
    if (!catalogue_offsets("cpg_set_catalogue", obj_size, n_obj, n_idx, off)) return CPG_ERR_ARGUMENT;
    CatScan sc;
    scan_catalogue(idx_cat, off.data(), n_obj, sc);
    return set_catalogue_scanned(c, idx_cat, n_idx, obj_size, n_obj, sc);
}

// cpg_set_particles + cpg_set_catalogue in one call: the host-side pass over the flat catalogue runs on its own threads
// WHILE the particle arrays cross PCIe (a page-locked source is pure DMA: the scan is free).
CPG_API int cpg_set_snapshot(cpg_ctx *c, const double *xyz, const double *vxyz, const double *masses, int64_t n_part,
                             const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size, int32_t n_obj)
{
    CPG_TRY(use_device(c));
    if (n_idx < 0 || n_obj < 0 || (n_idx > 0 && !idx_cat) || (n_obj > 0 && !obj_size)) {
        set_error("cpg_set_snapshot: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    const bool scan = c->opt_detect_ranges || c->opt_snapshot_cache;
    std::vector<int64_t> off;
    CatScan sc;
    std::thread worker;
    if (scan) {
        if (!catalogue_offsets("cpg_set_snapshot", obj_size, n_obj, n_idx, off)) return CPG_ERR_ARGUMENT;
        worker = std::thread([&]() { scan_catalogue(idx_cat, off.data(), n_obj, sc); });
    }
    const int rc = cpg_set_particles(c, xyz, vxyz, masses, n_part);
    if (worker.joinable()) worker.join();
    CPG_TRY(rc);
    if (!scan) return install_catalogue(c, idx_cat, nullptr, n_idx, obj_size, n_obj);
    return set_catalogue_scanned(c, idx_cat, n_idx, obj_size, n_obj, sc);
}

CPG_API int cpg_set_catalogue_ranges(cpg_ctx *c, const int64_t *obj_start, const int32_t *obj_size, int32_t n_obj)
{
    CPG_TRY(use_device(c));
    if (n_obj < 0 || (n_obj > 0 && (!obj_start || !obj_size))) {
        set_error("cpg_set_catalogue_ranges: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    int64_t n_idx = 0;
    for (int h = 0; h < n_obj; h++) {
        if (obj_size[h] < 0 || (obj_size[h] > 0 && obj_start[h] < 0)) {
            set_error("cpg_set_catalogue_ranges: object %d has a negative start or size", h);
            return CPG_ERR_ARGUMENT;
        }
        n_idx += obj_size[h];
    }
    static const int64_t none = 0;
    return install_catalogue(c, nullptr, n_obj ? obj_start : &none, n_idx, obj_size, n_obj);
}

CPG_API int cpg_catalogue_ranges(const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size, int32_t n_obj,
                                 int64_t *obj_start_out)
{
    if (n_idx < 0 || n_obj < 0 || (n_idx > 0 && !idx_cat) || (n_obj > 0 && (!obj_size || !obj_start_out))) {
        set_error("cpg_catalogue_ranges: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    std::vector<int64_t> off((size_t)n_obj + 1, 0);
    for (int h = 0; h < n_obj; h++) {
        if (obj_size[h] < 0) {
            set_error("cpg_catalogue_ranges: obj_size[%d] < 0", h);
            return CPG_ERR_ARGUMENT;
        }
        off[h + 1] = off[h] + obj_size[h];
    }
    if (off[n_obj] != n_idx) {
        set_error("cpg_catalogue_ranges: sum(obj_size) = %lld but idx_cat has %lld entries", (long long)off[n_obj], (long long)n_idx);
        return CPG_ERR_ARGUMENT;
    }
    CatScan sc;
    scan_catalogue(idx_cat, off.data(), n_obj, sc);
    for (int h = 0; h < n_obj; h++) obj_start_out[h] = sc.start[h];
    return sc.contiguous ? 1 : 0;
}

CPG_API int cpg_host_token(const void *p, size_t bytes, int32_t full, uint64_t *token_out)
{
    if (!token_out || (bytes > 0 && !p)) return CPG_ERR_ARGUMENT;
    *token_out = host_token(p, bytes, full != 0);
    return CPG_OK;
}

// ---- Gadget / FoF ingest ----------------------------------------------------------

}  // extern "C"

template <class Load>
static int device_scan(cpg_ctx *c, Load ld, int64_t n, int64_t *out_a, int64_t *out_b, Pair64 *h_total)
{
    const int n_tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    CPG_TRY(c->oc_tiles.ensure((size_t)n_tiles + 1));
    if (n_tiles > 0) {
        scan_tile_sum_kernel<Load><<<n_tiles, SCAN_THREADS, 0, c->stream>>>(ld, n, c->oc_tiles.p);
        scan_top_kernel<<<1, SCAN_THREADS, 0, c->stream>>>(c->oc_tiles.p, n_tiles);
        scan_apply_kernel<Load><<<n_tiles, SCAN_THREADS, 0, c->stream>>>(ld, n, c->oc_tiles.p, out_a, out_b);
        c->timing.kernel_launches += 3;
        c->timing.hot_kernel_launches += 3;
        CPG_CUDA(cudaGetLastError());
        CPG_CUDA(cudaMemcpyAsync(h_total, c->oc_tiles.p + n_tiles, sizeof(Pair64), cudaMemcpyDeviceToHost, c->stream));
        CPG_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        *h_total = Pair64{0, 0};
    }
    return CPG_OK;
}

extern "C" {

CPG_API int cpg_obj_cat(cpg_ctx *c, const int32_t *nb_shs, const int32_t *sh_len, const int32_t *fof_sizes,
                        const double *group_r200, int32_t n_fof, int64_t n_sh, int32_t min_number_ptcs,
                        int32_t *n_obj_out, int64_t *n_idx_out)
{
    CPG_TRY(use_device(c));
    if (n_fof < 0 || n_sh < 0 || !n_obj_out || !n_idx_out || (n_fof > 0 && (!nb_shs || !fof_sizes || !group_r200)) ||
        (n_sh > 0 && !sh_len)) {
        set_error("cpg_obj_cat: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    c->oc_valid = false;
    Timer T(c);
    T.mark(0);
    const size_t n = (size_t)n_fof;
    CPG_TRY(c->oc_nb.ensure(n)); CPG_TRY(c->oc_fof.ensure(n)); CPG_TRY(c->oc_r200in.ensure(n));
    CPG_TRY(c->oc_shlen.ensure((size_t)n_sh)); CPG_TRY(c->oc_err.ensure(1));
    CPG_TRY(c->oc_shoff.ensure(n)); CPG_TRY(c->oc_start.ensure(n)); CPG_TRY(c->oc_passoff.ensure(n));
    CPG_TRY(c->oc_idxoff.ensure(n));
    if (n) {
        CPG_CUDA(cudaMemcpyAsync(c->oc_nb.p, nb_shs, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
        CPG_CUDA(cudaMemcpyAsync(c->oc_fof.p, fof_sizes, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
        CPG_CUDA(cudaMemcpyAsync(c->oc_r200in.p, group_r200, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    }
    if (n_sh) CPG_CUDA(cudaMemcpyAsync(c->oc_shlen.p, sh_len, sizeof(int32_t) * n_sh, cudaMemcpyHostToDevice, c->stream));
    CPG_CUDA(cudaMemsetAsync(c->oc_err.p, 0, sizeof(int32_t), c->stream));
    c->timing.h2d_bytes = (int64_t)(n * 16 + (size_t)n_sh * 4);
    T.mark(1);
    T.mark(2);
    Pair64 tot1, tot2;
    CPG_TRY(device_scan(c, LoadGroup{c->oc_nb.p, c->oc_fof.p}, n_fof, c->oc_shoff.p, c->oc_start.p, &tot1));
    CPG_TRY(device_scan(c, LoadPass{c->oc_nb.p, c->oc_shlen.p, c->oc_shoff.p, n_sh, min_number_ptcs}, n_fof,
                        c->oc_passoff.p, c->oc_idxoff.p, &tot2));
    const int64_t n_pass = tot2.a, n_idx = tot2.b;
    if (n_pass > 0x7fffffff) {
        set_error("cpg_obj_cat: %lld objects do not fit the int32 object count", (long long)n_pass);
        return CPG_ERR_ARGUMENT;
    }
    CPG_TRY(c->oc_cstart.ensure((size_t)n_pass)); CPG_TRY(c->oc_coff.ensure((size_t)n_pass + 1));
    CPG_TRY(c->oc_size.ensure((size_t)n_pass)); CPG_TRY(c->oc_r200.ensure((size_t)n_pass));
    CPG_TRY(c->oc_idx.ensure((size_t)n_idx));
    ObjCatParams P;
    P.nb_shs = c->oc_nb.p; P.sh_len = c->oc_shlen.p; P.group_r200 = c->oc_r200in.p;
    P.sh_off = c->oc_shoff.p; P.start = c->oc_start.p; P.pass_off = c->oc_passoff.p; P.idx_off = c->oc_idxoff.p;
    P.n_fof = n_fof; P.n_sh = n_sh; P.n_part_limit = (int64_t)1 << 31; P.min_ptcs = min_number_ptcs;
    P.c_start = c->oc_cstart.p; P.c_off = c->oc_coff.p; P.c_size = c->oc_size.p; P.c_r200 = c->oc_r200.p;
    P.error = c->oc_err.p;
    objcat_compact_kernel<<<grid_for(std::max<int64_t>(n_fof, 1), 256, 1 << 30), 256, 0, c->stream>>>(P, n_pass, n_idx);
    c->timing.kernel_launches++;
    c->timing.hot_kernel_launches++;
    if (n_idx > 0) {
        objcat_fill_kernel<<<grid_for(n_idx, 256, c->sm_count * 16), 256, 0, c->stream>>>(c->oc_coff.p, c->oc_cstart.p, n_pass,
                                                                                         n_idx, c->oc_idx.p);
        c->timing.kernel_launches++;
        c->timing.hot_kernel_launches++;
    }
    CPG_CUDA(cudaGetLastError());
    T.mark(3);
    int32_t err = 0;
    CPG_CUDA(cudaMemcpyAsync(&err, c->oc_err.p, sizeof err, cudaMemcpyDeviceToHost, c->stream));
    T.mark(4);
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    c->timing.h2d_ms = Timer::ms(c->ev[0], c->ev[1]);
    c->timing.kernel_ms = Timer::ms(c->ev[2], c->ev[3]);
    c->timing.d2h_ms = Timer::ms(c->ev[3], c->ev[4]);
    if (err) {
        set_error("cpg_obj_cat: a central subhalo reaches past particle 2^31 - 1 (int32 index catalogue, "
                  "gen_catalogues.pyx:12) or the group tables are inconsistent");
        return CPG_ERR_ARGUMENT;
    }
    c->oc_n_pass = n_pass; c->oc_n_idx = n_idx; c->oc_valid = true;
    *n_obj_out = (int32_t)n_pass;
    *n_idx_out = n_idx;
    (void)tot1;
    return CPG_OK;
}

CPG_API int cpg_obj_cat_fetch(cpg_ctx *c, int32_t *idx_cat, double *obj_r200, int32_t *obj_size)
{
    CPG_TRY(use_device(c));
    if (!c->oc_valid) {
        set_error("cpg_obj_cat_fetch: no catalogue built (call cpg_obj_cat first)");
        return CPG_ERR_STATE;
    }
    const size_t np = (size_t)c->oc_n_pass, ni = (size_t)c->oc_n_idx;
    if (idx_cat && ni) CPG_CUDA(cudaMemcpyAsync(idx_cat, c->oc_idx.p, sizeof(int32_t) * ni, cudaMemcpyDeviceToHost, c->stream));
    if (obj_r200 && np) CPG_CUDA(cudaMemcpyAsync(obj_r200, c->oc_r200.p, sizeof(double) * np, cudaMemcpyDeviceToHost, c->stream));
    if (obj_size && np) CPG_CUDA(cudaMemcpyAsync(obj_size, c->oc_size.p, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, c->stream));
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    return CPG_OK;
}

CPG_API int cpg_use_obj_cat(cpg_ctx *c)
{
    CPG_TRY(use_device(c));
    if (!c->oc_valid) {
        set_error("cpg_use_obj_cat: no catalogue built (call cpg_obj_cat first)");
        return CPG_ERR_STATE;
    }
    std::vector<int32_t> sizes((size_t)c->oc_n_pass);
    if (c->oc_n_pass)
        CPG_CUDA(cudaMemcpyAsync(sizes.data(), c->oc_size.p, sizeof(int32_t) * sizes.size(), cudaMemcpyDeviceToHost, c->stream));
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    // the flat catalogue stays where it was built: it becomes the context's idx buffer (no copy, no PCIe traffic)
    CPG_TRY(c->idx.ensure((size_t)c->oc_n_idx));
    if (c->oc_n_idx)
        CPG_CUDA(cudaMemcpyAsync(c->idx.p, c->oc_idx.p, sizeof(int32_t) * (size_t)c->oc_n_idx, cudaMemcpyDeviceToDevice, c->stream));
    return install_catalogue(c, nullptr, nullptr, c->oc_n_idx, sizes.data(), (int32_t)c->oc_n_pass);
}

}  // extern "C"

template <class T>
static int convert_block(cpg_ctx *c, const void *src, double *dst, int64_t n, const BlockConvert &cv)
{
    // staged through a device copy of the raw block in pieces, so that the fp64 result is the only full-size buffer
    const int64_t piece = (int64_t)16 << 20;   // elements
    CPG_TRY(c->raw_block.ensure((size_t)std::min(piece, std::max<int64_t>(n, 1)) * sizeof(T)));
    for (int64_t o = 0; o < n; o += piece) {
        const int64_t m = std::min(piece, n - o);
        CPG_CUDA(cudaMemcpyAsync(c->raw_block.p, (const T *)src + o, sizeof(T) * m, cudaMemcpyHostToDevice, c->stream));
        block_convert_kernel<T><<<grid_for(m, 256, c->sm_count * 16), 256, 0, c->stream>>>((const T *)c->raw_block.p, dst + o, m, cv);
        c->timing.kernel_launches++;
    }
    CPG_CUDA(cudaGetLastError());
    return CPG_OK;
}

static int ingest_block(cpg_ctx *c, const void *src, int32_t is_f32, double *dst, int64_t n, double unit_fac, int has_post,
                        double post_mul)
{
    BlockConvert cv;
    cv.unit_fac = unit_fac; cv.post_mul = post_mul; cv.has_post = has_post;
    if (is_f32) return convert_block<float>(c, src, dst, n, cv);
    return convert_block<double>(c, src, dst, n, cv);
}

extern "C" {

CPG_API int cpg_set_particles_gadget(cpg_ctx *c, const void *pos, int32_t pos_is_f32, const void *vel, int32_t vel_is_f32,
                                     const void *mass, int32_t mass_is_f32, double mass_table, int64_t n_part,
                                     double length_unit_fac, double mass_unit_fac, double velocity_unit_fac, double time)
{
    CPG_TRY(use_device(c));
    if (n_part < 0 || (n_part > 0 && !pos) || !(length_unit_fac != 0.0) || !(mass_unit_fac != 0.0) ||
        (vel && !(velocity_unit_fac != 0.0))) {
        set_error("cpg_set_particles_gadget: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    if (!mass && mass_table == 0.0) {   // readgadget.py:109-110: 'Problem reading the block MASS'
        set_error("cpg_set_particles_gadget: no Masses block and MassTable entry is 0");
        return CPG_ERR_ARGUMENT;
    }
    Timer T(c);
    T.mark(0);
    snapshot_begin(c);
    CPG_TRY(c->xyz.ensure((size_t)n_part * 3));
    CPG_TRY(c->masses.ensure((size_t)n_part));
    CPG_TRY(ingest_block(c, pos, pos_is_f32, c->xyz.p, n_part * 3, length_unit_fac, 0, 1.0));
    c->timing.h2d_bytes = n_part * 3 * (pos_is_f32 ? 4 : 8);
    if (mass) {
        CPG_TRY(ingest_block(c, mass, mass_is_f32, c->masses.p, n_part, mass_unit_fac, 0, 1.0));
        c->timing.h2d_bytes += n_part * (mass_is_f32 ? 4 : 8);
    } else if (n_part > 0) {
        // np.ones(N, float64) * Masses[ptype] / unit_fac  (readgadget.py:108)
        block_fill_kernel<<<grid_for(n_part, 256, c->sm_count * 16), 256, 0, c->stream>>>(c->masses.p, n_part,
                                                                                         (1.0 * mass_table) / mass_unit_fac);
        c->timing.kernel_launches++;
    }
    if (vel) {
        CPG_TRY(c->vxyz.ensure((size_t)n_part * 3));
        // array *= np.sqrt(head.time)  (readgadget.py:116)
        CPG_TRY(ingest_block(c, vel, vel_is_f32, c->vxyz.p, n_part * 3, velocity_unit_fac, 1, sqrt(time)));
        c->timing.h2d_bytes += n_part * 3 * (vel_is_f32 ? 4 : 8);
    }
    T.mark(1);
    CPG_CUDA(cudaGetLastError());
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    c->timing.h2d_ms = Timer::ms(c->ev[0], c->ev[1]);
    snapshot_done(c, n_part, vel != nullptr);
    return CPG_OK;
}

CPG_API int cpg_fit_profiles(cpg_ctx *c, const double *radii, const double *lmed, const int32_t *n_valid, const double *avg,
                             const double *x0, const double *lb, const double *ub, int32_t n_obj, int32_t R,
                             int32_t profile, double xtol, double ftol, double *best, double *fun, int32_t *nfev,
                             int32_t *nit)
{
    CPG_TRY(use_device(c));
    const int npar = profile == FIT_EINASTO ? 3 : (profile == FIT_ABG ? 5 : 2);
    if (n_obj < 0 || R < 0 || R > FIT_MAX_BINS || profile < 0 || profile > FIT_ABG || !lb || !ub ||
        (n_obj > 0 && (!radii || !lmed || !n_valid || !avg || !x0 || !best))) {
        set_error("cpg_fit_profiles: bad arguments (0 <= R <= %d, profile 0..3)", FIT_MAX_BINS);
        return CPG_ERR_ARGUMENT;
    }
    for (int i = 0; i < npar; i++)
        if (!(lb[i] <= ub[i])) {   // scipy _validate_bounds -> ValueError
            set_error("cpg_fit_profiles: an upper bound is less than the corresponding lower bound");
            return CPG_ERR_ARGUMENT;
        }
    Timer T(c);
    if (n_obj == 0) return CPG_OK;
    const size_t n = (size_t)n_obj, nR = n * (size_t)std::max(R, 1);
    T.mark(0);
    CPG_TRY(c->fit_radii.ensure(nR)); CPG_TRY(c->fit_lmed.ensure(nR)); CPG_TRY(c->fit_avg.ensure(n));
    CPG_TRY(c->fit_x0.ensure(n * npar)); CPG_TRY(c->fit_best.ensure(n * npar)); CPG_TRY(c->fit_fun.ensure(n));
    CPG_TRY(c->fit_nvalid.ensure(n)); CPG_TRY(c->fit_nfev.ensure(n)); CPG_TRY(c->fit_nit.ensure(n));
    if (R > 0) {
        CPG_CUDA(cudaMemcpyAsync(c->fit_radii.p, radii, sizeof(double) * n * R, cudaMemcpyHostToDevice, c->stream));
        CPG_CUDA(cudaMemcpyAsync(c->fit_lmed.p, lmed, sizeof(double) * n * R, cudaMemcpyHostToDevice, c->stream));
    }
    CPG_CUDA(cudaMemcpyAsync(c->fit_avg.p, avg, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CPG_CUDA(cudaMemcpyAsync(c->fit_x0.p, x0, sizeof(double) * n * npar, cudaMemcpyHostToDevice, c->stream));
    CPG_CUDA(cudaMemcpyAsync(c->fit_nvalid.p, n_valid, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    c->timing.h2d_bytes = (int64_t)(sizeof(double) * (2 * n * R + n + n * npar) + sizeof(int32_t) * n);
    T.mark(1);
    T.mark(2);
    FitParams Q;
    Q.radii = c->fit_radii.p; Q.lmed = c->fit_lmed.p; Q.n_valid = c->fit_nvalid.p; Q.avg = c->fit_avg.p; Q.x0 = c->fit_x0.p;
    for (int i = 0; i < 5; i++) {
        Q.lb[i] = i < npar ? lb[i] : 0.0;
        Q.ub[i] = i < npar ? ub[i] : 0.0;
    }
    Q.n_obj = n_obj; Q.R = R; Q.profile = profile; Q.xtol = xtol; Q.ftol = ftol;
    Q.best = c->fit_best.p; Q.fun = c->fit_fun.p; Q.nfev = c->fit_nfev.p; Q.nit = c->fit_nit.p;
    // one warp per object.  (Two warps per object for more than 32 bins -- one bin per lane in every evaluation, a named
    // barrier per evaluation -- were measured on configs[2], 40 bins: slower for every model, 9.7 against 6.9 ms for
    // NFW: half as many objects in flight outweighs the shorter evaluation.  The template parameter is kept.)
    const int W = 1;
    const int grid = (n_obj + FIT_WARPS / W - 1) / (FIT_WARPS / W);
    const int fit_smem = 3 * (FIT_WARPS / W) * ((R + 1) & ~1) * (int)sizeof(double);   // 12 KB at 128 bins, 96 KB at 1024
#define CPG_FIT_LAUNCH(NP, WW)                                                                                              \
    do {                                                                                                                    \
        if (fit_smem > 48 * 1024)                                                                                           \
            CPG_CUDA(cudaFuncSetAttribute(fit_profiles_kernel<NP, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, fit_smem)); \
        fit_profiles_kernel<NP, WW><<<grid, FIT_WARPS * 32, fit_smem, c->stream>>>(Q);                                      \
    } while (0)
    if (W == 1) {
        if (npar == 2) CPG_FIT_LAUNCH(2, 1); else if (npar == 3) CPG_FIT_LAUNCH(3, 1); else CPG_FIT_LAUNCH(5, 1);
    } else {
        if (npar == 2) CPG_FIT_LAUNCH(2, 2); else if (npar == 3) CPG_FIT_LAUNCH(3, 2); else CPG_FIT_LAUNCH(5, 2);
    }
#undef CPG_FIT_LAUNCH
    CPG_CUDA(cudaGetLastError());
    c->timing.kernel_launches = 1;
    c->timing.hot_kernel_launches = 1;
    T.mark(3);
    CPG_CUDA(cudaMemcpyAsync(best, c->fit_best.p, sizeof(double) * n * npar, cudaMemcpyDeviceToHost, c->stream));
    if (fun) CPG_CUDA(cudaMemcpyAsync(fun, c->fit_fun.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (nfev) CPG_CUDA(cudaMemcpyAsync(nfev, c->fit_nfev.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    if (nit) CPG_CUDA(cudaMemcpyAsync(nit, c->fit_nit.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    c->timing.d2h_bytes = (int64_t)(sizeof(double) * (n * npar + n) + 2 * sizeof(int32_t) * n);
    T.mark(4);
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    c->timing.h2d_ms = Timer::ms(c->ev[0], c->ev[1]);
    c->timing.kernel_ms = Timer::ms(c->ev[2], c->ev[3]);
    c->timing.d2h_ms = Timer::ms(c->ev[3], c->ev[4]);
    return CPG_OK;
}

CPG_API int cpg_get_particles(cpg_ctx *c, double *xyz, double *vxyz, double *masses)
{
    CPG_TRY(use_device(c));
    const size_t n = (size_t)c->n_part;
    if (xyz && n) CPG_CUDA(cudaMemcpyAsync(xyz, c->xyz.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    if (masses && n) CPG_CUDA(cudaMemcpyAsync(masses, c->masses.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (vxyz && n) {
        if (!c->have_vel) {
            set_error("cpg_get_particles: no velocities on the device");
            return CPG_ERR_STATE;
        }
        CPG_CUDA(cudaMemcpyAsync(vxyz, c->vxyz.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    }
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    return CPG_OK;
}

CPG_API int cpg_shape(cpg_ctx *c, const double *centers, const double *r200, const double *r_over_r200, int32_t R,
                      int32_t local, double SAFE, float L_BOX, double IT_TOL, int32_t IT_WALL, int32_t IT_MIN,
                      int32_t reduced, int32_t shell_based, int32_t use_vel, double *out12, int32_t *n_iter,
                      int32_t *n_pts, double *obj_mass, double *vcenters)
{
    CPG_TRY(compute_ready(c));
    const int n = c->n_obj;
    if (!local) R = 1;
    if (R < 1 || R > 32767 || (n > 0 && (!centers || !r200 || !out12 || !n_iter || !n_pts)) || (local && !r_over_r200)) {
        set_error("cpg_shape: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    if (use_vel && !c->have_vel) {
        set_error("cpg_shape: use_vel set but no velocities were uploaded");
        return CPG_ERR_STATE;
    }
    if (!local) shell_based = 0;   // calcObjMorphGlobal always runs the ellipsoid variant (:1099)
    Timer T(c);
    if (n == 0) return CPG_OK;
    const size_t nR = (size_t)n * R;
    // Any error exit below must not leave kernels behind that still store into the caller's page-locked result
    // buffers (direct_results): drain both streams before returning a failure.
    struct DrainOnError {
        cpg_ctx *c;
        bool ok = false;
        ~DrainOnError()
        {
            if (ok) return;
            cudaStreamSynchronize(c->stream);
            cudaStreamSynchronize(c->stream2);
            (void)cudaGetLastError();
        }
    } drain{c};
    T.mark(0);
#ifdef CPG_TASK_TRACE
    CPG_TRY(c->morph_stats.ensure(2 + 4 * ((size_t)1 << 20)));
    CPG_CUDA(cudaMemsetAsync(c->morph_stats.p, 0, sizeof(unsigned long long) * (2 + 4 * ((size_t)1 << 20)), c->stream));
#else
    CPG_TRY(c->morph_stats.ensure(2));
#endif
    CPG_TRY(c->centers.ensure((size_t)n * 3)); CPG_TRY(c->r200.ensure(n)); CPG_TRY(c->rr.ensure(R));
    c->stage_used = 0;
    CPG_TRY(small_upload(c, c->centers.p, centers, sizeof(double) * 3 * n));
    CPG_TRY(small_upload(c, c->r200.p, r200, sizeof(double) * n));
    if (local) CPG_TRY(small_upload(c, c->rr.p, r_over_r200, sizeof(double) * R));
    c->timing.h2d_bytes = (int64_t)sizeof(double) * (4 * (int64_t)n + (local ? R : 0));

    // ---- task list: objects largest first; S shells of one object per task
    const int warp_slots = c->sm_count * 4 * CPG_WARP_MINB;   // resident warps of the warp kernel (CPG_WARP_MINB CTAs x 4 warps per SM)
    const int64_t n_units = (int64_t)n * R;
    int S = c->opt_shells_per_pass;
    if (S == 0) {
        // S shells share one read of the particles.  Worth it when there are still plenty of tasks to fill the
        // GPU afterwards, or when the objects are so large that streaming them dominates (cluster teams then
        // supply the parallelism).
        const int64_t mean_size = c->n_idx / std::max(1, n);
        S = 1;
        if (R >= 2 && (n_units / 2 >= 4 * (int64_t)warp_slots || mean_size >= 200000)) S = 2;
        if (R >= 4 && (n_units / 4 >= 4 * (int64_t)warp_slots || mean_size >= 200000)) S = 4;
    }
    if (S > R) S = pow2_floor(R);
    // (the warp kernel keeps prefix lengths as 16-bit pair counts: its objects must stay below 131072 particles)
    int64_t large_min = c->opt_large_halo_min > 0 ? std::min<int64_t>(c->opt_large_halo_min, 131072) : (int64_t)131072;
    // few tasks: a warp per task would leave the GPU to a handful of long-running warps (non-converging shells
    // iterate up to IT_WALL times); give every task of a non-tiny object a CTA / cluster instead
    if (c->opt_large_halo_min == 0 && (n_units + S - 1) / S < 8 * (int64_t)warp_slots) large_min = 4096;
    // the warp kernel passes over two groups of 4 shells per task when there are tasks to spare: the per-iteration
    // reduction + eigen-solve + bookkeeping (40 % of its time at 4 shells per task) is then shared by 8 shells
    // (measured on configs[1]: 41 ms against 29 ms with one group -- fewer, longer tasks, fewer of them resident in
    // shared memory -- so two groups are only taken when asked for)
    int G = 1;
    if (S == 4 && c->opt_groups == 2 && R >= 8) G = 2;
    // objects of millions of particles: one launch over the whole grid per Katz iteration (kernel 3)
    const int64_t grid_min = c->opt_grid_halo_min > 0 ? c->opt_grid_halo_min : ((int64_t)1 << 22);
    cpg_ctx::TaskKey *hit = nullptr, *victim = nullptr;
    for (auto &tk : c->task_cache)
        if (tk.valid && tk.cat == c->catalogue_version && tk.S == S * 16 + G + 256 * c->opt_batched_step && tk.R == R && tk.large_min == large_min &&
            tk.grid_min == grid_min && (c->opt_cluster_size == 0 || c->opt_cluster_size == tk.cluster_size))
            hit = &tk;
    for (auto &tk : c->task_cache)   // victim: a free slot, else the least recently used one
        if (!victim || (victim->valid && (!tk.valid || tk.used < victim->used))) victim = &tk;
    auto &TK = hit ? *hit : *victim;
    TK.used = ++c->task_clock;
    if (!hit) {
        TK.valid = false;
        std::vector<Task> warp_tasks, cl_tasks, grid_tasks;
        int64_t largest = 0;
        for (int oi = 0; oi < n; oi++) {
            const int h = c->order[oi];
            const int64_t N = c->h_size[h];
            if (N >= grid_min) {
                for (int k0 = 0; k0 < R; k0 += S) grid_tasks.push_back(Task{h, (int16_t)k0, (int16_t)std::min(S, R - k0)});
                continue;
            }
            largest = std::max(largest, N);
            const bool big = N >= large_min;
            std::vector<Task> &dst = big ? cl_tasks : warp_tasks;
            const int step = big ? S : S * G;
            for (int k0 = 0; k0 < R; k0 += step)
                dst.push_back(Task{h, (int16_t)k0, (int16_t)std::min(step, R - k0)});
        }
        if (c->opt_batched_step && G == 1 && local && R > 1) {
            // batched step: the warps of a CTA meet once per Katz iteration, so neighbours in the queue should stream
            // prefixes of similar length.  Estimated pass length of (object, shells k0..k0+n-1): N (r_last / r_max)^1.5
            // (between the r^2 of an inner cusp and the r^1 of an isothermal envelope); longest first.
            const double rmax = r_over_r200[R - 1] > 0.0 ? r_over_r200[R - 1] : 1.0;
            auto cost = [&](const Task &t) {
                const double x = std::min(1.0, std::max(1e-6, r_over_r200[t.shell0 + t.nshell - 1] / rmax));
                return (double)c->h_size[t.halo] * x * sqrt(x);
            };
            std::stable_sort(warp_tasks.begin(), warp_tasks.end(), [&](const Task &a, const Task &b) { return cost(a) > cost(b); });
        }
        int cluster_size = c->opt_cluster_size;
        if (cluster_size == 0) {
            cluster_size = 1;
            if (!cl_tasks.empty()) {
                // enough CTAs to cover the GPU about twice, bounded by the largest object (>= 32k particles per CTA)
                const int64_t want = (2 * (int64_t)c->sm_count + (int64_t)cl_tasks.size() - 1) / (int64_t)cl_tasks.size();
                const int64_t by_size = std::max<int64_t>(1, largest / 32768);
                cluster_size = pow2_floor(std::max<int64_t>(1, std::min<int64_t>(std::min(want, by_size), 16)));
            }
        }
        TK.nw = (int)warp_tasks.size(); TK.ncl = (int)cl_tasks.size(); TK.cluster_size = cluster_size;
        TK.ngrid = (int)grid_tasks.size(); TK.grid_min = grid_min;
        warp_tasks.insert(warp_tasks.end(), cl_tasks.begin(), cl_tasks.end());
        warp_tasks.insert(warp_tasks.end(), grid_tasks.begin(), grid_tasks.end());   // layout: warp | cluster | grid
        CPG_TRY(TK.tasks.ensure(warp_tasks.size()));
        CPG_TRY(small_upload(c, TK.tasks.p, warp_tasks.data(), sizeof(Task) * warp_tasks.size()));
        c->timing.h2d_bytes += (int64_t)(sizeof(Task) * warp_tasks.size());
        TK.valid = true; TK.cat = c->catalogue_version; TK.S = S * 16 + G + 256 * c->opt_batched_step; TK.R = R; TK.large_min = large_min;
    }
    const int nw = TK.nw, ncl = TK.ncl, ngrid = TK.ngrid, cluster_size = TK.cluster_size;
    // Page-locked result buffers (cpg_host_alloc): the kernels store every finished halo-shell straight into them
    // while they run, so the results cross PCIe under the kernels instead of in a 74 MB copy after them (configs[1]:
    // 1.3 ms per step on one GPU, 4.3 ms when 8 GPUs share the host's DMA bandwidth).  The buffers are initialised
    // on the host (zeros / -1: what the reference returns for failed shells) by a few threads while the GPU gathers.
    double *d_out12 = c->opt_direct_results ? (double *)pinned_device_alias(out12) : nullptr;
    int32_t *d_nit = c->opt_direct_results ? (int32_t *)pinned_device_alias(n_iter) : nullptr;
    int32_t *d_npt = c->opt_direct_results ? (int32_t *)pinned_device_alias(n_pts) : nullptr;
    const bool direct = d_out12 && d_nit && d_npt;
    std::vector<std::thread> fillers;
    if (direct) {
        const int nt = 4;
        for (int t = 0; t < nt; t++)
            fillers.emplace_back([=]() {
                const size_t a0 = nR * 12 * t / nt, a1 = nR * 12 * (t + 1) / nt, b0 = nR * t / nt, b1 = nR * (t + 1) / nt;
                memset(out12 + a0, 0, sizeof(double) * (a1 - a0));
                memset(n_iter + b0, 0xff, sizeof(int32_t) * (b1 - b0));
                memset(n_pts + b0, 0xff, sizeof(int32_t) * (b1 - b0));
            });
    } else {
        CPG_TRY(c->out12.ensure(nR * 12)); CPG_TRY(c->n_iter.ensure(nR)); CPG_TRY(c->n_pts.ensure(nR));
    }
    struct Joiner {   // every exit path waits for the fill threads
        std::vector<std::thread> &th;
        ~Joiner() { for (auto &x : th) if (x.joinable()) x.join(); }
    } joiner{fillers};
    // (a fill kernel, not cudaMemsetAsync: the copy-engine memset of the 67 MB result array of configs[1] took ~1 ms)
    init_results_kernel<<<grid_for(std::max<int64_t>((int64_t)(direct ? 0 : nR) * 6, 4), 256, c->sm_count * 16), 256, 0, c->stream>>>(
        reinterpret_cast<double2 *>(c->out12.p), c->n_iter.p, c->n_pts.p, c->task_counter.p, c->morph_stats.p,
        (int64_t)(direct ? 0 : nR));
    c->timing.kernel_launches++;
    CPG_CUDA(cudaGetLastError());
    T.mark(1);

    // ---- gather (positions, masses, statistics, velocity centres)
    bool radial = local && c->opt_radial_order && R <= SORT_MAX_B - 1;
    for (int k = 0; radial && k < R; k++)
        if (!(r_over_r200[k] > 0.0) || (k > 0 && !(r_over_r200[k] > r_over_r200[k - 1]))) radial = false;
    CPG_TRY(run_gather(c, centers, r200, r_over_r200, L_BOX, radial ? R : 0));
    CPG_TRY(run_stats(c));
    if (use_vel) CPG_TRY(run_vcenters(c, L_BOX));
    const bool inner = radial && !shell_based && !reduced && c->opt_prefix_table;
    if (inner) CPG_TRY(run_ptab(c, use_vel != 0));
    T.mark(2);

    // ---- the morphology loop
    MorphParams P;
    P.soa = soa_view(c);
    P.r200 = c->r200.p; P.rr = c->rr.p; P.skip = c->skip.p;
    P.n_eff = radial ? c->n_eff.p : nullptr;   // run_gather guarantees n_eff matches these radii when radial
    P.ptab = inner ? c->ptab.p : nullptr;
    P.tasks = TK.tasks.p; P.n_tasks = 0; P.task_counter = c->task_counter.p;
    P.R = R; P.local = local ? 1 : 0; P.SAFE = SAFE;
    P.tol = IT_TOL; P.wall = IT_WALL; P.itmin = IT_MIN;
    P.out12 = direct ? d_out12 : c->out12.p; P.n_iter = direct ? d_nit : c->n_iter.p; P.n_pts = direct ? d_npt : c->n_pts.p;
    P.stats = c->morph_stats.p;
    for (auto &x : fillers) x.join();   // the host-side initialisation ran under the gather; it must end before a kernel stores
    if (ngrid > 0) {
        int rcg;
        if (S == 4) rcg = dispatch_grid<4>(c, P, shell_based != 0, reduced != 0, use_vel != 0, nw + ncl, ngrid);
        else if (S == 2) rcg = dispatch_grid<2>(c, P, shell_based != 0, reduced != 0, use_vel != 0, nw + ncl, ngrid);
        else rcg = dispatch_grid<1>(c, P, shell_based != 0, reduced != 0, use_vel != 0, nw + ncl, ngrid);
        CPG_TRY(rcg);
    }
    int rc;
    if (S == 4 && G == 2) rc = dispatch_morph<4, 2>(c, P, shell_based != 0, reduced != 0, use_vel != 0, nw, ncl, cluster_size);
    else if (S == 4) rc = dispatch_morph<4, 1>(c, P, shell_based != 0, reduced != 0, use_vel != 0, nw, ncl, cluster_size);
    else if (S == 2) rc = dispatch_morph<2, 1>(c, P, shell_based != 0, reduced != 0, use_vel != 0, nw, ncl, cluster_size);
    else rc = dispatch_morph<1, 1>(c, P, shell_based != 0, reduced != 0, use_vel != 0, nw, ncl, cluster_size);
    CPG_TRY(rc);
    T.mark(3);

    // ---- results
    if (!direct) {
        CPG_CUDA(cudaMemcpyAsync(out12, c->out12.p, sizeof(double) * 12 * nR, cudaMemcpyDeviceToHost, c->stream));
        CPG_CUDA(cudaMemcpyAsync(n_iter, c->n_iter.p, sizeof(int32_t) * nR, cudaMemcpyDeviceToHost, c->stream));
        CPG_CUDA(cudaMemcpyAsync(n_pts, c->n_pts.p, sizeof(int32_t) * nR, cudaMemcpyDeviceToHost, c->stream));
    }
    // (direct: the same bytes crossed PCIe as stores of the kernels)
    c->timing.d2h_bytes = (int64_t)(nR * (12 * sizeof(double) + 2 * sizeof(int32_t)));
    if (obj_mass) {
        CPG_CUDA(cudaMemcpyAsync(obj_mass, c->obj_mass.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        c->timing.d2h_bytes += (int64_t)sizeof(double) * n;
    }
    if (use_vel && vcenters) {
        CPG_CUDA(cudaMemcpyAsync(vcenters, c->vcenters.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
        c->timing.d2h_bytes += (int64_t)sizeof(double) * 3 * n;
    }
    unsigned long long h_stats[2] = {0, 0};
    CPG_CUDA(cudaMemcpyAsync(h_stats, c->morph_stats.p, sizeof h_stats, cudaMemcpyDeviceToHost, c->stream));
    T.mark(4);
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    drain.ok = true;
    c->timing.h2d_ms = Timer::ms(c->ev[0], c->ev[1]);
    c->timing.gather_ms = Timer::ms(c->ev[1], c->ev[2]);
    c->timing.kernel_ms = Timer::ms(c->ev[2], c->ev[3]);
    c->timing.d2h_ms = Timer::ms(c->ev[3], c->ev[4]);
    c->timing.tile_visits = (int64_t)h_stats[0];
    c->timing.tile_shell_visits = (int64_t)h_stats[1];
    return CPG_OK;
}

static int dens_gather(cpg_ctx *c, const double *centers, float L_BOX)
{
    CPG_TRY(run_gather(c, centers, nullptr, nullptr, L_BOX, 0));
    // per-chunk largest mass: anchors the fixed-point grid of the bin masses (cpg_dens.cuh); once per gather
    if (!c->have.mmax) {
        const int nc = (int)c->h_chunks.size();
        CPG_TRY(c->chunk_mmax.ensure((size_t)std::max(nc, 1)));
        if (nc > 0) {
            chunk_mmax_kernel<<<nc, 256, 0, c->stream>>>(soa_view(c), c->chunks.p, nc, c->chunk_mmax.p);
            c->timing.kernel_launches++;
            CPG_CUDA(cudaGetLastError());
        }
        c->have.mmax = c->have.valid;
    }
    return CPG_OK;
}

static int dens_common_begin(cpg_ctx *c, Timer &T, const double *centers, const double *r200, float L_BOX)
{
    const int n = c->n_obj;
    T.mark(0);
    CPG_TRY(c->centers.ensure((size_t)n * 3)); CPG_TRY(c->r200.ensure(n));
    c->stage_used = 0;
    CPG_TRY(small_upload(c, c->centers.p, centers, sizeof(double) * 3 * n));
    if (r200) CPG_TRY(small_upload(c, c->r200.p, r200, sizeof(double) * n));
    c->timing.h2d_bytes = (int64_t)sizeof(double) * (r200 ? 4 : 3) * n;
    return CPG_OK;
}

static int dens_common_end(cpg_ctx *c, Timer &T, double *dens, int32_t *counts, size_t nR)
{
    T.mark(3);
    CPG_CUDA(cudaMemcpyAsync(dens, c->dens.p, sizeof(double) * nR, cudaMemcpyDeviceToHost, c->stream));
    c->timing.d2h_bytes = (int64_t)(sizeof(double) * nR);
    if (counts) {
        CPG_CUDA(cudaMemcpyAsync(counts, c->counts.p, sizeof(int32_t) * nR, cudaMemcpyDeviceToHost, c->stream));
        c->timing.d2h_bytes += (int64_t)(sizeof(int32_t) * nR);
    }
    T.mark(4);
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    c->timing.h2d_ms = Timer::ms(c->ev[0], c->ev[1]);
    c->timing.gather_ms = Timer::ms(c->ev[1], c->ev[2]);
    c->timing.kernel_ms = Timer::ms(c->ev[2], c->ev[3]);
    c->timing.d2h_ms = Timer::ms(c->ev[3], c->ev[4]);
    return CPG_OK;
}

CPG_API int cpg_dens_sph(cpg_ctx *c, const double *centers, const double *r200, const double *bin_edges, int32_t R,
                         float L_BOX, double *dens, int32_t *counts)
{
    CPG_TRY(compute_ready(c));
    const int n = c->n_obj;
    if (R < 1 || R > DENS_MAX_R || (n > 0 && (!centers || !r200 || !bin_edges || !dens))) {
        set_error("cpg_dens_sph: bad arguments (1 <= R <= %d)", DENS_MAX_R);
        return CPG_ERR_ARGUMENT;
    }
    Timer T(c);
    if (n == 0) return CPG_OK;
    const size_t nR = (size_t)n * R;
    const int nc = (int)c->h_chunks.size();
    CPG_TRY(dens_common_begin(c, T, centers, r200, L_BOX));
    CPG_TRY(c->d_edges.ensure(R + 1));
    CPG_CUDA(cudaMemcpyAsync(c->d_edges.p, bin_edges, sizeof(double) * (R + 1), cudaMemcpyHostToDevice, c->stream));
    CPG_TRY(c->bin_mass.ensure((size_t)nc * R)); CPG_TRY(c->bin_cnt.ensure((size_t)nc * R));
    CPG_TRY(c->dens.ensure(nR)); CPG_TRY(c->counts.ensure(nR));
    T.mark(1);
    CPG_TRY(dens_gather(c, centers, L_BOX));
    T.mark(2);
    DensParams D = {};
    D.soa = soa_view(c); D.chunks = c->chunks.p; D.n_chunks = nc; D.R = R; D.r200 = c->r200.p; D.edges = c->d_edges.p;
    D.part_mass = c->bin_mass.p; D.part_cnt = c->bin_cnt.p; D.chunk_mmax = c->chunk_mmax.p;
    D.monotone = 1;
    for (int k = 0; k < R; k++)
        if (!(bin_edges[k + 1] > bin_edges[k]) || !(bin_edges[k] >= 0.0)) D.monotone = 0;
    if (nc > 0) {
        dens_sph_kernel<<<nc, 256, 0, c->stream>>>(D);
        c->timing.kernel_launches++; c->timing.hot_kernel_launches++;
    }
    dens_bins_final_kernel<<<grid_for((int64_t)nR, 128, 1 << 30), 128, 0, c->stream>>>(D, c->chunk_first.p, n, 0,
                                                                                        c->dens.p, c->counts.p);
    c->timing.kernel_launches++;
    CPG_CUDA(cudaGetLastError());
    return dens_common_end(c, T, dens, counts, nR);
}

CPG_API int cpg_dens_ell(cpg_ctx *c, const double *centers, const double *a, const double *b, const double *cc,
                         const double *major, const double *inter, const double *minor, int32_t R, float L_BOX,
                         double *dens, int32_t *counts)
{
    CPG_TRY(compute_ready(c));
    const int n = c->n_obj;
    const bool resident = !a && !b && !cc && !major && !inter && !minor;   // profiles left by cpg_shape_interp
    if (R < 1 || R > DENS_MAX_R || (n > 0 && (!centers || !dens)) ||
        (n > 0 && !resident && (!a || !b || !cc || !major || !inter || !minor))) {
        set_error("cpg_dens_ell: bad arguments (1 <= R <= %d)", DENS_MAX_R);
        return CPG_ERR_ARGUMENT;
    }
    if (resident && n > 0 && (c->interp_R != R || c->interp_n != n)) {
        set_error("cpg_dens_ell: no interpolated profiles for %d objects x %d bins on the device (call cpg_shape_interp)", n, R);
        return CPG_ERR_STATE;
    }
    Timer T(c);
    if (n == 0) return CPG_OK;
    const size_t nR = (size_t)n * R, nR1 = (size_t)n * (R + 1);
    const int nc = (int)c->h_chunks.size();
    CPG_TRY(dens_common_begin(c, T, centers, nullptr, L_BOX));
    CPG_TRY(c->d_a.ensure(nR1)); CPG_TRY(c->d_b.ensure(nR1)); CPG_TRY(c->d_c.ensure(nR1));
    CPG_TRY(c->d_major.ensure(nR * 3)); CPG_TRY(c->d_inter.ensure(nR * 3)); CPG_TRY(c->d_minor.ensure(nR * 3));
    if (!resident) {
        c->interp_R = 0;
        CPG_CUDA(cudaMemcpyAsync(c->d_a.p, a, sizeof(double) * nR1, cudaMemcpyHostToDevice, c->stream));
        CPG_CUDA(cudaMemcpyAsync(c->d_b.p, b, sizeof(double) * nR1, cudaMemcpyHostToDevice, c->stream));
        CPG_CUDA(cudaMemcpyAsync(c->d_c.p, cc, sizeof(double) * nR1, cudaMemcpyHostToDevice, c->stream));
        CPG_CUDA(cudaMemcpyAsync(c->d_major.p, major, sizeof(double) * nR * 3, cudaMemcpyHostToDevice, c->stream));
        CPG_CUDA(cudaMemcpyAsync(c->d_inter.p, inter, sizeof(double) * nR * 3, cudaMemcpyHostToDevice, c->stream));
        CPG_CUDA(cudaMemcpyAsync(c->d_minor.p, minor, sizeof(double) * nR * 3, cudaMemcpyHostToDevice, c->stream));
        c->timing.h2d_bytes += (int64_t)(sizeof(double) * (3 * nR1 + 9 * nR));
    }
    CPG_TRY(c->bin_mass.ensure((size_t)nc * R)); CPG_TRY(c->bin_cnt.ensure((size_t)nc * R));
    CPG_TRY(c->dens.ensure(nR)); CPG_TRY(c->counts.ensure(nR));
    T.mark(1);
    CPG_TRY(dens_gather(c, centers, L_BOX));
    T.mark(2);
    DensParams D = {};
    D.soa = soa_view(c); D.chunks = c->chunks.p; D.n_chunks = nc; D.R = R;
    D.a = c->d_a.p; D.b = c->d_b.p; D.c = c->d_c.p;
    D.major = c->d_major.p; D.inter = c->d_inter.p; D.minor = c->d_minor.p;
    D.part_mass = c->bin_mass.p; D.part_cnt = c->bin_cnt.p; D.chunk_mmax = c->chunk_mmax.p;
    const size_t smem = (size_t)R * (EB_SLOTS * sizeof(double) + 2 * sizeof(unsigned long long) + sizeof(int));
    if (smem > 48 * 1024)
        CPG_CUDA(cudaFuncSetAttribute(dens_ell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (nc > 0) {
        dens_ell_kernel<<<nc, 256, smem, c->stream>>>(D);
        c->timing.kernel_launches++; c->timing.hot_kernel_launches++;
    }
    dens_bins_final_kernel<<<grid_for((int64_t)nR, 128, 1 << 30), 128, 0, c->stream>>>(D, c->chunk_first.p, n, 1,
                                                                                        c->dens.p, c->counts.p);
    c->timing.kernel_launches++;
    CPG_CUDA(cudaGetLastError());
    return dens_common_end(c, T, dens, counts, nR);
}

CPG_API int cpg_shape_interp(cpg_ctx *c, const double *r200, const double *r_over_r200, const double *bin_edges, int32_t R,
                             const double *a, const double *b, const double *cc, const double *major, const double *inter,
                             const double *minor, int32_t Rs, double *ai, double *bi, double *ci, double *mj, double *it,
                             double *mn)
{
    CPG_TRY(use_device(c));
    const int n = c->n_obj;
    if (R < 1 || R > DENS_MAX_R || Rs < 2 || Rs > INTERP_MAX_RS ||
        (n > 0 && (!r200 || !r_over_r200 || !bin_edges || !a || !b || !cc || !major || !inter || !minor))) {
        set_error("cpg_shape_interp: bad arguments (1 <= R <= %d, 2 <= Rs <= %d)", DENS_MAX_R, INTERP_MAX_RS);
        return CPG_ERR_ARGUMENT;
    }
    c->interp_R = 0;
    Timer T(c);
    if (n == 0) return CPG_OK;
    const size_t nRs = (size_t)n * Rs, nR = (size_t)n * R, nR1 = (size_t)n * (R + 1);
    T.mark(0);
    CPG_TRY(c->r200.ensure(n)); CPG_TRY(c->ip_rr.ensure(R)); CPG_TRY(c->ip_edges.ensure(R + 1));
    CPG_TRY(c->ip_a.ensure(nRs)); CPG_TRY(c->ip_b.ensure(nRs)); CPG_TRY(c->ip_c.ensure(nRs));
    CPG_TRY(c->ip_major.ensure(nRs * 3)); CPG_TRY(c->ip_inter.ensure(nRs * 3)); CPG_TRY(c->ip_minor.ensure(nRs * 3));
    CPG_TRY(c->d_a.ensure(nR1)); CPG_TRY(c->d_b.ensure(nR1)); CPG_TRY(c->d_c.ensure(nR1));
    CPG_TRY(c->d_major.ensure(nR * 3)); CPG_TRY(c->d_inter.ensure(nR * 3)); CPG_TRY(c->d_minor.ensure(nR * 3));
    const struct { double *dst; const double *src; size_t cnt; } up[] = {
        {c->r200.p, r200, (size_t)n}, {c->ip_rr.p, r_over_r200, (size_t)R}, {c->ip_edges.p, bin_edges, (size_t)R + 1},
        {c->ip_a.p, a, nRs}, {c->ip_b.p, b, nRs}, {c->ip_c.p, cc, nRs},
        {c->ip_major.p, major, nRs * 3}, {c->ip_inter.p, inter, nRs * 3}, {c->ip_minor.p, minor, nRs * 3}};
    for (const auto &u : up) {
        CPG_CUDA(cudaMemcpyAsync(u.dst, u.src, sizeof(double) * u.cnt, cudaMemcpyHostToDevice, c->stream));
        c->timing.h2d_bytes += (int64_t)(sizeof(double) * u.cnt);
    }
    T.mark(1);
    T.mark(2);
    InterpParams Q;
    Q.r200 = c->r200.p; Q.rr = c->ip_rr.p; Q.edges = c->ip_edges.p;
    Q.a = c->ip_a.p; Q.b = c->ip_b.p; Q.c = c->ip_c.p;
    Q.major = c->ip_major.p; Q.inter = c->ip_inter.p; Q.minor = c->ip_minor.p;
    Q.n_obj = n; Q.R = R; Q.Rs = Rs;
    Q.ai = c->d_a.p; Q.bi = c->d_b.p; Q.ci = c->d_c.p; Q.mj = c->d_major.p; Q.it = c->d_inter.p; Q.mn = c->d_minor.p;
    shape_interp_kernel<<<(n + INTERP_WARPS - 1) / INTERP_WARPS, INTERP_WARPS * 32, 0, c->stream>>>(Q);
    CPG_CUDA(cudaGetLastError());
    c->timing.kernel_launches = 1;
    c->timing.hot_kernel_launches = 1;
    T.mark(3);
    const struct { double *dst; const double *src; size_t cnt; } down[] = {
        {ai, c->d_a.p, nR1}, {bi, c->d_b.p, nR1}, {ci, c->d_c.p, nR1},
        {mj, c->d_major.p, nR * 3}, {it, c->d_inter.p, nR * 3}, {mn, c->d_minor.p, nR * 3}};
    for (const auto &d : down)
        if (d.dst) {
            CPG_CUDA(cudaMemcpyAsync(d.dst, d.src, sizeof(double) * d.cnt, cudaMemcpyDeviceToHost, c->stream));
            c->timing.d2h_bytes += (int64_t)(sizeof(double) * d.cnt);
        }
    T.mark(4);
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    c->timing.h2d_ms = Timer::ms(c->ev[0], c->ev[1]);
    c->timing.kernel_ms = Timer::ms(c->ev[2], c->ev[3]);
    c->timing.d2h_ms = Timer::ms(c->ev[3], c->ev[4]);
    c->interp_R = R;
    c->interp_n = n;
    return CPG_OK;
}

CPG_API int cpg_dens_kernel(cpg_ctx *c, const double *centers, const double *r200, const double *r_over_r200, int32_t R,
                            float L_BOX, double *dens)
{
    CPG_TRY(compute_ready(c));
    const int n = c->n_obj;
    if (R < 1 || R > DENS_MAX_R || (n > 0 && (!centers || !r200 || !r_over_r200 || !dens))) {
        set_error("cpg_dens_kernel: bad arguments (1 <= R <= %d)", DENS_MAX_R);
        return CPG_ERR_ARGUMENT;
    }
    Timer T(c);
    if (n == 0) return CPG_OK;
    const size_t nR = (size_t)n * R;
    const int nc = (int)c->h_kchunks.size();
    CPG_TRY(dens_common_begin(c, T, centers, r200, L_BOX));
    CPG_TRY(c->d_edges.ensure(R + 1));
    CPG_CUDA(cudaMemcpyAsync(c->d_edges.p, r_over_r200, sizeof(double) * R, cudaMemcpyHostToDevice, c->stream));
    CPG_TRY(c->bin_mass.ensure((size_t)nc * R));
    CPG_TRY(c->dens.ensure(nR));
    T.mark(1);
    CPG_TRY(dens_gather(c, centers, L_BOX));
    T.mark(2);
    DensParams D = {};
    D.soa = soa_view(c); D.chunks = c->kchunks.p; D.n_chunks = nc; D.R = R; D.r200 = c->r200.p; D.edges = c->d_edges.p;
    D.part_mass = c->bin_mass.p;
    D.ktilde_c0 = 1.0 / (2.0 * pow(2.0 * M_PI, 1.5));   // host libm, like the reference's generated C
    if (nc > 0) {
        dens_kernel_kernel<<<nc, 256, 0, c->stream>>>(D);
        c->timing.kernel_launches++; c->timing.hot_kernel_launches++;
    }
    dens_kernel_final_kernel<<<grid_for((int64_t)nR, 128, 1 << 30), 128, 0, c->stream>>>(D, c->kchunk_first.p, n,
                                                                                          c->dens.p);
    c->timing.kernel_launches++;
    CPG_CUDA(cudaGetLastError());
    return dens_common_end(c, T, dens, nullptr, nR);
}

CPG_API int cpg_obj_masses(cpg_ctx *c, double *obj_mass)
{
    CPG_TRY(compute_ready(c));
    const int n = c->n_obj;
    if (n > 0 && !obj_mass) return CPG_ERR_ARGUMENT;
    Timer T(c);
    if (n == 0) return CPG_OK;
    if (!c->have.valid) {
        std::vector<double> zeros((size_t)n * 3, 0.0);
        CPG_TRY(c->centers.ensure((size_t)n * 3));
        CPG_CUDA(cudaMemsetAsync(c->centers.p, 0, sizeof(double) * 3 * n, c->stream));
        CPG_TRY(run_gather(c, zeros.data(), nullptr, nullptr, 0.0f, 0));
    }
    CPG_TRY(run_stats(c));
    CPG_CUDA(cudaMemcpyAsync(obj_mass, c->obj_mass.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    return CPG_OK;
}

CPG_API int cpg_centers(cpg_ctx *c, int32_t mode, int32_t shape_path, double L_BOX, const double *ref_points,
                        double *centers_out)
{
    CPG_TRY(compute_ready(c));
    const int n = c->n_obj;
    if ((n > 0 && !centers_out) || (mode != 0 && mode != 1) || (L_BOX != 0.0 && !ref_points)) {
        set_error("cpg_centers: bad arguments (L_BOX != 0 needs the reference points of respectPBCNoRef)");
        return CPG_ERR_ARGUMENT;
    }
    Timer T(c);
    if (n == 0) return CPG_OK;
    {
        // The reference's class computes the centres anew in every driver (shape_profs_algos.pyx:586-591, dens_profs_algos.pyx:39-44)
        // from the same snapshot and catalogue: same inputs, same centres.  Returning the cached ones also keeps the resident
        // gathered catalogue valid -- getShapeCatGlobal right after getShapeCatLocal then neither recomputes the centres
        // (2.3 ms on configs[1]) nor gathers again (2.6 ms).
        auto &K = c->cen_cache;
        const size_t n3 = (size_t)n * 3;
        const bool ref_same = ref_points ? (K.ref.size() == n3 && !memcmp(K.ref.data(), ref_points, n3 * sizeof(double))) : K.ref.empty();
        if (c->opt_snapshot_cache && K.valid && K.snap == (int64_t)c->snapshot_version && K.cat == (int64_t)c->catalogue_version &&
            K.mode == mode && K.shape_path == shape_path && K.L_BOX == L_BOX && ref_same && K.centers.size() == n3) {
            memcpy(centers_out, K.centers.data(), n3 * sizeof(double));
            c->timing.h2d_ms = 0.0; c->timing.kernel_ms = 0.0; c->timing.d2h_ms = 0.0;
            c->timing.kernel_launches = 0; c->timing.hot_kernel_launches = 0;
            return CPG_OK;
        }
        K.valid = false;
    }
    T.mark(0);
    c->stage_used = 0;
    CPG_TRY(c->centers.ensure((size_t)n * 3));
    if (ref_points) {
        CPG_TRY(c->cref.ensure((size_t)n * 3));
        CPG_TRY(small_upload(c, c->cref.p, ref_points, sizeof(double) * 3 * n));
    }
    if (mode == 1) { CPG_TRY(c->cidxA.ensure(c->n_idx)); CPG_TRY(c->cidxB.ensure(c->n_idx)); }
    T.mark(1);
    T.mark(2);
    CenterParams P;
    P.G = gather_view(c, 0.0f);
    P.size = c->size.p; P.order = c->d_order.p; P.first_obj = 0; P.ref = ref_points ? c->cref.p : nullptr; P.L_BOX = L_BOX;
    P.mode = mode; P.shape_path = shape_path;
    P.idxA = c->cidxA.p; P.idxB = c->cidxB.p; P.centers = c->centers.p;
    // 'mode' on large objects: a cluster of 8 CTAs each (objects are listed largest first)
    int n_big = 0;
    if (mode == 1)
        while (n_big < n && c->h_size[c->order[n_big]] >= 32768) n_big++;
    if (n_big > 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_big * 8);
        cfg.blockDim = dim3(256);
        cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CPG_CUDA(cudaLaunchKernelEx(&cfg, centers_mode_cluster_kernel, P));
        c->timing.kernel_launches++; c->timing.hot_kernel_launches++;
    }
    if (n > n_big) {
        P.first_obj = n_big;
        centers_kernel<<<n - n_big, 256, 0, c->stream>>>(P);
    }
    c->timing.kernel_launches++; c->timing.hot_kernel_launches++;
    CPG_CUDA(cudaGetLastError());
    T.mark(3);
    CPG_CUDA(cudaMemcpyAsync(centers_out, c->centers.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    T.mark(4);
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    c->have.valid = false;   // c->centers was overwritten: the resident gather no longer matches it
    {
        auto &K = c->cen_cache;
        K.snap = (int64_t)c->snapshot_version; K.cat = (int64_t)c->catalogue_version; K.mode = mode; K.shape_path = shape_path;
        K.L_BOX = L_BOX;
        if (ref_points) K.ref.assign(ref_points, ref_points + (size_t)n * 3); else K.ref.clear();
        K.centers.assign(centers_out, centers_out + (size_t)n * 3);
        K.valid = true;
    }
    c->timing.h2d_ms = Timer::ms(c->ev[0], c->ev[1]);
    c->timing.kernel_ms = Timer::ms(c->ev[2], c->ev[3]);
    c->timing.d2h_ms = Timer::ms(c->ev[3], c->ev[4]);
    return CPG_OK;
}

// (n_rows, 12) results of cpg_shape -> the six arrays the reference's drivers return, zeros -> NaN
// (shape_profs_algos.pyx:616-624), on a few host threads.
CPG_API int cpg_unpack_shape(const double *out12, int64_t n_rows, double *d, double *q, double *s, double *major, double *inter,
                             double *minor)
{
    if (n_rows < 0 || (n_rows > 0 && (!out12 || !d || !q || !s || !major || !inter || !minor))) {
        set_error("cpg_unpack_shape: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const int T = host_threads((size_t)n_rows * 96 * 2);
    auto work = [&](int t) {
        const int64_t a = n_rows * t / T, b = n_rows * (t + 1) / T;
        for (int64_t i = a; i < b; i++) {
            const double *o = out12 + 12 * i;
            d[i] = o[0] == 0.0 ? nan : o[0];
            q[i] = o[1] == 0.0 ? nan : o[1];
            s[i] = o[2] == 0.0 ? nan : o[2];
            for (int r = 0; r < 3; r++) {
                major[3 * i + r] = o[3 + r] == 0.0 ? nan : o[3 + r];
                inter[3 * i + r] = o[6 + r] == 0.0 ? nan : o[6 + r];
                minor[3 * i + r] = o[9 + r] == 0.0 ? nan : o[9 + r];
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    return CPG_OK;
}

CPG_API int cpg_host_alloc(size_t bytes, void **out)
{
    if (!out) return CPG_ERR_ARGUMENT;
    *out = nullptr;
    if (cpg_device_count() <= 0) {
        set_error("cpg_host_alloc: no CUDA device");
        return CPG_ERR_NO_DEVICE;
    }
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        set_error("cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return CPG_ERR_NOMEM;
    }
    return CPG_OK;
}

CPG_API void cpg_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

CPG_API int cpg_get_prefix_lengths(cpg_ctx *c, int32_t *n_eff, int32_t R)
{
    CPG_TRY(use_device(c));
    if (!n_eff || !c->have.valid || c->have.sort_R != R || R <= 0) {
        set_error("cpg_get_prefix_lengths: the resident gather holds no radial ordering for R=%d", (int)R);
        return CPG_ERR_STATE;
    }
    CPG_CUDA(cudaMemcpyAsync(n_eff, c->n_eff.p, sizeof(int32_t) * (size_t)c->n_obj * R, cudaMemcpyDeviceToHost, c->stream));
    CPG_CUDA(cudaStreamSynchronize(c->stream));
    return CPG_OK;
}

CPG_API int cpg_get_stat(cpg_ctx *c, const char *name, int64_t *out)
{
    if (!c || !name || !out) return CPG_ERR_ARGUMENT;
    if (!strcmp(name, "uploads_skipped")) *out = (int64_t)c->uploads_skipped;
    else if (!strcmp(name, "catalogue_as_ranges")) *out = c->cat_ranges ? 1 : 0;
    else if (!strcmp(name, "f32_test")) *out = CPG_F32_TEST;
    else if (!strcmp(name, "n_part")) *out = c->n_part;
    else if (!strcmp(name, "n_obj")) *out = c->n_obj;
    else if (!strcmp(name, "snapshot_version")) *out = (int64_t)c->snapshot_version;
    else if (!strcmp(name, "catalogue_version")) *out = (int64_t)c->catalogue_version;
    else {
        set_error("cpg_get_stat: unknown name %s", name);
        return CPG_ERR_ARGUMENT;
    }
    return CPG_OK;
}

CPG_API int cpg_get_timing(cpg_ctx *c, cpg_timing *out)
{
    if (!c || !out) return CPG_ERR_ARGUMENT;
    *out = c->timing;
    return CPG_OK;
}

}  // extern "C"

#include "cpg_multi.inl"

#ifdef CPG_TASK_TRACE
// Diagnostics build only (make EXTRA=-DCPG_TASK_TRACE): per-task (start ns, end ns, smid << 32 | tiles, steps << 32 | N)
// of the last cpg_shape call; warp-kernel task t at word 2 + 4 t, cluster-kernel task t at 2 + 4 (2^19 + t).
extern "C" CPG_API int cpg_debug_trace(cpg_ctx *c, unsigned long long *out, int64_t n_words)
{
    using namespace cpg;
    if (!c || !out) return CPG_ERR_ARGUMENT;
    CPG_CUDA(cudaSetDevice(c->device));
    CPG_CUDA(cudaMemcpy(out, c->morph_stats.p, sizeof(unsigned long long) * (size_t)n_words, cudaMemcpyDeviceToHost));
    return CPG_OK;
}
#endif
