// cpg_multi.inl -- several GPUs of one box behind one handle (included at the end of cpg_api.cu).
//
// The reference parallelises ONE catalogue over the cores of one host (prange over objects,
// cosmic_profiles/shape_profs/shape_profs_algos.pyx:592, dens_profs/dens_profs_algos.pyx:102): objects are
// independent.  Here the objects of one catalogue are dealt to the devices by a particle-count-balanced LPT
// partition; every device receives only its own objects' particles (contiguous-range catalogues: the ranges are
// packed by host threads straight into page-locked bounce buffers; general catalogues: the whole snapshot and a
// sub-catalogue) and runs the single-device path; results land in disjoint rows of the caller's arrays.  No
// collective, no peer traffic.

#include <functional>
#include <queue>
#include <string>

namespace cpg {

// Greedy longest-processing-time assignment: objects by descending cost (ties: ascending id), each to the part with
// the smallest load so far (ties: lowest part).
static void lpt_assign(const double *cost, int32_t n, int32_t n_parts, int32_t *owner)
{
    std::vector<int32_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cost[a] > cost[b]; });
    typedef std::pair<double, int32_t> Load;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int32_t p = 0; p < n_parts; p++) heap.push(Load(0.0, p));
    for (int32_t h : order) {
        Load l = heap.top();
        heap.pop();
        owner[h] = l.second;
        l.first += cost[h];
        heap.push(l);
    }
}

struct HostPinned {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap && p) return CPG_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = std::max<size_t>(bytes, 4096);
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
            p = nullptr;
            set_error("cudaHostAlloc of %zu bytes failed: %s", want, cudaGetErrorString(cudaGetLastError()));
            return CPG_ERR_NOMEM;
        }
        cap = want;
        return CPG_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace cpg

struct cpg_multi {
    std::vector<cpg_ctx *> ctx;
    std::vector<std::vector<int32_t>> objs;   // per device: its objects, ascending catalogue ids
    std::vector<int64_t> n_part;              // per device: particles resident
    int32_t n_obj = 0;
    bool have_vel = false;
    struct Stage { HostPinned out, nit, npt, a, b, c, d, e, f; };
    std::vector<Stage> stage;
    // what was distributed last (same arrays again: nothing to do)
    struct {
        bool valid = false;
        const void *xyz = nullptr, *vxyz = nullptr, *masses = nullptr;
        int64_t n_part = 0, n_idx = 0;
        uint64_t tok[3] = {0, 0, 0}, cat_hash = 0;
        std::vector<int32_t> size;
    } key;
    double partition_ms = 0.0, upload_ms = 0.0, scatter_ms = 0.0, call_ms = 0.0;
    int opt_snapshot_cache = 1;
};

namespace cpg {

// Run fn(d) for every device on its own host thread; the first failure (status + message) is reported in the
// calling thread.
static int for_each_device(cpg_multi *m, const std::function<int(int)> &fn)
{
    const int D = (int)m->ctx.size();
    std::vector<int> rc((size_t)D, CPG_OK);
    std::vector<std::string> msg((size_t)D);
    auto work = [&](int d) {
        rc[d] = fn(d);
        if (rc[d] != CPG_OK) msg[d] = cpg_last_error();
    };
    std::vector<std::thread> th;
    for (int d = 1; d < D; d++) th.emplace_back(work, d);
    work(0);
    for (auto &x : th) x.join();
    for (int d = 0; d < D; d++)
        if (rc[d] != CPG_OK) {
            set_error("device %d (slot %d): %s", m->ctx[d]->device, d, msg[d].c_str());
            return rc[d];
        }
    return CPG_OK;
}

template <class T>
static void gather_rows(const T *src, const std::vector<int32_t> &objs, size_t row, T *dst)
{
    for (size_t i = 0; i < objs.size(); i++) memcpy(dst + i * row, src + (size_t)objs[i] * row, sizeof(T) * row);
}

template <class T>
static void scatter_rows(const T *src, const std::vector<int32_t> &objs, size_t row, T *dst)
{
    for (size_t i = 0; i < objs.size(); i++) memcpy(dst + (size_t)objs[i] * row, src + i * row, sizeof(T) * row);
}

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace cpg

extern "C" {

CPG_API int cpg_partition_lpt(const double *cost, int32_t n, int32_t n_parts, int32_t *owner_out)
{
    if (n < 0 || n_parts < 1 || (n > 0 && (!cost || !owner_out))) {
        set_error("cpg_partition_lpt: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    lpt_assign(cost, n, n_parts, owner_out);
    return CPG_OK;
}

CPG_API int cpg_multi_create(const int32_t *devices, int32_t n_dev, cpg_multi **out)
{
    if (!out || n_dev < 1 || !devices) {
        set_error("cpg_multi_create: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    *out = nullptr;
    // (a device may be listed more than once: every slot is a context of its own -- own streams, own buffers -- so
    // the slots are independent; they merely share that device's memory and SMs.  The single-GPU tests use this.)
    cpg_multi *m = new (std::nothrow) cpg_multi();
    if (!m) return CPG_ERR_NOMEM;
    for (int d = 0; d < n_dev; d++) {
        cpg_ctx *c = nullptr;
        const int rc = cpg_create(devices[d], &c);
        if (rc != CPG_OK) {
            cpg_multi_destroy(m);
            return rc;
        }
        m->ctx.push_back(c);
    }
    m->objs.resize((size_t)n_dev);
    m->n_part.assign((size_t)n_dev, 0);
    m->stage.resize((size_t)n_dev);
    *out = m;
    return CPG_OK;
}

CPG_API void cpg_multi_destroy(cpg_multi *m)
{
    if (!m) return;
    for (size_t d = 0; d < m->ctx.size(); d++) {
        cpg_destroy(m->ctx[d]);
        if (d < m->stage.size()) {
            auto &S = m->stage[d];
            HostPinned *all[] = {&S.out, &S.nit, &S.npt, &S.a, &S.b, &S.c, &S.d, &S.e, &S.f};
            for (auto *b : all) b->release();
        }
    }
    delete m;
}

CPG_API int cpg_multi_devices(cpg_multi *m)
{
    return m ? (int)m->ctx.size() : 0;
}

CPG_API cpg_ctx *cpg_multi_context(cpg_multi *m, int32_t slot)
{
    if (!m || slot < 0 || slot >= (int)m->ctx.size()) return nullptr;
    return m->ctx[slot];
}

CPG_API int cpg_multi_set_option(cpg_multi *m, const char *name, int64_t value)
{
    if (!m || !name) return CPG_ERR_ARGUMENT;
    if (!strcmp(name, "snapshot_cache")) {
        m->opt_snapshot_cache = (int)value;
        if (!value) m->key.valid = false;
        return CPG_OK;   // (the per-device contexts receive segments / sub-catalogues: their own caches do not apply)
    }
    for (auto *c : m->ctx) CPG_TRY(cpg_set_option(c, name, value));
    return CPG_OK;
}

// Distribute one snapshot + catalogue over the devices.
CPG_API int cpg_multi_set_snapshot(cpg_multi *m, const double *xyz, const double *vxyz, const double *masses,
                                   int64_t n_part, const int32_t *idx_cat, int64_t n_idx, const int32_t *obj_size,
                                   int32_t n_obj)
{
    if (!m || n_part < 0 || n_idx < 0 || n_obj < 0 || (n_part > 0 && (!xyz || !masses)) || (n_idx > 0 && !idx_cat) ||
        (n_obj > 0 && !obj_size)) {
        set_error("cpg_multi_set_snapshot: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    const int D = (int)m->ctx.size();
    const double t0 = now_ms();
    std::vector<int64_t> off((size_t)n_obj + 1, 0);
    for (int h = 0; h < n_obj; h++) {
        if (obj_size[h] < 0) {
            set_error("cpg_multi_set_snapshot: obj_size[%d] < 0", h);
            return CPG_ERR_ARGUMENT;
        }
        off[h + 1] = off[h] + obj_size[h];
    }
    if (off[n_obj] != n_idx) {
        set_error("cpg_multi_set_snapshot: sum(obj_size) = %lld but idx_cat has %lld entries", (long long)off[n_obj], (long long)n_idx);
        return CPG_ERR_ARGUMENT;
    }
    CatScan sc;
    scan_catalogue(idx_cat, off.data(), n_obj, sc);
    if (n_idx > 0 && (sc.min_idx < 0 || sc.max_idx >= n_part)) {
        set_error("cpg_multi_set_snapshot: idx_cat range [%lld, %lld] outside the %lld particles", (long long)sc.min_idx,
                  (long long)sc.max_idx, (long long)n_part);
        return CPG_ERR_ARGUMENT;
    }
    uint64_t tok[3] = {0, 0, 0};
    if (m->opt_snapshot_cache) {
        const bool full = m->opt_snapshot_cache == 2;
        tok[0] = host_token(xyz, sizeof(double) * 3 * (size_t)n_part, full);
        tok[1] = host_token(masses, sizeof(double) * (size_t)n_part, full);
        tok[2] = vxyz ? host_token(vxyz, sizeof(double) * 3 * (size_t)n_part, full) : 0;
        auto &K = m->key;
        if (K.valid && K.xyz == xyz && K.vxyz == vxyz && K.masses == masses && K.n_part == n_part && K.n_idx == n_idx &&
            K.tok[0] == tok[0] && K.tok[1] == tok[1] && K.tok[2] == tok[2] && K.cat_hash == sc.hash &&
            same_sizes(K.size, obj_size, n_obj)) {
            m->partition_ms = 0.0; m->upload_ms = 0.0;
            return CPG_OK;
        }
    }
    m->key.valid = false;
    // particle-count-balanced partition of the objects
    std::vector<double> cost((size_t)n_obj);
    for (int h = 0; h < n_obj; h++) cost[h] = (double)obj_size[h];
    std::vector<int32_t> owner((size_t)n_obj, 0);
    lpt_assign(cost.data(), n_obj, D, owner.data());
    for (auto &o : m->objs) o.clear();
    for (int h = 0; h < n_obj; h++) m->objs[(size_t)owner[h]].push_back(h);
    m->n_obj = n_obj;
    m->have_vel = vxyz != nullptr;
    const double t1 = now_ms();
    m->partition_ms = t1 - t0;

    const int rc = for_each_device(m, [&](int d) -> int {
        cpg_ctx *c = m->ctx[d];
        const auto &objs = m->objs[(size_t)d];
        const int32_t nd = (int32_t)objs.size();
        std::vector<int32_t> sizes((size_t)nd);
        for (int i = 0; i < nd; i++) sizes[i] = obj_size[objs[i]];
        CPG_TRY(use_device(c));
        if (sc.contiguous) {
            // only this shard's particles travel; the shard catalogue becomes the ranges [prefix, prefix + size)
            std::vector<int64_t> seg_start((size_t)nd), seg_len((size_t)nd), local((size_t)nd);
            int64_t run = 0;
            for (int i = 0; i < nd; i++) {
                seg_start[i] = sc.start[(size_t)objs[i]]; seg_len[i] = sizes[i];
                local[i] = run; run += sizes[i];
            }
            CPG_TRY(cpg_set_particles_segments(c, xyz, vxyz, masses, seg_start.data(), seg_len.data(), nd));
            m->n_part[(size_t)d] = run;
            static const int64_t none = 0;
            return install_catalogue(c, nullptr, nd ? local.data() : &none, run, sizes.data(), nd);
        }
        // general catalogue: whole snapshot + this shard's slices of the flat index array
        const int keep = c->opt_snapshot_cache;
        c->opt_snapshot_cache = 0;
        const int rc1 = cpg_set_particles(c, xyz, vxyz, masses, n_part);
        c->opt_snapshot_cache = keep;
        CPG_TRY(rc1);
        m->n_part[(size_t)d] = n_part;
        int64_t tot = 0;
        for (int i = 0; i < nd; i++) tot += sizes[i];
        std::vector<int32_t> sub((size_t)tot);
        int64_t run = 0;
        for (int i = 0; i < nd; i++) {
            if (sizes[i]) memcpy(sub.data() + run, idx_cat + off[(size_t)objs[i]], sizeof(int32_t) * (size_t)sizes[i]);
            run += sizes[i];
        }
        return install_catalogue(c, sub.data(), nullptr, tot, sizes.data(), nd);
    });
    m->upload_ms = now_ms() - t1;
    CPG_TRY(rc);
    if (m->opt_snapshot_cache) {
        auto &K = m->key;
        K.valid = true; K.xyz = xyz; K.vxyz = vxyz; K.masses = masses; K.n_part = n_part; K.n_idx = n_idx;
        K.tok[0] = tok[0]; K.tok[1] = tok[1]; K.tok[2] = tok[2]; K.cat_hash = sc.hash;
        K.size.assign(obj_size, obj_size + n_obj);
    }
    return CPG_OK;
}

CPG_API int cpg_multi_shard(cpg_multi *m, int32_t slot, int32_t *n_obj_out, int64_t *n_part_out, int32_t *objs_out)
{
    if (!m || slot < 0 || slot >= (int)m->ctx.size()) return CPG_ERR_ARGUMENT;
    const auto &o = m->objs[(size_t)slot];
    if (n_obj_out) *n_obj_out = (int32_t)o.size();
    if (n_part_out) *n_part_out = m->n_part[(size_t)slot];
    if (objs_out && !o.empty()) memcpy(objs_out, o.data(), sizeof(int32_t) * o.size());
    return CPG_OK;
}

// cpg_shape over all devices.  center_mode < 0: `centers` (n_obj,3) are given; 0 / 1: centres of mass / shrinking-sphere
// modes computed on every device for its own objects (cpg_centers, shape path; ref_points (n_obj,3) required when
// L_BOX != 0) and returned in centers_out (n_obj,3).  Everything else as cpg_shape; rows of every output belong to
// the object of the same row of the catalogue.
CPG_API int cpg_multi_shape(cpg_multi *m, const double *centers, int32_t center_mode, const double *ref_points,
                            const double *r200, const double *r_over_r200, int32_t R, int32_t local, double SAFE,
                            float L_BOX, double IT_TOL, int32_t IT_WALL, int32_t IT_MIN, int32_t reduced,
                            int32_t shell_based, int32_t use_vel, double *out12, int32_t *n_iter, int32_t *n_pts,
                            double *obj_mass, double *vcenters, double *centers_out)
{
    if (!m) return CPG_ERR_ARGUMENT;
    const int32_t n = m->n_obj;
    if (!local) R = 1;
    if (R < 1 || (n > 0 && (!r200 || !out12 || !n_iter || !n_pts)) || (center_mode < 0 && n > 0 && !centers) ||
        (center_mode >= 0 && L_BOX != 0.0f && !ref_points)) {
        set_error("cpg_multi_shape: bad arguments");
        return CPG_ERR_ARGUMENT;
    }
    const double t0 = now_ms();
    std::atomic<int64_t> scatter_us{0};
    const int rc = for_each_device(m, [&](int d) -> int {
        cpg_ctx *c = m->ctx[d];
        const auto &objs = m->objs[(size_t)d];
        const size_t nd = objs.size();
        if (nd == 0) return CPG_OK;
        auto &S = m->stage[(size_t)d];
        CPG_TRY(use_device(c));
        CPG_TRY(S.out.ensure(sizeof(double) * 12 * nd * R)); CPG_TRY(S.nit.ensure(sizeof(int32_t) * nd * R));
        CPG_TRY(S.npt.ensure(sizeof(int32_t) * nd * R));
        CPG_TRY(S.a.ensure(sizeof(double) * 3 * nd)); CPG_TRY(S.b.ensure(sizeof(double) * nd));
        CPG_TRY(S.c.ensure(sizeof(double) * nd)); CPG_TRY(S.d.ensure(sizeof(double) * 3 * nd));
        CPG_TRY(S.e.ensure(sizeof(double) * 3 * nd));
        double *cen = (double *)S.a.p, *r2 = (double *)S.b.p, *om = (double *)S.c.p, *vc = (double *)S.d.p, *rf = (double *)S.e.p;
        gather_rows(r200, objs, 1, r2);
        if (center_mode < 0) {
            gather_rows(centers, objs, 3, cen);
        } else {
            if (ref_points) gather_rows(ref_points, objs, 3, rf);
            CPG_TRY(cpg_centers(c, center_mode, 1, (double)L_BOX, ref_points ? rf : nullptr, cen));
        }
        CPG_TRY(cpg_shape(c, cen, r2, r_over_r200, R, local, SAFE, L_BOX, IT_TOL, IT_WALL, IT_MIN, reduced, shell_based,
                          use_vel, (double *)S.out.p, (int32_t *)S.nit.p, (int32_t *)S.npt.p, om, vc));
        const double s0 = now_ms();
        scatter_rows((const double *)S.out.p, objs, (size_t)12 * R, out12);
        scatter_rows((const int32_t *)S.nit.p, objs, (size_t)R, n_iter);
        scatter_rows((const int32_t *)S.npt.p, objs, (size_t)R, n_pts);
        if (obj_mass) scatter_rows(om, objs, 1, obj_mass);
        if (vcenters && use_vel) scatter_rows(vc, objs, 3, vcenters);
        if (centers_out) scatter_rows(cen, objs, 3, centers_out);
        scatter_us += (int64_t)((now_ms() - s0) * 1e3);
        return CPG_OK;
    });
    m->call_ms = now_ms() - t0;
    m->scatter_ms = (double)scatter_us.load() * 1e-3 / std::max<size_t>(1, m->ctx.size());
    return rc;
}

CPG_API int cpg_multi_get_info(cpg_multi *m, int32_t slot, int32_t *device, int32_t *n_obj, int64_t *n_part,
                               cpg_timing *timing, double *host_ms /* [4]: partition, upload, scatter (mean per device), last call */)
{
    if (!m || slot < 0 || slot >= (int)m->ctx.size()) return CPG_ERR_ARGUMENT;
    if (device) *device = m->ctx[(size_t)slot]->device;
    if (n_obj) *n_obj = (int32_t)m->objs[(size_t)slot].size();
    if (n_part) *n_part = m->n_part[(size_t)slot];
    if (timing) *timing = m->ctx[(size_t)slot]->timing;
    if (host_ms) { host_ms[0] = m->partition_ms; host_ms[1] = m->upload_ms; host_ms[2] = m->scatter_ms; host_ms[3] = m->call_ms; }
    return CPG_OK;
}

// ---- measured fp64 multiply-add peak of the device (the roofline the morphology kernels are held against) ----

}  // extern "C"

namespace cpg {
template <int ILP>
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace cpg

extern "C" {

// Dependent-free chains of fp64 FMAs on every SM (8 chains per thread, 16 and 32 warps per SM), best of a few
// repetitions, CUDA events: *tfma_per_s = 1e-12 * multiply-adds per second.
CPG_API int cpg_measure_fp64_peak(cpg_ctx *c, double *tfma_per_s)
{
    CPG_TRY(use_device(c));
    if (!tfma_per_s) return CPG_ERR_ARGUMENT;
    const int iters = 20000;
    double best = 0.0;
    for (int ctas = 2; ctas <= 4; ctas += 2) {
        const int grid = c->sm_count * ctas;
        CPG_TRY(c->peak_scratch.ensure((size_t)grid * 256));
        dfma_peak_kernel<8><<<grid, 256, 0, c->stream>>>(c->peak_scratch.p, 100, 1.0000001, 1e-9);
        for (int rep = 0; rep < 3; rep++) {
            CPG_CUDA(cudaEventRecord(c->ev[0], c->stream));
            dfma_peak_kernel<8><<<grid, 256, 0, c->stream>>>(c->peak_scratch.p, iters, 1.0000001, 1e-9);
            CPG_CUDA(cudaEventRecord(c->ev[1], c->stream));
            CPG_CUDA(cudaStreamSynchronize(c->stream));
            float ms = 0.f;
            CPG_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
            const double fmas = (double)grid * 256.0 * 8.0 * (double)iters;
            best = std::max(best, fmas / ((double)ms * 1e-3) / 1e12);
        }
    }
    CPG_CUDA(cudaGetLastError());
    *tfma_per_s = best;
    return CPG_OK;
}

}  // extern "C"
