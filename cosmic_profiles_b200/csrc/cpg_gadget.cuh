// cpg_gadget.cuh -- Gadget / FoF-layout ingest on the device.
//
// (1) Central-subhalo index catalogue.  Replaces calcObjCat + calcCSHIdxs
//     (cosmic_profiles/gadget/gen_catalogues.pyx:39-95, :12-35): for FoF group p with nb_shs[p] > 0 the
//     central subhalo is subhalo  sum_{q<p} nb_shs[q]  (:68-70) and its particles are the first
//     sh_len[...] particles of the group, i.e. indices  start_p = sum_{q<p} fof_sizes[q] ... start_p + len - 1
//     (:82-85, :29-35); the group passes when that length is >= MIN_NUMBER_PTCS (:71).  The reference
//     recomputes both prefix sums from scratch for every group (O(n^2)) and builds the flat catalogue with
//     one np.hstack per object (O(n * N)); here they are two device-wide exclusive scans and one fill.
// (2) Snapshot blocks as stored on disk (float32 or float64 Coordinates / Velocities / Masses, or a
//     MassTable constant) -> the fp64 arrays read_block hands to the drivers
//     (cosmic_profiles/gadget/readgadget.py:76-124, :127-194), unit division included, so that a float32
//     snapshot crosses PCIe at 12 instead of 24 bytes per particle and block.
#pragma once
#include "cpg_common.cuh"

namespace cpg {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // elements per CTA

struct Pair64 {
    int64_t a, b;
};
__device__ __forceinline__ Pair64 operator+(Pair64 x, Pair64 y) { return Pair64{x.a + y.a, x.b + y.b}; }

// element p of the first scan: (number of subhalos, FoF group length)
struct LoadGroup {
    const int32_t *nb_shs, *fof_sizes;
    __device__ __forceinline__ Pair64 operator()(int64_t p) const { return Pair64{nb_shs[p], fof_sizes[p]}; }
};

// element p of the second scan: (passes?, central-subhalo length if it passes)   gen_catalogues.pyx:66-73
struct LoadPass {
    const int32_t *nb_shs, *sh_len;
    const int64_t *sh_off;   // exclusive scan of nb_shs
    int64_t n_sh;
    int32_t min_ptcs;
    __device__ __forceinline__ Pair64 operator()(int64_t p) const
    {
        if (nb_shs[p] > 0) {
            const int64_t j = sh_off[p];
            if (j >= 0 && j < n_sh) {
                const int32_t len = sh_len[j];
                if (len >= min_ptcs) return Pair64{1, len};
            }
        }
        return Pair64{0, 0};
    }
};

__device__ __forceinline__ Pair64 shfl_up_pair(Pair64 v, int d)
{
    Pair64 r;
    r.a = __shfl_up_sync(0xffffffffu, v.a, d);
    r.b = __shfl_up_sync(0xffffffffu, v.b, d);
    return r;
}

// Inclusive scan of one value per thread over the CTA; returns the inclusive prefix and, in `total`, the CTA sum.
__device__ __forceinline__ Pair64 block_scan_inclusive(Pair64 v, Pair64 &total, Pair64 *s_warp /* [SCAN_THREADS/32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Pair64 u = shfl_up_pair(v, d);
        if (lane >= d) v = v + u;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    Pair64 off{0, 0}, tot{0, 0};
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        const Pair64 sw = s_warp[w];
        if (w < warp) off = off + sw;
        tot = tot + sw;
    }
    __syncthreads();
    total = tot;
    return v + off;
}

// Pass 1 of the device-wide scan: the sum of every SCAN_TILE-element tile.
template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sum_kernel(Load ld, int64_t n, Pair64 *tile_sum)
{
    __shared__ Pair64 s_warp[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    Pair64 v{0, 0};
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {   // coalesced: consecutive threads, consecutive elements
        const int64_t p = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
        if (p < n) v = v + ld(p);
    }
    Pair64 total;
    block_scan_inclusive(v, total, s_warp);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// Pass 2: exclusive scan of the tile sums by one CTA (n_tiles = n / 2048: 5e3 tiles for 1e7 groups);
// tile_sum[t] becomes the offset of tile t, tile_sum[n_tiles] the grand total.
__global__ void __launch_bounds__(SCAN_THREADS) scan_top_kernel(Pair64 *tile_sum, int n_tiles)
{
    __shared__ Pair64 s_warp[SCAN_THREADS / 32];
    Pair64 carry{0, 0};
    for (int t0 = 0; t0 < n_tiles; t0 += SCAN_THREADS) {
        const int t = t0 + threadIdx.x;
        const Pair64 v = t < n_tiles ? tile_sum[t] : Pair64{0, 0};
        Pair64 total;
        const Pair64 inc = block_scan_inclusive(v, total, s_warp);
        if (t < n_tiles) tile_sum[t] = Pair64{carry.a + inc.a - v.a, carry.b + inc.b - v.b};
        carry = carry + total;
    }
    if (threadIdx.x == 0) tile_sum[n_tiles] = carry;
}

// Pass 3: exclusive prefix of every element (blocked arrangement: thread t owns SCAN_ITEMS consecutive elements).
template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(Load ld, int64_t n, const Pair64 *tile_off, int64_t *out_a,
                                                                  int64_t *out_b)
{
    __shared__ Pair64 s_warp[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    Pair64 item[SCAN_ITEMS], sum{0, 0};
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        item[i] = (base + i < n) ? ld(base + i) : Pair64{0, 0};
        sum = sum + item[i];
    }
    Pair64 total;
    const Pair64 inc = block_scan_inclusive(sum, total, s_warp);
    Pair64 run = tile_off[blockIdx.x] + Pair64{inc.a - sum.a, inc.b - sum.b};
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) {
            out_a[base + i] = run.a;
            out_b[base + i] = run.b;
        }
        run = run + item[i];
    }
}

struct ObjCatParams {
    const int32_t *nb_shs, *sh_len;
    const double *group_r200;
    const int64_t *sh_off, *start;      // exclusive scans of nb_shs, fof_sizes
    const int64_t *pass_off, *idx_off;  // exclusive scans of pass, passing length
    int64_t n_fof, n_sh, n_part_limit;  // indices must stay below n_part_limit (<= 2^31: int32 catalogue)
    int32_t min_ptcs;
    // compacted per passing object
    int64_t *c_start, *c_off;   // first particle index; offset into the flat catalogue (n_pass + 1 entries)
    int32_t *c_size;
    double *c_r200;
    int32_t *error;             // 1: an index would not fit the int32 catalogue
};

// One thread per FoF group: passing groups write their compacted record (gen_catalogues.pyx:86-91).
__global__ void objcat_compact_kernel(ObjCatParams P, int64_t n_pass, int64_t n_idx)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) P.c_off[n_pass] = n_idx;
    if (p >= P.n_fof) return;
    LoadPass lp{P.nb_shs, P.sh_len, P.sh_off, P.n_sh, P.min_ptcs};
    const Pair64 e = lp(p);
    if (!e.a) return;
    const int64_t j = P.pass_off[p];
    P.c_start[j] = P.start[p];
    P.c_off[j] = P.idx_off[p];
    P.c_size[j] = (int32_t)e.b;
    P.c_r200[j] = P.group_r200[p];   // h_r200[p] = group_r200[p]  (:74)
    if (e.b > 0 && (P.start[p] < 0 || P.start[p] + e.b > P.n_part_limit)) atomicExch(P.error, 1);
}

// Flat catalogue: entry j of object o is start_o + j (calcCSHIdxs, gen_catalogues.pyx:29-35).  Every thread finds
// the object of its first entry by bisection in the (n_pass + 1) offsets, then walks forward; stores are coalesced.
__global__ void objcat_fill_kernel(const int64_t *__restrict__ c_off, const int64_t *__restrict__ c_start, int64_t n_pass,
                                   int64_t n_idx, int32_t *__restrict__ idx_cat)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_idx; j += stride) {
        int64_t lo = 0, hi = n_pass;   // c_off[lo] <= j < c_off[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (c_off[mid] <= j) lo = mid; else hi = mid;
        }
        idx_cat[j] = (int32_t)(c_start[lo] + (j - c_off[lo]));
    }
}

// ---- snapshot blocks -------------------------------------------------------------------------------
// readgadget.py:112-118 under NumPy >= 2 promotion rules (NEP 50), restated:
//   array = dataset[:] / unit_fac          float32 dataset: fp32 division by (float)unit_fac, float64: fp64 division
//   VEL:   array *= np.sqrt(head.time)     in place: fp64 product, rounded back to the array's dtype
// and read_block (:127-194) stores the result into a float64 array (exact widening).
struct BlockConvert {
    double unit_fac;     // divisor
    double post_mul;     // sqrt(time) for velocities
    int32_t has_post;    // apply post_mul (the reference multiplies only the VEL block, even by 1.0)
};

template <class T>
__device__ __forceinline__ double convert_one(T x, const BlockConvert &c);

template <>
__device__ __forceinline__ double convert_one<float>(float x, const BlockConvert &c)
{
    float y = __fdiv_rn(x, (float)c.unit_fac);
    if (c.has_post) y = (float)__dmul_rn((double)y, c.post_mul);
    return (double)y;
}

template <>
__device__ __forceinline__ double convert_one<double>(double x, const BlockConvert &c)
{
    double y = __ddiv_rn(x, c.unit_fac);
    if (c.has_post) y = __dmul_rn(y, c.post_mul);
    return y;
}

template <class T>
__global__ void block_convert_kernel(const T *__restrict__ src, double *__restrict__ dst, int64_t n, BlockConvert c)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = convert_one<T>(src[i], c);
}

__global__ void block_fill_kernel(double *__restrict__ dst, int64_t n, double v)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}

}  // namespace cpg
