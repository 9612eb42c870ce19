// cpg_gather.cuh -- catalogue gather: AoS snapshot + int32 index catalogue -> per-object, centre-relative,
// 16-byte aligned SoA streams, plus the per-object statistics the drivers need.
//
// Replaces the gather / PBC / mass loops of the reference drivers
// (cosmic_profiles/shape_profs/shape_profs_algos.pyx:593-599, :703-707, :820-827, :937-943;
//  cosmic_profiles/dens_profs/dens_profs_algos.pyx:45-47, :103-106, :211-214, :270-273) and
// CythonHelpers.respectPBCNoRef / calcCoM / calcLocalSpread (cython_helpers/helper_class.pyx:169-199,
// :85-105, :62-80).
#pragma once
#include "cpg_common.cuh"

namespace cpg {

constexpr int CHUNK = 16384;   // particles per statistics / binning chunk (one CTA each)

struct Chunk {
    int32_t halo;
    int32_t start;   // first particle of the chunk inside the object
};

struct GatherParams {
    const double *xyz;       // (n_part,3)
    const double *vxyz;      // (n_part,3) or nullptr
    const double *masses;    // (n_part)
    const int32_t *idx;      // (n_idx), or nullptr: every object is a contiguous range of the snapshot (FoF order,
                             // gadget/gen_catalogues.pyx:29-35) and entry j of object h is particle start[h] + (j - off[h])
    const int64_t *start;    // (n_obj) first particle of every object (ranges catalogue only)
    const int64_t *off;      // (n_obj+1) catalogue offsets
    const int64_t *poff;     // (n_obj+1) padded SoA offsets (even)
    const double *centers;   // (n_obj,3)
    const double *vcenters;  // (n_obj,3) (velocity gather only)
    int32_t n_obj;
    int64_t n_idx;
    float L_BOX;
    double *dx, *dy, *dz, *m, *vx, *vy, *vz;
    const int32_t *dest;     // (n_idx) slot of catalogue entry j inside its object (radial ordering) or nullptr
};

// Snapshot index of catalogue entry j, which belongs to object h whose entries start at `first`.
__device__ __forceinline__ int64_t cat_index(const GatherParams &G, int h, int64_t first, int64_t j)
{
    return G.idx ? (int64_t)G.idx[j] : G.start[h] + (j - first);
}

__device__ __forceinline__ int find_object(const int64_t *off, int n_obj, int64_t j)
{
    int lo = 0, hi = n_obj;   // invariant: off[lo] <= j < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= j) lo = mid; else hi = mid;
    }
    return lo;
}

// helper_class.pyx:169-199 for one particle: reflection about L/2 relative to the object's particle 0,
// distances truncated to fp32, the z test re-uses the (already updated) y distance (line 196).
__device__ __forceinline__ void pbc_reflect(double &x, double &y, double &z, double x0, double y0, float L_BOX)
{
    const float half = L_BOX / 2;
    const double L = (double)L_BOX;
    if (fabsf((float)__dsub_rn(x0, x)) > half) x = __dsub_rn(L, x);
    if (fabsf((float)__dsub_rn(y0, y)) > half) y = __dsub_rn(L, y);
    if (fabsf((float)__dsub_rn(y0, y)) > half) z = __dsub_rn(L, z);
}

// Positions and masses.  One thread per catalogue entry; consecutive threads write consecutive SoA slots.
__global__ void __launch_bounds__(256) gather_pos_kernel(const GatherParams G)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < G.n_idx; j += stride) {
        const int h = find_object(G.off, G.n_obj, j);
        const int64_t first = G.off[h], last = G.off[h + 1];
        const int64_t g = cat_index(G, h, first, j);
        double x = G.xyz[3 * g], y = G.xyz[3 * g + 1], z = G.xyz[3 * g + 2];
        if (G.L_BOX != 0.0f) {
            const int64_t g0 = cat_index(G, h, first, first);
            pbc_reflect(x, y, z, G.xyz[3 * g0], G.xyz[3 * g0 + 1], G.L_BOX);
        }
        const int64_t o = G.poff[h] + (G.dest ? (int64_t)G.dest[j] : (j - first));
        G.dx[o] = __dsub_rn(x, G.centers[3 * h]);
        G.dy[o] = __dsub_rn(y, G.centers[3 * h + 1]);
        G.dz[o] = __dsub_rn(z, G.centers[3 * h + 2]);
        G.m[o] = G.masses[g];
        if (j == last - 1 && ((last - first) & 1)) {   // pad element of an odd-sized object
            const int64_t pad = G.poff[h] + (last - first);
            G.dx[pad] = 1e150; G.dy[pad] = 1e150; G.dz[pad] = 1e150; G.m[pad] = 0.0;
        }
    }
}

// Velocities relative to the object's velocity centre (shape_profs_algos.pyx:342: vxyz[run,i]-vcenter[i]).
__global__ void __launch_bounds__(256) gather_vel_kernel(const GatherParams G)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < G.n_idx; j += stride) {
        const int h = find_object(G.off, G.n_obj, j);
        const int64_t first = G.off[h], last = G.off[h + 1];
        const int64_t g = cat_index(G, h, first, j);
        const int64_t o = G.poff[h] + (G.dest ? (int64_t)G.dest[j] : (j - first));
        G.vx[o] = __dsub_rn(G.vxyz[3 * g], G.vcenters[3 * h]);
        G.vy[o] = __dsub_rn(G.vxyz[3 * g + 1], G.vcenters[3 * h + 1]);
        G.vz[o] = __dsub_rn(G.vxyz[3 * g + 2], G.vcenters[3 * h + 2]);
        if (j == last - 1 && ((last - first) & 1)) {
            const int64_t pad = G.poff[h] + (last - first);
            G.vx[pad] = 0.0; G.vy[pad] = 0.0; G.vz[pad] = 0.0;
        }
    }
}

// Deterministic block sum of up to 4 values (fixed tree), result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch /* [8][NV] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0)
        for (int k = 0; k < NV; k++) scratch[warp * NV + k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < NV; k++) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += scratch[w * NV + k];
            v[k] = s;
        }
}

// Per-chunk partial statistics: mass, "some particle differs from particle 0" flag (the only thing the
// reference consumes of calcLocalSpread is ==0, i.e. an empty or fully coincident cloud).
__global__ void __launch_bounds__(256) stats_chunk_kernel(const SoA soa, const Chunk *chunks, int n_chunks,
                                                          double *part_mass, int32_t *part_spread, double *part_mminmax)
{
    __shared__ double scratch[8 * 1];
    __shared__ double s_mm[8][2];
    __shared__ int s_flag;
    const int c = blockIdx.x;
    if (c >= n_chunks) return;
    const Chunk ck = chunks[c];
    const int N = soa.size[ck.halo];
    const int64_t base = soa.poff[ck.halo];
    const int end = min(N, ck.start + CHUNK);
    if (threadIdx.x == 0) s_flag = 0;
    __syncthreads();
    const double x0 = soa.dx[base], y0 = soa.dy[base], z0 = soa.dz[base];
    double v[1] = {0.0};
    double mlo = 1e300, mhi = -1e300;
    int differs = 0;
    for (int i = ck.start + threadIdx.x; i < end; i += blockDim.x) {
        const double m = soa.m[base + i];
        v[0] += m;
        mlo = fmin(mlo, m); mhi = fmax(mhi, m);
        differs |= (soa.dx[base + i] != x0) | (soa.dy[base + i] != y0) | (soa.dz[base + i] != z0);
    }
    if (differs) s_flag = 1;
    for (int o = 16; o > 0; o >>= 1) {
        mlo = fmin(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
        mhi = fmax(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mm[threadIdx.x >> 5][0] = mlo; s_mm[threadIdx.x >> 5][1] = mhi; }
    block_sum<1>(v, scratch);
    if (threadIdx.x == 0) {
        part_mass[c] = v[0];
        part_spread[c] = s_flag;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { mlo = fmin(mlo, s_mm[w][0]); mhi = fmax(mhi, s_mm[w][1]); }
        part_mminmax[2 * c] = mlo; part_mminmax[2 * c + 1] = mhi;
    }
}

// Objects of more than BIG_OBJECT_CHUNKS chunks (millions of particles; BASELINE configs[4] is ONE object of 1e8) get
// kernels of their own wherever the regular ones give an object to one warp or one CTA: a 1e8-particle object has
// ~1e5 chunks, and one CTA walking them cost 12.4 ms in bucket_scan_kernel and 1.1 ms in stats_final_kernel.
constexpr int BIG_OBJECT_CHUNKS = 2048;

// Combine the chunk partials of every object in chunk order.
__global__ void stats_final_kernel(const int32_t *chunk_first /* n_obj+1 */, const int32_t *size, int n_obj,
                                   const double *part_mass, const int32_t *part_spread, const double *part_mminmax,
                                   double *obj_mass, uint8_t *skip, double *obj_equal_mass)
{
    // a warp per object, lanes over its chunks (fixed assignment and reduction tree: deterministic mass)
    const int h = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (h >= n_obj) return;
    if (chunk_first[h + 1] - chunk_first[h] > BIG_OBJECT_CHUNKS) return;   // stats_final_big_kernel's
    const int lane = threadIdx.x & 31;
    double m = 0.0, mlo = 1e300, mhi = -1e300;
    int spread = 0;
    for (int c = chunk_first[h] + lane; c < chunk_first[h + 1]; c += 32) {
        m += part_mass[c];
        spread |= part_spread[c];
        mlo = fmin(mlo, part_mminmax[2 * c]); mhi = fmax(mhi, part_mminmax[2 * c + 1]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        m += __shfl_xor_sync(0xffffffffu, m, o);
        mlo = fmin(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
        mhi = fmax(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
    }
    if (lane != 0) return;
    obj_mass[h] = m;
    obj_equal_mass[h] = (mlo == mhi && mlo > 0.0) ? mlo : 0.0;   // > 0: every member has exactly this mass
    // The reference leaves an object untouched when calcLocalSpread(xyz) == 0 (helper_class.pyx:62-80 with
    // fp32 accumulators): in practice only an empty object gets there -- for coincident points the fp32
    // rounding of the mean already makes the spread non-zero.  `spread` is kept for diagnostics.
    (void)spread;
    skip[h] = (size[h] == 0) ? 1 : 0;
}

// The same for one big object: a CTA of 256 threads over its chunks, fixed assignment and reduction order.
__global__ void __launch_bounds__(256) stats_final_big_kernel(const int32_t *chunk_first, const int32_t *size, const int32_t *big,
                                                              const double *part_mass, const int32_t *part_spread,
                                                              const double *part_mminmax, double *obj_mass, uint8_t *skip,
                                                              double *obj_equal_mass)
{
    __shared__ double s_m[8], s_lo[8], s_hi[8];
    const int h = big[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double m = 0.0, mlo = 1e300, mhi = -1e300;
    for (int c = chunk_first[h] + (int)threadIdx.x; c < chunk_first[h + 1]; c += 256) {
        m += part_mass[c];
        mlo = fmin(mlo, part_mminmax[2 * c]); mhi = fmax(mhi, part_mminmax[2 * c + 1]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        m += __shfl_xor_sync(0xffffffffu, m, o);
        mlo = fmin(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
        mhi = fmax(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
    }
    if (lane == 0) { s_m[warp] = m; s_lo[warp] = mlo; s_hi[warp] = mhi; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    m = 0.0; mlo = 1e300; mhi = -1e300;
    for (int w = 0; w < 8; w++) { m += s_m[w]; mlo = fmin(mlo, s_lo[w]); mhi = fmax(mhi, s_hi[w]); }
    obj_mass[h] = m;
    obj_equal_mass[h] = (mlo == mhi && mlo > 0.0) ? mlo : 0.0;
    (void)part_spread;
    skip[h] = (size[h] == 0) ? 1 : 0;
}

// helper_class.pyx:98-100: `cdef float mass_total` accumulated sequentially in catalogue order
// (mass_total = (float)((double)mass_total + m)).  The order is part of the result, so one lane walks the
// object while the warp streams the masses in coalesced 32-element batches.  Objects whose members all have
// the same mass (every dark-matter-only snapshot) take a closed-form route: inside one fp32 binade the
// accumulator is a multiple of its ulp, so every step adds the same rounded increment and a whole binade is
// advanced in O(1) -- bit-identical to the sequential loop, including its saturation once m < ulp/2.
__device__ __forceinline__ float f32_step(float acc, double m) { return (float)__dadd_rn((double)acc, m); }

__device__ float mass32_equal(double m, int N)
{
    float acc = 0.0f;
    int i = 0;
    while (i < N) {
        const float next = f32_step(acc, m);
        if (next == acc) break;                       // saturated: no later addition changes the sum
        const float inc = next - acc;                 // exact: both are multiples of ulp(acc)'s grid or acc == 0
        int e;
        frexpf(acc, &e);                              // acc in [2^(e-1), 2^e)
        const float top = ldexpf(1.0f, e);
        // the increment is constant while acc stays in this binade AND acc + m is rounded on this binade's grid
        // AND no tie-to-even case occurs (then the increment alternates with the parity of acc)
        const double u = (double)ldexpf(1.0f, e - 24);          // ulp of the binade
        const double frac = m / u - floor(m / u);
        const bool constant = acc > 0.0f && frac != 0.5 && f32_step(next, m) - next == inc;
        if (!constant) { acc = next; i++; continue; }
        // steps k = 1..n with acc + k*inc + (m - inc) still below `top` behave identically
        long long n = (long long)floor(((double)top - (double)acc - m) / (double)inc);
        if (n < 1) { acc = next; i++; continue; }
        if (n > N - i) n = N - i;
        acc = (float)((double)acc + (double)n * (double)inc);   // exact: a multiple of u below top
        i += (int)n;
    }
    return acc;
}

__global__ void __launch_bounds__(128) mass32_kernel(const GatherParams G, int n_obj, const double *obj_equal_mass,
                                                     float *mass32)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < n_obj; h += warps) {
        const int64_t first = G.off[h];
        const int N = (int)(G.off[h + 1] - first);
        if (N == 0) {
            if (lane == 0) mass32[h] = 0.0f;
            continue;
        }
        const double m0 = obj_equal_mass[h];   // from the per-chunk min/max of the statistics kernels
        const bool same = m0 > 0.0;
        float acc = 0.0f;
        if (same) {
            acc = mass32_equal(m0, N);   // uniform across the warp
        } else {
            // Unequal masses: 32 members at a time.  While the accumulator stays inside one fp32 binade it is a
            // multiple k*u of that binade's ulp, and adding m rounds to k*u + rne(m'/u)*u with m' = m rounded on the
            // fp64 grid of the binade -- independent of k unless m'/u is an exact tie.  So a block whose members
            // all avoid ties and whose sum keeps the accumulator below the top of the binade is added with ONE
            // integer warp reduction, bit-identically to the sequential loop (checked on the host against the
            // literal loop for wide dynamic ranges, dyadic masses and occasional giants); any other block (binade
            // crossings, ties, the very first ones) falls back to the literal sequential additions.
            constexpr int PF = 8;   // blocks fetched ahead: the catalogue gather has ~1 us of latency per dependent pair
            for (int s0 = 0; s0 < N; s0 += 32 * PF) {
                double pre[PF];
#pragma unroll
                for (int b = 0; b < PF; b++) {
                    const int i = s0 + 32 * b + lane;
                    pre[b] = (i < N) ? G.masses[cat_index(G, h, first, first + i)] : 0.0;
                }
#pragma unroll
                for (int b = 0; b < PF; b++) {
                    const int i0 = s0 + 32 * b;
                    if (i0 >= N) break;
                    const double mine = pre[b];
                    const int cntb = min(32, N - i0);
                    int e;
                    frexpf(acc, &e);   // acc in [2^(e-1), 2^e)
                    const double top = ldexp(1.0, e), base = ldexp(1.0, e - 1), u = ldexp(1.0, e - 24);
                    const double mr = __dsub_rn(__dadd_rn(base, mine), base);
                    const double t = mr / u, fl = floor(t);
                    const bool lane_ok = (mine >= 0.0) && (mine < base) && (t - fl != 0.5);
                    long long q = lane_ok ? (long long)rint(t) : 0;
                    double mmax = mine;
                    const bool ok = acc > 0.0f && __all_sync(0xffffffffu, lane_ok);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        q += __shfl_xor_sync(0xffffffffu, q, o);
                        mmax = fmax(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
                    }
                    if (ok && (double)acc + (double)q * u + mmax < top) {
                        acc = (float)((double)acc + (double)q * u);   // exact: a multiple of u below the top
                    } else {
                        for (int j = 0; j < cntb; j++) acc = f32_step(acc, __shfl_sync(0xffffffffu, mine, j));
                    }
                }
            }
        }
        if (lane == 0) mass32[h] = acc;
    }
}

// Per-chunk partial sums of m*v (raw velocities through the catalogue), for vcenter.
__global__ void __launch_bounds__(256) vsum_chunk_kernel(const GatherParams G, const SoA soa, const Chunk *chunks,
                                                         int n_chunks, double *part /* [n_chunks][3] */)
{
    __shared__ double scratch[8 * 3];
    const int c = blockIdx.x;
    if (c >= n_chunks) return;
    const Chunk ck = chunks[c];
    const int N = soa.size[ck.halo];
    const int64_t first = G.off[ck.halo];
    const int end = min(N, ck.start + CHUNK);
    double v[3] = {0.0, 0.0, 0.0};
    for (int i = ck.start + threadIdx.x; i < end; i += blockDim.x) {
        const int64_t g = cat_index(G, ck.halo, first, first + i);
        const double m = G.masses[g];
        v[0] = fma(m, G.vxyz[3 * g], v[0]);
        v[1] = fma(m, G.vxyz[3 * g + 1], v[1]);
        v[2] = fma(m, G.vxyz[3 * g + 2], v[2]);
    }
    block_sum<3>(v, scratch);
    if (threadIdx.x == 0) { part[3 * c] = v[0]; part[3 * c + 1] = v[1]; part[3 * c + 2] = v[2]; }
}

__global__ void vcenter_final_kernel(const int32_t *chunk_first, int n_obj, const double *part, const float *mass32,
                                     double *vcenters)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_obj) return;
    double s[3] = {0.0, 0.0, 0.0};
    for (int c = chunk_first[h]; c < chunk_first[h + 1]; c++) {
        s[0] += part[3 * c]; s[1] += part[3 * c + 1]; s[2] += part[3 * c + 2];
    }
    const double M = (double)mass32[h];
    vcenters[3 * h] = s[0] / M; vcenters[3 * h + 1] = s[1] / M; vcenters[3 * h + 2] = s[2] / M;
}

// ---------------------------------------------------------------------------------------------------
// Radial ordering.  With q, s <= 1 the ellipsoidal radius of the reference satisfies
//   p0^2 + p1^2/q^2 + p2^2/s^2 >= |x - c|^2,
// so a particle with |x-c|^2 >= d_k^2 can never be a member of shell k in any Katz iteration.  The gather
// therefore writes each object's particles bucketed by "first shell k with r^2 < d_k^2 (1+1e-12)" (a stable
// counting sort, catalogue order kept inside a bucket -> deterministic summation order), and the morphology
// kernels only stream the prefix n_eff[h][k] = #{r^2 < d_k^2 (1+1e-12)} of shell k.  Results are
// unchanged (the skipped particles fail the reference's test in every iteration); the work drops from N
// to n_eff per halo-shell iteration.  Needs ascending radii and R <= 255.
// ---------------------------------------------------------------------------------------------------
constexpr int SORT_MAX_B = 256;

struct SortParams {
    GatherParams G;
    const Chunk *chunks;
    const int32_t *chunk_first;   // n_obj + 1
    const int32_t *size;          // n_obj
    const double *r200;           // n_obj
    const double *rr;             // R ascending
    int32_t n_chunks;
    int32_t R;
    uint8_t *bucket;              // (n_idx)
    int32_t *tile_hist;           // (n_chunks, R+1): counts, then exclusive offsets inside the object
    int32_t *dest;                // (n_idx) or nullptr
    int32_t *n_eff;               // (n_obj, R)
    // per-chunk statistics, produced by the gather while the chunk sits in shared memory (what stats_chunk_kernel
    // computes from a second read of the gathered streams): mass sum, spread flag, (min, max) mass
    double *part_mass;            // (n_chunks)
    int32_t *part_spread;         // (n_chunks)
    double *part_mminmax;         // (n_chunks, 2)
};

__device__ __forceinline__ void load_relative(const GatherParams &G, int h, int64_t first, int64_t j, double &dx,
                                              double &dy, double &dz, int64_t &g)
{
    g = cat_index(G, h, first, j);
    double x = G.xyz[3 * g], y = G.xyz[3 * g + 1], z = G.xyz[3 * g + 2];
    if (G.L_BOX != 0.0f) {
        const int64_t g0 = cat_index(G, h, first, first);
        pbc_reflect(x, y, z, G.xyz[3 * g0], G.xyz[3 * g0 + 1], G.L_BOX);
    }
    dx = __dsub_rn(x, G.centers[3 * h]);
    dy = __dsub_rn(y, G.centers[3 * h + 1]);
    dz = __dsub_rn(z, G.centers[3 * h + 2]);
}

#ifndef CPG_SORT_CHUNK
#define CPG_SORT_CHUNK 1024
#endif
constexpr int SORT_CHUNK = CPG_SORT_CHUNK;   // particles per CTA of the ordering kernels

__global__ void __launch_bounds__(256) bucket_hist_kernel(const SortParams P)
{
    __shared__ double s_thr[SORT_MAX_B];
    __shared__ int s_hist[SORT_MAX_B];
    const Chunk ck = P.chunks[blockIdx.x];
    const int h = ck.halo, R = P.R, B = R + 1;
    const int N = P.size[h];
    const int64_t first = P.G.off[h];
    const int end = min(N, ck.start + SORT_CHUNK);
    const double r200 = P.r200[h];
    for (int k = threadIdx.x; k < B; k += blockDim.x) {
        if (k < R) {
            const double d = __dmul_rn(r200, P.rr[k]);
            s_thr[k] = __dmul_rn(__dmul_rn(d, d), 1.000000000001);
        }
        s_hist[k] = 0;
    }
    __syncthreads();
    for (int i = ck.start + threadIdx.x; i < end; i += blockDim.x) {
        double dx, dy, dz;
        int64_t g;
        load_relative(P.G, h, first, first + i, dx, dy, dz, g);
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        int lo = 0, hi = R;   // smallest k with r2 < thr[k]; R if none
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (r2 < s_thr[mid]) hi = mid; else lo = mid + 1;
        }
        P.bucket[first + i] = (uint8_t)lo;
        atomicAdd(&s_hist[lo], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < B; k += blockDim.x) P.tile_hist[(size_t)blockIdx.x * B + k] = s_hist[k];
}

// Exclusive offsets of every (chunk, bucket) inside its object, bucket-major then chunk order (= stable), and
// the prefix lengths n_eff.  A warp per object, lanes over buckets.
__global__ void __launch_bounds__(256) bucket_scan_kernel(const SortParams P, int n_obj)
{
    // A CTA per object.  The (chunk, bucket) histogram of the object is row-major in the chunk: the warps split the
    // object's chunks into 8 contiguous segments and walk them row by row with the LANES OVER THE BUCKETS (coalesced rows,
    // independent loads; lanes over chunks read one word per 284-byte row and made an object of 10^3 chunks a 0.27 ms tail).
    __shared__ int s_part[8][SORT_MAX_B];
    __shared__ int s_tot[SORT_MAX_B], s_start[SORT_MAX_B];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.x;
    if (h >= n_obj) return;
    const int R = P.R, B = R + 1;
    const int c0 = P.chunk_first[h], c1 = P.chunk_first[h + 1];
    if (c1 - c0 > BIG_OBJECT_CHUNKS) return;   // the three segment kernels below
    const int per = (c1 - c0 + 7) / 8;
    const int s0 = min(c1, c0 + warp * per), s1 = min(c1, s0 + per);   // this warp's chunks
    constexpr int NB = SORT_MAX_B / 32;                               // buckets per lane: b = lane + 32 j
    int sum[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) sum[j] = 0;
#pragma unroll 4
    for (int c = s0; c < s1; c++) {
        const int32_t *row = P.tile_hist + (size_t)c * B;
#pragma unroll
        for (int j = 0; j < NB; j++)
            if (lane + 32 * j < B) sum[j] += row[lane + 32 * j];
    }
#pragma unroll
    for (int j = 0; j < NB; j++)
        if (lane + 32 * j < B) s_part[warp][lane + 32 * j] = sum[j];
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int tot = 0;
        for (int w = 0; w < 8; w++) tot += s_part[w][b];
        s_tot[b] = tot;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan over the buckets: where each bucket starts inside the object
        int carry = 0;
        for (int b0 = 0; b0 < B; b0 += 32) {
            const int b = b0 + lane;
            const int tot = b < B ? s_tot[b] : 0;
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (b < B) s_start[b] = carry + incl - tot;
            if (b < R) P.n_eff[(size_t)h * R + b] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    // per chunk: offset of its share of bucket b (chunk order = catalogue order): bucket start + earlier segments + the
    // running count inside this warp's segment
    int run[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) {
        run[j] = 0;
        const int b = lane + 32 * j;
        if (b < B) {
            run[j] = s_start[b];
            for (int w = 0; w < warp; w++) run[j] += s_part[w][b];
        }
    }
    // (8 rows are loaded before the first of them is overwritten: the compiler may not move a load above a store to the
    // same array, and one row per L2 round trip made the largest object a 0.18 ms tail)
#pragma unroll
    for (int j = 0; j < NB; j++) {
        if (32 * j >= B) break;   // warp-uniform
        const int b = lane + 32 * j;
        const bool on = b < B;
        int r = run[j];
        for (int c = s0; c < s1; c += 8) {
            int t[8];
#pragma unroll
            for (int u = 0; u < 8; u++) t[u] = (on && c + u < s1) ? P.tile_hist[(size_t)(c + u) * B + b] : 0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (on && c + u < s1) P.tile_hist[(size_t)(c + u) * B + b] = r;
                r += t[u];
            }
        }
    }
}

// bucket_scan_kernel for ONE big object, spread over the GPU: its chunks in segments of SCAN_SEG.
//   (1) per segment and bucket, the sum over the segment's chunks (thread = bucket: coalesced rows of tile_hist);
//   (2) one CTA: bucket totals, bucket starts, n_eff, and the offset of every (segment, bucket) inside the object;
//   (3) per segment and bucket, the running offsets of its chunks (in place, chunk order = catalogue order).
// Same offsets as bucket_scan_kernel produces (integers: no ordering issue).
constexpr int SCAN_SEG = 256;

__global__ void __launch_bounds__(256) bucket_segsum_kernel(const SortParams P, int h, int32_t *seg_tot)
{
    const int B = P.R + 1;
    const int c0 = P.chunk_first[h] + blockIdx.x * SCAN_SEG, c1 = min(P.chunk_first[h + 1], c0 + SCAN_SEG);
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int tot = 0;
        for (int c = c0; c < c1; c++) tot += P.tile_hist[(size_t)c * B + b];
        seg_tot[(size_t)blockIdx.x * B + b] = tot;
    }
}

__global__ void __launch_bounds__(256) bucket_segscan_kernel(const SortParams P, int h, int nseg, int32_t *seg_tot)
{
    __shared__ int s_tot[SORT_MAX_B], s_start[SORT_MAX_B];
    const int R = P.R, B = R + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int tot = 0;
        for (int sg = 0; sg < nseg; sg++) tot += seg_tot[(size_t)sg * B + b];
        s_tot[b] = tot;
    }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int b0 = 0; b0 < B; b0 += 32) {
            const int b = b0 + lane;
            const int tot = b < B ? s_tot[b] : 0;
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (b < B) s_start[b] = carry + incl - tot;
            if (b < R) P.n_eff[(size_t)h * R + b] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int run = s_start[b];
        for (int sg = 0; sg < nseg; sg++) {
            const int t = seg_tot[(size_t)sg * B + b];
            seg_tot[(size_t)sg * B + b] = run;
            run += t;
        }
    }
}

__global__ void __launch_bounds__(256) bucket_segapply_kernel(const SortParams P, int h, const int32_t *seg_tot)
{
    const int B = P.R + 1;
    const int c0 = P.chunk_first[h] + blockIdx.x * SCAN_SEG, c1 = min(P.chunk_first[h + 1], c0 + SCAN_SEG);
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int run = seg_tot[(size_t)blockIdx.x * B + b];
        for (int c = c0; c < c1; c++) {
            const int t = P.tile_hist[(size_t)c * B + b];
            P.tile_hist[(size_t)c * B + b] = run;
            run += t;
        }
    }
}

// Stable bucket ordering of one chunk + the gather itself.  The chunk's particles are staged in shared memory
// in bucket order, so that the global stores are runs of consecutive addresses (one run per bucket present
// in the chunk) instead of 4 isolated 8-byte stores per particle.
__global__ void __launch_bounds__(256) gather_bucketed_kernel(const SortParams P)
{
    extern __shared__ __align__(16) unsigned char s_gb[];
    double *sx = reinterpret_cast<double *>(s_gb), *sy = sx + SORT_CHUNK, *sz = sy + SORT_CHUNK, *sm = sz + SORT_CHUNK;
    int *sdest = reinterpret_cast<int *>(sm + SORT_CHUNK);
    int(*s_wh)[SORT_MAX_B] = reinterpret_cast<int(*)[SORT_MAX_B]>(sdest + SORT_CHUNK);
    int *s_bstart = &s_wh[8][0];
    const Chunk ck = P.chunks[blockIdx.x];
    const int h = ck.halo, B = P.R + 1;
    const int N = P.size[h];
    const int64_t first = P.G.off[h], pbase = P.G.poff[h];
    const int end = min(N, ck.start + SORT_CHUNK);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 9 * SORT_MAX_B; i += blockDim.x) (&s_wh[0][0])[i] = 0;
    __syncthreads();
    // every warp owns a contiguous eighth of the chunk, so that "earlier warp" == "earlier in the catalogue"
    constexpr int PER_WARP = SORT_CHUNK / 8, ROUNDS = PER_WARP / 32;
    const int w_lo = ck.start + warp * PER_WARP;
    int bk[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int i = w_lo + r * 32 + lane;
        bk[r] = (i < end) ? (int)P.bucket[first + i] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, bk[r]);
        if (bk[r] >= 0 && lane == __ffs(peers) - 1) s_wh[warp][bk[r]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if ((int)threadIdx.x < B) {   // per bucket: exclusive prefix over the warps, total into s_bstart
        int run = 0;
        for (int w = 0; w < 8; w++) {
            const int t = s_wh[w][threadIdx.x];
            s_wh[w][threadIdx.x] = run;
            run += t;
        }
        s_bstart[threadIdx.x] = run;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the bucket totals (warp scan, 32 buckets per step): where each bucket starts inside the chunk
        int carry = 0;
        for (int k0 = 0; k0 < B; k0 += 32) {
            const int k = k0 + lane;
            const int t = k < B ? s_bstart[k] : 0;
            int incl = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (k < B) s_bstart[k] = carry + incl - t;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    int lp[ROUNDS], gd[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int b = bk[r];
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        int rin = 0;
        if (b >= 0) rin = s_wh[warp][b] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (b >= 0 && lane == __ffs(peers) - 1) s_wh[warp][b] += __popc(peers);
        __syncwarp();
        lp[r] = b >= 0 ? s_bstart[b] + rin : -1;
        gd[r] = b >= 0 ? P.tile_hist[(size_t)blockIdx.x * B + b] + rin : 0;
    }
    // the data movement itself: 8 independent gathers per thread in flight
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        if (lp[r] >= 0) {
            const int i = w_lo + r * 32 + lane;
            double dx, dy, dz;
            int64_t g;
            load_relative(P.G, h, first, first + i, dx, dy, dz, g);
            sx[lp[r]] = dx; sy[lp[r]] = dy; sz[lp[r]] = dz;
            sm[lp[r]] = P.G.masses[g];
            sdest[lp[r]] = gd[r];
            if (P.dest) P.dest[first + i] = gd[r];
        }
    }
    __syncthreads();
    const int n = end - ck.start;
    double msum = 0.0, mlo = 1e300, mhi = -1e300;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int64_t o = pbase + sdest[j];
        const double mj = sm[j];
        P.G.dx[o] = sx[j]; P.G.dy[o] = sy[j]; P.G.dz[o] = sz[j]; P.G.m[o] = mj;
        msum += mj;
        mlo = fmin(mlo, mj); mhi = fmax(mhi, mj);
    }
    if (P.part_mass) {   // fixed reduction order: deterministic chunk totals
        for (int o = 16; o > 0; o >>= 1) {
            msum += __shfl_xor_sync(0xffffffffu, msum, o);
            mlo = fmin(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
            mhi = fmax(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
        }
        __syncthreads();   // sx is free now: reuse its first 24 doubles as reduction scratch
        if (lane == 0) { sx[warp] = msum; sx[8 + warp] = mlo; sx[16 + warp] = mhi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0, lo = 1e300, hi = -1e300;
            for (int w = 0; w < 8; w++) { t += sx[w]; lo = fmin(lo, sx[8 + w]); hi = fmax(hi, sx[16 + w]); }
            P.part_mass[blockIdx.x] = t;
            P.part_spread[blockIdx.x] = 1;
            P.part_mminmax[2 * blockIdx.x] = lo; P.part_mminmax[2 * blockIdx.x + 1] = hi;
        }
    }
    if (threadIdx.x == 0 && end == N && (N & 1)) {   // pad element of an odd-sized object
        P.G.dx[pbase + N] = 1e150; P.G.dy[pbase + N] = 1e150; P.G.dz[pbase + N] = 1e150; P.G.m[pbase + N] = 0.0;
    }
}

// The same gather without the staging of the particle data in shared memory (default; build option CPG_GATHER_STAGED=1
// brings the kernel above back).  Only the permutation goes through shared memory -- sorted slot -> (entry of the chunk,
// destination inside the object), 6 bytes per particle instead of 36 -- and every thread then loads the particle of its
// sorted slot straight from the snapshot (a random 24-byte read inside the chunk's 24 KB window: L1 / L2 hits) and stores
// it to the SoA streams, consecutive threads to consecutive slots of a bucket.  15 KB instead of 46 KB of shared
// memory per CTA (8 resident CTAs per SM instead of 4), one shared-memory round trip less per particle.  Same slots, same
// chunk statistics (same summation order) as the staged kernel.
#ifndef CPG_GATHER_MLP
#define CPG_GATHER_MLP 0
#endif
#ifndef CPG_GATHER_MINB
#define CPG_GATHER_MINB 8   // 32 registers: 64 resident warps per SM (measured: gather 3.18 -> 2.99 ms on configs[1])
#endif
#ifndef CPG_GATHER_BALLOT
#define CPG_GATHER_BALLOT 1   // peers from one ballot per key bit instead of MATCH.ANY (measured: gather 2.99 -> 2.91 ms)
#endif
__global__ void __launch_bounds__(256, CPG_GATHER_MINB) gather_bucketed2_kernel(const SortParams P)
{
    __shared__ int s_dest[SORT_CHUNK];
    __shared__ unsigned short s_src[SORT_CHUNK];
    __shared__ int s_wh[9][SORT_MAX_B];
    __shared__ double s_redm[24];
    static_assert(SORT_CHUNK <= 65536, "16-bit entry numbers");
    int *s_bstart = &s_wh[8][0];
    const Chunk ck = P.chunks[blockIdx.x];
    const int h = ck.halo, B = P.R + 1;
    const int N = P.size[h];
    const int64_t first = P.G.off[h], pbase = P.G.poff[h];
    const int end = min(N, ck.start + SORT_CHUNK);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 9 * SORT_MAX_B; i += blockDim.x) (&s_wh[0][0])[i] = 0;
    __syncthreads();
    // every warp owns a contiguous eighth of the chunk, so that "earlier warp" == "earlier in the catalogue"
    constexpr int PER_WARP = SORT_CHUNK / 8, ROUNDS = PER_WARP / 32;
    const int w_lo = ck.start + warp * PER_WARP;
    int bk[ROUNDS];
    unsigned pe[ROUNDS];   // lanes of the warp with the same bucket in this round (kept for the ranking below)
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int i = w_lo + r * 32 + lane;
        bk[r] = (i < end) ? (int)P.bucket[first + i] : -1;
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
#if CPG_GATHER_BALLOT
        {   // the lanes with the same 9-bit key (bucket, or all ones for "no particle"), from one ballot per key bit
            unsigned m = 0xffffffffu;
#pragma unroll
            for (int bit = 0; bit < 9; bit++) {
                const bool on = (bk[r] >> bit) & 1;
                const unsigned v = __ballot_sync(0xffffffffu, on);
                m &= on ? v : ~v;
            }
            pe[r] = m;
        }
#else
        pe[r] = __match_any_sync(0xffffffffu, bk[r]);
#endif
        if (bk[r] >= 0 && lane == __ffs(pe[r]) - 1) s_wh[warp][bk[r]] += __popc(pe[r]);
        __syncwarp();
    }
    __syncthreads();
    if ((int)threadIdx.x < B) {   // per bucket: exclusive prefix over the warps, total into s_bstart
        int run = 0;
        for (int w = 0; w < 8; w++) {
            const int t = s_wh[w][threadIdx.x];
            s_wh[w][threadIdx.x] = run;
            run += t;
        }
        s_bstart[threadIdx.x] = run;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the bucket totals: where each bucket starts inside the chunk
        int carry = 0;
        for (int k0 = 0; k0 < B; k0 += 32) {
            const int k = k0 + lane;
            const int t = k < B ? s_bstart[k] : 0;
            int incl = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (k < B) s_bstart[k] = carry + incl - t;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int b = bk[r];
        const unsigned peers = pe[r];
        int rin = 0;
        if (b >= 0) rin = s_wh[warp][b] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (b >= 0 && lane == __ffs(peers) - 1) s_wh[warp][b] += __popc(peers);
        __syncwarp();
        if (b >= 0) {
            const int i = w_lo + r * 32 + lane;
            const int lp = s_bstart[b] + rin;
            const int gd = P.tile_hist[(size_t)blockIdx.x * B + b] + rin;
            s_src[lp] = (unsigned short)(i - ck.start);
            s_dest[lp] = gd;
            if (P.dest) P.dest[first + i] = gd;
        }
    }
    __syncthreads();
    const int n = end - ck.start;
    double msum = 0.0, mlo = 1e300, mhi = -1e300;
#if CPG_GATHER_MLP
    {   // all loads of the thread's SORT_CHUNK / 256 particles first, then the stores (the compiler may not move loads across
        // the stores itself: the streams could alias the snapshot)
        constexpr int NJ = SORT_CHUNK / 256;
        double dxs[NJ], dys[NJ], dzs[NJ], ms[NJ];
        int64_t os[NJ];
#pragma unroll
        for (int r = 0; r < NJ; r++) {
            const int j = threadIdx.x + 256 * r;
            if (j < n) {
                const int i = ck.start + (int)s_src[j];
                int64_t g;
                load_relative(P.G, h, first, first + i, dxs[r], dys[r], dzs[r], g);
                ms[r] = P.G.masses[g];
                os[r] = pbase + s_dest[j];
            }
        }
#pragma unroll
        for (int r = 0; r < NJ; r++) {
            const int j = threadIdx.x + 256 * r;
            if (j < n) {
                P.G.dx[os[r]] = dxs[r]; P.G.dy[os[r]] = dys[r]; P.G.dz[os[r]] = dzs[r]; P.G.m[os[r]] = ms[r];
                msum += ms[r];
                mlo = fmin(mlo, ms[r]); mhi = fmax(mhi, ms[r]);
            }
        }
    }
#else
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int i = ck.start + (int)s_src[j];
        double dx, dy, dz;
        int64_t g;
        load_relative(P.G, h, first, first + i, dx, dy, dz, g);
        const double mj = P.G.masses[g];
        const int64_t o = pbase + s_dest[j];
        P.G.dx[o] = dx; P.G.dy[o] = dy; P.G.dz[o] = dz; P.G.m[o] = mj;
        msum += mj;
        mlo = fmin(mlo, mj); mhi = fmax(mhi, mj);
    }
#endif
    if (P.part_mass) {   // fixed reduction order: deterministic chunk totals
        for (int o = 16; o > 0; o >>= 1) {
            msum += __shfl_xor_sync(0xffffffffu, msum, o);
            mlo = fmin(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
            mhi = fmax(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
        }
        if (lane == 0) { s_redm[warp] = msum; s_redm[8 + warp] = mlo; s_redm[16 + warp] = mhi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0, lo = 1e300, hi = -1e300;
            for (int w = 0; w < 8; w++) { t += s_redm[w]; lo = fmin(lo, s_redm[8 + w]); hi = fmax(hi, s_redm[16 + w]); }
            P.part_mass[blockIdx.x] = t;
            P.part_spread[blockIdx.x] = 1;
            P.part_mminmax[2 * blockIdx.x] = lo; P.part_mminmax[2 * blockIdx.x + 1] = hi;
        }
    }
    if (threadIdx.x == 0 && end == N && (N & 1)) {   // pad element of an odd-sized object
        P.G.dx[pbase + N] = 1e150; P.G.dy[pbase + N] = 1e150; P.G.dz[pbase + N] = 1e150; P.G.m[pbase + N] = 0.0;
    }
}

#ifndef CPG_GATHER_STAGED
#define CPG_GATHER_STAGED 0
#endif

constexpr size_t GATHER_BUCKETED_SMEM = (size_t)SORT_CHUNK * (4 * sizeof(double) + sizeof(int)) + 9 * SORT_MAX_B * sizeof(int);

// ------------------------------------------------------------------------------------------------
// Prefix tensor table.  With the particles of an object in radial order, the particles of the whole 128-particle
// tiles below |x-c| < s*d are members of the shell's ellipsoid whatever its orientation (cpg_morph.cuh, InnerPrefix),
// and their contribution to the shape tensor is the same in every Katz iteration of every shell of the object.
// ptab[(toff(h) + t) * 6 + c] = sum over the first 128 (t+1) particles of object h of m * (xx, xy, xz, yy, yz, zz)
// (positions, or velocities about the velocity centre for the dispersion variant): the morphology kernels stream
// only the tiles past the inner prefix and add one table row.  toff(h) = (poff[h] >> 7) + h.
// Two kernels: per-tile sums (a warp per tile, every particle read once), then an in-place scan along each object.
// ------------------------------------------------------------------------------------------------
constexpr int PT_TILE = 128;   // == 2 * TILE_PAIRS of cpg_morph.cuh

__device__ __forceinline__ int64_t ptab_row0(const int64_t *poff, int h) { return (poff[h] >> 7) + h; }

__global__ void __launch_bounds__(256) tile_tensor_kernel(const SoA soa, const Chunk *chunks, int n_chunks, int use_vel,
                                                          double *ptab)
{
    const Chunk ck = chunks[blockIdx.x];   // 2048-particle chunks = 16 tiles, two per warp
    const int h = ck.halo, N = soa.size[h];
    const int64_t base = soa.poff[h];
    const int64_t row0 = ptab_row0(soa.poff, h);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *px = use_vel ? soa.vx : soa.dx, *py = use_vel ? soa.vy : soa.dy, *pz = use_vel ? soa.vz : soa.dz;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int t = ck.start / PT_TILE + 2 * warp + r;
        if ((int64_t)(t + 1) * PT_TILE > N) break;   // whole tiles only (warp-uniform)
        double a[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int64_t o = base + (int64_t)t * PT_TILE + 2 * (lane + 32 * q);
            const double2 x = *reinterpret_cast<const double2 *>(px + o), y = *reinterpret_cast<const double2 *>(py + o);
            const double2 z = *reinterpret_cast<const double2 *>(pz + o), m = *reinterpret_cast<const double2 *>(soa.m + o);
            a[0] = fma(m.x, x.x * x.x, a[0]); a[1] = fma(m.x, x.x * y.x, a[1]); a[2] = fma(m.x, x.x * z.x, a[2]);
            a[3] = fma(m.x, y.x * y.x, a[3]); a[4] = fma(m.x, y.x * z.x, a[4]); a[5] = fma(m.x, z.x * z.x, a[5]);
            a[0] = fma(m.y, x.y * x.y, a[0]); a[1] = fma(m.y, x.y * y.y, a[1]); a[2] = fma(m.y, x.y * z.y, a[2]);
            a[3] = fma(m.y, y.y * y.y, a[3]); a[4] = fma(m.y, y.y * z.y, a[4]); a[5] = fma(m.y, z.y * z.y, a[5]);
        }
#pragma unroll
        for (int c = 0; c < 6; c++) a[c] = warp_sum(a[c]);
        if (lane == 0) {
            double *row = ptab + (row0 + t) * 6;
#pragma unroll
            for (int c = 0; c < 6; c++) row[c] = a[c];
        }
    }
}

// In-place inclusive scan of the tile rows of every object: a CTA per object, warp w scans the w-th eighth of the
// tiles (32 tiles per step, fixed tree), then every warp adds the totals of the segments before its own (fixed order).
__global__ void __launch_bounds__(256) tile_tensor_scan_kernel(const SoA soa, int n_obj, double *ptab)
{
    __shared__ double s_seg[8][6];
    const int h = blockIdx.x;
    if (h >= n_obj) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nt = soa.size[h] / PT_TILE;
    if (nt == 0) return;
    double *rows = ptab + ptab_row0(soa.poff, h) * 6;
    const int per = (nt + 7) / 8;
    const int t_lo = min(nt, warp * per), t_hi = min(nt, t_lo + per);
    double carry[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double nxt[6];   // the next 32 rows are loaded before this round's stores (no load may move above a store to the same array)
#pragma unroll
    for (int c = 0; c < 6; c++) nxt[c] = t_lo + lane < t_hi ? rows[(size_t)(t_lo + lane) * 6 + c] : 0.0;
    for (int t0 = t_lo; t0 < t_hi; t0 += 32) {
        const int t = t0 + lane;
        double v[6];
#pragma unroll
        for (int c = 0; c < 6; c++) v[c] = nxt[c];
#pragma unroll
        for (int c = 0; c < 6; c++) nxt[c] = t + 32 < t_hi ? rows[(size_t)(t + 32) * 6 + c] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double u = __shfl_up_sync(0xffffffffu, v[c], o);
                if (lane >= o) v[c] += u;
            }
        }
#pragma unroll
        for (int c = 0; c < 6; c++) {
            v[c] += carry[c];
            if (t < t_hi) rows[(size_t)t * 6 + c] = v[c];
            carry[c] = __shfl_sync(0xffffffffu, v[c], 31);
        }
    }
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < 6; c++) s_seg[warp][c] = carry[c];
    __syncthreads();
    if (warp == 0) return;
    double add[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < warp; w++)
#pragma unroll
        for (int c = 0; c < 6; c++) add[c] += s_seg[w][c];
    for (int t = t_lo + lane; t < t_hi; t += 32)
#pragma unroll
        for (int c = 0; c < 6; c++) rows[(size_t)t * 6 + c] += add[c];
}

}  // namespace cpg
