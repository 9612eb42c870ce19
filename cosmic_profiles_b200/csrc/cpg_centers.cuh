// cpg_centers.cuh -- object centres on the device (SURVEY.md section 8(f) "next" row 1).
//
// Replaces the serial Python loops that precede every driver of the reference
// (cosmic_profiles/shape_profs/shape_profs_algos.pyx:586-591, dens_profs/dens_profs_algos.pyx:39-44, :96-101)
// and their helpers in cosmic_profiles/common/python_routines.py: respectPBCNoRef (:499-528, translation by
// +-L_BOX relative to a reference point), calcCoM (:530-547) and calcMode (:287-306, shrinking sphere: keep
// the particles within `rad` of the current centre of mass, shrink rad by 17 %, stop when fewer than 5 remain).
// One CTA per object; every sum is a fixed-order block reduction (deterministic).  Results agree with the
// reference's numpy arithmetic to ~1e-15 relative (different summation order), far inside the 1e-9 gate.
#pragma once
#include <cooperative_groups.h>

#include "cpg_gather.cuh"

namespace cpg {

struct CenterParams {
    GatherParams G;            // xyz, masses, idx, off, n_obj, L_BOX (as double below)
    const int32_t *size;
    const int32_t *order;      // objects largest first: block b works on object order[first_obj + b / cluster size]
    int32_t first_obj;         // offset into `order` (the cluster launch takes the head of the list, the CTA launch the rest)
    const double *ref;         // (n_obj,3) reference points of respectPBCNoRef / calcCoM, or nullptr (L_BOX == 0)
    double L_BOX;
    int32_t mode;              // 0: centre of mass, 1: shrinking-sphere mode
    int32_t shape_path;        // mode: 1 = start from the bounding-box extent, 0 = start from rad = 1000
    int32_t *idxA, *idxB;      // (n_idx) scratch: local indices of the surviving particles (mode only)
    double *centers;           // (n_obj,3) out
};

__device__ __forceinline__ void unwrapped(const CenterParams &P, int h, int64_t first, int i, const double *ref, double &x, double &y,
                                          double &z, double &m)
{
    const int64_t g = cat_index(P.G, h, first, first + i);
    x = P.G.xyz[3 * g]; y = P.G.xyz[3 * g + 1]; z = P.G.xyz[3 * g + 2];
    m = P.G.masses[g];
    if (P.L_BOX != 0.0) {   // python_routines.py:514-525
        const double half = P.L_BOX / 2;
        const double dx = x - ref[0], dy = y - ref[1], dz = z - ref[2];
        if (dx > half) x = x - P.L_BOX;
        if (dx < -half) x = x + P.L_BOX;
        if (dy > half) y = y - P.L_BOX;
        if (dy < -half) y = y + P.L_BOX;
        if (dz > half) z = z - P.L_BOX;
        if (dz < -half) z = z + P.L_BOX;
    }
}

template <int NV>
__device__ __forceinline__ void block_sum_bcast(double (&v)[NV], double *scratch /* [8*NV + NV] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0)
        for (int k = 0; k < NV; k++) scratch[warp * NV + k] = v[k];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += scratch[w * NV + threadIdx.x];
        scratch[8 * NV + threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = scratch[8 * NV + k];
}

__global__ void __launch_bounds__(256) centers_kernel(const CenterParams P)
{
    __shared__ double scratch[8 * 6 + 6];
    __shared__ int s_scan[8];
    __shared__ int s_total;
    const int h = P.order[P.first_obj + blockIdx.x];
    const int N = P.size[h];
    const int64_t first = P.G.off[h];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double ref[3] = {0.0, 0.0, 0.0};
    if (N == 0) {
        if (tid < 3) P.centers[3 * h + tid] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    if (P.ref) { ref[0] = P.ref[3 * h]; ref[1] = P.ref[3 * h + 1]; ref[2] = P.ref[3 * h + 2]; }
    if (P.mode == 0) {
        // calcCoM: sum(m * (x - ref) / M) + ref; without a reference point use the object's first particle
        double o[3] = {ref[0], ref[1], ref[2]};
        if (!P.ref) {
            double x, y, z, m;
            unwrapped(P, h, first, 0, ref, x, y, z, m);
            o[0] = x; o[1] = y; o[2] = z;
        }
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < N; i += blockDim.x) {
            double x, y, z, m;
            unwrapped(P, h, first, i, ref, x, y, z, m);
            v[0] = fma(m, x - o[0], v[0]); v[1] = fma(m, y - o[1], v[1]); v[2] = fma(m, z - o[2], v[2]);
            v[3] += m;
        }
        block_sum_bcast<4>(v, scratch);
        if (tid < 3) P.centers[3 * h + tid] = v[tid] / v[3] + o[tid];
        return;
    }
    // ---- calcMode
    int *A = P.idxA + first, *B = P.idxB + first;
    double rad = 1000.0;
    if (P.shape_path) {   // max over the axes of (max - min), shape_profs_algos.pyx:589
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int i = tid; i < N; i += blockDim.x) {
            double x, y, z, m;
            unwrapped(P, h, first, i, ref, x, y, z, m);
            lo[0] = fmin(lo[0], x); hi[0] = fmax(hi[0], x);
            lo[1] = fmin(lo[1], y); hi[1] = fmax(hi[1], y);
            lo[2] = fmin(lo[2], z); hi[2] = fmax(hi[2], z);
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            for (int o = 16; o > 0; o >>= 1) {
                lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
                hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
            }
        __syncthreads();
        if (lane == 0)
            for (int k = 0; k < 3; k++) { scratch[warp * 6 + k] = lo[k]; scratch[warp * 6 + 3 + k] = hi[k]; }
        __syncthreads();
        double ext = 0.0;
        for (int k = 0; k < 3; k++) {
            double l = 1e300, u = -1e300;
            for (int w = 0; w < 8; w++) { l = fmin(l, scratch[w * 6 + k]); u = fmax(u, scratch[w * 6 + 3 + k]); }
            ext = fmax(ext, u - l);
        }
        rad = ext;
        __syncthreads();
    }
    for (int i = tid; i < N; i += blockDim.x) A[i] = i;
    int n_alive = N;
    __syncthreads();
    for (int iter = 0; iter < 4096; iter++) {
        // centre of mass of the surviving particles: sum(x*m)/sum(m), python_routines.py:298
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j = tid; j < n_alive; j += blockDim.x) {
            double x, y, z, m;
            unwrapped(P, h, first, A[j], ref, x, y, z, m);
            v[0] = fma(x, m, v[0]); v[1] = fma(y, m, v[1]); v[2] = fma(z, m, v[2]);
            v[3] += m;
        }
        block_sum_bcast<4>(v, scratch);
        const double cx = v[0] / v[3], cy = v[1] / v[3], cz = v[2] / v[3];
        // keep the particles with |x - com| < rad (stable compaction, catalogue order preserved)
        int base = 0;
        for (int j0 = 0; j0 < n_alive; j0 += blockDim.x) {
            const int j = j0 + tid;
            int p = -1;
            bool keep = false;
            if (j < n_alive) {
                p = A[j];
                double x, y, z, m;
                unwrapped(P, h, first, p, ref, x, y, z, m);
                const double ax = x - cx, ay = y - cy, az = z - cz;
                keep = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az))) < rad;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_scan[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < 8; w++) {
                const int c = s_scan[w];
                if (w < warp) before += c;
                total += c;
            }
            if (keep) B[base + before + __popc(bal & ((1u << lane) - 1u))] = p;
            base += total;
            __syncthreads();
        }
        if (base < 5) {   // fewer than 5 left: the centre of mass just computed is the mode
            if (tid == 0) { P.centers[3 * h] = cx; P.centers[3 * h + 1] = cy; P.centers[3 * h + 2] = cz; }
            return;
        }
        int *t = A; A = B; B = t;
        n_alive = base;
        rad *= 0.83;
        __syncthreads();
    }
    if (tid == 0) { P.centers[3 * h] = 0.0; P.centers[3 * h + 1] = 0.0; P.centers[3 * h + 2] = 0.0; }
}

// calcMode for large objects: a thread-block cluster per object.  CTA r owns the r-th slice of the object's
// members and compacts the survivors of its slice in place (slice-local index lists), the partial sums and
// counts are exchanged through distributed shared memory in rank order (deterministic), two cluster barriers
// per shrinking step.  Same arithmetic as centers_kernel.
__global__ void __launch_bounds__(256) centers_mode_cluster_kernel(const CenterParams P)
{
    namespace cg = cooperative_groups;
    __shared__ double scratch[8 * 6 + 6];
    __shared__ double s_pub[2][8];   // published partial sums (double-buffered)
    __shared__ int s_cnt[2];
    __shared__ int s_scan[8];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int h = P.order[P.first_obj + blockIdx.x / C];
    const int N = P.size[h];
    const int64_t first = P.G.off[h];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double ref[3] = {0.0, 0.0, 0.0};
    if (P.ref) { ref[0] = P.ref[3 * h]; ref[1] = P.ref[3 * h + 1]; ref[2] = P.ref[3 * h + 2]; }
    const int lo = (int)((int64_t)N * rank / C), hi = (int)((int64_t)N * (rank + 1) / C);
    int *A = P.idxA + first + lo, *B = P.idxB + first + lo;
    double rad = 1000.0;
    int buf = 0;
    if (P.shape_path) {
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int i = lo + tid; i < hi; i += blockDim.x) {
            double x, y, z, m;
            unwrapped(P, h, first, i, ref, x, y, z, m);
            mn[0] = fmin(mn[0], x); mx[0] = fmax(mx[0], x);
            mn[1] = fmin(mn[1], y); mx[1] = fmax(mx[1], y);
            mn[2] = fmin(mn[2], z); mx[2] = fmax(mx[2], z);
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            for (int o = 16; o > 0; o >>= 1) {
                mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
                mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
            }
        if (lane == 0)
            for (int k = 0; k < 3; k++) { scratch[warp * 6 + k] = mn[k]; scratch[warp * 6 + 3 + k] = mx[k]; }
        __syncthreads();
        if (tid < 6) {
            double v = scratch[tid];
            for (int w = 1; w < 8; w++) v = (tid < 3) ? fmin(v, scratch[w * 6 + tid]) : fmax(v, scratch[w * 6 + tid]);
            s_pub[buf][tid] = v;
        }
        cluster.sync();
        double ext = 0.0;
        for (int k = 0; k < 3; k++) {
            double l = 1e300, u = -1e300;
            for (int r = 0; r < C; r++) {
                const double *rp = cluster.map_shared_rank(&s_pub[buf][0], r);
                l = fmin(l, rp[k]); u = fmax(u, rp[3 + k]);
            }
            ext = fmax(ext, u - l);
        }
        rad = ext;
        buf ^= 1;
    }
    for (int i = lo + tid; i < hi; i += blockDim.x) A[i - lo] = i;
    int n_alive = hi - lo;
    __syncthreads();
    for (int iter = 0; iter < 4096; iter++) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j = tid; j < n_alive; j += blockDim.x) {
            double x, y, z, m;
            unwrapped(P, h, first, A[j], ref, x, y, z, m);
            v[0] = fma(x, m, v[0]); v[1] = fma(y, m, v[1]); v[2] = fma(z, m, v[2]);
            v[3] += m;
        }
        block_sum_bcast<4>(v, scratch);
        if (tid < 4) s_pub[buf][tid] = v[tid];
        cluster.sync();
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = 0; r < C; r++) {
            const double *rp = cluster.map_shared_rank(&s_pub[buf][0], r);
#pragma unroll
            for (int k = 0; k < 4; k++) t[k] += rp[k];
        }
        const double cx = t[0] / t[3], cy = t[1] / t[3], cz = t[2] / t[3];
        int base = 0;
        for (int j0 = 0; j0 < n_alive; j0 += blockDim.x) {
            const int j = j0 + tid;
            int p = -1;
            bool keep = false;
            if (j < n_alive) {
                p = A[j];
                double x, y, z, m;
                unwrapped(P, h, first, p, ref, x, y, z, m);
                const double ax = x - cx, ay = y - cy, az = z - cz;
                keep = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az))) < rad;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_scan[warp] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < 8; w++) {
                const int c = s_scan[w];
                if (w < warp) before += c;
                total += c;
            }
            if (keep) B[base + before + __popc(bal & ((1u << lane) - 1u))] = p;
            base += total;
            __syncthreads();
        }
        if (tid == 0) s_cnt[buf] = base;
        cluster.sync();
        int total_alive = 0;
        for (int r = 0; r < C; r++) total_alive += *cluster.map_shared_rank(&s_cnt[buf], r);
        buf ^= 1;
        if (total_alive < 5) {
            if (rank == 0 && tid == 0) { P.centers[3 * h] = cx; P.centers[3 * h + 1] = cy; P.centers[3 * h + 2] = cz; }
            break;
        }
        int *tmp = A; A = B; B = tmp;
        n_alive = base;
        rad *= 0.83;
        if (iter == 4095 && rank == 0 && tid == 0) { P.centers[3 * h] = 0.0; P.centers[3 * h + 1] = 0.0; P.centers[3 * h + 2] = 0.0; }
    }
    cluster.sync();   // nobody leaves while a peer may still read its shared memory
}

}  // namespace cpg
