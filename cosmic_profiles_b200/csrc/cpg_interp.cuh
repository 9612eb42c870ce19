// cpg_interp.cuh -- shape-profile interpolation onto the density bins.
//
// Replaces the per-object loop of calcDensProfsEllDirectBinning that calls scipy.interpolate.interp1d twelve times
// (cosmic_profiles/dens_profs/dens_profs_algos.pyx:163-195): abscissae r_vec = a[p] (major-axis profile of object p,
// NaN where a shell did not converge), ordinates a, b, c evaluated at the bin edges * r200 and the nine eigenvector
// components at the bin centres * r200, kind='linear', bounds_error=False, fill_value='extrapolate'.
// What scipy 1.x does for that call, operation for operation: stable argsort of the abscissae (NaNs last),
// searchsorted(side='left') of every query, clip to [1, len-1], y = w_hi * y_hi + w_lo * y_lo with
// w_hi = (x - x_lo) / (x_hi - x_lo), w_lo = (x_hi - x) / (x_hi - x_lo)  (scipy/interpolate/_interpolate.py,
// interp1d._call_linear).  No FMA contraction: results are bit-identical to scipy's (tests/test_interp_rows.py pins
// the NumPy twin of this kernel against scipy itself, tests/test_gpu_parity.py this kernel against the twin).
// A warp per object; the profile (<= INTERP_MAX_RS shells) lives in shared memory.
#pragma once
#include "cpg_common.cuh"

namespace cpg {

constexpr int INTERP_MAX_RS = 256;
constexpr int INTERP_WARPS = 4;

struct InterpParams {
    const double *r200;        // (n_obj)
    const double *rr;          // (R) bin centres, units of r200
    const double *edges;       // (R + 1) bin edges, units of r200 (dens_profs_algos.pyx:155-156)
    const double *a, *b, *c;   // (n_obj, Rs)
    const double *major, *inter, *minor;   // (n_obj, Rs, 3)
    int32_t n_obj, R, Rs;
    double *ai, *bi, *ci;      // (n_obj, R + 1)
    double *mj, *it, *mn;      // (n_obj, R, 3)
};

__device__ __forceinline__ bool sort_less(double u, double v)   // NaNs last, like np.argsort
{
    return (u == u && v != v) || (u < v);
}

__global__ void __launch_bounds__(INTERP_WARPS * 32) shape_interp_kernel(const InterpParams Q)
{
    __shared__ double s_raw[INTERP_WARPS][INTERP_MAX_RS], s_x[INTERP_WARPS][INTERP_MAX_RS];
    __shared__ int s_perm[INTERP_WARPS][INTERP_MAX_RS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * INTERP_WARPS + warp;
    if (p >= Q.n_obj) return;
    const int L = Q.Rs;
    double *raw = s_raw[warp], *xs = s_x[warp];
    int *perm = s_perm[warp];
    const double *ap = Q.a + (size_t)p * L;
    for (int i = lane; i < L; i += 32) raw[i] = ap[i];
    __syncwarp();
    int n_finite = 0;
    for (int i = lane; i < L; i += 32) {   // stable rank of element i
        const double xi = raw[i];
        int r = 0;
        for (int j = 0; j < L; j++) {
            const double xj = raw[j];
            const bool lt = sort_less(xj, xi);
            const bool eq = !lt && !sort_less(xi, xj);
            r += (lt || (eq && j < i)) ? 1 : 0;
        }
        perm[r] = i;
        xs[r] = xi;
        n_finite += (xi == xi) ? 1 : 0;
    }
    n_finite = __reduce_add_sync(0xffffffffu, n_finite);
    __syncwarp();
    const double r200 = Q.r200[p];
    const int R = Q.R;
    // queries: R + 1 edges (a, b, c) then R centres (eigenvectors)
    for (int qi = lane; qi < 2 * R + 1; qi += 32) {
        const bool edge = qi <= R;
        const double xq = __dmul_rn(edge ? Q.edges[qi] : Q.rr[qi - R - 1], r200);
        int idx = 0;
        if (xq != xq) idx = n_finite;   // a NaN query lands behind every finite entry
        else
            for (int j = 0; j < L; j++) idx += (xs[j] < xq) ? 1 : 0;   // searchsorted(side='left'); NaN entries compare false
        idx = min(max(idx, 1), L - 1);
        const int lo = idx - 1, hi = idx;
        const double x_lo = xs[lo], x_hi = xs[hi];
        const double den = __dsub_rn(x_hi, x_lo);
        const double w_hi = __ddiv_rn(__dsub_rn(xq, x_lo), den), w_lo = __ddiv_rn(__dsub_rn(x_hi, xq), den);
        const int s_lo = perm[lo], s_hi = perm[hi];
        if (edge) {
            const size_t o = (size_t)p * (R + 1) + qi, base = (size_t)p * L;
            Q.ai[o] = __dadd_rn(__dmul_rn(w_hi, Q.a[base + s_hi]), __dmul_rn(w_lo, Q.a[base + s_lo]));
            Q.bi[o] = __dadd_rn(__dmul_rn(w_hi, Q.b[base + s_hi]), __dmul_rn(w_lo, Q.b[base + s_lo]));
            Q.ci[o] = __dadd_rn(__dmul_rn(w_hi, Q.c[base + s_hi]), __dmul_rn(w_lo, Q.c[base + s_lo]));
        } else {
            const size_t o = ((size_t)p * R + (qi - R - 1)) * 3;
            const size_t b_lo = ((size_t)p * L + s_lo) * 3, b_hi = ((size_t)p * L + s_hi) * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                Q.mj[o + k] = __dadd_rn(__dmul_rn(w_hi, Q.major[b_hi + k]), __dmul_rn(w_lo, Q.major[b_lo + k]));
                Q.it[o + k] = __dadd_rn(__dmul_rn(w_hi, Q.inter[b_hi + k]), __dmul_rn(w_lo, Q.inter[b_lo + k]));
                Q.mn[o + k] = __dadd_rn(__dmul_rn(w_hi, Q.minor[b_hi + k]), __dmul_rn(w_lo, Q.minor[b_lo + k]));
            }
        }
    }
}

}  // namespace cpg
