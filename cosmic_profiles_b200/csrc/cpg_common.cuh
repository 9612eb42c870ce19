// cpg_common.cuh -- shared declarations of libcosmicgpu (sm_100a only, no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cosmic_gpu.h"

namespace cpg {

void set_error(const char *fmt, ...);

#define CPG_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            cpg::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return CPG_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// One unit of morphology work: up to `nshell` consecutive shells of one object share one read of the
// object's particles per Katz iteration.
struct Task {
    int32_t halo;
    int16_t shell0;
    int16_t nshell;
};

// Device-resident, centre-relative SoA view of the gathered catalogue.  Object h owns elements
// [poff[h], poff[h] + size[h]) of every stream; poff[h] is even (16-byte aligned double2 loads, TMA
// bulk copies) and the one pad element of odd-sized objects holds (1e150,1e150,1e150, m=0) so that no
// selection predicate can ever accept it.
struct SoA {
    const double *dx, *dy, *dz, *m;    // x - centre (bit-identical to the reference's per-use subtraction)
    const double *vx, *vy, *vz;        // v - vcentre (velocity-dispersion variant only)
    const int64_t *poff;               // n_obj + 1
    const int32_t *size;               // n_obj
};

struct MorphParams {
    SoA soa;
    const double *r200;        // n_obj
    const double *rr;          // R (local) or unused
    const int32_t *n_eff;      // (n_obj, R) radial-prefix lengths (cpg_gather.cuh) or nullptr = all particles
    const double *ptab;        // prefix tensor table (cpg_gather.cuh) or nullptr = no inner prefix (shell / reduced /
                               // no radial order): rows of 6 doubles, row (poff[h] >> 7) + h + t = first 128 (t+1) particles
    const uint8_t *skip;       // n_obj: 1 = per-object early exit (degenerate cloud), shape_profs_algos.pyx:1024
    const Task *tasks;
    int32_t n_tasks;
    int32_t *task_counter;     // dynamic work queue head
    int32_t R;
    int32_t local;
    double SAFE;
    double tol;
    int32_t wall;
    int32_t itmin;
    double *out12;             // (n_obj, R, 12), pre-zeroed
    int32_t *n_iter;           // (n_obj, R), pre-set to -1
    int32_t *n_pts;            // (n_obj, R), pre-set to -1
    unsigned long long *stats; // [0] tiles passed over, [1] (tile, shell) pairs among them; or nullptr
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;   // xor butterfly: every lane ends with the bit-identical total
}

__device__ __forceinline__ int warp_sum_int(int v)
{
    return __reduce_add_sync(0xffffffffu, v);
}

// ---------------------------------------------------------------------------------------------------
// 3x3 real symmetric eigen-solver (cyclic Jacobi), the in-register stand-in for the reference's LAPACK
// zheevr call (cython_helpers/helper_class.pyx:123-164).  Eigenvalues ascending in w[0..2]; v[3*k+r] is
// component r of the eigenvector of w[k] (unit length to ~1e-16, sign arbitrary like LAPACK's).
// ---------------------------------------------------------------------------------------------------
// Reciprocal and reciprocal square root to ~1 ulp from the SFU seed (MUFU.RCP64H / MUFU.RSQ64H, 2^-20ish)
// and two Newton steps -- a handful of DFMAs instead of the IEEE-exact division / sqrt sequences with their
// out-of-line special-case paths (the eigen-solve is instruction- and I-cache-bound).  Arguments are finite,
// non-zero and far from the under/overflow thresholds (the tensor is trace-normalised first).
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
}

#ifndef CPG_JACOBI_TOL2
#define CPG_JACOBI_TOL2 1e-33   // (off-diagonal norm / diagonal norm)^2 at which the sweeps stop: 3e-17 relative
#endif

__device__ __forceinline__ void jacobi_rotate(double &app, double &aqq, double &apq, double &arp, double &arq,
                                              double *vp, double *vq)
{
    if (apq != 0.0) {
        // tan of the rotation angle, smaller root: t = apq / (tau + sign(tau) * hypot(tau, apq)), tau = (aqq-app)/2.
        // (c, s) = (1, t) / sqrt(1 + t^2) is a rotation to rounding whatever the accuracy of t.
        const double tau = 0.5 * (aqq - app);
        const double h = fma(tau, tau, apq * apq);
        if (!(h > 1e-290)) {   // both the gap and the coupling vanish below 1e-145 of the trace: nothing to rotate
            apq = 0.0;
            return;
        }
        const double r = h * fast_rsqrt(h);
        const double t = apq * fast_rcp(tau + copysign(r, tau));
        const double c = fast_rsqrt(fma(t, t, 1.0)), s = t * c;
        app = fma(-t, apq, app);
        aqq = fma(t, apq, aqq);
        apq = 0.0;
        const double x = arp, y = arq;
        arp = c * x - s * y;
        arq = s * x + c * y;
#pragma unroll
        for (int r3 = 0; r3 < 3; r3++) {
            const double a = vp[r3], b = vq[r3];
            vp[r3] = c * a - s * b;
            vq[r3] = s * a + c * b;
        }
    }
}

// a: symmetric input (a00 a01 a02 a11 a12 a22).  v (in/out): on entry an orthonormal basis that is expected
// to nearly diagonalise `a` (the eigenvectors of the previous Katz iteration, or the identity); on exit the
// eigenvectors sorted by ascending eigenvalue.  Working on V^T a V starts the cyclic Jacobi sweeps close to
// convergence (quadratic), typically 2-3 sweeps instead of 5-6.
__device__ __forceinline__ void eigh3(double a00, double a01, double a02, double a11, double a12, double a22,
                                      double w[3], double v[9])
{
    double v0[3] = {v[0], v[1], v[2]}, v1[3] = {v[3], v[4], v[5]}, v2[3] = {v[6], v[7], v[8]};
    {
        // b_k = a * v_k, then a' = V^T a V
        double b0[3], b1[3], b2[3];
#define CPG_MATVEC(b, x)                                        \
        b[0] = fma(a02, x[2], fma(a01, x[1], a00 * x[0]));      \
        b[1] = fma(a12, x[2], fma(a11, x[1], a01 * x[0]));      \
        b[2] = fma(a22, x[2], fma(a12, x[1], a02 * x[0]));
        CPG_MATVEC(b0, v0) CPG_MATVEC(b1, v1) CPG_MATVEC(b2, v2)
#undef CPG_MATVEC
#define CPG_DOT(x, y) fma(x[2], y[2], fma(x[1], y[1], x[0] * y[0]))
        a00 = CPG_DOT(v0, b0); a01 = CPG_DOT(v0, b1); a02 = CPG_DOT(v0, b2);
        a11 = CPG_DOT(v1, b1); a12 = CPG_DOT(v1, b2); a22 = CPG_DOT(v2, b2);
#undef CPG_DOT
    }
    for (int sweep = 0; sweep < 40; sweep++) {
        const double off = a01 * a01 + a02 * a02 + a12 * a12;
        const double dia = a00 * a00 + a11 * a11 + a22 * a22;
        if (!(off > CPG_JACOBI_TOL2 * dia))
            break;   // off-diagonal below sqrt(CPG_JACOBI_TOL2) of the diagonal (eigenvectors exact to rounding), or NaN input
        jacobi_rotate(a00, a11, a01, a02, a12, v0, v1);
        jacobi_rotate(a00, a22, a02, a01, a12, v0, v2);
        jacobi_rotate(a11, a22, a12, a01, a02, v1, v2);
    }
    // sort ascending (3-element network), eigenvectors follow
#define CPG_CSWAP(wa, wb, va, vb)                     \
    if (wa > wb) {                                    \
        double t_ = wa; wa = wb; wb = t_;             \
        for (int r = 0; r < 3; r++) { t_ = va[r]; va[r] = vb[r]; vb[r] = t_; } \
    }
    CPG_CSWAP(a00, a11, v0, v1)
    CPG_CSWAP(a11, a22, v1, v2)
    CPG_CSWAP(a00, a11, v0, v1)
#undef CPG_CSWAP
    w[0] = a00; w[1] = a11; w[2] = a22;
    // re-orthonormalise cheaply (one Gram-Schmidt pass keeps the warm-started basis orthonormal to rounding
    // over many Katz iterations): v2 unit, v1 orthogonal to v2, v0 = v1 x v2
    {
        double n = fast_rsqrt(fma(v2[2], v2[2], fma(v2[1], v2[1], v2[0] * v2[0])));
        v2[0] *= n; v2[1] *= n; v2[2] *= n;
        const double d = fma(v1[2], v2[2], fma(v1[1], v2[1], v1[0] * v2[0]));
        v1[0] = fma(-d, v2[0], v1[0]); v1[1] = fma(-d, v2[1], v1[1]); v1[2] = fma(-d, v2[2], v1[2]);
        n = fast_rsqrt(fma(v1[2], v1[2], fma(v1[1], v1[1], v1[0] * v1[0])));
        v1[0] *= n; v1[1] *= n; v1[2] *= n;
        const double c0 = v1[1] * v2[2] - v1[2] * v2[1], c1 = v1[2] * v2[0] - v1[0] * v2[2],
                     c2 = v1[0] * v2[1] - v1[1] * v2[0];
        const double sgn = fma(c2, v0[2], fma(c1, v0[1], c0 * v0[0])) < 0.0 ? -1.0 : 1.0;
        v0[0] = sgn * c0; v0[1] = sgn * c1; v0[2] = sgn * c2;
    }
#pragma unroll
    for (int r = 0; r < 3; r++) { v[r] = v0[r]; v[3 + r] = v1[r]; v[6 + r] = v2[r]; }
}

// Sum NV = 8, 16 or 32 per-lane values over the warp with the halving exchange: after log2 steps lane l holds
// the warp total of value (l mod NV).  31 shuffles for 32 values instead of 160 with one butterfly each;
// the summation tree is fixed, so the result is deterministic.
template <int NV>
__device__ __forceinline__ double warp_sum_transpose(double (&val)[NV], int lane)
{
    static_assert(NV == 8 || NV == 16 || NV == 32, "NV");
    // lanes that differ in a bit >= log2(NV) hold the same value index: plain butterfly over those bits
#pragma unroll
    for (int o = 16; o >= NV; o >>= 1) {
#pragma unroll
        for (int j = 0; j < NV; j++) val[j] += __shfl_xor_sync(0xffffffffu, val[j], o);
    }
#pragma unroll
    for (int n = NV / 2; n >= 1; n >>= 1) {
        const bool upper = (lane & n) != 0;
#pragma unroll
        for (int j = 0; j < n; j++) {
            const double send = upper ? val[j] : val[j + n];
            const double keep = upper ? val[j + n] : val[j];
            val[j] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return val[0];
}

}  // namespace cpg
