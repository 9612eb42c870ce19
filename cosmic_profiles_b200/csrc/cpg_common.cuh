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
__device__ __forceinline__ void jacobi_rotate(double &app, double &aqq, double &apq, double &arp, double &arq,
                                              double *vp, double *vq)
{
    if (apq != 0.0) {
        const double theta = (aqq - app) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));   // |theta| = inf -> t = 0
        t = theta < 0.0 ? -t : t;
        const double c = rsqrt(t * t + 1.0), s = t * c;
        app -= t * apq;
        aqq += t * apq;
        apq = 0.0;
        const double x = arp, y = arq;
        arp = c * x - s * y;
        arq = s * x + c * y;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double a = vp[r], b = vq[r];
            vp[r] = c * a - s * b;
            vq[r] = s * a + c * b;
        }
    }
}

__device__ __forceinline__ void eigh3(double a00, double a01, double a02, double a11, double a12, double a22,
                                      double w[3], double v[9])
{
    double v0[3] = {1.0, 0.0, 0.0}, v1[3] = {0.0, 1.0, 0.0}, v2[3] = {0.0, 0.0, 1.0};
    for (int sweep = 0; sweep < 60; sweep++) {
        const double off = a01 * a01 + a02 * a02 + a12 * a12;
        const double dia = a00 * a00 + a11 * a11 + a22 * a22;
        if (!(off > 1e-40 * dia))
            break;   // converged, or NaN input
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
#pragma unroll
    for (int r = 0; r < 3; r++) { v[r] = v0[r]; v[3 + r] = v1[r]; v[6 + r] = v2[r]; }
}

}  // namespace cpg
