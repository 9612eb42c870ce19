// cpg_dens.cuh -- density profile estimators on sm_100a.
//
// Replaces CythonHelpers.calcDensProfBruteForceSph / calcDensProfBruteForceEll / calcKTilde
// (cosmic_profiles/cython_helpers/helper_class.pyx:204-241, :246-306, :417-428) and the prange bodies of
// calcDensProfs{SphDirectBinning,EllDirectBinning,KernelBased}
// (cosmic_profiles/dens_profs/dens_profs_algos.pyx:102-107, :210-215, :270-278).
//
// Membership predicates do not involve an eigen-solver, so they are evaluated with the reference's
// exact operand order through __d*_rn / __f*_rn intrinsics (never contracted): bin counts are bit-exact
// by construction.  One CTA per chunk of an object; per-chunk partial bins are combined in chunk order.
#pragma once
#include "cpg_gather.cuh"

namespace cpg {

constexpr int DENS_MAX_R = 1024;      // bins held in shared memory per CTA
constexpr int KDE_CHUNK = 2048;       // particles per CTA for the kernel estimator (8 per thread)

struct DensParams {
    SoA soa;
    const Chunk *chunks;
    int32_t n_chunks;
    int32_t R;
    const double *r200;        // n_obj
    const double *edges;       // sph: R+1 bin edges / kernel: R radii, in units of r200
    const double *a, *b, *c;   // ell: (n_obj, R+1)
    const double *major, *inter, *minor;   // ell: (n_obj, R, 3)
    int32_t monotone;          // sph: edges strictly increasing -> binary search is exact
    double ktilde_c0;          // 1/(2*(2*pi)^1.5) evaluated by the host libm like the reference
    const double *chunk_mmax;  // (n_chunks) largest |mass| of every chunk (chunk_mmax_kernel; anchors the fixed-point grid)
    double *part_mass;         // (n_chunks, R)
    int32_t *part_cnt;         // (n_chunks, R)
};

__device__ __forceinline__ double r2_exact(double x, double y, double z)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}


// ---- order-independent bin masses -----------------------------------------------------------------------
// A bin's mass is a sum over whatever threads find members: floating-point atomics would make its last bit depend on
// the arrival order.  Integer atomics are associative, so the masses are accumulated on a two-level fixed-point grid
// anchored at the chunk's largest mass: hi = rint(m / q1), lo = rint((m - hi q1) / q2), q1 = 2^(e+1-38), q2 = q1 2^-44,
// with 2^e <= max mass of the chunk < 2^(e+1).  |hi| <= 2^38 and |lo| <= 2^43, so a chunk of 16384 particles cannot
// overflow 63 bits; the rounding error per particle is <= 2^-83 of the chunk's largest mass (masses 25 orders of
// magnitude below it still keep 1e-10 relative).  The bin counts were integers all along.
struct MassGrid {
    double q1, inv_q1, q2, inv_q2;
};

__device__ __forceinline__ MassGrid mass_grid(double mmax)
{
    MassGrid g;
    const int e = (mmax > 0.0 && mmax < 1e300) ? ilogb(mmax) : 0;
    g.q1 = ldexp(1.0, e + 1 - 38); g.inv_q1 = ldexp(1.0, 38 - e - 1);
    g.q2 = ldexp(g.q1, -44); g.inv_q2 = ldexp(g.inv_q1, 44);
    return g;
}

__device__ __forceinline__ void grid_add(unsigned long long *hi, unsigned long long *lo, const MassGrid &g, double m)
{
    const long long h = __double2ll_rn(m * g.inv_q1);
    const double r = fma(-(double)h, g.q1, m);           // exact: |h| < 2^53 and the product is a multiple of q1
    const long long l = __double2ll_rn(r * g.inv_q2);
    atomicAdd(hi, (unsigned long long)h);
    atomicAdd(lo, (unsigned long long)l);
}

__device__ __forceinline__ double grid_value(unsigned long long hi, unsigned long long lo, const MassGrid &g)
{
    return (double)(long long)hi * g.q1 + (double)(long long)lo * g.q2;
}

// Largest |mass| of particles [start, end) of an object, identical in every thread of the CTA.
__device__ __forceinline__ double chunk_max_mass(const double *m, int64_t base, int start, int end, double *s_red /* [8] */)
{
    double mx = 0.0;
    for (int i = start + threadIdx.x; i < end; i += blockDim.x) mx = fmax(mx, fabs(m[base + i]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) mx = fmax(mx, s_red[w]);
    return mx;
}

// Largest |mass| per chunk of the gathered SoA: one pass over the masses, once per gather (the density kernels of every
// later call on the same gather reuse it).
__global__ void __launch_bounds__(256) chunk_mmax_kernel(const SoA soa, const Chunk *chunks, int n_chunks, double *chunk_mmax)
{
    __shared__ double s_red[8];
    if ((int)blockIdx.x >= n_chunks) return;
    const Chunk ck = chunks[blockIdx.x];
    const int end = min(soa.size[ck.halo], ck.start + CHUNK);
    const double mx = chunk_max_mass(soa.m, soa.poff[ck.halo], ck.start, end, s_red);
    if (threadIdx.x == 0) chunk_mmax[blockIdx.x] = mx;
}

// ---- spherical shells (helper_class.pyx:229-240) ---------------------------------------------------
__global__ void __launch_bounds__(256) dens_sph_kernel(const DensParams D)
{
    __shared__ double s_e2[DENS_MAX_R + 1];
    __shared__ unsigned long long s_hi[DENS_MAX_R], s_lo[DENS_MAX_R];
    __shared__ int s_cnt[DENS_MAX_R];
    const Chunk ck = D.chunks[blockIdx.x];
    const int N = D.soa.size[ck.halo];
    const int64_t base = D.soa.poff[ck.halo];
    const int end = min(N, ck.start + CHUNK);
    const int R = D.R;
    const double rf = (double)(float)D.r200[ck.halo];   // the callee's parameter is a C float (:204)
    for (int k = threadIdx.x; k <= R; k += blockDim.x) {
        const double e = __dmul_rn(D.edges[k], rf);
        s_e2[k] = __dmul_rn(e, e);
    }
    for (int k = threadIdx.x; k < R; k += blockDim.x) { s_hi[k] = 0ull; s_lo[k] = 0ull; s_cnt[k] = 0; }
    const MassGrid G = mass_grid(D.chunk_mmax[blockIdx.x]);
    __syncthreads();
    for (int i = ck.start + threadIdx.x; i < end; i += blockDim.x) {
        const double r2 = r2_exact(D.soa.dx[base + i], D.soa.dy[base + i], D.soa.dz[base + i]);
        const double m = D.soa.m[base + i];
        if (D.monotone) {
            if (r2 >= s_e2[0] && r2 < s_e2[R]) {
                int lo = 0, hi = R;   // s_e2[lo] <= r2 < s_e2[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_e2[mid] <= r2) lo = mid; else hi = mid;
                }
                grid_add(&s_hi[lo], &s_lo[lo], G, m);
                atomicAdd(&s_cnt[lo], 1);
            }
        } else {
            for (int k = 0; k < R; k++)
                if (r2 < s_e2[k + 1] && r2 >= s_e2[k]) {
                    grid_add(&s_hi[k], &s_lo[k], G, m);
                    atomicAdd(&s_cnt[k], 1);
                }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < R; k += blockDim.x) {
        D.part_mass[(size_t)blockIdx.x * R + k] = grid_value(s_hi[k], s_lo[k], G);
        D.part_cnt[(size_t)blockIdx.x * R + k] = s_cnt[k];
    }
}

// ---- ellipsoidal shells (helper_class.pyx:285-305) -------------------------------------------------
// per bin: 9 axis components divided by their fp32-stored norms, then a,b,c of both bounding ellipsoids
enum { EB_E2 = 0, EB_E1 = 3, EB_E0 = 6, EB_A1 = 9, EB_B1, EB_C1, EB_A0, EB_B0, EB_C0, EB_SLOTS = 16 };

__global__ void __launch_bounds__(256) dens_ell_kernel(const DensParams D)
{
    extern __shared__ double s_dyn[];   // [R][EB_SLOTS] frames, [R] + [R] fixed-point masses, then [R] int counts
    const int R = D.R;
    double *s_bin = s_dyn;
    unsigned long long *s_hi = reinterpret_cast<unsigned long long *>(s_dyn + (size_t)R * EB_SLOTS);
    unsigned long long *s_lo = s_hi + R;
    int *s_cnt = reinterpret_cast<int *>(s_lo + R);
    const Chunk ck = D.chunks[blockIdx.x];
    const int h = ck.halo;
    const int N = D.soa.size[h];
    const int64_t base = D.soa.poff[h];
    const int end = min(N, ck.start + CHUNK);
    for (int k = threadIdx.x; k < R; k += blockDim.x) {
        const double *M2 = D.major + 3 * ((size_t)h * R + k), *M1 = D.inter + 3 * ((size_t)h * R + k),
                     *M0 = D.minor + 3 * ((size_t)h * R + k);
        // `cdef float vec*_norm` (:282-284): the fp64 sqrt is stored in a C float
        const double n2 = (double)(float)sqrt(r2_exact(M2[0], M2[1], M2[2]));
        const double n1 = (double)(float)sqrt(r2_exact(M1[0], M1[1], M1[2]));
        const double n0 = (double)(float)sqrt(r2_exact(M0[0], M0[1], M0[2]));
        double *B = s_bin + (size_t)k * EB_SLOTS;
        for (int r = 0; r < 3; r++) {
            B[EB_E2 + r] = __ddiv_rn(M2[r], n2);
            B[EB_E1 + r] = __ddiv_rn(M1[r], n1);
            B[EB_E0 + r] = __ddiv_rn(M0[r], n0);
        }
        const double *A = D.a + (size_t)h * (R + 1), *Bx = D.b + (size_t)h * (R + 1), *C = D.c + (size_t)h * (R + 1);
        B[EB_A1] = A[k + 1]; B[EB_B1] = Bx[k + 1]; B[EB_C1] = C[k + 1];
        B[EB_A0] = A[k]; B[EB_B0] = Bx[k]; B[EB_C0] = C[k];
        s_hi[k] = 0ull; s_lo[k] = 0ull;
        s_cnt[k] = 0;
    }
    const MassGrid G = mass_grid(D.chunk_mmax[blockIdx.x]);
    __syncthreads();
    for (int i = ck.start + threadIdx.x; i < end; i += blockDim.x) {
        const double dx = D.soa.dx[base + i], dy = D.soa.dy[base + i], dz = D.soa.dz[base + i];
        const double m = D.soa.m[base + i];
        for (int k = 0; k < R; k++) {
            const double *B = s_bin + (size_t)k * EB_SLOTS;
            const double p0 = __dadd_rn(__dadd_rn(__dmul_rn(B[0], dx), __dmul_rn(B[1], dy)), __dmul_rn(B[2], dz));
            const double p1 = __dadd_rn(__dadd_rn(__dmul_rn(B[3], dx), __dmul_rn(B[4], dy)), __dmul_rn(B[5], dz));
            const double p2 = __dadd_rn(__dadd_rn(__dmul_rn(B[6], dx), __dmul_rn(B[7], dy)), __dmul_rn(B[8], dz));
            const double u0 = __ddiv_rn(p0, B[EB_A1]), u1 = __ddiv_rn(p1, B[EB_B1]), u2 = __ddiv_rn(p2, B[EB_C1]);
            if (r2_exact(u0, u1, u2) < 1.0) {
                const double w0 = __ddiv_rn(p0, B[EB_A0]), w1 = __ddiv_rn(p1, B[EB_B0]), w2 = __ddiv_rn(p2, B[EB_C0]);
                if (r2_exact(w0, w1, w2) >= 1.0) {
                    grid_add(&s_hi[k], &s_lo[k], G, m);
                    atomicAdd(&s_cnt[k], 1);
                }
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < R; k += blockDim.x) {
        D.part_mass[(size_t)blockIdx.x * R + k] = grid_value(s_hi[k], s_lo[k], G);
        D.part_cnt[(size_t)blockIdx.x * R + k] = s_cnt[k];
    }
}

// Combine chunk partials in chunk order and divide by the shell volume.
//   mode 0: spherical, vol = 4/3*pi*((e_{k+1} rf)^3 - (e_k rf)^3), empty bin -> 0   (:238-240)
//   mode 1: ellipsoidal, vol = 4/3*pi*a1*b1*c1 - 4/3*pi*a0*b0*c0, empty bin -> NaN   (:301-305)
__global__ void dens_bins_final_kernel(const DensParams D, const int32_t *chunk_first, int n_obj, int mode,
                                       double *dens, int32_t *counts)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int R = D.R;
    if (t >= (int64_t)n_obj * R) return;
    const int h = (int)(t / R), k = (int)(t % R);
    double m = 0.0;
    int n = 0;
    for (int c = chunk_first[h]; c < chunk_first[h + 1]; c++) {
        m += D.part_mass[(size_t)c * R + k];
        n += D.part_cnt[(size_t)c * R + k];
    }
    double vol;
    const double fourthirds_pi = __dmul_rn(4.0 / 3.0, 3.14159265358979323846);
    if (mode == 0) {
        const double rf = (double)(float)D.r200[h];
        const double hi = __dmul_rn(D.edges[k + 1], rf), lo = __dmul_rn(D.edges[k], rf);
        vol = __dmul_rn(fourthirds_pi, __dsub_rn(pow(hi, 3.0), pow(lo, 3.0)));
        dens[t] = n ? m / vol : 0.0;
    } else {
        const double *A = D.a + (size_t)h * (R + 1), *B = D.b + (size_t)h * (R + 1), *C = D.c + (size_t)h * (R + 1);
        const double v1 = __dmul_rn(__dmul_rn(__dmul_rn(fourthirds_pi, A[k + 1]), B[k + 1]), C[k + 1]);
        const double v0 = __dmul_rn(__dmul_rn(__dmul_rn(fourthirds_pi, A[k]), B[k]), C[k]);
        vol = __dsub_rn(v1, v0);
        dens[t] = n ? m / vol : __longlong_as_double(0x7ff8000000000000LL);
    }
    if (counts) counts[t] = n;
}

// ---- kernel-based estimator (dens_profs_algos.pyx:274-278, helper_class.pyx:428) ---------------------
// The reference compiles calcKTilde with float parameters and a float return: products and quotients of
// r, r_i, h_i are fp32, the two exp() are fp64, the result is truncated to fp32.
__device__ __forceinline__ float ktilde(float r, float ri, float hi, double c0)
{
    const float t3 = __fmul_rn(r, ri);
    const float t4 = __fmul_rn(hi, hi);
    const float dm = __fsub_rn(ri, r), dp = __fadd_rn(ri, r);
    const float t5 = -__fmul_rn(dm, dm);
    const float t6 = (float)__dmul_rn(2.0, (double)t4);
    const float t7 = -__fmul_rn(dp, dp);
    const double inv = __ddiv_rn(1.0, (double)__fdiv_rn(t3, t4));
    const double diff = __dsub_rn(exp((double)__fdiv_rn(t5, t6)), exp((double)__fdiv_rn(t7, t6)));
    return (float)__dmul_rn(__dmul_rn(c0, inv), diff);
}

constexpr int KDE_PER_THREAD = KDE_CHUNK / 256;

__global__ void __launch_bounds__(256) dens_kernel_kernel(const DensParams D)
{
    __shared__ double s_part[8][64];
    const Chunk ck = D.chunks[blockIdx.x];   // chunks of KDE_CHUNK particles here
    const int h = ck.halo;
    const int N = D.soa.size[h];
    const int64_t base = D.soa.poff[h];
    const int R = D.R;
    const double r200 = D.r200[h];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dist_f[KDE_PER_THREAD], h_f[KDE_PER_THREAD];
    double wgt[KDE_PER_THREAD];   // m / h^3
    const double hscale = __dmul_rn(0.005, r200), dscale = __dmul_rn(0.1, r200);
#pragma unroll
    for (int j = 0; j < KDE_PER_THREAD; j++) {
        const int i = ck.start + j * 256 + threadIdx.x;
        if (i < N && i < ck.start + KDE_CHUNK) {
            const double dist = sqrt(r2_exact(D.soa.dx[base + i], D.soa.dy[base + i], D.soa.dz[base + i]));
            const double hh = __dmul_rn(hscale, sqrt(__ddiv_rn(dist, dscale)));
            dist_f[j] = (float)dist;
            h_f[j] = (float)hh;
            wgt[j] = D.soa.m[base + i] / pow(hh, 3.0);
        } else {
            dist_f[j] = 1.0f; h_f[j] = 1.0f; wgt[j] = 0.0;
        }
    }
    for (int k0 = 0; k0 < R; k0 += 64) {
        const int kn = min(64, R - k0);
        for (int k = 0; k < kn; k++) {
            const float rf = (float)__dmul_rn(D.edges[k0 + k], r200);
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < KDE_PER_THREAD; j++)
                acc += wgt[j] == 0.0 ? 0.0 : wgt[j] * (double)ktilde(rf, dist_f[j], h_f[j], D.ktilde_c0);
            acc = warp_sum(acc);
            if (lane == 0) s_part[warp][k] = acc;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < kn; k += blockDim.x) {
            double s = 0.0;
            for (int w = 0; w < 8; w++) s += s_part[w][k];
            D.part_mass[(size_t)blockIdx.x * R + k0 + k] = s;
        }
        __syncthreads();
    }
}

__global__ void dens_kernel_final_kernel(const DensParams D, const int32_t *chunk_first, int n_obj, double *dens)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int R = D.R;
    if (t >= (int64_t)n_obj * R) return;
    const int h = (int)(t / R), k = (int)(t % R);
    double s = 0.0;
    for (int c = chunk_first[h]; c < chunk_first[h + 1]; c++) s += D.part_mass[(size_t)c * R + k];
    dens[t] = s;
}

}  // namespace cpg
