// cpg_fit.cuh -- batched density-profile fits.
//
// Replaces fitDensProfHelper -> fitDensProf -> toMinimize (cosmic_profiles/dens_profs/dens_profs_tools.py:325-359,
// :232-323, :227-230) and the four models (:22-79).  The reference maps scipy.optimize.minimize(method='Powell',
// bounds=...) over the objects in a process pool; the minimiser is scipy's (1.18.1, scipy/optimize/_optimize.py):
// _minimize_powell with _linesearch_powell / _line_for_search / _minimize_scalar_bounded.  Here one WARP (two for
// more than 32 bins) fits one object: the Powell / fmin control flow is uniform over its lanes (every lane carries the
// same scalars), the n <= 1024 terms of the objective are spread over the lanes and summed in NumPy's pairwise order (blocks of 128, halved recursively), so that every function value is the
// reference's up to the last-bit differences between CUDA's and NumPy's log / pow / exp / tan.
//
// Arithmetic is written with explicit round-to-nearest intrinsics: no FMA contraction (Python floats and NumPy
// scalars never fuse).  oracle/fit_oracle.py is the statement-by-statement CPU mirror of this file.
#pragma once
#include "cpg_common.cuh"

namespace cpg {

constexpr int FIT_MAX_BINS = 1024;   // bins per profile (shared memory: 3 arrays of R doubles per object)
constexpr int FIT_PW_BLOCK = 128;    // NumPy's pairwise sum is a single block up to 128 elements (PW_BLOCKSIZE)
constexpr int FIT_WARPS = 4;        // warps per CTA (4 objects, or 2 objects of two warps each)
enum { FIT_NFW = 0, FIT_HERNQUIST = 1, FIT_EINASTO = 2, FIT_ABG = 3 };

struct FitParams {
    const double *radii;    // (n_obj, R) kept radii r_over_r200 * r200, NaN bins squeezed out (host, NumPy)
    const double *lmed;     // (n_obj, R) np.log(median / np.average(median)) of the kept bins (host, NumPy)
    const int32_t *n_valid; // (n_obj) kept bins
    const double *avg;      // (n_obj) np.average(median)
    const double *x0;       // (n_obj, NPAR) initial guess (dens_profs_tools.py:276, :304, :317)
    double lb[5], ub[5];    // bounds (:277, :305, :318); ub = +inf when free
    int32_t n_obj, R, profile;
    double xtol, ftol;      // scipy defaults 1e-4, 1e-4
    double *best;           // (n_obj, NPAR)
    double *fun;            // (n_obj) final objective value
    int32_t *nfev, *nit;    // (n_obj)
};

namespace fit {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

// dens_profs_tools.py:22-79 on scalars.  k0 is the per-evaluation constant of the model (the same for every bin, so
// computing it once per evaluation is bit-identical): einasto -2/alpha, alpha_beta_gamma (beta-gamma)/alpha.
__device__ __forceinline__ double model_const(int profile, const double *x)
{
    if (profile == FIT_EINASTO) return dvd(-2.0, x[1]);
    if (profile == FIT_ABG) return dvd(sub(x[2], x[3]), x[1]);
    return 0.0;
}

__device__ __forceinline__ double model(int profile, double r, const double *x, double k0)
{
    if (profile == FIT_NFW) {            // rho_s/((r/r_s)*(1+r/r_s)**2)
        const double u = dvd(r, x[1]), w = add(1.0, u);
        return dvd(x[0], mul(u, mul(w, w)));
    }
    if (profile == FIT_HERNQUIST) {      // rho_s/((r/r_s)*(1+r/r_s)**3)
        const double u = dvd(r, x[1]), w = add(1.0, u);
        return dvd(x[0], mul(u, pow(w, 3.0)));
    }
    if (profile == FIT_EINASTO) {        // rho_s*np.exp(-2/alpha*((r/r_s)**alpha-1))
        const double u = dvd(r, x[2]);
        return mul(x[0], exp(mul(k0, sub(pow(u, x[1]), 1.0))));
    }
    // rho_s/((r/r_s)**gamma*(1+(r/r_s)**alpha)**((beta-gamma)/alpha))
    const double u = dvd(r, x[4]);
    return dvd(x[0], mul(pow(u, x[3]), pow(add(1.0, pow(u, x[1])), k0)));
}

// One object, W warps (W = 2 when it has more than 32 bins: a second round of bins on 8 of 32 lanes would double
// the latency of every evaluation).  All warps of an object carry identical state and take identical decisions:
// they read the same terms from shared memory.  All members are uniform over the object's lanes except lane / tid.
template <int N>
struct Problem {
    const double *radii, *lmed;
    double *terms;          // shared scratch, FIT_MAX_BINS doubles of this object
    int n, profile, lane;
    int tid, nthreads, bar_id;   // thread index inside the object's W warps, 32 W, named barrier (W > 1)

    __device__ __forceinline__ void obj_sync() const
    {
        if (nthreads == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    }

    int ncalls, maxfun;
    bool aborted;           // scipy's _MaxFuncCallError
    double lb[N], ub[N];

    // NumPy's pairwise sum (numpy/_core/src/umath/loops_utils.h.src) of a[0..n): one block of at most 128 elements --
    // n < 8 sequential, else 8 accumulators (lane j & 7 carries accumulator j), a fixed combine tree, a sequential tail
    __device__ double pw_block(const double *a, int n) const
    {
        double res;
        if (n < 8) {
            res = n > 0 ? a[0] : 0.0;
            for (int i = 1; i < n; i++) res = add(res, a[i]);
        } else {
            const int j = lane & 7;
            double r = a[j];
            int i = 8;
            for (; i < n - (n % 8); i += 8) r = add(r, a[i + j]);
            const double r0 = __shfl_sync(0xffffffffu, r, 0), r1 = __shfl_sync(0xffffffffu, r, 1);
            const double r2 = __shfl_sync(0xffffffffu, r, 2), r3 = __shfl_sync(0xffffffffu, r, 3);
            const double r4 = __shfl_sync(0xffffffffu, r, 4), r5 = __shfl_sync(0xffffffffu, r, 5);
            const double r6 = __shfl_sync(0xffffffffu, r, 6), r7 = __shfl_sync(0xffffffffu, r, 7);
            res = add(add(add(r0, r1), add(r2, r3)), add(add(r4, r5), add(r6, r7)));
            for (; i < n; i++) res = add(res, a[i]);
        }
        return res;
    }

    // ... and longer arrays: split at n2 = n / 2 rounded down to a multiple of 8, the two halves summed recursively
    // (depth 3 at FIT_MAX_BINS; uniform over the lanes)
    __device__ double pw_sum(const double *a, int n) const
    {
        if (n <= FIT_PW_BLOCK) return pw_block(a, n);
        int n2 = n / 2;
        n2 -= n2 % 8;
        const double lo = pw_sum(a, n2);
        return add(lo, pw_sum(a + n2, n - n2));
    }

    // toMinimize (dens_profs_tools.py:227-230) behind scipy's call counter; np.sum in NumPy's pairwise order
    __device__ double eval(const double (&x)[N])
    {
        if (ncalls >= maxfun) {
            aborted = true;
            return 0.0;
        }
        ncalls++;
        const double dn = (double)n;
        const double k0 = model_const(profile, x);
        for (int i = tid; i < n; i += nthreads) {
            const double t = sub(lmed[i], log(model(profile, radii[i], x, k0)));
            terms[i] = dvd(mul(t, t), dn);
        }
        obj_sync();
        const double res = n <= FIT_PW_BLOCK ? pw_block(terms, n) : pw_sum(terms, n);
        obj_sync();   // nobody overwrites the terms before every warp has summed them
        return res;
    }

    // scipy _line_for_search
    __device__ void line_for_search(const double (&x0)[N], const double (&alpha)[N], double &lmin, double &lmax) const
    {
        bool first = true;
        lmin = 0.0; lmax = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (alpha[i] == 0.0) continue;
            const double low = dvd(sub(lb[i], x0[i]), alpha[i]);
            const double high = dvd(sub(ub[i], x0[i]), alpha[i]);
            const double lo_i = alpha[i] > 0 ? low : high, hi_i = alpha[i] > 0 ? high : low;
            // np.max / np.min propagate NaN
            if (first) { lmin = lo_i; lmax = hi_i; first = false; }
            else {
                lmin = (lmin != lmin || lo_i != lo_i) ? nan("") : (lo_i > lmin ? lo_i : lmin);
                lmax = (lmax != lmax || hi_i != hi_i) ? nan("") : (hi_i < lmax ? hi_i : lmax);
            }
        }
        if (!(lmax >= lmin)) { lmin = 0.0; lmax = 0.0; }
    }
};

__device__ __forceinline__ double sgn(double v) { return (double)((v > 0) - (v < 0)); }

// scipy _minimize_scalar_bounded (Forsythe-Malcolm-Moler fmin); F: double -> double, may set *aborted.
template <class F>
__device__ void fminbound(F &func, const bool &aborted, double x1, double x2, double xatol, double &xmin, double &fmin)
{
    const double sqrt_eps = sqrt(2.2e-16);
    const double golden_mean = mul(0.5, sub(3.0, sqrt(5.0)));
    double a = x1, b = x2;
    double fulc = add(a, mul(golden_mean, sub(b, a)));
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = func(x);
    if (aborted) return;
    int num = 1;
    double ffulc = fx, fnfc = fx;
    double xm = mul(0.5, add(a, b));
    double tol1 = add(mul(sqrt_eps, fabs(xf)), dvd(xatol, 3.0));
    double tol2 = mul(2.0, tol1);
    while (fabs(sub(xf, xm)) > sub(tol2, mul(0.5, sub(b, a)))) {
        bool golden = true;
        if (fabs(e) > tol1) {
            golden = false;
            double r = mul(sub(xf, nfc), sub(fx, ffulc));
            double q = mul(sub(xf, fulc), sub(fx, fnfc));
            double p = sub(mul(sub(xf, fulc), q), mul(sub(xf, nfc), r));
            q = mul(2.0, sub(q, r));
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = rat;
            if ((fabs(p) < fabs(mul(mul(0.5, q), r))) && (p > mul(q, sub(a, xf))) && (p < mul(q, sub(b, xf)))) {
                rat = dvd(add(p, 0.0), q);
                x = add(xf, rat);
                if ((sub(x, a) < tol2) || (sub(b, x) < tol2)) {
                    const double d = sub(xm, xf);
                    const double si = add(sgn(d), d == 0 ? 1.0 : 0.0);
                    rat = mul(tol1, si);
                }
            } else {
                golden = true;
            }
        }
        if (golden) {
            e = (xf >= xm) ? sub(a, xf) : sub(b, xf);
            rat = mul(golden_mean, e);
        }
        const double si = add(sgn(rat), rat == 0 ? 1.0 : 0.0);
        const double ar = fabs(rat);
        // np.maximum propagates NaN
        const double step = (ar != ar || tol1 != tol1) ? nan("") : (ar > tol1 ? ar : tol1);
        x = add(xf, mul(si, step));
        const double fu = func(x);
        if (aborted) return;
        num++;
        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            fulc = nfc; ffulc = fnfc;
            nfc = xf; fnfc = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if ((fu <= fnfc) || (nfc == xf)) {
                fulc = nfc; ffulc = fnfc;
                nfc = x; fnfc = fu;
            } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
                fulc = x; ffulc = fu;
            }
        }
        xm = mul(0.5, add(a, b));
        tol1 = add(mul(sqrt_eps, fabs(xf)), dvd(xatol, 3.0));
        tol2 = mul(2.0, tol1);
        if (num >= 500) break;   // maxiter of _minimize_scalar_bounded
    }
    xmin = xf;
    fmin = fx;
}

template <int N>
struct Along {   // alpha -> f(p + alpha * xi), optionally alpha = tan(u)
    Problem<N> *P;
    const double *p, *xi;
    bool use_tan;
    __device__ double operator()(double u)
    {
        const double alpha = use_tan ? tan(u) : u;
        double y[N];
#pragma unroll
        for (int i = 0; i < N; i++) y[i] = add(p[i], mul(alpha, xi[i]));
        return P->eval(y);
    }
};

// scipy _linesearch_powell (bounded branches; with finite lower bounds the Brent branch is unreachable).
// Updates (fval, p, xi) in place unless the evaluation budget ran out (P.aborted), in which case nothing changes.
template <int N>
__device__ void linesearch(Problem<N> &P, double (&p)[N], double (&xi)[N], double tol, double &fval, bool &unsupported)
{
    bool any = false;
#pragma unroll
    for (int i = 0; i < N; i++) any = any || (xi[i] != 0.0);   // np.any: NaN counts as True
    if (!any) return;
    double b0, b1;
    P.line_for_search(p, xi, b0, b1);
    const bool inf0 = isinf(b0) && b0 < 0, inf1 = isinf(b1) && b1 > 0;
    if (inf0 && inf1) {
        unsupported = true;   // unbounded line search (Brent): cannot happen with the reference's bounds
        P.aborted = true;
        return;
    }
    Along<N> f{&P, p, xi, !(!inf0 && !inf1)};
    double amin = 0.0, fret = fval;
    if (!f.use_tan) fminbound(f, P.aborted, b0, b1, dvd(tol, 100.0), amin, fret);
    else {
        fminbound(f, P.aborted, atan(b0), atan(b1), dvd(tol, 100.0), amin, fret);
        amin = tan(amin);
    }
    if (P.aborted) return;
#pragma unroll
    for (int i = 0; i < N; i++) {
        xi[i] = mul(amin, xi[i]);
        p[i] = add(p[i], xi[i]);
    }
    fval = fret;
}

// scipy _minimize_powell with bounds, default options (maxiter = maxfev = N * 1000, direc = identity)
template <int N>
__device__ void powell(Problem<N> &P, double (&x)[N], double xtol, double ftol, double &fval_out, int &nit_out,
                       bool &unsupported)
{
    const int maxiter = N * 1000;
    double direc[N][N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) direc[i][j] = (i == j) ? 1.0 : 0.0;
    double fval = P.eval(x);
    double x1[N];
#pragma unroll
    for (int i = 0; i < N; i++) x1[i] = x[i];
    int iter = 0;
    const double tol = mul(xtol, 100.0);
    while (!P.aborted) {
        const double fx = fval;
        int bigind = 0;
        double delta = 0.0;
        for (int i = 0; i < N && !P.aborted; i++) {
            double direc1[N];
#pragma unroll
            for (int j = 0; j < N; j++) direc1[j] = direc[i][j];
            const double fx2 = fval;
            linesearch(P, x, direc1, tol, fval, unsupported);
            if (P.aborted) break;
            if (sub(fx2, fval) > delta) {
                delta = sub(fx2, fval);
                bigind = i;
            }
        }
        if (P.aborted) break;
        iter++;
        const double bnd = add(mul(ftol, add(fabs(fx), fabs(fval))), 1e-20);
        if (mul(2.0, sub(fx, fval)) <= bnd) break;
        if (P.ncalls >= P.maxfun) break;
        if (iter >= maxiter) break;
        if (fx != fx && fval != fval) break;
        double direc1[N], x2[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            direc1[i] = sub(x[i], x1[i]);
            x1[i] = x[i];
        }
        double lmin, lmax;
        P.line_for_search(x, direc1, lmin, lmax);
        // Python min(lmax, 1): 1 if (1 < lmax) else lmax -- a NaN lmax stays
        const double step = (1.0 < lmax) ? 1.0 : lmax;
#pragma unroll
        for (int i = 0; i < N; i++) x2[i] = add(x[i], mul(step, direc1[i]));
        const double fx2 = P.eval(x2);
        if (P.aborted) break;
        if (fx > fx2) {
            double t = mul(2.0, sub(add(fx, fx2), mul(2.0, fval)));
            double temp = sub(sub(fx, fval), delta);
            t = mul(t, mul(temp, temp));
            temp = sub(fx, fx2);
            t = sub(t, mul(delta, mul(temp, temp)));
            if (t < 0.0) {
                linesearch(P, x, direc1, tol, fval, unsupported);
                if (P.aborted) break;
                bool any = false;
#pragma unroll
                for (int i = 0; i < N; i++) any = any || (direc1[i] != 0.0);
                if (any) {
#pragma unroll
                    for (int j = 0; j < N; j++) {
                        // direc[bigind] = direc[-1]; direc[-1] = direc1   (bigind is data dependent: predicated copy)
#pragma unroll
                        for (int i = 0; i < N; i++)
                            if (i == bigind) direc[i][j] = direc[N - 1][j];
                        direc[N - 1][j] = direc1[j];
                    }
                }
            }
        }
    }
    fval_out = fval;
    nit_out = iter;
}

}  // namespace fit

template <int N, int W>
__global__ void __launch_bounds__(FIT_WARPS * 32) fit_profiles_kernel(const FitParams Q)
{
    constexpr int OBJS = FIT_WARPS / W;   // objects per CTA
    extern __shared__ __align__(16) double s_fit[];   // terms, radii, ln(median) of every object: 3 x OBJS rows of R (even) doubles
    const int rpad = (Q.R + 1) & ~1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp / W, sub = warp % W;
    const int obj = blockIdx.x * OBJS + slot;
    if (obj >= Q.n_obj) return;   // all W warps of the object
    fit::Problem<N> P;
    P.n = Q.n_valid[obj];
    P.lane = lane; P.tid = sub * 32 + lane; P.nthreads = 32 * W; P.bar_id = 1 + slot;
    for (int i = P.tid; i < P.n; i += P.nthreads) {
        s_fit[(OBJS + slot) * rpad + i] = Q.radii[(size_t)obj * Q.R + i];
        s_fit[(2 * OBJS + slot) * rpad + i] = Q.lmed[(size_t)obj * Q.R + i];
    }
    P.obj_sync();
    P.radii = s_fit + (OBJS + slot) * rpad; P.lmed = s_fit + (2 * OBJS + slot) * rpad; P.terms = s_fit + slot * rpad;
    P.profile = Q.profile; P.ncalls = 0; P.maxfun = N * 1000; P.aborted = false;
    double x[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        P.lb[i] = Q.lb[i]; P.ub[i] = Q.ub[i];
        x[i] = Q.x0[(size_t)obj * N + i];
    }
    double fval = 0.0;
    int nit = 0;
    bool unsupported = false;
    fit::powell<N>(P, x, Q.xtol, Q.ftol, fval, nit, unsupported);
    if (P.tid == 0) {
        // best_fit[0] *= np.average(median)   (dens_profs_tools.py:320)
        Q.best[(size_t)obj * N] = unsupported ? nan("") : fit::mul(x[0], Q.avg[obj]);
        for (int i = 1; i < N; i++) Q.best[(size_t)obj * N + i] = unsupported ? nan("") : x[i];
        Q.fun[obj] = fval;
        Q.nfev[obj] = unsupported ? -1 : P.ncalls;
        Q.nit[obj] = nit;
    }
}

}  // namespace cpg
