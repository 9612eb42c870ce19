/*
 * cp_oracle.c -- CPU restatement of CosmicProfiles' shape / density hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (cosmic_profiles_b200/) never links, imports or
 * falls back to anything in oracle/.
 *
 * Parity pin: the reference's own tests hold no golden values (they assert array
 * shapes only, cosmic_profiles/tests/test_shapes.py:94-124), so this restatement
 * is pinned against the reference itself, compiled from /root/reference by
 * oracle/build_ref.py into oracle/_ref, on seeded inputs
 * (tests/test_oracle_golden.py against the tests/golden npz fixtures written by tests/golden/make_golden.py).
 *
 * Everything that decides an integer (particle membership, iteration count) is
 * written with the reference's exact operand order; build with
 *   gcc -O2 -ffp-contract=off -fopenmp
 * so no FMA contraction or re-association happens.  The single deliberate
 * difference is the 3x3 eigen-solver: the reference calls LAPACK zheevr from
 * scipy's OpenBLAS (cython_helpers/helper_class.pyx:123-164, not vendored under
 * /root/reference); here a cyclic Jacobi solver is used.  Eigenvalues agree to
 * ~1e-16 relative, eigenvectors up to sign.
 *
 * Reference files restated (paths relative to /root/reference/cosmic_profiles):
 *   shape_profs/shape_profs_algos.pyx:16-145    runShellAlgo        -> katz_iterate (shell_based=1)
 *   shape_profs/shape_profs_algos.pyx:148-262   runEllAlgo          -> katz_iterate (shell_based=0)
 *   shape_profs/shape_profs_algos.pyx:265-514   run*VDispAlgo       -> katz_iterate (tens != pos)
 *   shape_profs/shape_profs_algos.pyx:975-1240  calcObjMorph*       -> cpo_morph per-object part
 *   shape_profs/shape_profs_algos.pyx:517-972   calcMorph* drivers  -> cpo_morph (gather, PBC, masses)
 *   cython_helpers/helper_class.pyx:19-57       calcShapeTensor     -> shape_tensor
 *   cython_helpers/helper_class.pyx:62-80       calcLocalSpread     -> local_spread
 *   cython_helpers/helper_class.pyx:85-105      calcCoM (vcenter)   -> com_f32norm
 *   cython_helpers/helper_class.pyx:110-120     cython_abs          -> absf
 *   cython_helpers/helper_class.pyx:169-199     respectPBCNoRef     -> pbc_reflect
 *   cython_helpers/helper_class.pyx:204-241     calcDensProfBruteForceSph -> cpo_dens_sph
 *   cython_helpers/helper_class.pyx:246-306     calcDensProfBruteForceEll -> cpo_dens_ell
 *   cython_helpers/helper_class.pyx:417-428     calcKTilde          -> ktilde
 *   dens_profs/dens_profs_algos.pyx:54-280      density drivers     -> cpo_dens_*
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CPO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ helpers */

/* helper_class.pyx:110-120: abs() of a C float; the fp64 argument is truncated to fp32 at the call. */
static float absf(float x) { return x >= 0.0f ? x : -x; }

/* Cyclic Jacobi for a real symmetric 3x3 (stand-in for LAPACK zheevr, helper_class.pyx:155).
 * a: full symmetric matrix (row-major); w: eigenvalues ascending; v[r][k]: component r of eigenvector k. */
static void eigh3(const double a_in[3][3], double w[3], double v[3][3])
{
    double a[3][3];
    int i, j, k, sweep;
    for (i = 0; i < 3; i++)
        for (j = 0; j < 3; j++) {
            a[i][j] = a_in[i][j];
            v[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (sweep = 0; sweep < 60; sweep++) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        double dia = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (!(off > 1e-40 * dia)) /* also leaves on NaN */
            break;
        for (i = 0; i < 2; i++)
            for (j = i + 1; j < 3; j++) {
                double apq = a[i][j];
                if (apq == 0.0)
                    continue;
                double theta = (a[j][j] - a[i][i]) / (2.0 * apq);
                double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                if (theta < 0.0)
                    t = -t;
                if (!isfinite(theta))
                    t = 0.0; /* |theta| overflowed: rotation angle is zero to working precision */
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                a[i][i] -= t * apq;
                a[j][j] += t * apq;
                a[i][j] = a[j][i] = 0.0;
                k = 3 - i - j;
                {
                    double aki = a[k][i], akj = a[k][j];
                    a[k][i] = a[i][k] = c * aki - s * akj;
                    a[k][j] = a[j][k] = s * aki + c * akj;
                }
                for (k = 0; k < 3; k++) {
                    double vki = v[k][i], vkj = v[k][j];
                    v[k][i] = c * vki - s * vkj;
                    v[k][j] = s * vki + c * vkj;
                }
            }
    }
    w[0] = a[0][0];
    w[1] = a[1][1];
    w[2] = a[2][2];
    /* ascending order, columns follow */
    for (i = 0; i < 2; i++)
        for (j = 0; j < 2 - i; j++)
            if (w[j] > w[j + 1]) {
                double t = w[j];
                w[j] = w[j + 1];
                w[j + 1] = t;
                for (k = 0; k < 3; k++) {
                    t = v[k][j];
                    v[k][j] = v[k][j + 1];
                    v[k][j + 1] = t;
                }
            }
}

/* helper_class.pyx:19-57.  T = coordinates the tensor is built from (positions, or velocities
 * for the dispersion variant), tc = their centre.  fp32 mass normaliser, six sequential sums. */
static void shape_tensor(const double *T, const int *sel, int npts, const double *m, const double tc[3],
                         int reduced, const double *r_ell, double S[3][3])
{
    float mass_tot = 0.0f;
    int run, i, j;
    for (run = 0; run < npts; run++)
        mass_tot = mass_tot + m[sel[run]]; /* (float)((double)mass_tot + m) */
    for (i = 0; i < 3; i++)
        for (j = 0; j <= i; j++) {
            double acc = 0.0;
            for (run = 0; run < npts; run++) {
                int p = sel[run];
                if (!reduced)
                    acc = acc + m[p] * (T[3 * p + i] - tc[i]) * (T[3 * p + j] - tc[j]) / mass_tot;
                else
                    acc = acc + m[p] * (T[3 * p + i] - tc[i]) * (T[3 * p + j] - tc[j]) /
                                    (mass_tot * (r_ell[p] * r_ell[p]));
            }
            S[i][j] = acc;
            S[j][i] = acc;
        }
}

/* helper_class.pyx:62-80 (fp32 accumulators; only the ==0 test is consumed). */
static float local_spread(const double *x, int n)
{
    float spread = 0, mx = 0, my = 0, mz = 0;
    int i;
    for (i = 0; i < n; i++) {
        mx += x[3 * i + 0] / n;
        my += x[3 * i + 1] / n;
        mz += x[3 * i + 2] / n;
    }
    for (i = 0; i < n; i++) {
        double a = x[3 * i + 0] - mx, b = x[3 * i + 1] - my, c = x[3 * i + 2] - mz;
        spread += sqrt(a * a + b * b + c * c) / n;
    }
    return spread;
}

/* helper_class.pyx:85-105: centre of mass with an fp32, sequentially accumulated normaliser. */
static void com_f32norm(const double *x, const double *m, int n, double com[3])
{
    float mass_total = 0.0f;
    int run;
    com[0] = com[1] = com[2] = 0.0;
    for (run = 0; run < n; run++)
        mass_total += m[run];
    for (run = 0; run < n; run++) {
        com[0] += m[run] * x[3 * run + 0] / mass_total;
        com[1] += m[run] * x[3 * run + 1] / mass_total;
        com[2] += m[run] * x[3 * run + 2] / mass_total;
    }
}

/* helper_class.pyx:169-199: reflection (not translation) about L/2 relative to particle 0, fp32
 * distances, and the z test re-uses the (already updated) y distance (line 196). */
static void pbc_reflect(double *x, int n, float L_BOX)
{
    int i;
    if (L_BOX == 0.0f)
        return;
    for (i = 0; i < n; i++) {
        float dist_x, dist_y, dist_z;
        dist_x = absf((float)(x[0] - x[3 * i + 0]));
        if (dist_x > L_BOX / 2)
            x[3 * i + 0] = L_BOX - x[3 * i + 0];
        dist_y = absf((float)(x[1] - x[3 * i + 1]));
        if (dist_y > L_BOX / 2)
            x[3 * i + 1] = L_BOX - x[3 * i + 1];
        dist_z = absf((float)(x[1] - x[3 * i + 1]));
        if (dist_z > L_BOX / 2)
            x[3 * i + 2] = L_BOX - x[3 * i + 2];
    }
}

/* ------------------------------------------------------------- Katz iteration */

/* One halo-shell: shape_profs_algos.pyx:148-262 (E1), :16-145 (S1), :265-514 (dispersion variants).
 * pos (N,3) selects and defines r_ell; tens (N,3) feeds the tensor (== pos for shapes).
 * out[12] = d, q, s, major(3), inter(3), minor(3); all zero on failure.
 * n_iter = shape-tensor evaluations performed; n_pts = particle count of the last evaluation
 * (or the count present when the IT_WALL / IT_MIN test failed). */
static void katz_iterate(const double *pos, const double *tens, const double *m, int N, const double c[3],
                         const double tc[3], double d, double delta_d, int shell_based, double IT_TOL,
                         int IT_WALL, int IT_MIN, int reduced, int *sel, double *r_ell, double *princ,
                         double out[12], int *n_iter, int *n_pts)
{
    int pts = 0, i, iteration = 1, last_pts = 0;
    double err = 1.0, q_new = 1.0, s_new = 1.0, q_old, s_old;
    double S[3][3], w[3], v[3][3];
    for (i = 0; i < N; i++) {
        double ax = c[0] - pos[3 * i], ay = c[1] - pos[3 * i + 1], az = c[2] - pos[3 * i + 2];
        double r2 = ax * ax + ay * ay + az * az;
        int in = shell_based ? (r2 < d * d && r2 >= (d - delta_d) * (d - delta_d)) : (r2 < d * d);
        if (in)
            sel[pts++] = i;
        r_ell[i] = sqrt(ax * ax + ay * ay / (q_new * q_new) + az * az / (s_new * s_new));
    }
    while (err > IT_TOL) {
        if (iteration > IT_WALL || pts < IT_MIN) {
            memset(out, 0, 12 * sizeof(double));
            *n_iter = iteration - 1;
            *n_pts = pts;
            return;
        }
        last_pts = pts;
        shape_tensor(tens, sel, pts, m, tc, reduced, r_ell, S);
        eigh3(S, w, v);
        q_old = q_new;
        s_old = s_new;
        q_new = sqrt(w[1] / w[2]);
        s_new = sqrt(w[0] / w[2]);
        {
            double eq = absf((float)(q_new - q_old)) / q_old, es = absf((float)(s_new - s_old)) / s_old;
            err = eq > es ? eq : es; /* Python max(a, b): b if b > a else a */
        }
        {
            double n2 = sqrt(v[0][2] * v[0][2] + v[1][2] * v[1][2] + v[2][2] * v[2][2]);
            double n1 = sqrt(v[0][1] * v[0][1] + v[1][1] * v[1][1] + v[2][1] * v[2][1]);
            double n0 = sqrt(v[0][0] * v[0][0] + v[1][0] * v[1][0] + v[2][0] * v[2][0]);
            out[0] = d;
            out[1] = q_new;
            out[2] = s_new;
            for (i = 0; i < 3; i++) {
                out[3 + i] = v[i][2] / n2;
                out[6 + i] = v[i][1] / n1;
                out[9 + i] = v[i][0] / n0;
            }
        }
        for (i = 0; i < N; i++) {
            double dx = pos[3 * i] - c[0], dy = pos[3 * i + 1] - c[1], dz = pos[3 * i + 2] - c[2];
            double p0 = out[3] * dx + out[4] * dy + out[5] * dz;
            double p1 = out[6] * dx + out[7] * dy + out[8] * dz;
            double p2 = out[9] * dx + out[10] * dy + out[11] * dz;
            princ[3 * i] = p0;
            princ[3 * i + 1] = p1;
            princ[3 * i + 2] = p2;
            r_ell[i] = sqrt(p0 * p0 + p1 * p1 / (q_new * q_new) + p2 * p2 / (s_new * s_new));
        }
        pts = 0;
        if (!shell_based || d == delta_d) {
            for (i = 0; i < N; i++) {
                double p0 = princ[3 * i], p1 = princ[3 * i + 1], p2 = princ[3 * i + 2];
                if (p0 * p0 + p1 * p1 / (q_new * q_new) + p2 * p2 / (s_new * s_new) < d * d)
                    sel[pts++] = i;
            }
        } else {
            for (i = 0; i < N; i++) {
                double p0 = princ[3 * i], p1 = princ[3 * i + 1], p2 = princ[3 * i + 2];
                if (p0 * p0 + p1 * p1 / (q_new * q_new) + p2 * p2 / (s_new * s_new) < d * d &&
                    p0 * p0 / ((d - delta_d) * (d - delta_d)) +
                            p1 * p1 / ((q_new * d - delta_d) * (q_new * d - delta_d)) +
                            p2 * p2 / ((s_new * d - delta_d) * (s_new * d - delta_d)) >=
                        1)
                    sel[pts++] = i;
            }
        }
        iteration += 1;
    }
    *n_iter = iteration - 1;
    *n_pts = last_pts;
}

/* ---------------------------------------------------------------- drivers */

static int max_size(const int64_t *off, int n)
{
    int64_t mx = 1;
    int p;
    for (p = 0; p < n; p++)
        if (off[p + 1] - off[p] > mx)
            mx = off[p + 1] - off[p];
    return (int)mx;
}

/* calcMorphLocal / calcMorphGlobal / calcMorph*VelDisp (shape_profs_algos.pyx:517-972) with the
 * centres given (the reference computes them in Python before the prange, :586-591).
 *   local != 0: R shells at d = r200*r_over_r200[k] (:1035-1044); local == 0: one ellipsoid at
 *   d = r200 + SAFE (:1096), R must be 1.
 *   vxyz == NULL: position tensor.  vcenters_out (n,3) receives helper_class.calcCoM of velocities.
 * out: (n, R, 12) zeros on failure; n_iter / n_pts: (n, R), -1 where the per-object early exit hit. */
CPO_EXPORT int cpo_morph(const double *xyz, const double *vxyz, const double *masses, const int32_t *idx_cat,
                         const int64_t *offsets, int n_obj, const double *centers, const double *r200,
                         const double *r_over_r200, int R, int local, double SAFE, float L_BOX, double IT_TOL,
                         int IT_WALL, int IT_MIN, int reduced, int shell_based, double *out, int32_t *n_iter,
                         int32_t *n_pts, double *obj_mass, double *vcenters_out, int nthreads)
{
    int largest = max_size(offsets, n_obj);
    int p;
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#endif
    memset(out, 0, sizeof(double) * 12 * (size_t)n_obj * R);
    for (p = 0; p < n_obj * R; p++) {
        n_iter[p] = -1;
        n_pts[p] = -1;
    }
#pragma omp parallel
    {
        double *pos = malloc(sizeof(double) * 3 * largest);
        double *vel = vxyz ? malloc(sizeof(double) * 3 * largest) : NULL;
        double *princ = malloc(sizeof(double) * 3 * largest);
        double *m = malloc(sizeof(double) * largest);
        double *r_ell = malloc(sizeof(double) * largest);
        int *sel = malloc(sizeof(int) * largest);
        int q;
#pragma omp for schedule(dynamic)
        for (q = 0; q < n_obj; q++) {
            int N = (int)(offsets[q + 1] - offsets[q]), n, k;
            double mass = 0.0, vc[3] = {0, 0, 0};
            const double *c = centers + 3 * q;
            for (n = 0; n < N; n++) {
                int64_t g = idx_cat[offsets[q] + n];
                pos[3 * n] = xyz[3 * g];
                pos[3 * n + 1] = xyz[3 * g + 1];
                pos[3 * n + 2] = xyz[3 * g + 2];
                if (vel) {
                    vel[3 * n] = vxyz[3 * g];
                    vel[3 * n + 1] = vxyz[3 * g + 1];
                    vel[3 * n + 2] = vxyz[3 * g + 2];
                }
                m[n] = masses[g];
                mass = mass + masses[g];
            }
            if (obj_mass)
                obj_mass[q] = mass;
            pbc_reflect(pos, N, L_BOX);
            if (vel) {
                com_f32norm(vel, m, N, vc);
                if (vcenters_out)
                    memcpy(vcenters_out + 3 * q, vc, sizeof vc);
            }
            /* per-object early exits, shape_profs_algos.pyx:1024-1029 / :1093 */
            if (local_spread(pos, N) == 0.0f)
                continue;
            if (local && r200[q] == 0.0)
                continue;
            for (k = 0; k < R; k++) {
                double d, delta;
                if (local) {
                    d = r200[q] * r_over_r200[k];
                    delta = (k == 0) ? d : r200[q] * r_over_r200[k] - r200[q] * r_over_r200[k - 1];
                } else {
                    d = r200[q] + SAFE;
                    delta = d;
                }
                katz_iterate(pos, vel ? vel : pos, m, N, c, vel ? vc : c, d, delta, local ? shell_based : 0,
                             IT_TOL, IT_WALL, IT_MIN, reduced, sel, r_ell, princ,
                             out + 12 * ((size_t)q * R + k), n_iter + (size_t)q * R + k,
                             n_pts + (size_t)q * R + k);
            }
        }
        free(pos);
        free(vel);
        free(princ);
        free(m);
        free(r_ell);
        free(sel);
    }
    return 0;
}

/* calcDensProfsSphDirectBinning (dens_profs_algos.pyx:54-109) + calcDensProfBruteForceSph
 * (helper_class.pyx:204-241).  bin_edges (R+1) are built by the caller exactly as :83. */
CPO_EXPORT int cpo_dens_sph(const double *xyz, const double *masses, const int32_t *idx_cat,
                            const int64_t *offsets, int n_obj, const double *centers, const double *r200,
                            const double *bin_edges, int R, float L_BOX, double *dens, int32_t *counts,
                            int nthreads)
{
    int largest = max_size(offsets, n_obj);
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#endif
    memset(dens, 0, sizeof(double) * (size_t)n_obj * R);
#pragma omp parallel
    {
        double *pos = malloc(sizeof(double) * 3 * largest);
        double *m = malloc(sizeof(double) * largest);
        int q;
#pragma omp for schedule(dynamic)
        for (q = 0; q < n_obj; q++) {
            int N = (int)(offsets[q + 1] - offsets[q]), n, k;
            const double *c = centers + 3 * q;
            float r_200 = (float)r200[q]; /* the callee's parameter is a C float */
            for (n = 0; n < N; n++) {
                int64_t g = idx_cat[offsets[q] + n];
                pos[3 * n] = xyz[3 * g];
                pos[3 * n + 1] = xyz[3 * g + 1];
                pos[3 * n + 2] = xyz[3 * g + 2];
                m[n] = masses[g];
            }
            pbc_reflect(pos, N, L_BOX);
            for (k = 0; k < R; k++) {
                double hi = bin_edges[k + 1] * r_200, lo = bin_edges[k] * r_200;
                double vol = 4.0 / 3.0 * M_PI * (pow(hi, 3.0) - pow(lo, 3.0));
                double acc = 0.0;
                int cnt = 0;
                for (n = 0; n < N; n++) {
                    double ax = c[0] - pos[3 * n], ay = c[1] - pos[3 * n + 1], az = c[2] - pos[3 * n + 2];
                    double r2 = ax * ax + ay * ay + az * az;
                    if (r2 < hi * hi && r2 >= lo * lo) {
                        acc = acc + m[n] / vol;
                        cnt++;
                    }
                }
                dens[(size_t)q * R + k] = acc;
                if (counts)
                    counts[(size_t)q * R + k] = cnt;
            }
        }
        free(pos);
        free(m);
    }
    return 0;
}

/* calcDensProfBruteForceEll (helper_class.pyx:246-306) behind calcDensProfsEllDirectBinning
 * (dens_profs_algos.pyx:112-218).  a,b,c: (n, R+1) and major/inter/minor: (n, R, 3) are the
 * scipy-interpolated profiles (:169-195), computed by the caller.  Empty bin -> NaN. */
CPO_EXPORT int cpo_dens_ell(const double *xyz, const double *masses, const int32_t *idx_cat,
                            const int64_t *offsets, int n_obj, const double *centers, const double *a,
                            const double *b, const double *c_ax, const double *major, const double *inter,
                            const double *minor, int R, float L_BOX, double *dens, int32_t *counts, int nthreads)
{
    int largest = max_size(offsets, n_obj);
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#endif
    memset(dens, 0, sizeof(double) * (size_t)n_obj * R);
#pragma omp parallel
    {
        double *pos = malloc(sizeof(double) * 3 * largest);
        double *m = malloc(sizeof(double) * largest);
        int q;
#pragma omp for schedule(dynamic)
        for (q = 0; q < n_obj; q++) {
            int N = (int)(offsets[q + 1] - offsets[q]), n, k;
            const double *c = centers + 3 * q;
            const double *A = a + (size_t)q * (R + 1), *B = b + (size_t)q * (R + 1), *C = c_ax + (size_t)q * (R + 1);
            for (n = 0; n < N; n++) {
                int64_t g = idx_cat[offsets[q] + n];
                pos[3 * n] = xyz[3 * g];
                pos[3 * n + 1] = xyz[3 * g + 1];
                pos[3 * n + 2] = xyz[3 * g + 2];
                m[n] = masses[g];
            }
            pbc_reflect(pos, N, L_BOX);
            for (k = 0; k < R; k++) {
                const double *M2 = major + 3 * ((size_t)q * R + k), *M1 = inter + 3 * ((size_t)q * R + k),
                             *M0 = minor + 3 * ((size_t)q * R + k);
                /* norms are stored in C floats (helper_class.pyx:282-292) */
                float n2 = sqrt(M2[0] * M2[0] + M2[1] * M2[1] + M2[2] * M2[2]);
                float n1 = sqrt(M1[0] * M1[0] + M1[1] * M1[1] + M1[2] * M1[2]);
                float n0 = sqrt(M0[0] * M0[0] + M0[1] * M0[1] + M0[2] * M0[2]);
                double vol = 4.0 / 3.0 * M_PI * A[k + 1] * B[k + 1] * C[k + 1] - 4.0 / 3.0 * M_PI * A[k] * B[k] * C[k];
                double acc = 0.0;
                int cnt = 0;
                for (n = 0; n < N; n++) {
                    double dx = pos[3 * n] - c[0], dy = pos[3 * n + 1] - c[1], dz = pos[3 * n + 2] - c[2];
                    double p0 = M2[0] / n2 * dx + M2[1] / n2 * dy + M2[2] / n2 * dz;
                    double p1 = M1[0] / n1 * dx + M1[1] / n1 * dy + M1[2] / n1 * dz;
                    double p2 = M0[0] / n0 * dx + M0[1] / n0 * dy + M0[2] / n0 * dz;
                    double u0 = p0 / A[k + 1], u1 = p1 / B[k + 1], u2 = p2 / C[k + 1];
                    double w0 = p0 / A[k], w1 = p1 / B[k], w2 = p2 / C[k];
                    if (u0 * u0 + u1 * u1 + u2 * u2 < 1 && w0 * w0 + w1 * w1 + w2 * w2 >= 1) {
                        acc = acc + m[n] / vol;
                        cnt++;
                    }
                }
                dens[(size_t)q * R + k] = cnt ? acc : NAN;
                if (counts)
                    counts[(size_t)q * R + k] = cnt;
            }
        }
        free(pos);
        free(m);
    }
    return 0;
}

/* helper_class.pyx:417-428 as Cython 3 compiles it: float parameters, fp32 products and quotients,
 * fp64 exp, fp32 return. */
static float ktilde(float r, float r_i, float h_i)
{
    float t3 = r * r_i;
    float t4 = h_i * h_i;
    float t5 = -((r_i - r) * (r_i - r));
    float t6 = (float)(2.0 * (double)(h_i * h_i));
    float t7 = -((r_i + r) * (r_i + r));
    return (float)(1.0 / (2.0 * pow(2.0 * M_PI, 1.5)) * pow((double)(t3 / t4), -1.0) *
                   (exp((double)(t5 / t6)) - exp((double)(t7 / t6))));
}

/* calcDensProfsKernelBased (dens_profs_algos.pyx:221-280). */
CPO_EXPORT int cpo_dens_kernel(const double *xyz, const double *masses, const int32_t *idx_cat,
                               const int64_t *offsets, int n_obj, const double *centers, const double *r200,
                               const double *r_over_r200, int R, float L_BOX, double *dens, int nthreads)
{
    int largest = max_size(offsets, n_obj);
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#endif
    memset(dens, 0, sizeof(double) * (size_t)n_obj * R);
#pragma omp parallel
    {
        double *pos = malloc(sizeof(double) * 3 * largest);
        double *m = malloc(sizeof(double) * largest);
        double *dist = malloc(sizeof(double) * largest);
        double *hs = malloc(sizeof(double) * largest);
        int q;
#pragma omp for schedule(dynamic)
        for (q = 0; q < n_obj; q++) {
            int N = (int)(offsets[q + 1] - offsets[q]), n, k;
            const double *c = centers + 3 * q;
            for (n = 0; n < N; n++) {
                int64_t g = idx_cat[offsets[q] + n];
                pos[3 * n] = xyz[3 * g];
                pos[3 * n + 1] = xyz[3 * g + 1];
                pos[3 * n + 2] = xyz[3 * g + 2];
                m[n] = masses[g];
            }
            pbc_reflect(pos, N, L_BOX);
            for (n = 0; n < N; n++) {
                /* np.linalg.norm(xyz - c, axis=1); hs = 0.005*r200*(dists/(0.1*r200))**0.5 (:274-275) */
                double dx = pos[3 * n] - c[0], dy = pos[3 * n + 1] - c[1], dz = pos[3 * n + 2] - c[2];
                dist[n] = sqrt(dx * dx + dy * dy + dz * dz);
                hs[n] = 0.005 * r200[q] * sqrt(dist[n] / (0.1 * r200[q]));
            }
            for (k = 0; k < R; k++) {
                double acc = 0.0;
                float rf = (float)(r_over_r200[k] * r200[q]);
                for (n = 0; n < N; n++)
                    acc += m[n] * ktilde(rf, (float)dist[n], (float)hs[n]) / pow(hs[n], 3.0);
                dens[(size_t)q * R + k] = acc;
            }
        }
        free(pos);
        free(m);
        free(dist);
        free(hs);
    }
    return 0;
}

CPO_EXPORT int cpo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
