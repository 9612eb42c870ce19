/* A plain-C caller of libcosmicgpu.so: what a Cython `cdef extern from "cosmic_gpu.h"` (or any other FFI) would bind.
 * Builds one synthetic triaxial object, runs the global shape (calcMorphGlobal, shape_profs_algos.pyx:631-736) and
 * the spherical density profile (calcDensProfsSphDirectBinning, dens_profs_algos.pyx:54-109) through the C ABI and
 * prints the results as one line of numbers; tests/test_c_abi.py compiles and runs it and checks the values against
 * the oracle.  No Python, no torch, no C++ types in the interface. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cosmic_gpu.h"

static double u01(unsigned long long *s)
{
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)((*s >> 11) & ((1ULL << 53) - 1)) / 9007199254740992.0;
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 20000;
    double *xyz = malloc(sizeof(double) * 3 * n), *m = malloc(sizeof(double) * n);
    int32_t *idx = malloc(sizeof(int32_t) * n), size = n;
    unsigned long long s = 12345;
    for (int i = 0; i < n; i++) {   /* uniform ellipsoid with axes 1 : 0.7 : 0.4 around (50, 50, 50) */
        double x, y, z;
        do { x = 2 * u01(&s) - 1; y = 2 * u01(&s) - 1; z = 2 * u01(&s) - 1; } while (x * x + y * y + z * z >= 1.0);
        xyz[3 * i] = 50.0 + x; xyz[3 * i + 1] = 50.0 + 0.7 * y; xyz[3 * i + 2] = 50.0 + 0.4 * z;
        m[i] = 0.01;
        idx[i] = i;
    }
    const double centre[3] = {50.0, 50.0, 50.0}, r200 = 1.0;
    cpg_ctx *ctx = NULL;
    if (cpg_create(0, &ctx) != CPG_OK) { fprintf(stderr, "cpg_create: %s\n", cpg_last_error()); return 2; }
    if (cpg_set_particles(ctx, xyz, NULL, m, n) != CPG_OK || cpg_set_catalogue(ctx, idx, n, &size, 1) != CPG_OK) {
        fprintf(stderr, "upload: %s\n", cpg_last_error());
        return 3;
    }
    double out12[12], mass = 0.0;
    int32_t n_iter = 0, n_pts = 0;
    if (cpg_shape(ctx, centre, &r200, NULL, 1, 0, 6.0, 0.0f, 1e-2, 100, 10, 0, 0, 0, out12, &n_iter, &n_pts, &mass, NULL) != CPG_OK) {
        fprintf(stderr, "cpg_shape: %s\n", cpg_last_error());
        return 4;
    }
    const double edges[5] = {1e-8, 0.25, 0.5, 0.75, 1.0};
    double dens[4];
    int32_t counts[4];
    if (cpg_dens_sph(ctx, centre, &r200, edges, 4, 0.0f, dens, counts) != CPG_OK) {
        fprintf(stderr, "cpg_dens_sph: %s\n", cpg_last_error());
        return 5;
    }
    cpg_timing t;
    cpg_get_timing(ctx, &t);
    /* n, q, s, n_iter, n_pts, mass, counts[0..3], dens[0..3] */
    printf("%d %.17g %.17g %d %d %.17g %d %d %d %d %.17g %.17g %.17g %.17g\n", n, out12[1], out12[2], n_iter, n_pts, mass,
           counts[0], counts[1], counts[2], counts[3], dens[0], dens[1], dens[2], dens[3]);
    /* error convention: a bad argument returns a negative status and leaves a message, no exception, no abort */
    if (cpg_shape(ctx, NULL, &r200, NULL, 1, 0, 6.0, 0.0f, 1e-2, 100, 10, 0, 0, 0, out12, &n_iter, &n_pts, NULL, NULL) != CPG_ERR_ARGUMENT)
        return 6;
    cpg_destroy(ctx);
    free(xyz); free(m); free(idx);
    (void)t;
    return 0;
}
