// cpg_morph.cuh -- the Katz-Dubinski morphology loop on sm_100a.
//
// Replaces runEllAlgo / runShellAlgo / runEllVDispAlgo / runShellVDispAlgo
// (cosmic_profiles/shape_profs/shape_profs_algos.pyx:148-262, :16-145, :265-383, :386-514),
// CythonHelpers.calcShapeTensor (cython_helpers/helper_class.pyx:19-57) and the LAPACK call at :155.
//
// Formulation: iteration k's tensor uses the membership and r_ell produced with
// frame k-1, so "rotate with frame k-1 -> test -> accumulate" is ONE pass over (dx,dy,dz,m[,dv]) per
// iteration; nothing per-particle is ever written back.  S shells of the same object share that pass
// (each shell has its own frame in shared memory and its own 7 fp64 + 1 int accumulators in registers),
// so the particle streams are read once per S halo-shell iterations.
//
// Arithmetic: pass 0 (identity frame, sphere / spherical shell) is evaluated with the reference's exact
// operand order (__dmul_rn/__dadd_rn, no contraction) -- its particle counts are bit-exact by
// construction.  Later passes rotate with eigenvectors that cannot be bit-identical to LAPACK's, so
// they use FMAs; q, s and the axes agree to ~1e-14 and the integer counts are verified equal in tests.
#pragma once
#include <cooperative_groups.h>

#include "cpg_common.cuh"

namespace cpg {
namespace cg = cooperative_groups;

#ifndef CPG_FP64_PAIRS
#define CPG_FP64_PAIRS 0   // 1: the all-fp64 path works on one particle pair at a time (rolled loop)
#endif
#ifndef CPG_F32_TEST
// 0 (default): all-fp64 membership tests.  1: fp32 pre-classification of the membership tests of the non-reduced
// variants with an exact fp64 re-test of the undecided particles (pair_f32 below).  Measured on configs[1], B200
// (profiles/README.md, round 2): the kernels are bound by the dependency chain of each of the 12 resident warps per SM,
// not by the fp64 pipe (48 % of what it sustains); the fp32 variant halves the fp64 instructions of the streaming loop
// but lengthens that chain with compare / select / convert instructions: 24.3 ms against 19.4 ms for the local-shape kernels.
// Kept as a build option (make EXTRA=-DCPG_F32_TEST=1); the GPU suite passes with either setting.
#define CPG_F32_TEST 0
#endif

// Frame slots (doubles) per shell in shared memory.  The ellipsoidal radius of the reference,
//   v = p0^2 + p1^2/q^2 + p2^2/s^2  with  p = R (x - c)            (shape_profs_algos.pyx:247-256)
// is the quadratic form v = (x-c)^T A (x-c), A = R^T diag(1, 1/q^2, 1/s^2) R.  The six products
// xx, yy, zz, xy, xz, yz of a particle are formed ONCE, shared by all S shells of the pass, and reused by
// the tensor accumulation (S_ab += w * ab): 6 + S*(6 + 1 + 7) fp64 ops per particle instead of S*25.
enum {
    F_AXX = 0, F_AYY, F_AZZ, F_AXY, F_AXZ, F_AYZ,   // A (off-diagonals stored doubled)
    F_D2,                                           // d^2
    F_THR,                                          // inner-ellipsoid threshold: 1 (shell), -1 (first shell: disabled)
    F_BXX, F_BYY, F_BZZ, F_BXY, F_BXZ, F_BYZ,       // B = R^T diag(1/(d-delta)^2, 1/(q d-delta)^2, 1/(s d-delta)^2) R
    F_QQ,                                           // q*q, s*s: the divisors of the reference's own expression
    F_SS,                                           //   p0^2 + p1^2/q^2 + p2^2/s^2 (exact_member)
    F_V,                                            // 9 slots: eigenvector basis of the previous iteration
    F_BA2 = F_V + 9, F_BB2, F_BC2,                  // (d-delta)^2, (q d-delta)^2, (s d-delta)^2 (shell variant, exact_member)
    F_F32,                                          // 8 slots = 16 floats: the fp32 pre-classification frame (FrameF32)
    // (even: rows stay 16-byte aligned for the double2 / float4 loads; the all-fp64 build stops after the basis)
    F_SLOTS = CPG_F32_TEST ? F_F32 + 8 : F_V + 10
};
// fp32 pre-classification frame, 16 floats at F_F32: A (as F_AXX.., rounded to fp32), lo, hi, B, loB, hiB.
// See classify_f32 for what lo/hi mean.
enum { G_A = 0, G_LO = 6, G_HI = 7, G_B = 8, G_LOB = 14, G_HIB = 15 };
static_assert(F_V == 16 && F_F32 % 2 == 0 && F_SLOTS % 2 == 0 && F_SLOTS >= F_V + 9, "frame layout");

constexpr int TILE_PAIRS = 64;   // particle pairs per tile of a stream (1 KB)

struct Particle {              // one particle as seen by the accumulators
    double xx, yy, zz, xy, xz, yz;   // products of the centre-relative position
    double m;
    double t[6];               // products the tensor is built from (velocities; VDISP only)
};



template <bool REDUCED, bool VDISP>
__device__ __forceinline__ void accumulate(double (&acc)[7], int &cnt, bool sel, double v, const Particle &p)
{
    // weight: m (non-reduced) or m / r_ell^2 (reduced, helper_class.pyx:54); the common 1/mass_tot factor
    // is applied once per iteration after the reduction (it only rescales the tensor).  Written as a guarded
    // block so that ptxas predicates the seven fp64 updates instead of selecting a zero weight: a particle
    // outside the ellipsoid costs an issue slot and (measured, profiles/r2_loop_lab_v0.txt) most of its fp64 pipe time.
#ifdef CPG_ACC_SELECT
    {
        double w = REDUCED ? p.m / v : p.m;
        w = sel ? w : 0.0;
        const double mm = sel ? p.m : 0.0;
        if (VDISP) {
#pragma unroll
            for (int c = 0; c < 6; c++) acc[c] = fma(w, p.t[c], acc[c]);
        } else {
            acc[0] = fma(w, p.xx, acc[0]);
            acc[1] = fma(w, p.xy, acc[1]);
            acc[2] = fma(w, p.xz, acc[2]);
            acc[3] = fma(w, p.yy, acc[3]);
            acc[4] = fma(w, p.yz, acc[4]);
            acc[5] = fma(w, p.zz, acc[5]);
        }
        (void)mm;
        cnt += sel ? 1 : 0;
    }
    return;
#endif
    if (sel) {
        const double w = REDUCED ? p.m / v : p.m;
        if (VDISP) {
#pragma unroll
            for (int c = 0; c < 6; c++) acc[c] = fma(w, p.t[c], acc[c]);
        } else {
            acc[0] = fma(w, p.xx, acc[0]);
            acc[1] = fma(w, p.xy, acc[1]);
            acc[2] = fma(w, p.xz, acc[2]);
            acc[3] = fma(w, p.yy, acc[3]);
            acc[4] = fma(w, p.yz, acc[4]);
            acc[5] = fma(w, p.zz, acc[5]);
        }
        cnt += 1;   // (no mass sum: the reference's 1/mass_tot only rescales the tensor, helper_class.pyx:52)
    }
}

// Pass 0 (spherical start, shape_profs_algos.pyx:206-212 / :81-87) runs through the same code with the identity
// frame (q = s = 1) and thr = (d-delta)^2 -- no separate code path.
//
// Membership of a particle in the ellipsoid (shell) of the previous eigen-decomposition
// (shape_profs_algos.pyx:247-260 / :122-143), fused with the next tensor accumulation.
//
// fp32 pre-classification (non-reduced variants).  fp64 instructions are the bulk of the streaming loop and the membership
// test was 6 of the ~14 fp64 instructions per particle-shell.  Here the quadratic form v = x^T A x is evaluated
// in fp32 (x, A rounded to fp32, 6 FFMA) and compared with TWO thresholds:
//      v32 < lo  -> member,      v32 > hi -> not a member,      otherwise -> undecided,
//   lo = d^2 (1 - K) rounded down, hi = d^2 (1 + K) rounded up,  K = 3 * 2^-20 / s^2.
// Error bound: x32 = x(1+e), |e| <= u = 2^-24; each product carries <= 3u, A32 <= u, the FFMA chain <= 6u, so
// |v32 - x^T A x| <= 10.1 u * sum_ij |A_ij||x_i||x_j| <= 10.1 u * sqrt(3) * lambda_max(A) * |x|^2, with
// lambda_max(A) = 1/s^2 (s <= q <= 1).  Every particle the pass tests lies in the shell's radial prefix,
// |x|^2 < d^2 (1 + 1e-12), or beyond it in the prefix's last tile, where |x|^2 >= d^2 and v >= |x|^2 (A >= I): in both
// cases a "member" / "not a member" verdict of the fp32 test is the verdict of ANY fp64 evaluation of the
// reference's expression (their differences are ~1e-15 |x|^2 / s^2, seven orders below the band), so the fp32 path
// never decides a case that rounding could decide differently.  The undecided particles (a fraction ~K of them) are
// re-tested by exact_member() with the reference's own expression and operand order.  Non-finite fp32 values
// (the 1e150 pad sentinel, NaN input) fail both comparisons of the "undecided" test and the "member" test: not a
// member, as in fp64.  K >= 1/4 (s < 3.4e-3) or d^2 outside [1e-30, 1e30] switches the fp32 test off for the shell
// (lo = -inf: everything is undecided).  The inner ellipsoid of the shell variant is handled the same way with
// B, thr and an absolute band 3 * 2^-20 * lambda_max(B) * d^2.
template <bool SHELL>
__device__ __noinline__ bool exact_member(const double *F, double x, double y, double z)
{
    // shape_profs_algos.pyx:247-256 (and :139 for the shell): rotate into the eigenframe, then
    // p0^2 + p1^2/q^2 + p2^2/s^2 < d^2, left to right, true divisions, no contraction.  With the identity frame of
    // pass 0 this is bit for bit (x^2 + y^2) + z^2 of :206 / :82.
    const double *V = F + F_V;
    const double p0 = __dadd_rn(__dadd_rn(__dmul_rn(V[6], x), __dmul_rn(V[7], y)), __dmul_rn(V[8], z));
    const double p1 = __dadd_rn(__dadd_rn(__dmul_rn(V[3], x), __dmul_rn(V[4], y)), __dmul_rn(V[5], z));
    const double p2 = __dadd_rn(__dadd_rn(__dmul_rn(V[0], x), __dmul_rn(V[1], y)), __dmul_rn(V[2], z));
    const double s0 = __dmul_rn(p0, p0), s1 = __dmul_rn(p1, p1), s2 = __dmul_rn(p2, p2);
    const double v = __dadd_rn(__dadd_rn(s0, __ddiv_rn(s1, F[F_QQ])), __ddiv_rn(s2, F[F_SS]));
    bool sel = v < F[F_D2];
    if (SHELL && sel && F[F_THR] >= 0.0) {   // thr < 0: first shell, no inner ellipsoid (:130)
        const double u = __dadd_rn(__dadd_rn(__ddiv_rn(s0, F[F_BA2]), __ddiv_rn(s1, F[F_BB2])), __ddiv_rn(s2, F[F_BC2]));
        sel = u >= F[F_THR];
    }
    return sel;
}

constexpr double F32_GAMMA3 = 3.0 / 1048576.0;   // 3 * 2^-20 (see above: 10.1 u sqrt(3) = 1.05e-6 < 2.86e-6)

// The fp32 frame of a shell from its fp64 frame.  amax = lambda_max(A) = 1/s^2, bmax = lambda_max(B).
__device__ __forceinline__ void write_f32_frame(double *F, double amax, double bmax, bool shell)
{
    float *G = reinterpret_cast<float *>(F + F_F32);
    const double d2 = F[F_D2];
    const double K = F32_GAMMA3 * amax;
    const bool ok = (K < 0.25) && (d2 > 1e-30) && (d2 < 1e30);
#pragma unroll
    for (int r = 0; r < 6; r++) G[G_A + r] = (float)F[F_AXX + r];
    G[G_LO] = ok ? __double2float_rd(d2 * (1.0 - K)) : -INFINITY;
    G[G_HI] = ok ? __double2float_ru(d2 * (1.0 + K)) : INFINITY;
    if (shell) {
        const double thr = F[F_THR];
        const double KB = F32_GAMMA3 * bmax * d2 * 1.000001;   // absolute: |x|^2 < d^2 (1 + 1e-12) for every member of the outer ellipsoid
        const bool okb = ok && (KB < 1e30) && (fabs(thr) < 1e30);
#pragma unroll
        for (int r = 0; r < 6; r++) G[G_B + r] = (float)F[F_BXX + r];
        G[G_LOB] = okb ? __double2float_rd(thr - KB) : -INFINITY;
        G[G_HIB] = okb ? __double2float_ru(thr + KB) : INFINITY;
    }
}

// The quadratic form x^T A x of NP particles from their six products, A = (axx, ayy, azz, axy, axz, ayz; off-diagonals
// doubled): one 6-term multiply-add chain per particle (default), or -- build option CPG_SPLIT_FORM=1 -- two 3-term chains
// (squares, then mixed products) joined by one addition, written stage by stage across the particles.  The split form
// halves the depth of the dependency chain; with the identity frame of pass 0 its first chain is the reference's
// (x^2 + y^2) + z^2 bit for bit and the second is exactly 0.
#ifndef CPG_SPLIT_FORM
#define CPG_SPLIT_FORM 0   // measured (profiles/README.md, round 2 second part): no gain in the kernels, 16 more fp64 instructions per tile
#endif
template <int NP>
__device__ __forceinline__ void quad_forms(double axx, double ayy, double azz, double axy, double axz, double ayz,
                                           const Particle (&p)[NP], double (&v)[NP])
{
#if CPG_SPLIT_FORM
    double a[NP], b[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) { a[j] = axx * p[j].xx; b[j] = axy * p[j].xy; }
#pragma unroll
    for (int j = 0; j < NP; j++) { a[j] = fma(ayy, p[j].yy, a[j]); b[j] = fma(axz, p[j].xz, b[j]); }
#pragma unroll
    for (int j = 0; j < NP; j++) { a[j] = fma(azz, p[j].zz, a[j]); b[j] = fma(ayz, p[j].yz, b[j]); }
#pragma unroll
    for (int j = 0; j < NP; j++) v[j] = a[j] + b[j];
#else
#pragma unroll
    for (int j = 0; j < NP; j++)
        v[j] = fma(ayz, p[j].yz, fma(axz, p[j].xz, fma(axy, p[j].xy, fma(azz, p[j].zz, fma(ayy, p[j].yy, axx * p[j].xx)))));
#endif
}

// All-fp64 membership + accumulation (the reduced variants, whose weight m / r_ell^2 needs the fp64 value anyway).
template <int S, bool SHELL, bool REDUCED, bool VDISP, int NP>
__device__ __forceinline__ void passk_particles(const Particle (&p)[NP], const double *frames, unsigned smask,
                                                double (&acc)[S][7], int (&cnt)[S])
{
#pragma unroll
    for (int s = 0; s < S; s++) {
        // smask: shells that are still iterating, whose radial prefix reaches this tile and whose inner prefix
        // (served by the prefix tensor table) ends before it (warp-uniform)
        if (smask & (1u << s)) {
            const double2 *F2 = reinterpret_cast<const double2 *>(frames + s * F_SLOTS);
            const double2 f0 = F2[0], f1 = F2[1], f2 = F2[2], f3 = F2[3];
            // f0=(Axx,Ayy) f1=(Azz,Axy) f2=(Axz,Ayz) f3=(d2,thr)
            double2 f4, f5, f6;
            if (SHELL) { f4 = F2[4]; f5 = F2[5]; f6 = F2[6]; }   // (Bxx,Byy) (Bzz,Bxy) (Bxz,Byz)
            double v[NP];
            bool sel[NP];
            quad_forms<NP>(f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, p, v);
            if (SHELL) {
                double u[NP];
                quad_forms<NP>(f4.x, f4.y, f5.x, f5.y, f6.x, f6.y, p, u);
#pragma unroll
                for (int j = 0; j < NP; j++) sel[j] = (v[j] < f3.x) && (u[j] >= f3.y);
            } else {
#pragma unroll
                for (int j = 0; j < NP; j++) sel[j] = v[j] < f3.x;
            }
#pragma unroll
            for (int j = 0; j < NP; j++) accumulate<REDUCED, VDISP>(acc[s], cnt[s], sel[j], v[j], p[j]);
        }
    }
}

// The same for the non-reduced ellipsoid variant, two shells at a time (build option CPG_SHELL_PAIRS=1): the frames of
// two shells are fetched together, their 2 NP forms run as independent chains, the 2 NP compares issue back to back and
// there is one uniform branch per pair; a shell of the pair that is not active on this tile takes no particle
// (predicate in the compare).  Shortens the static dependency chain of a 4-shell tile from 1322 to 821 cycles, but a
// pair with one active shell costs as much as two: measured (profiles/README.md) 1.13 instead of 1.19 us per tile in the
// warp kernel, 1.32 instead of 1.24 in the cluster kernel, whose tiles mostly carry all four shells -- not the default.
#ifndef CPG_SHELL_PAIRS
#define CPG_SHELL_PAIRS 0
#endif
template <int S, int NP>
__device__ __forceinline__ void passk_shell_pairs(const Particle (&p)[NP], const double *frames, unsigned smask,
                                                  double (&acc)[S][7], int (&cnt)[S])
{
    static_assert(S % 2 == 0, "shell pairs");
#pragma unroll
    for (int h = 0; h < S; h += 2) {
        const unsigned two = (smask >> h) & 3u;
        if (two) {   // warp-uniform
            const double2 *Fa = reinterpret_cast<const double2 *>(frames + h * F_SLOTS);
            const double2 *Fb = reinterpret_cast<const double2 *>(frames + (h + 1) * F_SLOTS);
            const double2 a0 = Fa[0], a1 = Fa[1], a2 = Fa[2], b0 = Fb[0], b1 = Fb[1], b2 = Fb[2];
            const double da = Fa[3].x, db = Fb[3].x;
            const bool ona = (two & 1u) != 0, onb = (two & 2u) != 0;
            double va[NP], vb[NP];
            quad_forms<NP>(a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, p, va);
            quad_forms<NP>(b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, p, vb);
            bool sa[NP], sb[NP];
#pragma unroll
            for (int j = 0; j < NP; j++) { sa[j] = (va[j] < da) && ona; sb[j] = (vb[j] < db) && onb; }
#pragma unroll
            for (int j = 0; j < NP; j++) {
                accumulate<false, false>(acc[h], cnt[h], sa[j], va[j], p[j]);
                accumulate<false, false>(acc[h + 1], cnt[h + 1], sb[j], vb[j], p[j]);
            }
        }
    }
}

// The undecided particles of the fp32 pre-classification, re-tested in fp64: with the quadratic form (what
// passk_particles evaluates; exact squares, so that pass 0 is the reference's ((xx+yy)+zz)), or -- for needle-like
// shells, K >= F32_NEEDLE_K, where the quadratic form itself loses digits to cancellation, and for shells whose fp32
// test is switched off -- with the reference's own rotate-then-sum expression (exact_member).  Out of line: it runs
// for a fraction ~K of the particles.  v0, v1 (u0, u1): the fp32 forms of the pair's two particles; m2: their fp32
// verdicts (bit e = particle e is a member); returns the corrected verdicts.
constexpr float F32_NEEDLE_K = 1e-3f;   // K = 3 * 2^-20 / s^2 >= 1e-3  <=>  s <= 0.053

struct PairForms { float v0, v1, u0, u1; };

// Returns the forms with every UNDECIDED particle's pair (v, u) replaced by a verdict the caller's own comparisons
// read back: (v, u) = (-1, +inf) for a member (v < lo_eff and u > hiB_eff always hold: lo_eff >= 0, hiB_eff <= FLT_MAX),
// (+inf, u) for a non-member.  Decided particles keep their values.
template <bool SHELL>
__device__ __noinline__ PairForms redecide_pair(const double *F, const unsigned char *tile, int pi, PairForms w)
{
    const float4 *G4 = reinterpret_cast<const float4 *>(F + F_F32);
    const float lo = G4[1].z, hi = G4[1].w;
    const bool off = !(lo > -INFINITY);
    float loB = 0.f, hiB = 0.f;
    bool offB = false;
    if (SHELL) { loB = G4[3].z; hiB = G4[3].w; offB = !(loB > -INFINITY); }
    const bool needle = off || (hi - lo > 2.0f * F32_NEEDLE_K * hi);   // hi - lo = 2 K d^2
    const double2 *T = reinterpret_cast<const double2 *>(tile);
    const double2 x = T[pi], y = T[TILE_PAIRS + pi], z = T[2 * TILE_PAIRS + pi];
#pragma unroll 1
    for (int e = 0; e < 2; e++) {
        const float vj = e ? w.v1 : w.v0;
        bool un = off || ((vj >= lo) && (vj <= hi));
        if (SHELL) {
            const float uj = e ? w.u1 : w.u0;
            un = un || ((vj < lo) && (((uj >= loB) && (uj <= hiB)) || offB));
        }
        if (un) {
            const double px = e ? x.y : x.x, py = e ? y.y : y.x, pz = e ? z.y : z.x;
            bool mem;
            if (needle) {
                mem = exact_member<SHELL>(F, px, py, pz);
            } else {
                const double xx = __dmul_rn(px, px), yy = __dmul_rn(py, py), zz = __dmul_rn(pz, pz);
                const double xy = px * py, xz = px * pz, yz = py * pz;
                const double v = fma(F[F_AYZ], yz, fma(F[F_AXZ], xz, fma(F[F_AXY], xy, fma(F[F_AZZ], zz, fma(F[F_AYY], yy, F[F_AXX] * xx)))));
                mem = v < F[F_D2];
                if (SHELL) {
                    const double u = fma(F[F_BYZ], yz, fma(F[F_BXZ], xz, fma(F[F_BXY], xy, fma(F[F_BZZ], zz, fma(F[F_BYY], yy, F[F_BXX] * xx)))));
                    mem = mem && (u >= F[F_THR]);
                }
            }
            const float nv = mem ? -1.0f : INFINITY;
            if (e) { w.v1 = nv; w.u1 = INFINITY; } else { w.v0 = nv; w.u0 = INFINITY; }
        }
    }
    return w;
}

// One particle pair (pair `pi` of the tile: 2 particles per lane) of the non-reduced variants: per shell the
// membership of the two particles is decided in fp32 (classification comment above) and the tensor updates are
// predicated fp64 FMAs on the six products formed once per particle.  Undecided particles (rare; warp vote) go through
// redecide_pair.  Called for pairs lane and lane + 32 from a loop that is NOT unrolled: half the registers and half
// the code of a 4-particle body (the kernels are bound by issue slots and the instruction cache, not by ILP: the
// other resident warps fill the latency of the two fp32 chains).  Branch-free by construction (bitwise bool
// operators: no short-circuit evaluation, which the compiler would turn into divergent branches).
template <int S, bool SHELL, bool VDISP>
__device__ __forceinline__ void pair_f32(const unsigned char *tile, int pi, const double *frames, unsigned smask,
                                         double (&acc)[S][7], int (&cnt)[S])
{
    const double2 *T = reinterpret_cast<const double2 *>(tile);   // stream st, pair pi: T[st * TILE_PAIRS + pi]
    float f[2][6];     // fp32 products of the position: xx, yy, zz, xy, xz, yz
    double t[2][6];    // fp64 products the tensor is built from (positions, or velocities): xx, xy, xz, yy, yz, zz
    double pm[2];
    {
        const double2 x = T[pi], y = T[TILE_PAIRS + pi], z = T[2 * TILE_PAIRS + pi], m = T[3 * TILE_PAIRS + pi];
        double2 a = x, b = y, c = z;
        if (VDISP) { a = T[4 * TILE_PAIRS + pi]; b = T[5 * TILE_PAIRS + pi]; c = T[6 * TILE_PAIRS + pi]; }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            // round to nearest (the error bound assumes it); 1e150 (pad sentinel) becomes inf
            const float fx = (float)(e ? x.y : x.x), fy = (float)(e ? y.y : y.x), fz = (float)(e ? z.y : z.x);
            f[e][0] = fx * fx; f[e][1] = fy * fy; f[e][2] = fz * fz; f[e][3] = fx * fy; f[e][4] = fx * fz; f[e][5] = fy * fz;
            const double ta = e ? a.y : a.x, tb = e ? b.y : b.x, tc = e ? c.y : c.x;
            t[e][0] = ta * ta; t[e][1] = ta * tb; t[e][2] = ta * tc; t[e][3] = tb * tb; t[e][4] = tb * tc; t[e][5] = tc * tc;
            pm[e] = e ? m.y : m.x;
        }
    }
#pragma unroll
    for (int s = 0; s < S; s++) {
        if (smask & (1u << s)) {   // warp-uniform
            const float4 *G4 = reinterpret_cast<const float4 *>(frames + s * F_SLOTS + F_F32);
            const float4 g0 = G4[0], g1 = G4[1];   // (Axx,Ayy,Azz,Axy) (Axz,Ayz,lo,hi)
            PairForms w;
            w.v0 = fmaf(g1.y, f[0][5], fmaf(g1.x, f[0][4], fmaf(g0.w, f[0][3], fmaf(g0.z, f[0][2], fmaf(g0.y, f[0][1], g0.x * f[0][0])))));
            w.v1 = fmaf(g1.y, f[1][5], fmaf(g1.x, f[1][4], fmaf(g0.w, f[1][3], fmaf(g0.z, f[1][2], fmaf(g0.y, f[1][1], g0.x * f[1][0])))));
            w.u0 = 0.f; w.u1 = 0.f;
            const bool off = !(g1.z > -INFINITY);    // fp32 test switched off for this shell: everything undecided
            const float lo_eff = off ? 0.0f : g1.z;
            bool und = off | ((w.v0 >= g1.z) & (w.v0 <= g1.w)) | ((w.v1 >= g1.z) & (w.v1 <= g1.w));
            float hiB_eff = 0.f;
            if (SHELL) {
                const float4 g2 = G4[2], g3 = G4[3];   // (Bxx,Byy,Bzz,Bxy) (Bxz,Byz,loB,hiB)
                w.u0 = fmaf(g3.y, f[0][5], fmaf(g3.x, f[0][4], fmaf(g2.w, f[0][3], fmaf(g2.z, f[0][2], fmaf(g2.y, f[0][1], g2.x * f[0][0])))));
                w.u1 = fmaf(g3.y, f[1][5], fmaf(g3.x, f[1][4], fmaf(g2.w, f[1][3], fmaf(g2.z, f[1][2], fmaf(g2.y, f[1][1], g2.x * f[1][0])))));
                const bool offB = !(g3.z > -INFINITY);
                hiB_eff = offB ? 3.0e38f : g3.w;
                und = und | ((w.v0 < g1.z) & (offB | ((w.u0 >= g3.z) & (w.u0 <= g3.w))))
                          | ((w.v1 < g1.z) & (offB | ((w.u1 >= g3.z) & (w.u1 <= g3.w))));
            }
            if (__any_sync(0xffffffffu, und))   // rare: fp64 decides the undecided ones
                w = redecide_pair<SHELL>(frames + s * F_SLOTS, tile, pi, w);
            bool in0 = w.v0 < lo_eff, in1 = w.v1 < lo_eff;
            if (SHELL) { in0 = in0 & (w.u0 > hiB_eff); in1 = in1 & (w.u1 > hiB_eff); }
            if (in0) {   // predicated updates (no mass sum: only ratios of the tensor are used)
#pragma unroll
                for (int c = 0; c < 6; c++) acc[s][c] = fma(pm[0], t[0][c], acc[s][c]);
                cnt[s] += 1;
            }
            if (in1) {
#pragma unroll
                for (int c = 0; c < 6; c++) acc[s][c] = fma(pm[1], t[1][c], acc[s][c]);
                cnt[s] += 1;
            }
        }
    }
}

// A = w2 * e2 e2^T + w1 * e1 e1^T + w0 * e0 e0^T as (xx, yy, zz, 2xy, 2xz, 2yz); v = eigenvectors of eigh3.
__device__ __forceinline__ void quad_form(double *A, const double (&v)[9], double w2, double w1, double w0)
{
    const double *e0 = v, *e1 = v + 3, *e2 = v + 6;
    A[0] = w2 * e2[0] * e2[0] + w1 * e1[0] * e1[0] + w0 * e0[0] * e0[0];
    A[1] = w2 * e2[1] * e2[1] + w1 * e1[1] * e1[1] + w0 * e0[1] * e0[1];
    A[2] = w2 * e2[2] * e2[2] + w1 * e1[2] * e1[2] + w0 * e0[2] * e0[2];
    A[3] = 2.0 * (w2 * e2[0] * e2[1] + w1 * e1[0] * e1[1] + w0 * e0[0] * e0[1]);
    A[4] = 2.0 * (w2 * e2[0] * e2[2] + w1 * e1[0] * e1[2] + w0 * e0[0] * e0[2]);
    A[5] = 2.0 * (w2 * e2[1] * e2[2] + w1 * e1[1] * e1[2] + w0 * e0[1] * e0[2]);
}

// Inner radial prefix of a shell: the particles of the buckets j with rr[j] <= s * rr[k] * (1 - 1e-9) lie at
// |x-c| < s*d_k; with q, s <= 1 their ellipsoidal radius obeys r_ell^2 <= |x-c|^2 / s^2 < d_k^2, so they are
// members of shell k's ellipsoid whatever its orientation.  The whole 128-particle tiles of that prefix are never
// streamed: their tensor sums are one row of the prefix tensor table (cpg_gather.cuh), added after the reduction.  rr: the R shell radii (shared
// memory copy), neff: this object's prefix lengths (shared memory copy); both null when there is no ordering.
struct InnerPrefix {
    const double *rr;
    const int *neff;
    // compact copies for the warp kernel's shared memory: radii rounded UP to fp32 (a bucket that passes the test with
    // the rounded-up radius passes it with the true one) and prefix lengths in pairs (objects < 131072 particles)
    const float *rr_up = nullptr;
    const uint16_t *npairs16 = nullptr;
    // ... or, with rr_up set and npairs16 null, the object's row of the n_eff table in global memory (kernel 1c: one
    // read per step instead of a shared-memory copy per task slot)
    const int32_t *neff_g = nullptr;
    __device__ __forceinline__ int pairs_inside(int k, double s) const
    {
        if (rr_up) {
            // rr_up[k] <= rr[k] (1 + 6e-8): still below s * rr[k] (1 - 1e-9); rounded DOWN to fp32 so that the search
            // compares floats (a bucket that passes with the rounded-down limit passes with the true one)
            const float lim = __double2float_rd(s * (double)rr_up[k] * (1.0 - 1e-7));
            int lo = -1, hi = k;   // rr_up[lo] <= lim < rr_up[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (rr_up[mid] <= lim) lo = mid; else hi = mid;
            }
            return lo >= 0 ? (npairs16 ? (int)npairs16[lo] : (int)(neff_g[lo] >> 1)) : 0;
        }
        if (!rr) return 0;
        const double lim = s * rr[k] * (1.0 - 1e-9);
        int lo = -1, hi = k;   // rr[lo] <= lim < rr[hi]   (rr ascending, rr[k] > lim)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (rr[mid] <= lim) lo = mid; else hi = mid;
        }
        return lo >= 0 ? neff[lo] >> 1 : 0;   // whole pairs only
    }
    // Shell variant (shape_profs_algos.pyx:130-143): a particle belongs to the homoeoid between the ellipsoid of radius d
    // and the inner one with semi-axes (d - delta, q d - delta, s d - delta).  With s <= q <= 1 and c = s d - delta > 0
    // the inner ellipsoid contains the ball of radius c: the reference's p0^2/a^2 + p1^2/b^2 + p2^2/c^2 is then at most
    // |x|^2 / c^2 < 1, so a particle at |x - centre| < c (1 - 1e-6) is NOT a member whatever the orientation, and no
    // evaluation of the expression can say otherwise.  The radially ordered buckets j with rr[j] (1 + 1e-6) <= c / r200
    // hold only such particles: their tiles are skipped (nothing to add: they carry no weight), exactly like the tiles
    // past the outer prefix.  c / r200 = (s - delta / d) rr[k].  Returns the pairs in those buckets (0: no skipping).
    __device__ __forceinline__ int pairs_below_shell(int k, double s, double d, double delta) const
    {
        if ((!rr_up && !rr) || k <= 0 || !(d > delta)) return 0;   // first shell (d == delta): no inner ellipsoid
        const double frac = s - delta / d;                         // c / d
        if (!(frac > 1e-3)) return 0;                              // c <= 0 (or tiny): the inner ball is no bound
        // rr_up[k] (1 - 2e-7) <= rr[k]: the radius of shell k from below
        const double lim = frac * (rr_up ? (double)rr_up[k] * (1.0 - 2e-7) : rr[k]) * (1.0 - 1e-6);
        const float limf = __double2float_rd(lim);   // (fp32 search over rr_up: conservative, see pairs_inside)
        int lo = -1, hi = k;   // radius[lo] <= lim < radius[hi]   (ascending; radius[k] > lim because frac < 1)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (rr_up ? (rr_up[mid] <= limf) : (rr[mid] <= lim)) lo = mid; else hi = mid;
        }
        if (lo < 0) return 0;
        return rr_up ? (npairs16 ? (int)npairs16[lo] : (int)(neff_g[lo] >> 1)) : neff[lo] >> 1;
    }
};

// Per halo-shell iteration state, replicated in every lane that serves the shell.
struct KatzState {
    double q, s;        // current axis ratios
    double d, delta;    // ellipsoidal radius, shell width (delta == d: first shell / ellipsoid mode)
    int iteration;      // the reference's `iteration` (starts at 1)
    bool active;
};

// One trip through the body of the reference's while-loop for one shell, given the reduced sums of the
// pass that just finished.  tot = {Sxx,Sxy,Sxz,Syy,Syz,Szz,M}.  Writes the next frame into `frame`
// (shared, 16 doubles) when the iteration goes on and the 12 outputs when it ends.  `writer` lanes do
// the global/shared stores; all lanes of a shell compute identical values.
template <bool SHELL>
__device__ __forceinline__ void katz_step(KatzState &st, const double (&tot)[7], int pts, const MorphParams &P,
                                          const double (&vprev)[9], double *frame, bool frame_writer, size_t out_idx,
                                          bool out_writer, const InnerPrefix &ip, int k, int *nin_slot)
{
    if (!st.active)
        return;
    if (st.iteration > P.wall || pts < P.itmin) {   // shape_profs_algos.pyx:214-219: zeros (already there)
        if (out_writer) {
            P.n_iter[out_idx] = st.iteration - 1;
            P.n_pts[out_idx] = pts;
        }
        st.active = false;
        return;
    }
    const double im = fast_rcp(tot[0] + tot[3] + tot[5]);   // any positive scale does: only ratios are consumed
    double w[3], v[9];
#pragma unroll
    for (int r = 0; r < 9; r++) v[r] = vprev[r];   // warm start: basis of the previous iteration
    eigh3(tot[0] * im, tot[1] * im, tot[2] * im, tot[3] * im, tot[4] * im, tot[5] * im, w, v);
#ifdef CPG_STEP_TWICE
    {   // measurement only: the eigen-solve a second time on (opaquely) the same input; results unchanged
        double w2[3], v2[9];
#pragma unroll
        for (int r = 0; r < 9; r++) v2[r] = vprev[r];
        const double one = 1.0 + 0.0 * P.tol;
        eigh3(tot[0] * im * one, tot[1] * im, tot[2] * im, tot[3] * im, tot[4] * im, tot[5] * im, w2, v2);
        w[0] = fma(0.0, w2[0] + v2[0] + v2[4] + v2[8], w[0]);
    }
#endif
    if (!(w[2] > 0.0)) {
        // all-zero or NaN tensor (coincident points, a particle exactly at the centre in reduced mode): the
        // reference divides by zero here and raises; report the shell as failed instead
        if (out_writer) {
            P.n_iter[out_idx] = st.iteration - 1;
            P.n_pts[out_idx] = pts;
        }
        st.active = false;
        return;
    }
    const double q_old = st.q, s_old = st.s;
    st.q = sqrt(w[1] / w[2]);
    st.s = sqrt(w[0] / w[2]);
    // helper_class.pyx:110-120: the difference is truncated to fp32 before abs()
    const double eq = (double)fabsf((float)(st.q - q_old)) / q_old;
    const double es = (double)fabsf((float)(st.s - s_old)) / s_old;
    const double err = es > eq ? es : eq;   // Python max(eq, es)
    if (!(err > P.tol)) {                   // converged (or NaN): these are the returned values
        if (out_writer) {
            double *o = P.out12 + 12 * out_idx;
            o[0] = st.d; o[1] = st.q; o[2] = st.s;
#pragma unroll
            for (int r = 0; r < 3; r++) { o[3 + r] = v[6 + r]; o[6 + r] = v[3 + r]; o[9 + r] = v[r]; }
            P.n_iter[out_idx] = st.iteration;
            P.n_pts[out_idx] = pts;
        }
        st.active = false;
        return;
    }
    if (frame_writer) {
#pragma unroll
        for (int r = 0; r < 9; r++) frame[F_V + r] = v[r];
        // v[6..8] major (e2), v[3..5] intermediate (e1), v[0..2] minor (e0)
        const double qq = __dmul_rn(st.q, st.q), ss = __dmul_rn(st.s, st.s);   // q_new**2, s_new**2 (:251)
        const double iq2 = 1.0 / qq, is2 = 1.0 / ss;
        quad_form(frame + F_AXX, v, 1.0, iq2, is2);
        frame[F_QQ] = qq; frame[F_SS] = ss;   // (inside the frame in every build: slots 14, 15)
        double bmax = 0.0;
        if (SHELL) {
            const bool first = (st.d == st.delta);   // shape_profs_algos.pyx:130: first shell -> ellipsoid test only
            // (d-delta_d)**2, (q_new*d-delta_d)**2, (s_new*d-delta_d)**2 of :139
            const double a = __dsub_rn(st.d, st.delta), b = __dsub_rn(__dmul_rn(st.q, st.d), st.delta),
                         c = __dsub_rn(__dmul_rn(st.s, st.d), st.delta);
            const double a2 = __dmul_rn(a, a), b2 = __dmul_rn(b, b), c2 = __dmul_rn(c, c);
            if (first) {
#pragma unroll
                for (int r = 0; r < 6; r++) frame[F_BXX + r] = 0.0;
            } else {
                quad_form(frame + F_BXX, v, 1.0 / a2, 1.0 / b2, 1.0 / c2);
                bmax = fmax(1.0 / a2, fmax(1.0 / b2, 1.0 / c2));
            }
            if (CPG_F32_TEST) { frame[F_BA2] = a2; frame[F_BB2] = b2; frame[F_BC2] = c2; }
            frame[F_THR] = first ? -1.0 : 1.0;
        }
        if (CPG_F32_TEST) write_f32_frame(frame, is2, bmax, SHELL);
        *nin_slot = SHELL ? ip.pairs_below_shell(k, st.s, st.d, st.delta) : ip.pairs_inside(k, st.s);
    }
    st.iteration += 1;
}

// Shell geometry (shape_profs_algos.pyx:1035-1044 local, :1096 global), exact reference arithmetic.
__device__ __forceinline__ void shell_geometry(const MorphParams &P, int halo, int k, double &d, double &delta)
{
    const double r200 = P.r200[halo];
    if (P.local) {
        d = __dmul_rn(r200, P.rr[k]);
        delta = (k == 0) ? d : __dsub_rn(d, __dmul_rn(r200, P.rr[k - 1]));
    } else {
        d = __dadd_rn(r200, P.SAFE);
        delta = d;
    }
}

__device__ __forceinline__ void make_particle(double x, double y, double z, double m, Particle &p)
{
    p.xx = __dmul_rn(x, x); p.yy = __dmul_rn(y, y); p.zz = __dmul_rn(z, z);   // exact squares: pass 0 needs them
    p.xy = x * y; p.xz = x * z; p.yz = y * z;
    p.m = m;
}

__device__ __forceinline__ void make_tensor_products(double u, double v, double w, Particle &p)
{
    // component order of accumulate(): xx, xy, xz, yy, yz, zz
    p.t[0] = u * u; p.t[1] = u * v; p.t[2] = u * w; p.t[3] = v * v; p.t[4] = v * w; p.t[5] = w * w;
}

// ------------------------------------------------------------------------------------------------
// Particle streaming: TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier) into a
// per-warp ring of RING_STAGES tiles.  A tile is TILE_PAIRS consecutive particle pairs of every stream
// (1 KB per stream); lane l consumes pairs l and l+32 of the tile with conflict-free 16-byte LDS.  One
// elected lane issues the copies RING_STAGES tiles ahead, so the warp never waits on global-memory latency
// with registers tied up, and an object whose radial prefix fits in the ring (<= 512 particles per warp)
// is loaded once and stays resident in shared memory for all Katz iterations of the task.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_SORT_R = 256;   // == SORT_MAX_B of cpg_gather.cuh: radial ordering exists for R <= 255
#ifndef CPG_RING_STAGES
#define CPG_RING_STAGES 4   // tiles of 4 KB per warp ring (positions); 3 leaves room for a fourth CTA of the warp kernel per SM
#endif
constexpr int RING_MAX_STAGES = 4;

template <bool VDISP, int STG = (VDISP ? 2 : CPG_RING_STAGES)>
struct RingCfg {
    static constexpr int NSTREAM = VDISP ? 7 : 4;
    static constexpr int STAGES = STG;   // default: 16 KB (positions) / 14 KB (with velocities) per warp
    static constexpr int STREAM_BYTES = TILE_PAIRS * 16;
    static constexpr int STAGE_BYTES = NSTREAM * STREAM_BYTES;
    static constexpr int WARP_BYTES = STAGES * STAGE_BYTES;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_s, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar_s, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar_s), "r"(parity) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst_s, const void *src, uint32_t bytes, uint32_t bar_s)
{
    // a CTA-local shared address is a valid shared::cluster address of the executing CTA
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
                 "l"(src), "r"(bytes), "r"(bar_s)
                 : "memory");
}

struct WarpRing {
    unsigned char *buf;   // this warp's tiles in shared memory
    uint64_t *bar;        // one mbarrier per stage
    uint32_t buf_s, bar_s;   // the same two as 32-bit shared-window addresses (computed once)
    uint32_t seq;         // tiles issued so far by this warp (stage = seq % STAGES, parity = (seq / STAGES) & 1)
    uint32_t res_seq;     // first sequence number of the resident tiles of the current task
    bool res_loaded;      // resident tiles are in place
    uint32_t st_tiles, st_tshells;   // statistics: tiles this warp passed over, (tile, shell) pairs among them
};

__device__ __forceinline__ void ring_flush_stats(WarpRing &r, unsigned long long *stats, int lane)
{
    if (stats && lane == 0 && r.st_tiles) {
        atomicAdd(stats, (unsigned long long)r.st_tiles);
        atomicAdd(stats + 1, (unsigned long long)r.st_tshells);
    }
    r.st_tiles = 0; r.st_tshells = 0;
}

// arrivals: 1 for the TMA ring (the elected lane's arrive.expect_tx), 32 for the per-lane cp.async ring (every lane's
// cp.async.mbarrier.arrive.noinc)
__device__ __forceinline__ void ring_init(WarpRing &r, unsigned char *buf, uint64_t *bar, int lane, uint32_t arrivals = 1)
{
    r.buf = buf; r.bar = bar; r.seq = 0; r.res_seq = 0; r.res_loaded = false; r.st_tiles = 0; r.st_tshells = 0;
    r.buf_s = smem_u32(buf); r.bar_s = smem_u32(bar);
    if (lane == 0) {
        for (int s = 0; s < RING_MAX_STAGES; s++) mbar_init(bar + s, arrivals);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
}

// Issue the bulk copies of tile `t` (pairs [64 t, 64 t + n)) of every stream into ring slot `seq`.
template <bool VDISP, int STG = RingCfg<VDISP>::STAGES>
__device__ __forceinline__ void ring_issue(const WarpRing &r, uint32_t seq, const SoA &soa, int64_t base, int t, int npairs)
{
    using C = RingCfg<VDISP, STG>;
    const uint32_t stage = seq % C::STAGES;
    const uint32_t bytes = (uint32_t)min(TILE_PAIRS, npairs - t * TILE_PAIRS) * 16u;
    const uint32_t dst = r.buf_s + stage * C::STAGE_BYTES;
    const uint32_t bar = r.bar_s + stage * 8u;
    const int64_t o = base + (int64_t)t * (2 * TILE_PAIRS);
    mbar_expect_tx(bar, bytes * C::NSTREAM);
    bulk_g2s(dst + 0 * C::STREAM_BYTES, soa.dx + o, bytes, bar);
    bulk_g2s(dst + 1 * C::STREAM_BYTES, soa.dy + o, bytes, bar);
    bulk_g2s(dst + 2 * C::STREAM_BYTES, soa.dz + o, bytes, bar);
    bulk_g2s(dst + 3 * C::STREAM_BYTES, soa.m + o, bytes, bar);
    if (VDISP) {
        bulk_g2s(dst + 4 * C::STREAM_BYTES, soa.vx + o, bytes, bar);
        bulk_g2s(dst + 5 * C::STREAM_BYTES, soa.vy + o, bytes, bar);
        bulk_g2s(dst + 6 * C::STREAM_BYTES, soa.vz + o, bytes, bar);
    }
}

// The same tile brought in by the whole warp with 16-byte cp.async (LDGSTS): lane l copies pairs l and l + 32 of every
// stream, then arrives on the stage's mbarrier once its copies have landed.  8 warp instructions per tile instead of
// the ~100 the elected-lane TMA issue costs (4 UBLKCP, each behind a uniform-register broadcast loop): used by the
// warp kernel, whose tasks average fewer than 8 tiles per pass (profiles/README.md, source-page breakdown).
template <bool VDISP, int STG = RingCfg<VDISP>::STAGES>
__device__ __forceinline__ void ring_issue_lanes(const WarpRing &r, uint32_t seq, const SoA &soa, int64_t base, int t, int npairs,
                                                 int lane)
{
    using C = RingCfg<VDISP, STG>;
    const uint32_t stage = seq % C::STAGES;
    const int n = min(TILE_PAIRS, npairs - t * TILE_PAIRS);
    const uint32_t dst = r.buf_s + stage * C::STAGE_BYTES;
    const uint32_t bar = r.bar_s + stage * 8u;
    const int64_t o = base + (int64_t)t * (2 * TILE_PAIRS);
    const double *src[7] = {soa.dx, soa.dy, soa.dz, soa.m, soa.vx, soa.vy, soa.vz};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int pi = lane + 32 * h;
        if (pi < n) {
#pragma unroll
            for (int st = 0; st < C::NSTREAM; st++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + st * C::STREAM_BYTES + pi * 16u),
                             "l"(src[st] + o + 2 * pi)
                             : "memory");
        }
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

template <bool VDISP, int STG = RingCfg<VDISP>::STAGES>
__device__ __forceinline__ const unsigned char *ring_tile(const WarpRing &r, uint32_t seq)
{
    return r.buf + (seq % STG) * RingCfg<VDISP, STG>::STAGE_BYTES;
}

// The two particles of pair `pi` of the tile in ring slot `seq` (this lane consumes pairs lane and lane+32).
template <bool VDISP, int STG = RingCfg<VDISP>::STAGES>
__device__ __forceinline__ void ring_read_pair(const WarpRing &r, uint32_t seq, int pi, Particle &a, Particle &b)
{
    using C = RingCfg<VDISP, STG>;
    const unsigned char *src = r.buf + (seq % C::STAGES) * C::STAGE_BYTES;
    const double2 x = reinterpret_cast<const double2 *>(src + 0 * C::STREAM_BYTES)[pi];
    const double2 y = reinterpret_cast<const double2 *>(src + 1 * C::STREAM_BYTES)[pi];
    const double2 z = reinterpret_cast<const double2 *>(src + 2 * C::STREAM_BYTES)[pi];
    const double2 m = reinterpret_cast<const double2 *>(src + 3 * C::STREAM_BYTES)[pi];
    make_particle(x.x, y.x, z.x, m.x, a);
    make_particle(x.y, y.y, z.y, m.y, b);
    if (VDISP) {
        const double2 u = reinterpret_cast<const double2 *>(src + 4 * C::STREAM_BYTES)[pi];
        const double2 v = reinterpret_cast<const double2 *>(src + 5 * C::STREAM_BYTES)[pi];
        const double2 w = reinterpret_cast<const double2 *>(src + 6 * C::STREAM_BYTES)[pi];
        make_tensor_products(u.x, v.x, w.x, a);
        make_tensor_products(u.y, v.y, w.y, b);
    }
}

// The bulk copy of a partial tile leaves the tail of the slot untouched: fill it with the never-selected
// sentinel the gather pads odd objects with, so that the compute loop needs no per-lane bounds checks.
template <bool VDISP, int STG = RingCfg<VDISP>::STAGES>
__device__ __forceinline__ void ring_fill_tail(const WarpRing &r, uint32_t seq, int lane, int pairs_in_tile)
{
    using C = RingCfg<VDISP, STG>;
    unsigned char *dst = r.buf + (seq % C::STAGES) * C::STAGE_BYTES;
    for (int pi = pairs_in_tile + lane; pi < TILE_PAIRS; pi += 32) {
#pragma unroll
        for (int st = 0; st < C::NSTREAM; st++)
            reinterpret_cast<double2 *>(dst + st * C::STREAM_BYTES)[pi] =
                (st < 3) ? make_double2(1e150, 1e150) : make_double2(0.0, 0.0);
    }
    __syncwarp();
}

// The per-lane copies of ring_issue_lanes as a cp.async GROUP of the issuing thread (no mbarrier): lane l copies -- and later
// consumes -- pairs l and l + 32 of every stream, so a thread only ever reads shared memory that it filled itself and
// cp.async.wait_group is all the synchronisation a streamed tile needs.
#ifndef CPG_RT_GROUPS
#define CPG_RT_GROUPS 1
#endif
template <bool VDISP, int STG>
__device__ __forceinline__ void ring_issue_group(const WarpRing &r, uint32_t stage, const SoA &soa, int64_t base, int t, int npairs, int lane)
{
    using C = RingCfg<VDISP, STG>;
    const int n = min(TILE_PAIRS, npairs - t * TILE_PAIRS);
    const uint32_t dst = r.buf_s + stage * C::STAGE_BYTES;
    const int64_t o = base + (int64_t)t * (2 * TILE_PAIRS);
    const double *src[7] = {soa.dx, soa.dy, soa.dz, soa.m, soa.vx, soa.vy, soa.vz};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int pi = lane + 32 * h;
        if (pi < n) {
#pragma unroll
            for (int st = 0; st < C::NSTREAM; st++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + st * C::STREAM_BYTES + pi * 16u),
                             "l"(src[st] + o + 2 * pi)
                             : "memory");
        }
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One full pass over the object's radial prefix.  This warp consumes tiles tile0, tile0+tstride, ... of the
// `npairs` pairs that the still-active shells need; npr[s] = pairs in the prefix of shell s.
// resident: every tile this warp will ever need for the task fits in its ring -> loaded by the first pass,
// re-read from shared memory by all later ones.
template <int S, bool SHELL, bool REDUCED, bool VDISP, bool LANES = false, int STG = RingCfg<VDISP>::STAGES>
__device__ __forceinline__ void run_pass(const SoA &soa, int64_t base, int npairs, int tile0, int tstride, int lane,
                                         WarpRing &ring, bool resident, int npairs_task, const double *frames,
                                         const int (&npr)[S], const int *nin, unsigned amask, double (&acc)[S][7],
                                         int (&cnt)[S])
{
    // tin[s]: whole tiles of shell s's inner prefix -- served by the prefix tensor table, not streamed
    int tin[S], tlow = 0x7fffffff;
#pragma unroll
    for (int s = 0; s < S; s++) {
        tin[s] = nin[s] / TILE_PAIRS;
        if (amask & (1u << s)) tlow = min(tlow, tin[s]);
    }
#pragma unroll
    for (int s = 0; s < S; s++) {
        cnt[s] = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) acc[s][k] = 0.0;
    }
    constexpr int RING_STAGES = STG;
    // a resident task keeps all of its tiles in the ring (slot j <-> tile tile0 + j * tstride) and just skips the
    // ones no shell needs; a streamed pass starts at the first tile some active shell still has to test
    const int tskip = resident ? 0 : tlow;
    tile0 += tskip;
    const int ntiles = (npairs + TILE_PAIRS - 1) / TILE_PAIRS;
    const int nmine = tile0 < ntiles ? (ntiles - tile0 + tstride - 1) / tstride : 0;
    const bool reuse = resident && ring.res_loaded;
    const int npairs_load = resident ? npairs_task : npairs;
    uint32_t seq0 = reuse ? ring.res_seq : ring.seq;
    int issue_n = nmine;
    if (resident && !reuse) {   // first pass of a resident task: bring in everything the task will ever need
        const int nt_task = (npairs_task + TILE_PAIRS - 1) / TILE_PAIRS;
        issue_n = tile0 < nt_task ? (nt_task - tile0 + tstride - 1) / tstride : 0;
    }
    constexpr bool GROUPS = LANES && CPG_RT_GROUPS;   // per-lane cp.async groups instead of mbarriers (see ring_issue_group)
    if (!reuse && GROUPS) {
        for (int j = 0; j < RING_STAGES; j++) {   // exactly RING_STAGES groups in flight at any time (empty ones included)
            if (j < issue_n) ring_issue_group<VDISP, STG>(ring, (seq0 + j) % RING_STAGES, soa, base, tile0 + j * tstride, npairs_load, lane);
            cp_async_commit();
        }
    } else if (!reuse && LANES) {
        for (int j = 0; j < min(issue_n, RING_STAGES); j++)
            ring_issue_lanes<VDISP, STG>(ring, seq0 + j, soa, base, tile0 + j * tstride, npairs_load, lane);
    } else if (!reuse && lane == 0)
        for (int j = 0; j < min(issue_n, RING_STAGES); j++)
            ring_issue<VDISP, STG>(ring, seq0 + j, soa, base, tile0 + j * tstride, npairs_load);
    for (int j = 0; j < nmine; j++) {
        const int t = tile0 + j * tstride;
        const uint32_t seq = seq0 + j;
        if (!reuse) {
            if (GROUPS) cp_async_wait<RING_STAGES - 1>();
            else mbar_wait(ring.bar_s + (seq % RING_STAGES) * 8u, (seq / RING_STAGES) & 1u);
            // only the last tile of the loaded range can be partial (particles past the active prefix but
            // inside the loaded range are real particles that fail every membership test)
            const int loaded = min(TILE_PAIRS, npairs_load - t * TILE_PAIRS);
            if (loaded < TILE_PAIRS) {
                if (GROUPS) __syncwarp();   // the fill touches other lanes' pairs: they are done with the slot's previous tile
                ring_fill_tail<VDISP, STG>(ring, seq, lane, loaded);
            }
        }
        unsigned smask = 0;
#pragma unroll
        for (int s = 0; s < S; s++)
            if (t * TILE_PAIRS < npr[s] && t >= tin[s]) smask |= 1u << s;
        smask &= amask;
        asm volatile("" : "+r"(smask));   // keep the mask in its register: ptxas otherwise re-derives it before every shell block
#ifndef CPG_NO_TILE_STATS
        if (smask != 0) { ring.st_tiles += 1; ring.st_tshells += __popc(smask); }
#endif
        if (smask == 0) {
            // (resident tasks only)
        } else if (CPG_F32_TEST && !REDUCED) {
            const unsigned char *tile = ring_tile<VDISP, STG>(ring, seq);
#pragma unroll 1
            for (int h = 0; h < 2; h++) pair_f32<S, SHELL, VDISP>(tile, lane + 32 * h, frames, smask, acc, cnt);
        } else if (VDISP) {
            // two particles at a time: the dispersion variant carries 13 doubles per particle
#pragma unroll
            for (int h = 0; h < 2; h++) {
                Particle p[2];
                ring_read_pair<VDISP, STG>(ring, seq, lane + 32 * h, p[0], p[1]);
                passk_particles<S, SHELL, REDUCED, VDISP, 2>(p, frames, smask, acc, cnt);
            }
        } else if (CPG_FP64_PAIRS) {
            // (fewer registers and half the code; two fp64 chains per shell instead of four)
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                Particle p[2];
                ring_read_pair<VDISP, STG>(ring, seq, lane + 32 * h, p[0], p[1]);
                passk_particles<S, SHELL, REDUCED, VDISP, 2>(p, frames, smask, acc, cnt);
            }
        } else {
            Particle p[4];
            ring_read_pair<VDISP, STG>(ring, seq, lane, p[0], p[1]);
            ring_read_pair<VDISP, STG>(ring, seq, lane + 32, p[2], p[3]);
            if (CPG_SHELL_PAIRS && !SHELL && !REDUCED && S % 2 == 0)
                passk_shell_pairs<(S % 2 == 0 ? S : 2), 4>(p, frames, smask, reinterpret_cast<double (&)[(S % 2 == 0 ? S : 2)][7]>(acc), reinterpret_cast<int (&)[(S % 2 == 0 ? S : 2)]>(cnt));
            else
                passk_particles<S, SHELL, REDUCED, VDISP, 4>(p, frames, smask, acc, cnt);
        }
        if (GROUPS) {
            if (!reuse) {   // (a lane overwrites only the pairs it has just read itself: no warp barrier)
                if (!resident && j + RING_STAGES < nmine)
                    ring_issue_group<VDISP, STG>(ring, seq % RING_STAGES, soa, base, t + RING_STAGES * tstride, npairs_load, lane);
                cp_async_commit();
            }
        } else if (!resident && j + RING_STAGES < nmine) {
            __syncwarp();   // every lane has taken its data out of this slot
            if (LANES)
                ring_issue_lanes<VDISP, STG>(ring, seq + RING_STAGES, soa, base, t + RING_STAGES * tstride, npairs_load, lane);
            else if (lane == 0)
                ring_issue<VDISP, STG>(ring, seq + RING_STAGES, soa, base, t + RING_STAGES * tstride, npairs_load);
        }
    }
    if (!reuse) {
        if (GROUPS) {
            cp_async_wait<0>();   // nothing in flight between passes (resident tiles beyond the active range included)
            __syncwarp();
        }
        if (resident) {
            // the barriers of resident tiles beyond the first pass's active range have not been waited on yet
            for (int j = nmine; j < issue_n; j++) {
                if (!GROUPS) mbar_wait(ring.bar_s + ((seq0 + j) % RING_STAGES) * 8u, ((seq0 + j) / RING_STAGES) & 1u);
                const int loaded = min(TILE_PAIRS, npairs_load - (tile0 + j * tstride) * TILE_PAIRS);
                if (loaded < TILE_PAIRS) ring_fill_tail<VDISP, STG>(ring, seq0 + j, lane, loaded);
            }
            ring.res_seq = seq0;
            ring.res_loaded = true;
            ring.seq = seq0 + issue_n;
        } else {
            ring.seq = seq0 + nmine;
        }
    }
}

// run_pass for kernel 1d: a warp-kernel pass (tile0 = 0, stride 1, per-lane cp.async loads) over a RUNTIME set of stages
// [sbase, sbase + nst) of the warp's RING_MAX_STAGES-stage ring, nst a power of two.  The two task slots of a warp share
// the ring: a task of <= 2 tiles stays resident in its slot's half, a streamed pass takes the whole ring when the other
// slot holds nothing resident and its own half otherwise.  Which phase of a stage's mbarrier comes next is tracked in
// `phase` (one bit per stage; every issue to a stage is waited for exactly once before the next issue to it).
template <int S, bool SHELL, bool REDUCED, bool VDISP>
__device__ __forceinline__ void run_pass_rt(const SoA &soa, int64_t base, int npairs, int lane, WarpRing &ring, unsigned &phase,
                                            int sbase, int nst, bool resident, bool &res_loaded, int npairs_task,
                                            const double *frames, const int (&npr)[S], const int *nin, unsigned amask,
                                            double (&acc)[S][7], int (&cnt)[S])
{
    constexpr int STG = RING_MAX_STAGES;
    static_assert(!VDISP, "position streams only");
    int tin[S], tlow = 0x7fffffff;
#pragma unroll
    for (int s = 0; s < S; s++) {
        tin[s] = nin[s] / TILE_PAIRS;
        if (amask & (1u << s)) tlow = min(tlow, tin[s]);
    }
#pragma unroll
    for (int s = 0; s < S; s++) {
        cnt[s] = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) acc[s][k] = 0.0;
    }
    const int tile0 = resident ? 0 : tlow;
    const int ntiles = (npairs + TILE_PAIRS - 1) / TILE_PAIRS;
    const int nmine = tile0 < ntiles ? ntiles - tile0 : 0;
    const bool reuse = resident && res_loaded;
    const int npairs_load = resident ? npairs_task : npairs;
    const int issue_n = (resident && !reuse) ? (npairs_task + TILE_PAIRS - 1) / TILE_PAIRS : nmine;   // (resident: <= nst tiles)
    const int smask_st = nst - 1;
    if (!reuse) {
        if (CPG_RT_GROUPS) {
            // exactly nst groups in flight at any time (empty ones where there is nothing to copy): tile j is complete when
            // at most nst - 1 younger groups are pending
            for (int j = 0; j < nst; j++) {
                if (j < issue_n) ring_issue_group<VDISP, STG>(ring, (uint32_t)(sbase + j), soa, base, tile0 + j, npairs_load, lane);
                cp_async_commit();
            }
        } else {
            for (int j = 0; j < min(issue_n, nst); j++)
                ring_issue_lanes<VDISP, STG>(ring, (uint32_t)(sbase + j), soa, base, tile0 + j, npairs_load, lane);
        }
    }
    for (int j = 0; j < nmine; j++) {
        const int t = tile0 + j;
        const uint32_t stage = (uint32_t)(sbase + (j & smask_st));
        if (!reuse) {
            if (CPG_RT_GROUPS) {
                if (nst == 4) cp_async_wait<3>(); else cp_async_wait<1>();
            } else {
                mbar_wait(ring.bar_s + stage * 8u, (phase >> stage) & 1u);
                phase ^= 1u << stage;
            }
            const int loaded = min(TILE_PAIRS, npairs_load - t * TILE_PAIRS);
            if (loaded < TILE_PAIRS) {
                if (CPG_RT_GROUPS) __syncwarp();   // the fill touches other lanes' pairs: they are done with the stage's previous tile
                ring_fill_tail<VDISP, STG>(ring, stage, lane, loaded);
            }
        }
        unsigned smask = 0;
#pragma unroll
        for (int s = 0; s < S; s++)
            if (t * TILE_PAIRS < npr[s] && t >= tin[s]) smask |= 1u << s;
        smask &= amask;
        asm volatile("" : "+r"(smask));
#ifndef CPG_NO_TILE_STATS
        if (smask != 0) { ring.st_tiles += 1; ring.st_tshells += __popc(smask); }
#endif
        if (smask != 0) {
            Particle p[4];
            ring_read_pair<VDISP, STG>(ring, stage, lane, p[0], p[1]);
            ring_read_pair<VDISP, STG>(ring, stage, lane + 32, p[2], p[3]);
            passk_particles<S, SHELL, REDUCED, VDISP, 4>(p, frames, smask, acc, cnt);
        }
        if (CPG_RT_GROUPS) {
            if (!reuse) {
                // (a lane overwrites only the pairs it has just read itself: no warp barrier)
                if (!resident && j + nst < nmine) ring_issue_group<VDISP, STG>(ring, stage, soa, base, t + nst, npairs_load, lane);
                cp_async_commit();
            }
        } else if (!resident && j + nst < nmine) {
            __syncwarp();   // every lane has taken its data out of this stage
            ring_issue_lanes<VDISP, STG>(ring, stage, soa, base, t + nst, npairs_load, lane);
        }
    }
    if (CPG_RT_GROUPS && !reuse) {
        cp_async_wait<0>();   // nothing in flight between passes (resident tiles beyond the active range included)
        __syncwarp();
    }
    if (!reuse && resident) {
        // the resident tiles beyond the first pass's active range have not been waited on / completed yet
        for (int j = nmine; j < issue_n; j++) {
            const uint32_t stage = (uint32_t)(sbase + j);
            if (!CPG_RT_GROUPS) {
                mbar_wait(ring.bar_s + stage * 8u, (phase >> stage) & 1u);
                phase ^= 1u << stage;
            }
            const int loaded = min(TILE_PAIRS, npairs_load - j * TILE_PAIRS);
            if (loaded < TILE_PAIRS) ring_fill_tail<VDISP, STG>(ring, stage, lane, loaded);
        }
        res_loaded = true;
    }
}

// run_pass out of line.  Inlined into the task loop of a kernel, the pass competes for registers with everything the
// task keeps alive (Katz state, shell geometry, queue and table indices ...): at the 168-register cap ptxas then serialises
// the membership chains, re-derives masks and addresses inside the tile loop and a lone warp needs ~1700 cycles per tile,
// against ~650 for the same source compiled on its own (profiles/micro/loop_lab.cu).  As a function of its own the pass
// gets the whole register budget; the caller's live values are saved around the call once per pass, and the 8 S sums
// come back through local memory (28 words per lane and pass).
#ifndef CPG_OUTLINE_PASS
#define CPG_OUTLINE_PASS 0   // measured slower (call + local-memory round trip of the sums: 7.4 us per step instead of 3.7)
#endif
template <int S, bool SHELL, bool REDUCED, bool VDISP, bool LANES>
__device__ __noinline__ void run_pass_outlined(SoA soa, int64_t base, int npairs, int tile0, int tstride, int lane,
                                               WarpRing *ringp, bool resident, int npairs_task, const double *frames,
                                               const int *nprp, const int *nin, unsigned amask, double *acc_out, int *cnt_out)
{
    WarpRing ring = *ringp;
    int npr[S];
#pragma unroll
    for (int s = 0; s < S; s++) npr[s] = nprp[s];
    double acc[S][7];
    int cnt[S];
    run_pass<S, SHELL, REDUCED, VDISP, LANES>(soa, base, npairs, tile0, tstride, lane, ring, resident, npairs_task, frames, npr,
                                              nin, amask, acc, cnt);
    *ringp = ring;
#pragma unroll
    for (int s = 0; s < S; s++) {
#pragma unroll
        for (int c = 0; c < 7; c++) acc_out[s * 7 + c] = acc[s][c];
        cnt_out[s] = cnt[s];
    }
}

template <int S>
__device__ __forceinline__ int active_npairs(const int (&npr)[S], unsigned amask)
{
    int n = 0;
#pragma unroll
    for (int s = 0; s < S; s++)
        if (amask & (1u << s)) n = max(n, npr[s]);
    return n;
}

// ------------------------------------------------------------------------------------------------
// Kernel 1: a WARP per task (object, group of <= S shells).  No block-level synchronisation at all:
// warp-shuffle reductions, eigen-solves spread over lanes (lane l serves shell l % S), persistent
// warps pulling tasks (sorted largest first) from an atomic queue.
// ------------------------------------------------------------------------------------------------
#ifndef CPG_WARP_KERNEL_WARPS
#define CPG_WARP_KERNEL_WARPS 4
#endif
constexpr int WARP_KERNEL_WARPS = CPG_WARP_KERNEL_WARPS;   // warps per CTA of the warp kernel
#ifndef CPG_WARP_RING_LANES
#define CPG_WARP_RING_LANES 1   // 1: per-lane cp.async tile loads in the warp kernel; 0: elected-lane TMA bulk copies
#endif
constexpr bool WARP_RING_LANES = CPG_WARP_RING_LANES != 0;
#ifndef CPG_WARP_MINB
#define CPG_WARP_MINB 3   // resident CTAs per SM the register allocator must leave room for
#endif

// Lane <-> shell mapping shared by both kernels.  The transposed warp reduction leaves the total of value
// (l mod 8S) in lane l; value 8s+c is component c (Sxx,Sxy,Sxz,Syy,Syz,Szz,M,count) of shell s.  So the lanes
// [8s, 8s+8) (and their images modulo 8S) serve shell s, and lane 8s is the one that writes its frame.
template <int S, int G = 1>
struct LaneMap {
    // G == 1: S shells per task, the transposed reduction leaves value (l mod 8S) in lane l, lanes [8s,8s+8) serve
    //         shell s.  G == 2 (S == 4): two groups of 4 shells are passed over one after the other, their sums
    //         sit in two registers, and lanes [4s,4s+4) serve shell s of the 8.
    static_assert(G == 1 || (G == 2 && S == 4), "two groups only for S == 4");
    static constexpr int NV = 8 * S;
    static constexpr int NS = S * G;
    int ls;        // shell served by this lane
    bool writer;   // this lane publishes frames / results of shell ls
    __device__ __forceinline__ explicit LaneMap(int lane)
        : ls(G == 1 ? ((lane & (NV - 1)) >> 3) : (lane >> 2)),
          writer(G == 1 ? (lane < NV && (lane & 7) == 0) : ((lane & 3) == 0)) {}
    __device__ static __forceinline__ unsigned active_mask(bool active_writer)
    {
        const unsigned b = __ballot_sync(0xffffffffu, active_writer);
        unsigned m = 0;
#pragma unroll
        for (int s = 0; s < NS; s++) m |= ((b >> ((G == 1 ? 8 : 4) * s)) & 1u) << s;
        return m;
    }
};

template <bool SHELL, class LM>
__device__ __forceinline__ void init_frames(double *frames, int *np, int *nin, const LM &L, const KatzState &st,
                                            int ne, const InnerPrefix &ip, int k)
{
    if (L.writer) {
        // pass 0: the sphere of radius d (its inner prefix comes from the table) / the spherical shell [d - delta, d)
        // (the particles below d - delta are skipped)
        nin[L.ls] = !st.active ? 0 : SHELL ? ip.pairs_below_shell(k, 1.0, st.d, st.delta) : ip.pairs_inside(k, 1.0);
        double *F = frames + L.ls * F_SLOTS;
        F[F_D2] = __dmul_rn(st.d, st.d);
        const double in = __dsub_rn(st.d, st.delta);
        F[F_THR] = __dmul_rn(in, in);   // pass 0: r2 >= (d-delta)^2 (shell variant; 0 for the first shell)
#pragma unroll
        for (int r = 0; r < 6; r++) F[F_AXX + r] = F[F_BXX + r] = (r < 3) ? 1.0 : 0.0;   // A = B = I
        // starting basis: major = x, intermediate = y, minor = z (v[6..8], v[3..5], v[0..2]), so that exact_member's
        // p0^2 + p1^2/1 + p2^2/1 is the reference's (x^2 + y^2) + z^2 of pass 0, same order
#pragma unroll
        for (int r = 0; r < 9; r++) F[F_V + r] = (r == 2 || r == 4 || r == 6) ? 1.0 : 0.0;
        F[F_QQ] = 1.0; F[F_SS] = 1.0;
        if (CPG_F32_TEST) {
            F[F_BA2] = 1.0; F[F_BB2] = 1.0; F[F_BC2] = 1.0;
            write_f32_frame(F, 1.0, 1.0, true);
        }
        np[L.ls] = (ne + 1) >> 1;
    }
}

// What the prefix tensor table contributes to value `comp` (0..5 tensor sums, 7 particle count) of a shell whose
// inner prefix holds `pairs_in` pairs: the row of its last whole tile.  row0 = table row of the object's tile 0.
__device__ __forceinline__ double table_term(const double *ptab, int64_t row0, int pairs_in, int comp)
{
    const int tin = pairs_in / TILE_PAIRS;
    if (!ptab || tin == 0) return 0.0;
    if (comp < 6) return ptab[(row0 + tin - 1) * 6 + comp];
    return comp == 7 ? (double)(2 * TILE_PAIRS) * (double)tin : 0.0;
}

template <int S>
__device__ __forceinline__ void pack_values(const double (&acc)[S][7], const int (&cnt)[S], double (&val)[8 * S])
{
#pragma unroll
    for (int s = 0; s < S; s++) {
#pragma unroll
        for (int c = 0; c < 7; c++) val[8 * s + c] = acc[s][c];
        val[8 * s + 7] = (double)cnt[s];   // exact below 2^53
    }
}

// Per halo-shell iteration state parked in shared memory while the particle pass runs (see CPG_PARK_STATE).
struct ShellSlot {
    double q, s, d, delta;
    unsigned long long out_idx;
    int iteration, active, k, pad;
};
// 1 (default): the warp kernel keeps the Katz state of its shells (q, s, d, delta, iteration, output index) in shared
// memory between the per-iteration steps instead of in registers.  The particle pass compiled on its own uses 156-164 of the
// 168 registers (profiles/micro/loop_lab.cu); every value the task loop keeps alive across it pushes ptxas into
// serialising the membership chains, re-deriving masks and addresses inside the tile loop and spilling -- a lone warp then
// needs ~1700 cycles per tile instead of ~1200 (streamed) / ~900 (ring-resident).
#ifndef CPG_PARK_STATE
#define CPG_PARK_STATE 1
#endif

template <int S, int G, bool SHELL, bool REDUCED, bool VDISP>
__global__ void __launch_bounds__(WARP_KERNEL_WARPS * 32, CPG_WARP_MINB) morph_warp_kernel(const MorphParams P)
{
    using LM = LaneMap<S, G>;
    constexpr int NS = S * G;   // shells per task; they are passed over in G groups of S
    extern __shared__ __align__(128) unsigned char s_ring_w[];   // WARP_KERNEL_WARPS rings
    __shared__ __align__(16) double s_frames[WARP_KERNEL_WARPS][NS * F_SLOTS];
    __shared__ __align__(8) uint64_t s_bar[WARP_KERNEL_WARPS][RING_MAX_STAGES];
    __shared__ int s_np[WARP_KERNEL_WARPS][NS], s_nin[WARP_KERNEL_WARPS][NS];
    __shared__ float s_rr[MAX_SORT_R];                              // shell radii rounded up (inner-prefix lookup)
    __shared__ uint16_t s_neff[WARP_KERNEL_WARPS][MAX_SORT_R];      // prefix lengths (pairs) of the warp's current object
    __shared__ ShellSlot s_slot[CPG_PARK_STATE ? WARP_KERNEL_WARPS : 1][NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *frames = s_frames[warp];
    int *np = s_np[warp], *nin = s_nin[warp];
    const LM L(lane);
    const int ls = L.ls;
    WarpRing ring;
    ring_init(ring, s_ring_w + warp * RingCfg<VDISP>::WARP_BYTES, s_bar[warp], lane, WARP_RING_LANES ? 32u : 1u);
    const bool use_table = !SHELL && !REDUCED && P.n_eff != nullptr && P.ptab != nullptr;   // inner prefix from the prefix tensor table
    const bool use_inner = use_table || (SHELL && P.n_eff != nullptr);                        // ... or skipped (shell variant)
    if (use_inner) {
        for (int j = threadIdx.x; j < P.R; j += blockDim.x) s_rr[j] = __double2float_ru(P.rr[j]);
        __syncthreads();
    }
    InnerPrefix ip;
    ip.rr = nullptr; ip.neff = nullptr;
    ip.rr_up = use_inner ? s_rr : nullptr;
    ip.npairs16 = s_neff[warp];
    const double *ptab = use_table ? P.ptab : nullptr;

    for (;;) {
        int t = 0;
        if (lane == 0)
            t = atomicAdd(P.task_counter, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= P.n_tasks)
            break;
        const Task tk = P.tasks[t];
        const int halo = tk.halo;
        if (P.skip[halo] || (P.local && P.r200[halo] == 0.0))
            continue;   // shape_profs_algos.pyx:1024-1029: outputs stay zero, counters stay -1
#ifdef CPG_TASK_TRACE
        unsigned long long trace_t0, trace_steps = 0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(trace_t0));
        const uint32_t trace_tiles0 = ring.st_tiles;
#endif
        const int N = P.soa.size[halo];
        const int64_t base = P.soa.poff[halo];

        KatzState st;
        st.q = 1.0; st.s = 1.0; st.iteration = 1;
        st.active = ls < tk.nshell;
        const int k = tk.shell0 + (st.active ? ls : 0);
        shell_geometry(P, halo, k, st.d, st.delta);
        const size_t out_idx = (size_t)halo * P.R + k;
        __syncwarp();
        if (use_inner) {   // the object's prefix lengths up to the task's last shell
            for (int j = lane; j < tk.shell0 + tk.nshell; j += 32)
                s_neff[warp][j] = (uint16_t)(P.n_eff[(size_t)halo * P.R + j] >> 1);   // objects of the warp kernel: < 131072 particles
            __syncwarp();
        }
        init_frames<SHELL>(frames, np, nin, L, st, (P.n_eff && st.active) ? P.n_eff[out_idx] : N, ip, k);
        if (CPG_PARK_STATE && L.writer) {
            ShellSlot &sl = s_slot[warp][ls];
            sl.q = 1.0; sl.s = 1.0; sl.d = st.d; sl.delta = st.delta; sl.out_idx = (unsigned long long)out_idx;
            sl.iteration = 1; sl.active = st.active ? 1 : 0; sl.k = k;
        }
        __syncwarp();
        unsigned amask = LM::active_mask(st.active && L.writer);
        int npr[G][S];
        int npairs_task = 0;
#pragma unroll
        for (int g = 0; g < G; g++) {
#pragma unroll
            for (int s = 0; s < S; s++) npr[g][s] = np[g * S + s];
            npairs_task = max(npairs_task, active_npairs<S>(npr[g], (amask >> (g * S)) & ((1u << S) - 1u)));
        }
        const bool resident = (npairs_task + TILE_PAIRS - 1) / TILE_PAIRS <= RingCfg<VDISP>::STAGES;
        ring.res_loaded = false;
        const int64_t row0 = (base >> 7) + halo;

        while (true) {
            // one pass per group of S shells; every group leaves its 8S sums transposed in one register
            double mine[G];
#pragma unroll
            for (int g = 0; g < G; g++) {
                const unsigned am = (amask >> (g * S)) & ((1u << S) - 1u);
                mine[g] = 0.0;
                if (am) {
                    // this lane's value (lane mod 8S) of the reduction = component (lane & 7) of shell (lane mod 8S) / 8:
                    // fetch what the table adds for the shell's inner prefix now, consume it after the pass
                    const double tadd = table_term(ptab, row0, nin[g * S + ((lane & (8 * S - 1)) >> 3)], lane & 7);
                    double acc[S][7];
                    int cnt[S];
                    if (CPG_OUTLINE_PASS)
                        run_pass_outlined<S, SHELL, REDUCED, VDISP, WARP_RING_LANES>(P.soa, base, active_npairs<S>(npr[g], am), 0, 1, lane,
                                                                                    &ring, resident, npairs_task, frames + g * S * F_SLOTS,
                                                                                    np + g * S, nin + g * S, am, &acc[0][0], cnt);
                    else
                    run_pass<S, SHELL, REDUCED, VDISP, WARP_RING_LANES>(P.soa, base, active_npairs<S>(npr[g], am), 0, 1, lane, ring, resident,
                                                       npairs_task, frames + g * S * F_SLOTS, npr[g], nin + g * S, am, acc,
                                                       cnt);
                    double val[8 * S];
                    pack_values<S>(acc, cnt, val);
                    mine[g] = warp_sum_transpose<8 * S>(val, lane) + tadd;
                }
            }
            double tot[7], vprev[9];
            int pts;
            if (G == 1) {
#pragma unroll
                for (int c = 0; c < 7; c++) tot[c] = __shfl_sync(0xffffffffu, mine[0], (lane & ~7) + c);
                pts = (int)__shfl_sync(0xffffffffu, mine[0], (lane & ~7) + 7);
            } else {
                const int src = (ls & (S - 1)) * 8;
                const bool hi = ls >= S;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const double a = __shfl_sync(0xffffffffu, mine[0], src + c);
                    const double b = __shfl_sync(0xffffffffu, mine[G - 1], src + c);
                    const double v = hi ? b : a;
                    if (c < 7) tot[c] = v; else pts = (int)v;
                }
            }
#pragma unroll
            for (int r = 0; r < 9; r++) vprev[r] = frames[ls * F_SLOTS + F_V + r];
            __syncwarp();
            if (CPG_PARK_STATE) {
                // the state of this lane's shell comes back from shared memory, goes through the step and is parked again
                ShellSlot &sl = s_slot[warp][ls];
                KatzState sv;
                sv.q = sl.q; sv.s = sl.s; sv.d = sl.d; sv.delta = sl.delta; sv.iteration = sl.iteration; sv.active = sl.active != 0;
                const int kk = sl.k;
                const size_t oidx = (size_t)sl.out_idx;
                __syncwarp();
                katz_step<SHELL>(sv, tot, pts, P, vprev, frames + ls * F_SLOTS, L.writer, oidx, L.writer, ip, kk, nin + ls);
                if (L.writer) { sl.q = sv.q; sl.s = sv.s; sl.iteration = sv.iteration; sl.active = sv.active ? 1 : 0; }
                __syncwarp();
                amask = LM::active_mask(sv.active && L.writer);
            } else {
            katz_step<SHELL>(st, tot, pts, P, vprev, frames + ls * F_SLOTS, L.writer, out_idx, L.writer, ip, k, nin + ls);
            __syncwarp();
            amask = LM::active_mask(st.active && L.writer);
            }
#ifdef CPG_TASK_TRACE
            trace_steps++;
#endif
            if (amask == 0)
                break;
        }
#ifdef CPG_TASK_TRACE
        if (lane == 0 && P.stats && t < (1 << 19)) {
            unsigned long long t1; unsigned smid;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            P.stats[2 + 4 * (size_t)t + 0] = trace_t0; P.stats[2 + 4 * (size_t)t + 1] = t1;
            P.stats[2 + 4 * (size_t)t + 2] = ((unsigned long long)smid << 32) | (unsigned)(ring.st_tiles - trace_tiles0);
            P.stats[2 + 4 * (size_t)t + 3] = (trace_steps << 32) | (unsigned)N;
        }
#else
        if (ring.st_tiles > (1u << 20)) ring_flush_stats(ring, P.stats, lane);
#endif
    }
    ring_flush_stats(ring, P.stats, lane);
}

// ------------------------------------------------------------------------------------------------
// Kernel 1b: the warp kernel with a BATCHED step.  The per-iteration step (eigen-solve, convergence test, next frame)
// costs the same fp64-pipe time for 4 useful lanes as for 32 (profiles/micro/fp64_lanes.cu); in kernel 1 every warp spends a third of its fp64 instructions on the step of its own 4 shells.
// Here the WB_WARPS warps of a CTA stream their own tasks as before but meet at a block barrier once per Katz iteration:
// every warp leaves its 8 S reduced sums in shared memory, ONE warp (they take turns) runs the step for all
// WB_WARPS x S shells at once -- lane l serves shell l % S of warp l / S -- and publishes frames, inner-prefix lengths
// and active masks for everybody.  Per-shell arithmetic is unchanged (same katz_step, same operands), so the results
// are those of kernel 1 bit for bit.  The shell state (q, s, iteration ...) lives in shared memory, where the stepping
// warp finds it.  A warp whose task has finished takes the next one from the queue and re-joins at the next barrier.
// The host orders the tasks by estimated pass length, so that the warps of a CTA wait for one another as little as possible.
// ------------------------------------------------------------------------------------------------
#ifndef CPG_WB_WARPS
#define CPG_WB_WARPS 4
#endif
constexpr int WB_WARPS = CPG_WB_WARPS;

template <int S, bool SHELL, bool REDUCED, bool VDISP>
__global__ void __launch_bounds__(WB_WARPS * 32, CPG_WARP_MINB) morph_wbatch_kernel(const MorphParams P)
{
    static_assert(WB_WARPS * S <= 32, "one lane per shell in the batched step");
    constexpr int NV = 8 * S;
    extern __shared__ __align__(128) unsigned char s_ring_b[];   // WB_WARPS rings
    __shared__ __align__(16) double s_frames[WB_WARPS][S * F_SLOTS];
    __shared__ __align__(8) uint64_t s_bar[WB_WARPS][RING_MAX_STAGES];
    __shared__ int s_np[WB_WARPS][S], s_nin[WB_WARPS][S];
    __shared__ float s_rr[MAX_SORT_R];
    __shared__ uint16_t s_neff[WB_WARPS][MAX_SORT_R];
    __shared__ double s_red[WB_WARPS][NV];
    __shared__ ShellSlot s_slot[WB_WARPS][S];
    __shared__ long long s_row0[WB_WARPS];
    __shared__ unsigned s_amask[WB_WARPS];
    __shared__ int s_have[WB_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *frames = s_frames[warp];
    int *np = s_np[warp], *nin = s_nin[warp];
    const LaneMap<S> L(lane);
    WarpRing ring;
    ring_init(ring, s_ring_b + warp * RingCfg<VDISP>::WARP_BYTES, s_bar[warp], lane, WARP_RING_LANES ? 32u : 1u);
    const bool use_table = !SHELL && !REDUCED && P.n_eff != nullptr && P.ptab != nullptr;   // inner prefix from the prefix tensor table
    const bool use_inner = use_table || (SHELL && P.n_eff != nullptr);                        // ... or skipped (shell variant)
    if (use_inner)
        for (int j = threadIdx.x; j < P.R; j += blockDim.x) s_rr[j] = __double2float_ru(P.rr[j]);
    __syncthreads();
    const double *ptab = use_table ? P.ptab : nullptr;

    bool have = false;
    unsigned amask = 0;
    int npr[S];
    int npairs_task = 0;
    bool resident = false;
    int64_t base = 0, row0 = 0;
#pragma unroll
    for (int s = 0; s < S; s++) npr[s] = 0;

    for (int iter = 0;; iter++) {
        if (!have) {
            // next live task of the queue (dead ones -- empty object, r200 == 0 -- are skipped: outputs stay zero / -1)
            for (;;) {
                int t = 0;
                if (lane == 0) t = atomicAdd(P.task_counter, 1);
                t = __shfl_sync(0xffffffffu, t, 0);
                if (t >= P.n_tasks) break;
                const Task tk = P.tasks[t];
                const int halo = tk.halo;
                if (P.skip[halo] || (P.local && P.r200[halo] == 0.0)) continue;
                const int N = P.soa.size[halo];
                base = P.soa.poff[halo];
                row0 = (base >> 7) + halo;
                KatzState st;
                st.q = 1.0; st.s = 1.0; st.iteration = 1;
                st.active = L.ls < tk.nshell;
                const int k = tk.shell0 + (st.active ? L.ls : 0);
                shell_geometry(P, halo, k, st.d, st.delta);
                const size_t out_idx = (size_t)halo * P.R + k;
                if (use_inner) {
                    for (int j = lane; j < tk.shell0 + tk.nshell; j += 32)
                        s_neff[warp][j] = (uint16_t)(P.n_eff[(size_t)halo * P.R + j] >> 1);
                    __syncwarp();
                }
                InnerPrefix ip;
                ip.rr = nullptr; ip.neff = nullptr;
                ip.rr_up = use_inner ? s_rr : nullptr;
                ip.npairs16 = s_neff[warp];
                init_frames<SHELL>(frames, np, nin, L, st, (P.n_eff && st.active) ? P.n_eff[out_idx] : N, ip, k);
                if (L.writer) {
                    ShellSlot &sl = s_slot[warp][L.ls];
                    sl.q = 1.0; sl.s = 1.0; sl.d = st.d; sl.delta = st.delta; sl.out_idx = (unsigned long long)out_idx;
                    sl.iteration = 1; sl.active = st.active ? 1 : 0; sl.k = k;
                }
                __syncwarp();
                amask = LaneMap<S>::active_mask(st.active && L.writer);
                npairs_task = 0;
#pragma unroll
                for (int s = 0; s < S; s++) npr[s] = np[s];
                npairs_task = active_npairs<S>(npr, amask);
                resident = (npairs_task + TILE_PAIRS - 1) / TILE_PAIRS <= RingCfg<VDISP>::STAGES;
                ring.res_loaded = false;
                if (lane == 0) s_row0[warp] = row0;
                have = true;
                break;
            }
        }
        if (lane == 0) s_have[warp] = have ? 1 : 0;
        if (!__syncthreads_or(have ? 1 : 0)) break;   // queue empty and every task of the CTA finished
        if (have) {
            const double tadd = table_term(ptab, row0, nin[(lane & (NV - 1)) >> 3], lane & 7);
            double acc[S][7];
            int cnt[S];
            run_pass<S, SHELL, REDUCED, VDISP, WARP_RING_LANES>(P.soa, base, active_npairs<S>(npr, amask), 0, 1, lane, ring, resident,
                                                                npairs_task, frames, npr, nin, amask, acc, cnt);
            double val[NV];
            pack_values<S>(acc, cnt, val);
            const double mine = warp_sum_transpose<NV>(val, lane) + tadd;
            if (lane < NV) s_red[warp][lane] = mine;
        }
        __syncthreads();   // all sums in place
        if (warp == (iter % WB_WARPS)) {
            const int w = lane / S, sh = lane % S;
            const bool valid = lane < WB_WARPS * S && s_have[w] != 0;
            KatzState st;
            st.q = 1.0; st.s = 1.0; st.d = 1.0; st.delta = 1.0; st.iteration = 1; st.active = false;
            double tot[7] = {0, 0, 0, 0, 0, 0, 0}, vprev[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            int pts = 0, k = 0;
            size_t out_idx = 0;
            const int ws = valid ? w : 0;   // (keeps every lane's addresses in range)
            if (valid) {
                const ShellSlot &sl = s_slot[w][sh];
                st.q = sl.q; st.s = sl.s; st.d = sl.d; st.delta = sl.delta; st.iteration = sl.iteration; st.active = sl.active != 0;
                k = sl.k; out_idx = (size_t)sl.out_idx;
#pragma unroll
                for (int c = 0; c < 7; c++) tot[c] = s_red[w][8 * sh + c];
                pts = (int)s_red[w][8 * sh + 7];
#pragma unroll
                for (int r = 0; r < 9; r++) vprev[r] = s_frames[w][sh * F_SLOTS + F_V + r];
            }
            InnerPrefix ip;
            ip.rr = nullptr; ip.neff = nullptr;
            ip.rr_up = use_inner ? s_rr : nullptr;
            ip.npairs16 = s_neff[ws];
            katz_step<SHELL>(st, tot, pts, P, vprev, s_frames[ws] + sh * F_SLOTS, valid, out_idx, valid, ip, k, &s_nin[ws][sh]);
            if (valid) {
                ShellSlot &sl = s_slot[w][sh];
                sl.q = st.q; sl.s = st.s; sl.iteration = st.iteration; sl.active = st.active ? 1 : 0;
            }
            const unsigned b = __ballot_sync(0xffffffffu, valid && st.active);
            if (lane < WB_WARPS) s_amask[lane] = (b >> (lane * S)) & ((1u << S) - 1u);
        }
        __syncthreads();   // next frames, inner-prefix lengths, active masks published
        if (have) {
            amask = s_amask[warp];
            if (amask == 0) have = false;
        }
    }
    ring_flush_stats(ring, P.stats, lane);
}

// ------------------------------------------------------------------------------------------------
// Kernel 1c: the warp kernel with an ASYNCHRONOUS step.  In kernel 1 the per-iteration step (reduction, eigen-solve,
// convergence test, next frame: 3.7 us, 28 % of a warp's time, a serial chain that keeps 4 of 32 lanes busy) sits in the
// dependency chain of every warp; kernel 1b shares it between the warps of a CTA but makes them wait for one another at a
// block barrier (slower).  Here the two halves of the Katz iteration are decoupled:
//   * WA_PASS "pass" warps per CTA only stream particles.  Each owns WA_SLOTS task slots (frames, shell state, a
//     WA_STAGES-tile ring per slot in shared memory).  After the pass of a slot the warp leaves the 8 S reduced sums in
//     shared memory, marks the slot POSTED and goes on with its other slot -- it never executes the step and never waits
//     for another pass warp;
//   * ONE "step" warp per CTA polls the slot states, takes up to 32 / S posted slots at a time -- one LANE per shell, so
//     a step costs the same chain for 32 shells as kernel 1 spends on 4 -- runs katz_step for them, publishes frames,
//     inner-prefix lengths and active masks and marks each slot READY (next pass) or EMPTY (task finished: the pass warp
//     pulls the next task from the queue into it).
// Per-shell arithmetic is that of kernel 1 (same run_pass, same reduction tree, same katz_step, same operands): results
// are identical bit for bit.  Slot states are plain shared-memory words, ordered with __threadfence_block().
// ------------------------------------------------------------------------------------------------
#ifndef CPG_WA_PASS
#define CPG_WA_PASS 5      // pass warps per CTA
#endif
#ifndef CPG_WA_SLOTS
#define CPG_WA_SLOTS 2     // task slots per pass warp
#endif
#ifndef CPG_WA_STAGES
#define CPG_WA_STAGES 2    // tiles (4 KB) in a slot's ring: tasks of <= 128 * WA_STAGES prefix particles stay resident
#endif
#ifndef CPG_WA_MINB
#define CPG_WA_MINB 2      // resident CTAs per SM the register allocator leaves room for
#endif
constexpr int WA_PASS = CPG_WA_PASS, WA_SLOTS = CPG_WA_SLOTS, WA_STAGES = CPG_WA_STAGES;
constexpr int WA_NSLOT = WA_PASS * WA_SLOTS;
constexpr int WA_THREADS = (WA_PASS + 1) * 32;
static_assert(WA_NSLOT <= 32, "one lane of the step warp watches one slot");
enum { WA_EMPTY = 0, WA_READY = 1, WA_POSTED = 2 };

struct SlotDesc {          // what a pass needs to know about its slot's task (shared memory; written by the pass warp)
    long long base, row0;
    int npairs_task, resident, res_loaded, pad;
    unsigned seq, res_seq;
};

template <int S, bool SHELL, bool REDUCED, bool VDISP>
__global__ void __launch_bounds__(WA_THREADS, CPG_WA_MINB) morph_wasync_kernel(const MorphParams P)
{
    using LM = LaneMap<S>;
    using RC = RingCfg<VDISP, WA_STAGES>;
    constexpr int NV = 8 * S;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char s_ring_a[];   // WA_NSLOT rings
    __shared__ __align__(16) double s_frames[WA_NSLOT][S * F_SLOTS];
    __shared__ __align__(8) uint64_t s_bar[WA_NSLOT][RING_MAX_STAGES];
    __shared__ double s_red[WA_NSLOT][NV];
    __shared__ ShellSlot s_slot[WA_NSLOT][S];
    __shared__ SlotDesc s_desc[WA_NSLOT];
    __shared__ int s_np[WA_NSLOT][S], s_nin[WA_NSLOT][S];
    __shared__ unsigned s_amask[WA_NSLOT];
    __shared__ int s_state[WA_NSLOT];
    __shared__ int s_done;
    __shared__ float s_rr[MAX_SORT_R];
    __shared__ uint16_t s_neff[WA_NSLOT][MAX_SORT_R];   // prefix lengths (pairs) of the slot's object
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    volatile int *state = s_state;
    const bool use_table = !SHELL && !REDUCED && P.n_eff != nullptr && P.ptab != nullptr;
    const bool use_inner = use_table || (SHELL && P.n_eff != nullptr);
    if (use_inner)
        for (int j = threadIdx.x; j < P.R; j += blockDim.x) s_rr[j] = __double2float_ru(P.rr[j]);
    if (threadIdx.x < WA_NSLOT) s_state[threadIdx.x] = WA_EMPTY;
    if (threadIdx.x == 0) s_done = 0;
    __syncthreads();
    const double *ptab = use_table ? P.ptab : nullptr;

    if (warp < WA_PASS) {
        // ---------------------------------------------------------------- pass warp
        const LM L(lane);
        WarpRing ring;
        for (int j = 0; j < WA_SLOTS; j++) {
            const int sl = warp * WA_SLOTS + j;
            ring_init(ring, s_ring_a + sl * RC::WARP_BYTES, s_bar[sl], lane, WARP_RING_LANES ? 32u : 1u);
            if (lane == 0) { s_desc[sl].seq = 0; s_desc[sl].res_seq = 0; s_desc[sl].res_loaded = 0; }
        }
        __syncwarp();
        bool queue_done = false;
        int cur = 0;
#ifdef CPG_TASK_TRACE
        long long tr_t0 = clock64(), tr_pass = 0, tr_idle = 0, tr_npass = 0, tr_nres = 0, tr_mark = tr_t0;
#endif
        for (;;) {
            int run = -1, nempty = 0;
            for (int jj = 0; jj < WA_SLOTS && run < 0; jj++) {
                const int j = (cur + jj) % WA_SLOTS, sl = warp * WA_SLOTS + j;
                const int stt = __shfl_sync(FULL, state[sl], 0);
                if (stt == WA_READY) {
                    run = sl;
                } else if (stt == WA_EMPTY) {
                    nempty++;
                    // next live task of the queue into this slot (dead ones -- empty object, r200 == 0 -- are skipped)
                    while (!queue_done && run < 0) {
                        int t = 0;
                        if (lane == 0) t = atomicAdd(P.task_counter, 1);
                        t = __shfl_sync(FULL, t, 0);
                        if (t >= P.n_tasks) { queue_done = true; break; }
                        const Task tk = P.tasks[t];
                        const int halo = tk.halo;
                        if (P.skip[halo] || (P.local && P.r200[halo] == 0.0)) continue;
                        const int N = P.soa.size[halo];
                        const int64_t base = P.soa.poff[halo];
                        KatzState st;
                        st.q = 1.0; st.s = 1.0; st.iteration = 1;
                        st.active = L.ls < tk.nshell;
                        const int k = tk.shell0 + (st.active ? L.ls : 0);
                        shell_geometry(P, halo, k, st.d, st.delta);
                        const size_t out_idx = (size_t)halo * P.R + k;
                        if (use_inner) {   // the object's prefix lengths up to the task's last shell
                            for (int j2 = lane; j2 < tk.shell0 + tk.nshell; j2 += 32)
                                s_neff[sl][j2] = (uint16_t)(P.n_eff[(size_t)halo * P.R + j2] >> 1);
                            __syncwarp();
                        }
                        InnerPrefix ip;
                        ip.rr = nullptr; ip.neff = nullptr;
                        ip.rr_up = use_inner ? s_rr : nullptr;
                        ip.npairs16 = s_neff[sl];
                        init_frames<SHELL>(s_frames[sl], s_np[sl], s_nin[sl], L, st, (P.n_eff && st.active) ? P.n_eff[out_idx] : N, ip, k);
                        if (L.writer) {
                            ShellSlot &ss = s_slot[sl][L.ls];
                            ss.q = 1.0; ss.s = 1.0; ss.d = st.d; ss.delta = st.delta; ss.out_idx = (unsigned long long)out_idx;
                            ss.iteration = 1; ss.active = st.active ? 1 : 0; ss.k = k;
                        }
                        __syncwarp();
                        const unsigned am = LM::active_mask(st.active && L.writer);
                        int npr0[S];
#pragma unroll
                        for (int s = 0; s < S; s++) npr0[s] = s_np[sl][s];
                        const int npt = active_npairs<S>(npr0, am);
                        if (lane == 0) {
                            s_amask[sl] = am;
                            SlotDesc &d = s_desc[sl];
                            d.base = base; d.row0 = (base >> 7) + halo; d.npairs_task = npt;
                            d.resident = ((npt + TILE_PAIRS - 1) / TILE_PAIRS <= WA_STAGES) ? 1 : 0;
                            d.res_loaded = 0;
                            s_state[sl] = WA_READY;
                        }
                        __syncwarp();
                        run = sl;
                    }
                }
            }
            if (run < 0) {
                if (queue_done && nempty == WA_SLOTS) break;
                __nanosleep(100);
                continue;
            }
#ifdef CPG_TASK_TRACE
            { const long long now = clock64(); tr_idle += now - tr_mark; tr_mark = now; }
#endif
            __threadfence_block();   // the step warp's frames, prefix lengths and mask of this slot are visible
            {
                const int sl = run;
                const unsigned amask = s_amask[sl];
                int npr[S];
#pragma unroll
                for (int s = 0; s < S; s++) npr[s] = s_np[sl][s];
                const SlotDesc d = s_desc[sl];
                ring.buf = s_ring_a + sl * RC::WARP_BYTES; ring.bar = s_bar[sl];
                ring.buf_s = smem_u32(ring.buf); ring.bar_s = smem_u32(ring.bar);
                ring.seq = d.seq; ring.res_seq = d.res_seq; ring.res_loaded = d.res_loaded != 0;
                const double tadd = table_term(ptab, d.row0, s_nin[sl][(lane & (NV - 1)) >> 3], lane & 7);
                double acc[S][7];
                int cnt[S];
                run_pass<S, SHELL, REDUCED, VDISP, WARP_RING_LANES, WA_STAGES>(P.soa, d.base, active_npairs<S>(npr, amask), 0, 1, lane, ring,
                                                                               d.resident != 0, d.npairs_task, s_frames[sl], npr, s_nin[sl],
                                                                               amask, acc, cnt);
                double val[NV];
                pack_values<S>(acc, cnt, val);
                const double mine = warp_sum_transpose<NV>(val, lane) + tadd;
                if (lane < NV) s_red[sl][lane] = mine;
                if (lane == 0) { s_desc[sl].seq = ring.seq; s_desc[sl].res_seq = ring.res_seq; s_desc[sl].res_loaded = ring.res_loaded ? 1 : 0; }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) state[sl] = WA_POSTED;
                cur = (sl - warp * WA_SLOTS + 1) % WA_SLOTS;
#ifdef CPG_TASK_TRACE
                { const long long now = clock64(); tr_pass += now - tr_mark; tr_mark = now; tr_npass++; tr_nres += d.res_loaded ? 1 : 0; }
#endif
            }
            if (ring.st_tiles > (1u << 20)) ring_flush_stats(ring, P.stats, lane);
        }
        ring_flush_stats(ring, P.stats, lane);
#ifdef CPG_TASK_TRACE
        if (lane == 0 && P.stats) {
            atomicAdd(P.stats + 2, (unsigned long long)(clock64() - tr_t0));
            atomicAdd(P.stats + 3, (unsigned long long)tr_pass);
            atomicAdd(P.stats + 4, (unsigned long long)tr_idle);   // (includes task fetch + init)
            atomicAdd(P.stats + 10, (unsigned long long)tr_npass);
            atomicAdd(P.stats + 11, (unsigned long long)tr_nres);
        }
#endif
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(&s_done, 1); }
    } else {
        // ---------------------------------------------------------------- step warp: one lane per shell
        volatile int *done = &s_done;
#ifdef CPG_TASK_TRACE
        long long tr_t0 = clock64(), tr_busy = 0, tr_batches = 0, tr_slots = 0;
#endif
        for (;;) {
            const int stt = lane < WA_NSLOT ? state[lane] : WA_EMPTY;
            const unsigned posted = __ballot_sync(FULL, stt == WA_POSTED);
            if (posted == 0) {
                if (*done >= WA_PASS) break;
                __nanosleep(40);
                continue;
            }
            __threadfence_block();   // the pass warps' sums are visible
#ifdef CPG_TASK_TRACE
            const long long tr_b0 = clock64();
#endif
            const int which = lane / S, sh = lane % S;
            const unsigned pos = __fns(posted, 0, which + 1);   // the (which + 1)-th posted slot
            const bool valid = pos != 0xffffffffu;
            const int sl = valid ? (int)pos : 0;                 // (keeps every lane's addresses in range)
            KatzState st;
            st.q = 1.0; st.s = 1.0; st.d = 1.0; st.delta = 1.0; st.iteration = 1; st.active = false;
            double tot[7] = {0, 0, 0, 0, 0, 0, 0}, vprev[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            int pts = 0, k = 0;
            size_t out_idx = 0;
            if (valid) {
                const ShellSlot &ss = s_slot[sl][sh];
                st.q = ss.q; st.s = ss.s; st.d = ss.d; st.delta = ss.delta; st.iteration = ss.iteration; st.active = ss.active != 0;
                k = ss.k; out_idx = (size_t)ss.out_idx;
#pragma unroll
                for (int c = 0; c < 7; c++) tot[c] = s_red[sl][8 * sh + c];
                pts = (int)s_red[sl][8 * sh + 7];
#pragma unroll
                for (int r = 0; r < 9; r++) vprev[r] = s_frames[sl][sh * F_SLOTS + F_V + r];
            }
            InnerPrefix ip;
            ip.rr = nullptr; ip.neff = nullptr;
            ip.rr_up = use_inner ? s_rr : nullptr;
            ip.npairs16 = s_neff[sl];
            katz_step<SHELL>(st, tot, pts, P, vprev, s_frames[sl] + sh * F_SLOTS, valid, out_idx, valid, ip, k, &s_nin[sl][sh]);
            if (valid) {
                ShellSlot &ss = s_slot[sl][sh];
                ss.q = st.q; ss.s = st.s; ss.iteration = st.iteration; ss.active = st.active ? 1 : 0;
            }
            const unsigned b = __ballot_sync(FULL, valid && st.active);
            const unsigned am = (b >> (which * S)) & ((1u << S) - 1u);
            __threadfence_block();   // frames, prefix lengths, shell state before the slot changes hands
            __syncwarp();
            if (valid && sh == 0) {
                s_amask[sl] = am;
                __threadfence_block();
                state[sl] = am ? WA_READY : WA_EMPTY;
            }
#ifdef CPG_TASK_TRACE
            tr_busy += clock64() - tr_b0; tr_batches++; tr_slots += min(__popc(posted), 32 / S);
#endif
        }
#ifdef CPG_TASK_TRACE
        if (lane == 0 && P.stats) {
            atomicAdd(P.stats + 6, (unsigned long long)(clock64() - tr_t0));
            atomicAdd(P.stats + 7, (unsigned long long)tr_busy);
            atomicAdd(P.stats + 8, (unsigned long long)tr_batches);
            atomicAdd(P.stats + 9, (unsigned long long)tr_slots);
        }
#endif
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 1d: a warp works on WD_SLOTS independent tasks at a time and runs ONE step for all of them.  Per round the
// warp passes over the particles of each of its live slots in turn (same run_pass as kernel 1, a WD_STAGES-tile ring per
// slot), keeps each slot's transposed sums in one register, and then executes katz_step once with one lane per shell
// for the WD_SLOTS * S shells of all slots: the serial chain of the step (3.7 us in kernel 1, 28 % of its time) is paid
// once per WD_SLOTS task-iterations, with no other warp to wait for (kernel 1b) and no warp set aside for stepping
// (kernel 1c).  A slot whose shells have all finished is refilled from the queue before the next round.  Per-shell
// arithmetic is that of kernel 1 (same pass, same reduction tree, same katz_step).
// ------------------------------------------------------------------------------------------------
#ifndef CPG_WD_WARPS
#define CPG_WD_WARPS 4
#endif
#ifndef CPG_WD_SLOTS
#define CPG_WD_SLOTS 2
#endif
#ifndef CPG_WD_STAGES
#define CPG_WD_STAGES 2
#endif
#ifndef CPG_WD_MINB
#define CPG_WD_MINB 3
#endif
constexpr int WD_WARPS = CPG_WD_WARPS, WD_SLOTS = CPG_WD_SLOTS, WD_STAGES = CPG_WD_STAGES;
constexpr int WD_NSLOT = WD_WARPS * WD_SLOTS;

struct DuoDesc {           // what a pass of kernel 1d needs to know about its slot's task
    long long base, row0;
    int npairs_task, resident, res_loaded, pad;
};

template <int S, bool SHELL, bool REDUCED, bool VDISP>
__global__ void __launch_bounds__(WD_WARPS * 32, CPG_WD_MINB) morph_wduo_kernel(const MorphParams P)
{
    using LM = LaneMap<S>;
    using RC = RingCfg<VDISP, RING_MAX_STAGES>;   // ONE ring of RING_MAX_STAGES tiles per warp, shared by its slots (run_pass_rt)
    static_assert(WD_SLOTS == 2 && RING_MAX_STAGES == 4, "two slots share a 4-stage ring");
    constexpr int NV = 8 * S;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(WD_SLOTS * S <= 32, "one lane per shell in the step");
    extern __shared__ __align__(128) unsigned char s_ring_d[];   // WD_WARPS rings
    __shared__ __align__(16) double s_frames[WD_NSLOT][S * F_SLOTS];
    __shared__ __align__(8) uint64_t s_bar[WD_WARPS][RING_MAX_STAGES];
    __shared__ ShellSlot s_slot[WD_NSLOT][S];
    __shared__ DuoDesc s_desc[WD_NSLOT];
    __shared__ int s_np[WD_NSLOT][S], s_nin[WD_NSLOT][S];
    __shared__ unsigned s_am[WD_NSLOT];   // active-shell mask of every slot (0: no task)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool use_table = !SHELL && !REDUCED && P.n_eff != nullptr && P.ptab != nullptr;
    const bool use_inner = use_table || (SHELL && P.n_eff != nullptr);
    // behind the rings (dynamic, sized by R so that three CTAs fit an SM for the usual R <= 100): the shell radii rounded
    // up to fp32 and the prefix lengths (pairs) of every slot's object, R rounded up to 8 entries each
    const int rpad = (P.R + 7) & ~7;
    float *s_rr = reinterpret_cast<float *>(s_ring_d + WD_WARPS * RC::WARP_BYTES);
    uint16_t *s_neff = reinterpret_cast<uint16_t *>(s_rr + rpad);
    if (use_inner) {
        for (int j = threadIdx.x; j < P.R; j += blockDim.x) s_rr[j] = __double2float_ru(P.rr[j]);
        __syncthreads();
    }
    const double *ptab = use_table ? P.ptab : nullptr;
    const LM L(lane);
    WarpRing ring;
    ring_init(ring, s_ring_d + warp * RC::WARP_BYTES, s_bar[warp], lane, 32u);
    unsigned phase = 0;   // next mbarrier phase of every stage of the ring
    __syncwarp();
    bool queue_done = false;
    if (lane < WD_SLOTS) s_am[warp * WD_SLOTS + lane] = 0;
    __syncwarp();
#ifdef CPG_TASK_TRACE
    long long tr_t0 = clock64(), tr_mark = tr_t0, tr_fill = 0, tr_pass = 0, tr_step = 0, tr_rounds = 0, tr_npass = 0, tr_nshell = 0;
#define CPG_TR(acc) { const long long now_ = clock64(); acc += now_ - tr_mark; tr_mark = now_; }
#else
#define CPG_TR(acc)
#endif

    for (;;) {
        // ---- refill the empty slots from the queue (dead tasks -- empty object, r200 == 0 -- are skipped)
        bool any = false;
#pragma unroll 1
        for (int j = 0; j < WD_SLOTS; j++) {
            const int sl = warp * WD_SLOTS + j;
            unsigned amj = s_am[sl];
            while (amj == 0 && !queue_done) {
                int t = 0;
                if (lane == 0) t = atomicAdd(P.task_counter, 1);
                t = __shfl_sync(FULL, t, 0);
                if (t >= P.n_tasks) { queue_done = true; break; }
                const Task tk = P.tasks[t];
                const int halo = tk.halo;
                if (P.skip[halo] || (P.local && P.r200[halo] == 0.0)) continue;
                const int N = P.soa.size[halo];
                const int64_t base = P.soa.poff[halo];
                KatzState st;
                st.q = 1.0; st.s = 1.0; st.iteration = 1;
                st.active = L.ls < tk.nshell;
                const int k = tk.shell0 + (st.active ? L.ls : 0);
                shell_geometry(P, halo, k, st.d, st.delta);
                const size_t out_idx = (size_t)halo * P.R + k;
                if (use_inner) {   // the object's prefix lengths up to the task's last shell
                    for (int j2 = lane; j2 < tk.shell0 + tk.nshell; j2 += 32)
                        s_neff[sl * rpad + j2] = (uint16_t)(P.n_eff[(size_t)halo * P.R + j2] >> 1);   // objects < 131072 particles
                    __syncwarp();
                }
                InnerPrefix ip;
                ip.rr = nullptr; ip.neff = nullptr;
                ip.rr_up = use_inner ? s_rr : nullptr;
                ip.npairs16 = s_neff + sl * rpad;
                init_frames<SHELL>(s_frames[sl], s_np[sl], s_nin[sl], L, st, (P.n_eff && st.active) ? P.n_eff[out_idx] : N, ip, k);
                if (L.writer) {
                    ShellSlot &ss = s_slot[sl][L.ls];
                    ss.q = 1.0; ss.s = 1.0; ss.d = st.d; ss.delta = st.delta; ss.out_idx = (unsigned long long)out_idx;
                    ss.iteration = 1; ss.active = st.active ? 1 : 0; ss.k = k;
                }
                __syncwarp();
                amj = LM::active_mask(st.active && L.writer);
                int npr0[S];
#pragma unroll
                for (int s = 0; s < S; s++) npr0[s] = s_np[sl][s];
                const int npt = active_npairs<S>(npr0, amj);
                if (lane == 0) {
                    s_am[sl] = amj;
                    DuoDesc &d = s_desc[sl];
                    d.base = base; d.row0 = (base >> 7) + halo; d.npairs_task = npt;
                    d.resident = ((npt + TILE_PAIRS - 1) / TILE_PAIRS <= RING_MAX_STAGES / WD_SLOTS) ? 1 : 0;
                    d.res_loaded = 0;
                }
                __syncwarp();
            }
            any = any || amj != 0;
        }
        if (!any) break;
        CPG_TR(tr_fill)

        // ---- one pass per live slot; each leaves its 8 S sums transposed in one register
        // (ONE copy of the pass in the instruction stream: the loop over the slots is not unrolled, the slots' sums are
        // steered into their registers with selects)
        double mine[WD_SLOTS];
#pragma unroll
        for (int j = 0; j < WD_SLOTS; j++) mine[j] = 0.0;
#pragma unroll 1
        for (int j = 0; j < WD_SLOTS; j++) {
            const int sl = warp * WD_SLOTS + j;
            const unsigned amj = s_am[sl];
            if (amj) {
                int npr[S];
#pragma unroll
                for (int s = 0; s < S; s++) npr[s] = s_np[sl][s];
                const DuoDesc d = s_desc[sl];
                // the other slot's resident tiles stay where they are: stream through this slot's half of the ring then
                const int so = warp * WD_SLOTS + (j ^ 1);
                const bool other_res = s_am[so] != 0 && s_desc[so].resident != 0;
                const bool whole = d.resident == 0 && !other_res;
                const int sbase = whole ? 0 : j * (RING_MAX_STAGES / WD_SLOTS), nst = whole ? RING_MAX_STAGES : RING_MAX_STAGES / WD_SLOTS;
                bool res_loaded = d.res_loaded != 0;
                const double tadd = table_term(ptab, d.row0, s_nin[sl][(lane & (NV - 1)) >> 3], lane & 7);
                double acc[S][7];
                int cnt[S];
                run_pass_rt<S, SHELL, REDUCED, VDISP>(P.soa, d.base, active_npairs<S>(npr, amj), lane, ring, phase, sbase, nst,
                                                      d.resident != 0, res_loaded, d.npairs_task, s_frames[sl], npr, s_nin[sl], amj, acc, cnt);
                if (lane == 0) s_desc[sl].res_loaded = res_loaded ? 1 : 0;
                double val[NV];
                pack_values<S>(acc, cnt, val);
                const double mj = warp_sum_transpose<NV>(val, lane) + tadd;
#pragma unroll
                for (int jj = 0; jj < WD_SLOTS; jj++) mine[jj] = (jj == j) ? mj : mine[jj];
            }
        }
        __syncwarp();
        CPG_TR(tr_pass)
#ifdef CPG_TASK_TRACE
        tr_rounds++;
        for (int j = 0; j < WD_SLOTS; j++) { const unsigned m_ = s_am[warp * WD_SLOTS + j]; tr_npass += m_ ? 1 : 0; tr_nshell += __popc(m_); }
#endif

        // ---- one step for the shells of all slots: lane l serves shell l % S of slot l / S
        {
            const int jl = (lane / S) % WD_SLOTS, sh = lane % S;
            const int sl = warp * WD_SLOTS + jl;
            const unsigned am_l = s_am[sl];
            double tot[7];
            int pts;
            {
                double got[8];
#pragma unroll
                for (int c = 0; c < 8; c++) got[c] = 0.0;
#pragma unroll
                for (int j = 0; j < WD_SLOTS; j++) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const double v = __shfl_sync(FULL, mine[j], (8 * sh + c) & 31);
                        if (j == jl) got[c] = v;
                    }
                }
#pragma unroll
                for (int c = 0; c < 7; c++) tot[c] = got[c];
                pts = (int)got[7];
            }
            const bool valid = lane < WD_SLOTS * S && am_l != 0;
            KatzState st;
            st.q = 1.0; st.s = 1.0; st.d = 1.0; st.delta = 1.0; st.iteration = 1; st.active = false;
            double vprev[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            int k = 0;
            size_t out_idx = 0;
            if (valid) {
                const ShellSlot &ss = s_slot[sl][sh];
                st.q = ss.q; st.s = ss.s; st.d = ss.d; st.delta = ss.delta; st.iteration = ss.iteration; st.active = ss.active != 0;
                k = ss.k; out_idx = (size_t)ss.out_idx;
#pragma unroll
                for (int r = 0; r < 9; r++) vprev[r] = s_frames[sl][sh * F_SLOTS + F_V + r];
            }
            InnerPrefix ip;
            ip.rr = nullptr; ip.neff = nullptr;
            ip.rr_up = use_inner ? s_rr : nullptr;
            ip.npairs16 = s_neff + sl * rpad;
            katz_step<SHELL>(st, tot, pts, P, vprev, s_frames[sl] + sh * F_SLOTS, valid, out_idx, valid, ip, k, &s_nin[sl][sh]);
            if (valid) {
                ShellSlot &ss = s_slot[sl][sh];
                ss.q = st.q; ss.s = st.s; ss.iteration = st.iteration; ss.active = st.active ? 1 : 0;
            }
            const unsigned b = __ballot_sync(FULL, valid && st.active);
            __syncwarp();
            if (lane < WD_SLOTS) s_am[warp * WD_SLOTS + lane] = (b >> (lane * S)) & ((1u << S) - 1u);
            __syncwarp();
        }
        CPG_TR(tr_step)
        if (ring.st_tiles > (1u << 20)) ring_flush_stats(ring, P.stats, lane);
    }
    ring_flush_stats(ring, P.stats, lane);
#ifdef CPG_TASK_TRACE
    if (lane == 0 && P.stats) {
        atomicAdd(P.stats + 2, (unsigned long long)(clock64() - tr_t0));
        atomicAdd(P.stats + 3, (unsigned long long)tr_pass);
        atomicAdd(P.stats + 4, (unsigned long long)tr_fill);
        atomicAdd(P.stats + 5, (unsigned long long)tr_step);
        atomicAdd(P.stats + 8, (unsigned long long)tr_rounds);
        atomicAdd(P.stats + 10, (unsigned long long)tr_npass);
        atomicAdd(P.stats + 12, (unsigned long long)tr_nshell);
    }
#endif
#undef CPG_TR
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: a thread-block CLUSTER per task, for objects too large for one warp.  The C CTAs of a
// cluster interleave over the object's particle pairs; per iteration every CTA publishes its S x 8
// partial sums in its own shared memory, one cluster barrier later every CTA reads all C partials
// through distributed shared memory in rank order, so all CTAs hold bit-identical totals and take the
// same convergence decisions without any global-memory traffic.  C == 1 degenerates to a plain CTA
// per task (no cluster barrier).  Partials are double-buffered: one barrier per iteration.
// ------------------------------------------------------------------------------------------------
#ifndef CPG_CLUSTER_WARPS
#define CPG_CLUSTER_WARPS 12
#endif
// 12 warps x 16 KB rings + statics = 207 KB of shared memory: one CTA per SM either way, but 12 resident warps instead of 8
// (the register file then allows 168 registers per thread)
constexpr int CLUSTER_KERNEL_WARPS = CPG_CLUSTER_WARPS;
#ifndef CPG_CLUSTER_RING_LANES
#define CPG_CLUSTER_RING_LANES 0   // 1: per-lane cp.async tile loads in the cluster kernel too; 0: elected-lane TMA bulk copies
#endif
constexpr bool CLUSTER_RING_LANES = CPG_CLUSTER_RING_LANES != 0;

template <int S, bool SHELL, bool REDUCED, bool VDISP>
__global__ void __launch_bounds__(CLUSTER_KERNEL_WARPS * 32) morph_cluster_kernel(const MorphParams P)
{
    constexpr int NW = CLUSTER_KERNEL_WARPS, NV = 8 * S;
    extern __shared__ __align__(128) unsigned char s_ring_c[];     // NW rings
    // ONE set of frames per CTA, written by warp 0 and published with a block barrier: the per-iteration step
    // (reduction, eigen-solve, next frame) costs the same time whether 1 or 32 lanes are active -- every warp of the CTA repeating it for private frames (round 1) was 12x the work.
    __shared__ __align__(16) double s_frames[S * F_SLOTS];
    __shared__ __align__(8) uint64_t s_bar[NW][RING_MAX_STAGES];
    __shared__ double s_part[2][NW][NV];                        // per-warp partial sums
    __shared__ double s_cta[2][NV];                             // per-CTA partial sums, read by the cluster
    __shared__ int s_np[S], s_nin[S];
    __shared__ unsigned s_amask;
    __shared__ double s_rr[MAX_SORT_R];
    __shared__ int s_neff[MAX_SORT_R];
    __shared__ ShellSlot s_slot[S];   // the Katz state of the task's shells between steps (CPG_PARK_STATE: not in registers)
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *frames = s_frames;
    int *np = s_np, *nin = s_nin;
    const LaneMap<S> L(lane);
    const int ls = L.ls;

    const int t = blockIdx.x / C;
    const Task tk = P.tasks[t];
    const int halo = tk.halo;
    if (P.skip[halo] || (P.local && P.r200[halo] == 0.0))
        return;   // uniform over the cluster
    const int N = P.soa.size[halo];
    const int64_t base = P.soa.poff[halo];
#ifdef CPG_TASK_TRACE
    unsigned long long trace_t0, trace_steps = 0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(trace_t0));
#endif
    const int tile0 = rank * NW + warp, tstride = C * NW;   // 64-pair tiles dealt round-robin to the cluster's warps
    WarpRing ring;
    ring_init(ring, s_ring_c + warp * RingCfg<VDISP>::WARP_BYTES, s_bar[warp], lane, CLUSTER_RING_LANES ? 32u : 1u);

    KatzState st;   // (kept by warp 0; the other warps only follow the published active mask)
    st.q = 1.0; st.s = 1.0; st.iteration = 1;
    st.active = ls < tk.nshell;
    const int k = tk.shell0 + (st.active ? ls : 0);
    shell_geometry(P, halo, k, st.d, st.delta);
    const size_t out_idx = (size_t)halo * P.R + k;
    const bool use_table = !SHELL && !REDUCED && P.n_eff != nullptr && P.ptab != nullptr;   // inner prefix from the prefix tensor table
    const bool use_inner = use_table || (SHELL && P.n_eff != nullptr);                        // ... or skipped (shell variant)
    const int64_t row0 = (base >> 7) + halo;
    if (use_inner) {
        for (int j = threadIdx.x; j < P.R; j += blockDim.x) {
            s_rr[j] = P.rr[j];
            s_neff[j] = P.n_eff[(size_t)halo * P.R + j];
        }
        __syncthreads();
    }
    InnerPrefix ip;
    ip.rr = use_inner ? s_rr : nullptr;
    ip.neff = s_neff;
    if (warp == 0) {
        init_frames<SHELL>(frames, np, nin, L, st, (P.n_eff && st.active) ? P.n_eff[out_idx] : N, ip, k);
        if (CPG_PARK_STATE && L.writer) {
            ShellSlot &sl = s_slot[ls];
            sl.q = 1.0; sl.s = 1.0; sl.d = st.d; sl.delta = st.delta; sl.out_idx = (unsigned long long)out_idx;
            sl.iteration = 1; sl.active = st.active ? 1 : 0; sl.k = k;
        }
        const unsigned am0 = LaneMap<S>::active_mask(st.active && L.writer);
        if (lane == 0) s_amask = am0;
    }
    __syncthreads();
    unsigned amask = s_amask;
    const bool out_writer = (rank == 0 && L.writer);   // (of warp 0)
    int npr[S];
#pragma unroll
    for (int s = 0; s < S; s++) npr[s] = np[s];
    const int npairs_task = active_npairs<S>(npr, amask);
    const int nt_task = (npairs_task + TILE_PAIRS - 1) / TILE_PAIRS;
    // uniform over the cluster's warps (a streamed pass deals the tiles from the first one some shell still tests,
    // a resident one from tile 0: the two dealings must not be mixed inside one task)
    const bool resident = (nt_task + tstride - 1) / tstride <= RingCfg<VDISP>::STAGES;

    double acc[S][7];
    int cnt[S];
    int buf = 0;
    run_pass<S, SHELL, REDUCED, VDISP, CLUSTER_RING_LANES>(P.soa, base, npairs_task, tile0, tstride, lane, ring, resident, npairs_task,
                                       frames, npr, nin, amask, acc, cnt);
    while (true) {
        {
            double val[NV];
            pack_values<S>(acc, cnt, val);
            const double mine = warp_sum_transpose<NV>(val, lane);
            if (lane < NV) s_part[buf][warp][lane] = mine;
        }
        __syncthreads();   // every warp's partial sums are in place (and every warp is done reading the frames)
        if (C > 1) {
            if (threadIdx.x < NV) {
                double v = 0.0;
                for (int w = 0; w < NW; w++) v += s_part[buf][w][threadIdx.x];
                s_cta[buf][threadIdx.x] = v;
            }
            cluster.sync();
        }
        if (warp == 0) {
            double tot[7] = {0, 0, 0, 0, 0, 0, 0}, vprev[9];
            double ptsd = 0.0;
            if (C == 1) {
#pragma unroll
                for (int w = 0; w < NW; w++) {
#pragma unroll
                    for (int c = 0; c < 7; c++) tot[c] += s_part[buf][w][8 * ls + c];
                    ptsd += s_part[buf][w][8 * ls + 7];
                }
            } else {
                // all CTAs read the C partials in rank order: bit-identical totals, identical decisions in every CTA
                for (int r = 0; r < C; r++) {
                    const double *remote = cluster.map_shared_rank(&s_cta[buf][8 * ls], r);
#pragma unroll
                    for (int c = 0; c < 7; c++) tot[c] += remote[c];
                    ptsd += remote[7];
                }
            }
            if (use_table) {   // the inner prefix of the pass just finished comes from the table
                const int pin = nin[ls];
#pragma unroll
                for (int c = 0; c < 6; c++) tot[c] += table_term(P.ptab, row0, pin, c);
                ptsd += table_term(P.ptab, row0, pin, 7);
            }
#pragma unroll
            for (int r = 0; r < 9; r++) vprev[r] = frames[ls * F_SLOTS + F_V + r];
            __syncwarp();
            unsigned am;
            if (CPG_PARK_STATE) {
                ShellSlot &sl = s_slot[ls];
                KatzState sv;
                sv.q = sl.q; sv.s = sl.s; sv.d = sl.d; sv.delta = sl.delta; sv.iteration = sl.iteration; sv.active = sl.active != 0;
                const int kk = sl.k;
                const size_t oidx = (size_t)sl.out_idx;
                __syncwarp();
                katz_step<SHELL>(sv, tot, (int)ptsd, P, vprev, frames + ls * F_SLOTS, L.writer, oidx, out_writer, ip, kk, nin + ls);
                if (L.writer) { sl.q = sv.q; sl.s = sv.s; sl.iteration = sv.iteration; sl.active = sv.active ? 1 : 0; }
                __syncwarp();
                am = LaneMap<S>::active_mask(sv.active && L.writer);
            } else {
                katz_step<SHELL>(st, tot, (int)ptsd, P, vprev, frames + ls * F_SLOTS, L.writer, out_idx, out_writer, ip, k,
                                 nin + ls);
                __syncwarp();
                am = LaneMap<S>::active_mask(st.active && L.writer);
            }
            if (lane == 0) s_amask = am;
        }
        __syncthreads();   // next frames, inner-prefix lengths and the active mask are published
        amask = s_amask;
        buf ^= 1;
#ifdef CPG_TASK_TRACE
        trace_steps++;
#endif
        if (amask == 0)
            break;
        run_pass<S, SHELL, REDUCED, VDISP, CLUSTER_RING_LANES>(P.soa, base, active_npairs<S>(npr, amask), tile0, tstride, lane, ring,
                                           resident, npairs_task, frames, npr, nin, amask, acc, cnt);
    }
#ifdef CPG_TASK_TRACE
    if (threadIdx.x == 0 && rank == 0 && P.stats && t < (1 << 19)) {
        unsigned long long t1; unsigned smid;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        const size_t o = 2 + 4 * ((size_t)(1 << 19) + t);
        P.stats[o] = trace_t0; P.stats[o + 1] = t1;
        P.stats[o + 2] = ((unsigned long long)smid << 32) | (unsigned)ring.st_tiles;
        P.stats[o + 3] = (trace_steps << 32) | (unsigned)N;
    }
#endif
    ring_flush_stats(ring, P.stats, lane);
    if (C > 1)
        cluster.sync();   // nobody leaves while a peer may still read its shared memory
}

// ------------------------------------------------------------------------------------------------
// Kernel 3: the whole GRID per task, for objects of millions of particles (BASELINE configs[4]: one 1e8-particle
// halo).  A cluster of 16 CTAs is the most one task can get from kernel 2, and a handful of such tasks leaves
// most of the GPU waiting for the slowest one.  Here every Katz iteration of a task is one launch of
// grid_pass_kernel over all SMs (each CTA publishes its 8S partial sums) followed by grid_step_kernel (one
// warp: fixed-order sum over the CTAs, eigen-solve, convergence test, next frames).  The iteration state
// lives in global memory between the launches; launches of a converged task return at once.
// ------------------------------------------------------------------------------------------------
struct GridState {
    double frames[4 * F_SLOTS];
    double q[4], s[4], d[4], delta[4];
    int iteration[4], active[4], np[4], nin[4];
    unsigned amask;
    int halo, shell0, nshell;
};

template <int S, bool SHELL>
__global__ void grid_init_kernel(const MorphParams P, const Task *tasks, GridState *states, int n_tasks)
{
    const int t = blockIdx.x;
    if (t >= n_tasks) return;
    const int lane = threadIdx.x & 31;
    const Task tk = tasks[t];
    GridState *gs = states + t;
    const LaneMap<S> L(lane);
    const int ls = L.ls, halo = tk.halo;
    const bool dead = P.skip[halo] || (P.local && P.r200[halo] == 0.0);
    KatzState st;
    st.q = 1.0; st.s = 1.0; st.iteration = 1;
    st.active = !dead && ls < tk.nshell;
    const int k = tk.shell0 + (ls < tk.nshell ? ls : 0);
    shell_geometry(P, halo, k, st.d, st.delta);
    const size_t out_idx = (size_t)halo * P.R + k;
    InnerPrefix ip;   // (the host passes ptab only for the non-reduced ellipsoid variant on radially ordered particles)
    ip.rr = (P.n_eff && (P.ptab || SHELL)) ? P.rr : nullptr;
    ip.neff = P.n_eff ? P.n_eff + (size_t)halo * P.R : nullptr;
    init_frames<SHELL>(gs->frames, gs->np, gs->nin, L, st, (P.n_eff && st.active) ? P.n_eff[out_idx] : P.soa.size[halo], ip, k);
    if (L.writer) {
        gs->q[ls] = 1.0; gs->s[ls] = 1.0; gs->d[ls] = st.d; gs->delta[ls] = st.delta;
        gs->iteration[ls] = 1; gs->active[ls] = st.active ? 1 : 0;
    }
    const unsigned amask = LaneMap<S>::active_mask(st.active && L.writer);
    if (lane == 0) { gs->amask = amask; gs->halo = halo; gs->shell0 = tk.shell0; gs->nshell = tk.nshell; }
}

constexpr int GRID_KERNEL_WARPS = 12;

template <int S, bool SHELL, bool REDUCED, bool VDISP>
__global__ void __launch_bounds__(GRID_KERNEL_WARPS * 32) grid_pass_kernel(const MorphParams P, const GridState *gs,
                                                                          double *partials /* [gridDim.x][8S] */)
{
    constexpr int NW = GRID_KERNEL_WARPS, NV = 8 * S;
    extern __shared__ __align__(128) unsigned char s_ring_g[];
    __shared__ __align__(16) double s_frames[S * F_SLOTS];
    __shared__ __align__(8) uint64_t s_bar[NW][RING_MAX_STAGES];
    __shared__ double s_part[NW][NV];
    __shared__ int s_nin[S];
    const unsigned amask = gs->amask;
    if (amask == 0) return;   // converged (uniform over the grid)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < S * F_SLOTS; i += blockDim.x) s_frames[i] = gs->frames[i];
    if (threadIdx.x < S) s_nin[threadIdx.x] = ((!SHELL && !REDUCED && P.ptab) || (SHELL && P.n_eff)) ? gs->nin[threadIdx.x] : 0;
    WarpRing ring;
    ring_init(ring, s_ring_g + warp * RingCfg<VDISP>::WARP_BYTES, s_bar[warp], lane);
    __syncthreads();
    int npr[S];
#pragma unroll
    for (int s = 0; s < S; s++) npr[s] = gs->np[s];
    const int halo = gs->halo;
    double acc[S][7];
    int cnt[S];
    run_pass<S, SHELL, REDUCED, VDISP>(P.soa, P.soa.poff[halo], active_npairs<S>(npr, amask), blockIdx.x * NW + warp,
                                       gridDim.x * NW, lane, ring, false, 0, s_frames, npr, s_nin, amask, acc, cnt);
    ring_flush_stats(ring, P.stats, lane);
    double val[NV];
    pack_values<S>(acc, cnt, val);
    const double mine = warp_sum_transpose<NV>(val, lane);
    if (lane < NV) s_part[warp][lane] = mine;
    __syncthreads();
    if (threadIdx.x < NV) {
        double v = 0.0;
        for (int w = 0; w < NW; w++) v += s_part[w][threadIdx.x];
        partials[(size_t)blockIdx.x * NV + threadIdx.x] = v;
    }
}

template <int S, bool SHELL>
__global__ void grid_step_kernel(const MorphParams P, GridState *gs, const double *partials, int n_cta)
{
    constexpr int NV = 8 * S;
    if (gs->amask == 0) return;
    const int lane = threadIdx.x & 31;
    const LaneMap<S> L(lane);
    const int ls = L.ls;
    double mine = 0.0;
    for (int b = 0; b < n_cta; b++) mine += partials[(size_t)b * NV + (lane & (NV - 1))];   // fixed order
    if (!SHELL && P.n_eff && P.ptab)   // inner prefix of the pass just finished
        mine += table_term(P.ptab, (P.soa.poff[gs->halo] >> 7) + gs->halo, gs->nin[(lane & (NV - 1)) >> 3], lane & 7);
    double tot[7], vprev[9];
#pragma unroll
    for (int c = 0; c < 7; c++) tot[c] = __shfl_sync(0xffffffffu, mine, (lane & ~7) + c);
    const int pts = (int)__shfl_sync(0xffffffffu, mine, (lane & ~7) + 7);
    KatzState st;
    st.q = gs->q[ls]; st.s = gs->s[ls]; st.d = gs->d[ls]; st.delta = gs->delta[ls];
    st.iteration = gs->iteration[ls]; st.active = gs->active[ls] != 0;
    const int k = gs->shell0 + (ls < gs->nshell ? ls : 0);
    const size_t out_idx = (size_t)gs->halo * P.R + k;
#pragma unroll
    for (int r = 0; r < 9; r++) vprev[r] = gs->frames[ls * F_SLOTS + F_V + r];
    InnerPrefix ip;
    ip.rr = (P.n_eff && (P.ptab || SHELL)) ? P.rr : nullptr;
    ip.neff = P.n_eff ? P.n_eff + (size_t)gs->halo * P.R : nullptr;
    __syncwarp();
    katz_step<SHELL>(st, tot, pts, P, vprev, gs->frames + ls * F_SLOTS, L.writer, out_idx, L.writer, ip, k, gs->nin + ls);
    __syncwarp();
    if (L.writer) { gs->q[ls] = st.q; gs->s[ls] = st.s; gs->iteration[ls] = st.iteration; gs->active[ls] = st.active ? 1 : 0; }
    const unsigned amask = LaneMap<S>::active_mask(st.active && L.writer);
    if (lane == 0) gs->amask = amask;
}

}  // namespace cpg
