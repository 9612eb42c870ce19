// Loop lab: the streaming-loop body of the morphology kernels (cpg_morph.cuh: ring_read_pair + passk_*) in isolation,
// on tiles that sit in shared memory, so that variants of it can be timed per tile with 1..3 warps per SM sub-partition
// and without the per-iteration step, the task queue or any global-memory traffic.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cosmic_profiles_b200/csrc -I../../include [-D...] -o loop_lab loop_lab.cu
#include <cstdio>
#include "cpg_morph.cuh"
using namespace cpg;

#ifndef LAB_VARIANT
#define LAB_VARIANT 0   // 0: passk_particles (one shell at a time), 1: passk_shell_pairs
#endif

template <int S>
__global__ void __launch_bounds__(384, 1) lab(double *out, long long *cyc, int tiles, unsigned smask, double d2scale, SoA gsoa)
{
    extern __shared__ __align__(128) unsigned char s_ring[];
    __shared__ __align__(16) double s_frames[12][S * F_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpRing ring;
    ring.buf = s_ring + warp * RingCfg<false>::WARP_BYTES;
    // synthetic particles: |x| in (0.5, 1.5)
    for (int st = 0; st < RingCfg<false>::STAGES; st++)
        for (int pi = lane; pi < TILE_PAIRS; pi += 32)
            for (int s4 = 0; s4 < 4; s4++) {
                double2 v;
                const double a = 0.37 * (pi + 1) + 1.3 * st + 0.11 * s4, b = a + 0.5;
                v.x = s4 < 3 ? 0.9 * sin(a * (s4 + 1)) : 0.01;
                v.y = s4 < 3 ? 0.9 * cos(b * (s4 + 2)) : 0.012;
                reinterpret_cast<double2 *>(ring.buf + st * RingCfg<false>::STAGE_BYTES + s4 * RingCfg<false>::STREAM_BYTES)[pi] = v;
            }
    if (lane < S) {
        double *F = s_frames[warp] + lane * F_SLOTS;
        for (int r = 0; r < F_SLOTS; r++) F[r] = 0.0;
        F[F_AXX] = 1.0; F[F_AYY] = 1.7 + 0.1 * lane; F[F_AZZ] = 3.1; F[F_AXY] = 0.2; F[F_AXZ] = -0.1; F[F_AYZ] = 0.05;
        F[F_D2] = d2scale * (1.0 + 0.1 * lane);
        F[F_THR] = -1.0;
    }
    __syncthreads();
    const double *frames = s_frames[warp];
    double acc[S][7];
    int cnt[S];
#pragma unroll
    for (int s = 0; s < S; s++) { cnt[s] = 0;
#pragma unroll
        for (int c = 0; c < 7; c++) acc[s][c] = 0.0; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < (LAB_VARIANT >= 2 ? 0 : tiles); it++) {
        unsigned m = smask;
        asm volatile("" : "+r"(m));
        Particle p[4];
        ring_read_pair<false>(ring, (uint32_t)it, lane, p[0], p[1]);
        ring_read_pair<false>(ring, (uint32_t)it, lane + 32, p[2], p[3]);
#if LAB_VARIANT == 0
        passk_particles<S, false, false, false, 4>(p, frames, m, acc, cnt);
#elif LAB_VARIANT == 1
        passk_shell_pairs<S, 4>(p, frames, m, acc, cnt);
#endif
    }
#if LAB_VARIANT == 3
    // the real pass in streaming mode: a 64-tile object read from global memory (L2) through the cp.async ring
    {
        __shared__ int s_nin[12][S];
        __shared__ __align__(8) uint64_t s_bar[12][RING_MAX_STAGES];
        if (lane < S) s_nin[warp][lane] = 0;
        __syncwarp();
        ring_init(ring, ring.buf, s_bar[warp], lane, 32u);
        const int NT = 64;
        int npr[S];
#pragma unroll
        for (int s = 0; s < S; s++) npr[s] = NT * TILE_PAIRS;
#pragma unroll 1
        for (int it = 0; it < tiles / NT; it++) {
            unsigned m = smask;
            asm volatile("" : "+r"(m));
            double a2[S][7]; int c2[S];
            run_pass<S, false, false, false, true>(gsoa, (int64_t)(blockIdx.x % 8) * NT * 128, NT * TILE_PAIRS, 0, 1, lane, ring, false, NT * TILE_PAIRS, frames, npr, s_nin[warp], m, a2, c2);
#pragma unroll
            for (int s = 0; s < S; s++) { cnt[s] += c2[s];
#pragma unroll
                for (int c = 0; c < 7; c++) acc[s][c] += a2[s][c]; }
        }
    }
#endif
#if LAB_VARIANT == 5
    // streaming mode from global memory, sums to shared memory (no extra live registers across run_pass)
    {
        __shared__ int s_nin[12][S];
        __shared__ double s_tot[12][32];
        __shared__ __align__(8) uint64_t s_bar[12][RING_MAX_STAGES];
        if (lane < S) s_nin[warp][lane] = 0;
        s_tot[warp][lane] = 0.0;
        __syncwarp();
        ring_init(ring, ring.buf, s_bar[warp], lane, 32u);
        const int NT = 64;
        int npr[S];
#pragma unroll
        for (int s = 0; s < S; s++) npr[s] = NT * TILE_PAIRS;
#pragma unroll 1
        for (int it = 0; it < tiles / NT; it++) {
            unsigned m = smask;
            asm volatile("" : "+r"(m));
            run_pass<S, false, false, false, true>(gsoa, (int64_t)(blockIdx.x % 8) * NT * 128, NT * TILE_PAIRS, 0, 1, lane, ring, false, NT * TILE_PAIRS, frames, npr, s_nin[warp], m, acc, cnt);
            double v = 0.0;
#pragma unroll
            for (int s = 0; s < S; s++) { v += cnt[s];
#pragma unroll
                for (int c = 0; c < 7; c++) v += acc[s][c]; }
            s_tot[warp][lane] += v;
            __syncwarp();
        }
        acc[0][0] = s_tot[warp][lane];
    }
#endif
#if LAB_VARIANT == 4
    {
        __shared__ int s_nin[12][S];
        __shared__ double s_tot[12][32][2];
        if (lane < S) s_nin[warp][lane] = 0;
        s_tot[warp][lane][0] = 0.0; s_tot[warp][lane][1] = 0.0;
        __syncwarp();
        ring.seq = 0; ring.res_seq = 0; ring.res_loaded = true; ring.st_tiles = 0; ring.st_tshells = 0;
        ring.buf_s = smem_u32(ring.buf); ring.bar = nullptr; ring.bar_s = 0;
        int npr[S];
#pragma unroll
        for (int s = 0; s < S; s++) npr[s] = 4 * TILE_PAIRS;
        SoA soa = {};
#pragma unroll 1
        for (int it = 0; it < tiles / 4; it++) {
            unsigned m = smask;
            asm volatile("" : "+r"(m));
            run_pass<S, false, false, false, true>(soa, 0, 4 * TILE_PAIRS, 0, 1, lane, ring, true, 4 * TILE_PAIRS, frames, npr, s_nin[warp], m, acc, cnt);
            double v = 0.0;
#pragma unroll
            for (int s = 0; s < S; s++) { v += cnt[s];
#pragma unroll
                for (int c = 0; c < 7; c++) v += acc[s][c]; }
            s_tot[warp][lane][0] += v;
            __syncwarp();
        }
        acc[0][0] = s_tot[warp][lane][0];
    }
#endif
#if LAB_VARIANT == 2
    // the real pass over a task that is resident in the ring (4 tiles): run_pass's own tile loop, no loads, no barriers
    {
        __shared__ int s_nin[12][S];
        if (lane < S) s_nin[warp][lane] = 0;
        __syncwarp();
        ring.seq = 0; ring.res_seq = 0; ring.res_loaded = true; ring.st_tiles = 0; ring.st_tshells = 0;
        ring.buf_s = smem_u32(ring.buf); ring.bar = nullptr; ring.bar_s = 0;
        int npr[S];
#pragma unroll
        for (int s = 0; s < S; s++) npr[s] = 4 * TILE_PAIRS;
        SoA soa = {};
        double tot[S][7];
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
            for (int c = 0; c < 7; c++) tot[s][c] = 0.0;
#pragma unroll 1
        for (int it = 0; it < tiles / 4; it++) {
            unsigned m = smask;
            asm volatile("" : "+r"(m));
            run_pass<S, false, false, false, true>(soa, 0, 4 * TILE_PAIRS, 0, 1, lane, ring, true, 4 * TILE_PAIRS, frames, npr, s_nin[warp], m, acc, cnt);
#pragma unroll
            for (int s = 0; s < S; s++)
#pragma unroll
                for (int c = 0; c < 7; c++) tot[s][c] += acc[s][c];
        }
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
            for (int c = 0; c < 7; c++) acc[s][c] = tot[s][c];
    }
#endif
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < S; k++) { s += cnt[k];
#pragma unroll
        for (int c = 0; c < 6; c++) s += acc[k][c]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0) cyc[blockIdx.x * 12 + warp] = t1 - t0;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * 384); cudaMalloc(&cyc, sizeof(long long) * sms * 12);
    const int smem = 12 * RingCfg<false>::WARP_BYTES;
    SoA gsoa = {};
    {
        const size_t n = 8 * 64 * 128 + 256;
        double *h = (double *)malloc(sizeof(double) * n), *d[4];
        for (int st = 0; st < 4; st++) {
            for (size_t i = 0; i < n; i++) h[i] = st < 3 ? 0.9 * sin(0.37 * (i % 977 + 1) * (st + 1)) : 0.01;
            cudaMalloc(&d[st], sizeof(double) * n); cudaMemcpy(d[st], h, sizeof(double) * n, cudaMemcpyHostToDevice);
        }
        gsoa.dx = d[0]; gsoa.dy = d[1]; gsoa.dz = d[2]; gsoa.m = d[3];
    }
    cudaFuncSetAttribute(lab<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int tiles = 4096;
    printf("%s variant %d split %d: cycles per 128-particle tile and warp (4 shells; 172 fp64 warp instructions at 2.85 shells, 232 at 4)\n", p.name, LAB_VARIANT, CPG_SPLIT_FORM);
    const unsigned masks[] = {0xF, 0x7, 0x3, 0x1};
    const double d2s[] = {1.6, 1e-9, 1e9};   // about half of the particles selected / none / all
    for (double d2 : d2s)
        for (unsigned m : masks) {
            printf("d2 %8.1e mask %x:", d2, m);
            for (int w = 1; w <= 3; w++) {
                lab<4><<<sms, 128 * w, smem>>>(out, cyc, tiles, m, d2, gsoa);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf(" error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[12];
                cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
                double mean = 0; for (int i = 0; i < 4 * w; i++) mean += (double)h[i] / (4 * w);
                printf("   %d w/SMSP %7.1f", w, mean / tiles);
            }
            printf("\n");
        }
    return 0;
}
