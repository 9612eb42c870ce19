// Microbenchmark: fp64 tensor-core MMA (mma.sync m8n8k4 f64, SASS DMMA) on sm_100a: its rate alone, the DFMA rate alone,
// and both interleaved in the same warps.  If the interleaved time is close to the larger of the two the pipes are
// independent, and part of the morphology loop's fp64 work (the S x 6 tensor updates are a small GEMM over the particles
// of a tile) could move off the saturated DFMA pipe; if it is close to their sum they share the units.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void k(double *out, int iters, double a, double b)
{
    double c[NM > 0 ? NM : 1][2], x[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); i++) x[i] = threadIdx.x * 1e-3 + i;
    const double fa = 1.0 + a * 1e-9, fb = b * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) dmma(c[i], fa, fb);
#pragma unroll
        for (int i = 0; i < NF; i++) x[i] = fma(x[i], fa, fb);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF>
float run(double *out, int sms, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NM, NF><<<sms, 384>>>(out, 100, 1.0, 1.0);
    cudaEventRecord(e0);
    k<NM, NF><<<sms, 384>>>(out, iters, 1.0, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double clk = p.clockRate * 1e-6;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 512);
    const int iters = 20000;
    const double warps = (double)sms * 12;
    printf("%s SMs %d clock %.3f GHz; 12 warps/SM, %d iterations\n", p.name, sms, clk, iters);
    float t;
    t = run<0, 8>(out, sms, iters);
    printf("8 DFMA chains            : %.3f ms  %.3f warp-DFMA/clk/SMSP  %.1f TFMA/s\n", t, warps * 8 * iters / (t * 1e-3) / (clk * 1e9) / sms / 4, warps * 8 * iters * 32 / (t * 1e-3) / 1e12);
    t = run<8, 0>(out, sms, iters);
    printf("8 DMMA chains (m8n8k4)   : %.3f ms  %.3f warp-DMMA/clk/SMSP  %.1f TFMA/s (256 FMA per DMMA)\n", t, warps * 8 * iters / (t * 1e-3) / (clk * 1e9) / sms / 4, warps * 8 * iters * 256 / (t * 1e-3) / 1e12);
    t = run<4, 0>(out, sms, iters);
    printf("4 DMMA chains            : %.3f ms  %.3f warp-DMMA/clk/SMSP\n", t, warps * 4 * iters / (t * 1e-3) / (clk * 1e9) / sms / 4);
    t = run<8, 8>(out, sms, iters);
    printf("8 DMMA + 8 DFMA together : %.3f ms\n", t);
    t = run<4, 8>(out, sms, iters);
    printf("4 DMMA + 8 DFMA together : %.3f ms\n", t);
    t = run<2, 8>(out, sms, iters);
    printf("2 DMMA + 8 DFMA together : %.3f ms\n", t);
    t = run<1, 8>(out, sms, iters);
    printf("1 DMMA + 8 DFMA together : %.3f ms\n", t);
    return 0;
}
