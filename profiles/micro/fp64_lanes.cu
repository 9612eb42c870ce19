// Microbenchmark: does the fp64 pipe of sm_100a spend time on inactive lanes?  Independent DFMA chains executed by
// all 32 lanes of every warp, by lanes 0-15 only, 0-7, 0-3, lane 0 (the others branch around the loop), 12 warps/SM.
// If a warp instruction whose upper half-warp is inactive costs half the pipe time, the per-iteration Katz step of
// the morphology kernels (4 shells per warp = 4 useful lanes, today replicated over 32) should run in lanes 0-15.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lanes fp64_lanes.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b, int active)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    if ((threadIdx.x & 31) < active) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// lanes active in a strided pattern: lane % stride == 0 (e.g. one lane per group of 8)
template <int ILP>
__global__ void ks(double *out, int iters, double a, double b, int stride)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    if (((threadIdx.x & 31) % stride) == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double clk = p.clockRate * 1e-6;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 512);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    printf("%s SMs %d clock %.3f GHz; 12 warps/SM, 8 chains per thread, %d iterations\n", p.name, sms, clk, iters);
    const int act[] = {32, 16, 8, 4, 1};
    for (int a : act) {
        k<8><<<sms, 384>>>(out, 100, 1.0000001, 1e-9, a);
        cudaEventRecord(e0);
        k<8><<<sms, 384>>>(out, iters, 1.0000001, 1e-9, a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double winst = (double)sms * 12 * 8 * iters;   // warp instructions
        printf("lanes 0..%2d active: %.3f ms  %.3f warp-DFMA/clk/SMSP\n", a - 1, ms, winst / (ms * 1e-3) / (clk * 1e9) / sms / 4);
    }
    const int str[] = {2, 4, 8, 16};
    for (int s : str) {
        ks<8><<<sms, 384>>>(out, 100, 1.0000001, 1e-9, s);
        cudaEventRecord(e0);
        ks<8><<<sms, 384>>>(out, iters, 1.0000001, 1e-9, s);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double winst = (double)sms * 12 * 8 * iters;
        printf("every %2d-th lane active: %.3f ms  %.3f warp-DFMA/clk/SMSP\n", s, ms, winst / (ms * 1e-3) / (clk * 1e9) / sms / 4);
    }
    return 0;
}
