// Microbenchmark: DFMA issue cadence and latency on sm_100a for the operand patterns of the morphology streaming loop.
//   kc: x = fma(x, a, b) with a, b in the constant bank (what fp64_lanes.cu measures: the ideal cadence)
//   kr: x[i] = fma(y[i], z[j], x[i]) with three distinct register operands per instruction (the loop's pattern)
//   kp: the same, predicated on a per-lane data-dependent predicate (the tensor updates)
// for ILP = 1, 2, 4, 8 independent chains per thread and 1, 2, 3, 4 warps per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_mix dfma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int MODE>
__global__ void k(double *out, int iters, double a, double b, const double *in)
{
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x * 1e-3 + i; y[i] = in[i] + 1e-9 * threadIdx.x; z[i] = in[8 + i]; }
    const bool pred = in[16 + (threadIdx.x & 7)] > 0.5;   // true for all lanes at run time, unknown to the compiler
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) x[i] = fma(x[i], a, b);
            else if (MODE == 1) x[i] = fma(y[i], z[(i + 1) % ILP], x[i]);
            else { if (pred) x[i] = fma(y[i], z[(i + 1) % ILP], x[i]); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int MODE>
void run(const char *name, int sms, double clk, double *out, const double *in)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    printf("%-28s ILP %d:", name, ILP);
    for (int wps = 1; wps <= 4; wps++) {
        k<ILP, MODE><<<sms, 128 * wps>>>(out, 100, 1.0000001, 1e-9, in);
        cudaEventRecord(e0);
        k<ILP, MODE><<<sms, 128 * wps>>>(out, iters, 1.0000001, 1e-9, in);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double winst = (double)sms * 4 * wps * ILP * iters;
        printf("  %d w/SMSP %.3f", wps, winst / (ms * 1e-3) / (clk * 1e9) / sms / 4);
    }
    printf("   warp-DFMA/clk/SMSP\n");
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double clk = p.clockRate * 1e-6;
    double *out, *in; cudaMalloc(&out, sizeof(double) * sms * 512); cudaMalloc(&in, sizeof(double) * 32);
    double h[32]; for (int i = 0; i < 32; i++) h[i] = 1.0 + 1e-7 * i;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    printf("%s SMs %d clock %.3f GHz\n", p.name, sms, clk);
    run<1, 0>("constant-bank operands", sms, clk, out, in); run<2, 0>("constant-bank operands", sms, clk, out, in);
    run<4, 0>("constant-bank operands", sms, clk, out, in); run<8, 0>("constant-bank operands", sms, clk, out, in);
    run<1, 1>("three register operands", sms, clk, out, in); run<2, 1>("three register operands", sms, clk, out, in);
    run<4, 1>("three register operands", sms, clk, out, in); run<8, 1>("three register operands", sms, clk, out, in);
    run<1, 2>("predicated, three registers", sms, clk, out, in); run<2, 2>("predicated, three registers", sms, clk, out, in);
    run<4, 2>("predicated, three registers", sms, clk, out, in); run<8, 2>("predicated, three registers", sms, clk, out, in);
    return 0;
}
