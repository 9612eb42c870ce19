// Microbenchmark: FP64 FMA throughput / latency on the target GPU (used to set the compute roofline of the
// morphology kernels in DESIGN.md).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_bench dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, int sms, double clk_ghz)
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 2048 * 1024);
    const int iters = 20000;
    dim3 grid(sms), block(warps_per_sm * 32);
    k<ILP><<<grid, block>>>(out, 100, 1.0000001, 1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP><<<grid, block>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)sms * warps_per_sm * 32 * ILP * iters;
    double per_clk_sm = fmas / (ms * 1e-3) / (clk_ghz * 1e9) / sms;
    printf("ILP %2d warps/SM %2d: %.3f ms  %.2f TFMA/s  %.1f DFMA/clk/SM (at %.3f GHz)  chain step %.1f clk\n", ILP, warps_per_sm, ms,
           fmas / (ms * 1e-3) / 1e12, per_clk_sm, clk_ghz, (ms * 1e-3) * clk_ghz * 1e9 / iters / 1.0);
    cudaFree(out);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double clk = p.clockRate * 1e-6;
    printf("%s SMs %d clock %.3f GHz\n", p.name, sms, clk);
    for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w, sms, clk); }
    for (int w : {4, 8, 12, 16}) { run<4>(w, sms, clk); }
    for (int w : {4, 8, 12}) { run<8>(w, sms, clk); }
    return 0;
}
