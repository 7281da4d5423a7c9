// micro-benchmark: DFMA throughput vs warps/SM and ILP (development tool)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(int iters, double a, double b, double *sink)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456) sink[0] = s;
}
template <int ILP>
void run(int warps_per_sm, int sms, double *sink)
{
    // one CTA per SM with warps_per_sm warps
    int threads = warps_per_sm * 32, ctas = sms;
    if (threads > 1024) { ctas = sms * (threads / 1024); threads = 1024; }
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<ctas, threads>>>(100, 0.999999, 1e-9, sink);
    cudaEventRecord(e0);
    k<ILP><<<ctas, threads>>>(iters, 0.999999, 1e-9, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)ctas * threads * iters * 8.0 * ILP;
    printf("warps/SM %2d ILP %d : %6.2f TFLOP/s (%5.1f%% of 37.2)\n", warps_per_sm, ILP, 2 * ops / ms / 1e9, 2 * ops / ms / 1e9 / 37.22 * 100);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink; cudaMalloc(&sink, 8);
    for (int w : {4, 8, 16, 32, 64}) { run<1>(w, p.multiProcessorCount, sink); run<2>(w, p.multiProcessorCount, sink); run<4>(w, p.multiProcessorCount, sink); run<8>(w, p.multiProcessorCount, sink); }
    return 0;
}
