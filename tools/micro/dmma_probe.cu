// micro-benchmark: FP64 tensor-core MMA (mma.sync m8n8k4 f64, SASS DMMA) on B200 -- its rate
// relative to the FP64 pipe (DFMA) and whether the two run concurrently (development tool).
//   ND DFMAs and NM DMMAs per loop iteration, all chains independent (8 DFMA chains, 4 DMMA
//   accumulator pairs per thread).  Reported: time per iteration pattern and the FMA rate.
// If DMMA is a separate unit, t(ND, NM) ~ max(t(ND, 0), t(0, NM)); if it shares the FP64 pipe,
// t(ND, NM) ~ t(ND, 0) + t(0, NM).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int ND, int NM>
__global__ void __launch_bounds__(256) k(int iters, double a, double b, double *sink)
{
    double x[8], c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < ND; ++i) x[(r * ND + i) & 7] = fma(x[(r * ND + i) & 7], a, b);
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma(c0[i & 3], c1[i & 3], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
    if (s == 123.456) sink[0] = s;
}

template <int ND, int NM>
void run(int sms, int warps_per_sm, double *sink)
{
    const int threads = 256, ctas = sms * (warps_per_sm / 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ND, NM><<<ctas, threads>>>(100, 0.999999, 1e-9, sink);
    cudaEventRecord(e0);
    k<ND, NM><<<ctas, threads>>>(iters, 0.999999, 1e-9, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)ctas * threads / 32;
    const double dfma = warps * 32.0 * iters * 4.0 * ND;          // scalar FMAs
    const double dm = warps * 256.0 * iters * 4.0 * NM;           // 8x8x4 FMAs per warp-wide DMMA
    printf("warps/SM %2d  DFMA x%d  DMMA x%d : %8.3f ms   DFMA %6.2f TFLOP/s   DMMA %6.2f TFLOP/s   sum %6.2f\n", warps_per_sm, ND, NM, ms,
           2 * dfma / ms / 1e9, 2 * dm / ms / 1e9, 2 * (dfma + dm) / ms / 1e9);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink; cudaMalloc(&sink, 8);
    const int sms = p.multiProcessorCount;
    for (int w : {8, 16, 32}) {
        run<8, 0>(sms, w, sink);
        run<0, 1>(sms, w, sink);
        run<0, 2>(sms, w, sink);
        run<0, 4>(sms, w, sink);
        run<8, 1>(sms, w, sink);
        run<8, 2>(sms, w, sink);
        run<4, 2>(sms, w, sink);
        run<2, 2>(sms, w, sink);
    }
    return 0;
}
