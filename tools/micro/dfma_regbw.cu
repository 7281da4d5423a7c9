// micro-benchmark: does DFMA throughput on B200 depend on how many of its three 64-bit source
// operands are fresh register reads (register-file bandwidth / operand reuse)?  (development tool)
//   MODE 0: acc[i] = fma(u0,   v0,   acc[i])   one fresh operand  (u0, v0 reusable)
//   MODE 1: acc[i] = fma(u[i], v0,   acc[i])   two fresh operands
//   MODE 2: acc[i] = fma(u[i], v[j], acc[i])   three fresh operands
//   MODE 3: acc[i] = fma(u[i], v[j], acc[i]) in "snake" order: consecutive instructions share u or v
// 16 independent accumulators (ILP 16), 8 warps/SM .. 32 warps/SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, const double *in, double *sink)
{
    __shared__ double2 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_double2(in[threadIdx.x & 31], in[(threadIdx.x + 7) & 31]);
    __syncthreads();
    int kk = threadIdx.x;
    double acc[16], u[8], v[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = in[i] + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { u[i] = in[16 + i] * threadIdx.x; v[i] = in[24 + i] * threadIdx.x; }
    double uv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) uv[i] = in[32 + i];      // warp-uniform values: the compiler keeps them in uniform registers
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(u[0], v[0], acc[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(u[i & 7], v[0], acc[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(u[i & 7], v[(i * 3 + 1) & 7], acc[i]);
        } else if (MODE == 7) {       // two fresh registers + one uniform-register operand
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(u[i & 7], uv[(i * 3 + 1) & 7], acc[i]);
        } else if (MODE == 8) {       // one fresh register (+ reused u0) + uniform operand
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(u[0], uv[(i * 3 + 1) & 7], acc[i]);
        } else if (MODE == 4) {       // same register in two operand slots
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(u[i & 7], u[i & 7], acc[i]);
        } else if (MODE == 5) {       // mode 3 + one broadcast LDS.128 per 4 DFMAs feeding the next group
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double2 w = sm[(it + a) & 63];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int bb = (a & 1) ? 3 - b : b;
                    acc[a * 4 + bb] = fma((b & 2) ? w.x : w.y, v[bb], acc[a * 4 + bb]);
                }
            }
        } else if (MODE == 6) {       // mode 3 + 4 integer instructions per 4 DFMAs
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                kk = (kk * 5 + a) ^ (kk >> 3);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int bb = (a & 1) ? 3 - b : b;
                    acc[a * 4 + bb] = fma(u[a], v[bb], acc[a * 4 + bb]);
                }
            }
        } else {
            // 4 x 4 grid of (u, v) pairs walked boustrophedon: neighbours share u or v
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int bb = (a & 1) ? 3 - b : b;
                    acc[a * 4 + bb] = fma(u[a], v[bb], acc[a * 4 + bb]);
                }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456 || kk == 12345) sink[0] = s;
}

template <int MODE>
void run(int sms, int warps_per_sm, const double *in, double *sink)
{
    const int threads = 256, ctas = sms * (warps_per_sm / 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, threads>>>(100, in, sink);
    cudaEventRecord(e0);
    k<MODE><<<ctas, threads>>>(iters, in, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)ctas * threads * iters * 16.0;
    printf("warps/SM %2d mode %d : %6.2f TFLOP/s (%5.1f%% of 37.2)\n", warps_per_sm, MODE, 2 * ops / ms / 1e9, 2 * ops / ms / 1e9 / 37.22 * 100);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink, *in; cudaMalloc(&sink, 8); cudaMalloc(&in, 64 * 8);
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1e-9 * (i + 1);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {8, 16, 32}) {
        run<0>(p.multiProcessorCount, w, in, sink);
        run<1>(p.multiProcessorCount, w, in, sink);
        run<2>(p.multiProcessorCount, w, in, sink);
        run<3>(p.multiProcessorCount, w, in, sink);
        run<4>(p.multiProcessorCount, w, in, sink);
        run<5>(p.multiProcessorCount, w, in, sink);
        run<6>(p.multiProcessorCount, w, in, sink);
        run<7>(p.multiProcessorCount, w, in, sink);
        run<8>(p.multiProcessorCount, w, in, sink);
    }
    return 0;
}
