// micro-benchmark: how fast can FP64 tensor-core MMAs (mma.sync m8n8k4 f64, SASS DMMA) be FED on B200?
// Phase B of psi1_wide_kernel in isolation: per group of 4 points a warp loads NA A fragments and NB B fragments from shared
// memory and issues NA x NB DMMAs on NA x NB accumulator pairs.  Modes: 0 operands fixed in registers, 1 operands loaded
// from shared memory (conflict-free layout), 2 loaded with the Y-tile layout of the kernel (row stride 56 doubles).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_feed dmma_feed.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int NA, int NB, int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, double *sink)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gi = lane >> 2, kk = lane & 3;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double C[NA][NB][2];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) C[a][b][0] = C[a][b][1] = 0.0;
    double ar[NA], br[NB];
#pragma unroll
    for (int a = 0; a < NA; ++a) ar[a] = 1.0 + a;
#pragma unroll
    for (int b = 0; b < NB; ++b) br[b] = 2.0 + b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int ks = 0; ks < 16; ++ks) {
            const int pt = 4 * ks + kk;
            if (MODE >= 1) {
#pragma unroll
                for (int b = 0; b < NB; ++b) br[b] = (MODE == 2) ? sm[8192 + pt * 56 + 8 * (2 * b + (warp & 1)) + gi] : sm[8192 + (b * 64 + pt) * 8 + gi];
            }
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                if (MODE >= 1) ar[a] = sm[((warp % 7) + 7 * a) * 512 % 8192 + pt * 8 + gi];
#pragma unroll
                for (int b = 0; b < NB; ++b) dmma(C[a][b], ar[a], br[b]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) s += C[a][b][0] + C[a][b][1];
    if (s == 123.456) sink[0] = s;
}

template <int NA, int NB, int MODE>
static void run(int warps)
{
    double *sink;
    cudaMalloc(&sink, 8);
    const int iters = 2000;
    cudaFuncSetAttribute(k<NA, NB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NA, NB, MODE><<<148, warps * 32, 16384 * 8>>>(10, sink);
    cudaEventRecord(e0);
    k<NA, NB, MODE><<<148, warps * 32, 16384 * 8>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 256.0 * NA * NB * 16.0 * iters * warps * 148.0;
    printf("warps/SM %2d  %d x %d DMMA per 4 points  mode %d : %8.3f ms  %6.2f TFLOP/s  %s\n", warps, NA, NB, MODE, ms, flop / ms / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main()
{
    run<3, 4, 0>(16); run<3, 4, 1>(16); run<3, 4, 2>(16);
    run<3, 4, 0>(8);  run<3, 4, 1>(8);  run<3, 4, 2>(8);
    run<3, 4, 1>(14); run<3, 4, 2>(14);
    run<1, 1, 0>(16); run<1, 1, 1>(16);
    run<2, 2, 1>(16); run<4, 4, 1>(16); run<3, 8, 1>(8);
    return 0;
}
