// micro-benchmark: the embed_grads Psi2 inner loop with the pair table in __constant__ memory --
// z operands arrive as uniform registers (LDCU -> DFMA R, R, UR, R), so every FMA reads two
// 64-bit registers.  Reports the FP64-pipe fraction (4Q + 11 = 51 instructions per point-pair)
// for the compiler-ordered loop (k) and for the pipelined, operand-ordered step of embed_x.cu (k2).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gparml_b200/csrc/gp_exp.cuh"

#define QQ 10
#ifndef NPP
#define NPP 2
#endif
#ifndef MINB
#define MINB 2
#endif
#define NPAIR 340
#define EMBX_PF 3
__constant__ double2 c_zz[NPAIR * QQ];
__constant__ double2 c_h[NPAIR];

// exp(x) split into its dependent steps (gp_exp.cuh: same constants, same result)
struct ExpState {
    double x, t, r, p, tab;
    int k;
};
#define GPX_SHIFT 6755399441055744.0

template <int Q, int NP, bool DO_E, bool DO_A>
__device__ __forceinline__ void embx_step(const double2 *__restrict__ zn, const double2 gn, const double2 *__restrict__ zc,
                                          const double (&hc)[NP], double (&hn)[NP], const double (&kn)[NP],
                                          const double (&A)[NP][Q], const double (&nW)[NP][Q], double (&bz)[NP][Q],
                                          double (&bzz)[NP][Q], double (&ah)[NP], const double *exp_tab)
{
    constexpr int PF = EMBX_PF;
    ExpState es[NP];
    if (DO_E) {
        // ---- block E: exponent of the next pair -------------------------------------------------
        constexpr int NC = (NP == 1) ? 2 : 1;      // sub-chains per sum: always >= 4 independent chains
        double e0[NP][NC], e1[NP][NC];
        double2 z[Q];
#pragma unroll
        for (int q = 0; q < PF && q < Q; ++q) z[q] = zn[q];
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            e0[v][0] = gn.x;
            e1[v][0] = kn[v];
            if (NC == 2) { e0[v][NC - 1] = 0.0; e1[v][NC - 1] = 0.0; }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (q + PF < Q) z[q + PF] = zn[q + PF];
#pragma unroll
            for (int v = 0; v < NP; ++v) e0[v][q % NC] = fma(z[q].x, A[v][q], e0[v][q % NC]);
#pragma unroll
            for (int v = NP - 1; v >= 0; --v) e1[v][q % NC] = fma(z[q].y, nW[v][q], e1[v][q % NC]);
        }
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            if (NC == 2) es[v].x = (e0[v][0] + e0[v][NC - 1]) + (e1[v][0] + e1[v][NC - 1]);
            else es[v].x = e0[v][0] + e1[v][0];
        }
    }
    // ---- block XA: exp steps of the next pair, each followed by one group of accumulations ---------
    double2 zz[Q];
    if (DO_A) {
#pragma unroll
        for (int q = 0; q < PF && q < Q; ++q) zz[q] = zc[q];
#pragma unroll
        for (int v = 0; v < NP; ++v) ah[v] += hc[v];
    }
    const int sg = DO_E ? (__double2hiint(gn.y) & 0x80000000) : 0;
#define EMBX_GROUP(q)                                                                        \
    if (DO_A && (q) < Q) {                                                                   \
        if ((q) + PF < Q) zz[((q) + PF) < Q ? ((q) + PF) : 0] = zc[((q) + PF) < Q ? ((q) + PF) : 0]; \
        _Pragma("unroll") for (int v = 0; v < NP; ++v) bz[v][(q) < Q ? (q) : 0] = fma(hc[v], zz[(q) < Q ? (q) : 0].x, bz[v][(q) < Q ? (q) : 0]); \
        _Pragma("unroll") for (int v = NP - 1; v >= 0; --v) bzz[v][(q) < Q ? (q) : 0] = fma(hc[v], zz[(q) < Q ? (q) : 0].y, bzz[v][(q) < Q ? (q) : 0]); \
    }
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].t = fma(es[v].x, 46.16624130844683, GPX_SHIFT);
    }
    EMBX_GROUP(0)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            es[v].k = __double2loint(es[v].t);
            es[v].t = es[v].t - GPX_SHIFT;
        }
    }
    EMBX_GROUP(1)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            es[v].r = fma(es[v].t, -0.02166084939249829, es[v].x);
            es[v].tab = exp_tab[es[v].k & (GP_EXP_TAB - 1)];
        }
    }
    EMBX_GROUP(2)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].r, 1.0 / 120.0, 1.0 / 24.0);
    }
    EMBX_GROUP(3)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].p, es[v].r, 1.0 / 6.0);
    }
    EMBX_GROUP(4)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].p, es[v].r, 0.5);
    }
    EMBX_GROUP(5)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].p, es[v].r, 1.0);
    }
    EMBX_GROUP(6)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].p, es[v].r, 1.0);
    }
    EMBX_GROUP(7)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = es[v].tab * es[v].p;      // in [1, 2.03)
    }
    EMBX_GROUP(8)
    EMBX_GROUP(9)
    EMBX_GROUP(10)
    EMBX_GROUP(11)
    EMBX_GROUP(12)
    EMBX_GROUP(13)
    EMBX_GROUP(14)
    EMBX_GROUP(15)
#undef EMBX_GROUP
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            int m = es[v].k >> 5;
            m = m < -1021 ? -1021 : m;
            hn[v] = __hiloint2double((__double2hiint(es[v].p) + (m << 20)) ^ sg, __double2loint(es[v].p));
        }
    }
}


#define PROBE_SETUP                                                                                      \
    __shared__ double exp_tab[GP_EXP_TAB];                                                               \
    gp_exp_load_table(exp_tab);                                                                          \
    __syncthreads();                                                                                     \
    double A[NPP][QQ], W[NPP][QQ], bz[NPP][QQ], bzz[NPP][QQ], kn[NPP], ah[NPP];                                      \
    for (int v = 0; v < NPP; ++v) {                                                                        \
        kn[v] = in[(threadIdx.x + v) & 63];                                                              \
        ah[v] = 0;                                                                                       \
        for (int q = 0; q < QQ; ++q) {                                                                   \
            A[v][q] = in[(v * 20 + q + threadIdx.x) & 63];                                               \
            W[v][q] = in[(v * 20 + 10 + q + threadIdx.x) & 63];                                          \
            bz[v][q] = 0;                                                                                \
            bzz[v][q] = 0;                                                                               \
        }                                                                                                \
    }
#define PROBE_FINISH                                                                                     \
    double s = 0;                                                                                        \
    for (int v = 0; v < NPP; ++v) { s += ah[v]; for (int q = 0; q < QQ; ++q) s += bz[v][q] + bzz[v][q]; } \
    out[blockIdx.x * 128 + threadIdx.x] = s;

__global__ void __launch_bounds__(128, MINB) k2(const double *in, double *out, int npairs, int reps)
{
    PROBE_SETUP
    for (int r = 0; r < reps; ++r) {
        double hc[NPP], hn[NPP];
        embx_step<QQ, NPP, true, false>(c_zz, c_h[0], c_zz, hc, hc, kn, A, W, bz, bzz, ah, exp_tab);
#pragma unroll 1
        for (int j = 0; j + 1 < npairs; ++j) {
            embx_step<QQ, NPP, true, true>(c_zz + (j + 1) * QQ, c_h[j + 1], c_zz + j * QQ, hc, hn, kn, A, W, bz, bzz, ah, exp_tab);
            for (int v = 0; v < NPP; ++v) hc[v] = hn[v];
        }
        embx_step<QQ, NPP, false, true>(c_zz, c_h[0], c_zz + (npairs - 1) * QQ, hc, hn, kn, A, W, bz, bzz, ah, exp_tab);
    }
    PROBE_FINISH
}

__global__ void __launch_bounds__(128, MINB) k(const double *in, double *out, int npairs, int reps)
{
    PROBE_SETUP
    for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int j = 0; j < npairs; ++j) {
            const double2 g = c_h[j];
            double e0[NPP], e1[NPP], h[NPP];
#pragma unroll
            for (int v = 0; v < NPP; ++v) { e0[v] = g.x; e1[v] = kn[v]; }
#pragma unroll
            for (int q = 0; q < QQ; ++q) {
                const double2 z = c_zz[j * QQ + q];
#pragma unroll
                for (int v = 0; v < NPP; ++v) { e0[v] = fma(A[v][q], z.x, e0[v]); e1[v] = fma(W[v][q], z.y, e1[v]); }
            }
            const int sg = __double2hiint(g.y) & 0x80000000;
#pragma unroll
            for (int v = 0; v < NPP; ++v) { h[v] = gp_exp_signed(e0[v] + e1[v], exp_tab, sg); ah[v] += h[v]; }
#pragma unroll
            for (int q = 0; q < QQ; ++q) {
                const double2 z = c_zz[j * QQ + q];
#pragma unroll
                for (int v = 0; v < NPP; ++v) { bz[v][q] = fma(h[v], z.x, bz[v][q]); bzz[v][q] = fma(h[v], z.y, bzz[v][q]); }
            }
        }
    }
    PROBE_FINISH
}

template <typename K>
static void run(const char *name, K kern, const double *in, double *out, int ctas)
{
    const int reps = 15;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<ctas, 128>>>(in, out, NPAIR, 1);
    cudaEventRecord(e0);
    kern<<<ctas, 128>>>(in, out, NPAIR, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lanes = (double)ctas * 128 * NPP * NPAIR * reps * (4.0 * QQ + 11);
    printf("%-44s %.3f ms, %.2f TFLOP/s executed = %.1f%% of 37.2 (%s)\n", name, ms, 2 * lanes / ms / 1e9, 2 * lanes / ms / 1e9 / 37.22 * 100,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double2 *hz = new double2[NPAIR * QQ], *hh = new double2[NPAIR];
    for (int i = 0; i < NPAIR * QQ; ++i) hz[i] = make_double2(0.01 * (i % 37) - 0.2, 0.0001 * (i % 31));
    for (int i = 0; i < NPAIR; ++i) hh[i] = make_double2(-0.3 - 0.001 * i, (i & 1) ? -1.0 : 1.0);
    cudaMemcpyToSymbol(c_zz, hz, sizeof(double2) * NPAIR * QQ);
    cudaMemcpyToSymbol(c_h, hh, sizeof(double2) * NPAIR);
    double hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 0.01 * i - 0.3;
    double *in, *out; cudaMalloc(&in, sizeof(hin)); cudaMalloc(&out, 8 * 128 * 4096);
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    const int ctas = p.multiProcessorCount * MINB;
    run("constant table, compiler-ordered loop", k, in, out, ctas);
    run("constant table, pipelined operand-ordered", k2, in, out, ctas);
    return 0;
}
