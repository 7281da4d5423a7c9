// Probe: register-resident symmetric sweep inversion of one M x M matrix (M <= 128) by ONE CTA.
// Finding that motivated it: the round-1 sweep spent ~2400 cycles per pivot whatever M was -- not latency but
// ISSUE slots: ~250 instructions per warp and pivot, of which 16 were the FMAs (per-element selects for the
// pivot row / column, 64-bit register moves between code variants).  This version issues the rank-1 update
// unconditionally and repairs row k / column k in two rare branches; thread-grid shape is a template parameter
// (fewer warps = less per-warp overhead, same FP64 work).  Prints cycles per pivot and the error against a host
// Gauss-Jordan.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sweep_probe sweep_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define GS_THREADS 1024

// A (M x M, row-major, SPD) <- -A^-1 by the first TY warps of the CTA; all GS_THREADS threads must call.
// Thread (ty = warp, tx = lane) owns rows ty + TY a (a < NA) and columns tx + 32 b (b < NB).
// cb: 2 x 128 doubles + 2 (reciprocal pivots), piv: M doubles.  Returns false on a non-positive pivot (uniform).
template <int TY, int NA, int NB>
__device__ __forceinline__ bool sweep_reg(double *A, int M, double *cb, double *piv)
{
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const bool active = ty < TY;
    double *pv = cb + 256;
    double e[NA][NB];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int i = ty + TY * a, j = tx + 32 * b;
            e[a][b] = (active && i < M && j < M) ? A[(size_t)i * M + j] : 0.0;
        }
    bool ok = true;
    for (int k = 0; k < M; ++k) {
        double *buf = cb + (k & 1) * 128;
        const int kx = k & 31, kb = k >> 5;          // column k: lanes tx == kx, slot kb
        const int ry = k % TY, ra = k / TY;          // row k: warp ry, slot ra
        const bool rowk_warp = (ty == ry), colk_lane = (tx == kx);
        if (active && colk_lane) {                   // owners of column k publish it (and the reciprocal pivot)
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < NA; ++a) buf[ty + TY * a] = e[a][b];
                    if (rowk_warp) {
#pragma unroll
                        for (int a = 0; a < NA; ++a)
                            if (a == ra) pv[k & 1] = 1.0 / e[a][b];
                    }
                }
        }
        __syncthreads();
        const double d = buf[k];
        if (!(d > 0.0) || !isfinite(d)) { ok = false; break; }          // uniform: every thread reads the same pivot
        if (!active) continue;
        const double pinv = pv[k & 1];
        if (tid == 0) piv[k] = d;
        double ti[NA], tj[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) tj[b] = buf[tx + 32 * b];
#pragma unroll
        for (int a = 0; a < NA; ++a) ti[a] = buf[ty + TY * a] * pinv;
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) e[a][b] = fma(-ti[a], tj[b], e[a][b]);
        if (rowk_warp) {                             // row k: A[k][j] = c_j / d
#pragma unroll
            for (int a = 0; a < NA; ++a)
                if (a == ra) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) e[a][b] = tj[b] * pinv;
                }
        }
        if (colk_lane) {                             // column k: A[i][k] = c_i / d, A[k][k] = -1 / d
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < NA; ++a) e[a][b] = (rowk_warp && a == ra) ? -pinv : ti[a];
                }
        }
    }
    __syncthreads();
    if (!ok) return false;
    if (active) {
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int i = ty + TY * a, j = tx + 32 * b;
                if (i < M && j < M) A[(size_t)i * M + j] = e[a][b];
            }
    }
    __syncthreads();
    return true;
}


// ---------------------------------------------------------------------------------------------------------
// Look-ahead version.  Pivot ORDER is free for a symmetric positive definite matrix, so the pivots are taken in
// the order k = 32 u + r (r = 0..31 outer, slot u = 0..NB-1 inner): consecutive pivots then sit in statically
// known register slots, and the column of the NEXT pivot can be updated and published first (its owner also
// takes the reciprocal), with the hand-over (mbarrier arrive -> wait) hidden behind the other slots' updates.
// 1024 threads, thread (ty, tx) owns rows ty + 32 a and columns tx + 32 b, a, b < NB.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// one pivot k = 32 U + r whose successor is k' = 32 UN + rn (has_next: there is one); all slots static
template <int NB, int U, int UN>
__device__ __forceinline__ bool la_pivot(double (&e)[NB][NB], int r, int rn, bool has_next, int t, double *cb, double *piv,
                                         uint64_t *bar, uint32_t &phase)
{
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int k = 32 * U + r;
    double *buf = cb + (t & 1) * 128, *nbuf = cb + ((t + 1) & 1) * 128, *pv = cb + 256;
    if (t > 0) {
#ifdef LA_MBAR
        mbar_wait(bar, phase & 1);
        ++phase;
#else
        asm volatile("bar.sync 1, 2048;" ::: "memory");     // split barrier: 32 warps arrive (below) + 32 warps sync
#endif
    }
    const double d = buf[k];
    if (!(d > 0.0) || !isfinite(d)) return false;        // uniform
    const double pinv = pv[t & 1];
    if (tid == 0) piv[t] = d;
    double ti[NB], tj[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) tj[b] = buf[tx + 32 * b];
#pragma unroll
    for (int a = 0; a < NB; ++a) ti[a] = buf[ty + 32 * a] * pinv;
    const bool rowk_warp = (ty == r), colk_lane = (tx == r);
    // ---- the slot of the next pivot's column first
#pragma unroll
    for (int a = 0; a < NB; ++a) e[a][UN] = fma(-ti[a], tj[UN], e[a][UN]);
    if (rowk_warp) e[U][UN] = tj[UN] * pinv;
    if (UN == U && colk_lane) {
#pragma unroll
        for (int a = 0; a < NB; ++a) e[a][U] = (rowk_warp && a == U) ? -pinv : ti[a];
    }
    if (has_next && tx == rn) {
#pragma unroll
        for (int a = 0; a < NB; ++a) nbuf[ty + 32 * a] = e[a][UN];
        if (ty == rn) pv[(t + 1) & 1] = 1.0 / e[UN][UN];
#ifdef LA_MBAR
        mbar_arrive(bar);
#endif
    }
#ifndef LA_MBAR
    if (has_next) asm volatile("bar.arrive 1, 2048;" ::: "memory");
#endif
    // ---- the other slots
#pragma unroll
    for (int b = 0; b < NB; ++b)
        if (b != UN) {
#pragma unroll
            for (int a = 0; a < NB; ++a) e[a][b] = fma(-ti[a], tj[b], e[a][b]);
            if (rowk_warp) e[U][b] = tj[b] * pinv;
        }
    if (UN != U && colk_lane) {
#pragma unroll
        for (int a = 0; a < NB; ++a) e[a][U] = (rowk_warp && a == U) ? -pinv : ti[a];
    }
    return true;
}

// the pivots 32 u + r, u < NV, of group r; `last`: no group follows
template <int NB, int NV>
__device__ __forceinline__ bool la_group(double (&e)[NB][NB], int r, bool last, int &t, double *cb, double *piv, uint64_t *bar,
                                         uint32_t &phase)
{
    bool ok = true;
    if (NV >= 1) { ok = ok && la_pivot<NB, 0, (NV > 1 ? 1 : 0)>(e, r, NV > 1 ? r : r + 1, NV > 1 || !last, t, cb, piv, bar, phase); ++t; }
    if (NV >= 2 && ok) { ok = la_pivot<NB, 1 % NB, (NV > 2 ? 2 % NB : 0)>(e, r, NV > 2 ? r : r + 1, NV > 2 || !last, t, cb, piv, bar, phase); ++t; }
    if (NV >= 3 && ok) { ok = la_pivot<NB, 2 % NB, (NV > 3 ? 3 % NB : 0)>(e, r, NV > 3 ? r : r + 1, NV > 3 || !last, t, cb, piv, bar, phase); ++t; }
    if (NV >= 4 && ok) { ok = la_pivot<NB, 3 % NB, 0>(e, r, r + 1, !last, t, cb, piv, bar, phase); ++t; }
    return ok;
}

// A (M x M, 32 (NB-1) < M <= 32 NB) <- -A^-1.  cb: 258 doubles, piv: M doubles (pivots in elimination order),
// bar: mbarrier initialised with count 32, phase: its running phase counter.  All 1024 threads call.
template <int NB>
__device__ __forceinline__ bool sweep_la(double *A, int M, double *cb, double *piv, uint64_t *bar, uint32_t &phase)
{
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    double e[NB][NB];
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int i = ty + 32 * a, j = tx + 32 * b;
            e[a][b] = (i < M && j < M) ? A[(size_t)i * M + j] : 0.0;
        }
    if (tx == 0) {                                   // pivot 0 = element (0, 0): publish column 0
#pragma unroll
        for (int a = 0; a < NB; ++a) cb[ty + 32 * a] = e[a][0];
        if (ty == 0) cb[256] = 1.0 / e[0][0];
    }
    __syncthreads();
    const int rem = M - 32 * (NB - 1);               // groups r < rem have NB pivots, the others NB - 1
    const int groups = NB > 1 ? 32 : M;
    int t = 0;
    bool ok = true;
    for (int r = 0; r < groups && ok; ++r) {
        const bool last = (r == groups - 1);
        if (r < rem) ok = la_group<NB, NB>(e, r, last, t, cb, piv, bar, phase);
        else ok = la_group<NB, (NB > 1 ? NB - 1 : 1)>(e, r, last, t, cb, piv, bar, phase);
    }
    __syncthreads();
    if (!ok) return false;
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int i = ty + 32 * a, j = tx + 32 * b;
            if (i < M && j < M) A[(size_t)i * M + j] = e[a][b];
        }
    __syncthreads();
    return true;
}

template <int NB>
__global__ void __launch_bounds__(GS_THREADS, 1) probe_la_kernel(const double *src, double *dst, int M, int reps, long long *cycles, int *status)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double cb[258], piv[128];
    __shared__ __align__(8) uint64_t bar;
    double *A = sm;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    long long t = 0;
    for (int r = 0; r < reps; ++r) {
        for (int idx = threadIdx.x; idx < M * M; idx += GS_THREADS) A[idx] = src[idx];
        __syncthreads();
        const long long t0 = clock64();
        const bool ok = sweep_la<NB>(A, M, cb, piv, &bar, phase);
        t += clock64() - t0;
        if (!ok && threadIdx.x == 0) *status = 1;
        if (!ok) return;
    }
    for (int idx = threadIdx.x; idx < M * M; idx += GS_THREADS) dst[idx] = -A[idx];
    if (threadIdx.x == 0) *cycles = t / reps;
}


// ---------------------------------------------------------------------------------------------------------
// Blocked version: pivots in blocks of 8.  Per block: (1) the owners publish the block's 8 columns, (2) ONE warp
// inverts the 8 x 8 pivot block with shuffles (the only sequential part: 8 dependent reciprocals), (3) all threads
// form T = A[:, K] P (one element each), (4) every thread applies the rank-8 update to its NS x NS elements and
// repairs the block's rows / columns.  Three barriers per 8 pivots instead of eight.
// smem: Cs[8][128] columns, Ts[8][128], Ps[8][8], flag.
// ---------------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ bool sweep_blk(double *A, int M, double *Cs, double *Ts, double *Ps, int *flag, double *piv)
{
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    double e[NS][NS];
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) {
            const int i = ty + 32 * a, j = tx + 32 * b;
            e[a][b] = (i < M && j < M) ? A[(size_t)i * M + j] : 0.0;
        }
    if (tid == 0) *flag = 0;
    for (int k0 = 0; k0 < M; k0 += 8) {
        const int nb = (M - k0 < 8) ? (M - k0) : 8;
        const int x0 = k0 & 31, kb = k0 >> 5;            // the block's columns: lanes x0 .. x0 + 7 of slot kb; rows: warps x0 .. x0 + 7, slot kb
        const int cx = tx - x0, ry = ty - x0;            // this lane's column / this warp's row inside the block (valid if 0 <= . < nb)
        const bool colK = cx >= 0 && cx < nb, rowK = ry >= 0 && ry < nb;
        // (1) publish the columns (zero for the columns a partial last block does not have)
        if (cx >= 0 && cx < 8) {
#pragma unroll
            for (int b = 0; b < NS; ++b)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < NS; ++a) Cs[cx * 128 + ty + 32 * a] = colK ? e[a][b] : 0.0;
                }
        }
        __syncthreads();
        // (2) warp 0: P = A_KK^-1 by an 8 x 8 sweep in registers; lane (r = lane / 8, c = lane % 8) holds rows r and r + 4 of column c
        if (ty == 0) {
            const int r = tx >> 3, c = tx & 7;
            double v0 = (r < nb && c < nb) ? Cs[c * 128 + k0 + r] : ((r == c) ? 1.0 : 0.0);
            double v1 = (r + 4 < nb && c < nb) ? Cs[c * 128 + k0 + r + 4] : ((r + 4 == c) ? 1.0 : 0.0);
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double dk = __shfl_sync(0xffffffffu, (k < 4) ? v0 : v1, (k & 3) * 8 + k);
                if (k < nb && (!(dk > 0.0) || !isfinite(dk))) ok = false;
                if (k < nb && tx == 0) piv[k0 + k] = dk;
                const double pinv = 1.0 / dk;
                const double rowk = __shfl_sync(0xffffffffu, (k < 4) ? v0 : v1, (k & 3) * 8 + c);     // A[k][c]
                const double c0 = __shfl_sync(0xffffffffu, v0, r * 8 + k);                           // A[r][k]
                const double c1 = __shfl_sync(0xffffffffu, v1, r * 8 + k);                           // A[r + 4][k]
                const double t0 = c0 * pinv, t1 = c1 * pinv;
                double n0 = fma(-t0, rowk, v0), n1 = fma(-t1, rowk, v1);
                if (c == k) { n0 = t0; n1 = t1; }
                if (r == k) n0 = (c == k) ? -pinv : rowk * pinv;
                if (r + 4 == k) n1 = (c == k) ? -pinv : rowk * pinv;
                v0 = n0; v1 = n1;
            }
            Ps[r * 8 + c] = -v0;
            Ps[(r + 4) * 8 + c] = -v1;
            if (!ok && tx == 0) *flag = 1;
        }
        __syncthreads();
        if (*flag) return false;                         // uniform
        // (3) T[i][c] = sum_c' Cs[c'][i] P[c'][c]: thread t -> (c = t / 128, i = t % 128)
        {
            const int c = tid >> 7, i = tid & 127;
            double s = 0.0;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) s = fma(Cs[cc * 128 + i], Ps[cc * 8 + c], s);
            Ts[c * 128 + i] = s;
        }
        __syncthreads();
        // (4) rank-8 update, then the block's rows and columns
#pragma unroll 2
        for (int c = 0; c < 8; ++c) {
            double ti[NS], tj[NS];
#pragma unroll
            for (int a = 0; a < NS; ++a) ti[a] = Ts[c * 128 + ty + 32 * a];
#pragma unroll
            for (int b = 0; b < NS; ++b) tj[b] = Cs[c * 128 + tx + 32 * b];
#pragma unroll
            for (int a = 0; a < NS; ++a)
#pragma unroll
                for (int b = 0; b < NS; ++b) e[a][b] = fma(-ti[a], tj[b], e[a][b]);
        }
        if (rowK) {                                       // rows of the block: A'[k0 + ry][j] = T[j][ry]
#pragma unroll
            for (int a = 0; a < NS; ++a)
                if (a == kb) {
#pragma unroll
                    for (int b = 0; b < NS; ++b) e[a][b] = Ts[ry * 128 + tx + 32 * b];
                }
        }
        if (colK) {                                       // columns of the block: A'[i][k0 + cx] = T[i][cx]; inside the block -P
#pragma unroll
            for (int b = 0; b < NS; ++b)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < NS; ++a) e[a][b] = (rowK && a == kb) ? -Ps[ry * 8 + cx] : Ts[cx * 128 + ty + 32 * a];
                }
        }
        __syncthreads();                                 // Cs / Ts / Ps are rewritten by the next block
    }
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) {
            const int i = ty + 32 * a, j = tx + 32 * b;
            if (i < M && j < M) A[(size_t)i * M + j] = e[a][b];
        }
    __syncthreads();
    return true;
}

template <int NS>
__global__ void __launch_bounds__(GS_THREADS, 1) probe_blk_kernel(const double *src, double *dst, int M, int reps, long long *cycles, int *status)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double Cs[8 * 128], Ts[8 * 128], Ps[64], piv[128];
    __shared__ int flag;
    double *A = sm;
    long long t = 0;
    for (int r = 0; r < reps; ++r) {
        for (int idx = threadIdx.x; idx < M * M; idx += GS_THREADS) A[idx] = src[idx];
        __syncthreads();
        const long long t0 = clock64();
        const bool ok = sweep_blk<NS>(A, M, Cs, Ts, Ps, &flag, piv);
        t += clock64() - t0;
        if (!ok && threadIdx.x == 0) *status = 1;
        if (!ok) return;
    }
    for (int idx = threadIdx.x; idx < M * M; idx += GS_THREADS) dst[idx] = -A[idx];
    if (threadIdx.x == 0) *cycles = t / reps;
}

template <int TY, int NA, int NB>
__global__ void __launch_bounds__(GS_THREADS, 1) probe_kernel(const double *src, double *dst, int M, int reps, long long *cycles, int *status)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double cb[258], piv[128];
    double *A = sm;
    long long t = 0;
    for (int r = 0; r < reps; ++r) {
        for (int idx = threadIdx.x; idx < M * M; idx += GS_THREADS) A[idx] = src[idx];
        __syncthreads();
        const long long t0 = clock64();
        const bool ok = sweep_reg<TY, NA, NB>(A, M, cb, piv);
        t += clock64() - t0;
        if (!ok && threadIdx.x == 0) *status = 1;
        if (!ok) return;
    }
    for (int idx = threadIdx.x; idx < M * M; idx += GS_THREADS) dst[idx] = -A[idx];
    if (threadIdx.x == 0) *cycles = t / reps;
}

template <int NB>
static void run_la(int M, const std::vector<double> &A, const std::vector<double> &I)
{
    if (M > 32 * NB || M <= 32 * (NB - 1)) return;
    std::vector<double> out((size_t)M * M);
    double *dA, *dO;
    long long *dC;
    int *dS;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dO, A.size() * 8); cudaMalloc(&dC, 8); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4);
    const size_t smem = (size_t)M * M * 8;
    cudaFuncSetAttribute(probe_la_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_la_kernel<NB><<<1, GS_THREADS, smem>>>(dA, dO, M, 20, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; int st = 0;
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dO, A.size() * 8, cudaMemcpyDeviceToHost);
    double err = 0.0, nrm = 0.0;
    for (size_t i = 0; i < out.size(); ++i) { err = fmax(err, fabs(out[i] - I[i])); nrm = fmax(nrm, fabs(I[i])); }
    printf("M=%3d LOOK-AHEAD slots %dx%d: %s status %d  %7lld cycles (%.0f per pivot, %.1f us)  max rel err %.2e\n", M, NB, NB,
           cudaGetErrorString(e), st, cyc, (double)cyc / M, cyc / 1965.0, err / nrm);
    cudaFree(dA); cudaFree(dO); cudaFree(dC); cudaFree(dS);
}

template <int NS>
static void run_blk(int M, const std::vector<double> &A, const std::vector<double> &I)
{
    if (M > 32 * NS || M <= 32 * (NS - 1)) return;
    std::vector<double> out((size_t)M * M);
    double *dA, *dO;
    long long *dC;
    int *dS;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dO, A.size() * 8); cudaMalloc(&dC, 8); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4);
    const size_t smem = (size_t)M * M * 8;
    cudaFuncSetAttribute(probe_blk_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_blk_kernel<NS><<<1, GS_THREADS, smem>>>(dA, dO, M, 20, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; int st = 0;
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dO, A.size() * 8, cudaMemcpyDeviceToHost);
    double err = 0.0, nrm = 0.0;
    for (size_t i = 0; i < out.size(); ++i) { err = fmax(err, fabs(out[i] - I[i])); nrm = fmax(nrm, fabs(I[i])); }
    printf("M=%3d BLOCKED (8 pivots) slots %dx%d: %s status %d  %7lld cycles (%.0f per pivot, %.1f us)  max rel err %.2e\n", M, NS, NS,
           cudaGetErrorString(e), st, cyc, (double)cyc / M, cyc / 1965.0, err / nrm);
    cudaFree(dA); cudaFree(dO); cudaFree(dC); cudaFree(dS);
}

template <int TY, int NA, int NB>
static void run(int M, const std::vector<double> &A, const std::vector<double> &I)
{
    if (TY * NA < M || 32 * NB < M) return;
    std::vector<double> out((size_t)M * M);
    double *dA, *dO;
    long long *dC;
    int *dS;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dO, A.size() * 8); cudaMalloc(&dC, 8); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4);
    const size_t smem = (size_t)M * M * 8;
    cudaFuncSetAttribute(probe_kernel<TY, NA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<TY, NA, NB><<<1, GS_THREADS, smem>>>(dA, dO, M, 20, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0; int st = 0;
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dO, A.size() * 8, cudaMemcpyDeviceToHost);
    double err = 0.0, nrm = 0.0;
    for (size_t i = 0; i < out.size(); ++i) { err = fmax(err, fabs(out[i] - I[i])); nrm = fmax(nrm, fabs(I[i])); }
    printf("M=%3d warps=%2d elems/thread=%2dx%d: %s status %d  %7lld cycles (%.0f per pivot, %.1f us)  max rel err %.2e\n", M, TY, NA, NB,
           cudaGetErrorString(e), st, cyc, (double)cyc / M, cyc / 1965.0, err / nrm);
    cudaFree(dA); cudaFree(dO); cudaFree(dC); cudaFree(dS);
}

int main(int argc, char **argv)
{
    const int Ms[7] = {2, 30, 50, 64, 97, 100, 128};
    for (int mi = 0; mi < 7; ++mi) {
        const int M = Ms[mi];
        std::vector<double> A((size_t)M * M);
        srand(7 + M);
        std::vector<double> B((size_t)M * M);
        for (auto &v : B) v = rand() / (double)RAND_MAX - 0.5;
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) {
                double s = (i == j) ? 0.5 : 0.0;
                for (int k = 0; k < M; ++k) s += B[(size_t)i * M + k] * B[(size_t)j * M + k] / M;
                A[(size_t)i * M + j] = s;
            }
        std::vector<double> W(A), I((size_t)M * M, 0.0);
        for (int i = 0; i < M; ++i) I[(size_t)i * M + i] = 1.0;
        for (int k = 0; k < M; ++k) {
            const double p = 1.0 / W[(size_t)k * M + k];
            for (int j = 0; j < M; ++j) { W[(size_t)k * M + j] *= p; I[(size_t)k * M + j] *= p; }
            for (int i = 0; i < M; ++i)
                if (i != k) {
                    const double f = W[(size_t)i * M + k];
                    for (int j = 0; j < M; ++j) { W[(size_t)i * M + j] -= f * W[(size_t)k * M + j]; I[(size_t)i * M + j] -= f * I[(size_t)k * M + j]; }
                }
        }
        run<32, 4, 4>(M, A, I);
        run<32, 2, 2>(M, A, I);
        run<32, 1, 1>(M, A, I);
        run_blk<1>(M, A, I);
        run_blk<2>(M, A, I);
        run_blk<3>(M, A, I);
        run_blk<4>(M, A, I);
        run_la<1>(M, A, I);
        run_la<2>(M, A, I);
        run_la<3>(M, A, I);
        run_la<4>(M, A, I);
    }
    return 0;
}
