// K4: global_step -- the O(M^3) master step as ONE single-CTA fp64 kernel.
//
// Replaces (citations relative to /root/reference)
//   local_MapReduce.py:383-394            cache(): Kmm, Kmm^-1
//   partial_terms.py:54-61                A^-1 = (Kmm + beta Psi2)^-1
//   partial_terms.py:436-473              logmarglik (F)
//   partial_terms.py:102-138              dF/dKmm, dF/dPsi1Y, dF/dPsi2, dF/dPsi0
//   partial_terms.py:146-160,207-240      grad_Z
//   partial_terms.py:247-254,286-299      grad_alpha
//   partial_terms.py:306-333              grad_sf2
//   partial_terms.py:340-360              grad_beta
//   parallel_GPLVM.py:302-369             the glue that sequences them
//
// The reference uses explicit LU inverses and slogdet.  Here both symmetric positive definite
// matrices are inverted in place with the symmetric sweep operator (M rank-1 steps; after
// sweeping every pivot the matrix holds -A^-1 and the pivots are the Schur-complement
// diagonals, so log det = sum log d_k and a non-positive pivot means "not positive definite").
// One barrier pair per pivot and M^2 / 1024 updates per thread: the kernel is latency-bound by
// design and was 10x slower with Cholesky + triangular inverse + L^T L (3-5 barriers per
// column, warp-serial inner products).  Differences to LU are O(cond * eps).  A
// failed pivot reports GPARML_ERR_NOT_PD (the caller raises LinAlgError, which the reference's
// optimiser wrapper turns into f = inf, scg_adapted.py:55).
//
// Two M x M work matrices (X, W) live in shared memory when 2 M^2 doubles fit (M <= 116);
// larger M takes the multi-kernel path of global_step_large.cu.  The kernels are replicated on
// every GPU after the all-reduce, so no second broadcast is needed.
//
// Three launches per evaluation: kmm_only (Kmm, Kmm^-1: side stream at set_globals), the head
// (A^-1, dF/dPsi1Y, dF/dPsi2, pair tables: everything embed_grads waits for -- one inversion, no
// M x M x M product) and the tail (dF/dKmm with its two products, the bound and the hyper-parameter
// gradients: side stream, concurrent with embed_grads).
#include <math.h>

#include "gs_common.cuh"

#define GS_JITTER 1e-7   // partial_terms.py:454,456
#define GS_TI 2          // register tile of the two M x M x M products: 2 x 5 outputs per thread
#define GS_TJ 5

// In-place symmetric sweep of every pivot: A <- -A^-1, returns sum log(pivot) (valid in every
// thread).  Memory-resident version for M > 128: every step streams the whole matrix through
// the CTA (shared-memory / L2 bandwidth bound).  tmp >= M doubles, piv >= M doubles.
__device__ bool gs_sweep_invert_mem(double *A, int M, double *tmp, double *piv, double *red, double *logdet)
{
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int k = 0; k < M; ++k) {
        for (int i = tid; i < M; i += GS_THREADS) tmp[i] = A[(size_t)i * M + k];
        __syncthreads();
        const double d = tmp[k];
        if (!(d > 0.0) || !isfinite(d)) return false;          // uniform: every thread reads the same pivot
        const double pinv = 1.0 / d;
        if (tid == 0) piv[k] = d;
        for (int i = ty; i < M; i += GS_THREADS / 32) {
            const double ti = tmp[i] * pinv;
            double *row = A + (size_t)i * M;
            if (i == k) {
                for (int j = tx; j < M; j += 32) row[j] = (j == k) ? -pinv : tmp[j] * pinv;
            } else {
                for (int j = tx; j < M; j += 32) row[j] = (j == k) ? ti : fma(-ti, tmp[j], row[j]);
            }
        }
        __syncthreads();
    }
    double v = 0.0;
    for (int i = tid; i < M; i += GS_THREADS) v += log(piv[i]);
    *logdet = gp_block_sum(v, red);
    return true;
}

// Register-resident version for M <= 128: thread (ty, tx) owns the elements (ty + 32 a, tx + 32 b), a, b < NS =
// ceil(M / 32), for the whole sweep; only the pivot column travels through shared memory (double buffered, one
// barrier per pivot) together with the reciprocal pivot, which its owner thread alone computes.
// The rank-1 update is issued unconditionally for all NS x NS elements and row k / column k are repaired in two
// rare branches: the first version selected per element ((tx + 32 b == k) ? ... : fma) and spent ~250 issue slots
// per warp and pivot on 16 FMAs -- 2400 cycles per pivot whatever M was (tools/micro/sweep_probe.cu; B200:
// M = 100 121 -> 75 us, M = 50 61 -> 25 us).
template <int NS>
__device__ __forceinline__ bool gs_sweep_invert_reg(double *A, int M, double *tmp2 /* 2 x 128 + 2 */, double *piv, double *red,
                                                    double *logdet)
{
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    double *pv = tmp2 + 256;
    double e[NS][NS];
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) {
            const int i = ty + 32 * a, j = tx + 32 * b;
            e[a][b] = (i < M && j < M) ? A[(size_t)i * M + j] : 0.0;
        }
    for (int k = 0; k < M; ++k) {
        double *buf = tmp2 + (k & 1) * 128;
        const int kx = k & 31, ks = k >> 5;          // column k: lanes tx == kx, slot ks; row k: warp ty == kx, slot ks
        const bool rowk_warp = (ty == kx), colk_lane = (tx == kx);
        if (colk_lane) {                             // owners of column k publish it (and the reciprocal pivot)
#pragma unroll
            for (int b = 0; b < NS; ++b)
                if (b == ks) {
#pragma unroll
                    for (int a = 0; a < NS; ++a) buf[ty + 32 * a] = e[a][b];
                    if (rowk_warp) {
#pragma unroll
                        for (int a = 0; a < NS; ++a)
                            if (a == ks) pv[k & 1] = 1.0 / e[a][b];
                    }
                }
        }
        __syncthreads();
        const double d = buf[k];
        if (!(d > 0.0) || !isfinite(d)) return false;          // uniform: every thread reads the same pivot
        const double pinv = pv[k & 1];
        if (tid == 0) piv[k] = d;
        double ti[NS], tj[NS];
#pragma unroll
        for (int b = 0; b < NS; ++b) tj[b] = buf[tx + 32 * b];
#pragma unroll
        for (int a = 0; a < NS; ++a) ti[a] = buf[ty + 32 * a] * pinv;
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int b = 0; b < NS; ++b) e[a][b] = fma(-ti[a], tj[b], e[a][b]);
        if (rowk_warp) {                             // row k: A[k][j] = c_j / d
#pragma unroll
            for (int a = 0; a < NS; ++a)
                if (a == ks) {
#pragma unroll
                    for (int b = 0; b < NS; ++b) e[a][b] = tj[b] * pinv;
                }
        }
        if (colk_lane) {                             // column k: A[i][k] = c_i / d, A[k][k] = -1 / d
#pragma unroll
            for (int b = 0; b < NS; ++b)
                if (b == ks) {
#pragma unroll
                    for (int a = 0; a < NS; ++a) e[a][b] = (rowk_warp && a == ks) ? -pinv : ti[a];
                }
        }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) {
            const int i = ty + 32 * a, j = tx + 32 * b;
            if (i < M && j < M) A[(size_t)i * M + j] = e[a][b];
        }
    double v = 0.0;
    for (int i = tid; i < M; i += GS_THREADS) v += log(piv[i]);
    *logdet = gp_block_sum(v, red);      // includes the barriers that publish A
    return true;
}

__device__ __forceinline__ bool gs_sweep_invert(double *A, int M, double *tmp, double *piv, double *red, double *logdet)
{
    if (M <= 32) return gs_sweep_invert_reg<1>(A, M, tmp, piv, red, logdet);
    if (M <= 64) return gs_sweep_invert_reg<2>(A, M, tmp, piv, red, logdet);
    if (M <= 96) return gs_sweep_invert_reg<3>(A, M, tmp, piv, red, logdet);
    if (M <= 128) return gs_sweep_invert_reg<4>(A, M, tmp, piv, red, logdet);
    return gs_sweep_invert_mem(A, M, tmp, piv, red, logdet);
}

// C[i][j] = sum_k A[i][k] B[k][j] for a GS_TI x GS_TJ register tile per thread; `emit` consumes
// each finished element.  A and B are M x M, row-major.
template <typename Emit>
__device__ __forceinline__ void gs_matmul(const double *__restrict__ A, const double *__restrict__ B, int M, Emit emit)
{
    const int tiles_j = (M + GS_TJ - 1) / GS_TJ, tiles_i = (M + GS_TI - 1) / GS_TI;
    for (int t = threadIdx.x; t < tiles_i * tiles_j; t += GS_THREADS) {
        const int i0 = (t / tiles_j) * GS_TI, j0 = (t % tiles_j) * GS_TJ;
        double acc[GS_TI][GS_TJ];
#pragma unroll
        for (int a = 0; a < GS_TI; ++a)
#pragma unroll
            for (int b = 0; b < GS_TJ; ++b) acc[a][b] = 0.0;
        for (int k = 0; k < M; ++k) {
            double av[GS_TI], bv[GS_TJ];
#pragma unroll
            for (int a = 0; a < GS_TI; ++a) av[a] = (i0 + a < M) ? A[(size_t)(i0 + a) * M + k] : 0.0;
#pragma unroll
            for (int b = 0; b < GS_TJ; ++b) bv[b] = (j0 + b < M) ? B[(size_t)k * M + j0 + b] : 0.0;
#pragma unroll
            for (int a = 0; a < GS_TI; ++a)
#pragma unroll
                for (int b = 0; b < GS_TJ; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < GS_TI; ++a)
#pragma unroll
            for (int b = 0; b < GS_TJ; ++b)
                if (i0 + a < M && j0 + b < M) emit(i0 + a, j0 + b, acc[a][b]);
    }
}

__global__ void __launch_bounds__(GS_THREADS, 1) global_step_kernel(GsParams p)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[33];
    __shared__ double qred[32 * GP_MAX_Q];
    __shared__ double ia2[GP_MAX_Q];
    const int M = p.M, Q = p.Q, D = p.D;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t MM = (size_t)M * M;
    double *X = p.use_smem ? sm : p.X;
    double *W = p.use_smem ? sm + MM : p.W;
    double *tmp = p.use_smem ? sm + 2 * MM : sm;          // max(M, 258) doubles: pivot column, double buffered, + 2 reciprocal pivots
    double *piv = tmp + (M > 258 ? M : 258);              // M doubles: the pivots
    const GlobalsDev g = *p.glob;
    const double sf2 = g.sf2, beta = g.beta;
    const double *S0 = p.stats + p.off_s0;
    const double *P1Y = p.stats + p.off_p1y;
    double *P2 = p.psi2_full;

    // The kernel runs in two launches.  kmm_only = 1 (launched from set_globals on the side stream,
    // concurrently with the statistics map): Kmm, Kmm^-1 and log det Kmm, which depend on the
    // hyper-parameters only (the reference's cache(), local_MapReduce.py:383-394).  kmm_only = 0:
    // everything that needs the reduced statistics.
    double *ldk_slot = p.out + 1 + M * Q + Q + 2;
    double ldK = 0.0, ldA = 0.0;
    if (p.kmm_only) {
        // ---- Kmm (kernels.py:108-111 with V = 2 ard^2 = 2 / alpha) ---------------------------
        for (size_t idx = tid; idx < MM; idx += GS_THREADS) {
            const int i = (int)(idx / M), j = (int)(idx % M);
            double s = 0.0;
            for (int q = 0; q < Q; ++q) {
                const double dz = p.Z[i * Q + q] - p.Z[j * Q + q];
                s = fma(g.alpha[q] * dz, dz, s);
            }
            const double k = sf2 * exp(-0.5 * s);
            X[idx] = k;
            p.kmm[idx] = k;
        }
        __syncthreads();
        bool ok = gs_sweep_invert(X, M, tmp, piv, red, &ldK);
        if (!ok) {
            // The reference retries once with 1e-7 on the diagonal when the factorisation says "not positive
            // definite" (partial_terms.py:453-454; there only the log-determinant is recomputed, here Kmm^-1 is
            // taken from the same jittered matrix).  Status bit 8 records the event; it is not an error.
            __syncthreads();
            for (size_t idx = tid; idx < MM; idx += GS_THREADS) X[idx] = p.kmm[idx] + ((idx / M == idx % M) ? GS_JITTER : 0.0);
            __syncthreads();
            ok = gs_sweep_invert(X, M, tmp, piv, red, &ldK);
            if (ok && tid == 0) atomicOr(p.status, 8);
        }
        if (!ok) {
            if (tid == 0) atomicOr(p.status, 1);
            return;
        }
        for (size_t idx = tid; idx < MM; idx += GS_THREADS) p.kmm_inv[idx] = -X[idx];
        if (tid == 0) ldk_slot[0] = ldK;
        return;
    }
    if (tid == 0 && p.stats[ST_FLAGS] != 0.0) atomicOr(p.status, 4);      // some shard (any rank) failed its input check
    if (*p.status & 1) return;                 // Kmm was not positive definite
    ldK = ldk_slot[0];

    // ---- head of the master step: what the embeddings map waits for (dF/dPsi1Y, dF/dPsi2 -> pair tables) --------
    // Neither needs an M x M x M product: dF/dPsi2 = 1/2 beta D (Kmm^-1 - A^-1) - 1/2 beta^3 C C^T
    // (partial_terms.py:123-131).  The two products behind dF/dKmm (Kmm^-1 Psi2 Kmm^-1, :102-113) only feed the
    // gradients of Z / alpha / sf2 and run in global_step_tail_kernel on the side stream, next to embed_grads.
    // full Psi2 and A = Kmm + beta Psi2 (partial_terms.py:60); row loops, no integer division
    long long clk[6];
    clk[0] = clock64();
    for (int i = wid; i < M; i += GS_THREADS / 32) {
        const size_t ro = (size_t)i * M;
        for (int j = lane; j < M; j += 32) {
            const double ps = S0[pidx(M, i, j)];
            P2[ro + j] = ps;
            X[ro + j] = fma(beta, ps, p.kmm[ro + j]);
        }
    }
    __syncthreads();
    clk[1] = clock64();
    bool ok = gs_sweep_invert(X, M, tmp, piv, red, &ldA);
    if (!ok) {                                  // partial_terms.py:455-457: one retry with A + 1e-7 I (status bit 16)
        __syncthreads();
        for (int i = wid; i < M; i += GS_THREADS / 32) {
            const size_t ro = (size_t)i * M;
            for (int j = lane; j < M; j += 32) X[ro + j] = fma(beta, P2[ro + j], p.kmm[ro + j]) + (i == j ? GS_JITTER : 0.0);
        }
        __syncthreads();
        ok = gs_sweep_invert(X, M, tmp, piv, red, &ldA);
        if (ok && tid == 0) atomicOr(p.status, 16);
    }
    if (!ok) {
        if (tid == 0) atomicOr(p.status, 2);
        return;
    }
    // X = -A^-1 from here on
    clk[2] = clock64();

    // ---- C = A^-1 Psi1Y, G1 = beta^2 C (partial_terms.py:115-121) -----------------------------
    // M D dot products of length M: four lanes per output (k strided by 4, fixed-order shuffle tree), Psi1Y staged
    // in shared memory behind the shared copy of C when both fit into W
    const int DS = D | 1;                      // odd row stride of the shared copy of C: conflict-free column reads
    const bool cs_smem = (size_t)M * DS <= MM; // W is free in the head: C (M x D) fits there unless D > M
    const bool py_smem = (size_t)M * (DS + D) <= MM;
    double *Ps = W + (size_t)M * DS;
    if (py_smem)
        for (int idx = tid; idx < M * D; idx += GS_THREADS) Ps[idx] = P1Y[idx];
    __syncthreads();
    const double *Pk = py_smem ? Ps : P1Y;
    const int ntask = ((4 * M * D + 31) / 32) * 32;
    for (int t = tid; t < ntask; t += GS_THREADS) {
        const int idx = t >> 2, part = t & 3;
        const bool valid = idx < M * D;
        const int i = valid ? idx / D : 0, d = valid ? idx - i * D : 0;
        double s = 0.0;
        for (int k = part; k < M; k += 4) s = fma(-X[(size_t)i * M + k], Pk[(size_t)k * D + d], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (valid && part == 0) {
            p.c_mat[idx] = s;
            p.g_1[idx] = beta * beta * s;
            if (cs_smem) W[i * DS + d] = s;
        }
    }
    __syncthreads();
    const double *Cs = cs_smem ? W : p.c_mat;
    const int cst = cs_smem ? DS : D;
    double v = 0.0;
    for (int idx = tid; idx < M * D; idx += GS_THREADS) v = fma(P1Y[idx], p.c_mat[idx], v);
    const double tr1 = gp_block_sum(v, red);               // tr(Psi1Y^T A^-1 Psi1Y)

    // ---- dF/dPsi2 (partial_terms.py:123-131) and the scalar contractions that need A^-1 only ----
    clk[3] = clock64();
    double s_ap = 0.0, s_cpc = 0.0, s_22 = 0.0;
    const double hD = 0.5 * (double)D, b3 = 0.5 * beta * beta * beta;
    for (int i = wid; i < M; i += GS_THREADS / 32) {
        const size_t ro = (size_t)i * M;
        for (int j = lane; j < M; j += 32) {
            double e = 0.0;
            for (int d = 0; d < D; ++d) e = fma(Cs[i * cst + d], Cs[j * cst + d], e);     // (C C^T)[i,j]
            const double ai = -X[ro + j], ps = P2[ro + j];
            const double g2 = hD * beta * (p.kmm_inv[ro + j] - ai) - b3 * e;
            p.a_inv[ro + j] = ai;
            p.g_2[ro + j] = g2;
            s_ap = fma(ai, ps, s_ap);
            s_cpc = fma(ps, e, s_cpc);
            s_22 = fma(g2, ps, s_22);
        }
    }
    s_ap = gp_block_sum(s_ap, red);      // tr(A^-1 Psi2)
    s_cpc = gp_block_sum(s_cpc, red);    // tr(C^T Psi2 C)
    s_22 = gp_block_sum(s_22, red);      // <dF/dPsi2, Psi2>
    if (tid == 0) {
        double *extra = p.out + 1 + M * Q + Q + 2;
        extra[0] = ldK; extra[1] = ldA; extra[3] = tr1;
        extra[4] = s_ap; extra[5] = s_cpc; extra[7] = s_22;
    }
    __syncthreads();                     // publishes g_2 to the whole CTA
    clk[4] = clock64();
    gs_pair_tables(p, p.g_2);
    __syncthreads();
    clk[5] = clock64();
    if (tid == 0) {                      // probe: SM cycles per section (GPARML_A_GS_EXTRA[8..12])
        double *extra = p.out + 1 + M * Q + Q + 2;
        for (int k = 0; k < 5; ++k) extra[8 + k] = (double)(clk[k + 1] - clk[k]);
    }
}

// Tail of the master step: dF/dKmm (partial_terms.py:102-113) with its two M x M x M products, then the bound
// and the hyper-parameter gradients (gs_tail).  One CTA; independent of embed_grads, so it runs concurrently with
// it on the side stream (capi.cu gparml_global_step_begin).
__global__ void __launch_bounds__(GS_THREADS, 1) global_step_tail_kernel(GsParams p)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[33];
    __shared__ double qred[32 * GP_MAX_Q];
    __shared__ double ia2[GP_MAX_Q];
    if (*p.status & 3) return;
    const int M = p.M, D = p.D;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t MM = (size_t)M * M;
    double *X = sm, *W = sm + MM;
    const double beta = p.glob->beta;
    const double *P2 = p.psi2_full;
    for (size_t idx = tid; idx < MM; idx += GS_THREADS) W[idx] = p.kmm_inv[idx];
    __syncthreads();
    // U = Psi2 Kmm^-1 -> X
    gs_matmul(P2, W, M, [&](int i, int j, double val) { X[(size_t)i * M + j] = val; });
    __syncthreads();
    double v = 0.0;
    for (int i = tid; i < M; i += GS_THREADS) v += X[(size_t)i * M + i];
    const double trKP = gp_block_sum(v, red);              // tr(Kmm^-1 Psi2)
    double s_kk = 0.0;
    const double hD = 0.5 * (double)D;
    gs_matmul(W, X, M, [&](int i, int j, double t) {       // t = (Kinv Psi2 Kinv)[i,j]
        const size_t idx = (size_t)i * M + j;
        double e = 0.0;
        for (int d = 0; d < D; ++d) e = fma(p.c_mat[i * D + d], p.c_mat[j * D + d], e);     // (C C^T)[i,j]
        const double gk = hD * W[idx] - hD * p.a_inv[idx] - hD * beta * t - 0.5 * beta * beta * e;
        p.g_k[idx] = gk;
        s_kk = fma(gk, p.kmm[idx], s_kk);
    });
    s_kk = gp_block_sum(s_kk, red);      // <dF/dKmm, Kmm>
    __syncthreads();
    (void)lane; (void)wid;
    const double *extra = p.out + 1 + p.M * p.Q + p.Q + 2;
    gs_tail(p, p.g_k, p.g_2, P2, extra[0], extra[1], extra[3], trKP, extra[4], extra[5], s_kk, extra[7], qred, ia2);
}

int gp_launch_global_step_large(gparml_ctx *c, GsParams &p);
size_t gp_global_step_large_ws_doubles(int M, int sm_count);

static int launch_gs(gparml_ctx *c, bool kmm_only, int phase);

// kmm_only launches go to the side stream `s` (concurrent with the statistics map), the rest to
// the context's main stream; all helpers launch on c->stream, which is swapped for the call.
int gp_launch_global_step(gparml_ctx *c, bool kmm_only, cudaStream_t s)
{
    cudaStream_t saved = c->stream;
    c->stream = s;
    const int r = launch_gs(c, kmm_only, 0);
    c->stream = saved;
    return r;
}

// phase 1 (head) on `s`; *split = true if the tail still has to be launched (gp_launch_global_step_tail)
int gp_launch_global_step_head(gparml_ctx *c, cudaStream_t s, bool *split)
{
    *split = true;                            // both paths (single-CTA and multi-kernel) have a separate tail
    cudaStream_t saved = c->stream;
    c->stream = s;
    const int r = launch_gs(c, false, 1);
    c->stream = saved;
    return r;
}

int gp_launch_global_step_tail(gparml_ctx *c, cudaStream_t s)
{
    cudaStream_t saved = c->stream;
    c->stream = s;
    const int r = launch_gs(c, false, 2);
    c->stream = saved;
    return r;
}

static int launch_gs(gparml_ctx *c, bool kmm_only, int phase)
{
    GsParams p;
    p.phase = phase;
    if (!kmm_only && phase != 2) c->pair_ra_stale = true;      // the head rewrites pair_h: embed_psi2m's scaled table follows
    p.M = c->M; p.Q = c->Q; p.D = c->D; p.P = c->L.P;
    p.n_total = (double)c->n_total;
    p.fixed_beta = (c->flags & GPARML_FLAG_FIXED_BETA) ? 1 : 0;
    p.kmm_only = kmm_only ? 1 : 0;
    const size_t MM = (size_t)c->M * c->M;
    const size_t vec = (size_t)(c->M > 258 ? c->M : 258) + c->M;      // pivot column buffers + pivots
    const size_t smem_full = (2 * MM + vec) * sizeof(double);
    p.use_smem = smem_full <= (size_t)216 * 1024 ? 1 : 0;
    const size_t smem = p.use_smem ? smem_full : vec * sizeof(double);
    p.stats = c->stats;
    p.off_p1y = c->L.off_p1y; p.off_d1z = c->L.off_d1z; p.off_d1a = c->L.off_d1a;
    p.off_s0 = c->L.off_s0; p.off_tz = c->L.off_tz; p.off_ta = c->L.off_ta;
    p.Z = c->Z; p.glob = c->d_glob; p.pair_lk = c->pair_lk;
    p.kmm = c->kmm; p.kmm_inv = c->kmm_inv; p.a_inv = c->a_inv;
    p.g_k = c->g_k; p.g_1 = c->g_1; p.g_2 = c->g_2; p.c_mat = c->c_mat;
    p.psi2_full = c->psi2_full;
    p.X = c->scratch_x; p.W = c->scratch_w;
    p.pair_g = c->pair_g; p.pair_h = c->pair_h;
    p.out = c->glob_out;
    p.status = c->d_status;
    if (!p.use_smem) {
        // M too large for one SM's shared memory: multi-kernel path on L2-resident matrices
        if (!c->gsl_ws) GP_CUDA(cudaMalloc((void **)&c->gsl_ws, gp_global_step_large_ws_doubles(c->M, c->sm_count) * sizeof(double)));
        return gp_launch_global_step_large(c, p);
    }
    if (phase == 2) {
        const size_t smem_tail = 2 * MM * sizeof(double);
        GP_CUDA(cudaFuncSetAttribute(global_step_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tail));
        global_step_tail_kernel<<<1, GS_THREADS, smem_tail, c->stream>>>(p);
        GP_LAUNCH_CHECK(c);
        return GPARML_OK;
    }
    GP_CUDA(cudaFuncSetAttribute(global_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    global_step_kernel<<<1, GS_THREADS, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    if (!kmm_only && phase == 0) {          // whole master step on one stream: head, then tail
        const size_t smem_tail = 2 * MM * sizeof(double);
        GP_CUDA(cudaFuncSetAttribute(global_step_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tail));
        global_step_tail_kernel<<<1, GS_THREADS, smem_tail, c->stream>>>(p);
        GP_LAUNCH_CHECK(c);
    }
    return GPARML_OK;
}
