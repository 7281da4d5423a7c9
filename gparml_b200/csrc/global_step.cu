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
// The reference uses explicit LU inverses and slogdet; here both symmetric positive definite
// matrices are Cholesky-factorised in place, the triangular factor is inverted in place and
// the inverse is formed as L^-T L^-1 (differences are O(cond * eps)).  A failed
// pivot reports GPARML_ERR_NOT_PD (the caller raises LinAlgError, which the reference's
// optimiser wrapper turns into f = inf, scg_adapted.py:55).
//
// Two M x M work matrices (X, W) live in shared memory when 2 M^2 doubles fit (M <= 118);
// larger M runs the same code on L2-resident global scratch.  Latency-bound by design: it is
// replicated on every GPU after the all-reduce, so no second broadcast is needed.
#include <math.h>

#include "common.cuh"

#define GS_THREADS 1024

struct GsParams {
    int M, Q, D;
    int64_t P;
    double n_total;
    int fixed_beta, kmm_only, use_smem;
    const double *stats;
    int64_t off_p1y, off_d1z, off_d1a, off_s0, off_tz, off_ta;
    const double *Z;
    const GlobalsDev *glob;
    const double *pair_lk;
    double *kmm, *kmm_inv, *a_inv, *g_k, *g_1, *g_2, *c_mat;
    double *X, *W;            // global scratch (used when !use_smem)
    double2 *pair_g;
    double *out;              // [0] F, [1 .. 1+MQ+Q+2) grad (Z, sf2, alpha, beta), then [logdetK, logdetA, trKP, tr1]
    int *status;
};

__device__ __forceinline__ int64_t pidx(int M, int i, int j)
{
    return (i <= j) ? gp_pair_index(M, i, j) : gp_pair_index(M, j, i);
}

// In-place lower Cholesky of the lower triangle of A (ld = M).  *fail set on a bad pivot.
__device__ void gs_chol(double *A, int M, int *fail)
{
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int k = 0; k < M; ++k) {
        if (tid == 0) {
            double d = A[(size_t)k * M + k];
            if (!(d > 0.0) || !isfinite(d)) { *fail = 1; d = 1.0; }
            A[(size_t)k * M + k] = sqrt(d);
        }
        __syncthreads();
        const double dk = A[(size_t)k * M + k];
        for (int i = k + 1 + tid; i < M; i += GS_THREADS) A[(size_t)i * M + k] /= dk;
        __syncthreads();
        for (int i = k + 1 + ty; i < M; i += GS_THREADS / 32) {
            const double aik = A[(size_t)i * M + k];
            for (int j = k + 1 + tx; j <= i; j += 32) A[(size_t)i * M + j] -= aik * A[(size_t)j * M + k];
        }
        __syncthreads();
    }
}

// In-place inverse of a lower-triangular matrix (column sweep from the last column; the
// original sub-diagonal column is staged in tmp).  tmp holds >= M doubles.
__device__ void gs_trtri(double *A, int M, double *tmp)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int j = M - 1; j >= 0; --j) {
        for (int k = j + 1 + tid; k < M; k += GS_THREADS) tmp[k] = A[(size_t)k * M + j];
        if (tid == 0) A[(size_t)j * M + j] = 1.0 / A[(size_t)j * M + j];
        __syncthreads();
        const double ajj = A[(size_t)j * M + j];
        for (int i = j + 1 + wid; i < M; i += GS_THREADS / 32) {
            double s = 0.0;
            for (int k = j + 1 + lane; k <= i; k += 32) s = fma(A[(size_t)i * M + k], tmp[k], s);
            s = gp_warp_sum(s);
            if (lane == 0) A[(size_t)i * M + j] = -ajj * s;
        }
        __syncthreads();
    }
}

// out = Linv^T Linv (full symmetric), Linv lower-triangular
__device__ void gs_ltl(const double *Li, int M, double *out)
{
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int i = ty; i < M; i += GS_THREADS / 32) {
        for (int j = tx; j <= i; j += 32) {
            double s = 0.0;
            for (int k = i; k < M; ++k) s = fma(Li[(size_t)k * M + i], Li[(size_t)k * M + j], s);
            out[(size_t)i * M + j] = s;
            out[(size_t)j * M + i] = s;
        }
    }
}

__global__ void __launch_bounds__(GS_THREADS, 1) global_step_kernel(GsParams p)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[33];
    __shared__ int fail;
    const int M = p.M, Q = p.Q, D = p.D;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const size_t MM = (size_t)M * M;
    double *X = p.use_smem ? sm : p.X;
    double *W = p.use_smem ? sm + MM : p.W;
    double *tmp = p.use_smem ? sm + 2 * MM : sm;          // M doubles
    const GlobalsDev g = *p.glob;
    const double sf2 = g.sf2, beta = g.beta;
    const double *S0 = p.stats + p.off_s0;
    const double *P1Y = p.stats + p.off_p1y;
    if (tid == 0) fail = 0;

    // ---- Kmm (kernels.py:108-111 with V = 2 ard^2 = 2 / alpha) -------------------------------
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
    gs_chol(X, M, &fail);
    if (fail) {
        if (tid == 0) atomicOr(p.status, 1);
        return;
    }
    double v = 0.0;
    for (int i = tid; i < M; i += GS_THREADS) v += log(X[(size_t)i * M + i]);
    const double ldK = 2.0 * gp_block_sum(v, red);
    gs_trtri(X, M, tmp);
    gs_ltl(X, M, W);
    __syncthreads();
    for (size_t idx = tid; idx < MM; idx += GS_THREADS) p.kmm_inv[idx] = W[idx];
    if (p.kmm_only) return;

    // ---- A = Kmm + beta Psi2, A^-1 (partial_terms.py:60) ---------------------------------------
    for (size_t idx = tid; idx < MM; idx += GS_THREADS) {
        const int i = (int)(idx / M), j = (int)(idx % M);
        X[idx] = p.kmm[idx] + beta * S0[pidx(M, i, j)];
    }
    __syncthreads();
    gs_chol(X, M, &fail);
    if (fail) {
        if (tid == 0) atomicOr(p.status, 2);
        return;
    }
    v = 0.0;
    for (int i = tid; i < M; i += GS_THREADS) v += log(X[(size_t)i * M + i]);
    const double ldA = 2.0 * gp_block_sum(v, red);
    gs_trtri(X, M, tmp);
    gs_ltl(X, M, p.a_inv);
    __syncthreads();

    // ---- C = A^-1 Psi1Y, G1 = beta^2 C (partial_terms.py:115-121) -----------------------------
    for (int idx = tid; idx < M * D; idx += GS_THREADS) {
        const int i = idx / D, d = idx % D;
        double s = 0.0;
        for (int k = 0; k < M; ++k) s = fma(p.a_inv[(size_t)i * M + k], P1Y[(size_t)k * D + d], s);
        p.c_mat[idx] = s;
        p.g_1[idx] = beta * beta * s;
    }
    __syncthreads();
    v = 0.0;
    for (int idx = tid; idx < M * D; idx += GS_THREADS) v = fma(P1Y[idx], p.c_mat[idx], v);
    const double tr1 = gp_block_sum(v, red);               // tr(Psi1Y^T A^-1 Psi1Y)

    // ---- U = Psi2 Kmm^-1 -> X ---------------------------------------------------------------------
    for (int i = ty; i < M; i += GS_THREADS / 32) {
        for (int j = tx; j < M; j += 32) {
            double s = 0.0;
            for (int k = 0; k < M; ++k) s = fma(S0[pidx(M, i, k)], W[(size_t)k * M + j], s);
            X[(size_t)i * M + j] = s;
        }
    }
    __syncthreads();
    v = 0.0;
    for (int i = tid; i < M; i += GS_THREADS) v += X[(size_t)i * M + i];
    const double trKP = gp_block_sum(v, red);              // tr(Kmm^-1 Psi2)

    // ---- dF/dKmm, dF/dPsi2 (partial_terms.py:102-131) and the scalar contractions ------------
    double s_ap = 0.0, s_cpc = 0.0, s_kk = 0.0, s_22 = 0.0;
    const double hD = 0.5 * (double)D;
    for (int i = ty; i < M; i += GS_THREADS / 32) {
        for (int j = tx; j < M; j += 32) {
            const size_t idx = (size_t)i * M + j;
            double t = 0.0;
            for (int k = 0; k < M; ++k) t = fma(W[(size_t)i * M + k], X[(size_t)k * M + j], t);   // Kinv Psi2 Kinv
            double e = 0.0;
            for (int d = 0; d < D; ++d) e = fma(p.c_mat[i * D + d], p.c_mat[j * D + d], e);     // (C C^T)[i,j]
            const double ai = p.a_inv[idx], wi = W[idx], ps = S0[pidx(M, i, j)];
            const double gk = hD * wi - hD * ai - hD * beta * t - 0.5 * beta * beta * e;
            const double g2 = hD * beta * (wi - ai) - 0.5 * beta * beta * beta * e;
            p.g_k[idx] = gk;
            p.g_2[idx] = g2;
            s_ap = fma(ai, ps, s_ap);
            s_cpc = fma(ps, e, s_cpc);
            s_kk = fma(gk, p.kmm[idx], s_kk);
            s_22 = fma(g2, ps, s_22);
        }
    }
    s_ap = gp_block_sum(s_ap, red);      // tr(A^-1 Psi2)
    s_cpc = gp_block_sum(s_cpc, red);    // tr(C^T Psi2 C)
    s_kk = gp_block_sum(s_kk, red);      // <dF/dKmm, Kmm>
    s_22 = gp_block_sum(s_22, red);      // <dF/dPsi2, Psi2>
    __syncthreads();

    const int nz = M * Q;
    double *grad = p.out + 1;
    if (tid == 0) {
        const double N = p.n_total;
        const double yyt = p.stats[ST_YYT], psi0 = p.stats[ST_PSI0], kl = p.stats[ST_KL], ncount = p.stats[ST_NLOCAL];
        // partial_terms.py:462-472
        const double F = -0.5 * N * D * log(2.0 * 3.14159265358979323846) + 0.5 * D * N * log(beta) + 0.5 * D * ldK
                         - 0.5 * D * ldA - 0.5 * beta * yyt - 0.5 * beta * D * psi0 + 0.5 * beta * D * trKP
                         + 0.5 * beta * beta * tr1 - kl;
        p.out[0] = F;
        // partial_terms.py:322-333 with dF/dPsi0 = -1/2 beta D (:133-138)
        grad[nz] = s_kk / sf2 + (-0.5 * beta * D) * ncount + beta * beta * tr1 / sf2 + 2.0 * s_22 / sf2;
        // partial_terms.py:340-360
        grad[nz + 1 + Q] = p.fixed_beta ? 0.0
                                        : (0.5 * N * D / beta - 0.5 * D * s_ap - 0.5 * yyt - 0.5 * D * psi0 + 0.5 * D * trKP
                                           + beta * tr1 - 0.5 * beta * beta * s_cpc);
        double *extra = p.out + 1 + nz + Q + 2;
        extra[0] = ldK; extra[1] = ldA; extra[2] = trKP; extra[3] = tr1;
    }

    // ---- grad_alpha (partial_terms.py:247-254, 286-299) ----------------------------------------
    for (int q = 0; q < Q; ++q) {
        const double al = g.alpha[q];
        const double *TAq = p.stats + p.off_ta + (int64_t)q * p.P;
        const double *D1Aq = p.stats + p.off_d1a + (int64_t)q * M * D;
        double s = 0.0;
        for (size_t idx = tid; idx < MM; idx += GS_THREADS) {
            const int i = (int)(idx / M), j = (int)(idx % M);
            const double dz = p.Z[i * Q + q] - p.Z[j * Q + q];
            const int64_t pp = pidx(M, i, j);
            s = fma(p.g_k[idx], -0.5 * p.kmm[idx] * dz * dz, s);
            s = fma(p.g_2[idx], -0.25 * dz * dz * S0[pp] - TAq[pp] / (al * al), s);
        }
        for (int idx = tid; idx < M * D; idx += GS_THREADS) s = fma(p.g_1[idx], D1Aq[idx], s);
        s = gp_block_sum(s, red);
        if (tid == 0) grad[nz + 1 + q] = s;
    }

    // ---- grad_Z (partial_terms.py:146-160, 207-240) --------------------------------------------
    for (int idx = tid; idx < nz; idx += GS_THREADS) {
        const int j = idx / Q, k = idx % Q;
        const double al = g.alpha[k], zjk = p.Z[idx];
        const double *TZk = p.stats + p.off_tz + (int64_t)k * p.P;
        double s = 0.0;
        for (int m = 0; m < M; ++m) {
            const double dz = zjk - p.Z[m * Q + k];
            const size_t jm = (size_t)j * M + m, mj = (size_t)m * M + j;
            const int64_t pp = pidx(M, j, m);
            s = fma(p.g_k[jm] + p.g_k[mj], -al * dz * p.kmm[jm], s);
            s = fma(2.0 * p.g_2[jm], -0.5 * al * dz * S0[pp] + TZk[pp], s);
        }
        const double *D1Z = p.stats + p.off_d1z + (int64_t)idx * D;
        for (int d = 0; d < D; ++d) s = fma(p.g_1[j * D + d], D1Z[d], s);
        grad[idx] = s;
    }

    // ---- pair table for embed_grads: (lk, Gs) ----------------------------------------------------
    for (size_t idx = tid; idx < MM; idx += GS_THREADS) {
        const int i = (int)(idx / M), j = (int)(idx % M);
        if (j < i) continue;
        const int64_t pp = gp_pair_index(M, i, j);
        const double gs = (i == j) ? p.g_2[idx] : (p.g_2[idx] + p.g_2[(size_t)j * M + i]);
        p.pair_g[pp] = make_double2(p.pair_lk[pp], gs);
    }
}

int gp_launch_global_step(gparml_ctx *c, bool kmm_only)
{
    GsParams p;
    p.M = c->M; p.Q = c->Q; p.D = c->D; p.P = c->L.P;
    p.n_total = (double)c->n_total;
    p.fixed_beta = (c->flags & GPARML_FLAG_FIXED_BETA) ? 1 : 0;
    p.kmm_only = kmm_only ? 1 : 0;
    const size_t MM = (size_t)c->M * c->M;
    const size_t smem_full = (2 * MM + c->M) * sizeof(double);
    p.use_smem = smem_full <= (size_t)220 * 1024 ? 1 : 0;
    const size_t smem = p.use_smem ? smem_full : (size_t)c->M * sizeof(double);
    p.stats = c->stats;
    p.off_p1y = c->L.off_p1y; p.off_d1z = c->L.off_d1z; p.off_d1a = c->L.off_d1a;
    p.off_s0 = c->L.off_s0; p.off_tz = c->L.off_tz; p.off_ta = c->L.off_ta;
    p.Z = c->Z; p.glob = c->d_glob; p.pair_lk = c->pair_lk;
    p.kmm = c->kmm; p.kmm_inv = c->kmm_inv; p.a_inv = c->a_inv;
    p.g_k = c->g_k; p.g_1 = c->g_1; p.g_2 = c->g_2; p.c_mat = c->c_mat;
    p.X = c->scratch_x; p.W = c->scratch_w;
    p.pair_g = c->pair_g;
    p.out = c->glob_out;
    p.status = c->d_status;
    GP_CUDA(cudaFuncSetAttribute(global_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    global_step_kernel<<<1, GS_THREADS, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}
