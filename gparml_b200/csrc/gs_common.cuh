// Shared pieces of the master step: parameter block and the O(M^2 Q) tail (bound, hyper-parameter
// gradients, pair table for embed_grads) used by the single-CTA kernel (global_step.cu) and the
// multi-kernel large-M path (global_step_large.cu).
#pragma once
#include <math.h>

#include "common.cuh"

#define GS_THREADS 1024

struct GsParams {
    int M, Q, D;
    int64_t P;
    double n_total;
    int fixed_beta, kmm_only, use_smem;
    int phase;                // 0: whole master step, 1: up to dF/dPsi2 + pair tables (what embed_grads needs), 2: tail only
    const double *stats;
    int64_t off_p1y, off_d1z, off_d1a, off_s0, off_tz, off_ta;
    const double *Z;
    const GlobalsDev *glob;
    const double *pair_lk;
    double *kmm, *kmm_inv, *a_inv, *g_k, *g_1, *g_2, *c_mat, *psi2_full;
    double *X, *W;            // global scratch (used when !use_smem)
    double2 *pair_g, *pair_h;
    double *out;              // [0] F, [1 .. 1+MQ+Q+2) grad (Z, sf2, alpha, beta), then [logdetK, logdetA, trKP, tr1]
    int *status;
};

__device__ __forceinline__ int64_t pidx(int M, int i, int j)
{
    return (i <= j) ? gp_pair_index(M, i, j) : gp_pair_index(M, j, i);
}

// Pair tables for embed_grads from dF/dPsi2 (W, M x M, shared or global memory): (lk, Gs) and
// (lk + log|Gs|, sign Gs).  Callable by any number of threads of one CTA.
__device__ __forceinline__ void gs_pair_tables(const GsParams &p, const double *W)
{
    const int M = p.M;
    const int tid = threadIdx.x;
    const size_t MM = (size_t)M * M;
    for (size_t idx = tid; idx < MM; idx += GS_THREADS) {
        const int i = (int)(idx / M), j = (int)(idx % M);
        if (j < i) continue;
        const int64_t pp = gp_pair_index(M, i, j);
        const double gs = (i == j) ? W[idx] : (W[idx] + W[(size_t)j * M + i]);
        p.pair_g[pp] = make_double2(p.pair_lk[pp], gs);
        // h = Gs Psi2_n = +-exp(lk + log|Gs| + ...): folds the multiplication by Gs into the exponent
        const double lg = log(fabs(gs));
        p.pair_h[pp] = make_double2(p.pair_lk[pp] + (lg > -700.0 ? lg : -700.0), gs < 0.0 ? -1.0 : 1.0);
    }
}

// The bound and the gradients of sf2 / beta (one thread): partial_terms.py:462-472, 322-333, 340-360.
__device__ __forceinline__ void gs_tail_scalars(const GsParams &p, double ldK, double ldA, double tr1, double trKP, double s_ap,
                                                double s_cpc, double s_kk, double s_22)
{
    const int M = p.M, Q = p.Q, D = p.D;
    const double sf2 = p.glob->sf2, beta = p.glob->beta;
    const int nz = M * Q;
    double *grad = p.out + 1;
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

// Bound + gradients from the finished partial derivatives.  GK = dF/dKmm, G2 = dF/dPsi2 (M x M,
// shared or global memory), P2 = full Psi2.  Must be called by all GS_THREADS threads of ONE CTA.
// qred: >= 32 * GP_MAX_Q doubles, ia2: >= GP_MAX_Q doubles of shared memory.
__device__ __forceinline__ void gs_tail(const GsParams &p, const double *X, const double *W, const double *P2, double ldK,
                                        double ldA, double tr1, double trKP, double s_ap, double s_cpc, double s_kk, double s_22,
                                        double *qred, double *ia2)
{
    const int M = p.M, Q = p.Q, D = p.D;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t MM = (size_t)M * M;
    const GlobalsDev g = *p.glob;
    const int nz = M * Q;
    double *grad = p.out + 1;
    if (tid == 0) gs_tail_scalars(p, ldK, ldA, tr1, trKP, s_ap, s_cpc, s_kk, s_22);

    // ---- grad_alpha (partial_terms.py:247-254, 286-299): one pass, Q partial sums per thread ----
    {
        double sq[GP_MAX_Q];
#pragma unroll
        for (int q = 0; q < GP_MAX_Q; ++q) sq[q] = 0.0;
        if (tid < Q) ia2[tid] = 1.0 / (g.alpha[tid] * g.alpha[tid]);
        __syncthreads();
        for (size_t idx = tid; idx < MM; idx += GS_THREADS) {
            const int i = (int)(idx / M), j = (int)(idx % M);
            const int64_t pp = pidx(M, i, j);
            const double gk = X[idx], g2 = W[idx], km = p.kmm[idx], ps = P2[idx];
#pragma unroll
            for (int q = 0; q < GP_MAX_Q; ++q) {
                if (q < Q) {
                    const double dz = p.Z[i * Q + q] - p.Z[j * Q + q];
                    const double ta = p.stats[p.off_ta + (int64_t)q * p.P + pp];
                    sq[q] = fma(gk, -0.5 * km * dz * dz, sq[q]);
                    sq[q] = fma(g2, -0.25 * dz * dz * ps - ta * ia2[q], sq[q]);
                }
            }
        }
        for (int idx = tid; idx < M * D; idx += GS_THREADS) {
            const double g1 = p.g_1[idx];
#pragma unroll
            for (int q = 0; q < GP_MAX_Q; ++q)
                if (q < Q) sq[q] = fma(g1, p.stats[p.off_d1a + (int64_t)q * M * D + idx], sq[q]);
        }
#pragma unroll
        for (int q = 0; q < GP_MAX_Q; ++q) {
            if (q < Q) {
                const double w = gp_warp_sum(sq[q]);
                if (lane == 0) qred[wid * GP_MAX_Q + q] = w;
            }
        }
        __syncthreads();
        if (tid < Q) {
            double s = 0.0;
            for (int w = 0; w < GS_THREADS / 32; ++w) s += qred[w * GP_MAX_Q + tid];
            grad[nz + 1 + tid] = s;
        }
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
            s = fma(X[jm] + X[mj], -al * dz * p.kmm[jm], s);
            s = fma(2.0 * W[jm], -0.5 * al * dz * P2[jm] + TZk[pidx(M, j, m)], s);
        }
        const double *D1Z = p.stats + p.off_d1z + (int64_t)idx * D;
        for (int d = 0; d < D; ++d) s = fma(p.g_1[j * D + d], D1Z[d], s);
        grad[idx] = s;
    }
}
