// K4 for large M (two M x M work matrices no longer fit one SM's shared memory, M > 116):
// the same master step as global_step.cu spread over the whole GPU as a short sequence of
// kernels on L2-resident matrices (M = 500: 2 MB each), split like the single-CTA path into
//
//   kmm_only (side stream at set_globals)   Kmm, block sweep -> Kmm^-1, pivots
//   head (context stream; what embed_grads waits for)
//       build / form A   full Psi2, A = Kmm + beta Psi2                  elementwise
//       block sweep      A <- -A^-1 in blocks of NB = 32 pivots: pivot (1 CTA, register-resident sweep),
//                        panel T = A[:, K] Pinv (one warp per row), rank-NB update (64 x 64 tiles)
//       gemm             C = A^-1 Psi1Y
//       head_finish      dF/dPsi1Y, dF/dPsi2 (no M^3 product needed), C C^T kept for the tail,
//                        partial sums of the contractions that need A^-1 only
//       pair tables      (lk, Gs), (lk + log|Gs|, sign) for embed_grads
//   tail (side stream, concurrent with embed_grads)
//       gemm x 2         U = Psi2 Kinv,  T2 = Kinv U
//       assemble_gk      dF/dKmm, <dF/dKmm, Kmm>, tr(Kinv Psi2)
//       grad_alpha / grad_Z   multi-CTA contractions (one warp per element of grad_Z)
//       final            fixed-order sums of the partials, bound, gradients of sf2 / alpha / beta
//
// Round 1 ran everything on the context's stream with a single-CTA O(M^2 Q) tail: 3.5 ms at M = 500, of
// which the tail 2 ms and the four-CTA panel kernel 0.26 ms (B200, c4).
// Same arithmetic as the single-CTA kernel (block sweep == sequential sweep of the same
// pivots), same error behaviour (non-positive pivot -> one retry with 1e-7 on the diagonal like the reference,
// partial_terms.py:453-457, then GPARML_ERR_NOT_PD).
// Reference lines replaced: see global_step.cu.
#include "gs_common.cuh"

#define GSL_NB 32
#define GSL_TILE 64
#define GSL_JITTER 1e-7   // partial_terms.py:454,456

// Every kernel of a block sweep starts here.  status[0]: error bits 1 (Kmm not positive definite), 2 (Kmm + beta Psi2),
// 4 (input check) stop the evaluation -- it raises anyway; bits 8 / 16 only record that a factorisation needed the
// reference's jitter.  status[2] != 0: the retry pass (pass 1) of the current sweep is live; without it the kernels
// of pass 1 return at once (they are launched unconditionally: the host never waits for the first pass).
__device__ __forceinline__ bool gsl_skip(const int *status, int pass)
{
    if (status[0] & 7) return true;
    return pass == 1 && status[2] == 0;
}

// After the first pass of a sweep: a non-positive pivot (fail_bit) becomes one retry with 1e-7 on the diagonal, like
// the reference (partial_terms.py:453-457: slogdet sign < 0 -> + 1e-7 I); event_bit records it.
__global__ void gsl_retry_decide_kernel(int *status, int fail_bit, int event_bit)
{
    const int st = status[0];
    if ((st & fail_bit) && !(st & 4)) {
        status[0] = (st & ~fail_bit) | event_bit;
        status[2] = 1;
    } else {
        status[2] = 0;
    }
}

// X = Kmm (+ beta Psi2) + 1e-7 I for the retry pass
__global__ void __launch_bounds__(256) gsl_retry_build_kernel(GsParams p, double *__restrict__ X, int with_psi2, const int *status)
{
    if (gsl_skip(status, 1)) return;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)p.M * p.M) return;
    const double a = with_psi2 ? fma(p.glob->beta, p.psi2_full[idx], p.kmm[idx]) : p.kmm[idx];
    X[idx] = a + ((idx / p.M == idx % p.M) ? GSL_JITTER : 0.0);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gsl_build_kernel(GsParams p, double *__restrict__ X)
{
    const int M = p.M, Q = p.Q;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0 && !p.kmm_only && p.stats[ST_FLAGS] != 0.0) atomicOr(p.status, 4);      // see common.cuh ST_FLAGS
    if (idx >= (size_t)M * M) return;
    const int i = (int)(idx / M), j = (int)(idx % M);
    double s = 0.0;
    for (int q = 0; q < Q; ++q) {
        const double dz = p.Z[i * Q + q] - p.Z[j * Q + q];
        s = fma(p.glob->alpha[q] * dz, dz, s);
    }
    const double k = p.glob->sf2 * exp(-0.5 * s);
    X[idx] = k;
    p.kmm[idx] = k;
    if (!p.kmm_only) p.psi2_full[idx] = p.stats[p.off_s0 + pidx(M, i, j)];
}

// X = Kmm + beta Psi2
__global__ void __launch_bounds__(256) gsl_form_a_kernel(GsParams p, double *__restrict__ X)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)p.M * p.M) return;
    X[idx] = fma(p.glob->beta, p.psi2_full[idx], p.kmm[idx]);
}

// dst = -src (and optionally a second copy)
__global__ void __launch_bounds__(256) gsl_negate_kernel(const double *__restrict__ src, size_t count, double *__restrict__ dst,
                                                         double *__restrict__ dst2)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const double v = -src[idx];
    dst[idx] = v;
    if (dst2) dst2[idx] = v;
}

// pivot block: Pinv = (A[k0:k0+nb, k0:k0+nb])^-1 by a register-resident sweep; pivots -> piv[]
__global__ void __launch_bounds__(1024) gsl_pivot_kernel(const double *__restrict__ A, int M, int k0, int nb,
                                                         double *__restrict__ pinv_out, double *__restrict__ piv, int *status,
                                                         int fail_bit, int pass)
{
    __shared__ double col[2][GSL_NB];
    if (gsl_skip(status, pass)) return;
    const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
    double e = (r < nb && c < nb) ? A[(size_t)(k0 + r) * M + k0 + c] : 0.0;
    for (int k = 0; k < nb; ++k) {
        double *cb = col[k & 1];
        if (c == k) cb[r] = e;
        __syncthreads();
        const double d = cb[k];
        if (!(d > 0.0) || !isfinite(d)) {
            if (threadIdx.x == 0) atomicOr(status, fail_bit);
            return;                                   // uniform
        }
        const double pi = 1.0 / d;
        if (threadIdx.x == 0) piv[k0 + k] = d;
        const double cr = cb[r], cc = cb[c];
        if (r == k) e = (c == k) ? -pi : cc * pi;
        else e = (c == k) ? cr * pi : fma(-cr * pi, cc, e);
    }
    if (r < nb && c < nb) pinv_out[r * GSL_NB + c] = -e;
}

// panel: Old[i][c] = A[i][k0 + c],  T[i][c] = sum_c' Old[i][c'] Pinv[c'][c].  One warp per row (lane = column c),
// 8 rows per CTA: ceil(M / 8) CTAs instead of the ceil(M / 128) of the thread-per-row version (4 CTAs at M = 500).
__global__ void __launch_bounds__(256) gsl_panel_kernel(const double *__restrict__ A, int M, int k0, int nb,
                                                        const double *__restrict__ pinv, double *__restrict__ T,
                                                        double *__restrict__ Old, const int *status, int pass)
{
    __shared__ double ps[GSL_NB * GSL_NB];
    __shared__ double as[8][GSL_NB];
    if (gsl_skip(status, pass)) return;
    for (int idx = threadIdx.x; idx < GSL_NB * GSL_NB; idx += 256) ps[idx] = pinv[idx];
    const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + r;
    const double a = (i < M && c < nb) ? A[(size_t)i * M + k0 + c] : 0.0;
    as[r][c] = a;
    __syncthreads();
    if (i >= M) return;
    Old[(size_t)i * GSL_NB + c] = a;
    double s = 0.0;
    if (c < nb) {
#pragma unroll
        for (int k = 0; k < GSL_NB; ++k) s = fma(as[r][k], ps[k * GSL_NB + c], s);
    }
    T[(size_t)i * GSL_NB + c] = s;
}

// rank-NB update of the whole matrix, 64 x 64 tile per CTA, 4 x 4 outputs per thread
__global__ void __launch_bounds__(256) gsl_update_kernel(double *__restrict__ A, int M, int k0, int nb,
                                                         const double *__restrict__ pinv, const double *__restrict__ T,
                                                         const double *__restrict__ Old, const int *status, int pass)
{
    __shared__ __align__(16) double Ts[GSL_NB][GSL_TILE];      // [c][i]
    __shared__ __align__(16) double Os[GSL_NB][GSL_TILE];      // [c][j]
    if (gsl_skip(status, pass)) return;
    const int i0 = blockIdx.y * GSL_TILE, j0 = blockIdx.x * GSL_TILE;
    for (int idx = threadIdx.x; idx < GSL_TILE * GSL_NB; idx += 256) {
        const int r = idx / GSL_NB, c = idx % GSL_NB;
        Ts[c][r] = (i0 + r < M) ? T[(size_t)(i0 + r) * GSL_NB + c] : 0.0;
        Os[c][r] = (j0 + r < M) ? Old[(size_t)(j0 + r) * GSL_NB + c] : 0.0;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
    for (int c = 0; c < GSL_NB; ++c) {
        const double2 t01 = *reinterpret_cast<const double2 *>(&Ts[c][ty * 4]), t23 = *reinterpret_cast<const double2 *>(&Ts[c][ty * 4 + 2]);
        const double2 o01 = *reinterpret_cast<const double2 *>(&Os[c][tx * 4]), o23 = *reinterpret_cast<const double2 *>(&Os[c][tx * 4 + 2]);
        const double tv[4] = {t01.x, t01.y, t23.x, t23.y}, ov[4] = {o01.x, o01.y, o23.x, o23.y};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(tv[a], ov[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i >= M) continue;
        const bool ik = i >= k0 && i < k0 + nb;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + tx * 4 + b;
            if (j >= M) continue;
            const bool jk = j >= k0 && j < k0 + nb;
            double v;
            if (ik && jk) v = -pinv[(i - k0) * GSL_NB + (j - k0)];
            else if (jk) v = Ts[j - k0][ty * 4 + a];                      // T[i][j - k0]
            else if (ik) v = T[(size_t)j * GSL_NB + (i - k0)];            // symmetric counterpart
            else v = A[(size_t)i * M + j] - acc[a][b];
            A[(size_t)i * M + j] = v;
        }
    }
}

// C (M x N, ldc) = alpha * A (M x K, lda) * B (K x N, ldb); 64 x 64 x 16 tiles, 4 x 4 per thread
__global__ void __launch_bounds__(256) gsl_gemm_kernel(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb,
                                                       double *__restrict__ C, int ldc, int M, int N, int K, double alpha,
                                                       const int *status)
{
    __shared__ __align__(16) double As[16][GSL_TILE];          // [k][i]
    __shared__ __align__(16) double Bs[16][GSL_TILE];          // [k][j]
    if (*status & 7) return;
    const int i0 = blockIdx.y * GSL_TILE, j0 = blockIdx.x * GSL_TILE;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int kk = 0; kk < K; kk += 16) {
        for (int idx = threadIdx.x; idx < GSL_TILE * 16; idx += 256) {
            const int r = idx / 16, k = idx % 16;                  // A tile: rows i0 + r, cols kk + k
            As[k][r] = (i0 + r < M && kk + k < K) ? A[(size_t)(i0 + r) * lda + kk + k] : 0.0;
            const int k2 = idx / GSL_TILE, c = idx % GSL_TILE;     // B tile: rows kk + k2, cols j0 + c
            Bs[k2][c] = (kk + k2 < K && j0 + c < N) ? B[(size_t)(kk + k2) * ldb + j0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double2 a01 = *reinterpret_cast<const double2 *>(&As[k][ty * 4]), a23 = *reinterpret_cast<const double2 *>(&As[k][ty * 4 + 2]);
            const double2 b01 = *reinterpret_cast<const double2 *>(&Bs[k][tx * 4]), b23 = *reinterpret_cast<const double2 *>(&Bs[k][tx * 4 + 2]);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = i0 + ty * 4 + a, j = j0 + tx * 4 + b;
            if (i < M && j < N) C[(size_t)i * ldc + j] = alpha * acc[a][b];
        }
}

// ---- head --------------------------------------------------------------------------------------
// dF/dPsi1Y = beta^2 C, dF/dPsi2 = 1/2 beta D (Kinv - A^-1) - 1/2 beta^3 C C^T (partial_terms.py:115-131); E = C C^T is
// kept (in the dF/dKmm buffer) for the tail.  part[b] = (tr(A^-1 Psi2), tr(C^T Psi2 C), <G2, Psi2>, <Psi1Y, C>)
__global__ void __launch_bounds__(256) gsl_head_finish_kernel(GsParams p, const double *__restrict__ Kinv, double *__restrict__ part)
{
    __shared__ double sh[33];
    const int M = p.M, D = p.D;
    const size_t MM = (size_t)M * M;
    const double beta = p.glob->beta, hD = 0.5 * (double)D;
    double s[4] = {0, 0, 0, 0};
    if ((*p.status & 7) == 0) {
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < MM; idx += (size_t)gridDim.x * blockDim.x) {
            const int i = (int)(idx / M), j = (int)(idx - (size_t)i * M);
            double e = 0.0;
            for (int d = 0; d < D; ++d) e = fma(p.c_mat[i * D + d], p.c_mat[j * D + d], e);
            const double ai = p.a_inv[idx], wi = Kinv[idx], ps = p.psi2_full[idx];
            const double g2 = hD * beta * (wi - ai) - 0.5 * beta * beta * beta * e;
            p.g_k[idx] = e;                       // C C^T, turned into dF/dKmm by the tail
            p.g_2[idx] = g2;
            s[0] = fma(ai, ps, s[0]);
            s[1] = fma(ps, e, s[1]);
            s[2] = fma(g2, ps, s[2]);
        }
        const double *P1Y = p.stats + p.off_p1y;
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)M * D; idx += (size_t)gridDim.x * blockDim.x) {
            const double c = p.c_mat[idx];
            p.g_1[idx] = beta * beta * c;
            s[3] = fma(P1Y[idx], c, s[3]);
        }
    }
    for (int k = 0; k < 4; ++k) {
        const double v = gp_block_sum(s[k], sh);
        if (threadIdx.x == 0) part[(size_t)blockIdx.x * 4 + k] = v;
    }
}

// pair tables for embed_grads from dF/dPsi2 (same entries as gs_pair_tables, any number of CTAs)
__global__ void __launch_bounds__(256) gsl_pair_tables_kernel(GsParams p)
{
    if (*p.status & 7) return;
    const int M = p.M;
    const size_t MM = (size_t)M * M;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < MM; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / M), j = (int)(idx - (size_t)i * M);
        if (j < i) continue;
        const int64_t pp = gp_pair_index(M, i, j);
        const double gs = (i == j) ? p.g_2[idx] : (p.g_2[idx] + p.g_2[(size_t)j * M + i]);
        p.pair_g[pp] = make_double2(p.pair_lk[pp], gs);
        const double lg = log(fabs(gs));
        p.pair_h[pp] = make_double2(p.pair_lk[pp] + (lg > -700.0 ? lg : -700.0), gs < 0.0 ? -1.0 : 1.0);
    }
}

// ---- tail --------------------------------------------------------------------------------------
// dF/dKmm = 1/2 D Kinv - 1/2 D A^-1 - 1/2 beta D Kinv Psi2 Kinv - 1/2 beta^2 C C^T (partial_terms.py:102-113);
// part[b] = (<dF/dKmm, Kmm>, tr(Kinv Psi2))
__global__ void __launch_bounds__(256) gsl_assemble_gk_kernel(GsParams p, const double *__restrict__ Kinv, const double *__restrict__ U,
                                                              const double *__restrict__ T2, double *__restrict__ part)
{
    __shared__ double sh[33];
    const int M = p.M;
    const size_t MM = (size_t)M * M;
    const double beta = p.glob->beta, hD = 0.5 * (double)p.D;
    double s[2] = {0, 0};
    if ((*p.status & 7) == 0) {
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < MM; idx += (size_t)gridDim.x * blockDim.x) {
            const int i = (int)(idx / M), j = (int)(idx - (size_t)i * M);
            const double e = p.g_k[idx];          // C C^T from the head
            const double gk = hD * Kinv[idx] - hD * p.a_inv[idx] - hD * beta * T2[idx] - 0.5 * beta * beta * e;
            p.g_k[idx] = gk;
            s[0] = fma(gk, p.kmm[idx], s[0]);
            if (i == j) s[1] += U[idx];
        }
    }
    for (int k = 0; k < 2; ++k) {
        const double v = gp_block_sum(s[k], sh);
        if (threadIdx.x == 0) part[(size_t)blockIdx.x * 2 + k] = v;
    }
}

// grad_alpha (partial_terms.py:247-254, 286-299): per-CTA partial sums part[b][q]
__global__ void __launch_bounds__(256) gsl_grad_alpha_kernel(GsParams p, double *__restrict__ part)
{
    __shared__ double sh[33];
    const int M = p.M, Q = p.Q, D = p.D;
    const size_t MM = (size_t)M * M;
    double sq[GP_MAX_Q], ia2[GP_MAX_Q];
#pragma unroll
    for (int q = 0; q < GP_MAX_Q; ++q) {
        sq[q] = 0.0;
        const double al = q < Q ? p.glob->alpha[q] : 1.0;
        ia2[q] = 1.0 / (al * al);
    }
    if ((*p.status & 7) == 0) {
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < MM; idx += (size_t)gridDim.x * blockDim.x) {
            const int i = (int)(idx / M), j = (int)(idx - (size_t)i * M);
            const int64_t pp = pidx(M, i, j);
            const double gk = p.g_k[idx], g2 = p.g_2[idx], km = p.kmm[idx], ps = p.psi2_full[idx];
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
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)M * D; idx += (size_t)gridDim.x * blockDim.x) {
            const double g1 = p.g_1[idx];
#pragma unroll
            for (int q = 0; q < GP_MAX_Q; ++q)
                if (q < Q) sq[q] = fma(g1, p.stats[p.off_d1a + (int64_t)q * M * D + idx], sq[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < GP_MAX_Q; ++q) {
        if (q < Q) {
            const double v = gp_block_sum(sq[q], sh);
            if (threadIdx.x == 0) part[(size_t)blockIdx.x * GP_MAX_Q + q] = v;
        }
    }
}

// grad_Z (partial_terms.py:146-160, 207-240): one warp per element (j, k), lanes over m', fixed-order shuffle tree
__global__ void __launch_bounds__(256) gsl_grad_z_kernel(GsParams p)
{
    if (*p.status & 7) return;
    const int M = p.M, Q = p.Q, D = p.D;
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (idx >= M * Q) return;
    const int j = idx / Q, k = idx - j * Q;
    const double al = p.glob->alpha[k], zjk = p.Z[idx];
    const double *TZk = p.stats + p.off_tz + (int64_t)k * p.P;
    double s = 0.0;
    for (int m = lane; m < M; m += 32) {
        const double dz = zjk - p.Z[m * Q + k];
        const size_t jm = (size_t)j * M + m, mj = (size_t)m * M + j;
        s = fma(p.g_k[jm] + p.g_k[mj], -al * dz * p.kmm[jm], s);
        s = fma(2.0 * p.g_2[jm], -0.5 * al * dz * p.psi2_full[jm] + TZk[pidx(M, j, m)], s);
    }
    const double *D1Z = p.stats + p.off_d1z + (int64_t)idx * D;
    for (int d = lane; d < D; d += 32) s = fma(p.g_1[j * D + d], D1Z[d], s);
    s = gp_warp_sum(s);
    if (lane == 0) p.out[1 + idx] = s;
}

// one CTA: fixed-order sums of all partials, log-determinants, then the bound and the gradients of sf2 / alpha / beta
__global__ void __launch_bounds__(256) gsl_final_kernel(GsParams p, const double *__restrict__ part_h, const double *__restrict__ part_t,
                                                        const double *__restrict__ part_a, int nparts, const double *__restrict__ pivK,
                                                        const double *__restrict__ pivA)
{
    __shared__ double red[33];
    if (*p.status & 3) return;
    const int tid = threadIdx.x, M = p.M, Q = p.Q;
    double h[4], t[2];
    for (int k = 0; k < 4; ++k) {
        double v = 0.0;
        for (int b = tid; b < nparts; b += 256) v += part_h[(size_t)b * 4 + k];
        h[k] = gp_block_sum(v, red);
    }
    for (int k = 0; k < 2; ++k) {
        double v = 0.0;
        for (int b = tid; b < nparts; b += 256) v += part_t[(size_t)b * 2 + k];
        t[k] = gp_block_sum(v, red);
    }
    double v = 0.0, w = 0.0;
    for (int i = tid; i < M; i += 256) { v += log(pivK[i]); w += log(pivA[i]); }
    const double ldK = gp_block_sum(v, red), ldA = gp_block_sum(w, red);
    for (int q = 0; q < Q; ++q) {
        double a = 0.0;
        for (int b = tid; b < nparts; b += 256) a += part_a[(size_t)b * GP_MAX_Q + q];
        a = gp_block_sum(a, red);
        if (tid == 0) p.out[1 + M * Q + 1 + q] = a;
    }
    if (tid == 0) gs_tail_scalars(p, ldK, ldA, h[3], t[1], h[0], h[1], t[0], h[2]);
}

// ---------------------------------------------------------------------------------------------
static int block_sweep_pass(gparml_ctx *c, double *X, double *pinv, double *T, double *Old, double *piv, int fail_bit, int pass)
{
    const int M = c->M;
    const int tiles = (M + GSL_TILE - 1) / GSL_TILE;
    for (int k0 = 0; k0 < M; k0 += GSL_NB) {
        const int nb = (M - k0 < GSL_NB) ? (M - k0) : GSL_NB;
        gsl_pivot_kernel<<<1, 1024, 0, c->stream>>>(X, M, k0, nb, pinv, piv, c->d_status, fail_bit, pass);
        GP_LAUNCH_CHECK(c);
        gsl_panel_kernel<<<(M + 7) / 8, 256, 0, c->stream>>>(X, M, k0, nb, pinv, T, Old, c->d_status, pass);
        GP_LAUNCH_CHECK(c);
        gsl_update_kernel<<<dim3(tiles, tiles), 256, 0, c->stream>>>(X, M, k0, nb, pinv, T, Old, c->d_status, pass);
        GP_LAUNCH_CHECK(c);
    }
    return GPARML_OK;
}

// X <- -X^-1 (pivots -> piv); if a pivot is not positive, once more from X + 1e-7 I (decided on the device: the
// kernels of the second pass are always launched and return at once when the first pass went through)
static int block_sweep(gparml_ctx *c, const GsParams &p, double *X, double *pinv, double *T, double *Old, double *piv, int fail_bit)
{
    GP_TRY(block_sweep_pass(c, X, pinv, T, Old, piv, fail_bit, 0));
    gsl_retry_decide_kernel<<<1, 1, 0, c->stream>>>(c->d_status, fail_bit, fail_bit == 1 ? 8 : 16);
    GP_LAUNCH_CHECK(c);
    gsl_retry_build_kernel<<<(int)(((size_t)c->M * c->M + 255) / 256), 256, 0, c->stream>>>(p, X, fail_bit == 2, c->d_status);
    GP_LAUNCH_CHECK(c);
    return block_sweep_pass(c, X, pinv, T, Old, piv, fail_bit, 1);
}

static int gemm(gparml_ctx *c, const double *A, int lda, const double *B, int ldb, double *C, int ldc, int M, int N, int K, double alpha)
{
    dim3 grid((N + GSL_TILE - 1) / GSL_TILE, (M + GSL_TILE - 1) / GSL_TILE);
    gsl_gemm_kernel<<<grid, 256, 0, c->stream>>>(A, lda, B, ldb, C, ldc, M, N, K, alpha, c->d_status);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// phase 0: whole master step on the context's stream; 1: head only; 2: tail only (p.phase)
int gp_launch_global_step_large(gparml_ctx *c, GsParams &p)
{
    const int M = c->M, D = c->D;
    const size_t MM = (size_t)M * M;
    const int eb = (int)((MM + 255) / 256);
    // scratch layout inside the context's large-M workspace
    double *X = c->scratch_x, *W = c->scratch_w;                  // X: matrix being inverted / U;  W: Kmm^-1
    double *T2 = c->gsl_ws;                                       // (M, M)  Kinv Psi2 Kinv
    double *pinv = T2 + MM;                                       // (NB, NB)
    double *T = pinv + GSL_NB * GSL_NB;                           // (M, NB)
    double *Old = T + (size_t)M * GSL_NB;                         // (M, NB)
    double *pivK = Old + (size_t)M * GSL_NB;                      // (M)
    double *pivA = pivK + M;                                      // (M)
    const int nparts = c->sm_count * 2;
    double *part_h = pivA + M;                                    // (nparts, 4)
    double *part_t = part_h + (size_t)nparts * 4;                 // (nparts, 2)
    double *part_a = part_t + (size_t)nparts * 2;                 // (nparts, GP_MAX_Q)

    if (p.kmm_only) {
        // from set_globals on the side stream: Kmm, Kmm^-1 (W and kmm_inv) and its pivots
        gsl_build_kernel<<<eb, 256, 0, c->stream>>>(p, X);
        GP_LAUNCH_CHECK(c);
        GP_TRY(block_sweep(c, p, X, pinv, T, Old, pivK, 1));
        gsl_negate_kernel<<<eb, 256, 0, c->stream>>>(X, MM, W, c->kmm_inv);
        GP_LAUNCH_CHECK(c);
        return GPARML_OK;
    }
    if (p.phase != 2) {
        // the build kernel is re-run only to expand Psi2 (it rewrites identical Kmm)
        gsl_build_kernel<<<eb, 256, 0, c->stream>>>(p, X);
        GP_LAUNCH_CHECK(c);
        gsl_form_a_kernel<<<eb, 256, 0, c->stream>>>(p, X);
        GP_LAUNCH_CHECK(c);
        GP_TRY(block_sweep(c, p, X, pinv, T, Old, pivA, 2));
        gsl_negate_kernel<<<eb, 256, 0, c->stream>>>(X, MM, c->a_inv, nullptr);
        GP_LAUNCH_CHECK(c);
        GP_TRY(gemm(c, c->a_inv, M, c->stats + c->L.off_p1y, D, c->c_mat, D, M, D, M, 1.0));      // C = A^-1 Psi1Y
        gsl_head_finish_kernel<<<nparts, 256, 0, c->stream>>>(p, W, part_h);
        GP_LAUNCH_CHECK(c);
        gsl_pair_tables_kernel<<<nparts, 256, 0, c->stream>>>(p);
        GP_LAUNCH_CHECK(c);
    }
    if (p.phase != 1) {
        GP_TRY(gemm(c, c->psi2_full, M, W, M, X, M, M, M, M, 1.0));                                 // U = Psi2 Kinv
        GP_TRY(gemm(c, W, M, X, M, T2, M, M, M, M, 1.0));                                           // T2 = Kinv U
        gsl_assemble_gk_kernel<<<nparts, 256, 0, c->stream>>>(p, W, X, T2, part_t);
        GP_LAUNCH_CHECK(c);
        gsl_grad_alpha_kernel<<<nparts, 256, 0, c->stream>>>(p, part_a);
        GP_LAUNCH_CHECK(c);
        gsl_grad_z_kernel<<<(M * c->Q + 7) / 8, 256, 0, c->stream>>>(p);
        GP_LAUNCH_CHECK(c);
        gsl_final_kernel<<<1, 256, 0, c->stream>>>(p, part_h, part_t, part_a, nparts, pivK, pivA);
        GP_LAUNCH_CHECK(c);
    }
    return GPARML_OK;
}

size_t gp_global_step_large_ws_doubles(int M, int sm_count)
{
    return (size_t)M * M + GSL_NB * GSL_NB + 2 * (size_t)M * GSL_NB + 2 * (size_t)M + (size_t)sm_count * 2 * (6 + GP_MAX_Q) + 64;
}
