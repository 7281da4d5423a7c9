// K1 support: the fixed-order reduction of psi1_stats' per-slice partial sums into the packed statistics
// buffer (the kernel itself is in psi1_mma.cu), and the on-demand Psi1 matrix.
//
// Replaces (citations relative to /root/reference)
//   kernel_exp.py:51-82                 Psi1 (n x M)             (psi1_matrix_kernel; partial_terms.exp_K_mi)
//   partial_terms.py:162-188, :256-271  layouts of sum_n dPsi1Y/dZ (M, Q, D) and sum_n dPsi1Y/dalpha (Q, M, D)
#include <math.h>

#include "common.cuh"
#include "gp_exp.cuh"

// sum over splits (fixed order) and scatter into the packed buffer with the reference layouts
__global__ void __launch_bounds__(256) psi1_reduce_kernel(const double *__restrict__ partial, int splits, int M, int Q, int D,
                                                          const GlobalsDev *__restrict__ glob, double *__restrict__ stats,
                                                          int64_t off_p1y, int64_t off_d1z, int64_t off_d1a)
{
    const int J = 1 + 2 * Q;
    const int64_t total = (int64_t)M * J * D;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += partial[(size_t)s * total + i];
    const int d = (int)(i % D);
    const int j = (int)((i / D) % J);
    const int m = (int)(i / ((int64_t)D * J));
    if (j == 0) {
        stats[off_p1y + (int64_t)m * D + d] = a;
    } else if (j <= Q) {
        const int q = j - 1;
        stats[off_d1z + ((int64_t)m * Q + q) * D + d] = a;
    } else {
        const int q = j - 1 - Q;
        const double al = glob->alpha[q];
        stats[off_d1a + ((int64_t)q * M + m) * D + d] = -0.5 * a / (al * al);
    }
}

// used by psi1_mma.cu: partial [splits][M (1+2Q)][D] in c->ws -> packed statistics
void gp_psi1_reduce(gparml_ctx *c, int splits)
{
    const int64_t total = (int64_t)c->M * (1 + 2 * c->Q) * c->D;
    psi1_reduce_kernel<<<(int)((total + 255) / 256), 256, 0, c->stream>>>(
        c->ws, splits, c->M, c->Q, c->D, c->d_glob, c->stats, c->L.off_p1y, c->L.off_d1z, c->L.off_d1a);
}

int gp_launch_psi1_stats_mma(gparml_ctx *c);      // psi1_mma.cu: warp-autonomous tasks, row entries in registers (D <= 16: one column chunk)
int gp_launch_psi1_stats_wide(gparml_ctx *c);     // psi1_wide.cu: stage 1 shared by the CTA through shared memory, all columns at once

#ifndef PSI1_WIDE_MIN_D
#define PSI1_WIDE_MIN_D 17
#endif

int gp_launch_psi1_stats(gparml_ctx *c)
{
    return c->D >= PSI1_WIDE_MIN_D ? gp_launch_psi1_stats_wide(c) : gp_launch_psi1_stats_mma(c);
}

// ---------------------------------------------------------------------------
// Psi1 matrix on demand (partial_terms.exp_K_mi, kernel_exp.py:51-82): (n, M)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) psi1_matrix_kernel(const double *__restrict__ rec1, int R, const double *__restrict__ Z,
                                                          int64_t n, int M, int Q, double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * M) return;
    const int64_t i = idx / M;
    const int m = (int)(idx % M);
    const double *rec = rec1 + i * R;
    double e = rec[3 * Q];
    for (int q = 0; q < Q; ++q) {
        const double d = rec[2 * q] - Z[(size_t)m * Q + q];
        e = fma(-0.5 * rec[2 * q + 1] * d, d, e);
    }
    out[idx] = exp(e);
}

int gp_launch_psi1_matrix(gparml_ctx *c)
{
    const int64_t total = c->n * c->M;
    if (total == 0) return GPARML_OK;
    psi1_matrix_kernel<<<(int)((total + 255) / 256), 256, 0, c->stream>>>(c->rec1, gp_rec_len(c->Q), c->Z, c->n, c->M, c->Q, c->psi1);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}
