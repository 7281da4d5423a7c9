// K1: psi1_stats -- Psi1 and its Y-contractions; plus the on-demand Psi1 matrix.
//
// Replaces (citations relative to /root/reference)
//   kernel_exp.py:51-82, :13-49         Psi1 (n x M) and Psi1^T Y (M x D)
//   partial_terms.py:162-188            sum_n dPsi1Y/dZ      (M, Q, D)
//   partial_terms.py:256-271            sum_n dPsi1Y/dalpha  (Q, M, D)
//
// With a_nq = alpha_q / (alpha_q S_nq + 1), ad_q = a_nq (mu_nq - z_mq):
//   Psi1[n,m]   = exp( lc1_n - 1/2 sum_q ad_q (mu_nq - z_mq) )
//   row (m, 0)      : Psi1                       -> Psi1^T Y
//   row (m, 1+q)    : Psi1 ad_q                  -> dPsi1Y/dZ[m,q,:]
//   row (m, 1+Q+q)  : Psi1 (ad_q^2 + v1_nq)      -> -2 alpha_q^2 dPsi1Y/dalpha[q,m,:]
// and every row is contracted with Y over the points: C[(m,j), d] = sum_n A[n,(m,j)] Y[n,d].
//
// Two stages per point tile inside one CTA: (1) one thread per (point, inducing point)
// evaluates Psi1 and its 1+2Q row entries into shared memory; (2) one thread per row keeps
// DC output columns in registers and runs the small GEMM over the tile.  Q is a run-time
// value here (row entries live in shared memory); DC (columns per thread) is the template.
// <4 % of the evaluation's work; FP64-pipe bound in stage 2.
#include <math.h>

#include "common.cuh"

#define PSI1_THREADS 256

struct Psi1Params {
    const double *rec1, *Y, *Z;
    int64_t n, n_per_split;
    int M, Q, D, R;
    int MB, TN;          // inducing points / points per tile: TN * MB <= 256, MB * (1+2Q) <= 256
    double *partial;     // [splits][M * (1+2Q)][D]
};

template <int DC>
__global__ void __launch_bounds__(PSI1_THREADS)
psi1_stats_kernel(Psi1Params p)
{
    extern __shared__ __align__(16) double sm[];
    const int Q = p.Q, J = 1 + 2 * Q, MB = p.MB, TN = p.TN, rows = MB * J;
    double *A = sm;                          // [TN][rows]
    double *Ys = A + (size_t)TN * rows;      // [TN][DC]
    double *zs = Ys + (size_t)TN * DC;       // [MB][Q]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * MB;
    const int d0 = blockIdx.z * DC;
    for (int idx = tid; idx < MB * Q; idx += PSI1_THREADS) {
        const int m = m0 + idx / Q;
        zs[idx] = (m < p.M) ? p.Z[(size_t)m * Q + idx % Q] : 0.0;
    }
    const int n_l = tid / MB, m_l = tid % MB;
    const bool s1 = tid < TN * MB;
    const bool s2 = tid < rows;
    const int64_t n_lo = (int64_t)blockIdx.y * p.n_per_split;
    const int64_t n_hi = (n_lo + p.n_per_split < p.n) ? (n_lo + p.n_per_split) : p.n;
    double acc[DC];
#pragma unroll
    for (int d = 0; d < DC; ++d) acc[d] = 0.0;
    __syncthreads();

    for (int64_t base = n_lo; base < n_hi; base += TN) {
        for (int idx = tid; idx < TN * DC; idx += PSI1_THREADS) {
            const int nn = idx / DC, dd = idx % DC;
            const int64_t i = base + nn;
            Ys[idx] = (i < n_hi && d0 + dd < p.D) ? p.Y[i * p.D + d0 + dd] : 0.0;
        }
        if (s1) {
            const int64_t i = base + n_l;
            double *ar = A + (size_t)n_l * rows + m_l * J;
            if (i < n_hi && m0 + m_l < p.M) {
                const double *rec = p.rec1 + i * p.R;
                const double *z = zs + m_l * Q;
                double e = rec[3 * Q];
                for (int q = 0; q < Q; ++q) {
                    const double2 ma = *reinterpret_cast<const double2 *>(rec + 2 * q);   // (mu_q, a_q)
                    const double d = ma.x - z[q];
                    const double ad = ma.y * d;
                    e = fma(-0.5 * ad, d, e);
                    ar[1 + q] = ad;
                }
                const double psi = exp(e);
                ar[0] = psi;
                for (int q = 0; q < Q; ++q) {
                    const double ad = ar[1 + q];
                    ar[1 + q] = psi * ad;
                    ar[1 + Q + q] = psi * fma(ad, ad, rec[2 * Q + q]);
                }
            } else {
                for (int j = 0; j < J; ++j) ar[j] = 0.0;
            }
        }
        __syncthreads();
        if (s2) {
            const int cnt = (int)((n_hi - base < TN) ? (n_hi - base) : TN);
            for (int nn = 0; nn < cnt; ++nn) {
                const double a = A[(size_t)nn * rows + tid];
                const double *y = Ys + nn * DC;
#pragma unroll
                for (int d = 0; d < DC; ++d) acc[d] = fma(a, y[d], acc[d]);
            }
        }
        __syncthreads();
    }

    if (s2) {
        const int m = m0 + tid / J, j = tid % J;
        if (m < p.M) {
            double *out = p.partial + ((size_t)blockIdx.y * p.M * J + (size_t)m * J + j) * p.D;
#pragma unroll
            for (int d = 0; d < DC; ++d)
                if (d0 + d < p.D) out[d0 + d] = acc[d];
        }
    }
}

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

template <int DC>
static int launch_dc(gparml_ctx *c, Psi1Params &p, int dchunks)
{
    const int J = 1 + 2 * c->Q;
    const size_t smem = ((size_t)p.TN * p.MB * J + (size_t)p.TN * DC + (size_t)p.MB * c->Q) * sizeof(double);
    GP_CUDA(cudaFuncSetAttribute(psi1_stats_kernel<DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, psi1_stats_kernel<DC>, PSI1_THREADS, smem));
    if (occ < 1) occ = 1;
    const int mblocks = (c->M + p.MB - 1) / p.MB;
    const int64_t per_split_rows = (int64_t)c->M * J * c->D;
    const int64_t slots = (int64_t)c->sm_count * occ;
    int64_t max_splits = (c->n + 8 * p.TN - 1) / (8 * p.TN);
    const int64_t ws_cap = ((int64_t)128 << 20) / (per_split_rows * (int64_t)sizeof(double));
    if (max_splits > ws_cap) max_splits = ws_cap;
    if (max_splits > 65535) max_splits = 65535;
    if (max_splits < 1) max_splits = 1;
    int64_t best = 1;
    double best_eff = -1.0;
    for (int64_t s = 1; s <= max_splits; ++s) {
        const int64_t total = (int64_t)mblocks * dchunks * s;
        const int64_t waves = (total + slots - 1) / slots;
        if (waves > 4) break;
        const double eff = (double)total / (double)(waves * slots);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    const int splits = (int)best;
    p.n_per_split = (c->n + splits - 1) / splits;
    GP_TRY(gp_ensure_ws(c, (size_t)splits * per_split_rows * sizeof(double)));
    p.partial = c->ws;
    dim3 grid(mblocks, splits, dchunks);
    psi1_stats_kernel<DC><<<grid, PSI1_THREADS, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    psi1_reduce_kernel<<<(int)((per_split_rows + 255) / 256), 256, 0, c->stream>>>(
        c->ws, splits, c->M, c->Q, c->D, c->d_glob, c->stats, c->L.off_p1y, c->L.off_d1z, c->L.off_d1a);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_psi1_stats(gparml_ctx *c)
{
    Psi1Params p;
    p.rec1 = c->rec1; p.Y = c->Y; p.Z = c->Z;
    p.n = c->n; p.M = c->M; p.Q = c->Q; p.D = c->D; p.R = gp_rec_len(c->Q);
    const int J = 1 + 2 * c->Q;
    int MB = PSI1_THREADS / J;
    if (MB > c->M) MB = c->M;
    if (MB < 1) MB = 1;
    int TN = PSI1_THREADS / MB;
    if (TN > 64) TN = 64;
    p.MB = MB; p.TN = TN;
    const int dchunks = (c->D + 15) / 16;
    const int per = (c->D + dchunks - 1) / dchunks;     // columns per chunk
    if (per <= 1) return launch_dc<1>(c, p, dchunks);
    if (per <= 2) return launch_dc<2>(c, p, dchunks);
    if (per <= 4) return launch_dc<4>(c, p, dchunks);
    if (per <= 6) return launch_dc<6>(c, p, dchunks);
    if (per <= 8) return launch_dc<8>(c, p, dchunks);
    if (per <= 10) return launch_dc<10>(c, p, dchunks);
    if (per <= 12) return launch_dc<12>(c, p, dchunks);
    if (per <= 14) return launch_dc<14>(c, p, dchunks);
    return launch_dc<16>(c, p, dchunks);
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
