// K1: psi1_stats -- Psi1 and its Y-contractions; plus the on-demand Psi1 matrix.
//
// Replaces (citations relative to /root/reference)
//   kernel_exp.py:51-82, :13-49         Psi1 (n x M) and Psi1^T Y (M x D)
//   partial_terms.py:162-188            sum_n dPsi1Y/dZ      (M, Q, D)
//   partial_terms.py:256-271            sum_n dPsi1Y/dalpha  (Q, M, D)
//
// With a_nq = alpha_q / (alpha_q S_nq + 1), ad_q = a_nq (mu_nq - z_mq):
//   Psi1[n,m]   = exp( lc1_n - 1/2 sum_q ad_q (mu_nq - z_mq) )
//   row (0,   m) : Psi1                       -> Psi1^T Y
//   row (1+q, m) : Psi1 ad_q                  -> dPsi1Y/dZ[m,q,:]
//   row (1+Q+q,m): Psi1 (ad_q^2 + v1_nq)      -> -2 alpha_q^2 dPsi1Y/dalpha[q,m,:]
// and every row is contracted with Y over the points: C[row, d] = sum_n A[n, row] Y[n, d].
//
// One CTA (128 threads) owns MB inducing points (MB * (1+2Q) <= 512 rows) and walks its slice of
// the points in tiles of 8, loaded with 16-byte cp.async (LDGSTS) into a double buffer while the
// previous tile is contracted.  Stage 1: one (point, inducing point) item per thread and round,
// record fields read from shared memory, Psi1 and the 1+2Q row entries written to shared memory
// laid out [point][row] with row = j * MB + m (conflict free).  Stage 2: a small register-blocked
// GEMM over the tile, each thread owning RB = 4 rows x DC columns, so that one 8-byte A read and
// one broadcast Y read feed DC resp. RB FMAs (0.45 shared-memory wavefronts per FP64
// instruction; the first version with one row per thread and a run-time Q was LSU- and
// issue-bound).  Templated on Q (stage 1 in registers) and DC (output columns per CTA, <= 10).
// ~4 % of the evaluation's work; FP64 pipe 43 %, LSU 72 % (profiles/).
#include <math.h>

#include "common.cuh"
#include "gp_exp.cuh"

#define PSI1_THREADS 128
#define PSI1_TN 8        // points per tile
#define PSI1_RB 4        // rows per thread in the contraction: MB * (1 + 2Q) <= THREADS * RB

struct Psi1Params {
    const double *rec1, *Y, *Z;
    int64_t n, n_per_split;
    int M, Q, D, R;
    int MB;              // inducing points per CTA
    double *partial;     // [splits][M * (1+2Q)][D]
};

__device__ __forceinline__ void psi1_cp_async8(double *dst_smem, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(gp_smem_u32(dst_smem)), "l"(src) : "memory");
}

__device__ __forceinline__ void psi1_cp_async16(double *dst_smem, const double *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gp_smem_u32(dst_smem)), "l"(src) : "memory");
}

template <int Q, int DC>
__global__ void __launch_bounds__(PSI1_THREADS, 4)
psi1_stats_kernel(Psi1Params p)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double exp_tab[GP_EXP_TAB];
    constexpr int DCP = (DC + 1) & ~1;                    // padded row of the Y tile (16-byte rows)
    constexpr int J = 1 + 2 * Q, R = (3 * Q + 2) & ~1, NV2 = (Q + 1) / 2;
    const int MB = p.MB;
    const int rows = MB * J;
    const int rows_pad = (rows + 1) & ~1;
    double *A = sm;                                      // [TN][rows_pad]
    double *Ys = A + (size_t)PSI1_TN * rows_pad;         // [2][TN][DCP]   (double buffered)
    double *recs = Ys + 2 * PSI1_TN * DCP;               // [2][TN][R]
    double *zs = recs + 2 * PSI1_TN * R;                 // [Q][MB]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * MB;
    const int d0 = blockIdx.z * DC;
    for (int idx = tid; idx < MB * Q; idx += PSI1_THREADS) {
        const int q = idx / MB, ml = idx % MB;
        zs[idx] = (m0 + ml < p.M) ? p.Z[(size_t)(m0 + ml) * Q + q] : 0.0;
    }
    for (int idx = tid; idx < 2 * PSI1_TN * DCP; idx += PSI1_THREADS) Ys[idx] = 0.0;   // columns >= D stay zero
    gp_exp_load_table(exp_tab);
    const int T2 = (rows + PSI1_RB - 1) / PSI1_RB;       // threads active in the contraction (<= THREADS)
    const bool s2 = tid < T2;
    const int64_t n_lo = (int64_t)blockIdx.y * p.n_per_split;
    const int64_t n_hi = (n_lo + p.n_per_split < p.n) ? (n_lo + p.n_per_split) : p.n;
    const int dcols = (p.D - d0 < DC) ? (p.D - d0) : DC; // valid output columns of this chunk
    double acc[PSI1_RB][DC];
#pragma unroll
    for (int k = 0; k < PSI1_RB; ++k)
#pragma unroll
        for (int d = 0; d < DC; ++d) acc[k][d] = 0.0;
    __syncthreads();

    // asynchronous tile loader: records are contiguous, Y rows are strided by D
    auto issue = [&](int64_t base, int buf) {
        const int cnt = (int)((n_hi - base < PSI1_TN) ? (n_hi - base) : PSI1_TN);
        double *rb = recs + buf * PSI1_TN * R, *yb = Ys + buf * PSI1_TN * DCP;
        for (int idx = tid; idx < cnt * (R / 2); idx += PSI1_THREADS) psi1_cp_async16(rb + 2 * idx, p.rec1 + base * R + 2 * idx);
        if (p.D == DCP && ((cnt * DCP) & 1) == 0) {
            // single chunk and unpadded rows: the Y tile is one contiguous, 16-byte aligned block
            for (int idx = tid; idx < cnt * (DCP / 2); idx += PSI1_THREADS) psi1_cp_async16(yb + 2 * idx, p.Y + base * p.D + 2 * idx);
        } else {
            for (int idx = tid; idx < cnt * DCP; idx += PSI1_THREADS) {
                const int nn = idx / DCP, dd = idx % DCP;               // compile-time divisor
                if (dd < dcols) psi1_cp_async8(yb + idx, p.Y + (base + nn) * p.D + d0 + dd);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (n_lo < n_hi) issue(n_lo, 0);

    int buf = 0;
    for (int64_t base = n_lo; base < n_hi; base += PSI1_TN, buf ^= 1) {
        const int cnt = (int)((n_hi - base < PSI1_TN) ? (n_hi - base) : PSI1_TN);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                  // tile data visible; previous contraction finished
        const double *rb = recs + buf * PSI1_TN * R, *yb = Ys + buf * PSI1_TN * DCP;
        // ---- stage 1: one (point, inducing point) item per thread and round ------------------
        for (int item = tid; item < cnt * MB; item += PSI1_THREADS) {
            const int nn = item / MB, ml = item % MB;
            if (m0 + ml < p.M) {
                const double2 *rec = reinterpret_cast<const double2 *>(rb + nn * R);
                double *ar = A + (size_t)nn * rows_pad + ml;                    // row (j, ml) at ar[j * MB]
                double ad[Q];
                double e = rb[nn * R + 3 * Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const double2 ma = rec[q];                                  // (mu_q, a_q)
                    const double d = ma.x - zs[q * MB + ml];
                    ad[q] = ma.y * d;
                    e = fma(-0.5 * ad[q], d, e);
                }
                const double psi = gp_exp(e, exp_tab);
                ar[0] = psi;
#pragma unroll
                for (int k = 0; k < NV2; ++k) {
                    const double2 v2 = rec[Q + k];                              // (v1_2k, v1_2k+1)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = 2 * k + h;
                        if (q < Q) {
                            ar[(1 + q) * MB] = psi * ad[q];
                            ar[(1 + Q + q) * MB] = psi * fma(ad[q], ad[q], h ? v2.y : v2.x);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (base + PSI1_TN < n_hi) issue(base + PSI1_TN, buf ^ 1);   // overlaps the contraction
        // ---- stage 2: RB x DC register tile per thread --------------------------------------
        if (s2) {
#pragma unroll 2
            for (int nn = 0; nn < cnt; ++nn) {
                const double *an = A + (size_t)nn * rows_pad + tid;
                double a[PSI1_RB];
#pragma unroll
                for (int k = 0; k < PSI1_RB; ++k) a[k] = (tid + k * T2 < rows) ? an[k * T2] : 0.0;
                const double2 *y2 = reinterpret_cast<const double2 *>(yb + nn * DCP);
#pragma unroll
                for (int d2 = 0; d2 < DCP / 2; ++d2) {
                    const double2 yy = y2[d2];                                   // broadcast
#pragma unroll
                    for (int k = 0; k < PSI1_RB; ++k) {
                        acc[k][2 * d2] = fma(a[k], yy.x, acc[k][2 * d2]);
                        if (2 * d2 + 1 < DC) acc[k][2 * d2 + 1] = fma(a[k], yy.y, acc[k][2 * d2 + 1]);
                    }
                }
            }
        }
    }

    if (s2) {
#pragma unroll
        for (int k = 0; k < PSI1_RB; ++k) {
            const int r = tid + k * T2;
            if (r < rows) {
                const int j = r / MB, m = m0 + r % MB;
                if (m < p.M) {
                    double *out = p.partial + ((size_t)blockIdx.y * p.M * J + (size_t)m * J + j) * p.D;
#pragma unroll
                    for (int d = 0; d < DC; ++d)
                        if (d0 + d < p.D) out[d0 + d] = acc[k][d];
                }
            }
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

template <int Q, int DC>
static int launch_qdc(gparml_ctx *c, Psi1Params &p, int dchunks)
{
    const int J = 1 + 2 * c->Q;
    const int rows_pad = (p.MB * J + 1) & ~1;
    constexpr int DCP = (DC + 1) & ~1;
    const size_t smem = ((size_t)PSI1_TN * rows_pad + 2 * (size_t)PSI1_TN * DCP + 2 * (size_t)PSI1_TN * p.R + (size_t)p.MB * c->Q) * sizeof(double);
    GP_CUDA(cudaFuncSetAttribute(psi1_stats_kernel<Q, DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, psi1_stats_kernel<Q, DC>, PSI1_THREADS, smem));
    if (occ < 1) occ = 1;
    const int mblocks = (c->M + p.MB - 1) / p.MB;
    const int64_t per_split_rows = (int64_t)c->M * J * c->D;
    const int64_t slots = (int64_t)c->sm_count * occ;
    int64_t max_splits = (c->n + 32 * PSI1_TN - 1) / (32 * PSI1_TN);
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
    psi1_stats_kernel<Q, DC><<<grid, PSI1_THREADS, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    psi1_reduce_kernel<<<(int)((per_split_rows + 255) / 256), 256, 0, c->stream>>>(
        c->ws, splits, c->M, c->Q, c->D, c->d_glob, c->stats, c->L.off_p1y, c->L.off_d1z, c->L.off_d1a);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// used by psi1_mma.cu: partial [splits][M (1+2Q)][D] in c->ws -> packed statistics
void gp_psi1_reduce(gparml_ctx *c, int splits)
{
    const int64_t total = (int64_t)c->M * (1 + 2 * c->Q) * c->D;
    psi1_reduce_kernel<<<(int)((total + 255) / 256), 256, 0, c->stream>>>(
        c->ws, splits, c->M, c->Q, c->D, c->d_glob, c->stats, c->L.off_p1y, c->L.off_d1z, c->L.off_d1a);
}

int gp_launch_psi1_stats_mma(gparml_ctx *c);

int gp_launch_psi1_stats(gparml_ctx *c)
{
#ifndef PSI1_LEGACY
    return gp_launch_psi1_stats_mma(c);
#endif
    Psi1Params p;
    p.rec1 = c->rec1; p.Y = c->Y; p.Z = c->Z;
    p.n = c->n; p.M = c->M; p.Q = c->Q; p.D = c->D; p.R = gp_rec_len(c->Q);
    // balanced blocks of inducing points with MB * (1 + 2Q) rows <= THREADS * RB
    int mb_max = PSI1_THREADS * PSI1_RB / (1 + 2 * c->Q);
    if (mb_max < 1) mb_max = 1;
    const int nblk = (c->M + mb_max - 1) / mb_max;
    p.MB = (c->M + nblk - 1) / nblk;
    // output columns per CTA: chunks of at most 10, rounded up to an instantiated width
    const int dchunks = (c->D + 9) / 10;
    const int per = (c->D + dchunks - 1) / dchunks;
    switch (c->Q) {
#define CASE_Q(q)                                                   \
    case q:                                                         \
        if (per <= 1) return launch_qdc<q, 1>(c, p, dchunks);       \
        if (per <= 2) return launch_qdc<q, 2>(c, p, dchunks);       \
        if (per <= 4) return launch_qdc<q, 4>(c, p, dchunks);       \
        if (per <= 8) return launch_qdc<q, 8>(c, p, dchunks);       \
        return launch_qdc<q, 10>(c, p, dchunks);
        CASE_Q(1) CASE_Q(2) CASE_Q(3) CASE_Q(4) CASE_Q(5) CASE_Q(6) CASE_Q(7) CASE_Q(8)
        CASE_Q(9) CASE_Q(10) CASE_Q(11) CASE_Q(12) CASE_Q(13) CASE_Q(14) CASE_Q(15) CASE_Q(16)
#undef CASE_Q
    }
    gp_set_error("psi1_stats: unsupported Q=%d (1..%d)", c->Q, GP_MAX_Q);
    return GPARML_ERR_ARG;
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
