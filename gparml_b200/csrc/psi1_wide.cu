// K1 for wide outputs (D > 16): psi1_stats with the stage-1 values shared by the whole CTA.
//
// Replaces the same reference lines as psi1_mma.cu (kernel_exp.py:13-82, partial_terms.py:162-188, :256-271) and
// writes the same per-slice partial layout, so psi1_reduce_kernel finishes both.
//
// Why a second kernel.  In psi1_mma.cu a warp is an autonomous task that keeps its 1 + 2Q row entries in
// registers and contracts them with at most 16 + 2 output columns; for wider D the columns are chunked and every
// chunk re-evaluates Psi1 and the row entries (D = 50: five times, 69 ms at c4 on B200, of which the repeated stage 1
// is a quarter and the 255-register serialisation of stage 1 and MMAs most of the rest).  Here a CTA of 8 warps owns
// 8 inducing points and a slice of the points and works in steps of 32 points:
//
//   warp 7 (producer)    evaluates Psi1 and the J = 1 + 2Q row entries of the step's 32 x 8 (point, inducing point)
//                        items ONCE (lane = (inducing point, point of a group of 4), 8 passes) and stores them in
//                        shared memory in A-fragment order; it also brings the step's Y rows in with cp.async;
//   warps 0..6 (consumers) own the rows j = w, w + 7, w + 14, ... and ALL ceil(D / 8) column tiles: per group of 4
//                        points they load the column tiles' B fragments once, each row's A fragment once, and issue
//                        rows x tiles FP64 tensor-core instructions (mma.sync m8n8k4, SASS DMMA) on accumulators that
//                        stay in registers for the whole slice.
//
// The two roles work on alternating halves of a double-buffered tile, one block barrier per step.  Per 32 items the
// FP64 pipe sees stage 1 once (6Q + 11 instructions) plus J ceil(D / 8) MMAs instead of chunks x (stage 1 + J (NT MMAs
// + DR FMAs)): D = 50, Q = 10: 2494 vs 2810 pipe cycles, and the MMA stream no longer waits for stage 1.
//
// Bound: FP64 pipe (DMMA shares it with DFMA, tools/micro/dmma_probe.cu).
#include <math.h>

#include "common.cuh"
#include "gp_exp.cuh"

#define P1W_WARPS 8
#define P1W_CONS 7           // consumer warps
#define P1W_TP 32            // points per step (8 MMA k-steps)
#define P1W_MAXNT 8          // column tiles per pass (D <= 64 per pass, more: blockIdx.y chunks)

struct Psi1WParams {
    const double *rec1, *Y, *Z;
    int64_t n, n_per_split;
    int M, D, G, S;
    double *partial;     // [S][M * (1+2Q)][D]
};

__device__ __forceinline__ void p1w_cp_async8(double *dst_smem, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(gp_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void p1w_cp_async16(double *dst_smem, const double *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gp_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void p1w_dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int Q>
__global__ void __launch_bounds__(P1W_WARPS * 32, 1)
psi1_wide_kernel(Psi1WParams p)
{
    constexpr int J = 1 + 2 * Q, R = (3 * Q + 2) & ~1;
    constexpr int RPW = (J + P1W_CONS - 1) / P1W_CONS;        // rows per consumer warp
    constexpr int AT = J * P1W_TP * 8;                        // doubles of one A tile: [j][point][inducing point]
    extern __shared__ __align__(16) double sm[];              // [2][AT] A tiles, then [2][P1W_TP][DP] Y tiles
    __shared__ double exp_tab[GP_EXP_TAB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = lane >> 2, kk = lane & 3;
    const int g = blockIdx.x % p.G, s = blockIdx.x / p.G;
    const int d0 = blockIdx.y * (8 * P1W_MAXNT);
    const int dcols = (p.D - d0 < 8 * P1W_MAXNT) ? (p.D - d0) : 8 * P1W_MAXNT;     // columns of this pass
    const int nt = (dcols + 7) / 8, DP = 8 * nt;
    double *As = sm, *Ys = sm + 2 * AT;
    const int64_t n_lo = (int64_t)s * p.n_per_split;
    const int64_t n_hi = (n_lo + p.n_per_split < p.n) ? (n_lo + p.n_per_split) : p.n;
    const int nsteps = (int)((n_hi - n_lo + P1W_TP - 1) / P1W_TP);

    gp_exp_load_table(exp_tab);
    for (int idx = threadIdx.x; idx < 2 * P1W_TP * DP; idx += P1W_WARPS * 32) Ys[idx] = 0.0;      // padding columns stay zero
    __syncthreads();

    const int m = g * 8 + gi;
    const bool mvalid = m < p.M;

    // ---- producer: stage 1 of one step into buffer `buf` ---------------------------------------------------------
    double z[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) z[q] = (warp == P1W_CONS && mvalid) ? p.Z[(size_t)m * Q + q] : 0.0;
    auto produce = [&](int step, int buf) {
        const int64_t base = n_lo + (int64_t)step * P1W_TP;
        const int cnt = (int)((n_hi - base < P1W_TP) ? (n_hi - base) : P1W_TP);
        // Y rows of the step (columns d0 .. d0 + dcols): 16-byte copies when the row pieces are 16-byte aligned
        double *yb = Ys + (size_t)buf * P1W_TP * DP;
        if (((p.D | d0 | dcols) & 1) == 0) {
            const int c2 = dcols / 2;
            for (int idx = lane; idx < cnt * c2; idx += 32) {
                const int pt = idx / c2, w = idx - pt * c2;
                p1w_cp_async16(yb + pt * DP + 2 * w, p.Y + (base + pt) * p.D + d0 + 2 * w);
            }
        } else {
            for (int idx = lane; idx < cnt * dcols; idx += 32) {
                const int pt = idx / dcols, w = idx - pt * dcols;
                p1w_cp_async8(yb + pt * DP + w, p.Y + (base + pt) * p.D + d0 + w);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        double *ab = As + (size_t)buf * AT;
#pragma unroll 1
        for (int pass = 0; pass < P1W_TP / 4; ++pass) {
            const int pl_raw = pass * 4 + kk;
            const bool valid = mvalid && pl_raw < cnt;
            const int pl = pl_raw < cnt ? pl_raw : (cnt > 0 ? cnt - 1 : 0);    // lanes past the end read a real record, weight 0
            const double2 *rec = reinterpret_cast<const double2 *>(p.rec1 + (base + pl) * R);
            double ad[Q];
            double es0 = 0.0, es1 = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double2 ma = rec[q];                // (mu_q, a_q)
                const double d = ma.x - z[q];
                ad[q] = ma.y * d;
                if (q & 1) es1 = fma(ad[q], d, es1);
                else es0 = fma(ad[q], d, es0);
            }
            const double e = fma(-0.5, es0 + es1, p.rec1[(base + pl) * R + 3 * Q]);
            const double psi = valid ? gp_exp(e, exp_tab) : 0.0;
            double *o = ab + pl_raw * 8 + gi;             // A-fragment order: [j][point][inducing point]
            o[0] = psi;
#pragma unroll
            for (int q = 0; q < Q; ++q) o[(size_t)(1 + q) * (P1W_TP * 8)] = psi * ad[q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double2 v2 = rec[Q + (q >> 1)];     // (v1_2k, v1_2k+1)
                o[(size_t)(1 + Q + q) * (P1W_TP * 8)] = psi * fma(ad[q], ad[q], (q & 1) ? v2.y : v2.x);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    };

    // ---- consumers: accumulators of rows cw + 7 r, all column tiles ----------------------------------------------
    double C[RPW][P1W_MAXNT][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int t = 0; t < P1W_MAXNT; ++t) { C[r][t][0] = 0.0; C[r][t][1] = 0.0; }

    if (warp == P1W_CONS && nsteps > 0) produce(0, 0);
    __syncthreads();
    for (int st = 0; st < nsteps; ++st) {
        const int buf = st & 1;
        if (warp == P1W_CONS) {
            if (st + 1 < nsteps) produce(st + 1, buf ^ 1);
        } else {
            const double *ab = As + (size_t)buf * AT, *yb = Ys + (size_t)buf * P1W_TP * DP;
#pragma unroll 2
            for (int ks = 0; ks < P1W_TP / 4; ++ks) {
                const int pt = 4 * ks + kk;
                double b[P1W_MAXNT];
#pragma unroll
                for (int t = 0; t < P1W_MAXNT; ++t) b[t] = (t < nt) ? yb[pt * DP + 8 * t + gi] : 0.0;     // B: (point kk, column gi)
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int j = warp + P1W_CONS * r;
                    if (j < J) {                                                                         // warp-uniform
                        const double a = ab[(size_t)j * (P1W_TP * 8) + pt * 8 + gi];                      // A: (inducing point gi, point kk)
#pragma unroll
                        for (int t = 0; t < P1W_MAXNT; ++t)
                            if (t < nt) p1w_dmma(C[r][t], a, b[t]);
                    }
                }
            }
        }
        __syncthreads();
    }

    // C fragment: (inducing point gi, columns 2 kk and 2 kk + 1 of tile t)
    if (warp < P1W_CONS && mvalid) {
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int j = warp + P1W_CONS * r;
            if (j >= J) continue;
            double *out = p.partial + ((size_t)s * p.M * J + (size_t)m * J + j) * p.D + d0;
#pragma unroll
            for (int t = 0; t < P1W_MAXNT; ++t) {
                const int col = 8 * t + 2 * kk;
                if (t < nt && col < dcols) out[col] = C[r][t][0];
                if (t < nt && col + 1 < dcols) out[col + 1] = C[r][t][1];
            }
        }
    }
}

void gp_psi1_reduce(gparml_ctx *c, int splits);   // psi1.cu

template <int Q>
static int launch_wide_q(gparml_ctx *c, Psi1WParams &p)
{
    constexpr int J = 1 + 2 * Q, AT = J * P1W_TP * 8;
    const int chunks = (c->D + 8 * P1W_MAXNT - 1) / (8 * P1W_MAXNT);
    const int dmax = c->D < 8 * P1W_MAXNT ? c->D : 8 * P1W_MAXNT;
    const int DP = 8 * ((dmax + 7) / 8);
    const size_t smem = ((size_t)2 * AT + (size_t)2 * P1W_TP * DP) * sizeof(double);
    GP_CUDA(cudaFuncSetAttribute(psi1_wide_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // slices: G x S x chunks CTAs (one per SM at a time) in whole waves, >= 8 steps per slice, bounded workspace
    const int64_t slots = c->sm_count;
    const int64_t per_s = (int64_t)p.G * chunks;
    int64_t max_s = (c->n + 8 * P1W_TP - 1) / (8 * P1W_TP);
    const int64_t ws_cap = ((int64_t)256 << 20) / ((int64_t)c->M * J * c->D * (int64_t)sizeof(double));
    if (max_s > ws_cap) max_s = ws_cap;
    if (max_s > 8 * slots / per_s + 1) max_s = 8 * slots / per_s + 1;
    if (max_s < 1) max_s = 1;
    int64_t S = 1;
    double best_eff = -1.0;
    for (int64_t cand = 1; cand <= max_s; ++cand) {
        const int64_t total = per_s * cand, waves = (total + slots - 1) / slots;
        const double eff = (double)total / (double)(waves * slots);
        if (eff > best_eff + 0.02) { best_eff = eff; S = cand; }
    }
    int64_t per = (c->n + S - 1) / S;
    per = (per + P1W_TP - 1) / P1W_TP * P1W_TP;
    if (per < P1W_TP) per = P1W_TP;
    p.n_per_split = per;
    p.S = (int)((c->n + per - 1) / per);
    if (p.S < 1) p.S = 1;
    GP_TRY(gp_ensure_ws(c, (size_t)p.S * c->M * J * c->D * sizeof(double)));
    p.partial = c->ws;
    dim3 grid((unsigned)(p.G * p.S), chunks);
    psi1_wide_kernel<Q><<<grid, P1W_WARPS * 32, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    gp_psi1_reduce(c, p.S);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_psi1_stats_wide(gparml_ctx *c)
{
    Psi1WParams p;
    p.rec1 = c->rec1; p.Y = c->Y; p.Z = c->Z;
    p.n = c->n; p.M = c->M; p.D = c->D;
    p.G = (c->M + 7) / 8;
    switch (c->Q) {
#define CASE_Q(q) case q: return launch_wide_q<q>(c, p);
        CASE_Q(1) CASE_Q(2) CASE_Q(3) CASE_Q(4) CASE_Q(5) CASE_Q(6) CASE_Q(7) CASE_Q(8)
        CASE_Q(9) CASE_Q(10) CASE_Q(11) CASE_Q(12) CASE_Q(13) CASE_Q(14) CASE_Q(15) CASE_Q(16)
#undef CASE_Q
    }
    gp_set_error("psi1_stats: unsupported Q=%d (1..%d)", c->Q, GP_MAX_Q);
    return GPARML_ERR_ARG;
}
