// K1 for wide outputs (D > 16): psi1_stats with the stage-1 values shared by the whole CTA.
//
// Replaces the same reference lines as psi1_mma.cu (kernel_exp.py:13-82, partial_terms.py:162-188, :256-271) and
// writes the same per-slice partial layout, so psi1_reduce_kernel finishes both.
//
// Why a second kernel.  In psi1_mma.cu a warp is an autonomous task that keeps its 1 + 2Q row entries in
// registers and contracts them with at most 16 + 2 output columns; for wider D the columns are chunked and every
// chunk re-evaluates Psi1 and the row entries (D = 50: five times, 69 ms at c4 on B200).  Here a CTA of 16 warps owns
// 8 inducing points and a slice of the points and works bulk-synchronously in steps of TP = 64 points:
//
//   phase A  every warp evaluates Psi1 and the J = 1 + 2Q row entries of 4 points x 8 inducing points ONCE (lane =
//            (inducing point, point)) and stores them in shared memory in A-fragment order;
//   phase B  the warps turn into consumers of the whole tile: warp w owns the rows j = w % 7 + 7 r and every second of
//            the ceil(D / 8) column tiles; per group of 4 points it loads its tiles' B fragments (Y) once, each row's A
//            fragment once, and issues rows x tiles FP64 tensor-core instructions (mma.sync m8n8k4, SASS DMMA) on
//            accumulators that stay in registers for the whole slice.
//
// The next step's Y rows and point records arrive by cp.async during both phases.  Two designs were measured first
// (c4, N = 100k, B200): one / two producer warps feeding 7 / 14 consumers through a double-buffered tile -- 10.5 / 8.1 ms,
// tensor pipe 46 % active, barrier stalls dominant: the producers' dependent DFMA chains queue behind the consumers'
// 16-cycle DMMAs on the SAME FP64 pipe and become the critical path; four producers: 6.7 ms.  Separating the phases
// lets stage 1 run at its 8-cycle dependent latency on an otherwise idle pipe.
//
// Bound: FP64 pipe (DMMA shares it with DFMA, tools/micro/dmma_probe.cu).
#include <math.h>

#include "common.cuh"
#include "gp_exp.cuh"

#define P1W_WARPS 16
#define P1W_RC 7             // row classes: in phase B warp c < 14 owns the rows j = c % 7 + 7 r ...
#define P1W_TC 2             // ... and the column tiles t = c / 7 + 2 u
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

// NS1: warps that evaluate stage 1 (4 points each) -> TP = 4 NS1 points per step
template <int Q, int NS1>
__global__ void __launch_bounds__(P1W_WARPS * 32, 1)
psi1_wide_kernel(Psi1WParams p)
{
    constexpr int J = 1 + 2 * Q, R = (3 * Q + 2) & ~1;
    constexpr int TP = 4 * NS1;
    constexpr int RPW = (J + P1W_RC - 1) / P1W_RC;            // rows per consumer warp
    constexpr int TPW = P1W_MAXNT / P1W_TC;                   // column tiles per consumer warp
    constexpr int AT = J * TP * 8;                            // doubles of the A tile: [j][point][inducing point]
    constexpr int NTH = P1W_WARPS * 32;
    extern __shared__ __align__(16) double sm[];              // [AT] A tile, [2][TP][DP] Y tiles, [2][TP][R] record tiles
    __shared__ double exp_tab[GP_EXP_TAB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = lane >> 2, kk = lane & 3;
    const int g = blockIdx.x % p.G, s = blockIdx.x / p.G;
    const int d0 = blockIdx.y * (8 * P1W_MAXNT);
    const int dcols = (p.D - d0 < 8 * P1W_MAXNT) ? (p.D - d0) : 8 * P1W_MAXNT;     // columns of this pass
    const int nt = (dcols + 7) / 8, DP = 8 * nt;
    double *As = sm, *Ys = sm + AT, *Rs = Ys + 2 * TP * DP;
    const int64_t n_lo = (int64_t)s * p.n_per_split;
    const int64_t n_hi = (n_lo + p.n_per_split < p.n) ? (n_lo + p.n_per_split) : p.n;
    const int nsteps = (int)((n_hi - n_lo + TP - 1) / TP);

    gp_exp_load_table(exp_tab);
    for (int idx = threadIdx.x; idx < 2 * TP * DP; idx += NTH) Ys[idx] = 0.0;      // padding columns / rows past the end stay finite
    __syncthreads();

    const int m = g * 8 + gi;
    const bool mvalid = m < p.M;
    double z[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) z[q] = (warp < NS1 && mvalid) ? p.Z[(size_t)m * Q + q] : 0.0;

    // Y rows (columns d0 .. d0 + dcols) and point records of `step` into buffer step & 1, asynchronously, by all threads
    auto fetch = [&](int step) {
        const int64_t base = n_lo + (int64_t)step * TP;
        const int cnt = (int)((n_hi - base < TP) ? (n_hi - base) : TP);
        double *yb = Ys + (size_t)(step & 1) * TP * DP, *rb = Rs + (size_t)(step & 1) * TP * R;
        if (((p.D | d0 | dcols) & 1) == 0) {              // row pieces are 16-byte aligned
            const int c2 = dcols / 2;
            for (int idx = threadIdx.x; idx < cnt * c2; idx += NTH) {
                const int pt = idx / c2, w = idx - pt * c2;
                p1w_cp_async16(yb + pt * DP + 2 * w, p.Y + (base + pt) * p.D + d0 + 2 * w);
            }
        } else {
            for (int idx = threadIdx.x; idx < cnt * dcols; idx += NTH) {
                const int pt = idx / dcols, w = idx - pt * dcols;
                p1w_cp_async8(yb + pt * DP + w, p.Y + (base + pt) * p.D + d0 + w);
            }
        }
        for (int idx = threadIdx.x; idx < cnt * (R / 2); idx += NTH) p1w_cp_async16(rb + 2 * idx, p.rec1 + base * R + 2 * idx);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double C[RPW][TPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int t = 0; t < TPW; ++t) { C[r][t][0] = 0.0; C[r][t][1] = 0.0; }
    const int rc = warp % P1W_RC, tc = warp / P1W_RC;         // consumer's row class and tile class (warps 14, 15: none)

    if (nsteps > 0) fetch(0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int st = 0; st < nsteps; ++st) {
        const int64_t base = n_lo + (int64_t)st * TP;
        const int cnt = (int)((n_hi - base < TP) ? (n_hi - base) : TP);
        if (st + 1 < nsteps) fetch(st + 1);
        // ---- phase A: stage 1 of 4 points x 8 inducing points per warp -> A tile ---------------------------------
        if (warp < NS1) {
            const double *rt = Rs + (size_t)(st & 1) * TP * R;
            const int pl_raw = warp * 4 + kk;
            const bool valid = mvalid && pl_raw < cnt;
            const int pl = pl_raw < cnt ? pl_raw : cnt - 1;           // lanes past the end read a real record, weight 0
            const double2 *rec = reinterpret_cast<const double2 *>(rt + pl * R);
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
            const double e = fma(-0.5, es0 + es1, rt[pl * R + 3 * Q]);
            const double psi = valid ? gp_exp(e, exp_tab) : 0.0;
            double *o = As + pl_raw * 8 + gi;             // A-fragment order: [j][point][inducing point]
            o[0] = psi;
#pragma unroll
            for (int q = 0; q < Q; ++q) o[(size_t)(1 + q) * (TP * 8)] = psi * ad[q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double2 v2 = rec[Q + (q >> 1)];     // (v1_2k, v1_2k+1)
                o[(size_t)(1 + Q + q) * (TP * 8)] = psi * fma(ad[q], ad[q], (q & 1) ? v2.y : v2.x);
            }
        }
        __syncthreads();
        // ---- phase B: rows x column tiles on the tensor-core instruction ------------------------------------------
        if (warp < P1W_RC * P1W_TC) {
            const double *yb = Ys + (size_t)(st & 1) * TP * DP;
#pragma unroll 2
            for (int ks = 0; ks < TP / 4; ++ks) {
                const int pt = 4 * ks + kk;
                double b[TPW];
#pragma unroll
                for (int u = 0; u < TPW; ++u) {
                    const int t = tc + P1W_TC * u;
                    b[u] = (t < nt) ? yb[pt * DP + 8 * t + gi] : 0.0;                                     // B: (point kk, column gi)
                }
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int j = rc + P1W_RC * r;
                    if (j < J) {                                                                         // warp-uniform
                        const double a = As[(size_t)j * (TP * 8) + pt * 8 + gi];                          // A: (inducing point gi, point kk)
#pragma unroll
                        for (int u = 0; u < TPW; ++u)
                            if (tc + P1W_TC * u < nt) p1w_dmma(C[r][u], a, b[u]);
                    }
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // the next step's tiles (issued before phase A)
        __syncthreads();
    }

    // C fragment: (inducing point gi, columns 2 kk and 2 kk + 1 of tile t)
    if (warp < P1W_RC * P1W_TC && mvalid) {
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int j = rc + P1W_RC * r;
            if (j >= J) continue;
            double *out = p.partial + ((size_t)s * p.M * J + (size_t)m * J + j) * p.D + d0;
#pragma unroll
            for (int u = 0; u < TPW; ++u) {
                const int t = tc + P1W_TC * u;
                const int col = 8 * t + 2 * kk;
                if (t < nt && col < dcols) out[col] = C[r][u][0];
                if (t < nt && col + 1 < dcols) out[col + 1] = C[r][u][1];
            }
        }
    }
}

void gp_psi1_reduce(gparml_ctx *c, int splits);   // psi1.cu

template <int Q>
static int launch_wide_q(gparml_ctx *c, Psi1WParams &p)
{
    constexpr int NS1 = (Q <= 12) ? 16 : 8;               // 64-point steps while the A tile fits shared memory
    constexpr int P1W_TP = 4 * NS1;
    constexpr int J = 1 + 2 * Q, AT = J * P1W_TP * 8;
    const int chunks = (c->D + 8 * P1W_MAXNT - 1) / (8 * P1W_MAXNT);
    const int dmax = c->D < 8 * P1W_MAXNT ? c->D : 8 * P1W_MAXNT;
    const int DP = 8 * ((dmax + 7) / 8);
    constexpr int R = (3 * Q + 2) & ~1;
    const size_t smem = ((size_t)AT + (size_t)2 * P1W_TP * DP + (size_t)2 * P1W_TP * R) * sizeof(double);
    GP_CUDA(cudaFuncSetAttribute(psi1_wide_kernel<Q, NS1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    psi1_wide_kernel<Q, NS1><<<grid, P1W_WARPS * 32, smem, c->stream>>>(p);
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
