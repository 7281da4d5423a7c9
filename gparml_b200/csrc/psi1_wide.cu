// K1 for wide outputs (D > 16): psi1_stats with the stage-1 values shared by the whole CTA.
//
// Replaces the same reference lines as psi1_mma.cu (kernel_exp.py:13-82, partial_terms.py:162-188, :256-271) and
// writes the same per-slice partial layout, so psi1_reduce_kernel finishes both.
//
// Why a second kernel.  In psi1_mma.cu a warp is an autonomous task that keeps its 1 + 2Q row entries in
// registers and contracts them with at most 16 + 2 output columns; for wider D the columns are chunked and every
// chunk re-evaluates Psi1 and the row entries (D = 50: five times, 69 ms at c4 on B200).  Here a CTA of 16 warps owns
// 8 inducing points and a slice of the points and works in steps of TP = 32 points on double-buffered tiles:
//
//   stage 1  warps 0..7 evaluate Psi1 and the J = 1 + 2Q row entries of 4 points x 8 inducing points of the NEXT step
//            ONCE (lane = (inducing point, point)) and store them in shared memory in A-fragment order;
//   MMAs     all warps consume the CURRENT step's tile: warp w owns every fourth row j and every fourth of the
//            ceil(D / 8) column tiles (classes chosen so that every SM sub-partition gets the same number of MMAs); per group of 4 points it loads its tiles' B fragments (Y) once, each row's A
//            fragment once, and issues rows x tiles FP64 tensor-core instructions (mma.sync m8n8k4, SASS DMMA) on
//            accumulators that stay in registers for the whole slice.
//
// Y rows and point records of later steps arrive by cp.async; one block barrier per step.  Designs measured on the way
// (c4, N = 100k, B200; tools/micro/dmma_feed.cu shows that 3 x 4 DMMAs per 4 points fed from shared memory run at the
// full 36.7 TFLOP/s with 8 or 16 warps): dedicated producer warps (1 / 2 / 4) feeding 7 / 14 / 12 consumers: 10.5 / 8.1 /
// 6.7 ms -- the producers' dependent DFMA chains queue behind the consumers' 16-cycle DMMAs on the SAME FP64 pipe and
// become the critical path (tensor pipe 46 % active, barrier stalls dominant); separate stage-1 and MMA phases of 64
// points: 6.9 ms (stage 1 + barrier 5.4k of 19.8k cycles per step with the pipe idle).
//
// Bound: FP64 pipe (DMMA shares it with DFMA, tools/micro/dmma_probe.cu).
#include <math.h>
#include <stdio.h>

#include "common.cuh"
#include "gp_exp.cuh"

#define P1W_WARPS 16
#define P1W_RC 4             // row classes: warp w = 4 a + b owns the rows j = a + 4 r ...
#define P1W_TC 4             // ... and the column tiles t = ((b - a) mod 4) + 4 u: the four warps of an SM sub-partition
                             // (same b) then hold one warp of every row class and of every tile class, which balances the
                             // DMMA work over the sub-partitions (J = 21, 7 tiles: 37 / 37 / 37 / 36 row-tiles per 4 points)
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

#define P1W_NS1 8            // warps that also evaluate stage 1 (4 points each): TP = 32 points per step
#define P1W_TP (4 * P1W_NS1)

template <int Q>
__global__ void __launch_bounds__(P1W_WARPS * 32, 1)
psi1_wide_kernel(Psi1WParams p)
{
    constexpr int J = 1 + 2 * Q, R = (3 * Q + 2) & ~1;
    constexpr int RS = R + 2;                                 // padded record stride: the 4 records a warp reads hit distinct banks
    constexpr int TP = P1W_TP;
    constexpr int RPW = (J + P1W_RC - 1) / P1W_RC;            // rows per consumer warp
    constexpr int TPW = P1W_MAXNT / P1W_TC;                   // column tiles per consumer warp
    constexpr int AT = J * TP * 8;                            // doubles of one A tile: [j][point][inducing point]
    constexpr int NTH = P1W_WARPS * 32;
    extern __shared__ __align__(16) double sm[];              // [2][AT] A tiles, [2][TP][DP] Y tiles, [2][TP][RS] record tiles
    __shared__ double exp_tab[GP_EXP_TAB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = lane >> 2, kk = lane & 3;
    const int g = blockIdx.x % p.G, s = blockIdx.x / p.G;
    const int d0 = blockIdx.y * (8 * P1W_MAXNT);
    const int dcols = (p.D - d0 < 8 * P1W_MAXNT) ? (p.D - d0) : 8 * P1W_MAXNT;     // columns of this pass
    const int nt = (dcols + 7) / 8, DP = 8 * nt;
    double *As = sm, *Ys = sm + 2 * AT, *Rs = Ys + 2 * TP * DP;
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
    for (int q = 0; q < Q; ++q) z[q] = (warp < P1W_NS1 && mvalid) ? p.Z[(size_t)m * Q + q] : 0.0;
    const bool y16 = ((p.D | d0 | dcols) & 1) == 0;           // Y row pieces are 16-byte aligned

    // asynchronous tile loads, one row (point) per warp and iteration: no integer division in the loops
    auto fetch_y = [&](int step) {
        if (step >= nsteps) return;
        const int64_t base = n_lo + (int64_t)step * TP;
        const int cnt = (int)((n_hi - base < TP) ? (n_hi - base) : TP);
        double *yb = Ys + (size_t)(step & 1) * TP * DP;
        for (int pt = warp; pt < cnt; pt += P1W_WARPS) {
            const double *src = p.Y + (base + pt) * p.D + d0;
            if (y16) {
                for (int w = lane; 2 * w < dcols; w += 32) p1w_cp_async16(yb + pt * DP + 2 * w, src + 2 * w);
            } else {
                for (int w = lane; w < dcols; w += 32) p1w_cp_async8(yb + pt * DP + w, src + w);
            }
        }
    };
    auto fetch_rec = [&](int step) {
        if (step >= nsteps) return;
        const int64_t base = n_lo + (int64_t)step * TP;
        const int cnt = (int)((n_hi - base < TP) ? (n_hi - base) : TP);
        double *rb = Rs + (size_t)(step & 1) * TP * RS;
        for (int pt = warp; pt < cnt; pt += P1W_WARPS)
            if (lane < R / 2) p1w_cp_async16(rb + pt * RS + 2 * lane, p.rec1 + (base + pt) * R + 2 * lane);      // R / 2 <= 25 pairs
    };
    // stage 1 of 4 points x 8 inducing points (this warp's share of `step`) -> A tile step & 1
    auto stage1 = [&](int step) {
        const int64_t base = n_lo + (int64_t)step * TP;
        const int cnt = (int)((n_hi - base < TP) ? (n_hi - base) : TP);
        const double *rt = Rs + (size_t)(step & 1) * TP * RS;
        const int pl_raw = warp * 4 + kk;
        const bool valid = mvalid && pl_raw < cnt;
        const int pl = pl_raw < cnt ? pl_raw : cnt - 1;           // lanes past the end read a real record, weight 0
        const double2 *rec = reinterpret_cast<const double2 *>(rt + pl * RS);
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
        const double e = fma(-0.5, es0 + es1, rt[pl * RS + 3 * Q]);
        const double psi = valid ? gp_exp(e, exp_tab) : 0.0;
        double *o = As + (size_t)(step & 1) * AT + pl_raw * 8 + gi;      // A-fragment order: [j][point][inducing point]
        o[0] = psi;
#pragma unroll
        for (int q = 0; q < Q; ++q) o[(size_t)(1 + q) * (TP * 8)] = psi * ad[q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double2 v2 = rec[Q + (q >> 1)];     // (v1_2k, v1_2k+1)
            o[(size_t)(1 + Q + q) * (TP * 8)] = psi * fma(ad[q], ad[q], (q & 1) ? v2.y : v2.x);
        }
    };

    double C[RPW][TPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int t = 0; t < TPW; ++t) { C[r][t][0] = 0.0; C[r][t][1] = 0.0; }
    const int rc = warp >> 2, tc = ((warp & 3) - rc) & 3;     // consumer's row class and tile class

    // prologue: records of steps 0 and 1, Y of step 0; stage 1 of step 0
    fetch_rec(0);
    fetch_rec(1);
    fetch_y(0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (warp < P1W_NS1 && nsteps > 0) stage1(0);
    __syncthreads();
#ifdef P1W_PROFILE
    long long tF = 0, tA = 0, tB = 0, tW = 0, c0, c1, c2, c3, c4;
#define P1W_CLK(x) x = clock64()
#else
#define P1W_CLK(x)
#endif
    for (int st = 0; st < nsteps; ++st) {
        P1W_CLK(c0);
        // copies for later steps (their buffers were last read before the barrier that ended step st - 1)
        fetch_rec(st + 2);
        fetch_y(st + 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        P1W_CLK(c1);
        // stage 1 of the NEXT step (8 of the 16 warps) runs while the other warps already issue this step's MMAs: the FP64
        // pipe always has tensor work queued, and no warp's stage-1 chain is longer than one group of 4 points
        if (warp < P1W_NS1 && st + 1 < nsteps) stage1(st + 1);
        P1W_CLK(c2);
        {
            const double *ab = As + (size_t)(st & 1) * AT, *yb = Ys + (size_t)(st & 1) * TP * DP;
#pragma unroll 4
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
                        const double a = ab[(size_t)j * (TP * 8) + pt * 8 + gi];                          // A: (inducing point gi, point kk)
#pragma unroll
                        for (int u = 0; u < TPW; ++u)
                            if (tc + P1W_TC * u < nt) p1w_dmma(C[r][u], a, b[u]);
                    }
                }
            }
        }
        P1W_CLK(c3);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#ifdef P1W_PROFILE
        c4 = clock64();
        tF += c1 - c0; tA += c2 - c1; tB += c3 - c2; tW += c4 - c3;
#endif
    }
#ifdef P1W_PROFILE
    if (blockIdx.x == 5 && blockIdx.y == 0 && lane == 0)
        printf("warp %2d steps %d: per step fetch %lld  stage1 %lld  MMAs %lld  wait+barrier %lld\n", warp, nsteps, tF / nsteps, tA / nsteps,
               tB / nsteps, tW / nsteps);
#endif

    // C fragment: (inducing point gi, columns 2 kk and 2 kk + 1 of tile t)
    if (mvalid) {
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
    constexpr int J = 1 + 2 * Q, AT = J * P1W_TP * 8;
    const int chunks = (c->D + 8 * P1W_MAXNT - 1) / (8 * P1W_MAXNT);
    const int dmax = c->D < 8 * P1W_MAXNT ? c->D : 8 * P1W_MAXNT;
    const int DP = 8 * ((dmax + 7) / 8);
    constexpr int R = (3 * Q + 2) & ~1;
    const size_t smem = ((size_t)2 * AT + (size_t)2 * P1W_TP * DP + (size_t)2 * P1W_TP * (R + 2)) * sizeof(double);
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
