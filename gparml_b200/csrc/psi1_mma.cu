// K1: psi1_stats on the FP64 tensor-core path -- Psi1 and its Y-contractions.
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
// and every row is contracted with Y over the points: C[row, d] = sum_n A[n, row] Y[n, d] -- the
// one dense contraction of the path (75 % of this kernel's FP64 work at D = 10).
//
// Mapping.  The contraction runs on the FP64 tensor-core instruction (mma.sync m8n8k4 f64, SASS
// DMMA): 8 inducing points x 8 output columns x 4 points per instruction.  Lane (g = lane / 4,
// k = lane % 4) of a warp evaluates Psi1 and its 1 + 2Q row entries for (inducing point g of the
// warp's group of 8, point k of the current group of 4) in registers -- which is exactly the
// A-fragment layout of the instruction, so the row entries go from the FP64 pipe into the MMA
// without touching shared memory; the B fragment is Y[point k][column g].  On B200 DMMA has the
// DFMA rate (tools/micro/dmma_probe.cu: 37.1 TFLOP/s, same pipe) but needs 4 register operands per
// 256 FMAs instead of 3 per 32, so it is not limited by register-file bandwidth the way a DFMA
// contraction is (tools/micro/dfma_regbw.cu).  Output columns beyond the 8-wide tiles (D = 10: two)
// are accumulated per lane with DFMA and reduced over the 4 lanes of a group at the end, which
// costs 2 pipe cycles per column instead of 16 for a zero-padded tile.
//
// A warp is an autonomous task (group of 8 inducing points, slice of the points): it streams its
// slice through a private double-buffered shared-memory tile with 16-byte cp.async (LDGSTS), no
// block-wide barrier anywhere; tasks = groups x slices are sized to one wave of 8 warps per SM.
// Partial results per slice go to a workspace and are summed in a fixed order (psi1_reduce_kernel).
//
// Bound: FP64 pipe.
#include <math.h>

#include "common.cuh"
#include "gp_exp.cuh"

#define P1M_WARPS 8
#ifndef P1M_TP
#define P1M_TP 16        // points per tile (4 MMA steps)
#endif

struct Psi1MParams {
    const double *rec1, *Y, *Z;
    int64_t n, n_per_split;
    int M, D, G, S;      // G groups of 8 inducing points, S point slices
    int d_chunk;         // output columns per chunk (blockIdx.y selects the chunk)
    double *partial;     // [S][M * (1+2Q)][D]
};

__device__ __forceinline__ void p1m_cp_async8(double *dst_smem, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(gp_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void p1m_cp_async16(double *dst_smem, const double *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gp_smem_u32(dst_smem)), "l"(src) : "memory");
}
// D (8x8, 2 per lane) += A (8x4, 1 per lane) * B (4x8, 1 per lane)
__device__ __forceinline__ void p1m_dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// NT: 8-column MMA tiles, DR: extra columns accumulated with DFMA; chunk width 8 NT + DR
template <int Q, int NT, int DR>
__global__ void __launch_bounds__(P1M_WARPS * 32, 1)
psi1_mma_kernel(Psi1MParams p)
{
    constexpr int J = 1 + 2 * Q, R = (3 * Q + 2) & ~1;
    constexpr int RS = R + 2;                             // padded record stride: the 4 records of a step hit distinct banks
    constexpr int DC = 8 * NT + DR, YS = (DC + 1) & ~1;   // Y tile row stride (16-byte rows)
    constexpr int NTA = NT > 0 ? NT : 1, DRA = DR > 0 ? DR : 1;
    constexpr int TILE = P1M_TP * RS + P1M_TP * YS;       // doubles per buffer
    extern __shared__ __align__(16) double sm[];          // [warps][2][TILE]
    __shared__ double exp_tab[GP_EXP_TAB];
    gp_exp_load_table(exp_tab);
    __syncthreads();                                      // the only block-wide barrier

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * P1M_WARPS + warp;
    const int g = task % p.G, s = task / p.G;
    if (s >= p.S) return;
    const int gi = lane >> 2, kk = lane & 3;
    const int m = g * 8 + gi;
    const bool mvalid = m < p.M;
    const int d0 = blockIdx.y * p.d_chunk;
    const int dcols = (p.D - d0 < DC) ? (p.D - d0) : DC;  // valid output columns of this chunk
    const int64_t n_lo = (int64_t)s * p.n_per_split;
    const int64_t n_hi = (n_lo + p.n_per_split < p.n) ? (n_lo + p.n_per_split) : p.n;

    double *wb = sm + (size_t)warp * 2 * TILE;
    for (int idx = lane; idx < 2 * TILE; idx += 32) wb[idx] = 0.0;      // Y columns >= dcols stay zero
    __syncwarp();

    double z[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) z[q] = mvalid ? p.Z[(size_t)m * Q + q] : 0.0;
    double C[NTA][J][2], X[DRA][J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int t = 0; t < NTA; ++t) { C[t][j][0] = 0.0; C[t][j][1] = 0.0; }
#pragma unroll
        for (int c = 0; c < DRA; ++c) X[c][j] = 0.0;
    }

    // asynchronous tile loader: records are contiguous (re-strided to RS), Y rows are strided by D
    auto issue = [&](int64_t base, int buf) {
        const int cnt = (int)((n_hi - base < P1M_TP) ? (n_hi - base) : P1M_TP);
        double *rb = wb + buf * TILE, *yb = rb + P1M_TP * RS;
        for (int idx = lane; idx < cnt * (R / 2); idx += 32) {
            const int pt = idx / (R / 2), w = idx % (R / 2);          // compile-time divisor
            p1m_cp_async16(rb + pt * RS + 2 * w, p.rec1 + (base + pt) * R + 2 * w);
        }
        if (p.D == YS && DC == YS) {
            // single chunk and unpadded rows: the Y tile is one contiguous, 16-byte aligned block
            for (int idx = lane; idx < cnt * (YS / 2); idx += 32) p1m_cp_async16(yb + 2 * idx, p.Y + base * p.D + 2 * idx);
        } else {
            for (int idx = lane; idx < cnt * DC; idx += 32) {
                const int pt = idx / DC, dd = idx % DC;               // compile-time divisor
                if (dd < dcols) p1m_cp_async8(yb + pt * YS + dd, p.Y + (base + pt) * p.D + d0 + dd);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int ntiles = (int)((n_hi - n_lo + P1M_TP - 1) / P1M_TP);
    if (ntiles > 0) issue(n_lo, 0);
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const int64_t base = n_lo + (int64_t)t * P1M_TP;
        const int cnt = (int)((n_hi - base < P1M_TP) ? (n_hi - base) : P1M_TP);
        __syncwarp();                                     // every lane is done with the buffer refilled next
        if (t + 1 < ntiles) {
            issue(base + P1M_TP, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();                                     // tile t visible to the whole warp
        const double *rb = wb + buf * TILE, *yb = rb + P1M_TP * RS;
#pragma unroll 1      // unrolling by 2 changes nothing: 255 registers, the steps stay serialised
        for (int st = 0; st < P1M_TP / 4; ++st) {
            if (st * 4 >= cnt) break;                     // warp-uniform
            const int pl_raw = st * 4 + kk;
            const bool valid = mvalid && pl_raw < cnt;
            const int pl = pl_raw < cnt ? pl_raw : cnt - 1;   // lanes past the end work on a real record, weight 0
            const double2 *rec = reinterpret_cast<const double2 *>(rb + pl * RS);
            const double *yr = yb + pl * YS;
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
            const double e = fma(-0.5, es0 + es1, rb[pl * RS + 3 * Q]);
            const double psi = valid ? gp_exp(e, exp_tab) : 0.0;
            double yt[NTA], yx[DRA];
#pragma unroll
            for (int tt = 0; tt < NT; ++tt) yt[tt] = yr[8 * tt + gi];         // B fragment: (point kk, column gi)
#pragma unroll
            for (int c = 0; c < DR; ++c) yx[c] = yr[8 * NT + c];
            // row 0
#pragma unroll
            for (int tt = 0; tt < NT; ++tt) p1m_dmma(C[tt][0], psi, yt[tt]);
#pragma unroll
            for (int c = 0; c < DR; ++c) X[c][0] = fma(psi, yx[c], X[c][0]);
            // rows 1 .. Q
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double a1 = psi * ad[q];
#pragma unroll
                for (int tt = 0; tt < NT; ++tt) p1m_dmma(C[tt][1 + q], a1, yt[tt]);
#pragma unroll
                for (int c = 0; c < DR; ++c) X[c][1 + q] = fma(a1, yx[c], X[c][1 + q]);
            }
            // rows 1 + Q .. 2Q
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double2 v2 = rec[Q + (q >> 1)];     // (v1_2k, v1_2k+1)
                const double a2 = psi * fma(ad[q], ad[q], (q & 1) ? v2.y : v2.x);
#pragma unroll
                for (int tt = 0; tt < NT; ++tt) p1m_dmma(C[tt][1 + Q + q], a2, yt[tt]);
#pragma unroll
                for (int c = 0; c < DR; ++c) X[c][1 + Q + q] = fma(a2, yx[c], X[c][1 + Q + q]);
            }
        }
    }

    // C fragment: (inducing point gi, columns 2 kk and 2 kk + 1 of tile tt); X: partial over this lane's points
    double *out = p.partial + ((size_t)s * p.M * J + (size_t)(mvalid ? m : 0) * J) * p.D + d0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) {
            const int col = 8 * tt + 2 * kk;
            if (mvalid && col < dcols) out[(size_t)j * p.D + col] = C[tt][j][0];
            if (mvalid && col + 1 < dcols) out[(size_t)j * p.D + col + 1] = C[tt][j][1];
        }
#pragma unroll
        for (int c = 0; c < DR; ++c) {
            double v = X[c][j];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (mvalid && kk == 0 && 8 * NT + c < dcols) out[(size_t)j * p.D + 8 * NT + c] = v;
        }
    }
}

void gp_psi1_reduce(gparml_ctx *c, int splits);   // psi1.cu: fixed-order sum over slices + scatter to the packed buffer

template <int Q, int NT, int DR>
static int launch_cfg(gparml_ctx *c, Psi1MParams &p, int dchunks)
{
    constexpr int R = (3 * Q + 2) & ~1, RS = R + 2, DC = 8 * NT + DR, YS = (DC + 1) & ~1;
    constexpr int TILE = P1M_TP * RS + P1M_TP * YS;
    const size_t smem = (size_t)P1M_WARPS * 2 * TILE * sizeof(double);
    GP_CUDA(cudaFuncSetAttribute(psi1_mma_kernel<Q, NT, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.d_chunk = DC;
    const int64_t tasks = (int64_t)p.G * p.S;
    dim3 grid((unsigned)((tasks + P1M_WARPS - 1) / P1M_WARPS), dchunks);
    psi1_mma_kernel<Q, NT, DR><<<grid, P1M_WARPS * 32, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// chunking of the D output columns: the cheapest of the instantiated (NT, DR) shapes, cost in FP64-pipe
// cycles per 32 (point, inducing point) items: stage 1 (6Q + 11 instructions) + J (16 NT + 2 DR) per chunk
template <int Q>
static int launch_q(gparml_ctx *c, Psi1MParams &p)
{
    constexpr int J = 1 + 2 * Q;
    constexpr bool wide = Q <= 10;      // accumulators J (2 NT + DR) doubles per lane: <= 84
    struct Cfg { int nt, dr; };
    const Cfg cfgs[5] = {{0, 1}, {0, 2}, {1, 0}, {1, 2}, {2, 0}};
    int best = -1, best_chunks = 1;
    double best_cost = 0.0;
    for (int i = 0; i < 5; ++i) {
        if (!wide && (2 * cfgs[i].nt + cfgs[i].dr) > 2) continue;
        const int cap = 8 * cfgs[i].nt + cfgs[i].dr;
        const int chunks = (c->D + cap - 1) / cap;
        const double cost = chunks * (2.0 * (6 * Q + 11) + J * (16.0 * cfgs[i].nt + 2.0 * cfgs[i].dr));
        if (best < 0 || cost < best_cost) { best = i; best_cost = cost; best_chunks = chunks; }
    }
    // slices: tasks = groups x slices x column chunks fill whole waves of P1M_WARPS warps per SM; among the
    // slice counts up to four waves the fewest that come within 2 % of the best fill, >= 4 tiles per slice
    const int64_t warps = (int64_t)c->sm_count * P1M_WARPS;
    const int64_t per_s = (int64_t)p.G * best_chunks;           // tasks per slice
    int64_t max_s = (c->n + 4 * P1M_TP - 1) / (4 * P1M_TP);
    const int64_t ws_cap = ((int64_t)256 << 20) / ((int64_t)c->M * J * c->D * (int64_t)sizeof(double));
    if (max_s > ws_cap) max_s = ws_cap;
    if (max_s > 4 * warps / per_s + 1) max_s = 4 * warps / per_s + 1;
    if (max_s < 1) max_s = 1;
    int64_t S = 1;
    double best_eff = -1.0;
    for (int64_t cand = 1; cand <= max_s; ++cand) {
        const int64_t total = per_s * cand, waves = (total + warps - 1) / warps;
        const double eff = (double)total / (double)(waves * warps);
        if (eff > best_eff + 0.02) { best_eff = eff; S = cand; }
    }
    if (S < 1) S = 1;
    int64_t per = (c->n + S - 1) / S;
    per = (per + P1M_TP - 1) / P1M_TP * P1M_TP;          // slices start on tile boundaries (16-byte aligned Y tiles)
    if (per < P1M_TP) per = P1M_TP;
    p.n_per_split = per;
    p.S = (int)((c->n + per - 1) / per);
    if (p.S < 1) p.S = 1;
    GP_TRY(gp_ensure_ws(c, (size_t)p.S * c->M * J * c->D * sizeof(double)));
    p.partial = c->ws;
    int r = GPARML_ERR_ARG;
    switch (best) {
        case 0: r = launch_cfg<Q, 0, 1>(c, p, best_chunks); break;
        case 1: r = launch_cfg<Q, 0, 2>(c, p, best_chunks); break;
        case 2: r = launch_cfg<Q, 1, 0>(c, p, best_chunks); break;
        case 3: r = launch_cfg<Q, (Q <= 10 ? 1 : 0), 2>(c, p, best_chunks); break;
        case 4: r = launch_cfg<Q, (Q <= 10 ? 2 : 1), 0>(c, p, best_chunks); break;
    }
    GP_TRY(r);
    gp_psi1_reduce(c, p.S);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_psi1_stats_mma(gparml_ctx *c)
{
    Psi1MParams p;
    p.rec1 = c->rec1; p.Y = c->Y; p.Z = c->Z;
    p.n = c->n; p.M = c->M; p.D = c->D;
    p.G = (c->M + 7) / 8;
    switch (c->Q) {
#define CASE_Q(q) case q: return launch_q<q>(c, p);
        CASE_Q(1) CASE_Q(2) CASE_Q(3) CASE_Q(4) CASE_Q(5) CASE_Q(6) CASE_Q(7) CASE_Q(8)
        CASE_Q(9) CASE_Q(10) CASE_Q(11) CASE_Q(12) CASE_Q(13) CASE_Q(14) CASE_Q(15) CASE_Q(16)
#undef CASE_Q
    }
    gp_set_error("psi1_stats: unsupported Q=%d (1..%d)", c->Q, GP_MAX_Q);
    return GPARML_ERR_ARG;
}
