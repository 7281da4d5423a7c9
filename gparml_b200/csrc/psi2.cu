// K2: psi2_stats -- the (n, m, m', q) exp-product reduction over n.
//
// Replaces (citations relative to /root/reference)
//   partial_terms.py:45-48 -> kernel_exp.py:126-148   Psi2_n (per point, M x M) and its sum :79
//   partial_terms.py:190-205                          sum_n dPsi2_n/dZ      (M, Q, M)
//   partial_terms.py:273-284                          sum_n dPsi2_n/dalpha  (Q, M, M)
// without ever materialising the (n, M, M) tensor the reference keeps (partial_terms.py:45).
//
// Formulation.  With w_nq = alpha_q / (2 alpha_q S_nq + 1), zbar = (z_m + z_m')/2,
// wd_q = w_nq (mu_nq - zbar_q):
//   Psi2_n[m,m'] = exp( lk[m,m'] + lc2_n - sum_q wd_q (mu_nq - zbar_q) )
//   S0[p]    = sum_n Psi2_n                     -> sum_exp_K_mi_K_im
//   TZ[q][p] = sum_n Psi2_n wd_q                -> symmetric part of dPsi2/dZ
//   TA[q][p] = sum_n Psi2_n (wd_q^2 + v_nq)     -> -alpha_q^2 * (point part of dPsi2/dalpha)
// over the P = M(M+1)/2 pairs m <= m' only (all three are symmetric in (m, m'); the
// antisymmetric Kmm-like part of dPsi2/dZ is point independent and added on expansion).
//
// Mapping.  One thread owns PP pairs (2 for Q <= 10) and keeps their 1 + 2Q accumulators, zbar (Q)
// and wd (Q) in registers; a CTA of 256 threads owns 256 PP pairs and walks its slice of the
// points.  Point records (prep_points) are staged through shared memory by 1-D bulk async copies
// (TMA unit, SASS UBLKCP) into a 2-stage ring guarded by mbarriers; every thread reads the same
// record at the same time, so all shared-memory reads are broadcasts and each read feeds PP
// pairs.  The grid is (pair tiles) x (n splits), sized to whole waves of resident CTAs; per-split
// partial sums go to a workspace and are added in a fixed order (deterministic, no atomics).
// With two pairs per thread the point loop is the hand-ordered software pipeline psi2_step below (B200,
// N = 250k: 6.24 -> 6.11 ms); one pair per thread (Q > 10) keeps the compiler-scheduled loop.  The
// launchers work on point ranges so that gparml_statistics can follow a row-range upload (capi.cu).
//
// Bound: FP64 pipe.  Algorithmic count 6Q + 20 per (point, pair) with exp = 18;
// executed: 6Q + 10 with the table-driven exp of gp_exp.cuh (8 FP64 instructions).
#include <math.h>

#include "common.cuh"

#include "gp_exp.cuh"

#ifndef PSI2_THREADS
#define PSI2_THREADS 256
#endif
#ifndef PSI2_MINB
#define PSI2_MINB 2      // resident CTAs per SM the register budget is tuned for
#endif
#ifndef PSI2_UNROLL
#define PSI2_UNROLL 2    // points in flight per thread
#endif
#ifndef PSI2_TN
#define PSI2_TN 64       // points per stage
#endif
#ifndef PSI2_STAGES
#define PSI2_STAGES 2
#endif

// Pairs per thread (register blocking: each shared-memory read feeds PP pairs) and the number
// of resident CTAs the register budget is tuned for.  Measured on B200 at Q=10 (tools/tune.py,
// N=250k): PP=1/2 CTAs 6.56 ms, PP=2/1 CTA 6.26 ms (LSU wavefronts 74 % -> 43 %, FP64 pipe
// 76 % -> 80 %).  Above Q=10 two pairs no longer fit the 255-register budget.
#ifdef PSI2_PAIRS
template <int Q> struct Psi2Cfg { static constexpr int PP = PSI2_PAIRS; static constexpr int MINB = PSI2_MINB; };
#else
template <int Q> struct Psi2Cfg {
    static constexpr int PP = (Q <= 10) ? 2 : 1;
    static constexpr int MINB = (Q <= 10) ? ((Q <= 3) ? 3 : ((Q <= 6) ? 2 : 1)) : 2;
};
#endif

// One software-pipelined iteration for PP = 2 pairs per thread, written in issue order (see embed_x.cu for
// the register-file argument: a DFMA with three fresh 64-bit register operands takes 3 issue cycles, one
// with an operand from the reuse cache of the previous instruction takes 2):
//   block E  (point i+1): per q  d0, d1 (mu shared), wd0, wd1 (w shared), e0, e1;
//   block XA (exp of point i+1, accumulation of point i): each exp step of the two pairs is followed by
//            half a chunk of accumulations; a chunk = two latent dimensions of one pair:
//            g_a, g_b (wd twice: one register read), then acc1[q], acc1[q+1], acc2[q], acc2[q+1] with psi
//            held in one operand slot (reuse) -- 2 fresh operands each after the first.
struct Psi2Exp {
    double x, t, r, p, tab;
    int k;
};

template <int Q, bool DO_E, bool DO_A>
__device__ __forceinline__ void psi2_step(const double *__restrict__ rn, const double *__restrict__ rc, const double (&lk)[2],
                                          const double (&zb)[2][Q], double (&wdn)[2][Q], const double (&wdc)[2][Q],
                                          const double (&psic)[2], double (&psin)[2], double (&acc)[2][1 + 2 * Q],
                                          const double *exp_tab)
{
    Psi2Exp es[2];
    if (DO_E) {
        const double2 *r = reinterpret_cast<const double2 *>(rn);
        const double lc2 = rn[3 * Q];
        double e[2][2];
        e[0][0] = lk[0]; e[1][0] = lk[1]; e[0][1] = lc2; e[1][1] = lc2;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double2 mw = r[q];           // (mu_q, w_q), broadcast; feeds both pairs
            const double d0 = mw.x - zb[0][q];
            const double d1 = mw.x - zb[1][q];
            wdn[0][q] = mw.y * d0;
            wdn[1][q] = mw.y * d1;
            e[0][q & 1] = fma(-wdn[0][q], d0, e[0][q & 1]);
            e[1][q & 1] = fma(-wdn[1][q], d1, e[1][q & 1]);
        }
        es[0].x = gp_exp_clamp(e[0][0] + e[0][1]);
        es[1].x = gp_exp_clamp(e[1][0] + e[1][1]);
    }
    const double2 *rv = reinterpret_cast<const double2 *>(rc) + Q;      // (v_2k, v_2k+1)
    if (DO_A) {
        acc[0][0] += psic[0];
        acc[1][0] += psic[1];
    }
    // chunk k of pair u: latent dimensions 2k, 2k+1
#define PSI2_CHUNK(k, u)                                                                             \
    if (DO_A && 2 * (k) < Q) {                                                                       \
        constexpr int qa = 2 * (k) < Q ? 2 * (k) : 0, qb = 2 * (k) + 1 < Q ? 2 * (k) + 1 : 0;        \
        const double2 v2 = rv[(k)];                                                                  \
        const double ga = fma(wdc[u][qa], wdc[u][qa], v2.x);                                         \
        const double gb = fma(wdc[u][qb], wdc[u][qb], v2.y);                                         \
        acc[u][1 + qa] = fma(psic[u], wdc[u][qa], acc[u][1 + qa]);                                   \
        if (2 * (k) + 1 < Q) acc[u][1 + qb] = fma(psic[u], wdc[u][qb], acc[u][1 + qb]);              \
        acc[u][1 + Q + qa] = fma(psic[u], ga, acc[u][1 + Q + qa]);                                   \
        if (2 * (k) + 1 < Q) acc[u][1 + Q + qb] = fma(psic[u], gb, acc[u][1 + Q + qb]);              \
    }
#define PSI2_EXP(stmt)                                          \
    if (DO_E) {                                                 \
        _Pragma("unroll") for (int u = 0; u < 2; ++u) { stmt; } \
    }
    PSI2_EXP(es[u].t = fma(es[u].x, GP_EXP_SCALE, GP_EXP_SHIFT))
    PSI2_CHUNK(0, 0)
    PSI2_EXP(es[u].k = __double2loint(es[u].t); es[u].t = es[u].t - GP_EXP_SHIFT)
    PSI2_CHUNK(0, 1)
    PSI2_EXP(es[u].r = fma(es[u].t, GP_EXP_NEG_STEP, es[u].x); es[u].tab = exp_tab[es[u].k & (GP_EXP_TAB - 1)])
    PSI2_CHUNK(1, 0)
    PSI2_EXP(es[u].p = fma(es[u].r, 1.0 / 24.0, 1.0 / 6.0))
    PSI2_CHUNK(1, 1)
    PSI2_EXP(es[u].p = fma(es[u].p, es[u].r, 0.5))
    PSI2_CHUNK(2, 0)
    PSI2_EXP(es[u].p = fma(es[u].p, es[u].r, 1.0))
    PSI2_CHUNK(2, 1)
    PSI2_EXP(es[u].p = fma(es[u].p, es[u].r, 1.0))
    PSI2_CHUNK(3, 0)
    PSI2_EXP(es[u].p = es[u].tab * es[u].p)
    PSI2_CHUNK(3, 1)
    PSI2_CHUNK(4, 0)
    PSI2_CHUNK(4, 1)
    PSI2_CHUNK(5, 0)
    PSI2_CHUNK(5, 1)
    PSI2_CHUNK(6, 0)
    PSI2_CHUNK(6, 1)
    PSI2_CHUNK(7, 0)
    PSI2_CHUNK(7, 1)
#undef PSI2_CHUNK
#undef PSI2_EXP
    if (DO_E) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            int m = es[u].k >> GP_EXP_LOG2_TAB;
            m = m < -1021 ? -1021 : m;
            psin[u] = __hiloint2double(__double2hiint(es[u].p) + (m << 20), __double2loint(es[u].p));
        }
    }
}

template <int Q>
__global__ void __launch_bounds__(PSI2_THREADS, Psi2Cfg<Q>::MINB)
psi2_stats_kernel(const double *__restrict__ rec2, int64_t n, const double *__restrict__ Z, int64_t P,
                  const int2 *__restrict__ pair_idx, const double *__restrict__ pair_lk, int64_t n_per_split,
                  double *__restrict__ partial)
{
    constexpr int R = (3 * Q + 2) & ~1;
    constexpr int NT2 = (Q + 2) / 2;          // double2 loads covering v_0..v_{Q-1} and lc2
    constexpr int UNR = PSI2_UNROLL;
    constexpr int PP = Psi2Cfg<Q>::PP;
    extern __shared__ __align__(16) double tile[];       // [STAGES][TN][R]
    __shared__ __align__(8) uint64_t bar[PSI2_STAGES];
    __shared__ double exp_tab[GP_EXP_TAB];

    const int tid = threadIdx.x;
    int64_t p[PP];
    bool valid[PP];
    double lk[PP], zb[PP][Q], acc[PP][1 + 2 * Q];
#pragma unroll
    for (int u = 0; u < PP; ++u) {
        p[u] = ((int64_t)blockIdx.x * PP + u) * PSI2_THREADS + tid;
        valid[u] = p[u] < P;
        const int2 ab = valid[u] ? pair_idx[p[u]] : make_int2(0, 0);
        lk[u] = valid[u] ? pair_lk[p[u]] : 0.0;
#pragma unroll
        for (int q = 0; q < Q; ++q) zb[u][q] = 0.5 * (Z[ab.x * Q + q] + Z[ab.y * Q + q]);
#pragma unroll
        for (int j = 0; j < 1 + 2 * Q; ++j) acc[u][j] = 0.0;
    }

    const int64_t n_lo = (int64_t)blockIdx.y * n_per_split;
    const int64_t n_hi = (n_lo + n_per_split < n) ? (n_lo + n_per_split) : n;
    const int64_t span = n_hi > n_lo ? n_hi - n_lo : 0;
    const int ntiles = (int)((span + PSI2_TN - 1) / PSI2_TN);

    gp_exp_load_table(exp_tab);
    if (tid == 0) {
        for (int s = 0; s < PSI2_STAGES; ++s) gp_mbar_init(&bar[s], 1);
        gp_fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < PSI2_STAGES && s < ntiles; ++s) {
            const int64_t base = n_lo + (int64_t)s * PSI2_TN;
            const int cnt = (int)((n_hi - base < PSI2_TN) ? (n_hi - base) : PSI2_TN);
            const uint32_t bytes = (uint32_t)cnt * R * sizeof(double);
            gp_mbar_expect_tx(&bar[s], bytes);
            gp_bulk_g2s(tile + (size_t)s * PSI2_TN * R, rec2 + base * R, bytes, &bar[s]);
        }
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % PSI2_STAGES;
        const uint32_t parity = (uint32_t)((t / PSI2_STAGES) & 1);
        const int64_t base = n_lo + (int64_t)t * PSI2_TN;
        const int cnt = (int)((n_hi - base < PSI2_TN) ? (n_hi - base) : PSI2_TN);
        gp_mbar_wait(&bar[s], parity);
        const double *tb = tile + (size_t)s * PSI2_TN * R;
#ifndef PSI2_COMPILER_ORDER
        if constexpr (PP == 2) {
            // software pipeline over the points of the tile (psi2_step): prologue, steady state, epilogue
            double wda[2][Q], wdb[2][Q], psa[2], psb[2];
            if (cnt > 0) {
                psi2_step<Q, true, false>(tb, tb, lk, zb, wda, wda, psa, psa, acc, exp_tab);
                int i = 0;
                for (; i + 2 < cnt; i += 2) {
                    psi2_step<Q, true, true>(tb + (i + 1) * R, tb + i * R, lk, zb, wdb, wda, psa, psb, acc, exp_tab);
                    psi2_step<Q, true, true>(tb + (i + 2) * R, tb + (i + 1) * R, lk, zb, wda, wdb, psb, psa, acc, exp_tab);
                }
                if (i + 1 < cnt) {
                    psi2_step<Q, true, true>(tb + (i + 1) * R, tb + i * R, lk, zb, wdb, wda, psa, psb, acc, exp_tab);
                    psi2_step<Q, false, true>(tb, tb + (i + 1) * R, lk, zb, wda, wdb, psb, psa, acc, exp_tab);
                } else {
                    psi2_step<Q, false, true>(tb, tb + i * R, lk, zb, wdb, wda, psa, psb, acc, exp_tab);
                }
            }
        } else
#endif
        {
#pragma unroll UNR
        for (int i = 0; i < cnt; ++i) {
            const double2 *r = reinterpret_cast<const double2 *>(tb + i * R);
            double wd[PP][Q], e0[PP], e1[PP], psi[PP];
            const double lc2 = tb[i * R + 3 * Q];
#pragma unroll
            for (int u = 0; u < PP; ++u) { e0[u] = lk[u]; e1[u] = lc2; }   // the two partial sums start at lk and lc2
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double2 mw = r[q];           // (mu_q, w_q), broadcast; feeds all PP pairs
#pragma unroll
                for (int u = 0; u < PP; ++u) {
                    const double d = mw.x - zb[u][q];
                    wd[u][q] = mw.y * d;
                    if (q & 1) e1[u] = fma(-wd[u][q], d, e1[u]);
                    else e0[u] = fma(-wd[u][q], d, e0[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < PP; ++u) {
#ifdef GP_USE_LIBM_EXP
                psi[u] = exp(e0[u] + e1[u]);
#else
                psi[u] = gp_exp(e0[u] + e1[u], exp_tab);
#endif
                acc[u][0] += psi[u];
            }
#pragma unroll
            for (int k = 0; k < NT2; ++k) {
                const double2 v2 = r[Q + k];       // (v_2k, v_2k+1); the last slot holds lc2 / padding
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int q = 2 * k + h;
                    if (q < Q) {
                        const double v = h ? v2.y : v2.x;
#pragma unroll
                        for (int u = 0; u < PP; ++u) {
                            acc[u][1 + q] = fma(psi[u], wd[u][q], acc[u][1 + q]);
                            const double g = fma(wd[u][q], wd[u][q], v);
                            acc[u][1 + Q + q] = fma(psi[u], g, acc[u][1 + Q + q]);
                        }
                    }
                }
            }
        }
        }
        __syncthreads();      // every thread is done reading stage s
        if (tid == 0 && t + PSI2_STAGES < ntiles) {
            const int64_t nb = n_lo + (int64_t)(t + PSI2_STAGES) * PSI2_TN;
            const int ncnt = (int)((n_hi - nb < PSI2_TN) ? (n_hi - nb) : PSI2_TN);
            const uint32_t bytes = (uint32_t)ncnt * R * sizeof(double);
            gp_mbar_expect_tx(&bar[s], bytes);
            gp_bulk_g2s(tile + (size_t)s * PSI2_TN * R, rec2 + nb * R, bytes, &bar[s]);
        }
    }

#pragma unroll
    for (int u = 0; u < PP; ++u) {
        if (valid[u]) {
            double *out = partial + (size_t)blockIdx.y * (1 + 2 * Q) * P + p[u];
#pragma unroll
            for (int j = 0; j < 1 + 2 * Q; ++j) out[(size_t)j * P] = acc[u][j];
        }
    }
}

// stats[off_s0 + j * P + p] = sum over splits (fixed order)
__global__ void __launch_bounds__(256) psi2_reduce_kernel(const double *__restrict__ partial, int splits, int64_t rows_x_P,
                                                          double *__restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_x_P) return;
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += partial[(size_t)s * rows_x_P + i];
    dst[i] = a;
}

// number of n-splits for `cnt` points: whole waves of resident CTAs, >= 4 point tiles per split, bounded workspace
template <int Q>
static int plan_q(gparml_ctx *c, int64_t cnt, int *splits_out)
{
    constexpr int R = (3 * Q + 2) & ~1;
    const size_t smem = (size_t)PSI2_STAGES * PSI2_TN * R * sizeof(double);
    int occ = 2;
    GP_CUDA(cudaFuncSetAttribute(psi2_stats_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, psi2_stats_kernel<Q>, PSI2_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t P = c->L.P;
    const int tiles = (int)((P + PSI2_THREADS * Psi2Cfg<Q>::PP - 1) / (PSI2_THREADS * Psi2Cfg<Q>::PP));
    const int64_t slots = (int64_t)c->sm_count * occ;
    const int64_t rows_x_P = (int64_t)(1 + 2 * Q) * P;
    int64_t max_splits = (cnt + 4 * PSI2_TN - 1) / (4 * PSI2_TN);
    const int64_t ws_cap = ((int64_t)512 << 20) / (rows_x_P * (int64_t)sizeof(double)) / GP_MAX_RANGES;   // all ranges of an evaluation share the workspace
    if (max_splits > ws_cap) max_splits = ws_cap;
    if (max_splits > 65535) max_splits = 65535;
    if (max_splits < 1) max_splits = 1;
    int64_t best = 1;
    double best_eff = -1.0;
    for (int64_t s = 1; s <= max_splits; ++s) {
        const int64_t total = (int64_t)tiles * s;
        const int64_t waves = (total + slots - 1) / slots;
        if (waves > 8) break;
        const double eff = (double)total / (double)(waves * slots);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    *splits_out = (int)best;
    return GPARML_OK;
}

// psi2_stats over the points [i0, i1): partial sums into the workspace slices [slice0, slice0 + splits)
template <int Q>
static int launch_range_q(gparml_ctx *c, int64_t i0, int64_t i1, int slice0, int splits)
{
    constexpr int R = (3 * Q + 2) & ~1;
    const size_t smem = (size_t)PSI2_STAGES * PSI2_TN * R * sizeof(double);
    const int64_t P = c->L.P, cnt = i1 - i0;
    const int tiles = (int)((P + PSI2_THREADS * Psi2Cfg<Q>::PP - 1) / (PSI2_THREADS * Psi2Cfg<Q>::PP));
    const int64_t rows_x_P = (int64_t)(1 + 2 * Q) * P;
    const int64_t n_per_split = (cnt + splits - 1) / splits;
    dim3 grid(tiles, splits);
    psi2_stats_kernel<Q><<<grid, PSI2_THREADS, smem, c->stream>>>(c->rec2 + i0 * R, cnt, c->Z, P, c->pair_idx, c->pair_lk, n_per_split,
                                                                  c->ws + (size_t)slice0 * rows_x_P);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

#define PSI2_ALL_Q(F) F(1) F(2) F(3) F(4) F(5) F(6) F(7) F(8) F(9) F(10) F(11) F(12) F(13) F(14) F(15) F(16)

int gp_psi2_plan_range(gparml_ctx *c, int64_t cnt, int *splits)
{
    switch (c->Q) {
#define CASE_Q(q) case q: return plan_q<q>(c, cnt, splits);
        PSI2_ALL_Q(CASE_Q)
#undef CASE_Q
    }
    gp_set_error("psi2_stats: unsupported Q=%d (1..%d)", c->Q, GP_MAX_Q);
    return GPARML_ERR_ARG;
}

int gp_launch_psi2_stats_range(gparml_ctx *c, int64_t i0, int64_t i1, int slice0, int splits)
{
    switch (c->Q) {
#define CASE_Q(q) case q: return launch_range_q<q>(c, i0, i1, slice0, splits);
        PSI2_ALL_Q(CASE_Q)
#undef CASE_Q
    }
    gp_set_error("psi2_stats: unsupported Q=%d (1..%d)", c->Q, GP_MAX_Q);
    return GPARML_ERR_ARG;
}

// fixed-order sum over the workspace slices -> packed statistics (S0, TZ, TA)
int gp_launch_psi2_reduce(gparml_ctx *c, int slices)
{
    const int64_t rows_x_P = (int64_t)(1 + 2 * c->Q) * c->L.P;
    psi2_reduce_kernel<<<(int)((rows_x_P + 255) / 256), 256, 0, c->stream>>>(c->ws, slices, rows_x_P, c->stats + c->L.off_s0);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_psi2_stats_f32(gparml_ctx *c);

int gp_launch_psi2_stats(gparml_ctx *c)
{
    if (c->flags & GPARML_FLAG_FP32_MAP) return gp_launch_psi2_stats_f32(c);      // opt-in fp32 evaluation
    int splits = 1;
    GP_TRY(gp_psi2_plan_range(c, c->n, &splits));
    GP_TRY(gp_ensure_ws(c, (size_t)splits * (1 + 2 * c->Q) * c->L.P * sizeof(double)));
    GP_TRY(gp_launch_psi2_stats_range(c, 0, c->n, 0, splits));
    return gp_launch_psi2_reduce(c, splits);
}
