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
// Two kernels:
//   psi2x_stats (Q <= 10, below)   5Q + 10 executed FP64 instructions per (point, pair), two pairs per thread;
//   psi2_stats  (Q = 11 .. 16)     6Q + 10, one pair per thread (two no longer fit the 255-register budget).
// Mapping (both).  One thread owns its pairs and keeps their 1 + 2Q accumulators and zbar (Q) in registers; a CTA of
// 256 threads walks its slice of the points.  Point records (prep_points) are staged through shared memory by 1-D bulk
// async copies (TMA unit, SASS UBLKCP) into a 2-stage ring guarded by mbarriers; every thread reads the same record
// at the same time, so all shared-memory reads are broadcasts.  The grid is (pair tiles) x (n splits), sized to whole
// waves of resident CTAs; per-split partial sums go to a workspace and are added in a fixed order (deterministic, no
// atomics).  The launchers work on point ranges so that gparml_statistics can follow a row-range upload (capi.cu).
//
// Bound: FP64 pipe.  Algorithmic count 6Q + 20 per (point, pair) with exp = 18.
#include <math.h>

#include "common.cuh"

#define GP_EXP_LOG2_TAB 6      // 64-entry table, degree-4 polynomial: faster here than the 256-entry / degree-3 default (gp_exp.cuh)
#include "gp_exp.cuh"

#ifndef PSI2_THREADS
#define PSI2_THREADS 256
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

template <int Q>
__global__ void __launch_bounds__(PSI2_THREADS, 2)
psi2_stats_kernel(const double *__restrict__ rec2, int64_t n, const double *__restrict__ Z, int64_t P,
                  const int2 *__restrict__ pair_idx, const double *__restrict__ pair_lk, int64_t n_per_split,
                  double *__restrict__ partial)
{
    constexpr int R = (3 * Q + 2) & ~1;
    constexpr int NT2 = (Q + 2) / 2;          // double2 loads covering v_0..v_{Q-1} and lc2
    constexpr int UNR = PSI2_UNROLL;
    constexpr int PP = 1;
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

// ---------------------------------------------------------------------------
// psi2x_stats (Q <= GP_PSI2X_MAX_Q): 5Q instead of 6Q FP64 instructions per (point, pair).
//
// With everything centred on c = column means of Z (mc = mu - c, zc = zbar - c; same envelope as embed_psi2x):
//   t_q = w (mc - zc)        = fma(-w, zc, w mc)                      (the old wd_q)
//   u_q = t_q^2 + v_q        = fma(t, t, v)                           (what TA accumulates)
//   exponent  -sum_q t_q^2 / w_q = sum_q u_q (-1/w_q) + sum_q alpha_q S_q   (v / w = alpha S)
// so the exponent is accumulated from the u the accumulation needs anyway: e = fma(u, -1/w, e), started at
// lk + lc2 + sum_q alpha_q S_q (prep_points).  Per (point, pair, q): t, u, e, TZ += psi t, TA += psi u.
// The records carry four numbers per (point, q) -- (-w, w mc) and (v, -1/w) -- read as two broadcast 16-byte
// loads that feed both pairs of the thread.  Rounding: the e sum cancels sum_q alpha_q S_q to ~2 eps sum_q alpha_q S_q
// absolute; when prep_points has seen alpha S > GP_PSI2X_ROBUST_AS the ROBUST instantiation runs instead
// (e = fma(t (-1/w), t, e) from lk + lc2: 6Q, no cancellation).  Both are launched; the device flag picks one.
//
// Issue order (operand reuse, see embed_x.cu): the three dependent steps of a latent dimension are skewed by one
// dimension each -- t(s), u(s-1), e(s-2) -- and each is issued for pair 0 then pair 1, which share the record operand
// in the same slot; the accumulations run pair-major so that psi stays in one operand slot.
// ---------------------------------------------------------------------------
#ifndef PSI2X_PF
#define PSI2X_PF 2        // latent dimensions whose record operands are requested ahead of their use
#endif
#ifndef PSI2X_TN
#define PSI2X_TN 128      // points per stage (B200, c3: 64 -> 21.47 ms, 128 -> 21.21 ms; three stages or 32 points are slower)
#endif

// One point for the two pairs of the thread.  Measured alternatives (B200, c3, ms per launch; this form 21.47 at 64
// points per stage): one exponent chain per pair instead of two 24.4; the first operands of the next point loaded
// during the accumulations 22.8; t / u / e skewed by two dimensions 22.2; 128-thread CTAs x 2 per SM 22.1.  ptxas
// reorders the FMAs of this function whatever the source order (also through volatile asm), so unlike embx_step (embed_x.cu) it is
// not written in issue order; what it keeps is the pairing (pair 0, pair 1) that shares the record operand.
template <int Q, bool ROBUST>
__device__ __forceinline__ void psi2x_point(const double *__restrict__ rp, const double (&lk)[2], const double (&zc)[2][Q],
                                            double (&acc)[2][1 + 2 * Q], const double *exp_tab)
{
    constexpr int PF = PSI2X_PF < Q ? PSI2X_PF : Q;
    const double2 *ra = reinterpret_cast<const double2 *>(rp);      // (-w, w mc)
    const double2 *rb = ra + Q;                                     // (v, -1/w)
    const double kn = rp[4 * Q + (ROBUST ? 1 : 0)];
    double t[2][Q], u[2][Q], e[2][2];
    e[0][0] = lk[0]; e[1][0] = lk[1]; e[0][1] = kn; e[1][1] = kn;
    double2 a[Q], b[Q];
#pragma unroll
    for (int q = 0; q < PF; ++q) { a[q] = ra[q]; b[q] = rb[q]; }
    // the three dependent steps of a latent dimension are skewed by one dimension each: t(s), u(s-1), e(s-2)
#pragma unroll
    for (int s = 0; s < Q + 2; ++s) {
        if (s + PF < Q) { a[s + PF] = ra[s + PF]; b[s + PF] = rb[s + PF]; }
        if (s < Q) {
            t[0][s] = fma(a[s].x, zc[0][s], a[s].y);
            t[1][s] = fma(a[s].x, zc[1][s], a[s].y);
        }
        if (s >= 1 && s - 1 < Q) {
            const int q = s - 1;
            if (ROBUST) {
                u[0][q] = t[0][q] * b[q].y;        // t (-1/w), only for the exponent
                u[1][q] = t[1][q] * b[q].y;
            } else {
                u[0][q] = fma(t[0][q], t[0][q], b[q].x);
                u[1][q] = fma(t[1][q], t[1][q], b[q].x);
            }
        }
        if (s >= 2) {
            const int q = s - 2;
            if (ROBUST) {
                e[0][q & 1] = fma(u[0][q], t[0][q], e[0][q & 1]);
                e[1][q & 1] = fma(u[1][q], t[1][q], e[1][q & 1]);
                u[0][q] = fma(t[0][q], t[0][q], b[q].x);
                u[1][q] = fma(t[1][q], t[1][q], b[q].x);
            } else {
                e[0][q & 1] = fma(u[0][q], b[q].y, e[0][q & 1]);
                e[1][q & 1] = fma(u[1][q], b[q].y, e[1][q & 1]);
            }
        }
    }
    // exp of both pairs, interleaved step by step
    double x[2], tt[2], r[2], pl[2], tab[2], psi[2];
    int k[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) x[v] = gp_exp_clamp(e[v][0] + e[v][1]);
#pragma unroll
    for (int v = 0; v < 2; ++v) tt[v] = fma(x[v], GP_EXP_SCALE, GP_EXP_SHIFT);
#pragma unroll
    for (int v = 0; v < 2; ++v) { k[v] = __double2loint(tt[v]); tt[v] = tt[v] - GP_EXP_SHIFT; }
#pragma unroll
    for (int v = 0; v < 2; ++v) { r[v] = fma(tt[v], GP_EXP_NEG_STEP, x[v]); tab[v] = exp_tab[k[v] & (GP_EXP_TAB - 1)]; }
#if GP_EXP_POLY_STEPS == 4
#pragma unroll
    for (int v = 0; v < 2; ++v) pl[v] = fma(r[v], 1.0 / 24.0, 1.0 / 6.0);
#pragma unroll
    for (int v = 0; v < 2; ++v) pl[v] = fma(pl[v], r[v], 0.5);
#else
#pragma unroll
    for (int v = 0; v < 2; ++v) pl[v] = fma(r[v], GP_EXP_C3, GP_EXP_C2);
#endif
#pragma unroll
    for (int v = 0; v < 2; ++v) pl[v] = fma(pl[v], r[v], GP_EXP_C1);
#pragma unroll
    for (int v = 0; v < 2; ++v) pl[v] = fma(pl[v], r[v], GP_EXP_C0);
#pragma unroll
    for (int v = 0; v < 2; ++v) pl[v] = tab[v] * pl[v];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        int m = k[v] >> GP_EXP_LOG2_TAB;
        m = m < -1021 ? -1021 : m;
        psi[v] = __hiloint2double(__double2hiint(pl[v]) + (m << 20), __double2loint(pl[v]));
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        acc[v][0] += psi[v];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            acc[v][1 + q] = fma(psi[v], t[v][q], acc[v][1 + q]);
            acc[v][1 + Q + q] = fma(psi[v], u[v][q], acc[v][1 + Q + q]);
        }
    }
}

template <int Q, bool ROBUST>
__global__ void __launch_bounds__(PSI2_THREADS, 1)
psi2x_stats_kernel(const double *__restrict__ recx, int64_t n, int64_t P, const double *__restrict__ pair_zc,
                   const double *__restrict__ pair_lk, int64_t n_per_split, double *__restrict__ partial,
                   const int *__restrict__ robust_flag)
{
    if ((*robust_flag != 0) != ROBUST) return;           // prep_points decided which instantiation this evaluation needs
    constexpr int RX = 4 * Q + 2;
    constexpr int QP = (Q + 1) & ~1;
    extern __shared__ __align__(16) double tile[];       // [STAGES][TN][RX]
    __shared__ __align__(8) uint64_t bar[PSI2_STAGES];
    __shared__ double exp_tab[GP_EXP_TAB];

    const int tid = threadIdx.x;
    int64_t p[2];
    double lk[2], zc[2][Q], acc[2][1 + 2 * Q];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        p[v] = ((int64_t)blockIdx.x * 2 + v) * blockDim.x + tid;       // 256 or 128 threads (gp_psi2x_threads)
        const int64_t pc = p[v] < P ? p[v] : P - 1;       // out-of-range threads work on a real pair and never store
        lk[v] = pair_lk[pc];
#pragma unroll
        for (int q = 0; q < Q; ++q) zc[v][q] = pair_zc[pc * QP + q];
#pragma unroll
        for (int j = 0; j < 1 + 2 * Q; ++j) acc[v][j] = 0.0;
    }

    const int64_t n_lo = (int64_t)blockIdx.y * n_per_split;
    const int64_t n_hi = (n_lo + n_per_split < n) ? (n_lo + n_per_split) : n;
    const int64_t span = n_hi > n_lo ? n_hi - n_lo : 0;
    const int ntiles = (int)((span + PSI2X_TN - 1) / PSI2X_TN);

    gp_exp_load_table(exp_tab);
    if (tid == 0) {
        for (int s = 0; s < PSI2_STAGES; ++s) gp_mbar_init(&bar[s], 1);
        gp_fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int s = t % PSI2_STAGES;
        const int64_t base = n_lo + (int64_t)t * PSI2X_TN;
        const int cnt = (int)((n_hi - base < PSI2X_TN) ? (n_hi - base) : PSI2X_TN);
        const uint32_t bytes = (uint32_t)cnt * RX * sizeof(double);
        gp_mbar_expect_tx(&bar[s], bytes);
        gp_bulk_g2s(tile + (size_t)s * PSI2X_TN * RX, recx + base * RX, bytes, &bar[s]);
    };
    if (tid == 0)
        for (int t = 0; t < PSI2_STAGES && t < ntiles; ++t) issue(t);

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % PSI2_STAGES;
        const int64_t base = n_lo + (int64_t)t * PSI2X_TN;
        const int cnt = (int)((n_hi - base < PSI2X_TN) ? (n_hi - base) : PSI2X_TN);
        gp_mbar_wait(&bar[s], (uint32_t)((t / PSI2_STAGES) & 1));
        const double *tb = tile + (size_t)s * PSI2X_TN * RX;
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) psi2x_point<Q, ROBUST>(tb + i * RX, lk, zc, acc, exp_tab);
        __syncthreads();      // every thread is done reading stage s
        if (tid == 0 && t + PSI2_STAGES < ntiles) issue(t + PSI2_STAGES);
    }

#pragma unroll
    for (int v = 0; v < 2; ++v) {
        if (p[v] < P) {
            double *out = partial + (size_t)blockIdx.y * (1 + 2 * Q) * P + p[v];
#pragma unroll
            for (int j = 0; j < 1 + 2 * Q; ++j) out[(size_t)j * P] = acc[v][j];
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

// CTA size of psi2x_stats: 128 threads where that wastes fewer pair slots in the last pair tile (c2: P = 1275 pairs fill
// 5 tiles of 256 pairs to 99.6 %, 3 tiles of 512 to 83 %)
static int gp_psi2x_threads(int64_t P)
{
    const int64_t w256 = (P + 511) / 512 * 512, w128 = (P + 255) / 256 * 256;
    return (w128 < w256) ? 128 : 256;
}

#define PSI2_CTA_SETUP_POINTS 24.0
// number of n-splits for `cnt` points: whole waves of resident CTAs, bounded workspace
template <int Q>
static int plan_q(gparml_ctx *c, int64_t cnt, int *splits_out)
{
    int occ = 1, pp, tn, threads = PSI2_THREADS;
    if constexpr (Q <= GP_PSI2X_MAX_Q) {
        const size_t smem = (size_t)PSI2_STAGES * PSI2X_TN * gp_recx_len(Q) * sizeof(double);
        GP_CUDA(cudaFuncSetAttribute(psi2x_stats_kernel<Q, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GP_CUDA(cudaFuncSetAttribute(psi2x_stats_kernel<Q, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        threads = gp_psi2x_threads(c->L.P);
        GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, psi2x_stats_kernel<Q, false>, threads, smem));
        pp = 2;
        tn = PSI2X_TN;
    } else {
        const size_t smem = (size_t)PSI2_STAGES * PSI2_TN * gp_rec_len(Q) * sizeof(double);
        GP_CUDA(cudaFuncSetAttribute(psi2_stats_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, psi2_stats_kernel<Q>, PSI2_THREADS, smem));
        pp = 1;
        tn = PSI2_TN;
    }
    if (occ < 1) occ = 1;
    const int64_t P = c->L.P;
    const int tiles = (int)((P + threads * pp - 1) / (threads * pp));
    const int64_t slots = (int64_t)c->sm_count * occ;
    const int64_t rows_x_P = (int64_t)(1 + 2 * Q) * P;
    int64_t max_splits = (cnt + tn - 1) / tn;             // at least one point tile per split
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
        // wave fill x the share of a CTA's life spent on points (its set-up -- pair registers, exp table, first
        // tile in flight -- costs about as much as PSI2_CTA_SETUP_POINTS points)
        const double per = (double)cnt / (double)s;
        const double eff = (double)total / (double)(waves * slots) * per / (per + PSI2_CTA_SETUP_POINTS);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    *splits_out = (int)best;
    return GPARML_OK;
}

// psi2_stats over the points [i0, i1): partial sums into the workspace slices [slice0, slice0 + splits)
template <int Q>
static int launch_range_q(gparml_ctx *c, int64_t i0, int64_t i1, int slice0, int splits)
{
    const int64_t P = c->L.P, cnt = i1 - i0;
    const int64_t rows_x_P = (int64_t)(1 + 2 * Q) * P;
    const int64_t n_per_split = (cnt + splits - 1) / splits;
    double *part = c->ws + (size_t)slice0 * rows_x_P;
    if constexpr (Q <= GP_PSI2X_MAX_Q) {
        constexpr int RX = 4 * Q + 2;
        const size_t smem = (size_t)PSI2_STAGES * PSI2X_TN * RX * sizeof(double);
        const int threads = gp_psi2x_threads(P);
        dim3 grid((unsigned)((P + threads * 2 - 1) / (threads * 2)), splits);
        // both instantiations are launched; the flag prep_points left in d_status[1] lets exactly one of them run
        psi2x_stats_kernel<Q, false><<<grid, threads, smem, c->stream>>>(c->rec2x + i0 * RX, cnt, P, c->pair_zc, c->pair_lk,
                                                                             n_per_split, part, c->d_status + 1);
        GP_LAUNCH_CHECK(c);
        psi2x_stats_kernel<Q, true><<<grid, threads, smem, c->stream>>>(c->rec2x + i0 * RX, cnt, P, c->pair_zc, c->pair_lk,
                                                                            n_per_split, part, c->d_status + 1);
        GP_LAUNCH_CHECK(c);
    } else {
        constexpr int R = (3 * Q + 2) & ~1;
        const size_t smem = (size_t)PSI2_STAGES * PSI2_TN * R * sizeof(double);
        dim3 grid((unsigned)((P + PSI2_THREADS - 1) / PSI2_THREADS), splits);
        psi2_stats_kernel<Q><<<grid, PSI2_THREADS, smem, c->stream>>>(c->rec2 + i0 * R, cnt, c->Z, P, c->pair_idx, c->pair_lk, n_per_split, part);
        GP_LAUNCH_CHECK(c);
    }
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
