// K5 (Psi2 part), hand-scheduled: embed_psi2x_kernel.
//
// Same arithmetic as the expanded-basis kernel described in embed.cu (4Q + 11 FP64 instructions
// per (point, pair)); this file exists because of how the FP64 pipe of B200 is fed:
//
//   * a warp-wide DFMA occupies the pipe for 2 cycles, but the register file delivers only two
//     fresh 64-bit operands in that time -- a DFMA whose three source operands are all fresh
//     register reads takes 3 cycles (tools/micro/dfma_regbw.cu on B200: 3 fresh 66 %, one operand
//     from the reuse cache 83 %, two 91 % of the DFMA peak).  The operand-reuse cache only serves
//     the instruction issued directly after the one that loaded it, in the same operand slot;
//   * ptxas keeps the FMA order of the source where latency does not force a change, and flags
//     operand reuse between neighbours, so the source below is written in the intended issue order
//     (the same arithmetic written q-outer / v-inner ran at 67 % of the pipe, this order at 74 %;
//     ptxas -O1 was tried and schedules worse).
//
// Order of one loop iteration (pair j accumulates while the exponent of pair j+1 is evaluated):
//   block E  (exponent of pair j+1): per latent dimension q the 2 NP FMAs are ordered
//            (z.x, v=0) (z.x, v=1) (z.y, v=1) (z.y, v=0): every second FMA takes z from the
//            reuse cache; 2 NP dependent chains, the same chain recurs every 2 NP instructions;
//   block XA (exp of pair j+1, accumulation of pair j): the ten dependent steps of the two exp
//            evaluations are each followed by the four accumulations of one latent dimension in
//            "snake" order (h0, z.x) (h1, z.x) (h1, z.y) (h0, z.y): consecutive FMAs share h or z,
//            so each reads two fresh operands; the independent accumulations cover the latency of
//            the exp chain.
// Shared-memory reads of the pair records are issued a few groups ahead of their use.
//
// Replaces partial_terms.py:367-431 (the Psi2 terms of grad_X_mu / grad_X_S), see embed.cu.
#include <math.h>

#include "embed.cuh"
#include "gp_exp.cuh"

#ifndef EMBX_CP
#define EMBX_CP 64             // pairs per stage
#endif
#ifndef EMBX_STAGES
#define EMBX_STAGES 2
#endif
#ifndef EMBX_PF
#define EMBX_PF 3              // shared-memory prefetch distance (latent dimensions)
#endif

template <int Q> struct EmbXCfg {
    static constexpr int NP = (Q <= 10) ? 2 : 1;
    static constexpr int MINB = (Q <= 4) ? 4 : 2;
};

// exp(x) split into its dependent steps (gp_exp.cuh: same constants, same result)
struct ExpState {
    double x, t, r, p, tab;
    int k;
};
#define GPX_SHIFT GP_EXP_SHIFT

#ifndef EMBX_HORNER_DEPTH
#define EMBX_HORNER_DEPTH 1    // software-pipeline depth of the Horner exponent (t leads e by this many latent dimensions)
#endif
#ifndef EMBX_HORNER
#define EMBX_HORNER 1          // exponent form: 0 = (zc, zc^2) dot product, 1 = Horner on zc, 2 = Horner + register prefetch
#endif

// zn: (zc, zc^2) record of the next pair; cn: its zc-only record; cnn: the zc-only record of the pair after it
// (EMBX_HORNER 2 loads it into zr during block XA); zc: (zc, zc^2) record of the current pair
template <int Q, int NP, bool DO_E, bool DO_A>
__device__ __forceinline__ void embx_step(const double2 *__restrict__ zn, const double2 *__restrict__ cn,
                                          const double2 *__restrict__ cnn, double2 (&zr)[(Q + 1) / 2], const double2 gn,
                                          const double2 *__restrict__ zc,
                                          const double (&hc)[NP], double (&hn)[NP], const double (&kn)[NP],
                                          const double (&A)[NP][Q], const double (&nW)[NP][Q], double (&bz)[NP][Q],
                                          double (&bzz)[NP][Q], double (&ah)[NP], const double *exp_tab)
{
    constexpr int PF = EMBX_PF;
    ExpState es[NP];
    constexpr int QP2 = (Q + 1) / 2;
#if EMBX_HORNER == 0
    if (DO_E) {
        // ---- block E: exponent of the next pair -------------------------------------------------
        constexpr int NC = (NP == 1) ? 2 : 1;      // sub-chains per sum: always >= 4 independent chains
        double e0[NP][NC], e1[NP][NC];
        double2 z[Q];
#pragma unroll
        for (int q = 0; q < PF && q < Q; ++q) z[q] = zn[q];
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            e0[v][0] = gn.x;
            e1[v][0] = kn[v];
            if (NC == 2) { e0[v][NC - 1] = 0.0; e1[v][NC - 1] = 0.0; }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if (q + PF < Q) z[q + PF] = zn[q + PF];
#pragma unroll
            for (int v = 0; v < NP; ++v) e0[v][q % NC] = fma(z[q].x, A[v][q], e0[v][q % NC]);
#pragma unroll
            for (int v = NP - 1; v >= 0; --v) e1[v][q % NC] = fma(z[q].y, nW[v][q], e1[v][q % NC]);
        }
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            if (NC == 2) es[v].x = gp_exp_clamp((e0[v][0] + e0[v][NC - 1]) + (e1[v][0] + e1[v][NC - 1]));
            else es[v].x = gp_exp_clamp(e0[v][0] + e1[v][0]);
        }
    }
#else
    if (DO_E) {
        // ---- block E: exponent of the next pair, in Horner form ---------------------------------
        //   kn + sum_q zc_q (A_q + nW_q zc_q):  t = fma(nW, zc, A), e = fma(t, zc, e)
        // -- still 2 FMAs per (point, q) but only zc is read (half the shared-memory bytes of the (zc, zc^2)
        // form: Q/2 16-byte loads from the zc-only ring), software-pipelined by one latent dimension so that e(q)
        // issues four instructions after the t(q) it depends on.
#if EMBX_HORNER == 1
#pragma unroll
        for (int k = 0; k < QP2; ++k) zr[k] = cn[k];
#endif
        double e[NP][2];
#if EMBX_HORNER_DEPTH == 2
        // t two latent dimensions ahead of e: 8 instructions between a t and the e that consumes it
        double tq[3][NP];
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            e[v][0] = gn.x;
            e[v][1] = kn[v];
            tq[0][v] = fma(nW[v][0], zr[0].x, A[v][0]);
            if (Q > 1) tq[1][v] = fma(nW[v][Q > 1 ? 1 : 0], zr[0].y, A[v][Q > 1 ? 1 : 0]);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double zq = (q & 1) ? zr[q >> 1].y : zr[q >> 1].x;
            if (q + 2 < Q) {
                const double zq2 = ((q + 2) & 1) ? zr[(q + 2) >> 1].y : zr[(q + 2) >> 1].x;
#pragma unroll
                for (int v = 0; v < NP; ++v) tq[(q + 2) % 3][v] = fma(nW[v][(q + 2) < Q ? (q + 2) : 0], zq2, A[v][(q + 2) < Q ? (q + 2) : 0]);
            }
#pragma unroll
            for (int v = NP - 1; v >= 0; --v) e[v][q & 1] = fma(tq[q % 3][v], zq, e[v][q & 1]);
        }
#else
        double t[NP], tn[NP];
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            e[v][0] = gn.x;
            e[v][1] = kn[v];
            t[v] = fma(nW[v][0], zr[0].x, A[v][0]);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double zq = (q & 1) ? zr[q >> 1].y : zr[q >> 1].x;
            if (q + 1 < Q) {
                const double zq1 = ((q + 1) & 1) ? zr[(q + 1) >> 1].y : zr[(q + 1) >> 1].x;
#pragma unroll
                for (int v = 0; v < NP; ++v) tn[v] = fma(nW[v][(q + 1) < Q ? (q + 1) : 0], zq1, A[v][(q + 1) < Q ? (q + 1) : 0]);
            }
#pragma unroll
            for (int v = NP - 1; v >= 0; --v) e[v][q & 1] = fma(t[v], zq, e[v][q & 1]);
#pragma unroll
            for (int v = 0; v < NP; ++v) t[v] = tn[v];
        }
#endif
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].x = gp_exp_clamp(e[v][0] + e[v][1]);
    }
#endif
    // ---- block XA: exp steps of the next pair, each followed by one group of accumulations ---------
    double2 zz[Q];
    if (DO_A) {
#pragma unroll
        for (int q = 0; q < PF && q < Q; ++q) zz[q] = zc[q];
#pragma unroll
        for (int v = 0; v < NP; ++v) ah[v] += hc[v];
    }
    const int sg = DO_E ? (__double2hiint(gn.y) & 0x80000000) : 0;
#if EMBX_HORNER == 2
#define EMBX_PREF(q) if (DO_E && (q) < QP2) zr[(q) < QP2 ? (q) : 0] = cnn[(q) < QP2 ? (q) : 0];
#else
#define EMBX_PREF(q)
#endif
#define EMBX_GROUP(q)                                                                        \
    EMBX_PREF(q)                                                                             \
    if (DO_A && (q) < Q) {                                                                   \
        if ((q) + PF < Q) zz[((q) + PF) < Q ? ((q) + PF) : 0] = zc[((q) + PF) < Q ? ((q) + PF) : 0]; \
        _Pragma("unroll") for (int v = 0; v < NP; ++v) bz[v][(q) < Q ? (q) : 0] = fma(hc[v], zz[(q) < Q ? (q) : 0].x, bz[v][(q) < Q ? (q) : 0]); \
        _Pragma("unroll") for (int v = NP - 1; v >= 0; --v) bzz[v][(q) < Q ? (q) : 0] = fma(hc[v], zz[(q) < Q ? (q) : 0].y, bzz[v][(q) < Q ? (q) : 0]); \
    }
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].t = fma(es[v].x, GP_EXP_SCALE, GPX_SHIFT);
    }
    EMBX_GROUP(0)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            es[v].k = __double2loint(es[v].t);
            es[v].t = es[v].t - GPX_SHIFT;
        }
    }
    EMBX_GROUP(1)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            es[v].r = fma(es[v].t, GP_EXP_NEG_STEP, es[v].x);
            es[v].tab = exp_tab[es[v].k & (GP_EXP_TAB - 1)];
        }
    }
    EMBX_GROUP(2)
#if GP_EXP_POLY_STEPS == 4
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].r, 1.0 / 24.0, 1.0 / 6.0);
    }
    EMBX_GROUP(3)
    EMBX_GROUP(4)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].p, es[v].r, 0.5);
    }
#else
    EMBX_GROUP(3)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].r, GP_EXP_C3, GP_EXP_C2);
    }
    EMBX_GROUP(4)
#endif
    EMBX_GROUP(5)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].p, es[v].r, GP_EXP_C1);
    }
    EMBX_GROUP(6)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = fma(es[v].p, es[v].r, GP_EXP_C0);
    }
    EMBX_GROUP(7)
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) es[v].p = es[v].tab * es[v].p;      // in [1, 2.03)
    }
    EMBX_GROUP(8)
    EMBX_GROUP(9)
    EMBX_GROUP(10)
    EMBX_GROUP(11)
    EMBX_GROUP(12)
    EMBX_GROUP(13)
    EMBX_GROUP(14)
    EMBX_GROUP(15)
#undef EMBX_GROUP
#undef EMBX_PREF
    if (DO_E) {
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            int m = es[v].k >> GP_EXP_LOG2_TAB;
            m = m < -1021 ? -1021 : m;
            hn[v] = __hiloint2double((__double2hiint(es[v].p) + (m << 20)) ^ sg, __double2loint(es[v].p));
        }
    }
}

template <int Q>
__global__ void __launch_bounds__(EMB_THREADS, EmbXCfg<Q>::MINB)
embed_psi2x_kernel(EmbedParams p)
{
    constexpr int R = (3 * Q + 2) & ~1;
    constexpr int NP = EmbXCfg<Q>::NP;
    constexpr int CP = (EMBX_HORNER && Q > 12) ? EMBX_CP / 2 : EMBX_CP, ST = EMBX_STAGES;      // static shared memory <= 48 kB
    constexpr int QP = (Q + 1) & ~1;
    __shared__ __align__(16) double2 ring_z[ST][CP * Q];          // (zc, zc^2): the accumulations
#if EMBX_HORNER
    __shared__ __align__(16) double ring_c[ST][CP * QP];          // zc alone: the exponent
#endif
    __shared__ __align__(16) double2 ring_g[ST][CP];
    __shared__ __align__(8) uint64_t bar[ST];
    __shared__ double exp_tab[GP_EXP_TAB];
    const int tid = threadIdx.x;

    const int64_t p_lo = p.p_bounds[blockIdx.y], p_hi = p.p_bounds[blockIdx.y + 1];
    const int nchunks = (int)((p_hi - p_lo + CP - 1) / CP);
    gp_exp_load_table(exp_tab);
    if (tid == 0) {
        for (int s = 0; s < ST; ++s) gp_mbar_init(&bar[s], 1);
        gp_fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int s = t % ST;
        const int64_t base = p_lo + (int64_t)t * CP;
        const int cnt = (int)((p_hi - base < CP) ? (p_hi - base) : CP);
        const uint32_t bz_ = (uint32_t)cnt * Q * sizeof(double2), bg = (uint32_t)cnt * sizeof(double2);
#if EMBX_HORNER
        const uint32_t bc = (uint32_t)cnt * QP * sizeof(double);
        gp_mbar_expect_tx(&bar[s], bz_ + bg + bc);
        gp_bulk_g2s(&ring_c[s][0], p.pair_zc + base * QP, bc, &bar[s]);
#else
        gp_mbar_expect_tx(&bar[s], bz_ + bg);
#endif
        gp_bulk_g2s(&ring_z[s][0], p.pair_zz + base * Q, bz_, &bar[s]);
        gp_bulk_g2s(&ring_g[s][0], p.pair_h + base, bg, &bar[s]);
    };
    if (tid == 0)
        for (int t = 0; t < ST && t < nchunks; ++t) issue(t);

    int64_t i[NP];
    bool valid[NP];
    double kn[NP], A[NP][Q], nW[NP][Q], bz[NP][Q], bzz[NP][Q], ah[NP];
#pragma unroll
    for (int v = 0; v < NP; ++v) {
        i[v] = p.i0 + ((int64_t)blockIdx.x * NP + v) * EMB_THREADS + tid;
        valid[v] = i[v] < p.i1;
        if (!valid[v]) i[v] = p.i1 - 1;                  // compute on a real record, never store
        const double2 *r2 = reinterpret_cast<const double2 *>(p.rec2 + i[v] * R);
        kn[v] = p.rec2[i[v] * R + 3 * Q];
        ah[v] = 0.0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double2 mw = r2[q];
            const double mc = mw.x - p.glob->center[q];
            const double wm = mw.y * mc;
            kn[v] = fma(-wm, mc, kn[v]);
            A[v][q] = 2.0 * wm;
            nW[v][q] = -mw.y;
            bz[v][q] = 0.0;
            bzz[v][q] = 0.0;
        }
    }

    for (int t = 0; t < nchunks; ++t) {
        const int s = t % ST;
        const int64_t base = p_lo + (int64_t)t * CP;
        const int cnt = (int)((p_hi - base < CP) ? (p_hi - base) : CP);
        gp_mbar_wait(&bar[s], (uint32_t)((t / ST) & 1));
        const double2 *zt = &ring_z[s][0];
#if EMBX_HORNER
        const double2 *ct = reinterpret_cast<const double2 *>(&ring_c[s][0]);      // QP / 2 pairs of zc per pair record
#else
        const double2 *ct = &ring_z[s][0];                                         // unused
#endif
        const double2 *gt = &ring_g[s][0];
        double hc[NP], hn[NP];
        double2 zr[(Q + 1) / 2];
        constexpr int CS = QP / 2;                 // double2 per zc-only record
#if EMBX_HORNER == 2
#pragma unroll
        for (int k = 0; k < (Q + 1) / 2; ++k) zr[k] = ct[k];                                        // zc of pair 0
#endif
        // h of pair 0 (prefetches the zc of pair 1)
        embx_step<Q, NP, true, false>(zt, ct, ct + (cnt > 1 ? 1 : 0) * CS, zr, gt[0], zt, hc, hc, kn, A, nW, bz, bzz, ah, exp_tab);
#pragma unroll 1
        for (int j = 0; j + 1 < cnt; ++j) {
            const int j2 = j + 2 < cnt ? j + 2 : cnt - 1;
            embx_step<Q, NP, true, true>(zt + (j + 1) * Q, ct + (j + 1) * CS, ct + j2 * CS, zr, gt[j + 1], zt + j * Q, hc, hn, kn, A,
                                         nW, bz, bzz, ah, exp_tab);
#pragma unroll
            for (int v = 0; v < NP; ++v) hc[v] = hn[v];
        }
        embx_step<Q, NP, false, true>(zt, ct, ct, zr, gt[0], zt + (cnt - 1) * Q, hc, hn, kn, A, nW, bz, bzz, ah, exp_tab);   // last pair
        __syncthreads();      // every thread is done reading stage s
        if (tid == 0 && t + ST < nchunks) issue(t + ST);
    }
    if (p.fuse_finish) {
        // one pair split: this thread holds the complete sums of its points -- finish here instead of a round trip
        // of (2Q + 1) doubles per point through HBM and a second kernel (embed.cu embed_finish_kernel, same arithmetic)
#pragma unroll
        for (int v = 0; v < NP; ++v) {
            if (valid[v]) {
                const double2 *r2 = reinterpret_cast<const double2 *>(p.rec2 + i[v] * R);
                const double *p1 = p.psi1_part + (size_t)i[v] * (2 * Q + 1);
                const int64_t o = i[v] * Q;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const double2 mw = r2[q];
                    gp_embed_finish_one(mw.x, mw.y, mw.x - p.glob->center[q], bz[v][q], bzz[v][q], ah[v], p1[q], p1[Q + q],
                                        p.s_pos[o + q], p.s_sig[o + q], p.gx_mu + o + q, p.gx_s + o + q, p.grad_latest + o + q,
                                        p.grad_latest + p.n * Q + o + q);
                }
            }
        }
        return;
    }
#pragma unroll
    for (int v = 0; v < NP; ++v) {
        if (valid[v]) {
            double *out = p.partial + ((size_t)blockIdx.y * p.pstride + (i[v] - p.pbase)) * (2 * Q + 1);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                out[q] = bz[v][q];
                out[Q + q] = bzz[v][q];
            }
            out[2 * Q] = ah[v];
        }
    }
}

int gp_embed_psi2x_points_per_cta(int Q) { return EMB_THREADS * ((Q <= 10) ? 2 : 1); }

template <int Q> static int occ_q(int *occ)
{
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, embed_psi2x_kernel<Q>, EMB_THREADS, 0));
    return GPARML_OK;
}
template <int Q> static int launch_x(gparml_ctx *c, const EmbedParams &p, int ntiles, int splits)
{
    dim3 grid((unsigned)ntiles, splits);
    embed_psi2x_kernel<Q><<<grid, EMB_THREADS, 0, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

#define ALL_Q(F) F(1) F(2) F(3) F(4) F(5) F(6) F(7) F(8) F(9) F(10) F(11) F(12) F(13) F(14) F(15) F(16)

int gp_embed_psi2x_occupancy(int Q, int *occ)
{
    switch (Q) {
#define CASE_Q(q) case q: return occ_q<q>(occ);
        ALL_Q(CASE_Q)
#undef CASE_Q
    }
    gp_set_error("embed_grads: unsupported Q=%d (1..%d)", Q, GP_MAX_Q);
    return GPARML_ERR_ARG;
}

int gp_launch_embed_psi2x(gparml_ctx *c, const EmbedParams &p, int ntiles, int splits)
{
    switch (c->Q) {
#define CASE_Q(q) case q: return launch_x<q>(c, p, ntiles, splits);
        ALL_Q(CASE_Q)
#undef CASE_Q
    }
    gp_set_error("embed_grads: unsupported Q=%d (1..%d)", c->Q, GP_MAX_Q);
    return GPARML_ERR_ARG;
}
