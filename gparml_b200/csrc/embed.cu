// K5: embed_grads -- per-point gradients of the bound w.r.t. the variational means and
// variances (the reference's second map, "embeddings_mapper").
//
// Replaces (citations relative to /root/reference)
//   partial_terms.py:367-398   grad_X_mu
//   partial_terms.py:400-431   grad_X_S
//   local_MapReduce.py:357-360 grad_latest = -[g_mu, g_S * sigmoid(S_raw)]
// and recomputes Psi2_n on the fly instead of re-building the (n, M, M) tensor a second time
// (local_MapReduce.py:348 -> partial_terms.py:45-48).
//
// With G1 = dF/dPsi1Y (M,D), Gs[p] = G2[m,m'] + G2[m',m] (m<m'), G2[m,m] (m=m'),
// B[n,m] = sum_d Y[n,d] G1[m,d], wd_q = w_nq (mu_nq - zbar_q), ad_q = a_nq (mu_nq - z_mq):
//   g_mu[n,q] = -mu_nq - sum_m B Psi1 ad_q - 2 sum_p Gs Psi2_n wd_q
//   g_S[n,q]  = -1/2 (1 - 1/S_nq) + 1/2 sum_m B Psi1 (ad_q^2 - a_nq) + sum_p Gs Psi2_n (2 wd_q^2 - w_nq)
//
// Psi2 part (>95 % of the work), all in the expanded basis described below: embed_psi2m_kernel in embed_m.cu (5 <= Q <= 10:
// exponent and sums as two products on the FP64 tensor-core instruction), embed_psi2x_kernel in embed_x.cu (other Q:
// DFMA), embed_psi2_f32_kernel in psi2_f32.cu for the opt-in fp32 path (sqrt(w) basis: u_q = sqrt(w_q) (mu_q - zbar_q),
// partial sums AM_q = sum_p h u_q, AS_q = sum_p h u_q^2, AH = sum_p h; embed_finish basis 0).  The reduction
// runs over pairs for each point (the opposite direction to psi2_stats), so a thread / a warp row owns points.
// Launch (launch_q): the point tiles that make whole rounds per SM take the full pair range and finish the gradients in
// their epilogue; the remaining tiles are split over the pair range so that they fill the machine, their partials are
// combined in a fixed order by embed_finish.
//
// Psi1 part (embed_psi1_kernel): thread per point, loop over the M inducing points.
//
// Bound: FP64 pipe.
#include <math.h>

#include "common.cuh"
#include "gp_exp.cuh"

#include "embed.cuh"

#ifdef GP_USE_LIBM_EXP
#define EMB_EXP(x) exp(x)
#else
#define EMB_EXP(x) gp_exp((x), exp_tab)
#endif

// ---------------------------------------------------------------------------------------------
// Expanded-basis Psi2 part (the fp64 default; kernels in embed_m.cu and embed_x.cu).  With mc = mu - center,
// zc = zbar - center:
//   -sum_q w_q (mc_q - zc_q)^2 = -sum_q w_q mc_q^2 + sum_q (2 w_q mc_q) zc_q - sum_q w_q zc_q^2
// so the exponent of a (point, pair) is a dot product of the per-point vector (A_q = 2 w mc, W_q = w;
// registers) with the per-pair vector (zc_q, zc_q^2; pair_zz table): 2Q FMAs.  The accumulators
// are kept in the pair basis too,
//   AH = sum_p h,  BZ_q = sum_p h zc_q,  BZZ_q = sum_p h zc_q^2          (h = Gs Psi2_n)
//   sum_p h wd_q   = w_q (mc_q AH - BZ_q)
//   sum_p h wd_q^2 = w_q^2 (mc_q^2 AH - 2 mc_q BZ_q + BZZ_q)            (embed_finish, basis 1)
// another 2Q FMAs; with Gs folded into the exponent (pair_h: lk + log|Gs| and the sign, applied by an
// integer XOR) 4Q + 9 FP64 instructions per (point, pair) in embed_psi2x (embed_psi2m: both dot products as MMAs, the
// pair factor folded into the second one's matrix).  Centring on the
// column means of Z keeps the cancellation in these differences at the scale of the spread of Z,
// not of its offset from the origin.  The pair table (P x 2Q doubles, 808 kB at M = 100, Q = 10) does
// not fit shared memory: every CTA walks its pair range in order and stages it through a ring of
// 1-D bulk async copies (TMA unit, SASS UBLKCP) guarded by mbarriers, together with the matching
// pair_h entries; all reads are shared-memory broadcasts feeding the NP points of a thread.
// Grid = (point tiles) x (pair-range splits), split partials combined in a fixed order.
// ---------------------------------------------------------------------------------------------

// Psi1 side (partial_terms.py:388-390, 421-423): h1 = B[n,m] Psi1[n,m], B = Y G1^T;
//   out[q] = sum_m h1 ad_q,  out[Q+q] = sum_m h1 (ad_q^2 - a_q),  out[2Q] = sum_m h1
// DR > 0: D <= DR, the point's Y row lives in registers and G1 (zero-padded to DR columns) in shared
// memory next to Z (B200, c3: 1.39 -> see DESIGN.md); DR = 0: any D, Y and G1 read through L1.
template <int Q, int DR>
__global__ void __launch_bounds__(128) embed_psi1_kernel(EmbedParams p)
{
    constexpr int R = (3 * Q + 2) & ~1;
    constexpr int DRA = DR > 0 ? DR : 1;
    extern __shared__ __align__(16) double zs[];         // [M][Q] = Z, then [M][DR] = G1
    __shared__ double exp_tab[GP_EXP_TAB];
    const int tid = threadIdx.x;
    const int M = p.M;
    double *g1s = zs + (((size_t)M * Q + 1) & ~(size_t)1);
    for (int idx = tid; idx < M * Q; idx += 128) zs[idx] = p.Z[idx];
    if (DR > 0)
        for (int idx = tid; idx < M * DR; idx += 128) {
            const int m = idx / DR, d = idx % DR;
            g1s[idx] = d < p.D ? p.G1[(size_t)m * p.D + d] : 0.0;
        }
    gp_exp_load_table(exp_tab);
    __syncthreads();
    const int64_t i = p.i0 + (int64_t)blockIdx.x * 128 + tid;
    if (i >= p.i1) return;
    const double2 *r1 = reinterpret_cast<const double2 *>(p.rec1 + i * R);
    const double lc1 = p.rec1[i * R + 3 * Q];
    const double *y = p.Y + i * p.D;
    double mu[Q], a[Q], ad[Q], s1[Q], s2[Q], yr[DRA];
    double s0 = 0.0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const double2 ma = r1[q];
        mu[q] = ma.x;
        a[q] = ma.y;
        s1[q] = 0.0;
        s2[q] = 0.0;
    }
#pragma unroll
    for (int d = 0; d < DR; ++d) yr[d] = d < p.D ? y[d] : 0.0;
    for (int m = 0; m < M; ++m) {
        const double *z = zs + m * Q;
        double e = lc1;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double d = mu[q] - z[q];
            ad[q] = a[q] * d;
            e = fma(-0.5 * ad[q], d, e);
        }
        double b = 0.0, b2 = 0.0;
        if (DR > 0) {
            const double *g1 = g1s + m * DR;
#pragma unroll
            for (int d = 0; d < DR; ++d) {
                if (d & 1) b2 = fma(yr[d], g1[d], b2);
                else b = fma(yr[d], g1[d], b);
            }
            b += b2;
        } else {
            const double *g1 = p.G1 + (size_t)m * p.D;
            for (int d = 0; d < p.D; ++d) b = fma(y[d], g1[d], b);
        }
        const double h1 = b * EMB_EXP(e);
        s0 += h1;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double t = h1 * ad[q];
            s1[q] += t;
            s2[q] = fma(t, ad[q], s2[q]);
        }
    }
    double *out = p.psi1_part + (size_t)i * (2 * Q + 1);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        out[q] = s1[q];
        out[Q + q] = fma(-a[q], s0, s2[q]);      // sum_m h1 (ad_q^2 - a_q): the finish needs no Psi1 record
    }
    out[2 * Q] = s0;
}

template <int Q, int DR>
static int launch_psi1_part(gparml_ctx *c, const EmbedParams &p, int64_t cnt)
{
    const size_t smem = ((((size_t)c->M * Q + 1) & ~(size_t)1) + (size_t)c->M * DR) * sizeof(double);
    GP_CUDA(cudaFuncSetAttribute(embed_psi1_kernel<Q, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_psi1_kernel<Q, DR><<<(unsigned)((cnt + 127) / 128), 128, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// Combine the split partials and the Psi1 part, add the KL terms (partial_terms.py:385,418),
// apply the softplus chain and the sign flip (local_MapReduce.py:357-360).  Runs when the pair range was split over
// several CTAs (or on the fp32 path); with one split the epilogue of embed_psi2x does this itself.
__global__ void __launch_bounds__(256) embed_finish_kernel(const double *__restrict__ partial, int splits, int64_t pstride, int64_t pbase,
                                                           const double *__restrict__ psi1_part, int64_t n, int64_t i0, int64_t cnt, int Q, int R,
                                                           const double *__restrict__ rec2,
                                                           const double *__restrict__ s_pos, const double *__restrict__ s_sig,
                                                           double *__restrict__ gx_mu, double *__restrict__ gx_s,
                                                           double *__restrict__ grad_latest, int basis,
                                                           const GlobalsDev *__restrict__ glob)
{
    const int64_t loc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (loc >= cnt * Q) return;
    const int64_t idx = i0 * Q + loc;
    const int64_t i = idx / Q;
    const int q = (int)(idx % Q);
    const int W = 2 * Q + 1;
    double am = 0.0, as = 0.0, ah = 0.0;
    for (int s = 0; s < splits; ++s) {
        const double *pr = partial + ((size_t)s * pstride + (i - pbase)) * W;
        am += pr[q];
        as += pr[Q + q];
        ah += pr[2 * Q];
    }
    const double *p1 = psi1_part + (size_t)i * W;
    const double2 mw = *reinterpret_cast<const double2 *>(rec2 + i * R + 2 * q);
    const double mu = mw.x, w = mw.y;
    double mc = mu - glob->center[q];
    if (basis == 0) {  // sqrt(w) basis (fp32 kernel): am = sum_p h u_q, as = sum_p h u_q^2 with u = sqrt(w) (mu - zbar);
        // in the shared formula t1 = w (mc ah - am'), t2 = w^2 (mc^2 ah - 2 mc am' + as') this is mc = 0, am' = -am / sqrt(w), as' = as / w
        mc = 0.0;
        const double rw = w > 0.0 ? 1.0 / sqrt(w) : 0.0;
        am = -am * rw;
        as = as * rw * rw;
    }
    gp_embed_finish_one(mu, w, mc, am, as, ah, p1[q], p1[Q + q], s_pos[idx], s_sig[idx], gx_mu + idx, gx_s + idx, grad_latest + idx,
                        grad_latest + n * Q + idx);
}

int gp_launch_embed_psi2_f32(gparml_ctx *c, const int *m_bounds, int splits, double *partial, int64_t i0, int64_t i1);

// fewest splits of the reduction range that fill whole waves of resident CTAs
static int pick_splits(int64_t ntiles, int64_t slots, int max_splits)
{
    int best = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= max_splits; ++s) {
        const int64_t total = ntiles * s;
        const int64_t waves = (total + slots - 1) / slots;
        const double eff = (double)total / (double)(waves * slots);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }   // prefer fewer splits unless clearly better
        if (waves >= 8) break;
    }
    return best;
}

template <int Q>
static int launch_q(gparml_ctx *c, int64_t i0, int64_t i1)
{
    const int64_t cnt = i1 - i0;
    const bool fp32 = (c->flags & GPARML_FLAG_FP32_MAP) != 0;
    int occ = 4, np = 1;                                  // fp32 kernel: one point per thread, 4 CTAs per SM
#ifdef EMB_NO_MMA
    const bool use_m = false;
#else
    // the tensor-core formulation (embed_m.cu) where its padding is small: 2Q + 1 features in tiles of 8
    const bool use_m = !fp32 && c->pair_r != nullptr && Q >= GP_PSI2M_MIN_Q && Q <= GP_PSI2M_MAX_Q;
#endif
    if (use_m) {
        GP_TRY(gp_embed_psi2m_occupancy(Q, &occ));
    } else if (!fp32) {
        np = gp_embed_psi2x_points_per_cta(Q) / EMB_THREADS;
        GP_TRY(gp_embed_psi2x_occupancy(Q, &occ));
    }
    if (occ < 1) occ = 1;
    const int64_t per_cta = use_m ? gp_embed_psi2m_points_per_cta() : (int64_t)EMB_THREADS * np;
    const int64_t pchunks = (c->L.P + GP_PAIR_CHUNK - 1) / GP_PAIR_CHUNK;      // pair range of embed_psi2m in chunks
    const int64_t ntiles = (cnt + per_cta - 1) / per_cta;
    const int64_t slots = (int64_t)c->sm_count * occ;
    const int64_t Pn = c->L.P;
    int max_splits = c->M < EMB_MAX_SPLITS ? c->M : EMB_MAX_SPLITS;
    if (!fp32 && Pn / 64 < max_splits) max_splits = (int)(Pn / 64 > 0 ? Pn / 64 : 1);   // >= 64 pairs per split
    if (use_m && max_splits > pchunks / 2) max_splits = (int)(pchunks / 2 > 0 ? pchunks / 2 : 1);   // >= 2 chunks per split
    // Two launches (fp64 path): the point tiles that make whole rounds -- every SM gets the same number -- run with the
    // whole pair range and finish the gradients in their epilogue; the remaining tiles (fewer than one per SM) are split
    // over the pair range so that they, too, fill the machine, and go through embed_finish.  One launch with a partial
    // last round costs up to one round of an SM's time (2 GPUs at c3: 13.2 tiles per SM took the time of 14).
    // (a number of tiles per SM that is no multiple of the resident CTAs leaves every SM with a partly filled last
    // round, which costs more the fewer rounds there are: below 4 rounds the whole part is cut to whole rounds)
    int64_t per_sm = fp32 ? 0 : ntiles / c->sm_count;
    if (per_sm < 4 * occ) per_sm -= per_sm % occ;
    // ... and the last whole round joins the split part: the tail of the master step (one CTA, ~0.2 ms, side stream)
    // takes an SM away from this kernel for a while, and only small, dynamically scheduled CTAs at the end absorb that
    // (8 GPUs at c3 with exactly 6 tiles per SM: 2.48 ms instead of 2.30)
#ifndef EMB_EXACT_FIT
    per_sm = per_sm >= 2 * occ ? per_sm - occ : 0;
#endif
    int64_t full_tiles = per_sm * c->sm_count;
#ifdef EMB_NO_FUSE
    full_tiles = 0;
#endif
    const int64_t t0 = i0 + full_tiles * per_cta < i1 ? i0 + full_tiles * per_cta : i1;      // first point of the split part
    const int64_t tail_cnt = i1 - t0, tail_tiles = (tail_cnt + per_cta - 1) / per_cta;
    const int splits = tail_tiles > 0 ? pick_splits(tail_tiles, slots, max_splits) : 1;
    EmbedParams p;
    p.rec1 = c->rec1; p.rec2 = c->rec2; p.Y = c->Y; p.Z = c->Z; p.G1 = c->g_1; p.pair_g = c->pair_g;
    p.pair_zz = c->pair_zz; p.pair_zc = c->pair_zc; p.pair_h = c->pair_h; p.pair_r = c->pair_r; p.pair_ra = c->pair_ra; p.glob = c->d_glob;
    p.n = c->n; p.i0 = i0; p.i1 = i1; p.M = c->M; p.D = c->D;
    // row splits (fp32 kernel): every split owns about P / splits pairs (row m has M - m pairs)
    const double P = (double)c->L.P;
    p.m_bounds[0] = 0;
    int m = 0;
    for (int s = 1; s < splits; ++s) {
        const double target = P * s / splits;
        while (m < c->M && (double)gp_pair_index(c->M, m, m) < target) ++m;
        if (m <= p.m_bounds[s - 1]) m = p.m_bounds[s - 1] + 1;
        if (m > c->M) m = c->M;
        p.m_bounds[s] = m;
    }
    p.m_bounds[splits] = c->M;
    const size_t W = 2 * Q + 1;
    // workspace: [n][W] Psi1 part, then [splits][pstride][W] partials of the split part (fp32 path: stride n, absolute rows)
    const int64_t pstride = fp32 ? c->n : (tail_cnt > 0 ? tail_cnt : 1), pbase = fp32 ? 0 : t0;
    GP_TRY(gp_ensure_ws(c, ((size_t)c->n + (size_t)splits * pstride) * W * sizeof(double)));
    p.psi1_part = c->ws;
    p.partial = c->ws + (size_t)c->n * W;
    p.pstride = pstride; p.pbase = pbase;
    p.s_pos = c->s_pos; p.s_sig = c->s_sig; p.gx_mu = c->gx_mu; p.gx_s = c->gx_s; p.grad_latest = c->grad_latest;
    p.fuse_finish = 0;
    {   // Psi1 side: Y row in registers / G1 in shared memory when they fit
        const bool fits = (size_t)c->M * (Q + 16) * sizeof(double) <= (size_t)96 * 1024;
        if (fits && c->D <= 4) GP_TRY((launch_psi1_part<Q, 4>(c, p, cnt)));
        else if (fits && c->D <= 10) GP_TRY((launch_psi1_part<Q, 10>(c, p, cnt)));
        else if (fits && c->D <= 16) GP_TRY((launch_psi1_part<Q, 16>(c, p, cnt)));
        else GP_TRY((launch_psi1_part<Q, 0>(c, p, cnt)));
    }
    if (fp32) {
        GP_TRY(gp_launch_embed_psi2_f32(c, p.m_bounds, splits, p.partial, i0, i1));   // opt-in fp32 evaluation of the Psi2 part
    } else {
        if (use_m) {
            if (c->pair_ra_stale) {
                GP_TRY(gp_launch_pair_ra(c));
                c->pair_ra_stale = false;
            }
            // one launch: whole tiles first, then the (tile, pair split) CTAs of the split part
            EmbedParams pm = p;
            pm.full_tiles = (int)full_tiles;
            pm.tail_splits = splits;
            for (int s = 0; s <= splits; ++s) pm.p_bounds[s] = (int)(pchunks * s / splits);      // chunk splits
#ifndef EMB_NO_FUSE
            pm.fuse_finish = splits == 1 ? 1 : 0;
#endif
            GP_TRY(gp_launch_embed_psi2m(c, pm, (int)(full_tiles + tail_tiles * splits)));
            if (pm.fuse_finish) return GPARML_OK;
        } else {
            if (full_tiles > 0) {                              // whole rounds, one pair range, fused finish
                EmbedParams pf = p;
                pf.i1 = t0;
                pf.p_bounds[0] = 0; pf.p_bounds[1] = (int)Pn;
                pf.fuse_finish = 1;
                GP_TRY(gp_launch_embed_psi2x(c, pf, (int)full_tiles, 1));
            }
            if (tail_cnt > 0) {
                EmbedParams pt = p;
                pt.i0 = t0;
                for (int s = 0; s <= splits; ++s) pt.p_bounds[s] = (int)(Pn * s / splits);   // pair splits
#ifndef EMB_NO_FUSE
                pt.fuse_finish = splits == 1 ? 1 : 0;
#endif
                GP_TRY(gp_launch_embed_psi2x(c, pt, (int)tail_tiles, splits));
                if (pt.fuse_finish) return GPARML_OK;
            }
        }
    }
    if (tail_cnt <= 0) return GPARML_OK;
    const int64_t f0 = fp32 ? i0 : t0, fcnt = fp32 ? cnt : tail_cnt;
    embed_finish_kernel<<<(int)((fcnt * Q + 255) / 256), 256, 0, c->stream>>>(p.partial, splits, pstride, pbase, p.psi1_part, c->n, f0, fcnt, Q,
                                                                          gp_rec_len(Q), c->rec2, c->s_pos, c->s_sig, c->gx_mu, c->gx_s,
                                                                          c->grad_latest, fp32 ? 0 : 1, c->d_glob);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_embed_grads(gparml_ctx *c) { return gp_launch_embed_grads_range(c, 0, c->n); }

int gp_launch_embed_grads_range(gparml_ctx *c, int64_t i0, int64_t i1)
{
    if (i1 <= i0) return GPARML_OK;
    switch (c->Q) {
#define CASE_Q(q) case q: return launch_q<q>(c, i0, i1);
        CASE_Q(1) CASE_Q(2) CASE_Q(3) CASE_Q(4) CASE_Q(5) CASE_Q(6) CASE_Q(7) CASE_Q(8)
        CASE_Q(9) CASE_Q(10) CASE_Q(11) CASE_Q(12) CASE_Q(13) CASE_Q(14) CASE_Q(15) CASE_Q(16)
#undef CASE_Q
    }
    gp_set_error("embed_grads: unsupported Q=%d (1..%d)", c->Q, GP_MAX_Q);
    return GPARML_ERR_ARG;
}
