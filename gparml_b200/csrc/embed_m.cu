// K5 (Psi2 part) on the FP64 tensor-core instruction: embed_psi2m_kernel.
//
// Replaces partial_terms.py:367-431 (the Psi2 terms of grad_X_mu / grad_X_S), like embed_psi2x (embed_x.cu), whose
// arithmetic it shares: in the expanded basis centred on c = column means of Z (mc = mu - c, zc = zbar - c)
//   exponent[n, p] = kn_n + sum_q (2 w mc)_nq zc_pq + sum_q (-w)_nq zc_pq^2              h[n, p] = exp(exponent)
//   sums[n, :]     = sum_p h[n, p] G_p (zc_p1 .. zc_pQ, zc_p1^2 .. zc_pQ^2, 1),   G_p = Gs_p exp(lk_p) = sign exp(lk + log|Gs|)
// Both lines are matrix products between a per-point feature matrix X (n x 2Q) and a per-pair feature matrix
// R (P x (2Q + 1)), with an exp between them:
//   E = X R^T        (8 points x 8 pairs per warp step, K = 2Q:  ceil(2Q / 4) mma.sync m8n8k4 f64)
//   S += h (G R)     (8 points x 8 NT features, K = 8 pairs:     2 NT MMAs, NT = ceil((2Q + 1) / 8))
// The pair factor G is folded into the second product's matrix (pair_ra = G R, rebuilt by pair_ra_kernel after every
// master step: 1 MB at c3) instead of into the exponent: no per-item add of lk + log|Gs|, no sign handling.
// On B200 the FP64 MMA runs at the DFMA rate (same pipe) but reads 4 register operands per 256 FMAs, so it does
// not hit the register-file limit that holds the DFMA formulation at ~75 % of the pipe.  Per
// 64 (point, pair) items at Q = 10: 5 + 6 MMAs (176 pipe cycles) + 2 x 6 DFMA-pipe instructions per lane for the
// two exps (24 cycles) = 200 cycles; embed_psi2x issues 49 DFMAs per item = 196 cycles at 100 % and needs 264.
// Measured (B200, c3, tools/tune.py, embed phase incl. 0.7 ms of embed_psi1): embed_psi2x 18.4 ms; this kernel 17.7 with
// the 7-instruction exp of gp_exp.cuh and lk + log|Gs| added to the exponent per item (255 cycles per 64 items), 17.1
// with the 6-instruction exp below, 16.3 with the pair factor folded into G R (231 cycles).  Without the exps the MMA
// stream runs at 96 % (184 cycles): every DFMA-pipe instruction between the MMAs costs ~4.4 cycles instead of 2 because
// it queues behind the other warps' 16-cycle MMAs (ncu: stall reason math_pipe_throttle) -- hence the effort to remove
// them.  Tried without gain: all warps of an SM in the same phase (block barriers around the exps), the exps of 2 / 4
// steps evaluated together (more independent chains), one row group per warp with 3 CTAs, 4- / 6- / 12- / 16-warp CTAs.
//
// Fragments (lane = 4 g + k):  A[row g][col k], B[row k][col g], C[row g][cols 2k, 2k + 1].
//   E step s:   A = X[point g][feature 4s + k] (registers, constant over the pair loop),
//               B = R[pair pi(g)][feature 4s + k], pi(c) = c / 2 + 4 (c % 2)
//   -> lane (g, k) holds the exponents of point g for pairs k and 4 + k of the step's 8 pairs: exactly the A
//      fragments of the two accumulation MMAs (K = pairs 0..3, then 4..7) -- h never leaves its lane.
//   S tile t:   B = R[pair k (resp. 4 + k)][feature 8t + g]
// R lives in shared memory as [feature / 4][pair][4]: both access patterns are conflict-free (32 lanes = 32
// consecutive doubles resp. two runs of 16).  R is point-independent and static per set_globals (pair_table_kernel
// writes it in 64-pair chunks; a stage = one 1-D bulk copy of its first ceil(2Q/4) blocks plus one of the chunk of G R).
#include <math.h>

#include "embed.cuh"
#include "gp_exp.cuh"

#ifndef EMBM_WARPS
#define EMBM_WARPS 8
#endif
#ifndef EMBM_MINB
#define EMBM_MINB 2
#endif
#ifndef EMBM_NG
#define EMBM_NG 2              // row groups of 8 points per warp: every B fragment read from shared memory feeds EMBM_NG MMAs
#endif
#ifndef EMBM_US
#define EMBM_US 1               // steps of 8 pairs whose exps are evaluated together
#endif
#define EMBM_STAGES 2
// exp of this kernel: every DFMA-pipe instruction of it queues behind the other warps' MMAs, so the table is larger
// (4096 entries = 2^(j/4096), built per CTA as the product of the 256-entry table and 16 sub-steps: one more rounding,
// 1.1e-16) and the polynomial shorter (degree 2, near-minimax on |r| <= ln2/8192: 2.5e-14) than in gp_exp.cuh:
// 6 instead of 7 FP64 instructions (EMBM_EXP12 0 keeps the 256-entry table in 16 conflict-free copies)
#ifndef EMBM_EXP12
#define EMBM_EXP12 1
#endif
#if EMBM_EXP12
#define EMBM_TAB_ENTRIES 4096
#define EMBM_TAB_REP 1
#define EMBM_LOG2_TAB 12
#define EMBM_SCALE 5909.278887481194              // 4096 / ln2
#define EMBM_NEG_STEP -0.0001692253858788929      // -ln2 / 4096
#define EMBM_C1 1.0000000008949057
#define EMBM_C2 0.500000000025057
static __device__ const double embm_sub_table[16] = {
    1.0, 1.0001692397053021, 1.0003385080526823, 1.0005078050469876, 1.0006771306930664, 1.0008464849957674,
    1.001015867959941, 1.0011852795904375, 1.0013547198921082, 1.0015241888698057, 1.0016936865283832,
    1.0018632128726943, 1.002032767907594, 1.002202351637938, 1.0023719640685822, 1.0025416052043845};
#else
#define EMBM_TAB_ENTRIES GP_EXP_TAB
#define EMBM_TAB_REP 16        // copies of the exp table, one per 8-byte bank slot: lane l reads copy l % 16, no bank conflicts
#define EMBM_LOG2_TAB GP_EXP_LOG2_TAB
#define EMBM_SCALE GP_EXP_SCALE
#define EMBM_NEG_STEP GP_EXP_NEG_STEP
#endif

__device__ __forceinline__ void embm_dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int Q>
__global__ void __launch_bounds__(EMBM_WARPS * 32, EMBM_MINB)
embed_psi2m_kernel(EmbedParams p)
{
    constexpr int R2 = (3 * Q + 2) & ~1;                 // Psi2 record length (prep_points)
    constexpr int KS = (2 * Q + 3) / 4;                  // K steps of the exponent product
    constexpr int NT = GP_PAIR_R_TILES(Q), NB = 2 * NT;  // feature tiles of 8 / blocks of 4
    constexpr int CP = GP_PAIR_CHUNK, NG = EMBM_NG, US = EMBM_US;
    constexpr int RD = NB * CP * 4;                      // doubles of a chunk of R resp. G R
    constexpr int RE = KS * CP * 4;                      // of which the exponent product reads the first KS blocks
    constexpr int STG = RE + RD;                         // doubles per stage: [R blocks 0..KS) | G R]
    constexpr int OUTW = 8 * NT + 1;                     // padded row of the epilogue staging
    constexpr int RING_D = EMBM_STAGES * STG;
    constexpr int OUT_D = EMBM_WARPS * NG * 8 * OUTW;
    extern __shared__ __align__(16) double smem[];       // ring: [stage][R | G R]; reused by the epilogue; then the exp table
    __shared__ __align__(8) uint64_t bar[EMBM_STAGES];
    double *exp_tab = smem + (RING_D > OUT_D ? RING_D : OUT_D);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, k = lane & 3;
    // whole tiles first (they are dispatched first), then the split ones: small CTAs that fill the machine at the end
    const bool whole = (int)blockIdx.x < p.full_tiles;
    const int rest = (int)blockIdx.x - p.full_tiles;
    const int tile = whole ? (int)blockIdx.x : p.full_tiles + rest / p.tail_splits;
    const int split = whole ? 0 : rest % p.tail_splits;
    const int c_lo = whole ? 0 : p.p_bounds[split], c_hi = whole ? p.p_bounds[p.tail_splits] : p.p_bounds[split + 1];      // chunks of CP pairs
    const int nchunks = c_hi - c_lo;

#if EMBM_EXP12
    static_assert(GP_EXP_LOG2_TAB == 8, "the 4096-entry table is built from the 256-entry one");
    for (int idx = tid; idx < EMBM_TAB_ENTRIES; idx += EMBM_WARPS * 32) exp_tab[idx] = gp_exp_table_const[idx >> 4] * embm_sub_table[idx & 15];
#else
    for (int idx = tid; idx < GP_EXP_TAB * EMBM_TAB_REP; idx += EMBM_WARPS * 32) exp_tab[idx] = gp_exp_table_const[idx / EMBM_TAB_REP];
#endif
    if (tid == 0) {
        for (int s = 0; s < EMBM_STAGES; ++s) gp_mbar_init(&bar[s], 1);
        gp_fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int s = t % EMBM_STAGES;
        const size_t chunk = (size_t)(c_lo + t);
        double *dst = smem + (size_t)s * STG;
        gp_mbar_expect_tx(&bar[s], (uint32_t)(STG * sizeof(double)));
        gp_bulk_g2s(dst, p.pair_r + chunk * RD, (uint32_t)(RE * sizeof(double)), &bar[s]);
        gp_bulk_g2s(dst + RE, p.pair_ra + chunk * RD, (uint32_t)(RD * sizeof(double)), &bar[s]);
    };
    if (tid == 0)
        for (int t = 0; t < EMBM_STAGES && t < nchunks; ++t) issue(t);

    // ---- per-point features: X[f] = 2 w mc (f < Q), -w (Q <= f < 2Q), 0 beyond; kn = lc2 - sum_q w mc^2 ----------
    const int64_t i_base = p.i0 + ((int64_t)tile * EMBM_WARPS + warp) * (8 * NG);
    double kn[NG], xa[NG][KS], acc[NG][NT][2];
#pragma unroll
    for (int u = 0; u < NG; ++u) {
        int64_t i = i_base + 8 * u + g;
        if (i >= p.i1) i = p.i1 - 1;                     // compute on a real record, never store
        const double2 *r2 = reinterpret_cast<const double2 *>(p.rec2 + i * R2);
        kn[u] = p.rec2[i * R2 + 3 * Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double2 mw = r2[q];
            const double mc = mw.x - p.glob->center[q];
            kn[u] = fma(-(mw.y * mc), mc, kn[u]);
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int f = 4 * s + k;
            double v = 0.0;
            if (f < 2 * Q) {
                const int q = f < Q ? f : f - Q;
                const double2 mw = r2[q];
                v = f < Q ? 2.0 * (mw.y * (mw.x - p.glob->center[q])) : -mw.y;
            }
            xa[u][s] = v;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[u][t][0] = acc[u][t][1] = 0.0;
    }

    // per-lane offsets into a stage (doubles)
    const int off_e = ((g >> 1) + 4 * (g & 1)) * 4 + k;                  // E product: pair pi(g), feature k of block s
    const int off_s = (g >> 2) * (CP * 4) + k * 4 + (g & 3);             // S product: pair k, feature g of tile t
    const double *tab = exp_tab + (lane & (EMBM_TAB_REP - 1));

    for (int t = 0; t < nchunks; ++t) {
        const int s = t % EMBM_STAGES;
        gp_mbar_wait(&bar[s], (uint32_t)((t / EMBM_STAGES) & 1));
        const double *R = smem + (size_t)s * STG;
        const double *RA = R + RE;                       // G R
        // EMBM_US steps of 8 pairs at a time: all their exponent MMAs, then all their exps interleaved step by step (a
        // dependent DFMA queues behind the other warps' MMAs -- ~36 cycles per step of the chain -- so the pipe stays
        // fed only with enough independent chains in flight: 2 NG EMBM_US per warp), then all their accumulation MMAs
#pragma unroll 1
        for (int j0 = 0; j0 < CP; j0 += 8 * US) {
            constexpr int NE = US * NG * 2;
            double e[NE];
#pragma unroll
            for (int v = 0; v < US; ++v) {
                const double *re = R + (j0 + 8 * v) * 4 + off_e;
#pragma unroll
                for (int u = 0; u < NG; ++u) e[(v * NG + u) * 2] = e[(v * NG + u) * 2 + 1] = kn[u];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const double b = re[ks * (CP * 4)];
#pragma unroll
                    for (int u = 0; u < NG; ++u) {
                        double (&c2)[2] = *reinterpret_cast<double (*)[2]>(&e[(v * NG + u) * 2]);
                        embm_dmma(c2, xa[u][ks], b);
                    }
                }
            }
            // exp of the NE exponents in lockstep (gp_exp.cuh, same constants and result as gp_exp_signed)
            double tt[NE], rr[NE], pl[NE];
            int kk[NE];
#pragma unroll
            for (int x = 0; x < NE; ++x) e[x] = gp_exp_clamp(e[x]);
#pragma unroll
            for (int x = 0; x < NE; ++x) tt[x] = fma(e[x], EMBM_SCALE, GP_EXP_SHIFT);
#pragma unroll
            for (int x = 0; x < NE; ++x) { kk[x] = __double2loint(tt[x]); tt[x] = tt[x] - GP_EXP_SHIFT; }
#pragma unroll
            for (int x = 0; x < NE; ++x) { rr[x] = fma(tt[x], EMBM_NEG_STEP, e[x]); tt[x] = tab[(kk[x] & (EMBM_TAB_ENTRIES - 1)) * EMBM_TAB_REP]; }
#if EMBM_EXP12
#pragma unroll
            for (int x = 0; x < NE; ++x) pl[x] = fma(rr[x], EMBM_C2, EMBM_C1);
#pragma unroll
            for (int x = 0; x < NE; ++x) pl[x] = fma(pl[x], rr[x], 1.0);
#else
#pragma unroll
            for (int x = 0; x < NE; ++x) pl[x] = fma(rr[x], GP_EXP_C3, GP_EXP_C2);
#pragma unroll
            for (int x = 0; x < NE; ++x) pl[x] = fma(pl[x], rr[x], GP_EXP_C1);
#pragma unroll
            for (int x = 0; x < NE; ++x) pl[x] = fma(pl[x], rr[x], GP_EXP_C0);
#endif
#pragma unroll
            for (int x = 0; x < NE; ++x) pl[x] = tt[x] * pl[x];
#pragma unroll
            for (int x = 0; x < NE; ++x) {
                int m = kk[x] >> EMBM_LOG2_TAB;
                m = m < -1021 ? -1021 : m;
                e[x] = __hiloint2double(__double2hiint(pl[x]) + (m << 20), __double2loint(pl[x]));
            }
#pragma unroll
            for (int v = 0; v < US; ++v) {
                const double *rs = RA + (j0 + 8 * v) * 4 + off_s;
#pragma unroll
                for (int tt2 = 0; tt2 < NT; ++tt2) {
                    const double b0 = rs[tt2 * (2 * CP * 4)], b1 = rs[tt2 * (2 * CP * 4) + 16];
#pragma unroll
                    for (int u = 0; u < NG; ++u) {
                        embm_dmma(acc[u][tt2], e[(v * NG + u) * 2], b0);
                        embm_dmma(acc[u][tt2], e[(v * NG + u) * 2 + 1], b1);
                    }
                }
            }
        }
        __syncthreads();      // every warp is done reading stage s
        if (tid == 0 && t + EMBM_STAGES < nchunks) issue(t + EMBM_STAGES);
    }

    // ---- epilogue: the sums of the warp through shared memory (the ring is free now), then one (point, q) per lane-task
    double *out = smem + (size_t)warp * NG * 8 * OUTW;
#pragma unroll
    for (int u = 0; u < NG; ++u)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            out[(8 * u + g) * OUTW + 8 * t + 2 * k] = acc[u][t][0];
            out[(8 * u + g) * OUTW + 8 * t + 2 * k + 1] = acc[u][t][1];
        }
    __syncwarp();
    if (whole || p.fuse_finish) {
        for (int task = lane; task < 8 * NG * Q; task += 32) {
            const int gg = task / Q, q = task - gg * Q;
            const int64_t ii = i_base + gg;
            if (ii >= p.i1) continue;
            const double2 mw = *reinterpret_cast<const double2 *>(p.rec2 + ii * R2 + 2 * q);
            const double *p1 = p.psi1_part + (size_t)ii * (2 * Q + 1);
            const int64_t o = ii * Q + q;
            gp_embed_finish_one(mw.x, mw.y, mw.x - p.glob->center[q], out[gg * OUTW + q], out[gg * OUTW + Q + q],
                                out[gg * OUTW + 2 * Q], p1[q], p1[Q + q], p.s_pos[o], p.s_sig[o], p.gx_mu + o, p.gx_s + o,
                                p.grad_latest + o, p.grad_latest + p.n * Q + o);
        }
        return;
    }
    for (int task = lane; task < 8 * NG * (2 * Q + 1); task += 32) {
        const int gg = task / (2 * Q + 1), f = task - gg * (2 * Q + 1);
        const int64_t ii = i_base + gg;
        if (ii >= p.i1) continue;
        p.partial[((size_t)split * p.pstride + (ii - p.pbase)) * (2 * Q + 1) + f] = out[gg * OUTW + f];
    }
}

// pair_ra = G_p pair_r, G_p = sign(Gs) exp(lk + log|Gs|) from pair_h (written by the head of the master step); one thread
// per (pair, block of 4 features).  Pairs beyond P keep their zeros.
__global__ void __launch_bounds__(256) pair_ra_kernel(const double *__restrict__ pair_r, const double2 *__restrict__ pair_h,
                                                      double *__restrict__ pair_ra, int64_t P, int NB)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pr = idx / NB;
    const int blk = (int)(idx - pr * NB);
    if (pr >= P) return;
    const double2 h = pair_h[pr];
    const double gf = h.y < 0.0 ? -exp(h.x) : exp(h.x);
    const size_t off = ((size_t)(pr / GP_PAIR_CHUNK) * NB + blk) * GP_PAIR_CHUNK * 4 + (size_t)(pr % GP_PAIR_CHUNK) * 4;
    const double2 a = *reinterpret_cast<const double2 *>(pair_r + off), b = *reinterpret_cast<const double2 *>(pair_r + off + 2);
    *reinterpret_cast<double2 *>(pair_ra + off) = make_double2(gf * a.x, gf * a.y);
    *reinterpret_cast<double2 *>(pair_ra + off + 2) = make_double2(gf * b.x, gf * b.y);
}

int gp_launch_pair_ra(gparml_ctx *c)
{
    const int NB = 2 * GP_PAIR_R_TILES(c->Q);
    const int64_t total = c->L.P * NB;
    pair_ra_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(c->pair_r, c->pair_h, c->pair_ra, c->L.P, NB);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_embed_psi2m_points_per_cta() { return EMBM_WARPS * 8 * EMBM_NG; }

template <int Q> static size_t smem_m()
{
    constexpr int NT = GP_PAIR_R_TILES(Q), RD = 2 * NT * GP_PAIR_CHUNK * 4, RE = ((2 * Q + 3) / 4) * GP_PAIR_CHUNK * 4;
    const size_t ring = (size_t)EMBM_STAGES * (RE + RD), outd = (size_t)EMBM_WARPS * EMBM_NG * 8 * (8 * NT + 1);
    return ((ring > outd ? ring : outd) + (size_t)EMBM_TAB_ENTRIES * EMBM_TAB_REP) * sizeof(double);
}
template <int Q> static int occ_m(int *occ)
{
    GP_CUDA(cudaFuncSetAttribute(embed_psi2m_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m<Q>()));
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, embed_psi2m_kernel<Q>, EMBM_WARPS * 32, smem_m<Q>()));
    return GPARML_OK;
}
template <int Q> static int launch_m(gparml_ctx *c, const EmbedParams &p, int ctas)
{
    embed_psi2m_kernel<Q><<<(unsigned)ctas, EMBM_WARPS * 32, smem_m<Q>(), c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

#define EMBM_ALL_Q(F) F(5) F(6) F(7) F(8) F(9) F(10)

int gp_embed_psi2m_occupancy(int Q, int *occ)
{
    switch (Q) {
#define CASE_Q(q) case q: return occ_m<q>(occ);
        EMBM_ALL_Q(CASE_Q)
#undef CASE_Q
    }
    gp_set_error("embed_psi2m: unsupported Q=%d", Q);
    return GPARML_ERR_ARG;
}

int gp_launch_embed_psi2m(gparml_ctx *c, const EmbedParams &p, int ctas)
{
    switch (c->Q) {
#define CASE_Q(q) case q: return launch_m<q>(c, p, ctas);
        EMBM_ALL_Q(CASE_Q)
#undef CASE_Q
    }
    gp_set_error("embed_psi2m: unsupported Q=%d", c->Q);
    return GPARML_ERR_ARG;
}
