// K2 / K5, opt-in fp32 evaluation (GPARML_FLAG_FP32_MAP).
//
// Same decomposition as psi2.cu / embed.cu, but every per-(point, pair) quantity is evaluated in
// fp32 (FFMA pipe, MUFU ex2) and only the sums are kept in fp64:
//   psi2_stats_f32 : fp32 partial sums over one 64-point tile, flushed into fp64 accumulators;
//   embed_psi2_f32 : fp32 partial sums over one row of pairs (m fixed), flushed into fp64.
// The fp32 records hold mu - c and the kernels use z - c (c = column means of Z), which keeps
// the subtraction mu - zbar away from fp32 cancellation when the latent space is not centred.
// psi1_stats, the Psi1 side of embed_grads, prep_points and the master step stay fp64.
//
// Stated tolerance of this path (tests/test_gpu_fp32.py): 2e-6 relative (max-norm) on the
// summed statistics, 2e-8 on F, 5e-6 on the gradients (observed 1e-7 / 9e-10 / 2e-7).  The fp64 path is the default and the
// only one the 1e-9 parity claim applies to.
#include <math.h>

#include "common.cuh"

#define F32_THREADS 256
#define F32_TN 64
#define F32_STAGES 2
#ifndef F32_PAIRS
#define F32_PAIRS 2
#endif

__device__ __forceinline__ float gp_expf_fast(float x) { return exp2f(x * 1.4426950408889634f); }

template <int Q>
__global__ void __launch_bounds__(F32_THREADS, (F32_PAIRS == 1) ? 2 : 1)
psi2_stats_f32_kernel(const float *__restrict__ rec2f, int64_t n, const double *__restrict__ Z,
                      const GlobalsDev *__restrict__ glob, int64_t P, const int2 *__restrict__ pair_idx,
                      const double *__restrict__ pair_lk, int64_t n_per_split, double *__restrict__ partial)
{
    constexpr int RF = (3 * Q + 4) & ~3;
    constexpr int PP = F32_PAIRS;
    extern __shared__ __align__(16) float tilef[];        // [STAGES][TN][RF]
    __shared__ __align__(8) uint64_t bar[F32_STAGES];
    const int tid = threadIdx.x;
    int64_t p[PP];
    bool valid[PP];
    float lk[PP], zb[PP][Q], accf[PP][1 + 2 * Q];
    double accd[PP][1 + 2 * Q];
#pragma unroll
    for (int u = 0; u < PP; ++u) {
        p[u] = ((int64_t)blockIdx.x * PP + u) * F32_THREADS + tid;
        valid[u] = p[u] < P;
        const int2 ab = valid[u] ? pair_idx[p[u]] : make_int2(0, 0);
        lk[u] = valid[u] ? (float)pair_lk[p[u]] : 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) zb[u][q] = (float)(0.5 * (Z[ab.x * Q + q] + Z[ab.y * Q + q]) - glob->center[q]);
#pragma unroll
        for (int j = 0; j < 1 + 2 * Q; ++j) { accf[u][j] = 0.f; accd[u][j] = 0.0; }
    }
    const int64_t n_lo = (int64_t)blockIdx.y * n_per_split;
    const int64_t n_hi = (n_lo + n_per_split < n) ? (n_lo + n_per_split) : n;
    const int64_t span = n_hi > n_lo ? n_hi - n_lo : 0;
    const int ntiles = (int)((span + F32_TN - 1) / F32_TN);
    if (tid == 0) {
        for (int s = 0; s < F32_STAGES; ++s) gp_mbar_init(&bar[s], 1);
        gp_fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < F32_STAGES && s < ntiles; ++s) {
            const int64_t base = n_lo + (int64_t)s * F32_TN;
            const int cnt = (int)((n_hi - base < F32_TN) ? (n_hi - base) : F32_TN);
            const uint32_t bytes = (uint32_t)cnt * RF * sizeof(float);
            gp_mbar_expect_tx(&bar[s], bytes);
            gp_bulk_g2s(tilef + (size_t)s * F32_TN * RF, rec2f + base * RF, bytes, &bar[s]);
        }
    }
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % F32_STAGES;
        const uint32_t parity = (uint32_t)((t / F32_STAGES) & 1);
        const int64_t base = n_lo + (int64_t)t * F32_TN;
        const int cnt = (int)((n_hi - base < F32_TN) ? (n_hi - base) : F32_TN);
        gp_mbar_wait(&bar[s], parity);
        const float *tb = tilef + (size_t)s * F32_TN * RF;
#pragma unroll 2
        for (int i = 0; i < cnt; ++i) {
            const float *rec = tb + i * RF;
            const float2 *r2 = reinterpret_cast<const float2 *>(rec);
            float wd[PP][Q], e[PP], psi[PP];
            const float lc2 = rec[3 * Q];
#pragma unroll
            for (int u = 0; u < PP; ++u) e[u] = lk[u] + lc2;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float2 mw = r2[q];               // (mu_q - c_q, w_q), broadcast
#pragma unroll
                for (int u = 0; u < PP; ++u) {
                    const float d = mw.x - zb[u][q];
                    wd[u][q] = mw.y * d;
                    e[u] = fmaf(-wd[u][q], d, e[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < PP; ++u) {
                psi[u] = gp_expf_fast(e[u]);
                accf[u][0] += psi[u];
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float v = rec[2 * Q + q];
#pragma unroll
                for (int u = 0; u < PP; ++u) {
                    accf[u][1 + q] = fmaf(psi[u], wd[u][q], accf[u][1 + q]);
                    accf[u][1 + Q + q] = fmaf(psi[u], fmaf(wd[u][q], wd[u][q], v), accf[u][1 + Q + q]);
                }
            }
        }
        // flush the tile's fp32 partial sums into the fp64 accumulators
#pragma unroll
        for (int u = 0; u < PP; ++u)
#pragma unroll
            for (int j = 0; j < 1 + 2 * Q; ++j) { accd[u][j] += (double)accf[u][j]; accf[u][j] = 0.f; }
        __syncthreads();
        if (tid == 0 && t + F32_STAGES < ntiles) {
            const int64_t nb = n_lo + (int64_t)(t + F32_STAGES) * F32_TN;
            const int ncnt = (int)((n_hi - nb < F32_TN) ? (n_hi - nb) : F32_TN);
            const uint32_t bytes = (uint32_t)ncnt * RF * sizeof(float);
            gp_mbar_expect_tx(&bar[s], bytes);
            gp_bulk_g2s(tilef + (size_t)s * F32_TN * RF, rec2f + nb * RF, bytes, &bar[s]);
        }
    }
#pragma unroll
    for (int u = 0; u < PP; ++u) {
        if (valid[u]) {
            double *out = partial + (size_t)blockIdx.y * (1 + 2 * Q) * P + p[u];
#pragma unroll
            for (int j = 0; j < 1 + 2 * Q; ++j) out[(size_t)j * P] = accd[u][j];
        }
    }
}

static __global__ void __launch_bounds__(256) psi2_reduce_f32path_kernel(const double *__restrict__ partial, int splits,
                                                                         int64_t rows_x_P, double *__restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_x_P) return;
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += partial[(size_t)s * rows_x_P + i];
    dst[i] = a;
}

template <int Q>
static int launch_psi2_f32(gparml_ctx *c)
{
    constexpr int RF = (3 * Q + 4) & ~3;
    const size_t smem = (size_t)F32_STAGES * F32_TN * RF * sizeof(float);
    int occ = 1;
    GP_CUDA(cudaFuncSetAttribute(psi2_stats_f32_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, psi2_stats_f32_kernel<Q>, F32_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t P = c->L.P;
    const int tiles = (int)((P + F32_THREADS * F32_PAIRS - 1) / (F32_THREADS * F32_PAIRS));
    const int64_t slots = (int64_t)c->sm_count * occ;
    const int64_t rows_x_P = (int64_t)(1 + 2 * Q) * P;
    int64_t max_splits = (c->n + 4 * F32_TN - 1) / (4 * F32_TN);
    const int64_t ws_cap = ((int64_t)256 << 20) / (rows_x_P * (int64_t)sizeof(double));
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
    const int splits = (int)best;
    const int64_t n_per_split = (c->n + splits - 1) / splits;
    GP_TRY(gp_ensure_ws(c, (size_t)splits * rows_x_P * sizeof(double)));
    dim3 grid(tiles, splits);
    psi2_stats_f32_kernel<Q><<<grid, F32_THREADS, smem, c->stream>>>(c->rec2f, c->n, c->Z, c->d_glob, P, c->pair_idx, c->pair_lk,
                                                                    n_per_split, c->ws);
    GP_LAUNCH_CHECK(c);
    psi2_reduce_f32path_kernel<<<(int)((rows_x_P + 255) / 256), 256, 0, c->stream>>>(c->ws, splits, rows_x_P, c->stats + c->L.off_s0);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_psi2_stats_f32(gparml_ctx *c)
{
    switch (c->Q) {
#define CASE_Q(q) case q: return launch_psi2_f32<q>(c);
        CASE_Q(1) CASE_Q(2) CASE_Q(3) CASE_Q(4) CASE_Q(5) CASE_Q(6) CASE_Q(7) CASE_Q(8)
        CASE_Q(9) CASE_Q(10) CASE_Q(11) CASE_Q(12) CASE_Q(13) CASE_Q(14) CASE_Q(15) CASE_Q(16)
#undef CASE_Q
    }
    gp_set_error("psi2_stats (fp32): unsupported Q=%d", c->Q);
    return GPARML_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------
// embed_psi2 in fp32: same sqrt(w) basis as embed.cu; fp32 sums over one row of pairs.
// Output layout identical to the fp64 kernel: partial[split][n][2Q + 1] (AM, AS, AH) in fp64.
// ---------------------------------------------------------------------------------------------
#define EMB32_THREADS 128
#define EMB32_MAX_SPLITS 32

struct Embed32Params {
    const float *rec2f;
    const double *Z;
    const GlobalsDev *glob;
    const double2 *pair_g;
    int64_t n, i0, i1;     // shard size (stride of the partial buffer) and the point range of this launch
    int M;
    int m_bounds[EMB32_MAX_SPLITS + 1];
    double *partial;
};

template <int Q>
__global__ void __launch_bounds__(EMB32_THREADS, 4) embed_psi2_f32_kernel(Embed32Params p)
{
    constexpr int RF = (3 * Q + 4) & ~3;
    extern __shared__ __align__(16) float hzf[];          // [M][Q] = (Z - c) / 2
    const int tid = threadIdx.x, M = p.M;
    for (int idx = tid; idx < M * Q; idx += EMB32_THREADS) hzf[idx] = (float)(0.5 * (p.Z[idx] - p.glob->center[idx % Q]));
    __syncthreads();
    int64_t i = p.i0 + (int64_t)blockIdx.x * EMB32_THREADS + tid;
    const bool valid = i < p.i1;
    if (!valid) i = p.i1 - 1;
    const float *rec = p.rec2f + i * RF;
    const float lc2 = rec[3 * Q];
    float sw[Q], mu[Q], sdm[Q], u[Q], amf[Q], asf[Q];
    double am[Q], as[Q], ah = 0.0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        mu[q] = rec[2 * q];
        sw[q] = sqrtf(rec[2 * q + 1]);
        am[q] = 0.0;
        as[q] = 0.0;
    }
    const int m_lo = p.m_bounds[blockIdx.y], m_hi = p.m_bounds[blockIdx.y + 1];
    for (int m = m_lo; m < m_hi; ++m) {
        const float *hm = hzf + m * Q;
        const double2 *pg = p.pair_g + gp_pair_index(M, m, m);
        double2 g = __ldg(pg);
        float ahf = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) { sdm[q] = sw[q] * (mu[q] - hm[q]); amf[q] = 0.f; asf[q] = 0.f; }
#pragma unroll 2
        for (int b = m; b < M; ++b) {
            const double2 gn = __ldg(pg + ((b + 1 < M) ? (b + 1 - m) : (b - m)));
            const float *hb = hzf + b * Q;
            float e = (float)g.x + lc2;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                u[q] = fmaf(-sw[q], hb[q], sdm[q]);
                e = fmaf(-u[q], u[q], e);
            }
            const float h = (float)g.y * gp_expf_fast(e);
            ahf += h;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float t = h * u[q];
                amf[q] += t;
                asf[q] = fmaf(t, u[q], asf[q]);
            }
            g = gn;
        }
        ah += (double)ahf;
#pragma unroll
        for (int q = 0; q < Q; ++q) { am[q] += (double)amf[q]; as[q] += (double)asf[q]; }
    }
    if (valid) {
        double *out = p.partial + ((size_t)blockIdx.y * p.n + i) * (2 * Q + 1);
#pragma unroll
        for (int q = 0; q < Q; ++q) { out[q] = am[q]; out[Q + q] = as[q]; }
        out[2 * Q] = ah;
    }
}

template <int Q>
static int launch_embed_f32(gparml_ctx *c, const int *m_bounds, int splits, double *partial, int64_t i0, int64_t i1)
{
    Embed32Params p;
    p.rec2f = c->rec2f; p.Z = c->Z; p.glob = c->d_glob; p.pair_g = c->pair_g; p.n = c->n; p.i0 = i0; p.i1 = i1; p.M = c->M; p.partial = partial;
    for (int s = 0; s <= splits; ++s) p.m_bounds[s] = m_bounds[s];
    const size_t smem = (size_t)c->M * Q * sizeof(float);
    GP_CUDA(cudaFuncSetAttribute(embed_psi2_f32_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((i1 - i0 + EMB32_THREADS - 1) / EMB32_THREADS), splits);
    embed_psi2_f32_kernel<Q><<<grid, EMB32_THREADS, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_embed_psi2_f32(gparml_ctx *c, const int *m_bounds, int splits, double *partial, int64_t i0, int64_t i1)
{
    switch (c->Q) {
#define CASE_Q(q) case q: return launch_embed_f32<q>(c, m_bounds, splits, partial, i0, i1);
        CASE_Q(1) CASE_Q(2) CASE_Q(3) CASE_Q(4) CASE_Q(5) CASE_Q(6) CASE_Q(7) CASE_Q(8)
        CASE_Q(9) CASE_Q(10) CASE_Q(11) CASE_Q(12) CASE_Q(13) CASE_Q(14) CASE_Q(15) CASE_Q(16)
#undef CASE_Q
    }
    gp_set_error("embed_psi2 (fp32): unsupported Q=%d", c->Q);
    return GPARML_ERR_ARG;
}
