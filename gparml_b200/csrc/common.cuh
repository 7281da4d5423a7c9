// Internal declarations shared by the gparml_b200 translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/gparml_b200.h"

#define GP_MAX_Q 16
#define GP_MAX_RANGES 8        // row ranges of one shard upload
#define GP_RANGE_ROWS 16384    // aim: one range per this many points (only the first range's transfer is exposed)

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void gp_set_error(const char *fmt, ...);
#define GP_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            gp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return GPARML_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)
#define GP_TRY(call)                      \
    do {                                  \
        int r__ = (call);                 \
        if (r__ != GPARML_OK) return r__; \
    } while (0)
#define GP_LAUNCH_CHECK(ctx)                                                          \
    do {                                                                              \
        (ctx)->launches++;                                                            \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            gp_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return GPARML_ERR_CUDA;                                                   \
        }                                                                             \
    } while (0)

// ---------------------------------------------------------------------------
// per-point record layout (written by prep_points, read by the Psi kernels)
//   rec[0 .. 2Q)   : (mu_q, c_q) interleaved;  c = w (Psi2 records) or a (Psi1 records)
//   rec[2Q .. 3Q)  : v_q = alpha_q * S_q * c_q
//   rec[3Q]        : log prefactor  (lc2 = log sf^4 - 1/2 sum log(2 a S + 1),  lc1 = log sf^2 - 1/2 sum log(a S + 1))
//   padded to an even number of doubles so every record is 16-byte aligned.
// ---------------------------------------------------------------------------
__host__ __device__ inline int gp_rec_len(int Q) { return (3 * Q + 2) & ~1; }
// fp32 record: same fields as floats, padded to a multiple of 4 (16-byte records)
__host__ __device__ inline int gp_rec_len_f32(int Q) { return (3 * Q + 4) & ~3; }
// Psi2 records of psi2x_stats (Q <= GP_PSI2X_MAX_Q), centred on the column means of Z (mc = mu - center):
//   recx[0 .. 2Q)    : (-w_q, w_q mc_q) interleaved
//   recx[2Q .. 4Q)   : (v_q, -1 / w_q) interleaved,  v = alpha S w
//   recx[4Q], [4Q+1] : lc2 + sum_q alpha_q S_q  (the exponent is accumulated as sum_q (t^2 + v)(-1/w), which
//                      carries -sum_q alpha_q S_q),  lc2
// pair feature table of embed_psi2m (embed_m.cu): 64-pair chunks, per chunk [feature / 4][pair][4] doubles with
// features (zc_1..zc_Q, zc_1^2..zc_Q^2, 1, 0...) padded to GP_PAIR_R_TILES(Q) tiles of 8
#define GP_PAIR_CHUNK 64
#define GP_PAIR_R_TILES(Q) ((2 * (Q) + 1 + 7) / 8)
#define GP_PSI2M_MIN_Q 5
#define GP_PSI2M_MAX_Q 10
#define GP_PSI2X_MAX_Q 10
#define GP_PSI2X_ROBUST_AS 1024.0   // alpha S above this: the cancellation-free variant of psi2x_stats runs instead
__host__ __device__ inline int gp_recx_len(int Q) { return 4 * Q + 2; }

// packed partial-sum buffer (see DESIGN.md "packed statistics")
struct StatLayout {
    int M, Q, D;
    int64_t P;        // M (M + 1) / 2 upper-triangular pairs (m <= m'), row-major
    int64_t off_p1y;  // (M, D)
    int64_t off_d1z;  // (M, Q, D)
    int64_t off_d1a;  // (Q, M, D)
    int64_t off_s0;   // (P)      Psi2
    int64_t off_tz;   // (Q, P)   sum_n Psi2_n w (mu - zbar)
    int64_t off_ta;   // (Q, P)   sum_n Psi2_n ((w (mu - zbar))^2 + alpha S w)
    int64_t count;
};
// ST_FLAGS: number of shards whose input check failed (unconstrained variance out of range); it is summed by the
// reducer / all-reduce like everything else, so EVERY rank's master step reports GPARML_ERR_RANGE together.
enum { ST_YYT = 0, ST_PSI0 = 1, ST_KL = 2, ST_NLOCAL = 3, ST_FLAGS = 4, ST_HEAD = 6 };

__host__ __device__ inline int64_t gp_pair_index(int M, int a, int b)  // a <= b
{
    return (int64_t)a * M - ((int64_t)a * (a - 1)) / 2 + (b - a);
}

inline StatLayout gp_make_layout(int M, int Q, int D)
{
    StatLayout L;
    L.M = M; L.Q = Q; L.D = D;
    L.P = (int64_t)M * (M + 1) / 2;
    L.off_p1y = ST_HEAD;
    L.off_d1z = L.off_p1y + (int64_t)M * D;
    L.off_d1a = L.off_d1z + (int64_t)M * Q * D;
    L.off_s0 = L.off_d1a + (int64_t)Q * M * D;
    L.off_tz = L.off_s0 + L.P;
    L.off_ta = L.off_tz + (int64_t)Q * L.P;
    L.count = L.off_ta + (int64_t)Q * L.P;
    return L;
}

// globals as the kernels see them
struct GlobalsDev {
    double sf2, beta;
    double log_sf2;
    double alpha[GP_MAX_Q];
    double center[GP_MAX_Q];   // column means of Z: the fp32 maps work on mu - center, z - center
};

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct gparml_ctx {
    int device = 0;
    int M = 0, Q = 0, D = 0;
    int64_t n_total = 0;
    int flags = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    int64_t jitter_events = 0;  // evaluations whose Kmm or Kmm + beta Psi2 needed the 1e-7 jitter retry

    // shard
    int64_t n = 0;         // points in this shard
    int64_t n_cap = 0;     // allocated capacity
    int variance_domain = GPARML_VARIANCE_UNCONSTRAINED;
    bool have_shard = false, have_globals = false, have_prep = false, have_stats = false, have_global_step = false;
    bool have_dir = false;
    double *Y = nullptr;        // (n, D)
    double *x_mu = nullptr;     // (n, Q)
    double *x_s = nullptr;      // (n, Q) uploaded domain
    double *grad_d = nullptr, *grad_latest = nullptr, *grad_new = nullptr, *grad_old = nullptr;  // (2, n, Q)
    double *rec1 = nullptr, *rec2 = nullptr;   // (n, R)
    double *rec2x = nullptr;    // (n, 4Q + 2) records of psi2x_stats (Q <= GP_PSI2X_MAX_Q, fp64 map)
    float *rec2f = nullptr;     // (n, RF) fp32 copy of the Psi2 records (GPARML_FLAG_FP32_MAP only)
    double *s_pos = nullptr;    // (n, Q) positive variance of this evaluation
    double *s_sig = nullptr;    // (n, Q) d softplus / d raw (sigmoid) of this evaluation, 1 if positive domain
    double *gx_mu = nullptr, *gx_s = nullptr;  // (n, Q) positive-domain gradients
    double *psi1 = nullptr;     // (n, M) on demand
    double *d_yyt = nullptr;    // [0] = sum_n y_n . y_n of the shard, [1..] partials of its reduction
    // second stream: the Y upload (needed only by psi1_stats) and the gradient download overlap compute
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_kmm = nullptr;      // Kmm / Kmm^-1 of the current globals are ready (side stream)
    // master step split: the tail (F, hyper-parameter gradients) and its download run on gs_stream next to embed_grads
    cudaStream_t gs_stream = nullptr;
    cudaEvent_t ev_gs_head = nullptr, ev_gs_tail = nullptr;
    bool gs_pending = false;           // gparml_global_step_begin without its _end
    double *glob_host = nullptr;       // pinned staging of [F, grad]
    // gparml_upload_shard sends X_mu / X_S in up to GP_MAX_RANGES row ranges; gparml_statistics consumes them range
    // by range (prep_points + psi2_stats of range k overlap the transfer of range k+1)
    cudaEvent_t ev_x[GP_MAX_RANGES] = {};
    int x_pending = 0;                 // ranges of the last upload the main stream has not been ordered behind yet
    int64_t x_bounds[GP_MAX_RANGES + 1] = {};
    // gparml_embedding_grads_download: timing events (kernels of all ranges; copy of the last range) and the ratio
    // (copy time per point) / (kernel time per point) they gave last time -- the next call sizes its ranges with it
    cudaEvent_t ev_dl[4] = {nullptr, nullptr, nullptr, nullptr};
    double dl_ratio = 0.7;
    cudaEvent_t ev_main = nullptr, ev_y = nullptr, ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    // globals
    double *Z = nullptr;        // (M, Q)
    GlobalsDev h_glob;          // host copy
    GlobalsDev *d_glob = nullptr;
    double step_size = 0.0;
    int2 *pair_idx = nullptr;   // (P) (m, m')
    double *pair_lk = nullptr;  // (P) -1/4 sum_q alpha_q (z_mq - z_m'q)^2
    double2 *pair_g = nullptr;  // (P) (lk, Gs) for embed_grads
    double2 *pair_h = nullptr;  // (P) (lk + log|Gs|, sign(Gs)) for the expanded-basis embed_grads kernel
    double2 *pair_zz = nullptr; // (P, Q) (zbar_q - center_q, (zbar_q - center_q)^2) for embed_grads
    double *pair_zc = nullptr;  // (P, Q rounded up to even) zbar_q - center_q alone: the exponent of embed_psi2x reads only this
    double *pair_r = nullptr;   // blocked pair feature table of embed_psi2m (GP_PSI2M_MIN_Q <= Q <= GP_PSI2M_MAX_Q), zero padded
    double *pair_ra = nullptr;  // the same table times sign(Gs) exp(lk + log|Gs|) per pair: rebuilt after every master step
    bool pair_ra_stale = true;

    // statistics
    StatLayout L;
    double *stats = nullptr;    // packed (L.count)
    double *ws = nullptr;       // workspace for split partial sums
    size_t ws_bytes = 0;
    double *red_ws = nullptr;   // small reduction workspace (4096 doubles)
    int *d_status = nullptr;    // [0] device status word (range / not-PD flags); [1] != 0: some alpha S of this evaluation
                                // exceeds GP_PSI2X_ROBUST_AS (set by prep_points, selects the psi2x_stats variant);
                                // [2] retry pass of the large-M block sweep is live (global_step_large.cu)

    // global step outputs (device)
    double *kmm = nullptr, *kmm_inv = nullptr, *a_inv = nullptr;
    double *g_k = nullptr, *g_1 = nullptr, *g_2 = nullptr;
    double *scratch_x = nullptr, *scratch_w = nullptr;   // (M, M) each, used when M*M does not fit shared memory
    double *c_mat = nullptr;    // (M, D)
    double *psi2_full = nullptr;  // (M, M) expanded Psi2 for the global step
    double *gsl_ws = nullptr;     // workspace of the large-M (multi-kernel) global step, allocated on first use
    double *glob_out = nullptr; // [0]=F, [1..] grad (M*Q + Q + 2), then scalars
    double *named_tmp = nullptr;  // expansion buffer
    size_t named_tmp_count = 0;

    // timing
    bool timing = false;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double phase_ms[5] = {0, 0, 0, 0, 0};
};

int gp_ensure_ws(gparml_ctx *c, size_t bytes);

// kernels' host launchers (each returns GPARML_* codes)
int gp_launch_yyt(gparml_ctx *c, cudaStream_t s);          // -> c->d_yyt[0], on stream s, no host sync
int gp_launch_set_yyt(gparml_ctx *c);                      // stats[ST_YYT] = d_yyt[0]
int gp_launch_embed_grads_range(gparml_ctx *c, int64_t i_begin, int64_t i_end);
int gp_launch_prep(gparml_ctx *c);
int gp_launch_prep_range(gparml_ctx *c, int64_t i0, int64_t i1, double *kl_partials, int max_blocks, int *blocks_used);
int gp_launch_prep_finish(gparml_ctx *c, const double *kl_partials, int blocks);
int gp_psi2_plan_range(gparml_ctx *c, int64_t cnt, int *splits);
int gp_launch_psi2_stats_range(gparml_ctx *c, int64_t i0, int64_t i1, int slice0, int splits);
int gp_launch_psi2_reduce(gparml_ctx *c, int slices);
int gp_launch_pair_table(gparml_ctx *c);
int gp_launch_psi1_stats(gparml_ctx *c);
int gp_launch_psi2_stats(gparml_ctx *c);
int gp_launch_psi1_matrix(gparml_ctx *c);
int gp_launch_global_step(gparml_ctx *c, bool kmm_only, cudaStream_t s);
int gp_launch_global_step_head(gparml_ctx *c, cudaStream_t s, bool *split);
int gp_launch_global_step_tail(gparml_ctx *c, cudaStream_t s);
int gp_launch_embed_grads(gparml_ctx *c);
int gp_launch_expand(gparml_ctx *c, double *dev_out, int which);
int gp_launch_compact(gparml_ctx *c, const double *dev_full_psi2, const double *dev_d2z, const double *dev_d2a);
int gp_launch_stats_add(gparml_ctx *c, const double *src, double scale);
int gp_scg_reduce(gparml_ctx *c, int op, double scale, double *host_out);
int gp_scg_update(gparml_ctx *c, int op, double scale);
int gp_measure_dfma(gparml_ctx *c, double *out);
int gp_init_colsum(gparml_ctx *c, double *out_host);
int gp_init_scatter(gparml_ctx *c, const double *mean_host, double *out_host);
int gp_init_project(gparml_ctx *c, const double *mean_host, const double *W_host);
int gp_init_random(gparml_ctx *c, int mode, uint64_t seed, int64_t row_offset);
int gp_init_kmeans_step(gparml_ctx *c, const double *cent_host, int k, double *out_host);

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double gp_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed tree); result valid in every thread.
// `sh` must hold >= 33 doubles.  blockDim.x must be a multiple of 32.
__device__ __forceinline__ double gp_block_sum(double v, double *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = gp_warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? sh[lane] : 0.0;
        t = gp_warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

// ---- mbarrier + 1-D bulk async copy (TMA unit, SASS UBLKCP) -----------------
__device__ __forceinline__ uint32_t gp_smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void gp_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gp_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gp_fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void gp_fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void gp_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gp_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void gp_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            gp_smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(gp_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void gp_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(gp_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
#endif
