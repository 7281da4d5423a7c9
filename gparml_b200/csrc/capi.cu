// C ABI (include/gparml_b200.h): context management, transfers, phase sequencing.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void gp_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *gparml_last_error(void) { return g_err; }
extern "C" int gparml_abi_version(void) { return 1; }

extern "C" int gparml_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// entry points that take part in the upload -> statistics pipeline (they never touch X_mu / X_S)
#define CHECK_CTX_PIPE(c)                             \
    do {                                              \
        if (!(c)) {                                   \
            gp_set_error("null context");             \
            return GPARML_ERR_ARG;                    \
        }                                             \
        GP_CUDA(cudaSetDevice((c)->device));          \
    } while (0)
// every other entry point first orders the context's stream behind a row-range upload of X_mu / X_S that
// gparml_statistics has not consumed (gparml_upload_shard sends them on the copy stream)
#define CHECK_CTX(c)                                  \
    do {                                              \
        CHECK_CTX_PIPE(c);                            \
        if ((c)->x_pending > 0) {                     \
            GP_CUDA(cudaStreamWaitEvent((c)->stream, (c)->ev_x[(c)->x_pending - 1], 0)); \
            (c)->x_pending = 0;                       \
        }                                             \
    } while (0)

// A gparml_global_step_begin whose _end has not been called yet still reads Z, the globals and the
// statistics on its side stream: entry points that overwrite those finish it first.
static int finish_pending_gs(gparml_ctx *c)
{
    if (c && c->gs_pending) {
        const int r = gparml_global_step_end(c, nullptr, nullptr);
        if (r != GPARML_OK && r != GPARML_ERR_NOT_PD && r != GPARML_ERR_RANGE) return r;
    }
    return GPARML_OK;
}

template <typename T>
static int dev_alloc(T **p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    GP_CUDA(cudaMalloc((void **)p, count * sizeof(T)));
    return GPARML_OK;
}

int gp_ensure_ws(gparml_ctx *c, size_t bytes)
{
    if (bytes <= c->ws_bytes) return GPARML_OK;
    // growing the workspace must not race with kernels still using the old one
    GP_CUDA(cudaStreamSynchronize(c->stream));
    if (c->ws) { cudaFree(c->ws); c->ws = nullptr; c->ws_bytes = 0; }
    size_t want = bytes + bytes / 4 + 4096;
    GP_CUDA(cudaMalloc((void **)&c->ws, want));
    c->ws_bytes = want;
    return GPARML_OK;
}

extern "C" int gparml_create(gparml_ctx **out, int device, int M, int Q, int D, int64_t n_total, int flags)
{
    if (!out) { gp_set_error("gparml_create: out is null"); return GPARML_ERR_ARG; }
    *out = nullptr;
    int ndev = gparml_device_count();
    if (ndev <= 0) {
        gp_set_error("gparml_create: no CUDA device visible -- gparml_b200 has no CPU path");
        return GPARML_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) { gp_set_error("gparml_create: device %d out of range (0..%d)", device, ndev - 1); return GPARML_ERR_ARG; }
    if (M < 1 || Q < 1 || Q > GP_MAX_Q || D < 1 || n_total < 0) {
        gp_set_error("gparml_create: unsupported shape M=%d Q=%d (1..%d) D=%d N=%lld", M, Q, GP_MAX_Q, D, (long long)n_total);
        return GPARML_ERR_ARG;
    }
    GP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        gp_set_error("gparml_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return GPARML_ERR_NO_DEVICE;
    }
    gparml_ctx *c = new gparml_ctx();
    c->device = device; c->M = M; c->Q = Q; c->D = D; c->n_total = n_total; c->flags = flags;
    c->sm_count = prop.multiProcessorCount;
    c->L = gp_make_layout(M, Q, D);
    int r = GPARML_OK;
    auto fail = [&](int code) { gparml_destroy(c); return code; };
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { gp_set_error("stream create failed"); return fail(GPARML_ERR_CUDA); }
    c->stream = c->own_stream;
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_kmm, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_y, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->gs_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_gs_head, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_gs_tail, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost((void **)&c->glob_host, ((size_t)M * Q + Q + 16) * sizeof(double)) != cudaSuccess) { gp_set_error("copy stream create failed"); return fail(GPARML_ERR_CUDA); }
    for (int i = 0; i < 8; ++i)
        if (cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming) != cudaSuccess) { gp_set_error("event create failed"); return fail(GPARML_ERR_CUDA); }
    for (int i = 0; i < 4; ++i)
        if (cudaEventCreate(&c->ev_dl[i]) != cudaSuccess) { gp_set_error("event create failed"); return fail(GPARML_ERR_CUDA); }
    for (int i = 0; i < GP_MAX_RANGES; ++i)
        if (cudaEventCreateWithFlags(&c->ev_x[i], cudaEventDisableTiming) != cudaSuccess) { gp_set_error("event create failed"); return fail(GPARML_ERR_CUDA); }
    const size_t MM = (size_t)M * M;
#define A_(ptr, count) if ((r = dev_alloc(&(ptr), (count))) != GPARML_OK) return fail(r)
    A_(c->Z, (size_t)M * Q);
    A_(c->d_glob, 1);
    A_(c->pair_idx, (size_t)c->L.P);
    A_(c->pair_lk, (size_t)c->L.P);
    A_(c->pair_g, (size_t)c->L.P);
    A_(c->pair_zz, (size_t)c->L.P * Q);
    A_(c->pair_zc, (size_t)c->L.P * ((Q + 1) & ~1));
    {   // padded to whole chunks of GP_PAIR_CHUNK pairs and zeroed: embed_psi2m copies whole chunks
        const size_t chunks = ((size_t)c->L.P + GP_PAIR_CHUNK - 1) / GP_PAIR_CHUNK;
        A_(c->pair_h, chunks * GP_PAIR_CHUNK);
        if (cudaMemsetAsync(c->pair_h, 0, chunks * GP_PAIR_CHUNK * sizeof(double2), c->stream) != cudaSuccess) { gp_set_error("memset failed"); return fail(GPARML_ERR_CUDA); }
        if (Q >= GP_PSI2M_MIN_Q && Q <= GP_PSI2M_MAX_Q) {
            const size_t count = chunks * GP_PAIR_R_TILES(Q) * 2 * GP_PAIR_CHUNK * 4;
            A_(c->pair_r, count);
            A_(c->pair_ra, count);
            if (cudaMemsetAsync(c->pair_r, 0, count * sizeof(double), c->stream) != cudaSuccess ||
                cudaMemsetAsync(c->pair_ra, 0, count * sizeof(double), c->stream) != cudaSuccess) { gp_set_error("memset failed"); return fail(GPARML_ERR_CUDA); }
        }
    }
    A_(c->stats, (size_t)c->L.count);
    A_(c->red_ws, 4096);
    A_(c->d_yyt, 1100);
    A_(c->d_status, 4);
    A_(c->kmm, MM); A_(c->kmm_inv, MM); A_(c->a_inv, MM);
    A_(c->g_k, MM); A_(c->g_2, MM); A_(c->g_1, (size_t)M * D); A_(c->c_mat, (size_t)M * D);
    A_(c->scratch_x, MM); A_(c->scratch_w, MM); A_(c->psi2_full, MM);
    A_(c->glob_out, (size_t)M * Q + Q + 16);
#undef A_
    if (cudaMemsetAsync(c->stats, 0, c->L.count * sizeof(double), c->stream) != cudaSuccess ||
        cudaMemsetAsync(c->d_status, 0, 4 * sizeof(int), c->stream) != cudaSuccess) { gp_set_error("memset failed"); return fail(GPARML_ERR_CUDA); }
    for (int i = 0; i < 8; ++i)
        if (cudaEventCreate(&c->ev[i]) != cudaSuccess) { gp_set_error("event create failed"); return fail(GPARML_ERR_CUDA); }
    *out = c;
    return GPARML_OK;
}

extern "C" int gparml_destroy(gparml_ctx *c)
{
    if (!c) return GPARML_OK;
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    void *ptrs[] = {c->Y, c->x_mu, c->x_s, c->grad_d, c->grad_latest, c->grad_new, c->grad_old, c->rec1, c->rec2, c->s_pos, c->s_sig,
                    c->gx_mu, c->gx_s, c->psi1, c->Z, c->d_glob, c->pair_idx, c->pair_lk, c->pair_g, c->pair_zz, c->pair_zc, c->pair_h, c->stats, c->ws, c->red_ws,
                    c->d_status, c->kmm, c->kmm_inv, c->a_inv, c->g_k, c->g_1, c->g_2, c->scratch_x, c->scratch_w, c->c_mat,
                    c->glob_out, c->named_tmp, c->psi2_full, c->gsl_ws, c->rec2f, c->d_yyt, c->rec2x, c->pair_r, c->pair_ra};
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->gs_stream) cudaStreamSynchronize(c->gs_stream);
    for (void *p : ptrs) if (p) cudaFree(p);
    if (c->glob_host) cudaFreeHost(c->glob_host);
    if (c->ev_gs_head) cudaEventDestroy(c->ev_gs_head);
    if (c->ev_gs_tail) cudaEventDestroy(c->ev_gs_tail);
    if (c->gs_stream) cudaStreamDestroy(c->gs_stream);
    for (int i = 0; i < 8; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 8; ++i) if (c->ev_chunk[i]) cudaEventDestroy(c->ev_chunk[i]);
    for (int i = 0; i < 4; ++i) if (c->ev_dl[i]) cudaEventDestroy(c->ev_dl[i]);
    for (int i = 0; i < GP_MAX_RANGES; ++i) if (c->ev_x[i]) cudaEventDestroy(c->ev_x[i]);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_y) cudaEventDestroy(c->ev_y);
    if (c->ev_kmm) cudaEventDestroy(c->ev_kmm);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return GPARML_OK;
}

extern "C" int gparml_set_stream(gparml_ctx *c, void *s)
{
    CHECK_CTX(c);
    GP_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = (cudaStream_t)s;          // taken literally: NULL is the legacy default stream
    return GPARML_OK;
}

extern "C" int gparml_use_own_stream(gparml_ctx *c)
{
    CHECK_CTX(c);
    GP_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = c->own_stream;
    return GPARML_OK;
}

extern "C" int gparml_synchronize(gparml_ctx *c)
{
    CHECK_CTX(c);
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}

extern "C" int gparml_set_n_total(gparml_ctx *c, int64_t n_total)
{
    CHECK_CTX(c);
    c->n_total = n_total;
    return GPARML_OK;
}

extern "C" int64_t gparml_n_local(const gparml_ctx *c) { return c ? c->n : -1; }
extern "C" int64_t gparml_stats_count(const gparml_ctx *c) { return c ? c->L.count : -1; }
extern "C" int64_t gparml_launch_count(const gparml_ctx *c) { return c ? c->launches : -1; }
extern "C" int64_t gparml_jitter_events(const gparml_ctx *c) { return c ? c->jitter_events : -1; }

static int ensure_shard_capacity(gparml_ctx *c, int64_t n)
{
    if (n <= c->n_cap) return GPARML_OK;
    GP_CUDA(cudaStreamSynchronize(c->stream));
    const size_t nq = (size_t)n * c->Q, R = gp_rec_len(c->Q);
    GP_TRY(dev_alloc(&c->Y, (size_t)n * c->D));
    GP_TRY(dev_alloc(&c->x_mu, nq));
    GP_TRY(dev_alloc(&c->x_s, nq));
    GP_TRY(dev_alloc(&c->grad_d, 2 * nq));
    GP_TRY(dev_alloc(&c->grad_latest, 2 * nq));
    GP_TRY(dev_alloc(&c->grad_new, 2 * nq));
    GP_TRY(dev_alloc(&c->grad_old, 2 * nq));
    GP_TRY(dev_alloc(&c->rec1, (size_t)n * R));
    GP_TRY(dev_alloc(&c->rec2, (size_t)n * R));
    if (c->flags & GPARML_FLAG_FP32_MAP) GP_TRY(dev_alloc(&c->rec2f, (size_t)n * gp_rec_len_f32(c->Q)));
    else if (c->Q <= GP_PSI2X_MAX_Q) GP_TRY(dev_alloc(&c->rec2x, (size_t)n * gp_recx_len(c->Q)));
    GP_TRY(dev_alloc(&c->s_pos, nq));
    GP_TRY(dev_alloc(&c->s_sig, nq));
    GP_TRY(dev_alloc(&c->gx_mu, nq));
    GP_TRY(dev_alloc(&c->gx_s, nq));
    if (c->psi1) { cudaFree(c->psi1); c->psi1 = nullptr; }
    c->n_cap = n;
    return GPARML_OK;
}

extern "C" int gparml_upload_shard(gparml_ctx *c, const double *Y, const double *X_mu, const double *X_S, int64_t n, int domain)
{
    CHECK_CTX_PIPE(c);
    if (n < 0 || (n > 0 && (!Y || !X_mu || !X_S))) { gp_set_error("upload_shard: null array or negative n"); return GPARML_ERR_ARG; }
    if (domain != GPARML_VARIANCE_UNCONSTRAINED && domain != GPARML_VARIANCE_POSITIVE) { gp_set_error("upload_shard: bad variance domain %d", domain); return GPARML_ERR_ARG; }
    GP_TRY(ensure_shard_capacity(c, n));
    const size_t nq = (size_t)n * c->Q;
    // Everything travels on the copy stream (ordered behind all work already queued on the main stream,
    // which may still read the old arrays): X_mu / X_S first, in up to GP_MAX_RANGES row ranges with an event each --
    // gparml_statistics runs prep_points + psi2_stats of range k while range k + 1 is still arriving --,
    // then Y, which only psi1_stats / the Psi1 part of embed_grads need.
    GP_CUDA(cudaEventRecord(c->ev_main, c->stream));
    GP_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
    // Row ranges that double in size: only the first (small) range's transfer is exposed, and the kernels of range k
    // (twice the work of range k - 1) cover the transfer of range k + 1 as long as the link moves a point faster than
    // half the time prep_points + psi2_stats need for it; few, large ranges keep the per-launch wave quantisation of
    // psi2_stats small (8 equal ranges cost it 4 % at c3).
    int ranges = 1;
    while (ranges < GP_MAX_RANGES && (n / GP_RANGE_ROWS + 1) >> ranges) ++ranges;      // floor(log2(n / 16384 + 1)), >= 1
    if (n / GP_RANGE_ROWS < 1 || (c->flags & GPARML_FLAG_FP32_MAP)) ranges = 1;
    c->x_bounds[0] = 0;
    for (int k = 1; k <= ranges; ++k) {
        int64_t b = (int64_t)((double)n * (double)((1 << k) - 1) / (double)((1 << ranges) - 1));
        b = (b + 127) / 128 * 128;
        c->x_bounds[k] = (k == ranges || b > n) ? n : b;
    }
    if (n > 0) {
        for (int k = 0; k < ranges; ++k) {
            const size_t off = (size_t)c->x_bounds[k] * c->Q, cnt = (size_t)(c->x_bounds[k + 1] - c->x_bounds[k]) * c->Q;
            GP_CUDA(cudaMemcpyAsync(c->x_mu + off, X_mu + off, cnt * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
            GP_CUDA(cudaMemcpyAsync(c->x_s + off, X_S + off, cnt * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
            GP_CUDA(cudaEventRecord(c->ev_x[k], c->copy_stream));
        }
        GP_CUDA(cudaMemcpyAsync(c->Y, Y, (size_t)n * c->D * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
    } else {
        GP_CUDA(cudaEventRecord(c->ev_x[0], c->copy_stream));
    }
    c->x_pending = ranges;
    if (n != c->n) {
        c->have_dir = false;
        if (c->psi1) { cudaFree(c->psi1); c->psi1 = nullptr; }
    }
    c->n = n;
    c->variance_domain = domain;
    c->have_shard = true;
    c->have_prep = c->have_stats = c->have_global_step = false;
    GP_TRY(gp_launch_yyt(c, c->copy_stream));
    GP_CUDA(cudaEventRecord(c->ev_y, c->copy_stream));
    // pageable host arrays: the async copies above are staged before they return, so the caller may
    // reuse its buffers; pinned arrays must stay valid until the next synchronising call.
    return GPARML_OK;
}

// ---- one-off initialisation on the device (init.cu) ----------------------------------------
extern "C" int gparml_upload_outputs(gparml_ctx *c, const double *Y, int64_t n)
{
    CHECK_CTX(c);
    if (n < 0 || (n > 0 && !Y)) { gp_set_error("upload_outputs: null array or negative n"); return GPARML_ERR_ARG; }
    GP_TRY(ensure_shard_capacity(c, n));
    const size_t nq = (size_t)n * c->Q;
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_y, 0));   // an earlier upload of Y on the copy stream
    if (n > 0) {
        GP_CUDA(cudaMemcpyAsync(c->Y, Y, (size_t)n * c->D * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        GP_CUDA(cudaMemsetAsync(c->x_mu, 0, nq * sizeof(double), c->stream));
        GP_CUDA(cudaMemsetAsync(c->x_s, 0, nq * sizeof(double), c->stream));
    }
    if (n != c->n) {
        c->have_dir = false;
        if (c->psi1) { cudaFree(c->psi1); c->psi1 = nullptr; }
    }
    c->n = n;
    c->variance_domain = GPARML_VARIANCE_UNCONSTRAINED;
    c->have_shard = true;
    c->have_prep = c->have_stats = c->have_global_step = false;
    GP_TRY(gp_launch_yyt(c, c->stream));
    GP_CUDA(cudaEventRecord(c->ev_y, c->stream));
    return GPARML_OK;
}

#define CHECK_SHARD(c, what)                                                              \
    do {                                                                                  \
        if (!(c)->have_shard) { gp_set_error(what ": no shard uploaded"); return GPARML_ERR_STATE; } \
    } while (0)

extern "C" int gparml_init_column_sums(gparml_ctx *c, double *out)
{
    CHECK_CTX(c);
    CHECK_SHARD(c, "init_column_sums");
    if (!out) { gp_set_error("init_column_sums: null output"); return GPARML_ERR_ARG; }
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_y, 0));
    return gp_init_colsum(c, out);
}

extern "C" int gparml_init_scatter(gparml_ctx *c, const double *mean, double *out)
{
    CHECK_CTX(c);
    CHECK_SHARD(c, "init_scatter");
    if (!mean || !out) { gp_set_error("init_scatter: null array"); return GPARML_ERR_ARG; }
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_y, 0));
    return gp_init_scatter(c, mean, out);
}

extern "C" int gparml_init_project(gparml_ctx *c, const double *mean, const double *W)
{
    CHECK_CTX(c);
    CHECK_SHARD(c, "init_project");
    if (!mean || !W) { gp_set_error("init_project: null array"); return GPARML_ERR_ARG; }
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_y, 0));
    c->have_prep = c->have_stats = c->have_global_step = false;
    GP_TRY(gp_init_project(c, mean, W));
    GP_CUDA(cudaStreamSynchronize(c->stream));     // W / mean may be pageable temporaries of the caller
    return GPARML_OK;
}

extern "C" int gparml_init_random(gparml_ctx *c, int what, uint64_t seed, int64_t row_offset)
{
    CHECK_CTX(c);
    CHECK_SHARD(c, "init_random");
    if (what != 0 && what != 1) { gp_set_error("init_random: what must be 0 (variances) or 1 (means)"); return GPARML_ERR_ARG; }
    if (row_offset < 0) { gp_set_error("init_random: negative row offset"); return GPARML_ERR_ARG; }
    c->have_prep = c->have_stats = c->have_global_step = false;
    if (what == 0) c->variance_domain = GPARML_VARIANCE_UNCONSTRAINED;
    return gp_init_random(c, what, seed, row_offset);
}

extern "C" int gparml_kmeans_step(gparml_ctx *c, const double *centroids, int k, double *out)
{
    CHECK_CTX(c);
    CHECK_SHARD(c, "kmeans_step");
    if (!centroids || !out || k < 1) { gp_set_error("kmeans_step: null array or k < 1"); return GPARML_ERR_ARG; }
    return gp_init_kmeans_step(c, centroids, k, out);
}

// make the main stream wait for the Y upload (and its sum of squares)
static int wait_y(gparml_ctx *c)
{
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_y, 0));
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_kmm, 0));     // and for Kmm / Kmm^-1 of the current globals
    return GPARML_OK;
}

extern "C" int gparml_set_globals(gparml_ctx *c, const double *Z, double sf2, const double *alpha, double beta)
{
    CHECK_CTX_PIPE(c);
    GP_TRY(finish_pending_gs(c));
    if (!Z || !alpha) { gp_set_error("set_globals: null array"); return GPARML_ERR_ARG; }
    if (!(sf2 > 0.0) || !(beta > 0.0)) { gp_set_error("set_globals: sf2 and beta must be positive (kernels.py:57 assert)"); return GPARML_ERR_ARG; }
    memset(&c->h_glob, 0, sizeof(c->h_glob));
    c->h_glob.sf2 = sf2;
    c->h_glob.beta = beta;
    c->h_glob.log_sf2 = log(sf2);
    for (int q = 0; q < c->Q; ++q) {
        if (!(alpha[q] >= 0.0)) { gp_set_error("set_globals: alpha[%d] negative (kernel_exp.py:30 assert)", q); return GPARML_ERR_ARG; }
        // alpha_q = 0 is legal in the reference (kernel_exp.py:30 asserts >= 0) and switches dimension q off.  The
        // packed buffer keeps the alpha-derivative sums scaled by alpha^2 (TA, rows 1+Q.. of K1), which is 0 * inf
        // there; 2^-400 instead of 0 gives bit-identical kernel values (exp(-2^-400 x) rounds to 1), alpha^2 = 2^-800
        // is an exact power of two far above the denormals, and the derivative comes out as the limit alpha -> 0+.
        c->h_glob.alpha[q] = alpha[q] > 0.0 ? alpha[q] : 0x1p-400;
    }
    // centre of the inducing inputs: the expanded basis of embed_grads (embed.cu) and the fp32 maps work on
    // mu - center, zbar - center, so their cancellation scales with the spread of Z, not its offset from the origin
    for (int q = 0; q < c->Q; ++q) {
        double s = 0.0;
        for (int m = 0; m < c->M; ++m) s += Z[(size_t)m * c->Q + q];
        c->h_glob.center[q] = s / c->M;
    }
    // pageable host memory: the async copies below stage synchronously, so Z/alpha may be reused by the caller on return
    GP_CUDA(cudaMemcpyAsync(c->Z, Z, (size_t)c->M * c->Q * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(cudaMemcpyAsync(c->d_glob, &c->h_glob, sizeof(GlobalsDev), cudaMemcpyHostToDevice, c->stream));
    GP_TRY(gp_launch_pair_table(c));
    // Kmm, Kmm^-1, log det Kmm depend on the hyper-parameters only (the reference's cache(),
    // local_MapReduce.py:383-394): computed now on the side stream, concurrently with the
    // statistics map, and joined by global_step / update_global_statistics through ev_kmm.
    GP_CUDA(cudaEventRecord(c->ev_main, c->stream));
    GP_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
    GP_TRY(gp_launch_global_step(c, true, c->copy_stream));
    GP_CUDA(cudaEventRecord(c->ev_kmm, c->copy_stream));
    c->have_globals = true;
    c->have_prep = c->have_stats = c->have_global_step = false;
    return GPARML_OK;
}

extern "C" int gparml_set_step(gparml_ctx *c, double step)
{
    CHECK_CTX_PIPE(c);
    c->step_size = step;
    c->have_prep = false;
    return GPARML_OK;
}

static int check_status(gparml_ctx *c, bool sync_needed)
{
    int st = 0;
    GP_CUDA(cudaMemcpyAsync(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    (void)sync_needed;
    if (st) {
        GP_CUDA(cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream));
        if (st & 24) c->jitter_events++;        // a factorisation needed the reference's 1e-7 jitter: not an error
        if (st & 4) { gp_set_error("unconstrained variance outside (-36.04, 36.04) (supporting_functions.py:154 assert)"); return GPARML_ERR_RANGE; }
        if (st & 1) { gp_set_error("Kmm is not positive definite, also with 1e-7 jitter (pivot <= 0)"); return GPARML_ERR_NOT_PD; }
        if (st & 2) { gp_set_error("Kmm + beta*Psi2 is not positive definite, also with 1e-7 jitter (pivot <= 0)"); return GPARML_ERR_NOT_PD; }
    }
    return GPARML_OK;
}

static int record(gparml_ctx *c, int i)
{
    if (c->timing) GP_CUDA(cudaEventRecord(c->ev[i], c->stream));
    return GPARML_OK;
}

static int prep_if_needed(gparml_ctx *c)
{
    if (!c->have_shard || !c->have_globals) { gp_set_error("upload_shard and set_globals must precede this call"); return GPARML_ERR_STATE; }
    if (!c->have_prep) {
        GP_TRY(gp_launch_prep(c));
        c->have_prep = true;
    }
    return GPARML_OK;
}

// The statistics map without any host synchronisation: everything is queued on the context's streams and the
// call returns.  Device-side failures (unconstrained variance out of range) stay in the device status word, which
// gparml_status / gparml_global_step_end read back -- so a host thread that drives several GPUs can launch all of
// its shards before it waits for any of them (the reference forks one mapper per shard, local_MapReduce.py:134).
extern "C" int gparml_statistics_launch(gparml_ctx *c)
{
    CHECK_CTX_PIPE(c);
    GP_TRY(finish_pending_gs(c));
    if (!c->have_shard || !c->have_globals) { gp_set_error("statistics: upload_shard and set_globals first"); return GPARML_ERR_STATE; }
    GP_TRY(record(c, 0));
    if (c->x_pending > 1 && c->n > 0) {
        // pipelined with the upload: range k is prepared and reduced while range k + 1 is still arriving
        const int ranges = c->x_pending;
        c->x_pending = 0;
        int splits[GP_MAX_RANGES], slice = 0, blocks = 0, total = 0;
        for (int k = 0; k < ranges; ++k) {
            splits[k] = 1;
            if (c->x_bounds[k + 1] > c->x_bounds[k]) GP_TRY(gp_psi2_plan_range(c, c->x_bounds[k + 1] - c->x_bounds[k], &splits[k]));
            total += splits[k];
        }
        GP_TRY(gp_ensure_ws(c, (size_t)total * (1 + 2 * c->Q) * c->L.P * sizeof(double)));
        for (int k = 0; k < ranges; ++k) {
            GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_x[k], 0));
            int b = 0;
            // KL block partials of all ranges share red_ws (2 x 2048 doubles): each range gets its share of the slots
            int kl_blocks = (int)(2000.0 * (double)(c->x_bounds[k + 1] - c->x_bounds[k]) / (double)c->n);
            if (kl_blocks < 4) kl_blocks = 4;
            GP_TRY(gp_launch_prep_range(c, c->x_bounds[k], c->x_bounds[k + 1], c->red_ws + 2 * blocks, kl_blocks, &b));
            blocks += b;
            if (k == 0) GP_TRY(record(c, 1));
            if (c->x_bounds[k + 1] > c->x_bounds[k]) {
                GP_TRY(gp_launch_psi2_stats_range(c, c->x_bounds[k], c->x_bounds[k + 1], slice, splits[k]));
                slice += splits[k];
            }
        }
        GP_TRY(gp_launch_prep_finish(c, c->red_ws, blocks));
        GP_TRY(gp_launch_psi2_reduce(c, slice));
        c->have_prep = true;
    } else {
        if (c->x_pending > 0) {
            GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_x[c->x_pending - 1], 0));
            c->x_pending = 0;
        }
        GP_TRY(gp_launch_prep(c));          // always: it also rewrites the header of the packed buffer
        c->have_prep = true;
        GP_TRY(record(c, 1));
        GP_TRY(gp_launch_psi2_stats(c));    // needs only X: runs while Y may still be arriving on the copy stream
    }
    GP_TRY(record(c, 2));
    GP_TRY(wait_y(c));
    GP_TRY(gp_launch_set_yyt(c));
    GP_TRY(gp_launch_psi1_stats(c));
    GP_TRY(record(c, 3));
    c->have_stats = true;
    c->have_global_step = false;
    return GPARML_OK;
}

// Waits for the context's stream and reports (and clears) the device status word.
extern "C" int gparml_status(gparml_ctx *c)
{
    CHECK_CTX(c);
    return check_status(c, true);
}

// Blocking form: the statistics map, then -- where a device-side check can fail (unconstrained variances,
// supporting_functions.py:154) -- the status word, so that the assert surfaces in this call like in the reference.
extern "C" int gparml_statistics(gparml_ctx *c)
{
    GP_TRY(gparml_statistics_launch(c));
    if (!(c->flags & GPARML_FLAG_FIXED_EMBEDDINGS) && c->variance_domain == GPARML_VARIANCE_UNCONSTRAINED)
        GP_TRY(check_status(c, true));
    return GPARML_OK;
}

extern "C" int gparml_stats_device_ptr(gparml_ctx *c, void **p)
{
    CHECK_CTX(c);
    if (!p) { gp_set_error("null out pointer"); return GPARML_ERR_ARG; }
    *p = c->stats;
    return GPARML_OK;
}

extern "C" int gparml_stats_add(gparml_ctx *c, const void *other, double scale)
{
    CHECK_CTX(c);
    GP_TRY(finish_pending_gs(c));
    if (!other) { gp_set_error("stats_add: null pointer"); return GPARML_ERR_ARG; }
    GP_TRY(gp_launch_stats_add(c, (const double *)other, scale));
    c->have_global_step = false;
    return GPARML_OK;
}

extern "C" int gparml_stats_copy(gparml_ctx *c, const void *other)
{
    CHECK_CTX(c);
    GP_TRY(finish_pending_gs(c));
    if (!other) { gp_set_error("stats_copy: null pointer"); return GPARML_ERR_ARG; }
    GP_CUDA(cudaMemcpyAsync(c->stats, other, (size_t)c->L.count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->have_stats = true;
    c->have_global_step = false;
    return GPARML_OK;
}

extern "C" int gparml_update_global_statistics(gparml_ctx *c)
{
    CHECK_CTX(c);
    GP_TRY(finish_pending_gs(c));
    if (!c->have_globals) { gp_set_error("update_global_statistics: set_globals first"); return GPARML_ERR_STATE; }
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_kmm, 0));     // launched by set_globals on the side stream
    return check_status(c, true);
}

// The master step in two halves (include/gparml_b200.h).  _begin launches everything and returns at once:
// the part embed_grads depends on (A^-1, dF/dPsi1Y, dF/dPsi2 -> pair tables) on the context's stream, the
// tail (F, gradients of Z / sf2 / alpha / beta) and its download on gs_stream, where it overlaps the
// embeddings map.  _end waits for the tail and hands out F and the gradient.
extern "C" int gparml_global_step_begin(gparml_ctx *c)
{
    CHECK_CTX(c);
    if (!c->have_globals) { gp_set_error("global_step: set_globals first"); return GPARML_ERR_STATE; }
    if (c->gs_pending) GP_TRY(gparml_global_step_end(c, nullptr, nullptr));
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_kmm, 0));     // Kmm^-1 from the side stream
    GP_TRY(record(c, 4));
    bool split = false;
    GP_TRY(gp_launch_global_step_head(c, c->stream, &split));
    GP_TRY(record(c, 5));
    GP_CUDA(cudaEventRecord(c->ev_gs_head, c->stream));
    GP_CUDA(cudaStreamWaitEvent(c->gs_stream, c->ev_gs_head, 0));
    if (split) GP_TRY(gp_launch_global_step_tail(c, c->gs_stream));
    const size_t ng = (size_t)c->M * c->Q + c->Q + 2;
    GP_CUDA(cudaMemcpyAsync(c->glob_host, c->glob_out, (1 + ng) * sizeof(double), cudaMemcpyDeviceToHost, c->gs_stream));
    GP_CUDA(cudaMemcpyAsync(c->glob_host + 1 + ng, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->gs_stream));
    GP_CUDA(cudaEventRecord(c->ev_gs_tail, c->gs_stream));
    c->have_global_step = true;
    c->gs_pending = true;
    return GPARML_OK;
}

extern "C" int gparml_global_step_end(gparml_ctx *c, double *F, double *grad)
{
    CHECK_CTX(c);
    if (!c->gs_pending) { gp_set_error("global_step_end: no global_step_begin pending"); return GPARML_ERR_STATE; }
    c->gs_pending = false;
    GP_CUDA(cudaEventSynchronize(c->ev_gs_tail));
    // later work on the context's stream (next set_globals / statistics) must not overtake the tail's reads
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_gs_tail, 0));
    const size_t ng = (size_t)c->M * c->Q + c->Q + 2;
    int st = 0;
    memcpy(&st, c->glob_host + 1 + ng, sizeof(int));           // device status word as of the end of the head
    if (st) {
        GP_CUDA(cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream));
        if (st & 24) c->jitter_events++;
        if (st & 4) { gp_set_error("unconstrained variance outside (-36.04, 36.04) (supporting_functions.py:154 assert)"); return GPARML_ERR_RANGE; }
        if (st & 1) { gp_set_error("Kmm is not positive definite, also with 1e-7 jitter (pivot <= 0)"); return GPARML_ERR_NOT_PD; }
        if (st & 2) { gp_set_error("Kmm + beta*Psi2 is not positive definite, also with 1e-7 jitter (pivot <= 0)"); return GPARML_ERR_NOT_PD; }
    }
    if (F) *F = c->glob_host[0];
    if (grad) memcpy(grad, c->glob_host + 1, ng * sizeof(double));
    return GPARML_OK;
}

extern "C" int gparml_global_step(gparml_ctx *c, double *F, double *grad)
{
    GP_TRY(gparml_global_step_begin(c));
    return gparml_global_step_end(c, F, grad);
}

extern "C" int gparml_embedding_grads(gparml_ctx *c)
{
    CHECK_CTX(c);
    if (c->flags & GPARML_FLAG_FIXED_EMBEDDINGS) { gp_set_error("embedding_grads: context has fixed embeddings"); return GPARML_ERR_STATE; }
    if (!c->have_global_step) { gp_set_error("embedding_grads: global_step first"); return GPARML_ERR_STATE; }
    GP_TRY(prep_if_needed(c));
    GP_TRY(wait_y(c));
    GP_TRY(record(c, 6));
    GP_TRY(gp_launch_embed_grads(c));
    GP_TRY(record(c, 7));
    return GPARML_OK;
}

// embedding_grads + download of GRAD_LATEST in one call, in `chunks` point ranges: the
// device-to-host copy of range k (copy stream) overlaps the kernels of range k + 1.
extern "C" int gparml_embedding_grads_download(gparml_ctx *c, double *host_grad_latest, int chunks)
{
    CHECK_CTX(c);
    if (c->flags & GPARML_FLAG_FIXED_EMBEDDINGS) { gp_set_error("embedding_grads: context has fixed embeddings"); return GPARML_ERR_STATE; }
    if (!c->have_global_step) { gp_set_error("embedding_grads: global_step first"); return GPARML_ERR_STATE; }
    if (!host_grad_latest) { gp_set_error("embedding_grads_download: null destination"); return GPARML_ERR_ARG; }
    if (chunks < 1) chunks = 1;
    if (chunks > 8) chunks = 8;
    GP_TRY(prep_if_needed(c));
    GP_TRY(wait_y(c));
    GP_TRY(record(c, 6));
    const int64_t n = c->n, Q = c->Q;
    // Geometric ranges (each r x the one before): the copy of range k hides behind the kernels of range k + 1 as long
    // as the link moves a point's gradient faster than r x the time the kernels need for it, and only the LAST range's
    // copy is exposed.  r is measured: the previous call timed its kernels and the (unobstructed) copy of its last
    // range; the first call assumes 0.7 (the value 8 ranks sharing one host link see at c3).  One rank alone on the
    // link gets r = 0.17: ranges of 83 / 14 / 2.4 % of the points instead of 46 / 32 / 22 %.
    int64_t bounds[9];
    {
        const double r = c->dl_ratio;
        double w[8], tot = 0.0, acc = 0.0;
        for (int k = 0; k < chunks; ++k) { w[k] = pow(r, k); tot += w[k]; }
        bounds[0] = 0;
        for (int k = 0; k < chunks; ++k) {
            acc += w[k];
            int64_t b = (int64_t)((double)n * acc / tot);
            b = (b + 255) / 256 * 256;                   // whole point tiles of the embeddings kernels
            if (b > n || k == chunks - 1) b = n;
            if (b < bounds[k]) b = bounds[k];
            bounds[k + 1] = b;
        }
    }
    int last = -1;
    for (int k = 0; k < chunks; ++k) if (bounds[k + 1] > bounds[k]) last = k;
    GP_CUDA(cudaEventRecord(c->ev_dl[0], c->stream));
    for (int k = 0; k < chunks; ++k) {
        const int64_t lo = bounds[k], hi = bounds[k + 1];
        if (hi <= lo) continue;
        GP_TRY(gp_launch_embed_grads_range(c, lo, hi));
        GP_CUDA(cudaEventRecord(c->ev_chunk[k], c->stream));
        if (k == last) GP_CUDA(cudaEventRecord(c->ev_dl[1], c->stream));
        GP_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_chunk[k], 0));
        const size_t bytes = (size_t)(hi - lo) * Q * sizeof(double);
        if (k == last) GP_CUDA(cudaEventRecord(c->ev_dl[2], c->copy_stream));
        GP_CUDA(cudaMemcpyAsync(host_grad_latest + lo * Q, c->grad_latest + lo * Q, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
        GP_CUDA(cudaMemcpyAsync(host_grad_latest + (n + lo) * Q, c->grad_latest + (n + lo) * Q, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
        if (k == last) GP_CUDA(cudaEventRecord(c->ev_dl[3], c->copy_stream));
    }
    GP_TRY(record(c, 7));
    GP_CUDA(cudaStreamSynchronize(c->copy_stream));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    if (last >= 0 && chunks > 1 && n >= 4096) {
        float t_kern = 0.f, t_copy = 0.f;
        if (cudaEventElapsedTime(&t_kern, c->ev_dl[0], c->ev_dl[1]) == cudaSuccess &&
            cudaEventElapsedTime(&t_copy, c->ev_dl[2], c->ev_dl[3]) == cudaSuccess && t_kern > 0.f && t_copy > 0.f) {
            const double n_last = (double)(bounds[last + 1] - bounds[last]);
            double r = 1.15 * ((double)t_copy / n_last) / ((double)t_kern / (double)n);      // 15 % margin
            r = r < 0.1 ? 0.1 : (r > 0.9 ? 0.9 : r);
            c->dl_ratio = r;
        }
    }
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// generic arrays
// ---------------------------------------------------------------------------
static int resolve(gparml_ctx *c, int id, double **ptr, int64_t *count, bool for_write)
{
    const int64_t nq = c->n * c->Q, MM = (int64_t)c->M * c->M;
    switch (id) {
        case GPARML_A_X_MU: *ptr = c->x_mu; *count = nq; break;
        case GPARML_A_X_S: *ptr = c->x_s; *count = nq; break;
        case GPARML_A_GRAD_D: *ptr = c->grad_d; *count = 2 * nq; break;
        case GPARML_A_GRAD_LATEST: *ptr = c->grad_latest; *count = 2 * nq; break;
        case GPARML_A_GRAD_NEW: *ptr = c->grad_new; *count = 2 * nq; break;
        case GPARML_A_GRAD_OLD: *ptr = c->grad_old; *count = 2 * nq; break;
        case GPARML_A_STATS: *ptr = c->stats; *count = c->L.count; break;
        case GPARML_A_KMM: *ptr = c->kmm; *count = MM; break;
        case GPARML_A_KMM_INV: *ptr = c->kmm_inv; *count = MM; break;
        case GPARML_A_A_INV: *ptr = c->a_inv; *count = MM; break;
        case GPARML_A_DF_DKMM: *ptr = c->g_k; *count = MM; break;
        case GPARML_A_DF_DPSI1Y: *ptr = c->g_1; *count = (int64_t)c->M * c->D; break;
        case GPARML_A_DF_DPSI2: *ptr = c->g_2; *count = MM; break;
        case GPARML_A_PSI1: *ptr = c->psi1; *count = c->n * c->M; break;
        case GPARML_A_GRAD_X_MU: *ptr = c->gx_mu; *count = nq; break;
        case GPARML_A_GRAD_X_S: *ptr = c->gx_s; *count = nq; break;
        case GPARML_A_Y: *ptr = c->Y; *count = c->n * c->D; break;
        case GPARML_A_GRAD_GLOBAL: *ptr = c->glob_out + 1; *count = (int64_t)c->M * c->Q + c->Q + 2; break;
        case GPARML_A_GS_EXTRA: *ptr = c->glob_out + 1 + (int64_t)c->M * c->Q + c->Q + 2; *count = 13; break;
        default: gp_set_error("unknown array id %d", id); return GPARML_ERR_ARG;
    }
    (void)for_write;
    return GPARML_OK;
}

extern "C" int64_t gparml_array_count(const gparml_ctx *c, int id)
{
    if (!c) return -1;
    double *p; int64_t n;
    if (resolve(const_cast<gparml_ctx *>(c), id, &p, &n, false) != GPARML_OK) return -1;
    return n;
}

extern "C" int gparml_array_device_ptr(gparml_ctx *c, int id, void **out)
{
    CHECK_CTX(c);
    double *p; int64_t n;
    GP_TRY(resolve(c, id, &p, &n, false));
    *out = p;
    return GPARML_OK;
}

extern "C" int gparml_download(gparml_ctx *c, int id, double *dst, int64_t count)
{
    CHECK_CTX(c);
    GP_TRY(wait_y(c));
    if (id == GPARML_A_PSI1) {
        GP_TRY(prep_if_needed(c));
        if (!c->psi1) GP_TRY(dev_alloc(&c->psi1, (size_t)c->n * c->M));
        GP_TRY(gp_launch_psi1_matrix(c));
    }
    double *p; int64_t n;
    GP_TRY(resolve(c, id, &p, &n, false));
    if (count != n) { gp_set_error("download(%d): count %lld != %lld", id, (long long)count, (long long)n); return GPARML_ERR_ARG; }
    if (n > 0 && (!p || !dst)) { gp_set_error("download(%d): array not available", id); return GPARML_ERR_STATE; }
    if (n > 0) GP_CUDA(cudaMemcpyAsync(dst, p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}

extern "C" int gparml_upload(gparml_ctx *c, int id, const double *src, int64_t count)
{
    CHECK_CTX(c);
    GP_TRY(finish_pending_gs(c));
    if (id >= GPARML_A_KMM && id != GPARML_A_KMM && id != GPARML_A_KMM_INV) { gp_set_error("upload(%d): array is read-only", id); return GPARML_ERR_ARG; }
    double *p; int64_t n;
    GP_TRY(resolve(c, id, &p, &n, true));
    if (count != n) { gp_set_error("upload(%d): count %lld != %lld", id, (long long)count, (long long)n); return GPARML_ERR_ARG; }
    if (n > 0 && (!p || !src)) { gp_set_error("upload(%d): array not available", id); return GPARML_ERR_STATE; }
    if (n > 0) GP_CUDA(cudaMemcpyAsync(p, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    if (id == GPARML_A_GRAD_D) c->have_dir = true;
    if (id == GPARML_A_X_MU || id == GPARML_A_X_S) c->have_prep = c->have_stats = false;
    if (id == GPARML_A_STATS) c->have_stats = true;
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// named statistics
// ---------------------------------------------------------------------------
static int ensure_named_tmp(gparml_ctx *c)
{
    const size_t need = (size_t)c->Q * c->M * c->M * 2 + (size_t)c->M * c->M + 16;
    if (c->named_tmp_count >= need) return GPARML_OK;
    GP_TRY(dev_alloc(&c->named_tmp, need));
    c->named_tmp_count = need;
    return GPARML_OK;
}

static int d2h(gparml_ctx *c, double *dst, const double *src, size_t count)
{
    GP_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return GPARML_OK;
}

extern "C" int gparml_stats_expand(gparml_ctx *c, const gparml_named_stats *o)
{
    CHECK_CTX(c);
    if (!o) { gp_set_error("null output struct"); return GPARML_ERR_ARG; }
    if (!c->have_globals) { gp_set_error("stats_expand: set_globals first"); return GPARML_ERR_STATE; }
    GP_TRY(ensure_named_tmp(c));
    const size_t MM = (size_t)c->M * c->M, MD = (size_t)c->M * c->D, QMM = MM * c->Q, MQD = MD * c->Q;
    double head[ST_HEAD];
    GP_TRY(d2h(c, head, c->stats, ST_HEAD));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    if (o->sum_YYT) *o->sum_YYT = head[ST_YYT];
    if (o->sum_exp_K_ii) *o->sum_exp_K_ii = head[ST_PSI0];
    if (o->sum_KL) *o->sum_KL = head[ST_KL];
    if (o->sum_d_exp_K_ii_d_sf2) *o->sum_d_exp_K_ii_d_sf2 = head[ST_NLOCAL];
    if (o->sum_exp_K_miY) GP_TRY(d2h(c, o->sum_exp_K_miY, c->stats + c->L.off_p1y, MD));
    if (o->sum_d_exp_K_miY_d_Z) GP_TRY(d2h(c, o->sum_d_exp_K_miY_d_Z, c->stats + c->L.off_d1z, MQD));
    if (o->sum_d_exp_K_miY_d_alpha) GP_TRY(d2h(c, o->sum_d_exp_K_miY_d_alpha, c->stats + c->L.off_d1a, MQD));
    if (o->sum_exp_K_mi_K_im) { GP_TRY(gp_launch_expand(c, c->named_tmp, 0)); GP_TRY(d2h(c, o->sum_exp_K_mi_K_im, c->named_tmp, MM)); GP_CUDA(cudaStreamSynchronize(c->stream)); }
    if (o->sum_d_exp_K_mi_K_im_d_Z) { GP_TRY(gp_launch_expand(c, c->named_tmp, 1)); GP_TRY(d2h(c, o->sum_d_exp_K_mi_K_im_d_Z, c->named_tmp, QMM)); GP_CUDA(cudaStreamSynchronize(c->stream)); }
    if (o->sum_d_exp_K_mi_K_im_d_alpha) { GP_TRY(gp_launch_expand(c, c->named_tmp, 2)); GP_TRY(d2h(c, o->sum_d_exp_K_mi_K_im_d_alpha, c->named_tmp, QMM)); GP_CUDA(cudaStreamSynchronize(c->stream)); }
    if (o->sum_d_exp_K_miY_d_sf2) { GP_TRY(gp_launch_expand(c, c->named_tmp, 3)); GP_TRY(d2h(c, o->sum_d_exp_K_miY_d_sf2, c->named_tmp, MD)); GP_CUDA(cudaStreamSynchronize(c->stream)); }
    if (o->sum_d_exp_K_mi_K_im_d_sf2) { GP_TRY(gp_launch_expand(c, c->named_tmp, 4)); GP_TRY(d2h(c, o->sum_d_exp_K_mi_K_im_d_sf2, c->named_tmp, MM)); GP_CUDA(cudaStreamSynchronize(c->stream)); }
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}

extern "C" int gparml_stats_set_named(gparml_ctx *c, const gparml_named_stats *in)
{
    CHECK_CTX(c);
    GP_TRY(finish_pending_gs(c));
    if (!in) { gp_set_error("null input struct"); return GPARML_ERR_ARG; }
    if (!c->have_globals) { gp_set_error("stats_set_named: set_globals first"); return GPARML_ERR_STATE; }
    GP_TRY(ensure_named_tmp(c));
    const size_t MM = (size_t)c->M * c->M, MD = (size_t)c->M * c->D, QMM = MM * c->Q, MQD = MD * c->Q;
    double head[ST_HEAD];
    GP_TRY(d2h(c, head, c->stats, ST_HEAD));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    if (in->sum_YYT) head[ST_YYT] = *in->sum_YYT;
    if (in->sum_exp_K_ii) head[ST_PSI0] = *in->sum_exp_K_ii;
    if (in->sum_KL) head[ST_KL] = *in->sum_KL;
    if (in->sum_d_exp_K_ii_d_sf2) head[ST_NLOCAL] = *in->sum_d_exp_K_ii_d_sf2;
    GP_CUDA(cudaMemcpyAsync(c->stats, head, sizeof(head), cudaMemcpyHostToDevice, c->stream));
    if (in->sum_exp_K_miY) GP_CUDA(cudaMemcpyAsync(c->stats + c->L.off_p1y, in->sum_exp_K_miY, MD * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (in->sum_d_exp_K_miY_d_Z) GP_CUDA(cudaMemcpyAsync(c->stats + c->L.off_d1z, in->sum_d_exp_K_miY_d_Z, MQD * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (in->sum_d_exp_K_miY_d_alpha) GP_CUDA(cudaMemcpyAsync(c->stats + c->L.off_d1a, in->sum_d_exp_K_miY_d_alpha, MQD * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (in->sum_exp_K_mi_K_im) {
        double *t_p2 = c->named_tmp, *t_dz = c->named_tmp + MM, *t_da = t_dz + QMM;
        GP_CUDA(cudaMemcpyAsync(t_p2, in->sum_exp_K_mi_K_im, MM * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (in->sum_d_exp_K_mi_K_im_d_Z) GP_CUDA(cudaMemcpyAsync(t_dz, in->sum_d_exp_K_mi_K_im_d_Z, QMM * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (in->sum_d_exp_K_mi_K_im_d_alpha) GP_CUDA(cudaMemcpyAsync(t_da, in->sum_d_exp_K_mi_K_im_d_alpha, QMM * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        GP_TRY(gp_launch_compact(c, t_p2, in->sum_d_exp_K_mi_K_im_d_Z ? t_dz : nullptr, in->sum_d_exp_K_mi_K_im_d_alpha ? t_da : nullptr));
    } else if (in->sum_d_exp_K_mi_K_im_d_Z || in->sum_d_exp_K_mi_K_im_d_alpha) {
        gp_set_error("stats_set_named: derivative tensors of Psi2 need sum_exp_K_mi_K_im as well");
        return GPARML_ERR_ARG;
    }
    GP_CUDA(cudaStreamSynchronize(c->stream));
    c->have_stats = true;
    c->have_global_step = false;
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// optimiser local state
// ---------------------------------------------------------------------------
#define SCG_GUARD(c)                                                                               \
    CHECK_CTX(c);                                                                                  \
    if (!(c)->have_shard) { gp_set_error("scg op: upload_shard first"); return GPARML_ERR_STATE; }

extern "C" int gparml_scg_set_grads(gparml_ctx *c) { SCG_GUARD(c); c->have_dir = true; return gp_scg_update(c, 0, 0.0); }
extern "C" int gparml_scg_get_mu(gparml_ctx *c, double *o) { SCG_GUARD(c); return gp_scg_reduce(c, 0, 0.0, o); }
extern "C" int gparml_scg_get_kappa(gparml_ctx *c, double *o) { SCG_GUARD(c); return gp_scg_reduce(c, 1, 0.0, o); }
extern "C" int gparml_scg_get_theta(gparml_ctx *c, double *o) { SCG_GUARD(c); return gp_scg_reduce(c, 2, 0.0, o); }
extern "C" int gparml_scg_get_current_grad(gparml_ctx *c, double *o) { SCG_GUARD(c); return gp_scg_reduce(c, 3, 0.0, o); }
extern "C" int gparml_scg_get_gamma(gparml_ctx *c, double *o) { SCG_GUARD(c); return gp_scg_reduce(c, 4, 0.0, o); }
extern "C" int gparml_scg_get_max_d(gparml_ctx *c, double a, double *o) { SCG_GUARD(c); return gp_scg_reduce(c, 5, a, o); }
extern "C" int gparml_scg_reset_d(gparml_ctx *c) { SCG_GUARD(c); c->have_dir = true; return gp_scg_update(c, 1, 0.0); }
extern "C" int gparml_scg_update_d(gparml_ctx *c, double g) { SCG_GUARD(c); return gp_scg_update(c, 2, g); }
extern "C" int gparml_scg_update_X(gparml_ctx *c, double a)
{
    SCG_GUARD(c);
    c->have_prep = c->have_stats = false;
    return gp_scg_update(c, 3, a);
}
extern "C" int gparml_scg_update_grad_old(gparml_ctx *c) { SCG_GUARD(c); return gp_scg_update(c, 4, 0.0); }
extern "C" int gparml_scg_update_grad_new(gparml_ctx *c) { SCG_GUARD(c); return gp_scg_update(c, 5, 0.0); }

// ---------------------------------------------------------------------------
// timing / probes
// ---------------------------------------------------------------------------
extern "C" int gparml_enable_timing(gparml_ctx *c, int on)
{
    CHECK_CTX(c);
    c->timing = on != 0;
    return GPARML_OK;
}

extern "C" int gparml_phase_times(gparml_ctx *c, double *out5)
{
    CHECK_CTX(c);
    if (!c->timing) { gp_set_error("phase_times: timing not enabled"); return GPARML_ERR_STATE; }
    GP_CUDA(cudaStreamSynchronize(c->stream));
    // out order: prep_points, psi1_stats, psi2_stats, global_step, embed_grads (psi2 runs before psi1)
    const int pairs[5][2] = {{0, 1}, {2, 3}, {1, 2}, {4, 5}, {6, 7}};
    for (int i = 0; i < 5; ++i) {
        float ms = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms, c->ev[pairs[i][0]], c->ev[pairs[i][1]]);
        if (e != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
        out5[i] = ms;
    }
    return GPARML_OK;
}

extern "C" int gparml_measure_dfma_peak(gparml_ctx *c, double *out)
{
    CHECK_CTX(c);
    return gp_measure_dfma(c, out);
}

// ---------------------------------------------------------------------------
// partial_terms helper surface: Kmm-side derivative tensors and chain-rule contractions
// ---------------------------------------------------------------------------
int gp_launch_kmm_deriv(gparml_ctx *c, int which, double *dev_out);
int gp_launch_grad_Z_contract(gparml_ctx *c, const double *A, const double *B, const double *C, const double *E, const double *G,
                              const double *H, double *out);
int gp_launch_grad_q_contract(gparml_ctx *c, int nq, const double *A, const double *B, const double *C, const double *E,
                              const double *G, const double *H, double *out);

extern "C" int gparml_kmm_derivative(gparml_ctx *c, int which, double *out)
{
    CHECK_CTX(c);
    if (!c->have_globals || !out || which < 0 || which > 2) { gp_set_error("kmm_derivative: set_globals first / bad args"); return GPARML_ERR_ARG; }
    GP_TRY(ensure_named_tmp(c));
    GP_CUDA(cudaStreamWaitEvent(c->stream, c->ev_kmm, 0));
    const size_t total = which == 2 ? (size_t)c->M * c->M : (size_t)c->M * c->M * c->Q;
    GP_TRY(gp_launch_kmm_deriv(c, which, c->named_tmp));
    GP_TRY(d2h(c, out, c->named_tmp, total));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}

// grad_Z (which = 0, out (M,Q)), grad_alpha (which = 1, out (Q)) or the matrix part of
// grad_sf2 (which = 2, out (1)) from six caller-owned host tensors in the reference's layouts.
extern "C" int gparml_grad_contract(gparml_ctx *c, int which, const double *dF_dKmm, const double *dKmm_dx,
                                    const double *dF_dPsi1Y, const double *dPsi1Y_dx, const double *dF_dPsi2,
                                    const double *dPsi2_dx, double *out)
{
    CHECK_CTX(c);
    if (!dF_dKmm || !dKmm_dx || !dF_dPsi1Y || !dPsi1Y_dx || !dF_dPsi2 || !dPsi2_dx || !out || which < 0 || which > 2) {
        gp_set_error("grad_contract: null tensor or bad selector");
        return GPARML_ERR_ARG;
    }
    const size_t MM = (size_t)c->M * c->M, MD = (size_t)c->M * c->D;
    const size_t mult = which == 2 ? 1 : (size_t)c->Q;
    const size_t need = 2 * MM + MD + mult * (2 * MM + MD) + (size_t)c->M * c->Q + 16;
    GP_TRY(gp_ensure_ws(c, need * sizeof(double)));
    double *A = c->ws, *G = A + MM, *C = G + MM, *B = C + MD, *H = B + mult * MM, *E = H + mult * MM, *o = E + mult * MD;
    GP_CUDA(cudaMemcpyAsync(A, dF_dKmm, MM * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(cudaMemcpyAsync(G, dF_dPsi2, MM * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(cudaMemcpyAsync(C, dF_dPsi1Y, MD * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(cudaMemcpyAsync(B, dKmm_dx, mult * MM * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(cudaMemcpyAsync(H, dPsi2_dx, mult * MM * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(cudaMemcpyAsync(E, dPsi1Y_dx, mult * MD * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    size_t nout;
    if (which == 0) { GP_TRY(gp_launch_grad_Z_contract(c, A, B, C, E, G, H, o)); nout = (size_t)c->M * c->Q; }
    else { GP_TRY(gp_launch_grad_q_contract(c, (int)mult, A, B, C, E, G, H, o)); nout = mult; }
    GP_TRY(d2h(c, out, o, nout));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// in-process reduce across the shard contexts of ONE host thread (any mix of devices); no host synchronisation
// ---------------------------------------------------------------------------
// order c's stream behind everything queued so far on other's stream (works across devices)
static int wait_for(gparml_ctx *c, gparml_ctx *other)
{
    if (other == c) return GPARML_OK;
    GP_CUDA(cudaSetDevice(other->device));
    GP_CUDA(cudaEventRecord(other->ev_main, other->stream));
    GP_CUDA(cudaSetDevice(c->device));
    GP_CUDA(cudaStreamWaitEvent(c->stream, other->ev_main, 0));
    return GPARML_OK;
}

// stats += packed buffer of a context on ANOTHER device (peer copy into scratch, then the same fixed-order add)
extern "C" int gparml_stats_add_peer(gparml_ctx *c, gparml_ctx *other, double scale)
{
    CHECK_CTX(c);
    GP_TRY(finish_pending_gs(c));
    if (!other || other->L.count != c->L.count) { gp_set_error("stats_add_peer: incompatible contexts"); return GPARML_ERR_ARG; }
    GP_TRY(gp_ensure_ws(c, (size_t)c->L.count * sizeof(double)));
    GP_TRY(wait_for(c, other));
    GP_CUDA(cudaMemcpyPeerAsync(c->ws, c->device, other->stats, other->device, (size_t)c->L.count * sizeof(double), c->stream));
    GP_TRY(gp_launch_stats_add(c, c->ws, scale));
    c->have_global_step = false;
    return GPARML_OK;
}

// stats = packed buffer of `other` (same or another device)
extern "C" int gparml_stats_copy_peer(gparml_ctx *c, gparml_ctx *other)
{
    CHECK_CTX(c);
    GP_TRY(finish_pending_gs(c));
    if (!other || other->L.count != c->L.count) { gp_set_error("stats_copy_peer: incompatible contexts"); return GPARML_ERR_ARG; }
    if (other == c) return GPARML_OK;
    GP_TRY(wait_for(c, other));
    GP_CUDA(cudaMemcpyPeerAsync(c->stats, c->device, other->stats, other->device, (size_t)c->L.count * sizeof(double), c->stream));
    c->have_stats = true;
    c->have_global_step = false;
    return GPARML_OK;
}

int gp_launch_stats_allreduce(gparml_ctx *root, double *const *bufs, int n, double scale);

// The reducer (local_MapReduce.py:250-277) for n shard contexts driven by one host thread: ONE kernel on the first
// context's device reads every context's packed buffer -- directly over NVLink peer memory when the contexts live on
// different GPUs -- adds them in list order (deterministic), scales (drop-out, :263-264) and stores the sum back into
// EVERY context's buffer, so each GPU can run its replicated master step without a second transfer.  Streams are
// ordered with events; the host never waits.  Falls back to peer copies when a pair of devices has no peer access.
extern "C" int gparml_stats_allreduce_peers(gparml_ctx **ctxs, int n, double scale)
{
    if (!ctxs || n < 1 || n > GPARML_MAX_PEERS) { gp_set_error("stats_allreduce_peers: 1..%d contexts", GPARML_MAX_PEERS); return GPARML_ERR_ARG; }
    gparml_ctx *root = ctxs[0];
    for (int g = 0; g < n; ++g) {
        if (!ctxs[g] || ctxs[g]->L.count != root->L.count) { gp_set_error("stats_allreduce_peers: incompatible contexts"); return GPARML_ERR_ARG; }
        for (int h = 0; h < g; ++h)
            if (ctxs[h] == ctxs[g]) { gp_set_error("stats_allreduce_peers: context listed twice"); return GPARML_ERR_ARG; }
        CHECK_CTX(ctxs[g]);
        GP_TRY(finish_pending_gs(ctxs[g]));
    }
    GP_CUDA(cudaSetDevice(root->device));
    bool direct = true;
    for (int g = 1; g < n && direct; ++g) {
        if (ctxs[g]->device == root->device) continue;
        int can = 0;
        GP_CUDA(cudaDeviceCanAccessPeer(&can, root->device, ctxs[g]->device));
        if (!can) { direct = false; break; }
        cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[g]->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) { cudaGetLastError(); direct = false; }
    }
    if (direct) {
        double *bufs[GPARML_MAX_PEERS];
        for (int g = 0; g < n; ++g) {
            bufs[g] = ctxs[g]->stats;
            GP_TRY(wait_for(root, ctxs[g]));
        }
        GP_TRY(gp_launch_stats_allreduce(root, bufs, n, scale));
        GP_CUDA(cudaEventRecord(root->ev_main, root->stream));
        for (int g = 1; g < n; ++g) {
            GP_CUDA(cudaSetDevice(ctxs[g]->device));
            GP_CUDA(cudaStreamWaitEvent(ctxs[g]->stream, root->ev_main, 0));
        }
    } else {
        for (int g = 1; g < n; ++g) GP_TRY(gparml_stats_add_peer(root, ctxs[g], g == n - 1 ? scale : 1.0));
        if (n == 1 && scale != 1.0) GP_TRY(gp_launch_stats_add(root, root->stats, 0.5 * scale));
        for (int g = 1; g < n; ++g) GP_TRY(gparml_stats_copy_peer(ctxs[g], root));
    }
    for (int g = 0; g < n; ++g) { ctxs[g]->have_stats = true; ctxs[g]->have_global_step = false; }
    return GPARML_OK;
}
