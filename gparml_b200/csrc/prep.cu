// K0: prep_points -- per-point derived quantities, KL, tr(YY^T), pair constants.
//
// Replaces the mapper glue of the reference (all citations relative to /root/reference):
//   local_MapReduce.py:205-214 / 333-341  X += step * grad_d (in memory only), S = softplus(S_raw)
//   partial_terms.py:40                    sum_YYT
//   partial_terms.py:83-87                 KL
// and hoists everything that depends on a point but not on an inducing point out of the
// Psi kernels (SURVEY.md section 7 "compute once per point in K0 and stream them").
//
// HBM-bound streaming kernels: each point reads (2 or 4)*Q doubles and writes
// 2*R + 2*Q doubles (R = 3Q+1 padded to even), fully coalesced through L1.
#include <math.h>

#include "common.cuh"

#define LIM_VAL 36.04365338911715  // -log(DBL_EPSILON), supporting_functions.py:125

// ---------------------------------------------------------------------------
// sum of squares of Y (partial_terms.py:40), deterministic two-stage reduction
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const double *__restrict__ y, int64_t count, double *__restrict__ partials)
{
    __shared__ double sh[33];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double v = y[i];
        acc = fma(v, v, acc);
    }
    acc = gp_block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// sums `k` partials (stride 1) in a fixed order and writes dst[0] = scale * sum
__global__ void __launch_bounds__(256) final_sum_kernel(const double *__restrict__ partials, int k, double scale, double *__restrict__ dst)
{
    __shared__ double sh[33];
    double acc = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) acc += partials[i];
    acc = gp_block_sum(acc, sh);
    if (threadIdx.x == 0) dst[0] = scale * acc;
}

// sum_n y_n . y_n -> d_yyt[0] on stream s (the copy stream right behind the Y upload); the partials
// live in d_yyt[1..] so that the reduction cannot collide with red_ws users on the main stream.
int gp_launch_yyt(gparml_ctx *c, cudaStream_t s)
{
    const int64_t count = c->n * c->D;
    int blocks = (int)((count + 256 * 8 - 1) / (256 * 8));
    if (blocks < 1) blocks = 1;
    if (blocks > 1024) blocks = 1024;
    sumsq_kernel<<<blocks, 256, 0, s>>>(c->Y, count, c->d_yyt + 1);
    GP_LAUNCH_CHECK(c);
    final_sum_kernel<<<1, 256, 0, s>>>(c->d_yyt + 1, blocks, 1.0, c->d_yyt);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

__global__ void set_yyt_kernel(const double *__restrict__ d_yyt, double *__restrict__ stats) { stats[ST_YYT] = d_yyt[0]; }

int gp_launch_set_yyt(gparml_ctx *c)
{
    set_yyt_kernel<<<1, 1, 0, c->stream>>>(c->d_yyt, c->stats);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// prep_points
// ---------------------------------------------------------------------------
struct PrepParams {
    int64_t n;              // points in the shard (offset of the variance half of grad_d)
    int64_t i0, i1;         // this launch covers points [i0, i1)
    int Q, R;
    const double *x_mu, *x_s, *grad_d;  // grad_d may be null
    double step;
    int mode;  // 0 = unconstrained variance (softplus), 1 = positive as given, 2 = fixed embeddings (S as given, KL = 0)
    const GlobalsDev *glob;
    double *rec1, *rec2, *s_pos, *s_sig;
    float *rec2f;           // optional fp32 copy of the Psi2 records (centred means), RF floats per point
    int RF;
    double *rec2x;          // optional records of psi2x_stats (common.cuh), 4Q + 2 doubles per point
    double *kl_partials;    // [gridDim.x][2]: (kl sum, number of exactly-zero variances)
    int *status;
};

#ifndef PREP_TP
#define PREP_TP 128        // points per tile
#endif
#define PREP_THREADS 256
#ifndef PREP_MINB
#define PREP_MINB 3        // resident CTAs per SM the register budget is capped for (B200, c3 with the psi2x records: 6 CTAs 0.505 ms,
                           // 5 (48 registers, spills) 0.429, 4 0.390, 3 0.385)
#endif

// Two phases per tile of 128 points:
//   A  thread per (point, q) element: reads mu / S_raw / direction, evaluates the per-element quantities and stores
//      them where they live in the records -- (mu_q, a_q) and (mu_q, w_q) are 16-byte pairs at offset 2q of a
//      256-byte-aligned record, so the 10 consecutive lanes of a point write 160 contiguous bytes (five full
//      sectors), the v_q 80 contiguous bytes -- plus S, sigmoid and the two denominators (shared memory);
//   B  thread per point: log-prefactors from the products over q, one 16-byte store per record.
// History (B200, c3, 992 B/point): one thread per point writing its own 256-byte records 1.05 TB/s; records
// staged through 77 kB of shared memory per CTA (2 CTAs/SM, 3 barriers per tile, 50 % bank conflicts) 2.6 TB/s;
// this version (20 kB, direct stores, one exp less per element): see DESIGN.md.
__global__ void __launch_bounds__(PREP_THREADS, PREP_MINB) prep_points_kernel(PrepParams p)
{
    extern __shared__ __align__(16) double dens[];      // [tile parity][den1 | den2][PREP_TP * Q]
    __shared__ double sh[33];
    __shared__ GlobalsDev g;
    __shared__ double ninv_al[GP_MAX_Q];                // -1 / alpha_q
    const int Q = p.Q, R = p.R, tid = threadIdx.x;
    const int RF = p.RF;
    if (tid == 0) g = *p.glob;
    if (tid < Q) ninv_al[tid] = -1.0 / p.glob->alpha[tid];
    __syncthreads();
    double kl = 0.0, zeros = 0.0;
    bool bad = false, big = false;
    const int RX = gp_recx_len(Q);
    const bool stepping = p.mode == 0 && p.grad_d != nullptr && p.step != 0.0;
    int parity = 0;
    for (int64_t base = p.i0 + (int64_t)blockIdx.x * PREP_TP; base < p.i1; base += (int64_t)gridDim.x * PREP_TP, parity ^= 1) {
        const int cnt = (int)((p.i1 - base < PREP_TP) ? (p.i1 - base) : PREP_TP);
        double *d1s = dens + (size_t)parity * 2 * PREP_TP * Q, *d2s = d1s + PREP_TP * Q;
        for (int e = tid; e < cnt * Q; e += PREP_THREADS) {
            const int pt = e / Q, q = e - pt * Q;
            const int64_t gi = base * Q + e;
            double mu = p.x_mu[gi], sr = p.x_s[gi], S, sig;
            if (p.mode == 0) {
                if (stepping) {                                            // local_MapReduce.py:205-211
                    mu = fma(p.grad_d[gi], p.step, mu);
                    sr = fma(p.grad_d[p.n * Q + gi], p.step, sr);
                }
                if (!(fabs(sr) < LIM_VAL)) bad = true;                      // supporting_functions.py:154
                const double t = exp(sr), u = 1.0 + t;
                S = log(u);                                                // supporting_functions.py:155 (same naive form)
                sig = t / u;                                               // = 1 / (exp(-sr) + 1), supporting_functions.py:167
            } else {
                S = sr;
                sig = 1.0;
            }
            const double al = g.alpha[q];
            const double den1 = fma(al, S, 1.0), den2 = fma(2.0 * al, S, 1.0);
            const double a = al / den1, w = al / den2;
            double *r1 = p.rec1 + (base + pt) * R, *r2 = p.rec2 + (base + pt) * R;
            *reinterpret_cast<double2 *>(r1 + 2 * q) = make_double2(mu, a);
            *reinterpret_cast<double2 *>(r2 + 2 * q) = make_double2(mu, w);
            r1[2 * Q + q] = al * S * a;
            r2[2 * Q + q] = al * S * w;
            if (p.rec2f) {
                float *rf = p.rec2f + (base + pt) * RF;
                *reinterpret_cast<float2 *>(rf + 2 * q) = make_float2((float)(mu - g.center[q]), (float)w);
                rf[2 * Q + q] = (float)(al * S * w);
            }
            if (p.rec2x) {
                double *rx = p.rec2x + (base + pt) * RX;
                *reinterpret_cast<double2 *>(rx + 2 * q) = make_double2(-w, w * (mu - g.center[q]));
                *reinterpret_cast<double2 *>(rx + 2 * Q + 2 * q) = make_double2(al * S * w, den2 * ninv_al[q]);
                if (al * S > GP_PSI2X_ROBUST_AS) big = true;
            }
            d1s[e] = den1;
            d2s[e] = den2;
            p.s_pos[gi] = S;
            p.s_sig[gi] = sig;
            if (p.mode != 2) {                                             // partial_terms.py:83-87, the -Q spread over q
                if (S == 0.0) { zeros += 1.0; kl += mu * mu - 1.0; }
                else kl += S - log(S) + mu * mu - 1.0;
            }
        }
        __syncthreads();
        for (int pt = tid; pt < cnt; pt += PREP_THREADS) {
            double prod1 = 1.0, prod2 = 1.0;
            for (int q = 0; q < Q; ++q) { prod1 *= d1s[pt * Q + q]; prod2 *= d2s[pt * Q + q]; }
            const double l1 = g.log_sf2 - 0.5 * log(prod1), l2 = 2.0 * g.log_sf2 - 0.5 * log(prod2);
            double *r1 = p.rec1 + (base + pt) * R + 3 * Q, *r2 = p.rec2 + (base + pt) * R + 3 * Q;
            if (3 * Q + 1 < R) {          // even Q: prefactor and the padding double are one aligned 16-byte store
                *reinterpret_cast<double2 *>(r1) = make_double2(l1, 0.0);
                *reinterpret_cast<double2 *>(r2) = make_double2(l2, 0.0);
            } else {
                *r1 = l1;
                *r2 = l2;
            }
            if (p.rec2f) {
                float *rf = p.rec2f + (base + pt) * RF;
                rf[3 * Q] = (float)l2;
                for (int k = 3 * Q + 1; k < RF; ++k) rf[k] = 0.f;
            }
            if (p.rec2x) {
                double as = 0.0;                                  // sum_q alpha_q S_q = sum_q (den1 - 1)
                for (int q = 0; q < Q; ++q) as += d1s[pt * Q + q] - 1.0;
                *reinterpret_cast<double2 *>(p.rec2x + (base + pt) * RX + 4 * Q) = make_double2(l2 + as, l2);
            }
        }
        // no second barrier: the next tile writes the other half of the denominator buffer
    }
    if (bad) atomicOr(p.status, 4);
    if (big) atomicOr(p.status + 1, 1);
    kl = gp_block_sum(bad ? NAN : 0.5 * kl, sh);      // NaN marks the failed input check for prep_finish (ST_FLAGS)
    zeros = gp_block_sum(zeros, sh);
    if (tid == 0) {
        p.kl_partials[2 * blockIdx.x] = kl;
        p.kl_partials[2 * blockIdx.x + 1] = zeros;
    }
}

// Final KL + header of the packed statistics buffer.  One block.
__global__ void __launch_bounds__(256) prep_finish_kernel(const double *__restrict__ kl_partials, int k, double n_local, int Q, int mode,
                                                          double sf2, double *__restrict__ stats)
{
    __shared__ double sh[33];
    double a = 0.0, z = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        a += kl_partials[2 * i];
        z += kl_partials[2 * i + 1];
    }
    a = gp_block_sum(a, sh);
    z = gp_block_sum(z, sh);
    if (threadIdx.x == 0) {
        double kl;
        stats[ST_FLAGS] = (a != a) ? 1.0 : 0.0;
        stats[ST_FLAGS + 1] = 0.0;
        if (a != a) kl = 0.0;                                         // input check failed: the evaluation raises
        else if (mode == 2 || z == n_local * (double)Q) kl = 0.0;        // all variances zero: fixed embeddings (partial_terms.py:86-87)
        else if (z > 0.0) kl = INFINITY;                              // -log(0) for some but not all entries, as numpy would give
        else kl = a;
        stats[ST_PSI0] = sf2 * n_local;                               // partial_terms.py:81 (ST_YYT: set_yyt_kernel)
        stats[ST_KL] = kl;
        stats[ST_NLOCAL] = n_local;                                   // partial_terms.py:318-320
    }
}

// prep_points over the points [i0, i1); block partials of the KL sum go to kl_partials[2 * blocks]
int gp_launch_prep_range(gparml_ctx *c, int64_t i0, int64_t i1, double *kl_partials, int max_blocks, int *blocks_used)
{
    PrepParams p;
    p.n = c->n; p.i0 = i0; p.i1 = i1;
    p.Q = c->Q;
    p.R = gp_rec_len(c->Q);
    p.x_mu = c->x_mu;
    p.x_s = c->x_s;
    p.grad_d = c->have_dir ? c->grad_d : nullptr;
    p.step = c->step_size;
    p.mode = (c->flags & GPARML_FLAG_FIXED_EMBEDDINGS) ? 2 : (c->variance_domain == GPARML_VARIANCE_POSITIVE ? 1 : 0);
    p.glob = c->d_glob;
    p.rec1 = c->rec1;
    p.rec2 = c->rec2;
    p.s_pos = c->s_pos;
    p.s_sig = c->s_sig;
    p.status = c->d_status;
    p.rec2f = (c->flags & GPARML_FLAG_FP32_MAP) ? c->rec2f : nullptr;
    p.RF = gp_rec_len_f32(c->Q);
    p.rec2x = (c->flags & GPARML_FLAG_FP32_MAP) ? nullptr : c->rec2x;
    if (i0 == 0) GP_CUDA(cudaMemsetAsync(c->d_status + 1, 0, sizeof(int), c->stream));     // ranges arrive in order
    int blocks = (int)((i1 - i0 + PREP_TP - 1) / PREP_TP);
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)4 * PREP_TP * p.Q * sizeof(double);
    GP_CUDA(cudaFuncSetAttribute(prep_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.kl_partials = kl_partials;
    prep_points_kernel<<<blocks, PREP_THREADS, smem, c->stream>>>(p);
    GP_LAUNCH_CHECK(c);
    *blocks_used = blocks;
    return GPARML_OK;
}

int gp_launch_prep_finish(gparml_ctx *c, const double *kl_partials, int blocks)
{
    const int mode = (c->flags & GPARML_FLAG_FIXED_EMBEDDINGS) ? 2 : (c->variance_domain == GPARML_VARIANCE_POSITIVE ? 1 : 0);
    prep_finish_kernel<<<1, 256, 0, c->stream>>>(kl_partials, blocks, (double)c->n, c->Q, mode, c->h_glob.sf2, c->stats);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_prep(gparml_ctx *c)
{
    int blocks = 1;
    // persistent grid: every CTA resident (PREP_MINB per SM), tiles dealt grid-stride
    GP_TRY(gp_ensure_ws(c, (size_t)c->sm_count * PREP_MINB * 2 * sizeof(double)));
    GP_TRY(gp_launch_prep_range(c, 0, c->n, c->ws, c->sm_count * PREP_MINB, &blocks));
    return gp_launch_prep_finish(c, c->ws, blocks);
}

// ---------------------------------------------------------------------------
// pair table: (m, m') indices of the upper triangle and the point-independent part
// of the Psi2 exponent, lk = -1/4 sum_q alpha_q (z_mq - z_m'q)^2  (kernel_exp.py:143); for
// embed_grads also the centred midpoints zc = (z_m + z_m')/2 - center and their squares.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pair_table_kernel(const double *__restrict__ Z, int M, int Q, const GlobalsDev *__restrict__ glob,
                                                         int2 *__restrict__ pair_idx, double *__restrict__ pair_lk,
                                                         double2 *__restrict__ pair_zz, double *__restrict__ pair_zc,
                                                         double *__restrict__ pair_r)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * M) return;
    const int a = (int)(idx / M), b = (int)(idx % M);
    if (b < a) return;
    double s = 0.0;
    for (int q = 0; q < Q; ++q) {
        const double dz = Z[a * Q + q] - Z[b * Q + q];
        s = fma(glob->alpha[q] * dz, dz, s);
    }
    const int64_t p = gp_pair_index(M, a, b);
    const int QP = (Q + 1) & ~1;
    pair_idx[p] = make_int2(a, b);
    pair_lk[p] = -0.25 * s;
    for (int q = 0; q < Q; ++q) {
        const double zc = 0.5 * (Z[a * Q + q] + Z[b * Q + q]) - glob->center[q];
        pair_zz[p * Q + q] = make_double2(zc, zc * zc);
        pair_zc[p * QP + q] = zc;
    }
    if (QP > Q) pair_zc[p * QP + Q] = 0.0;
    if (pair_r) {       // features (zc, zc^2, 1) of embed_psi2m: chunk [p / 64], element (feature / 4, p % 64, feature % 4)
        const int NB = 2 * GP_PAIR_R_TILES(Q);
        double *blk = pair_r + (size_t)(p / GP_PAIR_CHUNK) * NB * GP_PAIR_CHUNK * 4 + (p % GP_PAIR_CHUNK) * 4;
        for (int f = 0; f <= 2 * Q; ++f) {
            const int q = f < Q ? f : f - Q;
            const double zc = 0.5 * (Z[a * Q + q] + Z[b * Q + q]) - glob->center[q];
            blk[(size_t)(f >> 2) * GP_PAIR_CHUNK * 4 + (f & 3)] = f == 2 * Q ? 1.0 : (f < Q ? zc : zc * zc);
        }
    }
}

int gp_launch_pair_table(gparml_ctx *c)
{
    const int64_t total = (int64_t)c->M * c->M;
    pair_table_kernel<<<(int)((total + 255) / 256), 256, 0, c->stream>>>(c->Z, c->M, c->Q, c->d_glob, c->pair_idx, c->pair_lk, c->pair_zz, c->pair_zc, c->pair_r);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}
