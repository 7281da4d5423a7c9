// One-off initialisation on the device: the reference loads ALL outputs on the
// master to run an SVD-based PCA (local_MapReduce.py:52-65, supporting_functions.py:102-121), draws
// the initial variances there (local_MapReduce.py:88-93) and runs scipy's k-means on the first
// shard's embeddings for the inducing inputs (parallel_GPLVM.py:170-186).  Here every shard stays
// where it is: the kernels below produce per-shard partial sums (D column sums, the D x D centred
// scatter matrix, k (Q+1) cluster sums) that the host adds across shards / ranks, a D x D
// eigen-problem replaces the N x D SVD, and the projection writes X_mu in place.
//
// All kernels are HBM-bound streams over Y (n, D) or X_mu (n, Q); partial sums are reduced in a fixed
// order (except the shared-memory atomics of the k-means assignment, whose order only moves the
// last bits of the cluster means).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int INIT_THREADS = 256;
constexpr int INIT_MAX_CHUNKS = 256;

// ---- column sums of Y ------------------------------------------------------------------------
// chunk c owns rows [c*rows_per, ...); thread t walks the chunk's elements with stride
// blockDim*...: with the thread count a multiple of D each thread stays on one column.
__global__ void __launch_bounds__(INIT_THREADS)
colsum_kernel(const double *__restrict__ Y, int64_t n, int D, int64_t rows_per, double *__restrict__ partial)
{
    extern __shared__ double sh[];                      // (groups, D)
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
    if (D <= INIT_THREADS) {
        const int groups = INIT_THREADS / D, g = threadIdx.x / D, d = threadIdx.x % D;
        double acc = 0.0;
        if (g < groups)
            for (int64_t r = r0 + g; r < r1; r += groups) acc += Y[r * D + d];
        if (g < groups) sh[g * D + d] = acc;
        __syncthreads();
        if (threadIdx.x < D) {
            double t = 0.0;
            for (int k = 0; k < groups; ++k) t += sh[k * D + threadIdx.x];
            partial[(int64_t)blockIdx.x * D + threadIdx.x] = t;
        }
    } else {
        for (int d = threadIdx.x; d < D; d += INIT_THREADS) {
            double acc = 0.0;
            for (int64_t r = r0; r < r1; ++r) acc += Y[r * D + d];
            partial[(int64_t)blockIdx.x * D + d] = acc;
        }
    }
}

// out[j] = sum_c partial[c][j], fixed order
__global__ void sum_partials_kernel(const double *__restrict__ partial, int chunks, int64_t len, double *__restrict__ out)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    double t = 0.0;
    for (int c = 0; c < chunks; ++c) t += partial[(int64_t)c * len + j];
    out[j] = t;
}

// ---- centred scatter matrix sum_n (y_n - m)(y_n - m)^T -----------------------------------------
// grid = (chunks, tiles_i, tiles_j): a 16 x 16 output tile per CTA, rows staged through shared memory.
constexpr int SC_T = 16, SC_ROWS = 64;
__global__ void __launch_bounds__(SC_T *SC_T)
scatter_kernel(const double *__restrict__ Y, const double *__restrict__ mean, int64_t n, int D, int64_t rows_per,
               double *__restrict__ partial)
{
    __shared__ double a[SC_ROWS][SC_T + 1], b[SC_ROWS][SC_T + 1];
    const int ti = threadIdx.x / SC_T, tj = threadIdx.x % SC_T;
    const int i0 = blockIdx.y * SC_T, j0 = blockIdx.z * SC_T;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
    double acc = 0.0;
    for (int64_t rb = r0; rb < r1; rb += SC_ROWS) {
        const int rows = (int)min((int64_t)SC_ROWS, r1 - rb);
        __syncthreads();
        for (int e = threadIdx.x; e < SC_ROWS * SC_T; e += SC_T * SC_T) {
            const int r = e / SC_T, k = e % SC_T;
            double va = 0.0, vb = 0.0;
            if (r < rows) {
                if (i0 + k < D) va = Y[(rb + r) * D + i0 + k] - mean[i0 + k];
                if (j0 + k < D) vb = Y[(rb + r) * D + j0 + k] - mean[j0 + k];
            }
            a[r][k] = va;
            b[r][k] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < SC_ROWS; ++r) acc = fma(a[r][ti], b[r][tj], acc);
    }
    if (i0 + ti < D && j0 + tj < D)
        partial[(int64_t)blockIdx.x * D * D + (int64_t)(i0 + ti) * D + (j0 + tj)] = acc;
}

// ---- X_mu = (Y - mean) W ------------------------------------------------------------------------
__global__ void project_kernel(const double *__restrict__ Y, const double *__restrict__ mean, const double *__restrict__ W,
                               int64_t n, int D, int Q, double *__restrict__ x_mu)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * Q) return;
    const int64_t r = e / Q;
    const int q = (int)(e % Q);
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc = fma(Y[r * D + d] - __ldg(mean + d), __ldg(W + (int64_t)d * Q + q), acc);
    x_mu[e] = acc;
}

// ---- counter-based normal variates -------------------------------------------------------------
// element e of stream `seed` is a pure function of (seed, e): the draw does not depend on how the
// rows are sharded.  splitmix64 finaliser twice -> two 53-bit uniforms -> Box-Muller.
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double normal_at(uint64_t seed, uint64_t e)
{
    const uint64_t base = mix64(seed + 0x9E3779B97F4A7C15ull * (2 * e + 1));
    const uint64_t a = mix64(base), b = mix64(base + 0x9E3779B97F4A7C15ull);
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);           // [0, 1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// mode 0: X_S (uploaded domain) = softplus^-1(clip(0.5 + 0.01 z, 0.001, 1))  (local_MapReduce.py:90-93)
// mode 1: X_mu = z                                                            (init == 'random', :87)
__global__ void random_fill_kernel(double *__restrict__ out, int64_t count, uint64_t seed, int64_t first, int mode)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const double z = normal_at(seed, (uint64_t)(first + e));
    if (mode == 1) { out[e] = z; return; }
    const double s = fmin(fmax(0.5 + 0.01 * z, 0.001), 1.0);
    out[e] = log(expm1(s));
}

// ---- one k-means assignment + accumulation pass --------------------------------------------------
// thread = point: nearest of k centroids (first minimum on ties, like numpy.argmin in scipy's vq),
// cluster sums in shared memory, per-CTA partials out: [k][1+Q] (count, sum_q) then the summed
// Euclidean distance to the nearest centroid.
template <int Q>
__global__ void __launch_bounds__(INIT_THREADS)
kmeans_kernel(const double *__restrict__ x, int64_t n, const double *__restrict__ cent, int k, double *__restrict__ partial)
{
    extern __shared__ double sh[];
    double *cs = sh;                        // (k, Q)
    double *acc = cs + (size_t)k * Q;       // (k, 1+Q)
    double *red = acc + (size_t)k * (1 + Q);  // 33
    for (int e = threadIdx.x; e < k * Q; e += INIT_THREADS) cs[e] = cent[e];
    for (int e = threadIdx.x; e < k * (1 + Q); e += INIT_THREADS) acc[e] = 0.0;
    __syncthreads();
    double dist = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * INIT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * INIT_THREADS) {
        double p[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) p[q] = x[i * Q + q];
        double best = INFINITY;
        int arg = 0;
        for (int c = 0; c < k; ++c) {
            double d2 = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double t = p[q] - cs[c * Q + q];
                d2 = fma(t, t, d2);
            }
            if (d2 < best) { best = d2; arg = c; }
        }
        dist += sqrt(best);
        atomicAdd(acc + (size_t)arg * (1 + Q), 1.0);
#pragma unroll
        for (int q = 0; q < Q; ++q) atomicAdd(acc + (size_t)arg * (1 + Q) + 1 + q, p[q]);
    }
    const double tot = gp_block_sum(dist, red);
    const int64_t len = (int64_t)k * (1 + Q) + 1;
    double *o = partial + (int64_t)blockIdx.x * len;
    for (int e = threadIdx.x; e < k * (1 + Q); e += INIT_THREADS) o[e] = acc[e];
    if (threadIdx.x == 0) o[len - 1] = tot;
}

template <int Q>
int launch_kmeans_q(gparml_ctx *c, const double *cent, int k, int grid, size_t smem, double *partial)
{
    GP_CUDA(cudaFuncSetAttribute(kmeans_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kmeans_kernel<Q><<<grid, INIT_THREADS, smem, c->stream>>>(c->x_mu, c->n, cent, k, partial);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int chunks_for(const gparml_ctx *c, int64_t n, int tiles)
{
    int64_t want = (2 * (int64_t)c->sm_count + tiles - 1) / tiles;
    if (want > INIT_MAX_CHUNKS) want = INIT_MAX_CHUNKS;
    if (want > (n + 63) / 64) want = (n + 63) / 64;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace

// column sums of the shard's Y -> out_host (D)
int gp_init_colsum(gparml_ctx *c, double *out_host)
{
    const int D = c->D;
    const int chunks = chunks_for(c, c->n, 1);
    GP_TRY(gp_ensure_ws(c, ((size_t)chunks + 1) * D * sizeof(double)));
    double *partial = c->ws, *out = c->ws + (size_t)chunks * D;
    const int64_t rows_per = (c->n + chunks - 1) / chunks;
    const size_t smem = (size_t)(D <= INIT_THREADS ? (INIT_THREADS / D) * D : 1) * sizeof(double);
    colsum_kernel<<<chunks, INIT_THREADS, smem, c->stream>>>(c->Y, c->n, D, rows_per > 0 ? rows_per : 1, partial);
    GP_LAUNCH_CHECK(c);
    sum_partials_kernel<<<(D + 255) / 256, 256, 0, c->stream>>>(partial, chunks, D, out);
    GP_LAUNCH_CHECK(c);
    GP_CUDA(cudaMemcpyAsync(out_host, out, (size_t)D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}

// sum_n (y_n - mean)(y_n - mean)^T of the shard -> out_host (D, D)
int gp_init_scatter(gparml_ctx *c, const double *mean_host, double *out_host)
{
    const int D = c->D;
    const int tiles = (D + SC_T - 1) / SC_T;
    const int chunks = chunks_for(c, c->n, tiles * tiles);
    const size_t DD = (size_t)D * D;
    GP_TRY(gp_ensure_ws(c, (((size_t)chunks + 1) * DD + D) * sizeof(double)));
    double *partial = c->ws, *out = partial + (size_t)chunks * DD, *mean = out + DD;
    GP_CUDA(cudaMemcpyAsync(mean, mean_host, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int64_t rows_per = (c->n + chunks - 1) / chunks;
    scatter_kernel<<<dim3(chunks, tiles, tiles), SC_T * SC_T, 0, c->stream>>>(c->Y, mean, c->n, D, rows_per > 0 ? rows_per : 1, partial);
    GP_LAUNCH_CHECK(c);
    sum_partials_kernel<<<(unsigned)((DD + 255) / 256), 256, 0, c->stream>>>(partial, chunks, (int64_t)DD, out);
    GP_LAUNCH_CHECK(c);
    GP_CUDA(cudaMemcpyAsync(out_host, out, DD * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}

// X_mu = (Y - mean) W, W (D, Q) on the host
int gp_init_project(gparml_ctx *c, const double *mean_host, const double *W_host)
{
    const int D = c->D, Q = c->Q;
    GP_TRY(gp_ensure_ws(c, ((size_t)D * Q + D) * sizeof(double)));
    double *W = c->ws, *mean = W + (size_t)D * Q;
    GP_CUDA(cudaMemcpyAsync(W, W_host, (size_t)D * Q * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(cudaMemcpyAsync(mean, mean_host, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int64_t cnt = c->n * Q;
    if (cnt > 0) {
        project_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(c->Y, mean, W, c->n, D, Q, c->x_mu);
        GP_LAUNCH_CHECK(c);
    }
    return GPARML_OK;
}

int gp_init_random(gparml_ctx *c, int mode, uint64_t seed, int64_t row_offset)
{
    const int64_t cnt = c->n * c->Q;
    if (cnt > 0) {
        random_fill_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(mode == 1 ? c->x_mu : c->x_s, cnt, seed,
                                                                              row_offset * c->Q, mode);
        GP_LAUNCH_CHECK(c);
    }
    return GPARML_OK;
}

// one assignment pass of k-means over the shard's X_mu: out_host = [k][1+Q] (count, sums) + distance sum
int gp_init_kmeans_step(gparml_ctx *c, const double *cent_host, int k, double *out_host)
{
    const int Q = c->Q;
    const int64_t len = (int64_t)k * (1 + Q) + 1;
    int grid = 2 * c->sm_count;
    if ((int64_t)grid * INIT_THREADS > c->n) grid = (int)((c->n + INIT_THREADS - 1) / INIT_THREADS);
    if (grid < 1) grid = 1;
    const size_t smem = ((size_t)k * Q + (size_t)k * (1 + Q) + 40) * sizeof(double);
    if (smem > 200 * 1024) { gp_set_error("kmeans_step: k = %d centroids do not fit shared memory", k); return GPARML_ERR_ARG; }
    GP_TRY(gp_ensure_ws(c, (((size_t)grid + 1) * len + (size_t)k * Q) * sizeof(double)));
    double *partial = c->ws, *out = partial + (size_t)grid * len, *cent = out + len;
    GP_CUDA(cudaMemcpyAsync(cent, cent_host, (size_t)k * Q * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    switch (Q) {
#define KM_CASE(q) case q: GP_TRY(launch_kmeans_q<q>(c, cent, k, grid, smem, partial)); break;
        KM_CASE(1) KM_CASE(2) KM_CASE(3) KM_CASE(4) KM_CASE(5) KM_CASE(6) KM_CASE(7) KM_CASE(8)
        KM_CASE(9) KM_CASE(10) KM_CASE(11) KM_CASE(12) KM_CASE(13) KM_CASE(14) KM_CASE(15) KM_CASE(16)
#undef KM_CASE
        default: gp_set_error("kmeans_step: Q = %d unsupported", Q); return GPARML_ERR_ARG;
    }
    sum_partials_kernel<<<(unsigned)((len + 255) / 256), 256, 0, c->stream>>>(partial, grid, len, out);
    GP_LAUNCH_CHECK(c);
    GP_CUDA(cudaMemcpyAsync(out_host, out, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}
