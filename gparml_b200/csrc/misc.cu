// Layout conversion between the packed (symmetric-compact) statistics and the reference's
// named arrays, the optimiser's local-state vector operations (K6), and the DFMA peak probe.
#include <math.h>

#include "common.cuh"

// ---------------------------------------------------------------------------
// expand: packed -> reference layouts (parallel_GPLVM.py:142-151)
//   which 0: sum_exp_K_mi_K_im           (M, M)      = S0[p(m,m')] * scale
//   which 1: sum_d_exp_K_mi_K_im_d_Z     (M, Q, M)   partial_terms.py:201-203
//   which 2: sum_d_exp_K_mi_K_im_d_alpha (Q, M, M)   partial_terms.py:279-282
//   which 3: sum_exp_K_miY * scale       (M, D)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_kernel(const double *__restrict__ stats, int64_t off_s0, int64_t off_tz, int64_t off_ta,
                                                     int64_t off_p1y, int64_t P, const double *__restrict__ Z,
                                                     const GlobalsDev *__restrict__ glob, int M, int Q, int D, int which,
                                                     double scale, double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (which == 0) {
        if (idx >= (int64_t)M * M) return;
        const int a = (int)(idx / M), b = (int)(idx % M);
        const int64_t p = (a <= b) ? gp_pair_index(M, a, b) : gp_pair_index(M, b, a);
        out[idx] = scale * stats[off_s0 + p];
    } else if (which == 1) {
        if (idx >= (int64_t)M * Q * M) return;
        const int b = (int)(idx % M), q = (int)((idx / M) % Q), a = (int)(idx / ((int64_t)M * Q));
        const int64_t p = (a <= b) ? gp_pair_index(M, a, b) : gp_pair_index(M, b, a);
        const double dz = Z[a * Q + q] - Z[b * Q + q];
        out[idx] = -0.5 * glob->alpha[q] * dz * stats[off_s0 + p] + stats[off_tz + (int64_t)q * P + p];
    } else if (which == 2) {
        if (idx >= (int64_t)Q * M * M) return;
        const int b = (int)(idx % M), a = (int)((idx / M) % M), q = (int)(idx / ((int64_t)M * M));
        const int64_t p = (a <= b) ? gp_pair_index(M, a, b) : gp_pair_index(M, b, a);
        const double dz = Z[a * Q + q] - Z[b * Q + q];
        const double al = glob->alpha[q];
        out[idx] = -0.25 * dz * dz * stats[off_s0 + p] - stats[off_ta + (int64_t)q * P + p] / (al * al);
    } else {
        if (idx >= (int64_t)M * D) return;
        out[idx] = scale * stats[off_p1y + idx];
    }
}

int gp_launch_expand(gparml_ctx *c, double *dev_out, int which)
{
    int64_t total;
    double scale = 1.0;
    int w = which;
    switch (which) {
        case 0: total = (int64_t)c->M * c->M; break;
        case 1: total = (int64_t)c->M * c->Q * c->M; break;
        case 2: total = (int64_t)c->Q * c->M * c->M; break;
        case 3: total = (int64_t)c->M * c->D; scale = 1.0 / c->h_glob.sf2; break;               // partial_terms.py:310-312
        case 4: total = (int64_t)c->M * c->M; scale = 2.0 / c->h_glob.sf2; w = 0; break;        // partial_terms.py:314-316
        default: gp_set_error("expand: bad selector %d", which); return GPARML_ERR_ARG;
    }
    expand_kernel<<<(int)((total + 255) / 256), 256, 0, c->stream>>>(c->stats, c->L.off_s0, c->L.off_tz, c->L.off_ta, c->L.off_p1y,
                                                                    c->L.P, c->Z, c->d_glob, c->M, c->Q, c->D, w, scale, dev_out);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// compact: reference layouts -> packed (inverse of expand; for set_local_statistics)
__global__ void __launch_bounds__(256) compact_kernel(const double *__restrict__ psi2, const double *__restrict__ d2z,
                                                      const double *__restrict__ d2a, const double *__restrict__ Z,
                                                      const GlobalsDev *__restrict__ glob, int M, int Q, int64_t P,
                                                      const int2 *__restrict__ pair_idx, double *__restrict__ s0,
                                                      double *__restrict__ tz, double *__restrict__ ta)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int2 ab = pair_idx[p];
    const int a = ab.x, b = ab.y;
    const double ps = 0.5 * (psi2[(size_t)a * M + b] + psi2[(size_t)b * M + a]);
    s0[p] = ps;
    for (int q = 0; q < Q; ++q) {
        const double dz = Z[a * Q + q] - Z[b * Q + q];
        const double al = glob->alpha[q];
        if (d2z) tz[(int64_t)q * P + p] = d2z[((size_t)a * Q + q) * M + b] + 0.5 * al * dz * ps;
        if (d2a) ta[(int64_t)q * P + p] = -(al * al) * (d2a[((size_t)q * M + a) * M + b] + 0.25 * dz * dz * ps);
    }
}

int gp_launch_compact(gparml_ctx *c, const double *dev_full_psi2, const double *dev_d2z, const double *dev_d2a)
{
    const int64_t P = c->L.P;
    compact_kernel<<<(int)((P + 255) / 256), 256, 0, c->stream>>>(dev_full_psi2, dev_d2z, dev_d2a, c->Z, c->d_glob, c->M, c->Q, P,
                                                                 c->pair_idx, c->stats + c->L.off_s0, c->stats + c->L.off_tz,
                                                                 c->stats + c->L.off_ta);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// stats = (stats + other) * scale  -- in-process reduce of several shards on one device
__global__ void __launch_bounds__(256) stats_add_kernel(double *__restrict__ dst, const double *__restrict__ src, int64_t count, double scale)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = (dst[i] + src[i]) * scale;
}

int gp_launch_stats_add(gparml_ctx *c, const double *src, double scale)
{
    stats_add_kernel<<<(int)((c->L.count + 255) / 256), 256, 0, c->stream>>>(c->stats, src, c->L.count, scale);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// One-kernel all-reduce over the packed buffers of up to GPARML_MAX_PEERS contexts (the buffers of contexts on
// other GPUs are addressed directly: NVLink peer loads and stores).  Each thread owns two consecutive elements of
// every buffer: it reads them from all n buffers (all loads in flight before the first add), sums in list order,
// scales, and writes the result back to all n buffers -- element-wise, so reading and writing the same buffers in
// one kernel is race-free.
struct PeerBufs { double *p[GPARML_MAX_PEERS]; };

__global__ void __launch_bounds__(256) stats_allreduce_kernel(PeerBufs b, int n, int64_t count, double scale)
{
    const int64_t i = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= count) return;
    if (i + 1 < count) {        // packed buffers are cudaMalloc'ed: 16-byte aligned
        double2 v[GPARML_MAX_PEERS];
#pragma unroll
        for (int g = 0; g < GPARML_MAX_PEERS; ++g)
            if (g < n) v[g] = __ldcg(reinterpret_cast<const double2 *>(b.p[g] + i));      // L2 only: never a stale L1 line
        double2 a = v[0];
#pragma unroll
        for (int g = 1; g < GPARML_MAX_PEERS; ++g)
            if (g < n) { a.x += v[g].x; a.y += v[g].y; }
        a.x *= scale; a.y *= scale;
#pragma unroll
        for (int g = 0; g < GPARML_MAX_PEERS; ++g)
            if (g < n) *reinterpret_cast<double2 *>(b.p[g] + i) = a;
    } else {
        double a = __ldcg(b.p[0] + i);
        for (int g = 1; g < n; ++g) a += __ldcg(b.p[g] + i);
        a *= scale;
        for (int g = 0; g < n; ++g) b.p[g][i] = a;
    }
}

int gp_launch_stats_allreduce(gparml_ctx *root, double *const *bufs, int n, double scale)
{
    PeerBufs b;
    for (int g = 0; g < GPARML_MAX_PEERS; ++g) b.p[g] = g < n ? bufs[g] : nullptr;
    const int64_t count = root->L.count, threads = (count + 1) / 2;
    stats_allreduce_kernel<<<(int)((threads + 255) / 256), 256, 0, root->stream>>>(b, n, count, scale);
    GP_LAUNCH_CHECK(root);
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// K6: optimiser local state (scg_adapted_local_MapReduce.py:29-243) on the device-resident
// (2, n, Q) vectors.  HBM-bound streaming; reductions are two-stage and deterministic.
//   reduce ops: 0 mu = <new, d>   1 kappa = <d, d>   2 theta = <d, latest - new>
//               3 current_grad = <new, new>   4 gamma = <new, old>   5 max |scale * d|
//   update ops: 0 set_grads (new = old = latest, d = -latest)   1 reset_d (d = -new)
//               2 update_d (d = scale * d - new)   3 update_X (X += scale * d)
//               4 grad_old = grad_new   5 grad_new = grad_latest
// ---------------------------------------------------------------------------
// element-wise term of reduce op `OP`
template <int OP>
__device__ __forceinline__ double scg_term(double acc, double scale, double latest, double gnew, double gold, double d)
{
    switch (OP) {
        case 0: return fma(gnew, d, acc);
        case 1: return fma(d, d, acc);
        case 2: return fma(d, latest - gnew, acc);
        case 3: return fma(gnew, gnew, acc);
        case 4: return fma(gnew, gold, acc);
        default: return fmax(acc, fabs(scale * d));
    }
}

// 16-byte loads, two independent partial sums per thread, only the vectors an op needs are read
template <int OP>
__global__ void __launch_bounds__(256) scg_reduce_kernel(double scale, int64_t len, const double *__restrict__ latest,
                                                         const double *__restrict__ gnew, const double *__restrict__ gold,
                                                         const double *__restrict__ d, double *__restrict__ partials)
{
    __shared__ double sh[33];
    constexpr bool need_l = OP == 2, need_n = OP == 0 || OP == 2 || OP == 3 || OP == 4, need_o = OP == 4,
                   need_d = OP == 0 || OP == 1 || OP == 2 || OP == 5;
    double acc0 = 0.0, acc1 = 0.0;
    const int64_t len2 = len >> 1;           // len = 2 n Q is even, the vectors are cudaMalloc'ed (16-byte aligned)
    const double2 *l2 = reinterpret_cast<const double2 *>(latest), *n2 = reinterpret_cast<const double2 *>(gnew),
                  *o2 = reinterpret_cast<const double2 *>(gold), *d2 = reinterpret_cast<const double2 *>(d);
    const double2 z = make_double2(0.0, 0.0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 vl = need_l ? l2[i] : z, vn = need_n ? n2[i] : z, vo = need_o ? o2[i] : z, vd = need_d ? d2[i] : z;
        acc0 = scg_term<OP>(acc0, scale, vl.x, vn.x, vo.x, vd.x);
        acc1 = scg_term<OP>(acc1, scale, vl.y, vn.y, vo.y, vd.y);
    }
    double acc = (OP == 5) ? fmax(acc0, acc1) : acc0 + acc1;
    if (OP == 5) {
        // block max (order independent, exact)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, sh[w]);
            partials[blockIdx.x] = m;
        }
    } else {
        acc = gp_block_sum(acc, sh);
        if (threadIdx.x == 0) partials[blockIdx.x] = acc;
    }
}

__global__ void __launch_bounds__(256) scg_final_kernel(int op, const double *__restrict__ partials, int k, double *__restrict__ dst)
{
    __shared__ double sh[33];
    if (op == 5) {
        if (threadIdx.x == 0) {
            double m = 0.0;
            for (int i = 0; i < k; ++i) m = fmax(m, partials[i]);
            dst[0] = m;
        }
        return;
    }
    double acc = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) acc += partials[i];
    acc = gp_block_sum(acc, sh);
    if (threadIdx.x == 0) dst[0] = acc;
}

int gp_scg_reduce(gparml_ctx *c, int op, double scale, double *host_out)
{
    const int64_t len = 2 * c->n * c->Q;
    int blocks = (int)((len + 256 * 8 - 1) / (256 * 8));
    if (blocks < 1) blocks = 1;
    if (blocks > c->sm_count * 8) blocks = c->sm_count * 8;       // <= 1184 partials (red_ws holds 2048), 8 CTAs per SM resident
#define SCG_RED(OP) scg_reduce_kernel<OP><<<blocks, 256, 0, c->stream>>>(scale, len, c->grad_latest, c->grad_new, c->grad_old, c->grad_d, c->red_ws)
    switch (op) {
        case 0: SCG_RED(0); break;
        case 1: SCG_RED(1); break;
        case 2: SCG_RED(2); break;
        case 3: SCG_RED(3); break;
        case 4: SCG_RED(4); break;
        default: SCG_RED(5); break;
    }
#undef SCG_RED
    GP_LAUNCH_CHECK(c);
    scg_final_kernel<<<1, 256, 0, c->stream>>>(op, c->red_ws, blocks, c->red_ws + 2048);
    GP_LAUNCH_CHECK(c);
    GP_CUDA(cudaMemcpyAsync(host_out, c->red_ws + 2048, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(cudaStreamSynchronize(c->stream));
    return GPARML_OK;
}

__global__ void __launch_bounds__(256) scg_update_x_scalar_kernel(double scale, int64_t half, const double *__restrict__ d,
                                                                  double *__restrict__ x_mu, double *__restrict__ x_s)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * half; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < half) x_mu[i] = __dadd_rn(x_mu[i], __dmul_rn(scale, d[i]));
        else x_s[i - half] = __dadd_rn(x_s[i - half], __dmul_rn(scale, d[i]));
    }
}

// update ops on 16-byte pairs; op 3 (X += scale d) runs on the mean and the variance half separately
template <int OP>
__global__ void __launch_bounds__(256) scg_update_kernel(double scale, int64_t len, double *__restrict__ latest,
                                                         double *__restrict__ gnew, double *__restrict__ gold, double *__restrict__ d,
                                                         double *__restrict__ x)
{
    const int64_t len2 = len >> 1;
    double2 *l2 = reinterpret_cast<double2 *>(latest), *n2 = reinterpret_cast<double2 *>(gnew), *o2 = reinterpret_cast<double2 *>(gold),
            *d2 = reinterpret_cast<double2 *>(d), *x2 = reinterpret_cast<double2 *>(x);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len2; i += (int64_t)gridDim.x * blockDim.x) {
        switch (OP) {
            case 0: { const double2 g = l2[i]; n2[i] = g; o2[i] = g; d2[i] = make_double2(-g.x, -g.y); } break;
            case 1: { const double2 g = n2[i]; d2[i] = make_double2(-g.x, -g.y); } break;
            case 2: {   // no FMA contraction: bit-equal to numpy
                const double2 v = d2[i], g = n2[i];
                d2[i] = make_double2(__dsub_rn(__dmul_rn(scale, v.x), g.x), __dsub_rn(__dmul_rn(scale, v.y), g.y));
            } break;
            case 3: {
                const double2 v = d2[i], xx = x2[i];
                x2[i] = make_double2(__dadd_rn(xx.x, __dmul_rn(scale, v.x)), __dadd_rn(xx.y, __dmul_rn(scale, v.y)));
            } break;
            case 4: o2[i] = n2[i]; break;
            default: n2[i] = l2[i]; break;
        }
    }
    if ((len & 1) && blockIdx.x == 0 && threadIdx.x == 0) {      // odd tail (op 3 with n Q odd)
        const int64_t i = len - 1;
        if (OP == 3) x[i] = __dadd_rn(x[i], __dmul_rn(scale, d[i]));
    }
}

int gp_scg_update(gparml_ctx *c, int op, double scale)
{
    const int64_t half = c->n * c->Q, len = 2 * half;
    if (len == 0) return GPARML_OK;
    int blocks = (int)((len + 256 * 8 - 1) / (256 * 8));
    if (blocks > c->sm_count * 8) blocks = c->sm_count * 8;
    if (blocks < 1) blocks = 1;
#define SCG_UPD(OP, LEN, D, X) scg_update_kernel<OP><<<blocks, 256, 0, c->stream>>>(scale, LEN, c->grad_latest, c->grad_new, c->grad_old, D, X)
    switch (op) {
        case 0: SCG_UPD(0, len, c->grad_d, nullptr); break;
        case 1: SCG_UPD(1, len, c->grad_d, nullptr); break;
        case 2: SCG_UPD(2, len, c->grad_d, nullptr); break;
        case 3:
            if (half % 2 == 0) {
                // x_mu and x_s are separate allocations: two launches over 16-byte pairs
                SCG_UPD(3, half, c->grad_d, c->x_mu);
                GP_LAUNCH_CHECK(c);
                SCG_UPD(3, half, c->grad_d + half, c->x_s);
            } else {
                // n Q odd: the variance half of grad_d starts 8 bytes off a 16-byte boundary -> scalar kernel
                scg_update_x_scalar_kernel<<<blocks, 256, 0, c->stream>>>(scale, half, c->grad_d, c->x_mu, c->x_s);
            }
            break;
        case 4: SCG_UPD(4, len, c->grad_d, nullptr); break;
        default: SCG_UPD(5, len, c->grad_d, nullptr); break;
    }
#undef SCG_UPD
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// FP64 pipe peak probe: 8 independent DFMA chains per thread, no memory traffic.  Gives the
// denominator for "% of FP64 pipe peak" measured on the box (MEASURED_PEAKS.json has no FP64).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double a, double b, double *__restrict__ sink)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) sink[0] = s;
}

int gp_measure_dfma(gparml_ctx *c, double *out)
{
    cudaEvent_t e0, e1;
    GP_CUDA(cudaEventCreate(&e0));
    GP_CUDA(cudaEventCreate(&e1));
    const int blocks = c->sm_count * 8, iters = 4096;
    dfma_probe_kernel<<<blocks, 256, 0, c->stream>>>(64, 0.999999, 1e-9, c->red_ws);   // warm-up
    GP_LAUNCH_CHECK(c);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        GP_CUDA(cudaEventRecord(e0, c->stream));
        dfma_probe_kernel<<<blocks, 256, 0, c->stream>>>(iters, 0.999999, 1e-9, c->red_ws);
        GP_LAUNCH_CHECK(c);
        GP_CUDA(cudaEventRecord(e1, c->stream));
        GP_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        GP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)blocks * 256.0 * (double)iters * 16.0 * 8.0;
        const double rate = ops / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out = best;
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// Kmm-side derivative tensors (partial_terms.py:146-160, 247-254, 306-308)
//   which 0: dKmm_dZ     (M, Q, M) = -alpha_q (z_aq - z_bq) Kmm[a,b]
//   which 1: dKmm_dalpha (Q, M, M) = -1/2 Kmm[a,b] (z_aq - z_bq)^2
//   which 2: dKmm_dsf2   (M, M)    = Kmm / sf2
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kmm_deriv_kernel(const double *__restrict__ kmm, const double *__restrict__ Z,
                                                        const GlobalsDev *__restrict__ glob, int M, int Q, int which,
                                                        double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (which == 0) {
        if (idx >= (int64_t)M * Q * M) return;
        const int b = (int)(idx % M), q = (int)((idx / M) % Q), a = (int)(idx / ((int64_t)M * Q));
        out[idx] = -glob->alpha[q] * (Z[a * Q + q] - Z[b * Q + q]) * kmm[(size_t)a * M + b];
    } else if (which == 1) {
        if (idx >= (int64_t)Q * M * M) return;
        const int b = (int)(idx % M), a = (int)((idx / M) % M), q = (int)(idx / ((int64_t)M * M));
        const double dz = Z[a * Q + q] - Z[b * Q + q];
        out[idx] = -0.5 * kmm[(size_t)a * M + b] * dz * dz;
    } else {
        if (idx >= (int64_t)M * M) return;
        out[idx] = kmm[idx] / glob->sf2;
    }
}

int gp_launch_kmm_deriv(gparml_ctx *c, int which, double *dev_out)
{
    const int64_t total = which == 2 ? (int64_t)c->M * c->M : (int64_t)c->M * c->M * c->Q;
    kmm_deriv_kernel<<<(int)((total + 255) / 256), 256, 0, c->stream>>>(c->kmm, c->Z, c->d_glob, c->M, c->Q, which, dev_out);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

// ---------------------------------------------------------------------------
// Chain-rule contractions with caller-supplied tensors (partial_terms.grad_Z :207-240,
// grad_alpha :286-299, grad_sf2 :322-333).  One block per output element, fixed-order sums.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) grad_Z_contract_kernel(const double *__restrict__ A /*dF_dKmm (M,M)*/,
                                                              const double *__restrict__ B /*dKmm_dZ (M,Q,M)*/,
                                                              const double *__restrict__ C /*dF_dPsi1Y (M,D)*/,
                                                              const double *__restrict__ E /*dPsi1Y_dZ (M,Q,D)*/,
                                                              const double *__restrict__ G /*dF_dPsi2 (M,M)*/,
                                                              const double *__restrict__ H /*dPsi2_dZ (M,Q,M)*/,
                                                              int M, int Q, int D, double *__restrict__ out)
{
    __shared__ double sh[33];
    const int j = blockIdx.x / Q, k = blockIdx.x % Q;
    const double *Bjk = B + ((size_t)j * Q + k) * M, *Hjk = H + ((size_t)j * Q + k) * M, *Ejk = E + ((size_t)j * Q + k) * D;
    double s = 0.0;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        // row j and column j of an (M, M) mask; the doubly-hit [j, j] entry counts once (partial_terms.py:226-231)
        const double a = (m == j) ? A[(size_t)j * M + j] : (A[(size_t)j * M + m] + A[(size_t)m * M + j]);
        s = fma(a, Bjk[m], s);
        s = fma(2.0 * G[(size_t)j * M + m], Hjk[m], s);
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) s = fma(C[(size_t)j * D + d], Ejk[d], s);
    s = gp_block_sum(s, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// out[q] = sum A*B[q] + sum C*E[q] + sum G*H[q]   (B, H: (Q,M,M); E: (Q,M,D)).  With Q = 1 and
// scalars folded by the caller this also serves grad_sf2.
__global__ void __launch_bounds__(256) grad_q_contract_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                              const double *__restrict__ C, const double *__restrict__ E,
                                                              const double *__restrict__ G, const double *__restrict__ H,
                                                              int M, int D, double *__restrict__ out)
{
    __shared__ double sh[33];
    const int q = blockIdx.x;
    const size_t MM = (size_t)M * M, MD = (size_t)M * D;
    double s = 0.0;
    for (size_t i = threadIdx.x; i < MM; i += blockDim.x) {
        s = fma(A[i], B[q * MM + i], s);
        s = fma(G[i], H[q * MM + i], s);
    }
    for (size_t i = threadIdx.x; i < MD; i += blockDim.x) s = fma(C[i], E[q * MD + i], s);
    s = gp_block_sum(s, sh);
    if (threadIdx.x == 0) out[q] = s;
}

int gp_launch_grad_Z_contract(gparml_ctx *c, const double *A, const double *B, const double *C, const double *E, const double *G,
                              const double *H, double *out)
{
    grad_Z_contract_kernel<<<c->M * c->Q, 128, 0, c->stream>>>(A, B, C, E, G, H, c->M, c->Q, c->D, out);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}

int gp_launch_grad_q_contract(gparml_ctx *c, int nq, const double *A, const double *B, const double *C, const double *E,
                              const double *G, const double *H, double *out)
{
    grad_q_contract_kernel<<<nq, 256, 0, c->stream>>>(A, B, C, E, G, H, c->M, c->D, out);
    GP_LAUNCH_CHECK(c);
    return GPARML_OK;
}
