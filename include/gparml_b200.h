/* gparml_b200 -- C ABI of the B200-native (sm_100a) implementation of GParML's
 * map-reduce variational-bound hot path.
 *
 * The reference (markvdw/GParML) is pure Python and has NO native/FFI boundary of
 * its own: its extension point is the duck-typed
 * `partial_terms` class (partial_terms.py:15) driven by the mapper/reducer
 * functions of `local_MapReduce.py`.  This header is the boundary underneath the
 * Python mirrors of those two interfaces (gparml_b200/partial_terms.py and
 * gparml_b200/b200_MapReduce.py); each entry point names the reference code it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns int (0 = ok, <0 error);
 *     gparml_last_error() gives the message of the last failure.
 *   - the caller owns all host memory; the library owns all device memory.
 *   - one context = one shard (one "node"/input file of the reference) on one GPU;
 *     contexts are not thread-safe; all work of a context is ordered on one CUDA
 *     stream (its own, or the one given with gparml_set_stream()).
 *   - host arrays are float64, C-order, exactly the shapes the reference uses.
 *   - there is no CPU fallback: gparml_create() fails without a CUDA device.
 */
#ifndef GPARML_B200_H
#define GPARML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gparml_ctx gparml_ctx;

/* return codes */
#define GPARML_OK             0
#define GPARML_ERR_CUDA      -1   /* CUDA runtime error (message has the CUDA string)          */
#define GPARML_ERR_ARG       -2   /* bad argument / unsupported shape (e.g. Q > 16)            */
#define GPARML_ERR_NOT_PD    -3   /* Kmm or Kmm + beta*Psi2 not positive definite -> LinAlgError */
#define GPARML_ERR_STATE     -4   /* call sequence error (e.g. statistics before upload)       */
#define GPARML_ERR_NO_DEVICE -5   /* no usable CUDA device: the product has no CPU path        */
#define GPARML_ERR_RANGE     -6   /* |unconstrained variance| >= 36.04 (supporting_functions.py:154 assert) */

/* gparml_create flags */
#define GPARML_FLAG_FP32_MAP          1  /* opt-in: Psi-statistics maps evaluate in fp32, accumulate in fp64 */
#define GPARML_FLAG_FIXED_EMBEDDINGS  2  /* --fixed_embeddings: X_S == 0, KL = 0, no embeddings map
                                            (local_MapReduce.py:103,203; partial_terms.py:83-87)          */
#define GPARML_FLAG_FIXED_BETA        4  /* --fixed_beta: beta gradient zeroed (parallel_GPLVM.py:363-366)  */

/* variance domain of gparml_upload_shard */
#define GPARML_VARIANCE_UNCONSTRAINED 0  /* contents of <shard>.variance.npy; softplus applied on device
                                            (local_MapReduce.py:214,341)                                   */
#define GPARML_VARIANCE_POSITIVE      1  /* already positive, as partial_terms.set_data receives it
                                            (partial_terms.py:38)                                          */

/* device arrays addressable through gparml_download / gparml_upload (float64) */
enum gparml_array {
    GPARML_A_X_MU = 0,        /* (n, Q)    persisted means,            <shard>.embedding.npy            */
    GPARML_A_X_S = 1,         /* (n, Q)    persisted variances in the uploaded domain, .variance.npy   */
    GPARML_A_GRAD_D = 2,      /* (2, n, Q) local search direction,     .grad_d.npy                      */
    GPARML_A_GRAD_LATEST = 3, /* (2, n, Q) -[dF/dmu, dF/dS*sigmoid],   .grad_latest.npy (local_MapReduce.py:357-360) */
    GPARML_A_GRAD_NEW = 4,    /* (2, n, Q)                             .grad_new.npy                    */
    GPARML_A_GRAD_OLD = 5,    /* (2, n, Q)                             .grad_old.npy                    */
    GPARML_A_STATS = 6,       /* packed partial sums, gparml_stats_count() doubles (layout: DESIGN.md)  */
    GPARML_A_KMM = 7,         /* (M, M)    cache_Kmm       (local_MapReduce.py:386-388)                 */
    GPARML_A_KMM_INV = 8,     /* (M, M)    cache_Kmm_inv   (local_MapReduce.py:392)                     */
    GPARML_A_A_INV = 9,       /* (M, M)    (Kmm + beta Psi2)^-1  (partial_terms.py:60)                  */
    GPARML_A_DF_DKMM = 10,    /* (M, M)    partial_terms.py:102-113                                     */
    GPARML_A_DF_DPSI1Y = 11,  /* (M, D)    partial_terms.py:115-121                                     */
    GPARML_A_DF_DPSI2 = 12,   /* (M, M)    partial_terms.py:123-131                                     */
    GPARML_A_PSI1 = 13,       /* (n, M)    exp_K_mi, computed on demand (kernel_exp.py:51-82)           */
    GPARML_A_GRAD_X_MU = 14,  /* (n, Q)    dF/dX_mu, positive domain   (partial_terms.py:367-398)       */
    GPARML_A_GRAD_X_S = 15,   /* (n, Q)    dF/dX_S,  positive domain   (partial_terms.py:400-431)       */
    GPARML_A_Y = 16,          /* (n, D)    the shard's outputs                                           */
    GPARML_A_GRAD_GLOBAL = 17, /* (M*Q + Q + 2) dF/d[Z, sf2, alpha, beta], positive domain              */
    GPARML_A_GS_EXTRA = 18    /* (13) scalars of the master step: log det Kmm, log det A, tr(Kmm^-1 Psi2),
                                 tr(Psi1Y^T A^-1 Psi1Y), four contractions, then (probe) SM clock cycles of the
                                 head kernel's sections: form A, inversion, C / tr, dF/dPsi2, pair tables   */
};

/* the 12 accumulated statistics in the reference's layouts (parallel_GPLVM.py:142-151),
 * filled by gparml_stats_expand(); any pointer may be NULL to skip that statistic. */
typedef struct gparml_named_stats {
    double *sum_YYT;                      /* (1)       partial_terms.py:40            */
    double *sum_exp_K_ii;                 /* (1)       partial_terms.py:81            */
    double *sum_exp_K_mi_K_im;            /* (M, M)    partial_terms.py:79            */
    double *sum_exp_K_miY;                /* (M, D)    kernel_exp.py:13-49            */
    double *sum_KL;                       /* (1)       partial_terms.py:83-87         */
    double *sum_d_exp_K_miY_d_Z;          /* (M, Q, D) partial_terms.py:162-188       */
    double *sum_d_exp_K_mi_K_im_d_Z;      /* (M, Q, M) partial_terms.py:190-205       */
    double *sum_d_exp_K_miY_d_alpha;      /* (Q, M, D) partial_terms.py:256-271       */
    double *sum_d_exp_K_mi_K_im_d_alpha;  /* (Q, M, M) partial_terms.py:273-284       */
    double *sum_d_exp_K_ii_d_sf2;         /* (1)       partial_terms.py:318-320       */
    double *sum_d_exp_K_miY_d_sf2;        /* (M, D)    partial_terms.py:310-312       */
    double *sum_d_exp_K_mi_K_im_d_sf2;    /* (M, M)    partial_terms.py:314-316       */
} gparml_named_stats;

/* ---- library / context -------------------------------------------------- */
int gparml_abi_version(void);
int gparml_device_count(void);                       /* <= 0: no CUDA device                  */
const char *gparml_last_error(void);                 /* message of the last failing call      */

/* One shard context on CUDA device `device`.  Replaces partial_terms.__init__
 * (partial_terms.py:16-36) + load_partial_terms (local_MapReduce.py:403-409).
 * n_total is the global number of points N (options['N'], local_MapReduce.py:46). */
int gparml_create(gparml_ctx **out, int device, int M, int Q, int D, int64_t n_total, int flags);
int gparml_destroy(gparml_ctx *ctx);
/* Order all work of the context on the caller's stream (cudaStream_t passed as void*).  The
 * handle is taken literally: NULL is CUDA's legacy default stream 0 (what torch uses unless a
 * stream context is active), NOT "the context's own stream" -- use gparml_use_own_stream()
 * to go back to the private non-blocking stream the context was created with. */
int gparml_set_stream(gparml_ctx *ctx, void *cuda_stream);
int gparml_use_own_stream(gparml_ctx *ctx);
int gparml_synchronize(gparml_ctx *ctx);
int gparml_set_n_total(gparml_ctx *ctx, int64_t n_total);

/* ---- shard data (device-resident between evaluations) ------------------- */
/* Copies Y (n,D), X_mu (n,Q), X_S (n,Q) host -> device; replaces the per-evaluation
 * genfromtxt + load of local_MapReduce.py:195-201 / 323-329.  Also computes
 * sum_n y_n.y_n (partial_terms.py:40).  May be called again with a different n.
 * Asynchronous: the copies run on the context's copy stream (X_mu / X_S in up to eight row ranges that double in size,
 * then Y) and the next gparml_statistics starts on the first range while the others are still in
 * flight; every other entry point waits for them.  Pageable host arrays may be reused when the call
 * returns, pinned ones must stay valid until the next call that synchronises (gparml_statistics with
 * unconstrained variances, gparml_global_step[_end], gparml_download, gparml_synchronize). */
int gparml_upload_shard(gparml_ctx *ctx, const double *Y, const double *X_mu, const double *X_S,
                        int64_t n_local, int variance_domain);
int64_t gparml_n_local(const gparml_ctx *ctx);
/* Number of evaluations so far in which Kmm or Kmm + beta Psi2 only factorised after the reference's jitter
 * retry (+1e-7 on the diagonal, partial_terms.py:453-457; single-CTA master step, M <= 116).  A matrix that is
 * not positive definite even then gives GPARML_ERR_NOT_PD. */
int64_t gparml_jitter_events(const gparml_ctx *ctx);

/* ---- per-evaluation globals --------------------------------------------- */
/* Z (M,Q), alpha (Q,), sf2, beta: the `global_statistics_*_<i>.npy` broadcast
 * (parallel_GPLVM.py:236-238).  Also builds Kmm-side pair constants. */
int gparml_set_globals(gparml_ctx *ctx, const double *Z, double sf2, const double *alpha, double beta);
/* options['step_size'] (parallel_GPLVM.py:228): X += step * grad_d in memory only
 * (local_MapReduce.py:205-211 / 333-338).  0 disables the step. */
int gparml_set_step(gparml_ctx *ctx, double step_size);

/* ---- map 1: statistics_mapper (local_MapReduce.py:183-248) --------------- */
/* prep_points + psi1_stats + psi2_stats -> packed partial sums of THIS shard on device.
 * gparml_statistics blocks until a device-side input check can be reported: with unconstrained variances
 * it returns GPARML_ERR_RANGE where supporting_functions.py:154 asserts.
 * gparml_statistics_launch only queues the work and returns (no host synchronisation): one host thread can
 * start the map on every shard / GPU before it waits for any (the reference forks one mapper per shard,
 * local_MapReduce.py:134-137).  The check result stays in the device status word and is reported by the next
 * gparml_status (waits for the context's stream) or gparml_global_step_end. */
int gparml_statistics(gparml_ctx *ctx);
int gparml_statistics_launch(gparml_ctx *ctx);
int gparml_status(gparml_ctx *ctx);
int64_t gparml_stats_count(const gparml_ctx *ctx);            /* doubles in the packed buffer */
/* Device pointer of the packed buffer: the caller sum-reduces it in place across
 * shards (NCCL all-reduce, replacing statistics_reducer local_MapReduce.py:250-277). */
int gparml_stats_device_ptr(gparml_ctx *ctx, void **dev_ptr);
/* In-process reduce for several shards on ONE device: stats += packed buffer of another
 * context on the same device (fixed order = call order, so the sum is reproducible).
 * `scale` multiplies the result afterwards (1.0 normally; total/kept for the reference's
 * --drop_out_fraction rescaling, local_MapReduce.py:263-264). */
int gparml_stats_add(gparml_ctx *ctx, const void *other_dev_ptr, double scale);
/* stats = packed buffer of another context on the same device: hands the reduced sums to
 * every shard before the embeddings map (the accumulated_statistics_*.npy files each
 * embeddings_mapper loads, local_MapReduce.py:318-321). */
int gparml_stats_copy(gparml_ctx *ctx, const void *other_dev_ptr);
/* packed buffer -> the reference's 12 named arrays (host). */
int gparml_stats_expand(gparml_ctx *ctx, const gparml_named_stats *out);
/* set_local_statistics (partial_terms.py:54-61) from reference-layout host arrays:
 * only the five statistics the reference passes are needed for the global step's F
 * and partial derivatives; the derivative tensors are optional (NULL keeps the
 * device values). */
int gparml_stats_set_named(gparml_ctx *ctx, const gparml_named_stats *in);

/* ---- master: calculate_global_statistics / _derivatives ------------------ */
/* (parallel_GPLVM.py:302-369; partial_terms.py:54-61,102-160,207-360,436-473).
 * Uses the (reduced) packed buffer.  F and grad are host outputs; grad has
 * M*Q + Q + 2 entries ordered Z, sf2, alpha, beta (positive domain, no softplus
 * chain).  Either may be NULL.  Returns GPARML_ERR_NOT_PD on a failed Cholesky. */
int gparml_global_step(gparml_ctx *ctx, double *F, double *grad);

/* The same master step in two halves, so that the caller can put the embeddings map between them:
 *   gparml_global_step_begin(ctx);  gparml_embedding_grads(ctx);  gparml_global_step_end(ctx, &F, grad);
 * _begin is asynchronous.  What embeddings_mapper needs from the master (partial_derivatives_*.npy:
 * dF/dPsi1Y, dF/dPsi2; parallel_GPLVM.py:322-332) is produced on the context's stream; the bound and the
 * gradients of Z, sf2, alpha, beta (parallel_GPLVM.py:336-369) are finished and downloaded on a side
 * stream, concurrently with gparml_embedding_grads.  _end blocks until they are on the host and reports a
 * non-positive-definite Kmm / Kmm + beta Psi2 (GPARML_ERR_NOT_PD) or a variance out of range
 * (GPARML_ERR_RANGE).  F / grad may be NULL.  gparml_global_step == _begin followed by _end. */
int gparml_global_step_begin(gparml_ctx *ctx);
int gparml_global_step_end(gparml_ctx *ctx, double *F, double *grad);
/* cache() only: Kmm and Kmm^-1 (local_MapReduce.py:383-394). */
int gparml_update_global_statistics(gparml_ctx *ctx);

/* ---- map 2: embeddings_mapper (local_MapReduce.py:310-363) --------------- */
/* grad_X_mu / grad_X_S (partial_terms.py:367-431) with the softplus chain and the
 * sign flip of local_MapReduce.py:357-360; result stays on device (GRAD_LATEST). */
int gparml_embedding_grads(gparml_ctx *ctx);
/* The same map followed by the download of GRAD_LATEST (2, n, Q) into caller-owned (ideally
 * pinned) host memory -- the `.grad_latest.npy` the reference writes (local_MapReduce.py:359-360).
 * The points are processed in `chunks` ranges (1..8) and the device-to-host copy of one range
 * overlaps the kernels of the next; the ranges shrink geometrically by the copy / kernel time ratio
 * the previous call measured (0.7 on the first call).  Returns when the host array is complete. */
int gparml_embedding_grads_download(gparml_ctx *ctx, double *host_grad_latest, int chunks);

/* ---- partial_terms helper surface ------------------------------------------ */
/* Kmm-side derivative tensors to caller-owned host memory:
 *   which 0: dKmm_dZ     (M,Q,M)  partial_terms.py:146-160
 *   which 1: dKmm_dalpha (Q,M,M)  partial_terms.py:247-254
 *   which 2: dKmm_dsf2   (M,M)    partial_terms.py:306-308
 * Needs Kmm (gparml_update_global_statistics or gparml_global_step). */
int gparml_kmm_derivative(gparml_ctx *ctx, int which, double *out);
/* Chain-rule contraction of six caller-supplied host tensors (reference layouts):
 *   which 0: grad_Z     -> out (M,Q)   partial_terms.py:207-240
 *   which 1: grad_alpha -> out (Q)     partial_terms.py:286-299
 *   which 2: matrix part of grad_sf2 -> out (1)  partial_terms.py:322-333
 *            (the caller adds dF_dexp_K_ii * dexp_K_ii_dsf2, a product of two scalars) */
int gparml_grad_contract(gparml_ctx *ctx, int which, const double *dF_dKmm, const double *dKmm_dx,
                         const double *dF_dPsi1Y, const double *dPsi1Y_dx, const double *dF_dPsi2,
                         const double *dPsi2_dx, double *out);
/* stats = (stats + packed buffer of a context on ANOTHER device) * scale (peer copy + add); stream-ordered
 * behind the other context's queued work, no host synchronisation. */
int gparml_stats_add_peer(gparml_ctx *ctx, gparml_ctx *other, double scale);
/* stats = packed buffer of another context on any device of this process (stream-ordered, asynchronous). */
int gparml_stats_copy_peer(gparml_ctx *ctx, gparml_ctx *other);
/* statistics_reducer (local_MapReduce.py:250-277) for n shard contexts driven by ONE host thread, on any mix of
 * GPUs: every context's packed buffer becomes scale * (sum over the n buffers, added in list order).  One kernel
 * on ctxs[0]'s device reads and writes the other GPUs' buffers directly through NVLink peer memory; streams are
 * ordered with events and the host never waits.  scale: 1.0, or total/kept shards for --drop_out_fraction
 * (local_MapReduce.py:263-264).  n <= GPARML_MAX_PEERS. */
#define GPARML_MAX_PEERS 16
int gparml_stats_allreduce_peers(gparml_ctx **ctxs, int n, double scale);

/* ---- generic transfers ---------------------------------------------------- */
int64_t gparml_array_count(const gparml_ctx *ctx, int array_id);
int gparml_download(gparml_ctx *ctx, int array_id, double *dst, int64_t count);
int gparml_upload(gparml_ctx *ctx, int array_id, const double *src, int64_t count);
int gparml_array_device_ptr(gparml_ctx *ctx, int array_id, void **dev_ptr);

/* ---- optimiser local state: scg_adapted_local_MapReduce.py:29-243 --------- */
int gparml_scg_set_grads(gparml_ctx *ctx);                          /* :29-55   */
int gparml_scg_get_mu(gparml_ctx *ctx, double *out);                /* :60-75   */
int gparml_scg_get_kappa(gparml_ctx *ctx, double *out);             /* :77-90   */
int gparml_scg_get_theta(gparml_ctx *ctx, double *out);             /* :92-109  */
int gparml_scg_get_current_grad(gparml_ctx *ctx, double *out);      /* :111-124 */
int gparml_scg_get_gamma(gparml_ctx *ctx, double *out);             /* :126-141 */
int gparml_scg_get_max_d(gparml_ctx *ctx, double alpha, double *out); /* :143-156 */
int gparml_scg_reset_d(gparml_ctx *ctx);                            /* :161-174 */
int gparml_scg_update_d(gparml_ctx *ctx, double gamma);             /* :176-191 */
int gparml_scg_update_X(gparml_ctx *ctx, double alpha);             /* :193-216 */
int gparml_scg_update_grad_old(gparml_ctx *ctx);                    /* :218-230 */
int gparml_scg_update_grad_new(gparml_ctx *ctx);                    /* :232-243 */

/* ---- one-off initialisation on the device ------------------
 * Replaces the master-side PCA over ALL outputs (local_MapReduce.py:52-65 ->
 * supporting_functions.py:102-121), the variance draw (local_MapReduce.py:88-93) and the
 * k-means for the inducing inputs (parallel_GPLVM.py:170-186 -> scipy.cluster.vq.kmeans).
 * Every function works on ONE shard and returns partial sums the caller adds over shards. */
/* upload only the outputs of a shard (n, D); X_mu / X_S are zero-filled (unconstrained domain)
 * until gparml_init_project / gparml_init_random or gparml_upload write them. */
int gparml_upload_outputs(gparml_ctx *ctx, const double *Y, int64_t n);
/* out (D): sum_n y_n */
int gparml_init_column_sums(gparml_ctx *ctx, double *out);
/* out (D, D): sum_n (y_n - mean)(y_n - mean)^T, mean (D) = global column means */
int gparml_init_scatter(gparml_ctx *ctx, const double *mean, double *out);
/* X_mu = (Y - mean) W with W (D, Q): for PCA W[:, q] = v_q sqrt(N / lambda_q)
 * (supporting_functions.py:117-120: U[:, :Q] / std) */
int gparml_init_project(gparml_ctx *ctx, const double *mean, const double *W);
/* what = 0: X_S = softplus^-1(clip(0.5 + 0.01 N(0,1), 0.001, 1)) (local_MapReduce.py:90-93);
 * what = 1: X_mu = N(0,1) (init == 'random', local_MapReduce.py:86-87).  Counter-based:
 * element (row_offset + i, q) of stream `seed` does not depend on the sharding. */
int gparml_init_random(gparml_ctx *ctx, int what, uint64_t seed, int64_t row_offset);
/* one assignment pass of k-means over the shard's X_mu against centroids (k, Q):
 * out (k*(1+Q) + 1) = per cluster [count, sum of members (Q)], then the summed Euclidean
 * distance to the nearest centroid (scipy.cluster.vq.vq + update_cluster_means). */
int gparml_kmeans_step(gparml_ctx *ctx, const double *centroids, int k, double *out);

/* ---- introspection for bench.py / tests ----------------------------------- */
/* kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t gparml_launch_count(const gparml_ctx *ctx);
/* milliseconds of the last gparml_statistics / global_step / embedding_grads
 * phases measured with CUDA events on the context's stream:
 * out[0]=prep_points out[1]=psi1_stats out[2]=psi2_stats out[3]=global_step
 * out[4]=embed_grads.  Only valid with gparml_enable_timing(ctx, 1). */
int gparml_enable_timing(gparml_ctx *ctx, int on);
int gparml_phase_times(gparml_ctx *ctx, double *out5);
/* sustained FP64-pipe peak measured with a pure-DFMA kernel, in DFMA lane-ops/s. */
int gparml_measure_dfma_peak(gparml_ctx *ctx, double *lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GPARML_B200_H */
