// Shared declarations of the embed_grads translation units (embed.cu, embed_x.cu).
#pragma once
#include "common.cuh"

#ifndef EMB_THREADS
#define EMB_THREADS 128
#endif
#define EMB_MAX_SPLITS 32

struct EmbedParams {
    const double *rec1, *rec2, *Y, *Z, *G1;
    const double2 *pair_g;
    const double2 *pair_h;     // (P) (lk + log|Gs|, sign Gs)
    const double2 *pair_zz;    // (P, Q) (zc, zc^2), zc = zbar - center
    const double *pair_zc;     // (P, Q rounded up to even) zc alone
    const double *pair_r;      // blocked feature table of embed_psi2m (common.cuh)
    const double *pair_ra;     // pair_r times sign(Gs) exp(lk + log|Gs|)
    const GlobalsDev *glob;
    int64_t n;           // points in the shard (stride of the partial buffers)
    int64_t i0, i1;      // this launch covers points [i0, i1)
    int M, D;
    int m_bounds[EMB_MAX_SPLITS + 1];   // row splits (sqrt(w)-basis kernels)
    int p_bounds[EMB_MAX_SPLITS + 1];   // pair splits (expanded-basis kernel)
    double *partial;     // [splits][pstride][2Q + 1]  (AM, AS, AH) resp. (BZ, BZZ, AH); row of point i: i - pbase
    int64_t pstride, pbase;
    // embed_psi2m, one launch: CTAs [0, full_tiles) take a whole point tile with the full pair range and finish the
    // gradients themselves; the others are (tile, pair split) = (full_tiles + r / tail_splits, r % tail_splits)
    int full_tiles, tail_splits;
    double *psi1_part;   // [n][2Q + 1]          (sum_m h1 ad_q, sum_m h1 (ad_q^2 - a_q), -),  h1 = B Psi1
    // fused finish (expanded-basis kernel with ONE pair split): the epilogue of embed_psi2x writes the gradients itself
    int fuse_finish;
    const double *s_pos, *s_sig;
    double *gx_mu, *gx_s, *grad_latest;
};

// The last step of the embeddings map for one (point, q): combine the Psi2 sums (expanded basis: am = sum_p h zc_q,
// as = sum_p h zc_q^2, ah = sum_p h) with the Psi1 part and the KL terms (partial_terms.py:385,418), apply the softplus
// chain and the sign flip (local_MapReduce.py:357-360).  Shared by embed_finish_kernel and the fused epilogue.
__device__ __forceinline__ void gp_embed_finish_one(double mu, double w, double mc, double am, double as, double ah, double p1_q,
                                                    double p1_Qq, double S, double sig, double *gmu, double *gs, double *gl_mu,
                                                    double *gl_s)
{
    const double t1 = w * fma(mc, ah, -am);                                 // sum_p h wd_q
    const double t2 = w * (w * fma(mc, fma(mc, ah, -2.0 * am), as));        // sum_p h wd_q^2
    const double g_m = -mu - p1_q - 2.0 * t1;
    const double g_s = -0.5 * (1.0 - 1.0 / S) + 0.5 * p1_Qq + (2.0 * t2 - w * ah);
    *gmu = g_m;
    *gs = g_s;
    *gl_mu = -g_m;
    *gl_s = -(g_s * sig);
}

// embed_x.cu: hand-scheduled expanded-basis Psi2 part (compiled with ptxas -O1 so that the
// instruction order written in the source is the order that is issued)
int gp_embed_psi2x_points_per_cta(int Q);
int gp_embed_psi2x_occupancy(int Q, int *occ);
int gp_launch_embed_psi2x(gparml_ctx *c, const EmbedParams &p, int ntiles, int splits);
// embed_m.cu: the same sums on the FP64 tensor-core instruction; p_bounds are in chunks of GP_PAIR_CHUNK pairs
int gp_embed_psi2m_points_per_cta();
int gp_embed_psi2m_occupancy(int Q, int *occ);
int gp_launch_embed_psi2m(gparml_ctx *c, const EmbedParams &p, int ctas);
int gp_launch_pair_ra(gparml_ctx *c);      // pair_ra from pair_r and pair_h, once per master step
