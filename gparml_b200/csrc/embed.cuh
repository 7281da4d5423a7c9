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
    const GlobalsDev *glob;
    int64_t n;           // points in the shard (stride of the partial buffers)
    int64_t i0, i1;      // this launch covers points [i0, i1)
    int M, D;
    int m_bounds[EMB_MAX_SPLITS + 1];   // row splits (sqrt(w)-basis kernels)
    int p_bounds[EMB_MAX_SPLITS + 1];   // pair splits (expanded-basis kernel)
    double *partial;     // [splits][n][2Q + 1]  (AM, AS, AH) resp. (BZ, BZZ, AH)
    double *psi1_part;   // [n][2Q + 1]          (sum_m h1 ad, sum_m h1 ad^2, sum_m h1),  h1 = B Psi1
};

// embed_x.cu: hand-scheduled expanded-basis Psi2 part (compiled with ptxas -O1 so that the
// instruction order written in the source is the order that is issued)
int gp_embed_psi2x_points_per_cta(int Q);
int gp_embed_psi2x_occupancy(int Q, int *occ);
int gp_launch_embed_psi2x(gparml_ctx *c, const EmbedParams &p, int ntiles, int splits);
